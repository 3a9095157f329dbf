"""Opt-in kernel variants (each read once from the environment by the library, so each set runs in its own
process): the CTA-pair (cta_group::2) GEMM / vocabulary-NLL kernels, the 4-CTA-cluster sampled GEMM, the
cluster-multicast / staggered persistent LSTM kernel, and the direct-store (no TMA store) epilogue.  They are
kept correct and measurable even where the production default is another path."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(env_extra, selection):
    env = dict(os.environ, **env_extra)
    cmd = [sys.executable, "-m", "pytest", "-q", "-x", "-m", "gpu", "-p", "no:cacheprovider"] + selection
    res = subprocess.run(cmd, cwd=ROOT, env=env, capture_output=True, text=True, timeout=900)
    assert res.returncode == 0, res.stdout[-3000:] + res.stderr[-1000:]


def test_cta_pair_kernels():
    _run({"BLM_GEMM2": "1"}, ["tests/test_gpu_kernels.py", "-k", "pair_kernel or vocab_nll or fast_gelu",
                              "tests/test_gpu_full_size.py"])


def test_direct_store_epilogue():
    _run({"BLM_TMA_STORE": "0"}, ["tests/test_gpu_kernels.py", "-k", "pair_kernel or fast_gelu or gemm_epilogues"])


def test_cluster_sampled_gemm():
    _run({"BLM_SAMPLED_CLUSTER": "1"}, ["tests/test_gpu_kernels.py", "-k", "cluster_sampled"])


def test_lstm_cluster_multicast_and_stagger():
    _run({"BLM_LSTM_CLUSTER": "1", "BLM_LSTM_STAGGER": "1"}, ["tests/test_gpu_lstm.py"])


def test_training_with_materialised_transposes():
    """BLM_TRAIN_TRANSPOSE=1: the fine-tune step with transposed operand copies instead of MN-major descriptors."""
    _run({"BLM_TRAIN_TRANSPOSE": "1"}, ["tests/test_gpu_train.py", "-k", "bayes_tm or lstm_finetune_step"])
