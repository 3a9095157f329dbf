"""bayeslms_b200 -- B200 (sm_100a) kernels behind the BayesLMs n-best rescoring hot path.

Host code is Python/PyTorch (memory, streams, torch.distributed); all arithmetic runs in
hand-written CUDA behind the C ABI of ``include/bayeslm_b200.h``.
"""
__version__ = "0.1.0"
