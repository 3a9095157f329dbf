"""CPU-side checks: C-ABI surface, state_dict compatibility, host logic, sharding (gloo)."""
import ctypes
import os
import re
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_symbols():
    text = open(os.path.join(ROOT, "include", "bayeslm_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(blm_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    from bayeslms_b200 import _lib
    handle = ctypes.CDLL(_lib.LIB_PATH)
    declared = _header_symbols()
    assert len(declared) >= 18
    for name in declared:
        assert hasattr(handle, name), f"{name} declared in include/bayeslm_b200.h but not exported"
    assert sorted(_lib.SIGNATURES) == declared, "ctypes binding and header disagree"
    assert _lib.lib().blm_version() >= 100


def test_no_cpu_fallback():
    """Product ops refuse CPU tensors instead of silently computing on the host."""
    from bayeslms_b200 import _lib, model as M
    net = M.BayesTransformerModel(50, 32, 4, 64, 2, 0.5, True, "FFN")
    with pytest.raises(_lib.BlmError):
        net(torch.zeros(4, 1, dtype=torch.long))


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "bayeslms_b200")
    for fn in os.listdir(pkg):
        if fn.endswith(".py"):
            src = open(os.path.join(pkg, fn)).read()
            assert "import oracle" not in src and "from oracle" not in src, fn


@pytest.mark.parametrize("name", ["bayes_tm_FFN", "bayes_tm_MHA", "bayes_tm_EMB", "bayes_tm_none", "gauss_tm_0",
                                  "gauss_tm_3", "v_tm_0", "v_tm_1", "v_tm_2", "v_tm_3", "bayes_lstm_0",
                                  "bayes_lstm_3", "std_tm", "std_tm_relu", "std_lstm"])
def test_state_dict_layout_matches_reference(golden, name):
    from tests.util import build_from_cfg
    rec = golden(name + ".pt")
    net = build_from_cfg(rec["cfg"])
    mine = {k: tuple(v.shape) for k, v in net.state_dict().items()}
    ref = {k: tuple(v.shape) for k, v in rec["state_dict"].items()}
    if "pos_encoder.pe" in mine:
        assert mine.pop("pos_encoder.pe")[1:] == (1, rec["cfg"]["ninp"])
    assert mine == ref
    net.load_state_dict(rec["state_dict"], strict=False)


def test_load_checkpoint_fails_loudly_on_the_wrong_architecture(golden, tmp_path):
    """A reference TransformerModel checkpoint given to BayesTransformerModel(..., 'none') (different key names) used to
    load as random init; now it raises, and so does a Bayesian checkpoint given to the wrong --T_bayes_pos."""
    from bayeslms_b200 import model as M, scorer as S
    rec = golden("std_tm.pt")
    c = rec["cfg"]
    path = str(tmp_path / "std_tm.pt")
    torch.save(rec["state_dict"], path)
    good = M.TransformerModel(c["ntoken"], c["ninp"], c["nhead"], c["nhid"], c["nlayers"], 0.5, "gelu", True)
    assert S.load_checkpoint(good, path) == {"missing": [], "unexpected": []}
    assert torch.equal(good.transformerlayers[1].self_attn.in_proj_weight,
                       rec["state_dict"]["transformerlayers.layers.1.self_attn.in_proj_weight"])
    wrong = M.BayesTransformerModel(c["ntoken"], c["ninp"], c["nhead"], c["nhid"], c["nlayers"], 0.5, True, "none")
    with pytest.raises(KeyError, match="does not match BayesTransformerModel"):
        S.load_checkpoint(wrong, path)
    left = S.load_checkpoint(wrong, path, partial=True)      # --prior semantics: intersect, report the rest
    assert "transformerlayers.0.linear1.weight" in left["missing"] and left["unexpected"]
    rec = golden("bayes_tm_FFN.pt")
    path = str(tmp_path / "ffn.pt")
    torch.save(rec["state_dict"], path)
    c = rec["cfg"]
    with pytest.raises(KeyError):
        S.load_checkpoint(M.BayesTransformerModel(c["ntoken"], c["ninp"], c["nhead"], c["nhid"], c["nlayers"], 0.5, True, "MHA"), path)
    bad = dict(rec["state_dict"])
    bad["decoder.bias"] = bad["decoder.bias"][:-1]
    torch.save(bad, path)
    with pytest.raises(ValueError, match="decoder.bias"):
        S.load_checkpoint(M.BayesTransformerModel(c["ntoken"], c["ninp"], c["nhead"], c["nhid"], c["nlayers"], 0.5, True, "FFN"), path)


def _recipe_train_args(script, **env):
    """The stage-1 ``python steps/pytorchnn/train.py ...`` command line of a run_nnlm_*.sh recipe, with the shell
    variables it declares at the top (plus ``env``) substituted -- what the reference pipeline really passes."""
    text = open(script).read()
    cmd = re.search(r"python steps/pytorchnn/train\.py(.*?)> \$pytorch_path/train\.log", text, flags=re.S).group(1)
    variables = dict(re.findall(r"^([A-Za-z_0-9]+)=([^\s#]*)", text, flags=re.M))
    variables.update(env)
    words = cmd.replace("\\\n", " ").split()
    out = []
    for w in words:
        m = re.fullmatch(r"\$\{?([A-Za-z_0-9]+)\}?", w)
        out.append(variables.get(m.group(1), "x").strip('"') if m else w)
    return out


@pytest.mark.skipif(not os.path.exists("/root/reference/run_nnlm_ami_tm.sh"), reason="needs the reference checkout")
@pytest.mark.parametrize("script", ["run_nnlm_ami_tm.sh", "run_nnlm_ami_lstm.sh", "run_nnlm_lrs2_tm.sh", "run_nnlm_lrs2_lstm.sh"])
def test_train_cli_parses_the_recipe_command_lines(script):
    """run_nnlm_ami_tm.sh:89-110 / run_nnlm_ami_lstm.sh stage 1 pass --prior, --prior_path, --mark, --epoch (a prefix of
    --epochs), --cuda: the trainer must accept exactly that line."""
    from bayeslms_b200 import train as T
    argv = _recipe_train_args(os.path.join("/root/reference", script), data_dir="d", nn_model="m.pt", prior_path="p")
    assert "--prior" in argv and "--prior_path" in argv and "--tied" in argv
    args = T.build_parser().parse_args(argv)
    assert args.epochs == 32 and args.batch_size == 32 and args.clip == 1.0 and args.tied and args.cuda
    assert args.prior in ("True", "False") and args.prior_path == "p"


def test_train_cli_recipe_line_verbatim():
    """The same check without the reference checkout: the argv of run_nnlm_ami_tm.sh:89-110 with its default variables."""
    from bayeslms_b200 import train as T
    argv = ("--data data/pytorchnn_ami+fisher --model Transformer --emsize 512 --nhid 4096 --nlayers 6 --nhead 8 --lr 0.1 "
            "--dropout 0.2 --seq_len 100 --clip 1.0 --batch-size 32 --epoch 32 --seed 1111 --save exp/m/model.pt "
            "--prior False --prior_path steps/pytorchnn/prior/transformer --uncertainty Bayesian --T_bayes_pos FFN "
            "--T_gauss_pos 3 --T_v_pos 0 --tied --cuda").split()
    a = T.build_parser().parse_args(argv)
    assert (a.epochs, a.prior, a.prior_path, a.dropout, a.T_bayes_pos) == (32, "False", "steps/pytorchnn/prior/transformer", 0.2, "FFN")
    argv = ("--data d --model LSTM --emsize 1024 --nhid 1024 --nlayers 2 --nhead 8 --lr 5 --dropout 0.2 --seq_len 100 "
            "--clip 1.0 --batch-size 32 --epoch 32 --seed 1111 --save m.pt --uncertainty Bayesian --L_bayes_pos 3 "
            "--L_gauss_pos 00 --L_v_pos 00 --prior True --prior_path steps/pytorchnn/prior/lstm --tied --mark no --cuda").split()
    a = T.build_parser().parse_args(argv)
    assert (a.mark, a.prior, a.L_bayes_pos, a.optimizer, a.debug, a.work_dir) == ("no", "True", 3, "SGD", False, "TFM")
    assert T.pruned_length(1000, "base-0.25set") == 250 and T.pruned_length(1000, "no") == 1000


def test_build_model_follows_the_reference_switch():
    """score.py:374-448 / train.py:193-224: --uncertainty none builds TransformerModel / RNNModel (torch key names)."""
    import argparse
    from bayeslms_b200 import model as M
    ns = argparse.Namespace(model="Transformer", uncertainty="none", emsize=32, nhead=4, nhid=64, nlayers=2)
    m = M.build_model(ns, 40)
    assert isinstance(m, M.TransformerModel) and "transformerlayers.layers.1.self_attn.in_proj_weight" in m.state_dict()
    assert m.decoder.weight is m.encoder.weight          # the scorer ties (score.py:377)
    ns = argparse.Namespace(model="LSTM", uncertainty="none", emsize=32, nhid=32, nlayers=2, tied=False)
    m = M.build_model(ns, 40)
    assert isinstance(m, M.RNNModel) and "rnn.weight_hh_l1" in m.state_dict() and m.decoder.weight is not m.encoder.weight
    assert M.TransformerModel(40, 32, 4, 64, 2, 0.5, "relu", True).transformerlayers[0].activation == "relu"
    assert M.TransformerModel(40, 32, 4, 64, 2).activation == "relu"           # the constructor's default (model.py:124)
    with pytest.raises(ValueError):
        M.TransformerModel(40, 32, 4, 64, 2, 0.5, "tanh", True)
    with pytest.raises(NotImplementedError):
        M.RNNModel("GRU", 40, 32, 32, 2)


def test_import_model_shim_exposes_the_reference_classes():
    """with shim/steps/pytorchnn first on the path, ``import model`` gives the reference's
    class names with their positional constructors (score.py:377-440) -- checked in a fresh interpreter."""
    import subprocess
    code = (
        "import model, torch\n"
        "m = model.BayesTransformerModel(40, 32, 4, 64, 2, 0.5, True, 'FFN')\n"
        "assert 'transformerlayers.0.linear2.weight_lgstd' in m.state_dict()\n"
        "model.TransformerModel(40, 32, 4, 64, 2, 0.5, 'gelu', True); model.RNNModel('LSTM', 40, 32, 32, 2, 0.5, True)\n"
        "model.BayesRNNModel('LSTM', 40, 32, 32, 2, 0.5, True, 3); model.GaussRNNModel('LSTM', 40, 32, 32, 2, 0.5, False, '31')\n"
        "model.GaussTransformerModel(40, 32, 4, 64, 2, 0.5, True, 3); model.VTransformerModel(40, 32, 4, 64, 6, 0.5, True, 3)\n"
        "model.VariationalRNNModel('LSTM', 40, 32, 32, 2, 0.5, True, '11')\n"
        "assert model.__file__.endswith('shim/steps/pytorchnn/model.py'); print('ok')\n")
    env = dict(os.environ, PYTHONPATH=os.path.join(ROOT, "shim", "steps", "pytorchnn"))
    r = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True, cwd="/")
    assert r.returncode == 0 and r.stdout.strip() == "ok", r.stderr
    for script in ("compute_sentence_scores_bayes_jianwei.py", "train.py"):
        r = subprocess.run([sys.executable, os.path.join(ROOT, "shim", "steps", "pytorchnn", script), "--help"],
                           capture_output=True, text=True, cwd="/")
        assert r.returncode == 0 and "--uncertainty" in r.stdout, r.stderr


def test_v_pos_normalisation():
    from bayeslms_b200.model import normalise_v_pos, VTransformerModel
    assert [normalise_v_pos(v) for v in (0, 1, 2, 3, "00", "01", "10", "11", 11, "3")] == [0, 1, 2, 3, 0, 1, 2, 3, 3, 3]
    with pytest.raises(ValueError):
        normalise_v_pos(7)
    assert len(VTransformerModel(20, 16, 2, 32, 6, 0.5, True, "11").transformerlayers) == 5


def test_host_parsers_match_oracle(golden, tmp_path):
    from bayeslms_b200 import scorer as S
    from oracle import bayeslm_oracle as O
    rec = golden("scorer_loop.pt")
    vp, npth = tmp_path / "words.txt", tmp_path / "words_text"
    vp.write_text("".join(f"{w} {i}\n" for i, w in enumerate(rec["vocab_words"])) + "w000 999\n")
    npth.write_text("\n".join(rec["nbest_lines"]) + "\n")
    assert S.read_vocab(str(vp)) == O.read_vocab(str(vp))
    a, b = S.load_nbest(str(npth)), O.load_nbest(str(npth))
    assert list(a.items()) == list(b.items())
    vocab = S.read_vocab(str(vp))
    for hyps in a.values():
        for h in hyps:
            assert S.ids_for(h, vocab) == O.get_input_and_target(h, vocab)
    scores = {k: [(h, 1.23456 + i) for i, h in enumerate(v)] for k, v in a.items()}
    S.write_scores(scores, str(tmp_path / "a.nn"))
    O.write_scores(scores, str(tmp_path / "b.nn"))
    assert (tmp_path / "a.nn").read_text() == (tmp_path / "b.nn").read_text()
    with pytest.raises(ValueError):
        (tmp_path / "bad.txt").write_text("only-one-column\n")
        S.read_vocab(str(tmp_path / "bad.txt"))


def test_c_text_helpers_match_the_reference_parsers(tmp_path):
    """The library's one-pass tokeniser / formatter (blm_vocab_from_text, blm_nbest_scan / _tokenize / _group,
    blm_scores_format) against the oracle's restatement of score.py:20-120, 283-303 on awkward text: empty
    hypotheses, OOV words, tabs and runs of spaces, keys with several dashes, CRLF, blank lines, no final newline,
    duplicate vocabulary entries, non-ASCII words."""
    from bayeslms_b200 import scorer as S
    from oracle import bayeslm_oracle as O
    rs = np.random.RandomState(0)
    words = ["<s>", "<unk>"] + [f"w{i}" for i in range(200)] + ["naïve", "日本語", "w5"]      # w5 twice: first index wins
    vp = tmp_path / "words.txt"
    vp.write_text("".join(f"{w} {i}\n" for i, w in enumerate(words)), encoding="utf-8")
    lines = []
    for u in range(40):
        for n in range(rs.randint(1, 9)):
            L = rs.randint(0, 12)
            ws = [words[rs.randint(2, len(words))] if rs.rand() > 0.1 else "oov%d" % rs.randint(9) for _ in range(L)]
            sep = ["  ", " ", "\t", " \t "][rs.randint(4)]
            lines.append(f"sp-k_{u}-seg-{u % 3}-{n + 1} " + sep.join(ws) if ws else f"sp-k_{u}-seg-{u % 3}-{n + 1}")
    lines[3] += "   "
    lines[7] = "  " + lines[7]
    lines.insert(20, "")
    lines.insert(31, "lonekey")
    text = "\r\n".join(lines[:50]) + "\r\n" + "\n".join(lines[50:])        # no trailing newline
    npth = tmp_path / "words_text"
    npth.write_bytes(text.encode("utf-8"))
    vocab_py, vocab_c = O.read_vocab(str(vp)), S.Vocab(str(vp))
    assert len(vocab_c) == len(vocab_py) and all(vocab_c.id(w) == i for w, i in vocab_py.items())
    assert vocab_c.id("missing") == -1
    nbest = O.load_nbest(str(npth))
    nb = S.NbestText(str(npth), threads=3)
    assert not nb.needs_slow_path
    want = [O.get_input_and_target(h, vocab_py) for hyps in nbest.values() for h in hyps]
    assert nb.n_lines == len(want) and nb.n_utts == len(nbest) and nb.n_tokens == sum(len(x) for x, _ in want)
    assert nb.contiguous == (list(nbest.keys()) == [k for i, k in enumerate(
        [ln.strip().partition(" ")[0].rsplit("-", 1)[0] for ln in text.replace("\r\n", "\n").split("\n")])
        if i == 0 or k != [ln.strip().partition(" ")[0].rsplit("-", 1)[0] for ln in text.replace("\r\n", "\n").split("\n")][i - 1]])
    order = np.argsort(nb.utt_of_line, kind="stable")
    for (l0, l1) in ((0, nb.n_lines), (5, 77)):
        m = int(nb.offs[l1] - nb.offs[l0])
        tok, tgt, pos = (np.full(m, -7, dtype=np.int32) for _ in range(3))
        nb.tokenize(vocab_c, l0, l1, tok, tgt, pos)
        for i in range(l0, l1):
            a, b = int(nb.offs[i] - nb.offs[l0]), int(nb.offs[i + 1] - nb.offs[l0])
            x, y = want[int(np.nonzero(order == i)[0][0])] if not nb.contiguous else want[i]
            assert tok[a:b].tolist() == x and tgt[a:b].tolist() == y and pos[a:b].tolist() == list(range(b - a)), i
    scores_in_ref_order = rs.randn(len(want)).astype(np.float32) * 37
    ref = {}
    it = iter(scores_in_ref_order.tolist())
    for k, hyps in nbest.items():
        ref[k] = [(h, next(it)) for h in hyps]
    O.write_scores(ref, str(tmp_path / "ref.nn"))
    by_line = np.empty(len(want), dtype=np.float32)
    by_line[order] = scores_in_ref_order
    assert nb.format_scores(by_line) == (tmp_path / "ref.nn").read_bytes()
    # vocabulary errors and the flags of the slow path
    (tmp_path / "bad.txt").write_text("a 0\nb\n")
    with pytest.raises(ValueError, match="line 2"):
        S.Vocab(str(tmp_path / "bad.txt"))
    (tmp_path / "nbsp").write_bytes("u-1 a\u00a0b\n".encode("utf-8"))
    assert S.NbestText(str(tmp_path / "nbsp")).needs_slow_path
    (tmp_path / "cr").write_bytes(b"u-1 a\ru-2 b\n")
    assert S.NbestText(str(tmp_path / "cr")).needs_slow_path
    (tmp_path / "novocab.txt").write_text("<s> 0\na 1\n")
    (tmp_path / "oov").write_bytes(b"u-1 a zzz\n")
    nb2 = S.NbestText(str(tmp_path / "oov"))
    t = np.zeros(4, dtype=np.int32)
    from bayeslms_b200._lib import BlmError
    with pytest.raises(BlmError, match="<unk>"):
        nb2.tokenize(S.Vocab(str(tmp_path / "novocab.txt")), 0, 1, t, t.copy(), None)


def test_shard_ranges_cover_and_balance():
    from bayeslms_b200.scorer import shard_ranges
    w = list(np.random.RandomState(0).randint(1, 50, size=103))
    for world in (1, 2, 3, 8):
        r = shard_ranges(w, world)
        assert r[0][0] == 0 and r[-1][1] == len(w)
        assert all(r[i][1] == r[i + 1][0] for i in range(world - 1))
        loads = [sum(w[a:b]) for a, b in r]
        assert max(loads) - min(loads) <= 2 * max(w)
    assert shard_ranges([], 2) == [(0, 0), (0, 0)]
    assert shard_ranges([5], 4)[-1][1] == 1


def test_packed_batch_layout():
    from bayeslms_b200.engine import PackedBatch
    b = PackedBatch.from_lists([[0, 5, 6], [0], [0, 9]], [[5, 6, 0], [0], [9, 0]], "cpu")
    assert b.tokens.tolist() == [0, 5, 6, 0, 0, 9] and b.targets.tolist() == [5, 6, 0, 0, 9, 0]
    assert b.pos.tolist() == [0, 1, 2, 0, 0, 1] and b.offsets.tolist() == [0, 3, 4, 6]
    assert (b.max_len, b.n_tokens, b.n_hyp) == (3, 6, 3)


def _gloo_worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    import torch.distributed as dist
    from collections import OrderedDict
    from bayeslms_b200 import scorer as S
    dist.init_process_group("gloo", rank=rank, world_size=world)

    class FakeRescorer:  # host-side sharding logic only: the "score" is a function of the token ids
        is_rnn, device = False, torch.device("cpu")

        def score_transformer(self, hyps):
            return np.asarray([float(sum(x) * 0.5 + len(y)) for x, y in hyps], dtype=np.float32)

    vocab = {"<s>": 0, "<unk>": 1, **{f"w{i}": i + 2 for i in range(30)}}
    rs = np.random.RandomState(3)
    nbest = OrderedDict((f"utt{u}", [" ".join(f"w{rs.randint(30)}" for _ in range(rs.randint(0, 7))) or " "
                                     for _ in range(rs.randint(1, 5))]) for u in range(17))
    res = S.score_nbest(None, nbest, vocab, rank=rank, world=world, rescorer=FakeRescorer())
    single = S.score_nbest(None, nbest, vocab, rank=0, world=1, rescorer=FakeRescorer())
    q.put((rank, res == single))
    dist.destroy_process_group()


def test_sharded_scoring_equals_single_rank_gloo():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_gloo_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(60)
    assert got == [(0, True), (1, True)]


def test_balanced_chunks_cover_and_respect_the_limit():
    """scorer._chunks_by_tokens: contiguous, complete, every chunk <= max_tokens, fewest chunks, balanced sizes."""
    import numpy as np
    from bayeslms_b200.scorer import _chunks_by_tokens
    rng = np.random.default_rng(3)
    for n, mx in ((12800, 65536), (1000, 500), (7, 30), (1, 5), (50, 27)):
        lens = rng.integers(1, 27, size=n).tolist()
        cuts = _chunks_by_tokens(lens, mx)
        assert cuts[0][0] == 0 and cuts[-1][1] == n and all(a[1] == b[0] for a, b in zip(cuts, cuts[1:]))
        sizes = [sum(lens[a:b]) for a, b in cuts]
        assert max(sizes) <= max(mx, max(lens))
        greedy, tot = 1, 0                      # never more chunks than a greedy fill
        for x in lens:
            if tot and tot + x > mx:
                greedy, tot = greedy + 1, 0
            tot += x
        assert len(cuts) <= greedy
        if len(sizes) > 1 and mx >= 500:        # balanced when items are small against the limit
            assert max(sizes) - min(sizes) <= 2 * 26 + sum(lens) // len(sizes) // 10
    assert _chunks_by_tokens([], 10) == []
    assert [sum([10, 1, 1, 1, 10][a:b]) <= 11 for a, b in _chunks_by_tokens([10, 1, 1, 1, 10], 11)] == [True] * 3


def test_flatten_sessions_layout():
    """engine.flatten_sessions: nested [session][utterance][(input, target)] -> flat arrays in row order."""
    import numpy as np
    from bayeslms_b200.engine import flatten_sessions
    sessions = [[[([0, 5, 6], [5, 6, 0]), ([0, 7], [7, 0])], [([0], [0])]], [[([0, 9, 9, 9], [9, 9, 9, 0])]]]
    tok, tgt, offs, sess_of, utt_of = flatten_sessions(sessions)
    assert tok.tolist() == [0, 5, 6, 0, 7, 0, 0, 9, 9, 9] and tgt.tolist() == [5, 6, 0, 7, 0, 0, 9, 9, 9, 0]
    assert offs.tolist() == [0, 3, 5, 6, 10]
    assert sess_of.tolist() == [0, 0, 0, 1] and utt_of.tolist() == [0, 0, 1, 0]
    assert tok.dtype == np.int32 and tgt.dtype == np.int32


def test_bench_reference_arm_contract():
    """bench.py --impl reference prints exactly one JSON line on stdout with the keys of the contract."""
    import json
    import subprocess
    import sys
    res = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1",
                          "--warmup", "0"], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert res.returncode == 0, res.stderr[-2000:]
    lines = [l for l in res.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    line = json.loads(lines[0])
    for key in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better",
                "scaling", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert key in line, key
    assert line["impl"] == "reference" and line["cpu_baseline"]["kind"] == "port" and line["value"] > 0
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["e2e"]["d2h_bytes_per_step"] == 0


def test_wave_quantised_chunks():
    """_chunks_by_tokens with a wave quantum: chunks filled to whole waves of 128-row tiles when that needs fewer
    waves in total than the balanced cut, never more chunks, always complete and within the limit."""
    import numpy as np
    from bayeslms_b200.scorer import _chunks_by_tokens
    rng = np.random.default_rng(0)
    lens = rng.integers(7, 28, size=12800).tolist()
    q = 128 * 74
    base, alt = _chunks_by_tokens(lens, 65536), _chunks_by_tokens(lens, 65536, q)
    waves = lambda cuts: sum(-(-(-(-sum(lens[a:b]) // 128)) // 74) for a, b in cuts)  # noqa: E731
    assert alt[0][0] == 0 and alt[-1][1] == len(lens) and all(a[1] == b[0] for a, b in zip(alt, alt[1:]))
    assert max(sum(lens[a:b]) for a, b in alt) <= 65536 and len(alt) <= len(base)
    assert waves(alt) <= waves(base)
    assert all(sum(lens[a:b]) > 6 * q - 28 for a, b in alt[:-1]) or alt == base
    assert _chunks_by_tokens(lens[:100], 65536, q) == _chunks_by_tokens(lens[:100], 65536)


def test_ranking_agreement_counts_ties_at_the_stated_gap():
    """synth.ranking_agreement: a 1-best that differs between two candidates closer than ``min_gap`` in the reference
    scores is a tie at that tolerance (one_best_within_gap), a real flip is not."""
    from bayeslms_b200 import synth
    ref = [np.array([1.0, 1.01, 5.0]), np.array([2.0, 3.0, 4.0])]
    got = [np.array([1.02, 1.0, 5.0]), np.array([2.0, 3.0, 4.0])]         # utterance 0: near-tie flipped
    a = synth.ranking_agreement(got, ref, min_gap=0.05)
    assert a["one_best"] == 0.5 and a["one_best_within_gap"] == 1.0 and a["pair_order"] == 1.0
    assert abs(a["largest_flipped_gap"] - 0.01) < 1e-9
    bad = [np.array([5.0, 1.0, 0.5]), np.array([2.0, 3.0, 4.0])]          # utterance 0: a real flip
    b = synth.ranking_agreement(bad, ref, min_gap=0.05)
    assert b["one_best_within_gap"] == 0.5 and b["pair_order"] < 1.0


def test_rows32_layout_round_trip():
    """ops.rows32_to_dense inverts the 32-row-block layout of blm_gemm_desc.f32_rows32 (pure index arithmetic)."""
    from bayeslms_b200 import ops
    M, N = 70, 24
    dense = torch.arange(M * N, dtype=torch.float32).view(M, N)
    pad = torch.zeros(96, N)
    pad[:M] = dense
    blocked = pad.view(3, 32, N // 4, 4).permute(0, 2, 1, 3).contiguous().view(96, N)
    assert torch.equal(ops.rows32_to_dense(blocked, M), dense)
    # element (m, col) sits at float offset ((m // 32) * (N // 4) + col // 4) * 128 + (m % 32) * 4 + col % 4
    flat = blocked.reshape(-1)
    for m, col in ((0, 0), (31, 23), (32, 5), (69, 20)):
        off = ((m // 32) * (N // 4) + col // 4) * 128 + (m % 32) * 4 + col % 4
        assert flat[off] == dense[m, col]
