#!/usr/bin/env python
"""Headline benchmark: n-best rescoring tokens/s (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W [--impl reference]

Workload at N=1 (BASELINE.json configs[1]): Bayesian Transformer LM, 6 layers, d=512, FFN=4096,
8 heads, V=30000, T_bayes_pos=FFN, 50-best lists, random-init weights of that architecture,
synthetic n-best lists (bayeslms_b200/synth.py).  One *step* scores UTTS_PER_STEP utterances x
50 hypotheses (~0.2 M scored positions) in balanced packed batches of <= 65536 tokens; with N ranks every
rank scores its own UTTS_PER_STEP utterances (weak scaling, no data-path collective; one score
gather per step).  Reported:

  value        tokens/s, posterior mean, bf16 operands / fp32 accumulate, ids already in HBM
  e2e          same through Rescorer.score_packed_host: pinned host ids -> device, scores -> host
  cli_e2e      same through scorer.score_files: words_text + words.txt on disk -> lmwt.nn on disk (the call the
               reference's stage-6 command line maps to), vocabulary parsed every step
  sampled_k4   K=4 Philox posterior samples (Monte-Carlo predictive), device-resident
  precise      bf16x3 (hi/lo split) mode that holds the 1e-3 parity bar, with its own roofline and slowdown
  roofline     the vocabulary-streaming projection+NLL kernel, 2*M*d*V FLOP per launch over the
               CUDA-event time of its launches, against MEASURED_PEAKS.json (sustained bf16);
               roofline_dominant_kernel = the kernel with the largest share of the step (FFN1+GELU),
               whole_step_roofline = 93.8 MFLOP per token x tokens/s, kernel_rooflines = every GEMM of the step
  ranking_peaked_model   fast vs precise rankings on a model fine-tuned to sharp distributions (evidence.py)
  gp_tm_rescoring / gp_tm_sampled_k1 / gp_tm_strong_scaling   BASELINE config 3 (one fixed 8000-utterance list
               sharded over the ranks for the strong-scaling entry)
  finetune_step  BASELINE config 4: captured V-Transformer step (roofline, with_dropout_0.2, CPU port of the step)
  lstm_rescoring BASELINE config 5 at its per-GPU size and config 1 as its cpu_baseline; recurrence_roofline
  cpu_baseline the CPU oracle port (reference algorithm, torch CPU fp32, batch 1 per hypothesis)
               on a bounded sample of the same lists, rank 0, N=1 only
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

V, D, NHEAD, FF, NLAYERS, NBEST = 30000, 512, 8, 4096, 6, 50
UTTS_PER_STEP = 256
MAX_TOKENS = 65536
WORKLOAD = "Bayesian Transformer LM 6L d512 FFN4096 h8 V30000 T_bayes_pos=FFN, 50-best rescoring"


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        p = json.load(open(path))
        return {"tensor": p.get("bf16_tflops_sustained", 1397.3), "tensor_burst": p.get("bf16_tflops", 1663.3),
                "hbm": p.get("hbm_gbs", 6550.4), "src": "measured"}
    return {"tensor": 1400.0, "tensor_burst": 1590.0, "hbm": 6650.0, "src": "fallback"}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms during the timed region."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows, self.proc, self.index = [], None, index

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-i", str(self.index), "-lms", "50"], stdout=subprocess.PIPE, text=True)
            self.t = threading.Thread(target=lambda: [self.rows.append(l) for l in self.proc.stdout], daemon=True)
            self.t.start()
        except Exception:
            self.proc = None
        return self

    def __exit__(self, *a):
        if self.proc:
            time.sleep(0.25)
            self.proc.terminate()
            self.t.join(2)

    def summary(self):
        sm, mx, reasons = [], 0.0, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            f = [x.strip() for x in r.split(",")]
            try:
                sm.append(float(f[0]))
                mx = max(mx, float(f[1]))
                for n, v in zip(names, f[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(n)
            except (ValueError, IndexError):
                pass
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx or None, "reasons": sorted(reasons),
                "samples": len(sm)}


def build_model(device):
    from bayeslms_b200 import model as M
    torch.manual_seed(1111)
    net = M.BayesTransformerModel(V, D, NHEAD, FF, NLAYERS, 0.5, True, "FFN")
    with torch.no_grad():
        net.decoder.bias.uniform_(-0.1, 0.1)   # exercise the bias term
    return net.to(device).eval()


def bench_finetune(dev, world, steps, warmup=3, dropout=0.0):
    """BASELINE config 4: Variational Transformer (T_v_pos=11 -> both first layers variational, 5 layers),
    d=512, FFN=4096, V=30000; one fine-tune step = CE + KL, backward, clip, SGD momentum on a per-GPU
    batch of 32 x 100 tokens (train.py step), data parallel with one NCCL all-reduce of the gradients."""
    import torch.distributed as dist
    from bayeslms_b200 import model as M
    from bayeslms_b200.trainer import FineTuner
    torch.manual_seed(1111)
    net = M.VTransformerModel(V, D, NHEAD, FF, NLAYERS, dropout, True, "11").to(dev).train()
    ft = FineTuner(net, 0.01, clip=0.25, prec="bf16")
    g = torch.Generator().manual_seed(1111 + int(os.environ.get("RANK", 0)))
    T, B = 100, 32
    x = torch.randint(0, V, (T, B), generator=g).to(dev)
    y = torch.randint(0, V, (T, B), generator=g).to(dev)
    ft.capture(T, B, 1e-3)           # forward+backward and optimiser as two CUDA graphs; noise refreshed per step
    losses = [float(ft.step_captured(x, y, 7 + i)[0]) for i in range(warmup)]
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(steps):
        ft.step_captured(x, y, 100 + i)
    e1.record()
    torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1) / steps], device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    losses.append(float(ft.step_captured(x, y, 999)[0]))
    # algorithmic work of one step: forward + backward = 3 x the forward FLOP of the 5-layer model (SURVEY.md 8d:
    # 83.3 MFLOP per token forward: 5 layers of QKV / o_net / FFN1 / FFN2 projections + the vocabulary projection);
    # bytes: parameters read, gradients written + read, momentum read + written, parameters + bf16 copies written
    n_par = sum(p.numel() for p in net.parameters())
    flop = 3.0 * 83.3e6 * T * B
    byts = n_par * (4 + 4 + 4 + 4 + 4 + 4 + 2.0)
    pk = peaks()
    tf = flop / (ms.item() / 1e3) / 1e12
    in_graph = bool(ft._cap.get("ar_in_graph"))
    ft._cap = None            # release the graphs (captured NCCL work) before anything else touches the communicator
    del ft
    torch.cuda.synchronize()
    return {"workload": "Variational Transformer LM (T_v_pos=11) 5L d512 FFN4096 V30000 fine-tune step, "
                        "32 x 100 tokens per GPU, CE + KL, clip, SGD momentum", "dtype": "bf16",
            "tokens_per_s": T * B * world / (ms.item() / 1e3), "ms_per_step": ms.item(),
            "parallelism": f"dp{world}: replicated weights, " + ("per-layer NCCL all-reduces captured inside the backward graph "
                                                                   "(second stream), embedding / decoder range at the end"
                                                                   if in_graph else "one NCCL all-reduce of the flat gradient buffer per step"),
            "loss_first": losses[0], "loss_last": losses[-1], "dropout": dropout,
            "roofline": {"bound": "tensor", "achieved": tf, "peak": pk["tensor"], "unit": "TFLOP/s", "frac": tf / pk["tensor"],
                         "flop_per_step": flop, "hbm_bytes_per_step": byts,
                         "floor_ms": 1e3 * (flop / (pk["tensor"] * 1e12) + byts / (pk["hbm"] * 1e9)),
                         "note": "whole step (154 launches in two CUDA graphs, weight gradients on a second stream); "
                                 "3 x forward FLOP of the model; M = 3200 tokens fills 16-128 of 148 SMs per GEMM"}}


def bench_lstm(dev, world=1, rank=0, steps=1):
    """BASELINE config 5 at its real per-GPU size: Bayesian LSTM 2x1024 (emb 1024, L_bayes_pos=3), V=30000, 100-best
    lists, 10 000 utterances as 100 sessions of 100 utterances (hidden carry through hypothesis #0, score.py:271-274)
    over 8 GPUs = 12 sessions x 100 utterances per GPU (session-sharded, weak scaling, no data-path collective), mean
    and K = 8 Philox samples.  End to end: flat host id arrays -> lock-step batches -> persistent recurrence kernel ->
    vocabulary NLL -> scores on the host; wall clock between two barriers, max over ranks.  Also the recurrence
    kernel's roofline (effective GB/s on the algorithmic bytes, W_hh counted once per step although it
    never leaves shared memory) from CUDA-event brackets of its launches."""
    import torch.distributed as dist
    from bayeslms_b200 import model as M, ops, synth
    from bayeslms_b200.scorer import Rescorer
    torch.manual_seed(1111)
    net = M.BayesRNNModel("LSTM", V, 1024, 1024, 2, 0.5, True, 3).to(dev).eval()
    n_sess, per_sess, nbest = 12, 100, 100
    data = synth.make_nbest(n_sess * per_sess, nbest, V, seed=1112 + rank)
    tok, tgt, _, offs = data.flat_host()
    utt = np.repeat(np.arange(n_sess * per_sess), [len(u) for u in data.hyps])
    sess_of, utt_of = (utt // per_sess).astype(np.int32), (utt % per_sess).astype(np.int32)
    n_tok = torch.tensor([float(data.n_tokens())], device=dev)
    if world > 1:
        dist.all_reduce(n_tok)
    out = {}
    pk = peaks()
    for name, kw in (("mean", {}), ("sampled_k8", {"K": 8, "seed": 1111})):
        rs = Rescorer(net, prec="bf16", max_tokens=MAX_TOKENS, **kw)
        if name == "mean":
            rs.score_sessions_flat(tok, tgt, offs, sess_of, utt_of)       # warm-up (plans, workspaces)
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        ops.STATS.timing, ops.STATS.bytes = ({}, {}) if name == "mean" else (None, {})
        t0 = time.perf_counter()
        for _ in range(steps):
            rs.score_sessions_flat(tok, tgt, offs, sess_of, utt_of)
        torch.cuda.synchronize()
        dt = torch.tensor([(time.perf_counter() - t0) / steps], device=dev)
        if world > 1:
            dist.all_reduce(dt, op=dist.ReduceOp.MAX)
        out[name] = {"tokens_per_s": n_tok.item() / dt.item(), "ms": dt.item() * 1e3}
        if name == "mean":
            timing, ops.STATS.timing = ops.STATS.timing, None
            ev = timing.get("lstm_layer", [])            # the lock-step hypothesis batches (B up to 2048 rows)
            ms_k = sum(a.elapsed_time(b) for a, b, _ in ev)
            flop = sum(w for _, _, w in ev)
            byts = ops.STATS.bytes.get("lstm_layer", 0.0)
            tot = sum(sum(a.elapsed_time(b) for a, b, _ in v) for v in timing.values()) or 1.0
            if ms_k:
                chain = timing.get("lstm_layer:chain", [])
                out["recurrence_roofline"] = {
                    "kernel": "lstm_pair_kernel<16,2> (persistent recurrence of cta_group::2 pairs, W_hh resident in shared "
                              "memory), the launches of the lock-step hypothesis batches",
                    "bound": "hbm", "achieved": byts / (ms_k / 1e3) / 1e9, "peak": pk["hbm"], "unit": "GB/s",
                    "frac": byts / (ms_k / 1e3) / 1e9 / pk["hbm"], "launches": len(ev), "share_of_kernel_time": ms_k / tot,
                    "tensor_tflops": flop / (ms_k / 1e3) / 1e12, "tensor_frac": flop / (ms_k / 1e3) / 1e12 / pk["tensor"],
                    # ncu --set full of one 40-step launch at B = 2048 (profiles/r02d_lstm_pair_vs_single_ncu.txt): DRAM
                    # read 1371 MB (= gates_x, once) + write 241 MB; W_hh is never re-read
                    "traffic": (1371.2e6 + 240.8e6) / 40, "traffic_unit": "DRAM bytes per step at B=2048 (ncu, read + write)",
                    "hypothesis0_chains": {"kernel": "lstm_layer_kernel<16,1,2>, one row per session, all K blocks of h at once",
                                           "launches": len(chain), "ms": sum(a.elapsed_time(b) for a, b, _ in chain),
                                           "share_of_kernel_time": sum(a.elapsed_time(b) for a, b, _ in chain) / tot,
                                           "note": "sequential by definition (score.py:271-274): latency-bound, ~10 us per step"},
                    "note": "effective bandwidth on the algorithmic bytes of SURVEY.md 8d (W_hh 8 MiB counted once per step "
                            "+ gates_x + h, c); it may exceed what DRAM delivers because W_hh stays in shared memory"}
            out["kernel_ms_total"] = round(tot / steps, 2)
            out["kernel_time_shares"] = {k: round(sum(a.elapsed_time(b) for a, b, _ in v) / tot, 4)
                                         for k, v in sorted(timing.items(), key=lambda kv: -sum(a.elapsed_time(b) for a, b, _ in kv[1]))[:6]}
    # the 1e-3-parity mode (bf16x3, the scorer CLI's default) on the first 3 sessions of the same lists
    n_h = 3 * per_sess * nbest
    m_p = int(offs[n_h])
    sub = (tok[:m_p], tgt[:m_p], offs[:n_h + 1], sess_of[:n_h], utt_of[:n_h])
    rs_p = Rescorer(net, prec="bf16x3", max_tokens=MAX_TOKENS)
    rs_f = Rescorer(net, prec="bf16", max_tokens=MAX_TOKENS)
    res = {}
    for name, r in (("bf16x3", rs_p), ("bf16", rs_f)):
        r.score_sessions_flat(*sub)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        sc = r.score_sessions_flat(*sub)
        torch.cuda.synchronize()
        res[name] = (m_p / (time.perf_counter() - t0), sc)
    out["precise"] = {"tokens_per_s": res["bf16x3"][0], "dtype": "bf16x3", "slowdown_vs_bf16_same_sample": res["bf16"][0] / res["bf16x3"][0],
                      "max_abs_score_diff_vs_bf16": float(np.abs(res["bf16x3"][1] - res["bf16"][1]).max()),
                      "sample": f"3 sessions x {per_sess} utterances x {nbest}-best ({m_p} tokens) per GPU; with 3 instead of 12 "
                                "sessions the sequential hypothesis-#0 chains weigh more (3-row batches)"}
    out["workload"] = (f"Bayesian LSTM 2x1024 L_bayes_pos=3 V30000, {nbest}-best, {n_sess} sessions x {per_sess} utterances "
                       f"per GPU ({int(n_tok.item())} tokens over {world} GPU(s); config 5 = 100 sessions over 8 GPUs), end to "
                       "end from flat host id arrays incl. batch packing (wall clock, max over ranks)")
    return out


def cpu_finetune_tokens_per_s(n_seq=4, steps=2):
    """BASELINE config 4 on the host cores: the oracle's restatement of one train.py step (forward in train mode with
    seeded hidden noise, CE + KL, autograd backward, clip, SGD momentum) of the 5-layer Variational Transformer on
    ``n_seq`` sequences of 100 tokens (the GPU step takes 32), torch CPU fp32 on all cores."""
    from bayeslms_b200 import model as M
    from oracle import bayeslm_oracle as O
    torch.manual_seed(1111)
    net = M.VTransformerModel(V, D, NHEAD, FF, NLAYERS, 0.0, True, "11")
    sd = {k: v.detach().clone() for k, v in net.state_dict().items()}
    cfg = O.Config(family="v_tm", v_pos=3, ntoken=V, ninp=D, nhead=NHEAD, nhid=FF, nlayers=NLAYERS)   # "11" = 3
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    T = 100
    g = torch.Generator().manual_seed(3)
    leaf = {k: (v.clone().requires_grad_(True) if v.dtype.is_floating_point and k != "pos_encoder.pe" else v)
            for k, v in sd.items() if k != "decoder.weight"}
    leaf["decoder.weight"] = leaf["encoder.weight"]
    params = [v for k, v in leaf.items() if k != "decoder.weight" and torch.is_tensor(v) and v.requires_grad]
    opt = torch.optim.SGD(params, lr=0.01, momentum=0.9)
    dt = []
    for it in range(steps + 1):
        x = torch.randint(0, V, (T, n_seq), generator=g)
        y = torch.randint(0, V, (T, n_seq), generator=g)
        eps = {f"layer{i}": torch.randn(T, n_seq, D, generator=g) * 0.1 for i in (0, 1)}
        t0 = time.perf_counter()
        opt.zero_grad(set_to_none=True)
        loss, _, _ = O.finetune_loss(leaf, x, y.view(-1), cfg, eps, 1e-3)
        loss.backward()
        torch.nn.utils.clip_grad_norm_(params, 0.25)
        opt.step()
        if it:                      # the first step warms the allocator and the thread pool
            dt.append(time.perf_counter() - t0)
    per = sum(dt) / len(dt)
    return {"value": T * n_seq / per, "unit": "tokens/s", "cores": cores, "kind": "port",
            "sample": f"{steps} steps of {n_seq} x {T} tokens ({per:.2f} s per step), oracle port of the train.py step "
                      "(train-mode forward, CE + KL, autograd backward, clip, SGD momentum), torch CPU fp32"}


def cpu_lstm_tokens_per_s(n_utts=100, nbest=20, budget_s=20.0):
    """BASELINE config 1: Bayesian LSTM 2x1024 (emb 1024, L_bayes_pos=3), V=30000, 20-best lists of 100 utterances, the
    reference loop on the host cores (oracle port: one hypothesis at a time, batch 1, hidden carried through hypothesis
    #0, torch CPU fp32).  Bounded: stops at the first utterance boundary after ``budget_s`` seconds."""
    from bayeslms_b200 import model as M, synth
    from oracle import bayeslm_oracle as O
    torch.manual_seed(1111)
    net = M.BayesRNNModel("LSTM", V, 1024, 1024, 2, 0.5, True, 3)
    sd = {k: v.detach() for k, v in net.state_dict().items()}
    cfg = O.Config(family="bayes_lstm", bayes_pos=3, ntoken=V, ninp=1024, nhid=1024, nlayers=2)
    data = synth.make_nbest(n_utts, nbest, V, seed=1111)
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    hidden = O.init_hidden(cfg, 1)
    toks, utts, t0 = 0, 0, time.perf_counter()
    with torch.no_grad():
        for utt in data.tokenised():
            first = None
            for x, y in utt:
                lg, nh = O.rnn_forward(sd, torch.tensor(x).view(-1, 1), hidden, cfg)
                O.sentence_nll(lg, torch.tensor(y))
                first = first or nh
                toks += len(x)
            hidden = first
            utts += 1
            if time.perf_counter() - t0 > budget_s:
                break
    dt = time.perf_counter() - t0
    return {"value": toks / dt, "unit": "tokens/s", "cores": cores, "kind": "port",
            "sample": f"first {utts} of {n_utts} utterances x {nbest}-best ({toks} tokens, {dt:.1f} s), oracle port of the "
                      "reference loop: batch 1 per hypothesis, hidden carry, torch CPU fp32"}


def cpu_port_tokens_per_s(data, n_utts, state_dict):
    """The reference algorithm on the host: one hypothesis at a time, batch 1, fp32 torch CPU."""
    from oracle import bayeslm_oracle as O
    cfg = O.Config(family="bayes_tm", bayes_pos="FFN", ntoken=V, ninp=D, nhead=NHEAD, nhid=FF, nlayers=NLAYERS)
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    toks = 0
    t0 = time.perf_counter()
    with torch.no_grad():
        for utt in data.tokenised(0, n_utts):
            for x, y in utt:
                lg = O.transformer_forward(state_dict, torch.tensor(x).view(-1, 1), cfg)
                O.sentence_nll(lg, torch.tensor(y))
                toks += len(x)
    dt = time.perf_counter() - t0
    return toks / dt, toks, dt, cores


def run_reference(args, rank, world, emit):
    """--impl reference: the reference's CPU implementation of the path (oracle port; the Python
    reference cannot travel to the GPU box), all host threads, bounded sample per step."""
    if rank != 0:
        return
    from bayeslms_b200 import synth
    from bayeslms_b200 import model as M   # parameter container only: same random-init weights, no GPU
    torch.manual_seed(1111)
    net = M.BayesTransformerModel(V, D, NHEAD, FF, NLAYERS, 0.5, True, "FFN")
    sd = {k: v.detach() for k, v in net.state_dict().items()}
    n_utts = 2  # ~1.6 k tokens, a few seconds per step on 8+ cores
    data = synth.make_nbest(n_utts, NBEST, V, seed=1111)
    for _ in range(max(args.warmup, 0) and 1):
        cpu_port_tokens_per_s(data, 1, sd)
    vals, tok = [], 0
    t0 = time.perf_counter()
    for _ in range(args.steps):
        v, tok, _, cores = cpu_port_tokens_per_s(data, n_utts, sd)
        vals.append(v)
    dt = time.perf_counter() - t0
    value = tok * args.steps / dt
    sample = f"{n_utts} utterances x {NBEST}-best ({tok} tokens) per step, batch 1 per hypothesis"
    emit({
        "impl": "reference", "metric": "nbest_rescoring_tokens_per_sec", "value": value, "unit": "tokens/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "sample": sample},
        "cpu_baseline": {"value": value, "unit": "tokens/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "tokens/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    })


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", type=str, default="ours", choices=["ours", "reference"])
    ap.add_argument("--utts-per-step", type=int, default=UTTS_PER_STEP)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    local_rank = int(os.environ.get("LOCAL_RANK", 0))
    # stdout carries exactly ONE line, the JSON result: everything else that writes to fd 1 (NCCL's version banner,
    # library chatter) is sent to stderr, and the result is written to the saved descriptor at the end
    sys.stdout.flush()
    result_fd = os.dup(1)
    os.dup2(2, 1)

    def emit(obj):
        os.write(result_fd, (json.dumps(obj) + "\n").encode())

    if args.impl == "reference":
        run_reference(args, rank, world, emit)
        return 0

    import torch.distributed as dist
    from bayeslms_b200 import _lib, ops, synth
    from bayeslms_b200.engine import PackedBatch
    from bayeslms_b200.scorer import Rescorer, _chunks_by_tokens, wave_quantum

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    _lib.init(local_rank)
    warmup = max(args.warmup, 3)
    net = build_model(dev)
    # every rank scores its own utterances of one global synthetic set (weak scaling)
    data = synth.make_nbest(args.utts_per_step * world, NBEST, V, seed=1111)
    lo, hi = rank * args.utts_per_step, (rank + 1) * args.utts_per_step
    tok, tgt, pos, offs = data.flat_host(lo, hi)
    n_tokens, n_hyp = int(offs[-1]), len(offs) - 1
    lengths = np.diff(offs)
    chunks = _chunks_by_tokens(lengths, MAX_TOKENS, wave_quantum())

    def device_batches():
        out = []
        for a, b in chunks:
            t0, t1 = int(offs[a]), int(offs[b])
            mk = lambda x: torch.from_numpy(np.ascontiguousarray(x)).to(dev)  # noqa: E731
            out.append(PackedBatch(mk(tok[t0:t1]), mk(tgt[t0:t1]), mk(pos[t0:t1]), mk(offs[a:b + 1] - offs[a]),
                                   int(lengths[a:b].max()), t1 - t0, b - a))
        return out

    batches = device_batches()
    gather_buf = torch.zeros(world * n_hyp, dtype=torch.float32, device=dev) if world > 1 else None

    def step(prec="bf16", K=0, seed=None):
        outs = [net.score(b, prec=prec, K=K, seed=seed) for b in batches]
        res = torch.cat(outs)
        if world > 1:  # the one collective of the path: gather the score vector
            dist.all_gather_into_tensor(gather_buf, res)
        return res

    def timed(fn, steps, sync_each=False):
        """max-over-ranks device time of `steps` calls (CUDA events on the launching stream)."""
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return ms.item()

    for _ in range(warmup):
        step()
    # ---- timed region: posterior mean, bf16, ids resident in HBM
    ops.STATS.launches = 0
    ops.STATS.timing = {}
    with ClockSampler(local_rank) as clk:
        ms = timed(step, args.steps)
    launches = ops.STATS.launches
    timing, ops.STATS.timing = ops.STATS.timing, None
    value = n_tokens * world * args.steps / (ms / 1e3)

    # per-kernel shares from the event brackets recorded inside the timed region
    shares, nll_ms, nll_flops, nll_n = {}, 0.0, 0.0, 0
    for name, evs in timing.items():
        t = sum(a.elapsed_time(b) for a, b, _ in evs)
        shares[name] = t
        if name == "vocab_nll":
            nll_ms, nll_flops, nll_n = t, sum(w for _, _, w in evs), len(evs)
    tot = sum(shares.values()) or 1.0
    pk = peaks()
    # achieved TFLOP/s of every tensor-bound kernel of the step (algorithmic FLOP of its launches / event time)
    krf = {}
    for name, evs in timing.items():
        w = sum(x for _, _, x in evs)
        if w > 0 and shares[name] > 0:
            tf = w / (shares[name] / 1e3) / 1e12
            krf[name] = {"achieved": round(tf, 1), "frac": round(tf / pk["tensor"], 3), "launches": len(evs)}
    achieved = (nll_flops / nll_n) / (nll_ms / nll_n / 1e3) / 1e12 if nll_n else 0.0
    # the kernel with the largest share of the step's time (FFN1 + GELU in r01), next to the vocabulary kernel
    dom = max((n for n in timing if sum(w for _, _, w in timing[n]) > 0), key=lambda n: shares[n], default=None)
    dom_rf = None
    if dom is not None:
        dw, dn = sum(w for _, _, w in timing[dom]), len(timing[dom])
        dtf = dw / (shares[dom] / 1e3) / 1e12
        dom_rf = {"kernel": dom, "bound": "tensor", "achieved": dtf, "peak": pk["tensor"], "unit": "TFLOP/s",
                  "frac": dtf / pk["tensor"], "launches": dn, "avg_launch_ms": shares[dom] / dn, "share_of_step": shares[dom] / tot}

    # ---- e2e: host ids -> pinned staging -> device -> scores back on the host, every step
    rs = Rescorer(net, prec="bf16", max_tokens=MAX_TOKENS)
    for _ in range(2):
        rs.score_packed_host(tok, tgt, pos, offs)
    rs.h2d_bytes = rs.d2h_bytes = 0

    def e2e_step():
        res = rs.score_packed_host(tok, tgt, pos, offs)
        if world > 1:
            dist.all_gather_into_tensor(gather_buf, torch.from_numpy(res).to(dev))

    e2e_ms = timed(e2e_step, args.steps)
    e2e_val = n_tokens * world * args.steps / (e2e_ms / 1e3)
    h2d, d2h = rs.h2d_bytes // args.steps, rs.d2h_bytes // args.steps

    # ---- cli_e2e: what stage 6 of the pipeline runs -- words_text + words.txt on disk -> lmwt.nn on disk, through the
    # scorer's file path (C tokeniser into pinned staging, batches overlapped with the kernels, C formatter)
    import tempfile
    from bayeslms_b200.scorer import score_files
    tmpd = tempfile.mkdtemp(prefix=f"blm_cli_r{rank}_")
    vocab_path, nbest_path, out_path = (os.path.join(tmpd, n) for n in ("words.txt", "words_text", "lmwt.nn"))
    with open(vocab_path, "w") as f:
        f.write("\n".join(synth.vocab_lines(V)) + "\n")
    with open(nbest_path, "w") as f:
        f.write("\n".join(data.words_text(lo, hi)) + "\n")

    def cli_step():
        res = score_files(net, nbest_path, vocab_path, out_path, rescorer=rs)
        if world > 1:
            dist.all_gather_into_tensor(gather_buf, torch.from_numpy(res).to(dev))
        return res

    cli_scores = cli_step()
    cli_step()
    cli_ms = timed(cli_step, args.steps)
    cli_val = n_tokens * world * args.steps / (cli_ms / 1e3)
    cli_same = bool(np.array_equal(cli_scores, rs.score_packed_host(tok, tgt, pos, offs)))
    cli_lines = sum(1 for _ in open(out_path))

    # ---- K = 4 sampled and precise-mode numbers (same lists)
    k4_steps = max(1, args.steps // 2)
    step(K=4, seed=1111)
    ops.STATS.timing = {}
    k4_ms = timed(lambda: step(K=4, seed=1111), k4_steps)
    k4_timing, ops.STATS.timing = ops.STATS.timing, None
    sg = k4_timing.get("gemm_sampled:ffn2", [])
    sg_ms = sum(a.elapsed_time(b) for a, b, _ in sg)
    sg_tf = (sum(w for _, _, w in sg) / (sg_ms / 1e3) / 1e12) if sg_ms else 0.0
    step(prec="bf16x3")
    ops.STATS.timing = {}
    px_ms = timed(lambda: step(prec="bf16x3"), k4_steps)
    px_timing, ops.STATS.timing = ops.STATS.timing, None
    # precise mode issues three bf16 products per GEMM (hi*lo + lo*hi + hi*hi): the tensor pipe does 3 x the
    # algorithmic FLOP, which is what its roofline fraction is measured on
    px_flop = sum(w for evs in px_timing.values() for _, _, w in evs)     # (K-concatenated segments are already counted)
    px_gemm_ms = sum(a.elapsed_time(b) for evs in px_timing.values() for a, b, w in evs if w > 0)
    px_tf = px_flop / (px_gemm_ms / 1e3) / 1e12 if px_gemm_ms else 0.0

    # ---- BASELINE config 3: GP Transformer (T_gauss_pos=3) on the same lists, posterior mean, bf16
    from bayeslms_b200 import model as M
    torch.manual_seed(1111)
    gp_net = M.GaussTransformerModel(V, D, NHEAD, FF, NLAYERS, 0.5, True, 3).to(dev).eval()

    def gp_step():
        res = torch.cat([gp_net.score(b, prec="bf16") for b in batches])
        if world > 1:
            dist.all_gather_into_tensor(gather_buf, res)

    gp_step()
    gp_ms = timed(gp_step, k4_steps)
    # ... and with gpnn.sample = True, K = 1: coefficients, weights and bias of the GP unit drawn per call (Philox)
    gp_net.transformerlayers[0].gpnn.sample = True

    def gp_sampled_step():
        res = torch.cat([gp_net.score(b, prec="bf16", K=1, seed=1111) for b in batches])
        if world > 1:
            dist.all_gather_into_tensor(gather_buf, res)

    gp_sampled_step()
    gp_s_ms = timed(gp_sampled_step, k4_steps)
    gp_net.transformerlayers[0].gpnn.sample = False

    # ---- STRONG scaling (BASELINE config 3 as written): ONE fixed 8000-utterance 50-best list, utterances sharded over
    # the N ranks, flat host ids -> device -> scores -> one all-gather of the full score vector on every rank
    strong_utts, blk = 8000, 250          # the list is 32 independently seeded blocks: a rank only generates its shard
    per = -(-strong_utts // world)
    per = -(-per // blk) * blk
    s_lo, s_hi = min(rank * per, strong_utts), min((rank + 1) * per, strong_utts)
    parts = [synth.make_nbest(blk, NBEST, V, seed=2222 + b).flat_host() for b in range(s_lo // blk, s_hi // blk)]
    stok = np.concatenate([p[0] for p in parts])
    stgt = np.concatenate([p[1] for p in parts])
    spos = np.concatenate([p[2] for p in parts])
    soffs = np.concatenate([[0]] + [p[3][1:].astype(np.int64) + sum(int(q[3][-1]) for q in parts[:i])
                                    for i, p in enumerate(parts)]).astype(np.int32)
    strong_local = torch.tensor([float(soffs[-1])], device=dev)
    if world > 1:
        dist.all_reduce(strong_local)
    strong_tokens = int(strong_local.item())
    rs_gp = Rescorer(gp_net, prec="bf16", max_tokens=MAX_TOKENS)
    sgather = torch.zeros(world * per * NBEST, dtype=torch.float32, device=dev)
    spad = torch.zeros(per * NBEST, dtype=torch.float32, device=dev)

    def strong_step():
        res = rs_gp.score_packed_host(stok, stgt, spos, soffs)
        if world > 1:
            spad[:len(res)].copy_(torch.from_numpy(res), non_blocking=True)
            dist.all_gather_into_tensor(sgather, spad)

    strong_step()
    strong_ms = timed(strong_step, 1)
    del gp_net, rs_gp

    # ---- fast-vs-precise agreement on this step's lists (ranking evidence)
    fast = step().float().cpu().numpy()
    precise = step(prec="bf16x3").float().cpu().numpy()
    per_utt = lambda s: [s[i * NBEST:(i + 1) * NBEST] for i in range(args.utts_per_step)]  # noqa: E731
    _, picks_f = synth.wer(data, per_utt(fast), lo=lo)
    wer_p, picks_p = synth.wer(data, per_utt(precise), lo=lo)

    # ---- ranking evidence on a PEAKED model (a random-init LM ranks by length): bayeslms_b200.evidence fine-tunes a
    # Bayesian Transformer with the BASELINE layer sizes on a synthetic Markov corpus (~1.5 s), then fast vs precise
    # mode on chain-based 50-best lists; both modes against the CPU oracle: tests/test_gpu_ranking.py,
    # profiles/r02_ranking_evidence.json
    from bayeslms_b200 import evidence
    pk_net, pk_mk, pk_losses = evidence.peaked_model(device=dev)
    pk_data = pk_mk.nbest(100, NBEST, seed=5 + rank)
    rank_rep, _, _ = evidence.fast_vs_precise(pk_net, pk_data)
    rank_rep["model"] = (f"BayesTransformerModel FFN 2L d512 FF4096 h8 V500, 1000 fine-tune steps on a 3-successor Markov corpus "
                         f"(loss {pk_losses[0]:.2f} -> {pk_losses[-1]:.2f} nats/token), 100 x {NBEST}-best chain-based lists per rank")
    rank_rep["tie_policy"] = "scores compared on the %.4f grid of lmwt.nn, ties broken by hypothesis index"
    del pk_net

    finetune = bench_finetune(dev, world, max(10, args.steps))
    finetune["with_dropout_0.2"] = {k: v for k, v in bench_finetune(dev, world, max(10, args.steps), dropout=0.2).items()
                                    if k in ("ms_per_step", "tokens_per_s", "dropout")}
    lstm = bench_lstm(dev, world, rank)

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        lstm["cpu_baseline"] = cpu_lstm_tokens_per_s()          # BASELINE config 1: 20-best x 100 utterances on the host
        finetune["cpu_baseline"] = cpu_finetune_tokens_per_s()  # BASELINE config 4: the train.py step on the host
        sd = {k: v.detach().cpu() for k, v in net.state_dict().items()}
        n_cpu = 40                                      # ~30 k tokens: 8-15 s on the box's host cores
        v_cpu, t_cpu, dt_cpu, cores = cpu_port_tokens_per_s(data, n_cpu, sd)
        cpu = {"value": v_cpu, "unit": "tokens/s", "cores": cores, "kind": "port",
               "sample": f"first {n_cpu} utterances x {NBEST}-best of the same lists ({t_cpu} tokens, {dt_cpu:.1f} s), "
                         "oracle port of the reference loop: batch 1 per hypothesis, torch CPU fp32"}

    if rank == 0:
        line = {
            "metric": "nbest_rescoring_tokens_per_sec", "value": value, "unit": "tokens/s", "n_gpus": world,
            "steps": args.steps, "warmup": warmup, "ms_per_step": ms / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": {"workload": WORKLOAD, "posterior": "mean", "utterances_per_step_per_gpu": args.utts_per_step,
                       "tokens_per_step_per_gpu": n_tokens, "max_tokens_per_batch": MAX_TOKENS,
                       "parallelism": f"n-best sharded over {world} rank(s), replicated weights, one score gather",
                       "l2": "activations per batch (>= 0.5 GB) exceed the 126 MB L2; no explicit flush"},
            "e2e": {"value": e2e_val, "unit": "tokens/s", "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                    "ms_per_step": e2e_ms / args.steps},
            "cli_e2e": {"value": cli_val, "unit": "tokens/s", "ms_per_step": cli_ms / args.steps,
                        "ratio_to_e2e": cli_val / e2e_val if e2e_val else None,
                        "path": "words_text + words.txt on disk -> lmwt.nn on disk (bayeslms_b200.scorer.score_files, the "
                                "call `python -m bayeslms_b200.scorer` makes), vocabulary parsed every step",
                        "scores_equal_packed_path": cli_same, "lines_written": cli_lines},
            "gpu_launches": launches,
            "clocks": clk.summary(),
            "roofline": {"kernel": "gemm_kernel<256,4,EPI_NLL> (vocab projection + online LSE + target gather)",
                         "bound": "tensor", "achieved": achieved, "peak": pk["tensor"], "unit": "TFLOP/s",
                         "frac": achieved / pk["tensor"] if pk["tensor"] else None,
                         # dram__bytes_read + write of one launch at M = 52950 rows, ncu --set full
                         # (profiles/r01as_top_kernels_ncu_full.txt: 85.38 MB read + 3.35 MB written); algorithmic
                         # bytes at that M: V*d*2 + M*d*2 + 12*M = 85.6 MB
                         "traffic": 88725504, "traffic_unit": "bytes/launch at M=52950",
                         "peak_source": pk["src"] + " bf16_tflops_sustained", "launches": nll_n,
                         "avg_launch_ms": nll_ms / nll_n if nll_n else None,
                         "share_of_step": nll_ms / tot},
            "roofline_dominant_kernel": dom_rf,
            "whole_step_roofline": {"bound": "tensor", "achieved": 93.8e6 * value / 1e12, "peak": pk["tensor"], "unit": "TFLOP/s",
                                    "frac": 93.8e6 * value / 1e12 / pk["tensor"] / world,
                                    "note": "93.8 MFLOP per token (SURVEY.md 8d) x tokens/s, per GPU"},
            "kernel_time_shares": {k: round(v / tot, 4) for k, v in sorted(shares.items(), key=lambda kv: -kv[1])},
            "kernel_rooflines": {"unit": "TFLOP/s", "peak": pk["tensor"], **krf},
            "sampled_k4": {"value": n_tokens * world * k4_steps / (k4_ms / 1e3), "unit": "tokens/s", "K": 4,
                           "noise": "Philox4x32-10 on device",
                           # north_star (a): the reparameterised-weight GEMM of the sampled FFN (layer 0 linear2,
                           # [M, 512] x [512, 4096]^T), W~ drawn inside the launch; 2*M*N*K FLOP per launch over the
                           # CUDA-event time of its launches in this K = 4 run
                           "sampled_gemm_roofline": {
                               "kernel": "gemm_kernel + generate_weights (blm_gemm_sampled, generate-once)",
                               "bound": "tensor", "achieved": sg_tf, "peak": pk["tensor"], "unit": "TFLOP/s",
                               "frac": sg_tf / pk["tensor"] if pk["tensor"] else None, "launches": len(sg),
                               "avg_launch_ms": sg_ms / len(sg) if sg else None}},
            "precise": {"value": n_tokens * world * k4_steps / (px_ms / 1e3), "unit": "tokens/s", "dtype": "bf16x3",
                        "slowdown_vs_bf16": (px_ms / k4_steps) / (ms / args.steps),
                        "roofline": {"bound": "tensor", "achieved": px_tf, "peak": pk["tensor"], "unit": "TFLOP/s",
                                     "frac": px_tf / pk["tensor"], "algorithmic_tflops": px_tf / 3.0,
                                     "note": "all GEMM-shaped kernels of the step; the tensor pipe does 3 x the algorithmic "
                                             "FLOP (hi*lo + lo*hi + hi*hi), which is what `achieved` counts"},
                        "max_abs_score_diff_vs_bf16": float(np.abs(fast - precise).max()),
                        "one_best_agreement": float(np.mean(np.asarray(picks_f) == np.asarray(picks_p))),
                        "synthetic_wer": wer_p},
            "ranking_peaked_model": rank_rep,
            "gp_tm_strong_scaling": {"value": strong_tokens / (strong_ms / 1e3), "unit": "tokens/s", "ms": strong_ms,
                                     "scaling": "strong", "utterances": strong_utts, "tokens": strong_tokens,
                                     "workload": f"GP Transformer LM T_gauss_pos=3, ONE fixed {strong_utts}-utterance 50-best list "
                                                 f"sharded over {world} rank(s), host ids -> scores + all-gather of the full vector"},
            "gp_tm_sampled_k1": {"value": n_tokens * world * k4_steps / (gp_s_ms / 1e3), "unit": "tokens/s",
                                 "workload": "same lists, gpnn.sample = True, one Philox sample of (coef, weights, bias) per call"},
            "gp_tm_rescoring": {"value": n_tokens * world * k4_steps / (gp_ms / 1e3), "unit": "tokens/s",
                                "workload": "GP Transformer LM T_gauss_pos=3 (GP activation mixture in layer 0), same lists, "
                                            "posterior mean, bf16"},
            "finetune_step": finetune,
            "lstm_rescoring": lstm,
            "cpu_baseline": cpu,
        }
        emit(line)
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
