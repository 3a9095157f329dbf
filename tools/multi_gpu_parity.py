#!/usr/bin/env python
"""On-hardware multi-GPU parity (north_star: identical rankings at 1/2/4/8 GPUs): ONE fixed synthetic n-best file is
scored (a) by every rank alone (the single-rank result) and (b) sharded over the N ranks of this torchrun job through
the scorer's file path (utterances / sessions split over ranks, one all-reduce of the score vector); the two score
vectors must be BIT-IDENTICAL on every rank -- posterior mean and K > 0 device-Philox samples, Transformer and LSTM.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 \\
        tools/multi_gpu_parity.py > profiles/r02_multi_gpu_parity_nN.json
"""
import json
import os
import sys
import tempfile

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    from bayeslms_b200 import model as M, synth
    from bayeslms_b200.scorer import score_files
    sys.stdout.flush()
    result_fd = os.dup(1)      # stdout carries the JSON only: NCCL's banner and library chatter go to stderr
    os.dup2(2, 1)
    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    V, NB = 30000, 50
    data = synth.make_nbest(96, NB, V, seed=4321)
    tmpd = tempfile.mkdtemp(prefix=f"blm_par_r{rank}_")
    vp, npth = os.path.join(tmpd, "words.txt"), os.path.join(tmpd, "words_text")
    open(vp, "w").write("\n".join(synth.vocab_lines(V)) + "\n")
    open(npth, "w").write("\n".join(data.words_text()) + "\n")
    results = {}
    cases = [("bayes_tm_FFN_mean", "tm", dict(prec="bf16")), ("bayes_tm_FFN_mean_precise", "tm", dict(prec="bf16x3")),
             ("bayes_tm_FFN_K2_philox", "tm", dict(prec="bf16", K=2, seed=1111)),
             ("gp_tm_3_mean", "gp", dict(prec="bf16")),
             ("bayes_lstm_3_mean", "lstm", dict(prec="bf16", session_size=6)),
             ("bayes_lstm_3_K2_philox", "lstm", dict(prec="bf16", K=2, seed=1111, session_size=6))]
    nets = {}
    for name, fam, kw in cases:
        if fam not in nets:
            torch.manual_seed(1111)      # same weights on every rank (replicated model)
            if fam == "tm":
                net = M.BayesTransformerModel(V, 512, 8, 4096, 6, 0.5, True, "FFN")
            elif fam == "gp":
                net = M.GaussTransformerModel(V, 512, 8, 4096, 6, 0.5, True, 3)
            else:
                net = M.BayesRNNModel("LSTM", V, 1024, 1024, 2, 0.5, True, 3)
            with torch.no_grad():
                net.decoder.bias.uniform_(-0.1, 0.1)
            nets[fam] = net.to(dev).eval()
        net = nets[fam]
        single = score_files(net, npth, vp, None, **kw)                                   # this rank alone
        sharded = score_files(net, npth, vp, None, rank=rank, world=world, **kw) if world > 1 else single
        same = bool(np.array_equal(single, sharded))
        per = lambda s: [s[i * NB:(i + 1) * NB] for i in range(data.n_utts)]  # noqa: E731
        ranks_same = all(np.array_equal(np.argsort(a, kind="stable"), np.argsort(b, kind="stable"))
                         for a, b in zip(per(single), per(sharded)))
        flag = torch.tensor([int(same), int(ranks_same)], device=dev)
        if world > 1:
            dist.all_reduce(flag, op=dist.ReduceOp.MIN)
            # every rank's single-rank vector is also identical to rank 0's (replicas agree bit for bit)
            ref = torch.from_numpy(single).to(dev)
            dist.broadcast(ref, 0)
            flag2 = torch.tensor([int(torch.equal(ref, torch.from_numpy(single).to(dev)))], device=dev)
            dist.all_reduce(flag2, op=dist.ReduceOp.MIN)
        else:
            flag2 = torch.tensor([1])
        results[name] = {"bit_identical_scores": bool(flag[0].item()), "identical_rankings": bool(flag[1].item()),
                         "replicas_agree": bool(flag2[0].item()), "hypotheses": int(len(single)),
                         "score_checksum": float(np.float64(single.astype(np.float64).sum()))}
    if rank == 0:
        ok = all(r["bit_identical_scores"] and r["identical_rankings"] and r["replicas_agree"] for r in results.values())
        os.write(result_fd, (json.dumps({"n_gpus": world, "all_ok": ok, "file": f"{data.n_utts} utterances x {NB}-best, V={V}",
                                         "cases": results}, indent=1) + "\n").encode())
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
