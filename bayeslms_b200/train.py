"""Epoch-level fine-tuning loop on B200: the host side of ``steps/pytorchnn/train.py`` around the CUDA
fine-tune step (:class:`bayeslms_b200.trainer.FineTuner`).

What is kept from the reference (file:line = steps/pytorchnn/train.py unless noted):
  * corpus format and tokenisation -- ``words.txt`` with two columns, index = order of first occurrence;
    every line of ``train/valid/test.txt`` gets ``<s>`` appended, OOV -> ``<unk>`` (data.py:14-54);
  * ``batchify`` (trim to a multiple of the batch size, column-major streams, :164-176) and ``get_batch``
    (``seq_len`` rows, targets shifted by one, :293-297);
  * one step = CE + KL * seq_len / len(train_data) (:335-399), global-norm clip (:419), SGD momentum 0.9
    without weight decay (:466); LSTM state carried (detached) across batches (:316-321);
  * ``evaluate`` = sum over batches of len(data) * mean CE, divided by len(source) - 1 (:440-457);
  * the schedule: save the state_dict when the validation loss improves, otherwise halve the learning rate,
    build a fresh optimiser (momentum reset) and reload the best checkpoint; stop after 8 such reloads
    (:470-512); finally reload the best checkpoint and report the test loss (:518-530);
  * the CLI flags of :28-103, including ``--prior True --prior_path DIR`` (start from DIR/model.pt, keeping the
    tensors both sides have, :239-258) and the ``--mark base-<frac>set`` corpus pruning (:151-165), so the stage-1
    command lines of ``run_nnlm_ami_{tm,lstm}.sh`` parse and run unchanged.
  * ``--dropout`` is the reference's: the step draws Philox keep masks at every nn.Dropout site of the training
    forward (bayeslms_b200.trainer), seeded per (seed, epoch, batch).
There is no CPU path: the model must live on a B200.
"""
from __future__ import annotations

import argparse
import math
import os
import time
from typing import List, Optional

import torch

from . import engine
from .engine import PackedBatch
from .scorer import read_vocab

_USE_GRAPHS = os.environ.get("BLM_TRAIN_NO_GRAPHS") is None     # A/B switch: plain launches instead of graph replay


# ------------------------------------------------------------------ data (data.py:9-54, train.py:164-183,293-297)
class Corpus:
    def __init__(self, path: str):
        self.word2idx = read_vocab(os.path.join(path, "words.txt"))
        self.train = self.tokenize(os.path.join(path, "train.txt"))
        self.valid = self.tokenize(os.path.join(path, "valid.txt"))
        self.test = self.tokenize(os.path.join(path, "test.txt"))

    def __len__(self):
        return len(self.word2idx)

    def tokenize(self, path: str) -> torch.Tensor:
        if not os.path.exists(path):
            raise FileNotFoundError(path)
        unk = self.word2idx["<unk>"]
        ids: List[int] = []
        with open(path, "r", encoding="utf-8") as f:
            for line in f:
                ids.extend(self.word2idx.get(w, unk) for w in line.split() + ["<s>"])
        return torch.tensor(ids, dtype=torch.int64)


def batchify(data: torch.Tensor, bsz: int, device=None) -> torch.Tensor:
    nbatch = data.size(0) // bsz
    data = data.narrow(0, 0, nbatch * bsz).view(bsz, -1).t().contiguous()
    return data if device is None else data.to(device)


def get_batch(source: torch.Tensor, i: int, seq_len: int):
    n = min(seq_len, len(source) - 1 - i)
    return source[i:i + n], source[i + 1:i + 1 + n].reshape(-1)


# ------------------------------------------------------------------ evaluation (train.py:440-457)
@torch.no_grad()
def evaluate(model, source: torch.Tensor, seq_len: int, prec: str = "bf16x3") -> float:
    """Mean per-token NLL of ``source`` (batchified, on the device) under the posterior means."""
    was_training = model.training
    model.eval()
    total = torch.zeros((), dtype=torch.float64, device=source.device)
    is_rnn = model.family.endswith("lstm")
    hidden = model.init_hidden(source.size(1)) if is_rnn else None
    for i in range(0, source.size(0) - 1, seq_len):
        data, targets = get_batch(source, i, seq_len)
        T, B = data.shape
        if is_rnn:
            nll, hidden = engine.lstm_token_nll(model, data, targets, hidden, prec=prec)
            total += nll.double().sum() / B          # len(data) * mean over T*B tokens
        else:
            cols = data.t().contiguous().to(torch.int32).view(-1)          # hypothesis-major packing
            tgt = targets.view(T, B).t().contiguous().to(torch.int32).view(-1)
            offs = torch.arange(0, (B + 1) * T, T, dtype=torch.int32, device=data.device)
            pos = torch.arange(T, dtype=torch.int32, device=data.device).repeat(B)
            nll = model.score(PackedBatch(cols, tgt, pos, offs, T, T * B, B), prec=prec, return_token_nll=True)
            total += nll.double().sum() / B
    model.train(was_training)
    return float(total) / (len(source) - 1)


# ------------------------------------------------------------------ one epoch (train.py:306-438)
def train_epoch(ft, train_data: torch.Tensor, seq_len: int, epoch: int, *, log_interval: int = 200,
                seed: int = 1111, log=print) -> float:
    """One pass over ``train_data`` with the CUDA fine-tune step; returns the mean loss (CE + scaled KL, the quantity
    train.py:422-431 accumulates and logs) of the epoch."""
    model = ft.model
    model.train()
    kl_scale = float(seq_len) / len(train_data)          # kl / len(train_data) * seq_len, train.py:338
    is_rnn = model.family.endswith("lstm")
    if model.family == "v_tm" and getattr(model, "v_pos", 0) and seq_len != 100:
        raise ValueError(f"--seq_len {seq_len}: the variational Transformer layers carry (100, 1, d) parameters and only "
                         "exist at sequence length 100 (model.py:2754-2761, 2784); the reference's KL raises here too")
    hidden = None                                        # zeros at the start of every epoch (train.py:314)
    tot = torch.zeros((), dtype=torch.float64, device=train_data.device)
    tot_kl = torch.zeros((), dtype=torch.float64, device=train_data.device)
    n_since, n_batches, t0 = 0, 0, time.time()
    for batch, i in enumerate(range(0, train_data.size(0) - 1, seq_len)):
        data, targets = get_batch(train_data, i, seq_len)
        if model.family == "v_tm" and getattr(model, "v_pos", 0) and data.size(0) != 100:
            continue     # ragged last batch: the variational layers only exist at T = 100 (model.py:2784)
        step_seed = (seed * 1000003 + epoch * 100003 + batch) & 0x7FFFFFFFFFFF
        if is_rnn:
            loss, _, kl = ft.step(data, targets, kl_scale, seed=step_seed, hidden=hidden)
            hidden = ft.hidden                           # detached by construction (train.py:320)
        elif data.size(0) == seq_len and _USE_GRAPHS:
            # full-size batches replay the step as two CUDA graphs (~250 small launches otherwise); the graphs bake
            # in the shapes, the KL scale and the learning rate, so they are re-captured when any of them changes
            cap = getattr(ft, "_cap", None)
            key = (data.size(0), data.size(1), kl_scale, ft.lr)
            if cap is None or cap.get("key") != key:
                ft.capture(data.size(0), data.size(1), kl_scale)
                ft._cap["key"] = key
            loss, _, kl = ft.step_captured(data, targets, step_seed)
        else:
            loss, _, kl = ft.step(data, targets, kl_scale, seed=step_seed)
        tot += loss.double()
        tot_kl += kl.double() * kl_scale
        n_since += 1
        n_batches += 1
        if log_interval and batch % log_interval == 0 and batch > 0:
            cur = float(tot) / n_batches
            log(f"| epoch {epoch:3d} | {batch:5d}/{len(train_data) // seq_len:5d} batches | lr {ft.lr:02.4f} | "
                f"ms/batch {(time.time() - t0) * 1000 / n_since:5.2f} | loss {cur:5.2f} | "
                f"kl_loss {float(tot_kl) / n_batches:5.4f} | ppl {math.exp(min(cur, 50)):8.2f}")
            n_since, t0 = 0, time.time()
    if n_batches == 0:
        raise RuntimeError(f"the epoch ran no step: {train_data.size(0)} rows of training data at --seq_len {seq_len}")
    return float(tot) / n_batches


def train_steps(ft, ids: torch.Tensor, bsz: int, seq_len: int, steps: int, *, seed: int = 1111) -> List[float]:
    """``steps`` fine-tune steps over a token stream (batchified like train.py:164-176, wrapping around), through the
    captured CUDA graphs; returns the loss of every 10th step.  Used to make peaked models for the ranking evidence
    of bench.py / tests (a random-init model ranks n-best lists by length only)."""
    data = batchify(ids, bsz, ft.device)
    kl_scale = float(seq_len) / len(data)
    ft.model.train()
    ft.capture(seq_len, bsz, kl_scale)
    losses, n_rows = [], data.size(0) - 1
    for k in range(steps):
        i = (k * seq_len) % max(n_rows - seq_len, 1)
        x, y = data[i:i + seq_len], data[i + 1:i + 1 + seq_len].reshape(-1)
        loss, _, _ = ft.step_captured(x, y, seed + k)
        if k % 10 == 0 or k == steps - 1:
            losses.append(float(loss))
    return losses


# ------------------------------------------------------------------ schedule (train.py:464-512)
def fit(model, train_data, val_data, *, lr: float, epochs: int, seq_len: int, clip: float, save: str,
        prec: str = "bf16x3", log_interval: int = 200, seed: int = 1111, patience: int = 8, log=print):
    """Train with the reference schedule.  Returns (best validation loss, history) where history is a list of
    dicts (epoch, train_loss, val_loss, lr, reloaded)."""
    from .trainer import FineTuner
    ft = FineTuner(model, lr, clip=clip, prec=prec)
    best: Optional[float] = None
    counter, history = 0, []
    for epoch in range(1, epochs + 1):
        t0 = time.time()
        tr = train_epoch(ft, train_data, seq_len, epoch, log_interval=log_interval, seed=seed, log=log)
        val = evaluate(model, val_data, seq_len, prec=prec)
        log("-" * 89)
        log(f"| end of epoch {epoch:3d} | time: {time.time() - t0:5.2f}s | valid loss {val:5.2f} | "
            f"valid ppl {math.exp(min(val, 50)):8.2f}")
        log("-" * 89)
        reloaded = False
        if best is None or val < best:
            torch.save({k: v.detach().cpu() for k, v in model.state_dict().items()}, save)
            best = val
        else:
            # lr /= 2, fresh optimiser (momentum buffers start from zero), best weights back (train.py:500-506)
            ft.lr /= 2.0
            ft.flat_v.zero_()
            model.load_state_dict(torch.load(save, map_location="cpu", weights_only=True))
            model.__dict__.pop("_blm_plans", None)
            counter += 1
            reloaded = True
        history.append({"epoch": epoch, "train_loss": tr, "val_loss": val, "lr": ft.lr, "reloaded": reloaded})
        if counter == patience:
            break
    model.load_state_dict(torch.load(save, map_location="cpu", weights_only=True))
    model.__dict__.pop("_blm_plans", None)
    return best, history


_MARK_FRACTIONS = {"base-0.5set": 2, "base-0.25set": 4, "base-0.1set": 10, "base-0.05set": 20}


def pruned_length(train_len: int, mark: str) -> int:
    """``--mark``: the data-size experiments train on the leading 1/2, 1/4, 1/10 or 1/20 of the corpus (train.py:151-165)."""
    return int(train_len / _MARK_FRACTIONS.get(mark, 1))


def build_parser() -> argparse.ArgumentParser:
    p = argparse.ArgumentParser(description="Fine-tune a Bayesian / GP / Variational neural LM on B200.")
    p.add_argument("--data", type=str, default="./data/pytorchnn")
    p.add_argument("--model", type=str, default="LSTM")
    p.add_argument("--emsize", type=int, default=200)
    p.add_argument("--nhid", type=int, default=200)
    p.add_argument("--nlayers", type=int, default=2)
    p.add_argument("--nhead", type=int, default=2)
    p.add_argument("--uncertainty", type=str, default="none")
    p.add_argument("--T_bayes_pos", type=str, default="none")
    p.add_argument("--L_bayes_pos", type=int, default=0)
    p.add_argument("--L_gauss_pos", type=str, default="00")
    p.add_argument("--L_v_pos", type=str, default="11")
    p.add_argument("--T_gauss_pos", type=int, default=3)
    p.add_argument("--T_v_pos", type=str, default="0", help="0..3, or the per-layer bit string 00/01/10/11")
    p.add_argument("--mark", type=str, default="none",
                   help="base-0.5set / -0.25set / -0.1set / -0.05set train on that leading fraction (train.py:151-165)")
    p.add_argument("--lr", type=float, default=0.1)
    p.add_argument("--batch-size", type=int, default=20)
    p.add_argument("--epochs", type=int, default=20)
    p.add_argument("--seq_len", type=int, default=35)
    p.add_argument("--clip", type=float, default=0.25)
    p.add_argument("--dropout", type=float, default=0.2, help="dropout applied to layers (train.py:75)")
    p.add_argument("--tied", action="store_true")
    p.add_argument("--optimizer", type=str, default="SGD", help="parsed like the reference, which builds SGD regardless "
                                                                 "(train.py:466)")
    p.add_argument("--log-interval", type=int, default=200)
    p.add_argument("--cuda", action="store_true", help="accepted for compatibility: there is no CPU path")
    p.add_argument("--save", type=str, default="model.pt")
    p.add_argument("--seed", type=int, default=1111)
    p.add_argument("--resume", type=str, default="", help="checkpoint to start from (the fine-tuning recipe, README)")
    p.add_argument("--debug", action="store_true", help="parsed for compatibility (unused by the reference, train.py:94)")
    p.add_argument("--work_dir", type=str, default="TFM", help="parsed for compatibility (unused, train.py:96)")
    p.add_argument("--prior", type=str, default="False", help='"True": start from <prior_path>/model.pt (train.py:239-258)')
    p.add_argument("--prior_path", type=str, default="steps/pytorchnn/prior")
    p.add_argument("--prior2_path", type=str, default="steps/pytorchnn/prior/transformer2/",
                   help="parsed for compatibility (unused, train.py:102)")
    p.add_argument("--precision", type=str, default="bf16x3", choices=["bf16", "bf16x3"])
    return p


def main(argv=None) -> int:
    from . import model as models
    from .scorer import load_checkpoint
    args = build_parser().parse_args(argv)
    torch.manual_seed(args.seed)
    dev = torch.device("cuda", int(os.environ.get("LOCAL_RANK", 0)))
    torch.cuda.set_device(dev)
    corpus = Corpus(args.data)
    print("train set:", len(corpus.train), "\nvalid set:", len(corpus.valid), "\ntest set:", len(corpus.test),
          "\nnum tokens:", len(corpus))
    train_data = batchify(corpus.train[:pruned_length(len(corpus.train), args.mark)], args.batch_size, dev)
    val_data, test_data = batchify(corpus.valid, 20, dev), batchify(corpus.test, 20, dev)   # eval_batch_size = 20 (:178)
    net = models.build_model(args, len(corpus))
    if args.prior == "True":
        # pre-trained prior: keep the checkpoint tensors the model has, everything else stays as initialised
        # (a baseline model's means under a Bayesian model's fresh log-sigmas; train.py:239-258)
        left = load_checkpoint(net, os.path.join(args.prior_path, "model.pt"), partial=True)
        print(f"prior: {len(left['missing'])} model tensors keep their initialisation, "
              f"{len(left['unexpected'])} checkpoint tensors unused")
    if args.resume:
        load_checkpoint(net, args.resume)
    net = net.to(dev)
    print("Model total parameters:", sum(p.numel() for p in net.parameters()))
    best, _ = fit(net, train_data, val_data, lr=args.lr, epochs=args.epochs, seq_len=args.seq_len, clip=args.clip,
                  save=args.save, prec=args.precision, log_interval=args.log_interval, seed=args.seed)
    test = evaluate(net, test_data, args.seq_len, prec=args.precision)
    print("=" * 89)
    print(f"| End of training | test loss {test:5.2f} | test ppl {math.exp(min(test, 50)):8.2f}")
    print("=" * 89)
    return 0


if __name__ == "__main__":
    raise SystemExit(main())
