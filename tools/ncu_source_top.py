"""Top stall locations of an ncu --page source --csv dump (SASS level).  usage: ncu_source_top.py rep [N]"""
import csv, subprocess, sys
rep, n = sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
lines = out.splitlines()
# first line names the kernel; possibly several kernels concatenated
blocks, cur = [], []
for l in lines:
    if l.startswith('"Kernel Name"'):
        if cur: blocks.append(cur)
        cur = [l]
    else:
        cur.append(l)
if cur: blocks.append(cur)
for b in blocks[:1]:
    print(b[0][:120])
    rows = list(csv.reader(b[1:]))
    hdr = rows[0]
    ia, isrc, isamp, iexec = hdr.index("Address"), hdr.index("Source"), hdr.index("# Samples"), hdr.index("Instructions Executed")
    stall_cols = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
    body = rows[1:]
    tot = sum(int(r[isamp] or 0) for r in body)
    print("total samples", tot, "instructions", len(body))
    agg = {}
    for r in body:
        for i in stall_cols:
            agg[hdr[i]] = agg.get(hdr[i], 0) + int(r[i] or 0)
    print({k: v for k, v in sorted(agg.items(), key=lambda kv: -kv[1]) if v})
    idx = sorted(range(len(body)), key=lambda i: -int(body[i][isamp] or 0))[:n]
    for i in sorted(idx):
        r = body[i]
        st = {hdr[c][6:]: int(r[c]) for c in stall_cols if int(r[c] or 0)}
        print(f"{i:5d} {int(r[isamp]):6d} {100*int(r[isamp])/max(tot,1):5.1f}% exec={r[iexec]:>8s} {r[isrc].strip()[:70]:70s} {st}")

if len(sys.argv) > 4:
    lo, hi = int(sys.argv[3]), int(sys.argv[4])
    sel = [r for r in body if lo <= int(r[iexec] or 0) <= hi]
    tot2 = sum(int(r[isamp] or 0) for r in sel)
    agg = {}
    ops = {}
    for r in sel:
        for i in stall_cols:
            agg[hdr[i]] = agg.get(hdr[i], 0) + int(r[i] or 0)
        op = r[isrc].strip().split()[0]
        if op.startswith("@"): op = r[isrc].strip().split()[1]
        o = ops.setdefault(op.split(".")[0], [0, 0]); o[0] += 1; o[1] += int(r[isamp] or 0)
    print(f"instructions with exec in [{lo},{hi}]: {len(sel)}, samples {tot2} ({100*tot2/max(tot,1):.1f}%)")
    print({k: v for k, v in sorted(agg.items(), key=lambda kv: -kv[1]) if v})
    print({k: tuple(v) for k, v in sorted(ops.items(), key=lambda kv: -kv[1][1])})
