"""Aggregate the ncu source page (SASS) of a report: stall samples by opcode and by code region."""
import csv, collections, subprocess, sys
rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr = rows[1]
si, ci, ei = hdr.index('Source'), hdr.index('Warp Stall Sampling (All Samples)'), hdr.index('Instructions Executed')
stall_cols = [(i, h) for i, h in enumerate(hdr) if h.startswith('stall_')]
agg = collections.defaultdict(lambda: [0, 0, 0])
tot = 0
seq = []
for r in rows[2:]:
    if len(r) <= ci: continue
    toks = r[si].split()
    if not toks: continue
    op = toks[1] if toks[0].startswith('@') else toks[0]
    op = op.split('.')[0].rstrip(';')
    s = int(r[ci] or 0); e = int(r[ei] or 0)
    agg[op][0] += s; agg[op][1] += e; agg[op][2] += 1; tot += s
    seq.append((op, s, e, r))
print("total samples", tot)
for op, (s, e, n) in sorted(agg.items(), key=lambda x: -x[1][0])[:16]:
    print(f"{op:12s} samples {s:7d} {100*s/max(tot,1):5.1f}%  exec {e:11d}  static {n}")
# stall reason totals
tots = collections.Counter()
for op, s, e, r in seq:
    for i, h in stall_cols:
        try: tots[h] += int(r[i] or 0)
        except ValueError: pass
print({k: v for k, v in tots.most_common(10)})
# regions of 64 instructions
if "--regions" in sys.argv:
    W = 64
    for b in range(0, len(seq), W):
        blk = seq[b:b + W]
        s = sum(x[1] for x in blk)
        ops = collections.Counter(x[0] for x in blk).most_common(3)
        print(f"[{b:5d}] {100*s/max(tot,1):5.1f}% exec0={blk[0][2]:9d} {ops}")
