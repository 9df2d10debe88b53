"""Aggregate an `ncu --page source --csv` dump: opcode mix and hot regions per warp."""
import csv, collections, sys
path, nwarps = sys.argv[1], float(sys.argv[2])
rows = list(csv.reader(open(path)))
hdr = rows[1]; ix = {h: i for i, h in enumerate(hdr)}
ops = collections.Counter(); tot = 0; seq = []
for r in rows[2:]:
    if len(r) < 10 or not r[ix['Instructions Executed']].isdigit():
        if seq: break   # only the first kernel instance
        continue
    src = r[ix['Source']].strip(); n = int(r[ix['Instructions Executed']])
    toks = src.split()
    op = toks[0] if not toks[0].startswith('@') else toks[1]
    ops[op.split('.')[0]] += n; tot += n
    seq.append((src, n, int(r[ix['# Samples']])))
print('total warp-instr', tot, 'per warp', round(tot / nwarps, 1), 'static', len(seq))
for op, n in ops.most_common(30):
    print(f"  {op:12s} {n / nwarps:9.1f} per warp  {100 * n / tot:5.1f}%")
print('--- regions')
start = 0
def flush(a, b):
    dyn = sum(n for _, n, _ in seq[a:b]) / nwarps; smp = sum(s for _, _, s in seq[a:b])
    print(f"  [{a:4d},{b:4d}) exec/warp of first {seq[a][1] / nwarps:7.2f}  dyn {dyn:8.1f}  samples {smp:6d}   {seq[a][0][:60]}")
for i in range(1, len(seq)):
    c0, c1 = seq[i - 1][1], seq[i][1]
    if abs(c1 - c0) > 0.15 * max(c0, c1, 1):
        flush(start, i); start = i
flush(start, len(seq))
