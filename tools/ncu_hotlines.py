"""Per CUDA source line: share of executed warp instructions and of warp-stall samples, per kernel, from an .ncu-rep captured with
--import-source on.   python tools/ncu_hotlines.py report.ncu-rep [top]"""
import csv, subprocess, sys
raw = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
top = int(sys.argv[2]) if len(sys.argv) > 2 else 16
cur_file = cur_fn = hdr = None
agg, order = {}, []
def toi(x):
    try:
        return int(x)
    except ValueError:
        return 0
for r in csv.reader(raw.splitlines()):
    if not r:
        continue
    if r[0] == "File Path":
        cur_file = r[1].split("/")[-1]
    elif r[0] == "Function Name":
        cur_fn = r[1].split("(")[0].split("::")[-1]
        if cur_fn not in order:
            order.append(cur_fn)
    elif r[0] == "Line No":
        hdr = r
    elif r[0] != "" and hdr is not None:
        try:
            line = int(r[0])
        except ValueError:
            continue
        a = agg.setdefault((cur_fn, cur_file, line), [0, 0, r[1].strip()[:120]])
        a[0] += toi(r[hdr.index("Instructions Executed")])
        a[1] += toi(r[hdr.index("Warp Stall Sampling (All Samples)")])
for fn in order:
    items = [(k, v) for k, v in agg.items() if k[0] == fn]
    ti, ts = sum(v[0] for _, v in items) or 1, sum(v[1] for _, v in items) or 1
    print("=====", fn, "warp instructions", ti, "stall samples", ts)
    for k, v in sorted(items, key=lambda kv: -kv[1][1])[:top]:
        print(f"{k[1][:20]:20s} L{k[2]:5d} inst {100 * v[0] / ti:5.1f}% stall {100 * v[1] / ts:5.1f}%  {v[2]}")
    print()
