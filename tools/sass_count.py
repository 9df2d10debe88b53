"""static SASS instruction counts per kernel: python tools/sass_count.py [obj] [kernel-substring for opcode mix]"""
import collections, re, subprocess, sys
obj = sys.argv[1] if len(sys.argv) > 1 else "sequential_monte_carlo_b200/lib/smcb_filter.o"
pat = sys.argv[2] if len(sys.argv) > 2 else None
out = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True).stdout
name, n, mix = None, collections.Counter(), collections.defaultdict(collections.Counter)
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        name = re.sub(r"smcb::\(anonymous namespace\)::|smcb::", "", name).split("(")[0]
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(@!?U?P\d\s+)?([A-Z0-9_]+)", line)
    if m and name:
        n[name] += 1
        mix[name][m.group(2)] += 1
for k, v in sorted(n.items(), key=lambda kv: kv[1]):
    print(f"{v:6d}  {k}")
if pat:
    for k in mix:
        if pat in k:
            print("==", k, mix[k].most_common(14))
