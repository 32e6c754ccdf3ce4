"""Issue-cost model (tools/sass_cost.py) per CUDA source line: python tools/sass_cost_lines.py <disassembly> <function
substring> <lo address hex> <hi address hex> [top-n]; the disassembly is `nvdisasm --print-line-info` of the cubin
(cuobjdump -xelf all lib.so), built with -lineinfo."""
import re, sys, collections
path, pat, lo, hi = sys.argv[1], sys.argv[2], int(sys.argv[3], 16), int(sys.argv[4], 16)
top = int(sys.argv[5]) if len(sys.argv) > 5 else 50
cost = collections.Counter(); cnt = collections.Counter(); ops = collections.defaultdict(collections.Counter)
infun = False; cur = ("?", 0); total = 0
for line in open(path):
    if line.startswith(".text."):
        infun = pat in line
        continue
    if not infun:
        continue
    m = re.match(r'\s*//## File "([^"]+)", line (\d+)', line)
    if m:
        cur = (m.group(1).split("/")[-1], int(m.group(2)))
        continue
    m = re.match(r"\s+/\*([0-9a-f]{4,5})\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)\s*([^;]*);", line)
    if not m:
        continue
    a = int(m.group(1), 16)
    if not (lo <= a <= hi):
        continue
    op, args = m.group(2), m.group(3)
    base = op.split(".")[0]
    c = 1
    if base in ("DFMA", "DMUL", "DADD", "DSETP"):
        c = 2
        if base == "DFMA":
            srcs = [x.strip().lstrip("-|") for x in args.split(",")][1:]
            if len({x.split(".")[0] for x in srcs if re.match(r"R\d", x) and ".reuse" not in x}) == 3:
                c = 3
    cost[cur] += c; cnt[cur] += 1; ops[cur][base] += 1; total += c
print("total cost", total, "instructions", sum(cnt.values()))
for k, v in cost.most_common(top):
    print(f"{k[0]:20s}:{k[1]:4d}  cost {v:4d} ({100.0 * v / total:4.1f}%)  n {cnt[k]:3d}  " + " ".join(f"{o}={n}" for o, n in ops[k].most_common(5)))
