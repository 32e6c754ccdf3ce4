"""Issue-cost model of a kernel's SASS: 1 cycle per non-FP64 instruction, 2 per FP64 instruction,
3 per DFMA with three distinct 64-bit register operands none of which carries .reuse.
python tools/sass_cost.py <library.so> <substring of the mangled name> -- prints the cost of every loop body found
(a backward branch and its target) and of the whole function."""
import re, subprocess, sys
so, pat = sys.argv[1], sys.argv[2]
out = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
for b in re.split(r"\n\s*Function : ", out)[1:]:
    name = b.split("\n", 1)[0]
    if pat not in name:
        continue
    ins = []
    for line in b.split("\n"):
        m = re.match(r"\s+/\*([0-9a-f]{4,5})\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)\s*([^;]*);", line)
        if m:
            ins.append((int(m.group(1), 16), m.group(2), m.group(3)))
    def cost(lo, hi):
        n = c = f = d3 = 0
        for a, op, args in ins:
            if not (lo <= a <= hi):
                continue
            n += 1
            base = op.split(".")[0]
            if base in ("DFMA", "DMUL", "DADD", "DSETP"):
                f += 1
                c += 2
                if base == "DFMA":
                    srcs = [x.strip().lstrip("-|") for x in args.split(",")][1:]
                    regs = {x.split(".")[0] for x in srcs if re.match(r"R\d", x) and ".reuse" not in x}
                    if len(regs) == 3:
                        d3 += 1
                        c += 1
            else:
                c += 1
        return n, f, d3, c
    print(name[:100])
    print("  whole function: %d instructions, %d FP64 (%d three-register DFMAs), cost %d" % cost(0, 1 << 30))
    for a, op, args in ins:
        if op.startswith("BRA"):
            m = re.search(r"0x([0-9a-f]+)", args)
            if m and int(m.group(1), 16) < a and a - int(m.group(1), 16) > 0x400:
                t = int(m.group(1), 16)
                print("  loop %#x .. %#x: %d instructions, %d FP64 (%d three-register DFMAs), cost %d" % ((t, a) + cost(t, a)))
