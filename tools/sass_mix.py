"""Static SASS instruction mix of one kernel: python tools/sass_mix.py <substring of mangled name> [--loop]"""
import re, subprocess, sys, collections
import os
so = os.environ.get("SASS_SO", "climaland.jl_b200/libclimaland_b200.so")
out = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
blocks = re.split(r"\n\s*Function : ", out)
for b in blocks[1:]:
    name = b.split("\n", 1)[0]
    if sys.argv[1] in name:
        ops = collections.Counter()
        n = 0
        for line in b.split("\n"):
            m = re.match(r"\s+/\*[0-9a-f]{4,5}\*/\s+(@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
            if m:
                ops[m.group(2)] += 1
                n += 1
        print(name, "total", n)
        fp64 = sum(v for k, v in ops.items() if k in ("DFMA", "DMUL", "DADD", "DSETP"))
        print("  FP64-pipe:", fp64, " ".join(f"{k}={v}" for k, v in ops.most_common(24)))
