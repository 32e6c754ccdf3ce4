"""Dump the SASS of one kernel: python tools/sass_fn.py <substring of mangled name> > out.sass"""
import re, subprocess, sys
import os
so = os.environ.get("SASS_SO", "climaland.jl_b200/libclimaland_b200.so")
out = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
for b in re.split(r"\n\s*Function : ", out)[1:]:
    if sys.argv[1] in b.split("\n", 1)[0]:
        for line in b.split("\n"):
            m = re.match(r"\s+/\*([0-9a-f]{4,5})\*/\s+(.*?);", line)
            if m:
                print(m.group(1), m.group(2).strip())
        break
