"""Dynamic instruction mix and stall samples per opcode from an ncu report's source page:
python tools/ncu_opmix.py file.ncu-rep [kernel-index]"""
import csv, subprocess, sys, io, collections, re
out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "source", "--csv", "--print-source", "sass"],
                     capture_output=True, text=True).stdout
# reports with several kernels: blocks start with "Kernel Name"
blocks = out.split('"Kernel Name"')
blk = blocks[1 + (int(sys.argv[2]) if len(sys.argv) > 2 else 0)]
rows = list(csv.reader(io.StringIO('"Kernel Name"' + blk)))
print(rows[0][1])
hdr = rows[1]
iS, iE, iSm = hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("# Samples")
ex, sm = collections.Counter(), collections.Counter()
tot = 0
for r in rows[2:]:
    if len(r) <= iE: continue
    m = re.match(r"\s*(@!?U?P\d+\s+)?([A-Z0-9_]+)", r[iS])
    if not m: continue
    op = m.group(2)
    e = int(r[iE] or 0); s = int(r[iSm] or 0)
    ex[op] += e; sm[op] += s; tot += e
nw = max(ex.values()) and rows[2][iE]
nw = int(rows[2][iE])
print(f"total executed {tot}  per warp {tot/nw:.0f}  (warps {nw})")
for op, e in ex.most_common(30):
    print(f"  {op:10s} {e/nw:8.1f} /warp  {100*e/tot:5.1f}%   samples {sm[op]}")
