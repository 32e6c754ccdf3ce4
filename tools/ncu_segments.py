"""Stall samples of an ncu report's SASS page aggregated over consecutive instruction segments:
python tools/ncu_segments.py file.ncu-rep [segment-length]   (phases of a long straight-line kernel)"""
import csv, io, re, subprocess, sys
out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "source", "--csv", "--print-source", "sass"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
for i, r in enumerate(rows):
    if 'Source' in r and '# Samples' in r:
        hdr, start = r, i + 1
        break
iS, iE, iSm = hdr.index('Source'), hdr.index('Instructions Executed'), hdr.index('# Samples')
body = rows[start:]
stall_cols = [c for c in hdr if c.startswith('stall_') and 'Not Issued' not in c]
tot = sum(int(r[iSm] or 0) for r in body)
nw = int(body[0][iE])
print('total samples', tot, 'static instructions', len(body), 'warps', nw)
seg = int(sys.argv[2]) if len(sys.argv) > 2 else 250
for a in range(0, len(body), seg):
    chunk = body[a:a + seg]
    s = sum(int(r[iSm] or 0) for r in chunk)
    ex = sum(int(r[iE] or 0) for r in chunk)
    ops = {}
    for r in chunk:
        m = re.match(r"\s*(@!?U?P\d+\s+)?([A-Z0-9_]+)", r[iS])
        if m:
            ops[m.group(2)] = ops.get(m.group(2), 0) + 1
    st = {c: sum(int(r[hdr.index(c)] or 0) for r in chunk) for c in stall_cols}
    top = sorted(st.items(), key=lambda x: -x[1])[:4]
    mem = {k: ops.get(k, 0) for k in ('LDG', 'LDGSTS', 'LDS', 'STS', 'STG', 'SHFL', 'BRA', 'MUFU')}
    print(f"{a:5d} samples {s:5d} ({100*s/tot:4.1f}%) exec/warp {ex/nw:7.1f} {mem} {top}")
