"""Stall samples and executed instructions of an ncu report aggregated per CUDA source line
(needs -lineinfo and --import-source on): python tools/ncu_lines.py file.ncu-rep [top-n]"""
import sys
sys.path.insert(0, "/opt/nvidia/nsight-compute/2025.2.1/extras/python")
import ncu_report

ctx = ncu_report.load_report(sys.argv[1])
act = ctx.range_by_idx(0).action_by_idx(0)
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
samples = act.metric_by_name("smsp__pcsamp_sample_buffer") if False else None
m_s = act.metric_by_name("smsp__pcsamp_warps_issue_stalled_all") if "smsp__pcsamp_warps_issue_stalled_all" in act.metric_names() else None
names = [n for n in act.metric_names() if n.startswith("smsp__pcsamp_warps_issue_stalled_") and not n.endswith("_not_issued")]
inst = act.metric_by_name("inst_executed")
per_line = {}
tot = 0
for n in names:
    m = act.metric_by_name(n)
    reason = n.replace("smsp__pcsamp_warps_issue_stalled_", "")
    for i in range(m.num_instances()):
        pc = m.correlation_ids().as_uint64(i)
        v = m.as_uint64(i)
        if not v:
            continue
        info = act.source_info(pc)
        key = (info.file_name().split("/")[-1], info.line()) if info else ("?", 0)
        d = per_line.setdefault(key, {"n": 0, "inst": 0})
        d["n"] += v
        d[reason] = d.get(reason, 0) + v
        tot += v
for i in range(inst.num_instances()):
    pc = inst.correlation_ids().as_uint64(i)
    info = act.source_info(pc)
    key = (info.file_name().split("/")[-1], info.line()) if info else ("?", 0)
    per_line.setdefault(key, {"n": 0, "inst": 0})["inst"] += inst.as_uint64(i)
print(act.name(), "total samples", tot)
for key, d in sorted(per_line.items(), key=lambda kv: -kv[1]["n"])[:top]:
    rs = sorted(((k, v) for k, v in d.items() if k not in ("n", "inst")), key=lambda kv: -kv[1])[:3]
    print(f"{key[0]:22s}:{key[1]:4d} {d['n']:6d} ({100 * d['n'] / tot:4.1f}%) inst {d['inst']:9d}  " + " ".join(f"{k}={v}" for k, v in rs))

# coarse phases of soil_pair.cuh (line ranges of k_step_lanes) and of the math headers
PH = [("pair: tile request / wait / scalars", 282, 353), ("pair: set-up (constants, W22 factor)", 354, 469),
      ("pair: closures + T (loop head)", 470, 531), ("pair: neighbour exchange", 532, 540),
      ("pair: faces, residual, W11 rows", 541, 593), ("pair: W11 twisted Thomas", 594, 626),
      ("pair: update + W21 x1", 627, 652), ("pair: W22 solve", 653, 682), ("pair: store", 683, 720)]
agg = {}
for (f, ln), d in per_line.items():
    name = f
    if f == "soil_pair.cuh":
        name = next((p for p, a, b in PH if a <= ln <= b), "pair: other")
    a = agg.setdefault(name, {"n": 0, "inst": 0})
    a["n"] += d["n"]
    a["inst"] += d["inst"]
print()
for name, a in sorted(agg.items(), key=lambda kv: -kv[1]["n"]):
    print(f"{name:42s} samples {a['n']:6d} ({100 * a['n'] / tot:4.1f}%)  inst {a['inst']:10d}")
