"""CPU-side checks: the C-ABI library builds, loads and exports every symbol the header
declares; the ctypes mirror of clb_config matches the C struct; without a CUDA device the
product path fails loudly (there is no CPU fallback); synthetic workloads are deterministic."""
import ctypes as C
import os
import re
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "climaland_b200.h")


def _cl():
    import climaland_b200
    return climaland_b200


def test_library_exports_every_declared_symbol():
    cl = _cl()
    L = cl._lib.lib()
    src = re.sub(r"/\*.*?\*/", "", open(HEADER).read(), flags=re.S)
    declared = sorted(set(re.findall(r"\b(clb_[a-z0-9_]+)\s*\(", src)))
    assert len(declared) >= 24
    for name in declared:
        assert hasattr(L, name), f"{name} is declared in include/climaland_b200.h but not exported"
    assert L.clb_abi_version() == cl._lib.K["CLB_ABI_VERSION"]


def test_ctypes_config_matches_c_struct(tmp_path):
    cl = _cl()
    prog = tmp_path / "layout.c"
    prog.write_text('#include <stdio.h>\n#include <stddef.h>\n#include "climaland_b200.h"\n'
                    'int main(void){printf("%zu %zu %zu %zu %zu %zu %d %d\\n", sizeof(clb_config), offsetof(clb_config, n_columns),'
                    ' offsetof(clb_config, stream), offsetof(clb_config, rho_l), offsetof(clb_config, layout),'
                    ' sizeof(clb_stats), (int)CLB_F_NUM_CELL, (int)CLB_F_NUM); return 0;}\n')
    exe = tmp_path / "layout"
    subprocess.run(["gcc", "-I", os.path.join(ROOT, "include"), "-o", str(exe), str(prog)], check=True)
    vals = [int(v) for v in subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.split()]
    Cfg, St = cl._lib.Config, cl._lib.Stats
    assert vals[0] == C.sizeof(Cfg)
    assert vals[1] == Cfg.n_columns.offset and vals[2] == Cfg.stream.offset
    assert vals[3] == Cfg.rho_l.offset and vals[4] == Cfg.layout.offset
    assert vals[5] == C.sizeof(St)
    assert vals[6] == cl._lib.K["CLB_F_NUM_CELL"] and vals[7] == cl._lib.K["CLB_F_NUM"]


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    cl = _cl()
    with pytest.raises(cl.ClbError) as e:
        cl.SoilColumnSolver(model=cl.RICHARDS, n_columns=4, z_f=[-1.0, -0.5, 0.0])
    assert e.value.code == cl._lib.K["CLB_ERR_NO_DEVICE"]
    assert "no CPU path" in str(e.value)


def test_product_package_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "climaland.jl_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert "soil_oracle" not in text and "import oracle" not in text, f"{f} references the oracle"


def test_workloads_are_deterministic_and_shaped():
    _cl()
    from climaland_b200 import workloads as wl
    a = wl.make_workload("energy_hydrology", 64, N=15, seed=3, topmodel=True)
    b = wl.make_workload("energy_hydrology", 64, N=15, seed=3, topmodel=True)
    for k, v in a.items():
        if isinstance(v, np.ndarray):
            assert np.array_equal(v, b[k]), k
    assert a["y_theta_l"].shape == (64, 15) and a["top_bc_w"].shape == (64,)
    assert np.all(a["z_f"][1:] > a["z_f"][:-1]) and a["z_f"][0] == -50.0 and a["z_f"][-1] == 0.0
    assert abs((a["z_f"][-1] - a["z_f"][-2]) - 0.05) < 0.02      # top cell ~ dz_tuple[1] (Domains.jl:1296-1298)
    assert np.all(a["y_theta_l"] > a["theta_r"]) and np.all(a["nu"] - a["y_theta_i"] > a["theta_r"])
    # algorithmic bytes per column-step quoted in BASELINE.md section 2
    assert wl.algorithmic_bytes("richards", 15, topmodel=True) == 8 * (9 * 15 + 6) + 8 * 15  # + is_saturated field
    assert wl.algorithmic_bytes("energy_hydrology", 15, topmodel=True) == 2008
    assert wl.algorithmic_bytes("energy_hydrology", 50, topmodel=True) == 6488


def test_shard_ranges_partition_the_columns():
    _cl()
    from climaland_b200 import parallel
    for n in (1, 7, 61206, 64800):
        for world in (1, 2, 4, 8):
            edges = [parallel.shard_range(n, world, r) for r in range(world)]
            assert edges[0][0] == 0 and edges[-1][1] == n
            assert all(edges[r][1] == edges[r + 1][0] for r in range(world - 1))
            sizes = [hi - lo for lo, hi in edges]
            assert max(sizes) - min(sizes) <= 1


def test_julia_binding_matches_header():
    """julia/ClimaLandB200.jl (the reference-side ccall binding) cannot run here (no Julia): keep its
    field enum, ABI version and every ccall'ed symbol in step with include/climaland_b200.h."""
    import re
    import climaland_b200 as cl
    src = open(os.path.join(ROOT, "julia", "ClimaLandB200.jl")).read()
    K = cl._lib.K
    assert int(re.search(r"const CLB_ABI_VERSION = Int32\((\d+)\)", src).group(1)) == K["CLB_ABI_VERSION"]
    body = re.search(r"@enum ClbField::Int32 begin(.*?)\nend", src, flags=re.S).group(1)
    names = [t.split("=")[0].strip() for t in re.split(r"[;\n]", body) if t.strip()]
    header = sorted((v, k) for k, v in K.items() if k.startswith("CLB_F_") and k not in ("CLB_F_NUM", "CLB_F_NUM_CELL"))
    assert ["CLB_" + n for n in names] == [k for _, k in header]
    for sym in set(re.findall(r"ccall\(\(:(clb_\w+)", src)):
        assert sym in cl._lib.EXPORTS, sym


if __name__ == "__main__":
    sys.exit(pytest.main([__file__, "-q"]))
