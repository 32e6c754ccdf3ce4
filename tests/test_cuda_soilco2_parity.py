"""GPU parity of SoilCO2Model's implicit diffusion (clb_soilco2_*; Biogeochemistry.jl:320-413, 1119-1195)
against the CPU oracle on identical seeded inputs, through the C ABI.  1e-12 norm-wise relative per call."""
import numpy as np
import pytest

from helpers import assert_close

pytestmark = pytest.mark.gpu
TOL = 1e-12


def _setup(ncol, N, seed, atm, layout):
    import oracle as orc
    import climaland_b200 as cl
    from climaland_b200 import workloads
    rng = np.random.default_rng(seed)
    z_f, z_c = workloads.stretched_grid(N, depth=10.0, dz_top=0.05)
    P = orc.Problem(model=orc.RICHARDS, z_f=z_f, z_c=z_c, ncol=ncol, nthreads=4, nu=0.5, theta_r=0.1, K_sat=1e-6,
                    S_s=1e-3, hcm_a=2.0, hcm_b=2.0, hcm_m=0.5)
    s = cl.SoilColumnSolver(model=cl.RICHARDS, n_columns=ncol, z_f=z_f, z_c=z_c, layout=layout)
    sp, state = {}, {}
    for name in ("co2", "o2"):
        D = rng.uniform(1e-8, 2e-6, (ncol, N))
        th = rng.uniform(0.02, 0.45, (ncol, N))
        c_atm = rng.uniform(1e-4, 4e-4, ncol) if atm else None
        C = rng.uniform(5e-5, 2e-3, (ncol, N))
        C[rng.random((ncol, N)) < 0.02] = -1e-6          # the max(C, 0) clip is exercised
        top = np.zeros(ncol) if atm else rng.normal(0, 1e-9, ncol)
        bot = rng.normal(0, 1e-10, ncol)
        sp[name] = P.co2_species(D, th, c_atm)
        state[name] = dict(C=C, top=top, bot=bot)
        s.set(f"{name}_y", C)
        s.set(f"{name}_d", D)
        s.set(f"{name}_theta_eff", th)
        s.set(f"{name}_bot_bc", bot)
        if atm:
            s.set(f"{name}_c_atm", c_atm)
        else:
            s.set(f"{name}_top_bc", top)
    s.set_co2_top_state(co2=atm, o2=atm)
    return P, s, sp, state


@pytest.mark.parametrize("layout", [1, 2], ids=["colfast", "levfast"])
@pytest.mark.parametrize("atm", [False, True], ids=["fluxbc", "atmos_state_bc"])
@pytest.mark.parametrize("N,ncol", [(15, 1000), (50, 130), (7, 33)])
def test_hooks(N, ncol, atm, layout):
    P, s, sp, st = _setup(ncol, N, 4, atm, layout)
    dtg = 900.0
    s.soilco2_update_boundary_fluxes()
    s.soilco2_compute_imp_tendency()
    s.soilco2_compute_jacobian(dtg)
    for name in ("co2", "o2"):
        S, x = sp[name], st[name]
        top, dfl = x["top"].copy(), np.zeros(ncol)
        P.co2_boundary_flux(S, x["C"], top, dfl)
        if atm:
            assert_close(s.get(f"{name}_top_bc"), top, TOL, "top_bc")
            assert_close(s.get(f"{name}_dfluxbcdy"), dfl, TOL, "dfluxBCdY")
        assert_close(s.get(f"{name}_dy"), P.co2_imp_tendency(S, x["C"], top, x["bot"]), TOL, "tendency")
        lo, di, up = P.co2_jacobian(S, dtg, dfl)
        assert_close(s.get(f"{name}_w_lo"), lo, TOL, "lower")
        assert_close(s.get(f"{name}_w_di"), di, TOL, "diagonal")
        assert_close(s.get(f"{name}_w_up"), up, TOL, "upper")
    s.close()


@pytest.mark.parametrize("atm", [False, True], ids=["fluxbc", "atmos_state_bc"])
@pytest.mark.parametrize("N,ncol,iters", [(15, 2000, 3), (50, 100, 2), (64, 17, 1)])
def test_fused_stage(N, ncol, iters, atm):
    P, s, sp, st = _setup(ncol, N, 5, atm, 0)
    dtg = 1800.0
    s.soilco2_implicit_step(dtg, iters)
    for name in ("co2", "o2"):
        S, x = sp[name], st[name]
        C, top = x["C"].copy(), x["top"].copy()
        P.co2_implicit_step(S, C, top, x["bot"], dtg, iters)
        assert_close(s.get(f"{name}_y"), C, TOL, f"{name} after the stage")
        if atm:
            assert_close(s.get(f"{name}_top_bc"), top, TOL, "top_bc left in the cache")
        assert np.max(np.abs(s.get(f"{name}_y") - x["C"])) > 0
    s.close()


def test_errors():
    import climaland_b200 as cl
    from climaland_b200 import workloads
    z_f, z_c = workloads.stretched_grid(15)
    s = cl.SoilColumnSolver(model=cl.RICHARDS, n_columns=8, z_f=z_f, z_c=z_c)
    with pytest.raises(cl.ClbError, match="never set"):
        s.soilco2_implicit_step(900.0, 3)
    s.close()
