# julia/parity_check.jl -- closes the "parity unpinned" rows (the linear solve and the Newton / ARS111 stage, DESIGN.md
# section 5) against the REAL reference on a box that has Julia >= 1.10, a CUDA GPU and the reference's pinned
# Manifest (ClimaCore 0.15.2, ClimaTimeSteppers 0.10.6: .buildkite/Manifest.toml:510-514, 562-566).
#
# STATUS: NOT executed -- the build image has no Julia.  Run from a ClimaLand.jl v1.11.2 checkout:
#
#     CLIMALAND_B200_LIB=/path/to/libclimaland_b200.so \
#     julia --project=.buildkite /path/to/repo/julia/parity_check.jl [richards|energy_hydrology] [fused]
#
# What it does, in the order the rows are pinned:
#   1. per call  -- one update_implicit_cache! / compute_imp_tendency! / compute_jacobian! / ldiv! of the stock model and
#                   of the B200 hooks on the SAME Y, p: max relative difference of p.soil.psi, dY.soil.theta_l, the
#                   Jacobian rows (through clb_get_field of CLB_F_W11_*) and the ldiv! output, tolerance 1e-12
#                   (call shape: test/integrated/full_land.jl:586-637);
#   2. per stage -- one step! of both simulations from the same state: theta_l (rho_e_int) after the stage, 1e-12;
#   3. one day   -- solve! of both to t0 + 86400 s: prognostic state 1e-9 relative (BASELINE.json north_star).
# The set-up is the reference's own single-column Richards / soil-only EnergyHydrology benchmark configuration
# (experiments/benchmarks/richards.jl, experiments/benchmarks/soil.jl:47-73) on a small sphere.
import ClimaComms
ClimaComms.@import_required_backends
using ClimaCore, ClimaLand, ClimaTimeSteppers, LinearAlgebra, Dates
import ClimaLand.Parameters as LP
import ClimaParams
include(joinpath(@__DIR__, "ClimaLandB200.jl"))
using .ClimaLandB200

const FT = Float64
const which = length(ARGS) >= 1 ? ARGS[1] : "richards"
const fused = "fused" in ARGS

relerr(a, b) = maximum(abs, Array(parent(a)) .- Array(parent(b))) / max(maximum(abs, Array(parent(b))), floatmin(FT))

function make_model()
    toml = LP.create_toml_dict(FT)
    domain = ClimaLand.Domains.SphericalShell(; radius = FT(6.3781e6), depth = FT(50), nelements = (10, 15),
                                              dz_tuple = FT.((10.0, 0.05)))
    vg = ClimaLand.Soil.vanGenuchten{FT}(; α = FT(2.6), n = FT(2))
    if which == "richards"
        params = ClimaLand.Soil.RichardsParameters(; ν = FT(0.495), hydrology_cm = vg, K_sat = FT(0.0443 / 3600 / 100),
                                                   S_s = FT(1e-3), θ_r = FT(0.124))
        bcs = (; top = ClimaLand.Soil.WaterFluxBC((p, t) -> -1e-7), bottom = ClimaLand.Soil.FreeDrainage())
        model = ClimaLand.Soil.RichardsModel{FT}(; parameters = params, domain, boundary_conditions = bcs, sources = ())
        ic! = (Y, p, t, m) -> (Y.soil.ϑ_l .= FT(0.24))
    else
        params = ClimaLand.Soil.EnergyHydrologyParameters(toml; ν = FT(0.495), ν_ss_om = FT(0.1), ν_ss_quartz = FT(0.4),
            ν_ss_gravel = FT(0.0), hydrology_cm = vg, K_sat = FT(0.0443 / 3600 / 100), S_s = FT(1e-3), θ_r = FT(0.124))
        zero_w = ClimaLand.Soil.WaterFluxBC((p, t) -> -1e-7); zero_h = ClimaLand.Soil.HeatFluxBC((p, t) -> 0.0)
        bcs = (; top = ClimaLand.Soil.WaterHeatBC(; water = zero_w, heat = zero_h),
                 bottom = ClimaLand.Soil.WaterHeatBC(; water = ClimaLand.Soil.FreeDrainage(), heat = zero_h))
        model = ClimaLand.Soil.EnergyHydrology{FT}(; parameters = params, domain, boundary_conditions = bcs,
                                                   sources = (ClimaLand.Soil.PhaseChange{FT}(),))
        ic! = function (Y, p, t, m)
            Y.soil.ϑ_l .= FT(0.24); Y.soil.θ_i .= FT(0)
            ρc = ClimaLand.Soil.volumetric_heat_capacity.(Y.soil.ϑ_l, Y.soil.θ_i, m.parameters.ρc_ds, m.parameters.earth_param_set)
            Y.soil.ρe_int .= ClimaLand.Soil.volumetric_internal_energy.(Y.soil.θ_i, ρc, FT(285), m.parameters.earth_param_set)
        end
    end
    return model, ic!
end

model, ic! = make_model()
t0, Δt = 0.0, which == "richards" ? 1800.0 : 900.0
mk(tf) = (ClimaLand.Simulations.LandSimulation(t0, tf, Δt, model; set_ic! = ic!, user_callbacks = (), diagnostics = ()),
          ClimaLandB200.LandSimulationB200(t0, tf, Δt, model; fused, set_ic! = ic!, user_callbacks = (), diagnostics = ()))

# ---- 1. per call --------------------------------------------------------------------------------------------
let (ref, new) = mk(t0 + Δt)
    Y, p = ref._integrator.u, ref._integrator.p
    b = ClimaLandB200.B200Soil(model, Y, p)
    dtγ = Δt
    # stock hooks
    cache_ref! = ClimaLand.make_update_implicit_cache(model); imp_ref! = ClimaLand.make_compute_imp_tendency(model)
    jac_ref! = ClimaLand.make_compute_jacobian(model)
    p_ref, dY_ref, W_ref = deepcopy(p), similar(Y), ClimaLand.initialize_jacobian(Y)
    cache_ref!(p_ref, Y, t0); imp_ref!(dY_ref, Y, p_ref, t0); jac_ref!(W_ref, Y, p_ref, dtγ, t0)
    x_ref = similar(Y); ldiv!(x_ref, W_ref, dY_ref)
    # B200 hooks
    p_new, dY_new, W_new = deepcopy(p), similar(Y), ClimaLandB200.initialize_jacobian(b)
    ClimaLandB200.make_update_implicit_cache(b)(p_new, Y, t0)
    ClimaLandB200.make_compute_imp_tendency(b)(dY_new, Y, p_new, t0)
    ClimaLandB200.make_compute_jacobian(b)(W_new, Y, p_new, dtγ, t0)
    x_new = similar(Y); ldiv!(x_new, W_new, dY_ref)   # same right-hand side for both solves
    e = Dict("psi" => relerr(p_new.soil.ψ, p_ref.soil.ψ), "dY.theta_l" => relerr(dY_new.soil.ϑ_l, dY_ref.soil.ϑ_l),
             "ldiv!.theta_l" => relerr(x_new.soil.ϑ_l, x_ref.soil.ϑ_l))
    which == "richards" || (e["ldiv!.rho_e_int"] = relerr(x_new.soil.ρe_int, x_ref.soil.ρe_int))
    @info "per call" e
    @assert all(v -> v <= 1e-12, values(e))
end

# ---- 2. one stage, 3. one day --------------------------------------------------------------------------------
for (label, tf, tol) in (("one stage", t0 + Δt, 1e-12), ("one day", t0 + 86400.0, 1e-9))
    ref, new = mk(tf)
    ClimaLand.Simulations.solve!(ref); ClimaLand.Simulations.solve!(new)
    e = Dict("theta_l" => relerr(new._integrator.u.soil.ϑ_l, ref._integrator.u.soil.ϑ_l))
    which == "richards" || (e["rho_e_int"] = relerr(new._integrator.u.soil.ρe_int, ref._integrator.u.soil.ρe_int))
    @info label e
    @assert all(v -> v <= tol, values(e))
end
println("parity_check: ", which, fused ? " (fused)" : " (fine-grained)", " OK")
