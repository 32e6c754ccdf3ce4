# ClimaLandB200.jl -- the reference-side binding of libclimaland_b200.so.
#
# STATUS: written against ClimaLand.jl v1.11.2 / ClimaCore 0.15 / ClimaTimeSteppers 0.10 sources;
# NOT executed -- the build image has no Julia toolchain (DESIGN.md "Boundary").  The C ABI it
# binds (include/climaland_b200.h) is exercised by the same call sequence from Python/ctypes
# (climaland.jl_b200/solver.py, soil.py) in tests/.
#
# What a ClimaLand maintainer does with it (INTEGRATION.md has the walk-through):
#
#     using ClimaLand, ClimaLandB200
#     soil = ClimaLand.Soil.EnergyHydrology{FT}(...)                  # unchanged
#     sim  = ClimaLandB200.LandSimulationB200(t0, tf, Δt, soil; ...)   # swaps the implicit hooks
#     ClimaLand.Simulations.solve!(sim)
#
# Two drop-in levels (SURVEY 8b):
#   fine-grained  make_compute_imp_tendency / make_compute_jacobian / make_update_implicit_cache
#                 return closures that `ccall` one entry point each; `jac_prototype` is a
#                 B200SoilJacobian whose `ldiv!` is clb_ldiv.  ClimaTimeSteppers' own Newton loop
#                 drives them (src/simulations/Simulations.jl:177-199).
#   fused         FusedSoilNewton, a ClimaTimeSteppers.NewtonsMethod-like object whose `solve_newton!` takes dtγ, p, t
#                 from the integrator's Jacobian closure and calls clb_implicit_step once per stage
#                 (experimental until run under Julia; `fused = false` is the default).
module ClimaLandB200

using LinearAlgebra
import ClimaLand
import ClimaLand: Soil
import ClimaLand.Parameters as LP
import ClimaCore: Fields, Spaces
import ClimaTimeSteppers
import CUDA

const libclb = get(ENV, "CLIMALAND_B200_LIB", "libclimaland_b200.so")

# ---- enums of include/climaland_b200.h ------------------------------------------------------
const CLB_ABI_VERSION = Int32(3)
const CLB_RICHARDS, CLB_ENERGY_HYDROLOGY = Int32(0), Int32(1)
const CLB_VAN_GENUCHTEN, CLB_BROOKS_COREY = Int32(0), Int32(1)
const CLB_TOP_FLUX, CLB_TOP_MOISTURE_STATE = Int32(0), Int32(1)
const CLB_BOT_FLUX, CLB_BOT_FREE_DRAINAGE, CLB_BOT_MOISTURE_STATE = Int32(0), Int32(1), Int32(2)
const CLB_HOST, CLB_DEVICE = Int32(0), Int32(1)

# clb_field ids, in header order
@enum ClbField::Int32 begin
    F_NU = 0; F_THETA_R; F_K_SAT; F_S_S; F_HCM_A; F_HCM_B; F_HCM_M; F_RHO_C_DS
    F_K_LAG; F_KAPPA_LAG; F_THETA_L_LAG; F_IS_SATURATED
    F_Y_THETA_L; F_Y_RHO_E_INT; F_Y_THETA_I
    F_P_K; F_P_PSI; F_P_T
    F_DY_THETA_L; F_DY_RHO_E_INT; F_DY_THETA_I
    F_W11_LO; F_W11_DI; F_W11_UP; F_W21_LO; F_W21_DI; F_W21_UP; F_W22_LO; F_W22_DI; F_W22_UP
    F_B_THETA_L; F_B_RHO_E_INT; F_B_THETA_I; F_X_THETA_L; F_X_RHO_E_INT; F_X_THETA_I
    F_U_THETA_L; F_U_RHO_E_INT
    F_KAPPA_DRY; F_KAPPA_SAT_UNFROZEN; F_KAPPA_SAT_FROZEN; F_NU_SS_OM; F_NU_SS_QUARTZ; F_NU_SS_GRAVEL
    F_P_TF_DEPRESSED; F_DYE_THETA_L; F_DYE_THETA_I
    F_CO2_Y; F_O2_Y; F_CO2_D; F_O2_D; F_CO2_THETA_EFF; F_O2_THETA_EFF; F_CO2_DY; F_O2_DY
    F_CO2_W_LO; F_CO2_W_DI; F_CO2_W_UP; F_O2_W_LO; F_O2_W_DI; F_O2_W_UP
    F_CO2_B; F_O2_B; F_CO2_X; F_O2_X
    F_R_SS; F_R_ESS; F_H_GRAD; F_THETA_BC_TOP; F_THETA_BC_BOT
    F_TOP_BC_W; F_BOT_BC_W; F_TOP_BC_H; F_BOT_BC_H; F_DFLUXBCDY; F_TOTAL_WATER
    F_Y_INTF_W; F_Y_INTF_E; F_DY_INTF_W; F_DY_INTF_E; F_B_INTF_W; F_B_INTF_E; F_X_INTF_W; F_X_INTF_E
    F_AREA_WEIGHT; F_U_INTF_W; F_U_INTF_E; F_TOTAL_ENERGY
    F_F_MAX; F_PRECIP; F_INFILTRATION; F_R_S
    F_CO2_TOP_BC; F_CO2_BOT_BC; F_O2_TOP_BC; F_O2_BOT_BC; F_CO2_C_ATM; F_O2_C_ATM; F_CO2_DFLUXBCDY; F_O2_DFLUXBCDY
    F_SFC_W_DI; F_SFC_B; F_SFC_X
    F_VAPOR_FLUX_LIQ; F_LHF; F_SHF; F_R_N; F_T_AIR
end

# struct clb_config (same field order and widths as the header)
struct ClbConfig
    abi_version::Int32
    model::Int32
    closure::Int32
    top_bc::Int32
    bottom_bc::Int32
    has_topmodel_source::Int32
    n_levels::Int32
    device::Int32
    n_columns::Int64
    stream::Ptr{Cvoid}
    math_mode::Int32
    kernel_variant::Int32
    rho_l::Float64
    rho_i::Float64
    cp_l::Float64
    cp_i::Float64
    T_ref::Float64
    LH_f0::Float64
    layout::Int32
    reserved::Int32
end

# struct clb_explicit_params
struct ClbExplicitParams
    Omega::Float64
    gamma::Float64
    gammaT_ref::Float64
    alpha::Float64
    beta::Float64
    T_freeze::Float64
    grav::Float64
end

struct ClbStats
    iterations::Int32
    converged::Int32
    dx_norm::Float64
    nan_count::Int64
end

struct ClbError <: Exception
    code::Int32
    msg::String
end
Base.showerror(io::IO, e::ClbError) = print(io, "libclimaland_b200 error ", e.code, ": ", e.msg)

# The library never throws across the ABI: a negative status becomes a Julia exception here,
# which is how the reference's hooks report errors.
function check(rc::Cint)
    rc == 0 && return nothing
    throw(ClbError(rc, unsafe_string(ccall((:clb_last_error, libclb), Cstring, ()))))
end

mutable struct Handle
    ptr::Ptr{Cvoid}
    N::Int
    ncol::Int
    function Handle(cfg::ClbConfig)
        out = Ref{Ptr{Cvoid}}(C_NULL)
        check(ccall((:clb_create, libclb), Cint, (Ref{Ptr{Cvoid}}, Ref{ClbConfig}), out, cfg))
        h = new(out[], cfg.n_levels, cfg.n_columns)
        finalizer(x -> ccall((:clb_destroy, libclb), Cint, (Ptr{Cvoid},), x.ptr), h)
        return h
    end
end

# ---- field transfer -------------------------------------------------------------------------
# ClimaCore stores a scalar 3-D field as parent(field)::CuArray{FT,5} of size (Nv, Ni, Nj, 1, Nh):
# level fastest, the Ni*Nj*Nh columns follow with stride Nv (test/standalone/Soil/mask_test.jl:61).
# A 2-D (surface) field is (Ni, Nj, 1, Nh): stride 1 per column.
strides_of(f::Fields.Field) = ndims(parent(f)) == 5 ? (1, size(parent(f), 1)) : (0, 1)

function set_field!(h::Handle, id::ClbField, f::Fields.Field)
    sl, sc = strides_of(f)
    A = parent(f)
    GC.@preserve A check(ccall((:clb_set_field, libclb), Cint,
        (Ptr{Cvoid}, Int32, CUDA.CuPtr{Float64}, Int64, Int64, Int32),
        h.ptr, Int32(id), pointer(A), sl, sc, CLB_DEVICE))
end
set_field!(h::Handle, id::ClbField, v::Real) =
    check(ccall((:clb_fill_field, libclb), Cint, (Ptr{Cvoid}, Int32, Float64), h.ptr, Int32(id), Float64(v)))

function get_field!(f::Fields.Field, h::Handle, id::ClbField)
    sl, sc = strides_of(f)
    A = parent(f)
    GC.@preserve A check(ccall((:clb_get_field, libclb), Cint,
        (Ptr{Cvoid}, Int32, CUDA.CuPtr{Float64}, Int64, Int64, Int32),
        h.ptr, Int32(id), pointer(A), sl, sc, CLB_DEVICE))
end

# ---- model -> handle --------------------------------------------------------------------------
water_bc(bc) = bc isa Soil.WaterHeatBC ? bc.water : bc
closure_id(::Soil.vanGenuchten) = CLB_VAN_GENUCHTEN
closure_id(::Soil.BrooksCorey) = CLB_BROOKS_COREY
closure_id(f::Fields.Field) = closure_id(first(Array(parent(f)))) # per-cell closures share a type
top_id(bc) = water_bc(bc) isa Soil.MoistureStateBC ? CLB_TOP_MOISTURE_STATE : CLB_TOP_FLUX
bot_id(bc) = water_bc(bc) isa Soil.MoistureStateBC ? CLB_BOT_MOISTURE_STATE :
             water_bc(bc) isa Soil.FreeDrainage ? CLB_BOT_FREE_DRAINAGE : CLB_BOT_FLUX
has_topmodel(model) = any(s -> s isa ClimaLand.Soil.Runoff.TOPMODELSubsurfaceRunoff || (
    hasproperty(s, :runoff) && s.runoff isa ClimaLand.Soil.Runoff.TOPMODELRunoff), model.sources)

"""
    B200Soil(model, Y, p)

The handle plus what the hooks need to find their fields.  Built once, next to
`initialize(model)` (src/shared_utilities/models.jl:493-499); uploads the time-invariant
parameters (RichardsParameters rre.jl:22-47 / EnergyHydrologyParameters energy_hydrology.jl:60-170),
the vertical grid as ClimaCore produced it (Domains.jl:636-667) and the land-sea mask as the
active-column list (inactive columns are never read or written: mask_test.jl:53-61).
"""
struct B200Soil{M}
    model::M
    h::Handle
    energy::Bool
    runoff_set::Base.RefValue{Bool}   # TOPMODEL parameters uploaded (update_infiltration_water_flux!)
end
B200Soil(model, h::Handle, energy::Bool) = B200Soil(model, h, energy, Ref(false))

function B200Soil(model::Union{Soil.RichardsModel, Soil.EnergyHydrology}, Y, p)
    FT = eltype(Y)
    FT === Float64 || error("libclimaland_b200 is FP64 only")
    energy = model isa Soil.EnergyHydrology
    prm = model.parameters
    ϑ = Y.soil.ϑ_l
    Nv = size(parent(ϑ), 1)
    ncol_total = length(parent(ϑ)) ÷ Nv
    mask = ClimaLand.Domains.landsea_mask(ClimaLand.get_domain(model))
    active = isnothing(mask) ? nothing : Int64.(findall(>(0.5), vec(Array(parent(mask)))) .- 1)
    ncol = isnothing(active) ? ncol_total : length(active)
    earth = energy ? prm.earth_param_set : nothing
    LP = ClimaLand.Parameters
    cfg = ClbConfig(CLB_ABI_VERSION, energy ? CLB_ENERGY_HYDROLOGY : CLB_RICHARDS,
        closure_id(prm.hydrology_cm), top_id(model.boundary_conditions.top),
        bot_id(model.boundary_conditions.bottom), Int32(has_topmodel(model)), Int32(Nv),
        Int32(CUDA.deviceid(CUDA.device())), Int64(ncol), Ptr{Cvoid}(CUDA.stream().handle),
        Int32(0), Int32(0),
        energy ? LP.ρ_cloud_liq(earth) : 1000.0, energy ? LP.ρ_cloud_ice(earth) : 917.0,
        energy ? LP.cp_l(earth) : 4181.0, energy ? LP.cp_i(earth) : 2100.0,
        energy ? LP.T_0(earth) : 273.16, energy ? LP.LH_f0(earth) : 333600.0, Int32(0), Int32(0))
    h = Handle(cfg)
    # vertical grid: one column of z (cell centres) and the face heights
    z_c = Array(parent(Fields.coordinate_field(axes(ϑ)).z))[:, 1, 1, 1, 1]
    z_f = Array(parent(Fields.coordinate_field(Spaces.face_space(axes(ϑ))).z))[:, 1, 1, 1, 1]
    check(ccall((:clb_set_grid, libclb), Cint, (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}), h.ptr, z_c, z_f))
    isnothing(active) || check(ccall((:clb_set_active_columns, libclb), Cint,
        (Ptr{Cvoid}, Ptr{Int64}, Int64), h.ptr, active, length(active)))
    # parameters: scalars are broadcast, fields are transposed on the device.  Struct-valued
    # closure fields (Nf = 4) are split into component fields first, so every upload is a scalar field.
    up(id, v::Real) = set_field!(h, id, v)
    up(id, v::Fields.Field) = set_field!(h, id, v)
    up(F_NU, prm.ν); up(F_THETA_R, prm.θ_r); up(F_K_SAT, prm.K_sat); up(F_S_S, prm.S_s)
    cm = prm.hydrology_cm
    comp(name) = cm isa Fields.Field ? (@. getproperty(cm, name)) : getproperty(cm, name)
    if closure_id(cm) == CLB_VAN_GENUCHTEN
        up(F_HCM_A, comp(:α)); up(F_HCM_B, comp(:n)); up(F_HCM_M, comp(:m))
    else
        up(F_HCM_A, comp(:c)); up(F_HCM_B, comp(:ψb))
    end
    energy && up(F_RHO_C_DS, prm.ρc_ds)
    return B200Soil(model, h, energy)
end

# state / lagged cache -> library mirrors (device-to-device transposes on the handle's stream)
function push_state!(b::B200Soil, Y)
    set_field!(b.h, F_Y_THETA_L, Y.soil.ϑ_l)
    set_field!(b.h, F_Y_INTF_W, Y.soil.∫F_vol_liq_water_dt)
    if b.energy
        set_field!(b.h, F_Y_RHO_E_INT, Y.soil.ρe_int)
        set_field!(b.h, F_Y_THETA_I, Y.soil.θ_i)
        set_field!(b.h, F_Y_INTF_E, Y.soil.∫F_e_dt)
    end
end

function push_lagged!(b::B200Soil, p, t)
    h, m = b.h, b.model
    top, bot = water_bc(m.boundary_conditions.top), water_bc(m.boundary_conditions.bottom)
    if b.energy
        set_field!(h, F_TOP_BC_W, p.soil.top_bc.water); set_field!(h, F_BOT_BC_W, p.soil.bottom_bc.water)
        set_field!(h, F_TOP_BC_H, p.soil.top_bc.heat);  set_field!(h, F_BOT_BC_H, p.soil.bottom_bc.heat)
        set_field!(h, F_K_LAG, p.soil.K); set_field!(h, F_KAPPA_LAG, p.soil.κ); set_field!(h, F_THETA_L_LAG, p.soil.θ_l)
    else
        set_field!(h, F_TOP_BC_W, p.soil.top_bc); set_field!(h, F_BOT_BC_W, p.soil.bottom_bc)
    end
    top isa Soil.MoistureStateBC && set_field!(h, F_THETA_BC_TOP, top.bc(p, t))
    bot isa Soil.MoistureStateBC && set_field!(h, F_THETA_BC_BOT, bot.bc(p, t))
    if has_topmodel(m)
        set_field!(h, F_R_SS, p.soil.R_ss); set_field!(h, F_H_GRAD, p.soil.h∇)
        set_field!(h, F_IS_SATURATED, p.soil.is_saturated)
        b.energy && set_field!(h, F_R_ESS, p.soil.R_ess)
    end
end

# ---- the hooks (same names and argument order as the reference) ---------------------------------
"update_implicit_cache!(p, Y, t): src/shared_utilities/models.jl:238-246"
function make_update_implicit_cache(b::B200Soil)
    function update_implicit_cache!(p, Y, t)
        push_state!(b, Y); push_lagged!(b, p, t)
        check(ccall((:clb_update_implicit_cache, libclb), Cint, (Ptr{Cvoid},), b.h.ptr))
        get_field!(p.soil.ψ, b.h, F_P_PSI)
        if b.energy
            get_field!(p.soil.T, b.h, F_P_T)
        else
            get_field!(p.soil.K, b.h, F_P_K)
            get_field!(p.soil.total_water, b.h, F_TOTAL_WATER)
            if haskey(p.soil, :dfluxBCdY)   # rre.jl:460-468
                get_field!(p.soil.top_bc, b.h, F_TOP_BC_W); get_field!(p.soil.bottom_bc, b.h, F_BOT_BC_W)
            end
        end
        return nothing
    end
end

# ---- the explicit stage of EnergyHydrology (SURVEY 8f rank 1) ------------------------------------
"""
    set_explicit_params!(b, model)

Uploads what only `update_aux!` reads: the six thermal-conductivity / composition fields of
`EnergyHydrologyParameters` (energy_hydrology.jl:60-170) and its scalar closure constants.
Called once (the parameters are time-invariant).
"""
function set_explicit_params!(b::B200Soil, model::Soil.EnergyHydrology)
    q = model.parameters
    eps_ = q.earth_param_set
    for (id, f) in ((F_KAPPA_DRY, q.κ_dry), (F_KAPPA_SAT_UNFROZEN, q.κ_sat_unfrozen),
                    (F_KAPPA_SAT_FROZEN, q.κ_sat_frozen), (F_NU_SS_OM, q.ν_ss_om),
                    (F_NU_SS_QUARTZ, q.ν_ss_quartz), (F_NU_SS_GRAVEL, q.ν_ss_gravel))
        set_field!(b.h, id, f)
    end
    x = Ref(ClbExplicitParams(q.Ω, q.γ, q.γT_ref, q.α, q.β, LP.T_freeze(eps_), LP.grav(eps_)))
    check(ccall((:clb_set_explicit_params, libclb), Cint, (Ptr{Cvoid}, Ref{ClbExplicitParams}), b.h.ptr, x))
end

"update_aux!(p, Y, t) of EnergyHydrology: src/standalone/Soil/energy_hydrology.jl:722-814"
function make_update_aux(b::B200Soil)
    function update_aux!(p, Y, t)
        push_state!(b, Y)
        check(ccall((:clb_update_aux, libclb), Cint, (Ptr{Cvoid},), b.h.ptr))
        # the lagged inputs of the implicit stage now sit in the library's mirrors; Julia's copies of the
        # cache are refreshed for the other explicit tendencies (boundary fluxes, runoff, canopy coupling)
        get_field!(p.soil.θ_l, b.h, F_THETA_L_LAG); get_field!(p.soil.κ, b.h, F_KAPPA_LAG)
        get_field!(p.soil.K, b.h, F_K_LAG); get_field!(p.soil.T, b.h, F_P_T); get_field!(p.soil.ψ, b.h, F_P_PSI)
        get_field!(p.soil.Tf_depressed, b.h, F_P_TF_DEPRESSED)
        get_field!(p.soil.total_water, b.h, F_TOTAL_WATER); get_field!(p.soil.total_energy, b.h, F_TOTAL_ENERGY)
        return nothing
    end
end

"source!(dY, src::PhaseChange, Y, p, model): energy_hydrology.jl:846-906 (adds into dY.soil.ϑ_l, dY.soil.θ_i)"
function ClimaLand.source!(dY::Fields.FieldVector, src::Soil.PhaseChange, Y::Fields.FieldVector, p::NamedTuple,
                           b::B200Soil)
    set_field!(b.h, F_DYE_THETA_L, dY.soil.ϑ_l); set_field!(b.h, F_DYE_THETA_I, dY.soil.θ_i)
    check(ccall((:clb_phase_change_source, libclb), Cint, (Ptr{Cvoid},), b.h.ptr))
    get_field!(dY.soil.ϑ_l, b.h, F_DYE_THETA_L); get_field!(dY.soil.θ_i, b.h, F_DYE_THETA_I)
    return nothing
end

struct ClbRunoffParams
    f_over::Float64
    R_sb::Float64
    depth::Float64
end

"""
update_infiltration_water_flux!(p, runoff::TOPMODELRunoff, input, Y, t, model):
src/standalone/Soil/Runoff/Runoff.jl:234-283.  Leaves R_ss, R_ess, h∇ and is_saturated in the mirrors the
implicit TOPMODELSubsurfaceRunoff source reads, and refreshes Julia's copies.
"""
function update_infiltration_water_flux!(p, runoff::Soil.Runoff.TOPMODELRunoff, input, Y, t, b::B200Soil, depth)
    if !b.runoff_set[]
        r = Ref(ClbRunoffParams(runoff.f_over, runoff.subsurface_source.R_sb, depth))
        check(ccall((:clb_set_runoff_params, libclb), Cint, (Ptr{Cvoid}, Ref{ClbRunoffParams}), b.h.ptr, r))
        set_field!(b.h, F_F_MAX, runoff.f_max)
        b.runoff_set[] = true
    end
    set_field!(b.h, F_PRECIP, input)
    check(ccall((:clb_update_runoff, libclb), Cint, (Ptr{Cvoid},), b.h.ptr))
    get_field!(p.soil.is_saturated, b.h, F_IS_SATURATED); get_field!(p.soil.h∇, b.h, F_H_GRAD)
    get_field!(p.soil.R_ss, b.h, F_R_SS); get_field!(p.soil.infiltration, b.h, F_INFILTRATION)
    get_field!(p.soil.R_s, b.h, F_R_S)
    b.energy && get_field!(p.soil.R_ess, b.h, F_R_ESS)
    return nothing
end

const CLB_RUNOFF_NONE, CLB_RUNOFF_SURFACE, CLB_RUNOFF_TOPMODEL = Int32(0), Int32(1), Int32(2)
runoff_id(::Soil.Runoff.NoRunoff) = CLB_RUNOFF_NONE
runoff_id(::Soil.Runoff.SurfaceRunoff) = CLB_RUNOFF_SURFACE
runoff_id(::Soil.Runoff.TOPMODELRunoff) = CLB_RUNOFF_TOPMODEL

"""
    soil_boundary_fluxes!(bc::AtmosDrivenFluxBC, Val((:soil,)), model, Y, p, t, b::B200Soil, depth)

src/standalone/Soil/boundary_conditions.jl:901-936 with the runoff and the assembly of top_bc on the device:
`turbulent_fluxes!` and `net_radiation!` stay the reference's (SurfaceFluxes.jl, the radiation drivers); their results,
the liquid influx and the air temperature are uploaded, clb_update_atmos_driven_fluxes partitions the influx
(NoRunoff / SurfaceRunoff / TOPMODELRunoff) and writes top_bc.{water, heat}; Julia's copies are refreshed.
"""
function soil_boundary_fluxes!(bc::Soil.AtmosDrivenFluxBC, ::Val{(:soil,)}, model::Soil.EnergyHydrology, Y, p, t,
                               b::B200Soil, depth)
    ClimaLand.turbulent_fluxes!(p.soil.turbulent_fluxes, bc.atmos, model, Y, p, t)
    ClimaLand.net_radiation!(p.soil.R_n, bc.radiation, model, Y, p, t)
    tf = p.soil.turbulent_fluxes
    set_field!(b.h, F_PRECIP, p.drivers.P_liq); set_field!(b.h, F_T_AIR, p.drivers.T)
    set_field!(b.h, F_VAPOR_FLUX_LIQ, tf.vapor_flux_liq); set_field!(b.h, F_LHF, tf.lhf); set_field!(b.h, F_SHF, tf.shf)
    set_field!(b.h, F_R_N, p.soil.R_n)
    if bc.runoff isa Soil.Runoff.TOPMODELRunoff && !b.runoff_set[]
        r = Ref(ClbRunoffParams(bc.runoff.f_over, bc.runoff.subsurface_source.R_sb, depth))
        check(ccall((:clb_set_runoff_params, libclb), Cint, (Ptr{Cvoid}, Ref{ClbRunoffParams}), b.h.ptr, r))
        set_field!(b.h, F_F_MAX, bc.runoff.f_max)
        b.runoff_set[] = true
    end
    check(ccall((:clb_update_atmos_driven_fluxes, libclb), Cint, (Ptr{Cvoid}, Int32), b.h.ptr, runoff_id(bc.runoff)))
    get_field!(p.soil.top_bc.water, b.h, F_TOP_BC_W); get_field!(p.soil.top_bc.heat, b.h, F_TOP_BC_H)
    get_field!(p.soil.infiltration, b.h, F_INFILTRATION)
    return nothing
end

"soil_boundary_fluxes!(::EnergyWaterFreeDrainage, ::BottomBoundary, ...): boundary_conditions.jl:590-608"
function soil_boundary_fluxes!(bc::Soil.EnergyWaterFreeDrainage, ::ClimaLand.BottomBoundary, soil::Soil.EnergyHydrology,
                               Δz, Y, p, t, b::B200Soil)
    check(ccall((:clb_update_energy_water_free_drainage, libclb), Cint, (Ptr{Cvoid},), b.h.ptr))
    get_field!(p.soil.bottom_bc.water, b.h, F_BOT_BC_W); get_field!(p.soil.bottom_bc.heat, b.h, F_BOT_BC_H)
    return nothing
end

"""
    soil_step!(b, dt; max_iters = 3)

A whole EnergyHydrology soil step on the library's resident state (clb_soil_step): update_aux! + PhaseChange, the runoff
and the boundary fluxes selected with clb_set_option (CLB_OPT_RUNOFF_MODEL, _TOP_ATMOS_DRIVEN, _BOTTOM_EWFD), the explicit
update and the implicit ARS111 stage; nothing crosses the ABI but the call.
"""
soil_step!(b::B200Soil, dt; max_iters = 3) =
    check(ccall((:clb_soil_step, libclb), Cint, (Ptr{Cvoid}, Float64, Int32), b.h.ptr, Float64(dt), Int32(max_iters)))

"""
The implicit stage of SoilCO2Model (src/standalone/Soil/Biogeochemistry/Biogeochemistry.jl:320-413, 1119-1195)
as one call: uploads Y.soilco2.{CO2, O2} and the lagged p.soilco2.{D, D_o2, θ_eff, θ_eff_o2}, boundary values,
runs max_iters Newton iterations of both tridiagonals in one kernel, downloads the new CO2 / O2.
`c_atm_co2`, `c_atm_o2`: surface fields of the atmosphere's air-equivalent concentrations
(`p.drivers.c_co2 * P * M_C / (R * T)`, `O2_f_atm * P * M_O2 / (R * T)`) when the top BCs are the Atmos*StateBC.
"""
function soilco2_implicit_step!(Y, p, b::B200Soil, dtγ; max_iters = 3, c_atm_co2 = nothing, c_atm_o2 = nothing)
    set_field!(b.h, F_CO2_Y, Y.soilco2.CO2); set_field!(b.h, F_O2_Y, Y.soilco2.O2)
    set_field!(b.h, F_CO2_D, p.soilco2.D); set_field!(b.h, F_O2_D, p.soilco2.D_o2)
    set_field!(b.h, F_CO2_THETA_EFF, p.soilco2.θ_eff); set_field!(b.h, F_O2_THETA_EFF, p.soilco2.θ_eff_o2)
    set_field!(b.h, F_CO2_BOT_BC, p.soilco2.bottom_bc); set_field!(b.h, F_O2_BOT_BC, p.soilco2.bottom_bc_o2)
    for (opt, c, idc, idf, f) in ((Int32(2), c_atm_co2, F_CO2_C_ATM, F_CO2_TOP_BC, p.soilco2.top_bc),
                                  (Int32(3), c_atm_o2, F_O2_C_ATM, F_O2_TOP_BC, p.soilco2.top_bc_o2))
        check(ccall((:clb_set_option, libclb), Cint, (Ptr{Cvoid}, Int32, Int64), b.h.ptr, opt, c === nothing ? 0 : 1))
        c === nothing ? set_field!(b.h, idf, f) : set_field!(b.h, idc, c)
    end
    check(ccall((:clb_soilco2_implicit_step, libclb), Cint, (Ptr{Cvoid}, Float64, Int32), b.h.ptr, float(dtγ), Int32(max_iters)))
    get_field!(Y.soilco2.CO2, b.h, F_CO2_Y); get_field!(Y.soilco2.O2, b.h, F_O2_Y)
    return nothing
end

"compute_imp_tendency!(dY, Y, p, t): rre.jl:161-203, energy_hydrology.jl:363-425"
function make_compute_imp_tendency(b::B200Soil)
    function compute_imp_tendency!(dY, Y, p, t)
        # the mirrors already hold Y and p from the cache_imp! call that precedes T_imp! in
        # every Newton iteration (SURVEY 3.2); only the outputs move
        check(ccall((:clb_compute_imp_tendency, libclb), Cint, (Ptr{Cvoid},), b.h.ptr))
        get_field!(dY.soil.ϑ_l, b.h, F_DY_THETA_L)
        get_field!(dY.soil.∫F_vol_liq_water_dt, b.h, F_DY_INTF_W)
        if b.energy
            get_field!(dY.soil.ρe_int, b.h, F_DY_RHO_E_INT)
            get_field!(dY.soil.θ_i, b.h, F_DY_THETA_I)
            get_field!(dY.soil.∫F_e_dt, b.h, F_DY_INTF_E)
        end
        return nothing
    end
end

"""
    B200SoilJacobian

`jac_prototype` replacing `initialize_jacobian(Y)`'s FieldMatrixWithSolver
(src/shared_utilities/implicit_timestepping.jl:63-172): the tridiagonal blocks live in the library,
`ldiv!` is BlockDiagonalSolve (Richards) / BlockLowerTriangularSolve(@name(soil.ϑ_l)) (EnergyHydrology).
"""
struct B200SoilJacobian{B}
    b::B
end
Base.similar(w::B200SoilJacobian) = w
initialize_jacobian(b::B200Soil) = B200SoilJacobian(b)

"compute_jacobian!(W, Y, p, dtγ, t): rre.jl:391-458, energy_hydrology.jl:466-576"
function make_compute_jacobian(b::B200Soil)
    function compute_jacobian!(W::B200SoilJacobian, Y, p, dtγ, t)
        check(ccall((:clb_compute_jacobian, libclb), Cint, (Ptr{Cvoid}, Float64), b.h.ptr, float(dtγ)))
        return nothing
    end
end

function LinearAlgebra.ldiv!(x::Fields.FieldVector, W::B200SoilJacobian, rhs::Fields.FieldVector)
    b = W.b
    set_field!(b.h, F_B_THETA_L, rhs.soil.ϑ_l); set_field!(b.h, F_B_INTF_W, rhs.soil.∫F_vol_liq_water_dt)
    if b.energy
        set_field!(b.h, F_B_RHO_E_INT, rhs.soil.ρe_int); set_field!(b.h, F_B_THETA_I, rhs.soil.θ_i)
        set_field!(b.h, F_B_INTF_E, rhs.soil.∫F_e_dt)
    end
    check(ccall((:clb_ldiv, libclb), Cint, (Ptr{Cvoid},), b.h.ptr))
    get_field!(x.soil.ϑ_l, b.h, F_X_THETA_L); get_field!(x.soil.∫F_vol_liq_water_dt, b.h, F_X_INTF_W)
    if b.energy
        get_field!(x.soil.ρe_int, b.h, F_X_RHO_E_INT); get_field!(x.soil.θ_i, b.h, F_X_THETA_I)
        get_field!(x.soil.∫F_e_dt, b.h, F_X_INTF_E)
    end
    return x
end

# ---- fused level: one call per implicit stage ---------------------------------------------------
"""
    FusedSoilNewton(b; max_iters = 3, tol = nothing)

Passed as `IMEXAlgorithm(ARS111(), FusedSoilNewton(...))` where `NewtonsMethod` goes.  ClimaTimeSteppers
(0.10, `src/solvers/imex_ark.jl` / `newtons_method.jl`) touches a Newton object in three places, all provided here:

  * `step_u!` reads `newtons_method.update_j` and `newtons_method_cache.j` (to refresh the Jacobian at a new time step
    when `update_j` asks for it): the `update_j` field below is the reference's choice
    `UpdateEvery(NewNewtonIteration)` (Simulations.jl:127-135), for which that refresh is a no-op, and the cache
    carries `j = B200SoilJacobian(b)`;
  * `allocate_cache(alg, x_prototype, j_prototype)`;
  * `solve_newton!(alg, cache, x, f!, j!, pre_iteration!, post_implicit!)`, once per implicit stage with x = U
    (initialised to temp, `cache_imp!(U)` already called).  This method calls `j!(cache.j, x)` ONCE -- the integrator's
    closure around `T_imp!.Wfact(j, U, p, dt a_ii, t)`, which is the only place dtγ, p and t are visible -- with
    `Wfact = record_stage!` (below), which stores them in the Newton object instead of building a matrix; then the
    whole loop (max_iters x (Wfact, T_imp!, residual, ldiv!, update, cache_imp!)) runs as ONE kernel and
    `post_implicit!(x)` is called as the reference's loop does after its last iteration.

STATUS: experimental until it has run under Julia (`LandSimulationB200(...; fused = false)` is the default).
"""
mutable struct FusedSoilNewton{B, U}
    b::B
    max_iters::Int
    tol::Float64       # < 0: fixed iteration count (the reference default, Simulations.jl:127-135)
    update_j::U        # read by ClimaTimeSteppers' step_u!
    dtγ::Float64
    t::Any
    p::Any
end
FusedSoilNewton(b; max_iters = 3, tol = nothing) =
    FusedSoilNewton(b, max_iters, isnothing(tol) ? -1.0 : Float64(tol),
                    ClimaTimeSteppers.UpdateEvery(ClimaTimeSteppers.NewNewtonIteration), NaN, nothing, nothing)
ClimaTimeSteppers.allocate_cache(alg::FusedSoilNewton, x_prototype, j_prototype = nothing) = (; j = B200SoilJacobian(alg.b))

"Wfact of the fused level: remembers the stage's dtγ, p, t for solve_newton! (no matrix is built on the host)"
record_stage!(alg::FusedSoilNewton) = (W, Y, p, dtγ, t) -> (alg.dtγ = float(dtγ); alg.t = t; alg.p = p; nothing)

function ClimaTimeSteppers.solve_newton!(alg::FusedSoilNewton, cache, x, f!, j! = nothing, pre_iteration! = nothing,
                                         post_implicit! = nothing)
    b = alg.b
    isnothing(j!) && error("FusedSoilNewton needs the integrator's Jacobian closure (it carries dtγ, p, t)")
    j!(cache.j, x)                      # -> record_stage!: alg.dtγ, alg.p, alg.t of THIS stage
    isfinite(alg.dtγ) && !isnothing(alg.p) || error("FusedSoilNewton: T_imp!.Wfact is not record_stage!(alg)")
    push_state!(b, x); push_lagged!(b, alg.p, alg.t)
    stats = Ref(ClbStats(0, 0, 0.0, 0))
    check(ccall((:clb_implicit_step, libclb), Cint, (Ptr{Cvoid}, Float64, Int32, Float64, Ptr{ClbStats}),
        b.h.ptr, alg.dtγ, alg.max_iters, alg.tol, alg.tol < 0 ? C_NULL : Base.unsafe_convert(Ptr{ClbStats}, stats)))
    get_field!(x.soil.ϑ_l, b.h, F_Y_THETA_L); get_field!(x.soil.∫F_vol_liq_water_dt, b.h, F_Y_INTF_W)
    if b.energy
        get_field!(x.soil.ρe_int, b.h, F_Y_RHO_E_INT); get_field!(x.soil.∫F_e_dt, b.h, F_Y_INTF_E)
    end
    # The reference leaves p at the iterate BEFORE the last update (cache_imp! is the loop's pre_iteration!, not
    # called after the last iteration); nothing downstream reads that state of p (the next explicit stage recomputes
    # the cache from Y), so p is left as the integrator's cache_imp!(U) call before the solve wrote it.
    isnothing(post_implicit!) || post_implicit!(x)
    return nothing
end

# ---- resident state (SURVEY 8f rank 4): the integrator's vector operations and the surface blocks on the device ----
"y .+= a .* x on library mirrors (the integrator's `@. U = u + dt * T_exp`): clb_field_axpy"
field_axpy!(b::B200Soil, y::ClbField, a::Real, x::ClbField) =
    check(ccall((:clb_field_axpy, libclb), Cint, (Ptr{Cvoid}, Int32, Float64, Int32), b.h.ptr, Int32(y), Float64(a), Int32(x)))
"dst .= src on library mirrors: clb_field_copy"
field_copy!(b::B200Soil, dst::ClbField, src::ClbField) =
    check(ccall((:clb_field_copy, libclb), Cint, (Ptr{Cvoid}, Int32, Int32), b.h.ptr, Int32(dst), Int32(src)))

"""
    ldiv_diagonal!(x, w, rhs, b)

`ldiv!` of a DiagonalMatrixRow block of a surface variable -- `canopy.energy.T` with ∂Tres∂T from
src/standalone/Vegetation/canopy_energy.jl:222-250, solved by the reference inside `field_matrix_solve!`
(implicit_timestepping.jl:111-152): x = rhs ./ w with the block's entries uploaded from the host model.
"""
function ldiv_diagonal!(x::Fields.Field, w::Fields.Field, rhs::Fields.Field, b::B200Soil)
    set_field!(b.h, F_SFC_W_DI, w); set_field!(b.h, F_SFC_B, rhs)
    check(ccall((:clb_ldiv_diagonal, libclb), Cint, (Ptr{Cvoid}, Int32, Int32, Int32), b.h.ptr,
        Int32(F_SFC_W_DI), Int32(F_SFC_B), Int32(F_SFC_X)))
    get_field!(x, b.h, F_SFC_X)
    return x
end

"""
    ldiv_all!(x, rhs, b; canopy_T_block = nothing)

`ldiv!(x, W, rhs)` of an integrated model's FieldMatrixWithSolver (src/shared_utilities/implicit_timestepping.jl:63-172)
as ONE launch (clb_ldiv_all): the soil blocks (BlockLowerTriangularSolve(soil.ϑ_l) for EnergyHydrology), the
(soilco2.CO2, soilco2.CO2) / (soilco2.O2, soilco2.O2) tridiagonals when `Y` has a `soilco2` component, and the
DiagonalMatrixRow block of `canopy.energy.T` (∂Tres∂T from canopy_energy.jl:222-250, passed as `canopy_T_block`).
The Jacobian rows are the ones clb_compute_jacobian / clb_soilco2_compute_jacobian left in the mirrors.  The other
explicit variables keep the reference's `x = -rhs` broadcast.
"""
function ldiv_all!(x::Fields.FieldVector, rhs::Fields.FieldVector, b::B200Soil; canopy_T_block = nothing)
    mask = UInt32(1)
    set_field!(b.h, F_B_THETA_L, rhs.soil.ϑ_l); set_field!(b.h, F_B_INTF_W, rhs.soil.∫F_vol_liq_water_dt)
    if b.energy
        set_field!(b.h, F_B_RHO_E_INT, rhs.soil.ρe_int); set_field!(b.h, F_B_THETA_I, rhs.soil.θ_i)
        set_field!(b.h, F_B_INTF_E, rhs.soil.∫F_e_dt)
    end
    if hasproperty(rhs, :soilco2)
        mask |= UInt32(2)
        set_field!(b.h, F_CO2_B, rhs.soilco2.CO2); set_field!(b.h, F_O2_B, rhs.soilco2.O2)
    end
    if !isnothing(canopy_T_block)
        mask |= UInt32(4)
        set_field!(b.h, F_SFC_W_DI, canopy_T_block); set_field!(b.h, F_SFC_B, rhs.canopy.energy.T)
    end
    check(ccall((:clb_ldiv_all, libclb), Cint, (Ptr{Cvoid}, UInt32), b.h.ptr, mask))
    get_field!(x.soil.ϑ_l, b.h, F_X_THETA_L); get_field!(x.soil.∫F_vol_liq_water_dt, b.h, F_X_INTF_W)
    if b.energy
        get_field!(x.soil.ρe_int, b.h, F_X_RHO_E_INT); get_field!(x.soil.θ_i, b.h, F_X_THETA_I)
        get_field!(x.soil.∫F_e_dt, b.h, F_X_INTF_E)
    end
    if hasproperty(rhs, :soilco2)
        get_field!(x.soilco2.CO2, b.h, F_CO2_X); get_field!(x.soilco2.O2, b.h, F_O2_X)
        x.soilco2.SOC .= .-rhs.soilco2.SOC
    end
    isnothing(canopy_T_block) || get_field!(x.canopy.energy.T, b.h, F_SFC_X)
    return x
end

"∫ field dz per column into a surface field (ClimaCore column_integral_definite!, rre.jl:502-511): clb_column_integral"
function column_integral!(out::Fields.Field, b::B200Soil, cell::ClbField, col::ClbField)
    check(ccall((:clb_column_integral, libclb), Cint, (Ptr{Cvoid}, Int32, Int32), b.h.ptr, Int32(cell), Int32(col)))
    get_field!(out, b.h, col)
end

"global sums (Σ area-weighted column water, ∫F_w, column energy, ∫F_e) over all ranks: clb_global_balance"
function global_balance(b::B200Soil)
    out = zeros(Float64, 4)
    check(ccall((:clb_global_balance, libclb), Cint, (Ptr{Cvoid}, Ptr{Float64}), b.h.ptr, out))
    return out
end

"""
    soil_step_host!(b, dt, max_iters, ins, outs)

A WHOLE EnergyHydrology soil step from / to host `Array{Float64}`s in ClimaCore's `parent` layout (level fastest),
for callers without CUDA.jl: `ins` / `outs` are vectors of `ClbField => Array` pairs (normally the prognostic state,
`F_PRECIP` and the heat / bottom boundary fluxes in; the new state out).  clb_soil_step_host; pinned arrays
(`CUDA.pin`) are read and written in place over PCIe, column chunk by column chunk.
"""
function soil_step_host!(b::B200Soil, dt, max_iters, ins::Vector{<:Pair}, outs::Vector{<:Pair})
    fi = Int32[Int32(first(q)) for q in ins]; fo = Int32[Int32(first(q)) for q in outs]
    ai = [last(q) for q in ins]; ao = [last(q) for q in outs]
    GC.@preserve ai ao begin
        pin_ = Ptr{Float64}[pointer(a) for a in ai]; pout = Ptr{Float64}[pointer(a) for a in ao]
        check(ccall((:clb_soil_step_host, libclb), Cint,
            (Ptr{Cvoid}, Float64, Int32, Ptr{Int32}, Ptr{Ptr{Float64}}, Int32, Ptr{Int32}, Ptr{Ptr{Float64}}, Int32),
            b.h.ptr, Float64(dt), Int32(max_iters), fi, pin_, Int32(length(fi)), fo, pout, Int32(length(fo))))
    end
    return nothing
end

"""
    LandSimulationB200(t0, tf, Δt, model; fused = false, kwargs...)

`ClimaLand.Simulations.LandSimulation` (src/simulations/Simulations.jl:115-245) with the implicit
hooks of the soil model replaced; everything else (set_ic!, exp_tendency!, drivers, diagnostics,
callbacks, ClimaTimeSteppers.init) is the reference's own code.  `fused = false` (default): ClimaTimeSteppers' own
Newton loop drives the four hooks, one `ccall` each; `fused = true`: FusedSoilNewton, one kernel per stage
(experimental until it has run under Julia, see its docstring).
"""
function LandSimulationB200(t0, tf, Δt, model; fused = false, max_iters = 3, kwargs...)
    Y, p, _ = ClimaLand.initialize(model)
    b = B200Soil(model, Y, p)
    newton = fused ? FusedSoilNewton(b; max_iters) :
             ClimaTimeSteppers.NewtonsMethod(; max_iters,
                 update_j = ClimaTimeSteppers.UpdateEvery(ClimaTimeSteppers.NewNewtonIteration))
    ts = ClimaTimeSteppers.IMEXAlgorithm(ClimaTimeSteppers.ARS111(), newton)
    sim = ClimaLand.Simulations.LandSimulation(t0, tf, Δt, model; timestepper = ts, kwargs...)
    # swap the implicit side of the ClimaODEFunction the reference built (Simulations.jl:177-199)
    cache! = make_update_implicit_cache(b)
    imp_hook! = make_compute_imp_tendency(b)
    # Fine-grained level: cache_imp!(U) precedes every T_imp! call inside the Newton loop, so the mirrors hold U and p.
    # Fused level: the lane kernel keeps psi / T / K in shared memory and never writes the p mirrors, so a T_imp! call
    # from outside the solve (the integrator's stage tendency) refreshes them first.
    imp! = fused ? ((dY, Y, p, t) -> (cache!(p, Y, t); imp_hook!(dY, Y, p, t))) : imp_hook!
    Wfact = fused ? record_stage!(newton) : make_compute_jacobian(b)
    f = sim._integrator.sol.prob.f
    T_imp! = ClimaTimeSteppers.ODEFunction(imp!; jac_prototype = initialize_jacobian(b), Wfact)
    newf = ClimaTimeSteppers.ClimaODEFunction(; T_exp! = f.T_exp!, T_imp!, dss! = f.dss!,
        cache_imp! = (Y, p, t) -> cache!(p, Y, t))
    prob = ClimaTimeSteppers.ODEProblem(newf, sim._integrator.u, (t0, tf), sim._integrator.p)
    integ = ClimaTimeSteppers.init(prob, ts; dt = Δt, callback = sim.callbacks, adaptive = false)
    return ClimaLand.Simulations.LandSimulation(sim.model, ts, sim.start_date, sim.user_callbacks,
        sim.diagnostics, sim.required_callbacks, sim.callbacks, integ)
end

end # module
