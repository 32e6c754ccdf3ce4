"""B200-native implicit soil-column path of ClimaLand.jl (RichardsModel /
EnergyHydrology implicit tendency, Jacobian, Newton linear solve, fused ARS111
stage).  The product is `libclimaland_b200.so` (hand-written CUDA for sm_100a behind
the C ABI of include/climaland_b200.h); this package is the host-side mirror of the
reference's interface for that path.  There is no CPU fallback."""
from . import _lib
from ._lib import ClbError
from .solver import (BOT_FLUX, BOT_FREE_DRAINAGE, BOT_MOISTURE_STATE, BROOKS_COREY, EARTH, ENERGY_HYDROLOGY,
                     FIELDS, MATH_FAST, MATH_LIBM, RICHARDS, TOP_FLUX, TOP_MOISTURE_STATE, VAN_GENUCHTEN,
                     VARIANT_AUTO, VARIANT_GENERIC, VARIANT_LANE_QUAD, VARIANT_LANE_QUAD_PIPELINED, VARIANT_LANE_OCTET, VARIANT_LANE_PER_CELL, VARIANT_REGISTER_COLUMN, LAYOUT_AUTO,
                     LAYOUT_COLUMN_FASTEST, LAYOUT_LEVEL_FASTEST, SoilColumnSolver)

from .soil import (B200SoilJacobian, BrooksCorey, Column, EnergyHydrology, EnergyHydrologyParameters, FreeDrainage,
                   FusedSoilNewton, HeatFluxBC, IMEXAlgorithm, LandSimulation, MoistureStateBC, NewtonsMethod, PhaseChange,
                   RichardsModel, RichardsParameters, TOPMODELSubsurfaceRunoff, WaterFluxBC, WaterHeatBC,
                   initialize, initialize_jacobian, ldiv, make_compute_imp_tendency, make_compute_jacobian,
                   make_phase_change_source, make_update_aux, make_update_boundary_fluxes, make_update_implicit_cache,
                   vanGenuchten)

__all__ = [n for n in dir() if not n.startswith("_")]
