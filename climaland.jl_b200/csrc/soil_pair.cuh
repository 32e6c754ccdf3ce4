// soil_pair.cuh -- the lane kernels of the fused implicit stage: a column is split over 2 * PARTS lanes of a warp
// (PARTS = 1 lane pair, 2 lane QUAD -- the bench kernel, N = 15 / 16 -- , 4 lane OCTET, N = 15 / 16 / 50 compiled in
// and 17 .. 64 read at run time), Q cells per lane, 32 / (2 PARTS) columns per warp.
//
// The lanes of a column form two halves: the "bottom" half holds the levels from the bottom boundary up to the seam,
// the "top" half the levels from the top boundary down to the seam (NR = 2 PARTS Q level rows; rows >= N are pads at
// the top).  Each lane holds its Q cells ordered from the column's boundary towards the seam, so both halves run the
// SAME straight-line code:
//
//   * the closures (the FP64-heavy part) are evaluated up to four cells at a time with the dependency chains
//     interleaved statement by statement (soil_mathv.cuh): a sub-partition issues one instruction per cycle and one
//     FP64 instruction per 2 cycles with ~8 cycles of latency
//     (tools/ubench/fp64_*.cu), and gets its independent work from ILP because the registers allow 2 warps;
//   * the stencil needs no shuffles inside a lane; only the values at lane boundaries cross lanes
//     (8 double shuffles per Newton iteration in the quad against ~150 of the lane-per-cell kernel);
//   * the tridiagonal systems are solved by the TWISTED (two-sided) Thomas factorisation: each half eliminates from
//     its boundary towards the seam (part after part, the carry crossing lanes by shfl_up), the two seam rows give a
//     2x2 system solved redundantly by both halves, and each half back-substitutes outwards (shfl_down).  The
//     elimination runs on the leading minors (one fma per cell in the dependency chain, the reciprocals afterwards
//     and in parallel);
//   * per-cell stage constants (prepared closure parameters, lagged face coefficients, the factored (rho_e, rho_e)
//     block) live in SHARED memory, in a warp-private tile (a lane only ever touches its own column: no block
//     barriers); with the half in the low lane-group bit the quad's slot accesses are bank-conflict free.  The
//     iterate and the residual's constant part stay in registers;
//   * every field is read from HBM once, by TMA, straight into that tile, where it is transformed in place into the
//     stage constants: the double-buffered kernels fetch a tile of ALL fields with two cp.async.bulk.tensor.3d boxes
//     [fields x NR level rows x CPW columns] of the handle's arena of equally spaced mirrors (a UTMALDG costs the
//     issuing warp ~75 cycles, so one box per field was 7 % of a tile's time), the single-buffered ones and
//     level-fastest mirrors with one [N levels x CPW columns] box per field; persistent warps walk over tiles and
//     request the tile after next (double-buffered) or prefetch the next one to L2 (single-buffered) while they
//     compute; the new state is written once.
//
// Mirror layout: column-fastest (sl = ld, sc = 1) by default; LF = true reads level-fastest mirrors (sl = 1, sc = N,
// N even: TMA needs 16-byte global strides), where a tile is one contiguous piece of each field.
//
// Inward flux convention: F_f is the flux through face f in the direction boundary -> seam
// (upward in the bottom half, downward in the top half), so for both halves
//     F_f = -a_f (h_inner - h_outer) [+ boundary flux],   T_q = (F_q - F_{q+1}) / dz_q
// which is the reference's -(q_hi - q_lo)/dz with exact sign flips (rre.jl:161-203,
// energy_hydrology.jl:363-425); Jacobian rows follow rre.jl:391-458 and
// energy_hydrology.jl:466-576; the block solve implicit_timestepping.jl:160-171.
#pragma once
#include <cuda.h>

#include <type_traits>

#include "soil_device.cuh"
#include "soil_fused.cuh"
#include "soil_mathv.cuh"

namespace clb {

constexpr int kPairQ = 8;  // cell slots per half column
#ifndef CLB_PAIR_W
#define CLB_PAIR_W 4
#endif
constexpr int kPairW = CLB_PAIR_W;  // cells evaluated together (interleaved dependency chains)

// Per-launch grid constants in the lane-local orientation, [half][slot]; a kernel
// parameter (constant bank), indexed with the lane's half at run time.
// HR: cell slots per half column (= PARTS * Q; 8 for the lane pair / quad / octet of 15-16 levels).
template <int HR>
struct PairGridT {
    double z[2][HR];          // z_c of the cell
    double dti[2][HR];        // dtgamma / dz_c of the cell; 0 for pad slots
    double hidzf[2][HR + 1];  // 1/(2 dz_f) of face f (f = q: outer face of slot q, HR: seam); 0 for boundary / pad faces
    int nlev;                     // N, for the instantiations that take the level count at run time (template N = 0)
    int col0;                     // first column of this launch within the mirrors the TMA descriptors describe
                                  // (a launch over a column sub-range: the pointers of DevView are shifted, the
                                  // descriptors are not)
};
using PairGrid = PairGridT<kPairQ>;

// Stage constants per cell: slots [0, NS) in shared memory, the rest in registers.
//   Richards (11)         theta_r, nu, ca, ca2, cb, 1/S_s, cc, cd, K_sat, t1, 1/range
//   EnergyHydrology (18)  theta_r, nu_eff, ice energy, rho_c base, ca, K_lag rho_l c_l, cb, 1/S_s, cc, cd,
//                         aK_o, aC_o (outer-face coefficients), den22, od22 | c22, t1, t2, 1/range
// t1 / t2: constant part of the residual, temp - dtgamma * (implicit source).
// Raw fields (what TMA brings in, same slots before the in-place transform):
//   Richards (9)          nu, theta_r, K_sat, S_s, a, b, m, theta_l, is_sat
//   EnergyHydrology (14)  nu, theta_r, S_s, a, b, m, theta_l, is_sat, theta_i, rho_c_ds, K, kappa, theta_l_lag, rho_e
template <int MODEL>
struct PairSlots {
    static constexpr int kConst = (MODEL == 1) ? 18 : 11;
    static constexpr int kRaw = (MODEL == 1) ? 14 : 9;
};
enum { R_THETA_R = 0, R_NU, R_CA, R_CA2, R_CB, R_INV_SS, R_CC, R_CD, R_KSAT, R_T1, R_IRANGE };
enum { E_THETA_R = 0, E_NU_EFF, E_ICE, E_RCBASE, E_CA, E_KC, E_CB, E_INV_SS, E_CC, E_CD, E_AK, E_AC, E_DEN22, E_OD22,
       E_C22, E_T1, E_T2, E_IRANGE };


// TMA descriptors of the raw fields (column-fastest mirrors: dims {ncol, N}, box {16, N})
struct PairMaps {
    CUtensorMap m[14];
    // When the raw fields lie at a uniform stride in one allocation (the handle's arena) a tile is two boxes of a
    // 3-D tensor {columns, levels, fields}: a = the fa time-invariant parameter fields, b = the others (box {CPW, NR, .}:
    // level rows >= N are zero-filled, so a box lands as whole [NR][CPW] slots).
    CUtensorMap a, b;
    int arena, fa;
};

// Q cells per lane, CPW columns per warp.  A slot of the warp tile is [NR level rows][CPW columns] when the boxes
// come from column-fastest mirrors (RS = 0), or [CPW columns][RS >= NR level rows] when they come from level-fastest
// mirrors (the reference's own layout: a tile is then ONE contiguous piece of each field).
template <int NS, int NTOT, int Q, int CPW, int NR = 16, int RS = 0, int SE = 0>
struct PairStore {
    double *base;  // warp tile + this lane's column
    int half, r0;  // r0: first slot of this lane within its half (0, or Q for the inner lane of a quad)
    double r[(NTOT > NS) ? (NTOT - NS) : 1][Q];
    // level row of cell q: slot r0 + q of the half, counted from the column's boundary
    __device__ __forceinline__ double *at(int q, int slot) const
    {
        const int row = half ? NR - 1 - (r0 + q) : r0 + q;
        if constexpr (RS > 0) return base + slot * SE + row;  // base = tile + column * RS; SE doubles per slot
        return base + (slot * NR + row) * CPW;                        // base = tile + column
    }
    template <int SLOT>
    __device__ __forceinline__ double get(int q) const
    {
        if constexpr (SLOT < NS)
            return *at(q, SLOT);
        else
            return r[SLOT - NS][q];
    }
    template <int SLOT>
    __device__ __forceinline__ void put(int q, double v)
    {
        if constexpr (SLOT < NS)
            *at(q, SLOT) = v;
        else
            r[SLOT - NS][q] = v;
    }
};

template <int MASK>
__device__ __forceinline__ double xchg(double v) { return __shfl_xor_sync(0xffffffffu, v, MASK); }

// ---- TMA / mbarrier (sm_90+ PTX; one barrier per warp, single phase) -------------------------
__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned bar, unsigned count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned bar, unsigned bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned bar, unsigned parity)
{
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE_%=;\n\t"
        "bra WAIT_%=;\n\t"
        "DONE_%=:\n\t}" ::"r"(bar), "r"(parity)
        : "memory");
}
// box {16 columns, N levels} of a column-fastest mirror at (column c0, level 0) -> dense [level][16] tile
__device__ __forceinline__ void tma_load_2d(unsigned dst, const CUtensorMap *map, int c0, int l0, unsigned bar)
{
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(dst),
        "l"(map), "r"(c0), "r"(l0), "r"(bar)
        : "memory");
}

__device__ __forceinline__ void tma_load_3d(unsigned dst, const CUtensorMap *map, int c0, int l0, int f0, unsigned bar)
{
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];" ::"r"(dst),
        "l"(map), "r"(c0), "r"(l0), "r"(f0), "r"(bar)
        : "memory");
}

// the same box, only as far as L2 (no shared memory is committed to it)
__device__ __forceinline__ void tma_prefetch_l2_2d(const CUtensorMap *map, int c0, int l0)
{
    asm volatile("cp.async.bulk.prefetch.tensor.2d.L2.global.tile [%0, {%1, %2}];" ::"l"(map), "r"(c0), "r"(l0) : "memory");
}
__device__ __forceinline__ void prefetch_l2(const void *p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }

// Prepared closure constants:
//   van Genuchten  ca = 1/m, ca2 = m, cb = 1/n, cc = 1/alpha, cd = 1/(alpha m n range)
//   Brooks-Corey   ca = -1/c, ca2 = 2/c + 3, cb = psi_b, cc = -psi_b/(c range), cd unused
struct ClosureConst {
    double ca, ca2, cb, cc, cd, inv_Ss, inv_range;
};
// The time-invariant reciprocals are not recomputed every stage: the kernel reads PREPARED mirrors
// (k_prepare_params, written when a parameter field changes) in the slots of S_s / a / b / m:
//   van Genuchten  1/S_s, 1/alpha, 1/n, 1/m          Brooks-Corey  1/S_s, -1/c, psi_b, (unused)
// so a stage needs one reciprocal per cell (1/range, range depends on theta_i) instead of five.
template <int CLOSURE, bool WANT_M>
__device__ __forceinline__ ClosureConst pair_prepare(double inv_Ss, double pa, double pb, double pm, double theta_r, double nu_eff)
{
    const double theta_lo = theta_r + kSqrtEps;
    const double inv_range = fm::rcp(fmv::max_nn(nu_eff, theta_lo) - theta_r);  // fmax is 11 instructions, compare + select 3
    ClosureConst c;
    c.inv_Ss = inv_Ss;
    c.inv_range = inv_range;
    if (CLOSURE == kVanGenuchten) {
        c.ca = pm;
        c.ca2 = WANT_M ? fm::rcp(pm) : 0.0;  // m itself: only the conductivity (Richards) needs it
        c.cb = pb;
        c.cc = pa;
        c.cd = ((pa * pm) * pb) * inv_range;
    } else {
        c.ca = pa;
        c.ca2 = fma(-2.0, pa, 3.0);
        c.cb = pb;
        c.cc = (pb * pa) * inv_range;
        c.cd = 0.0;
    }
    return c;
}

// elementwise over the (padded) per-cell mirrors; exact IEEE divisions, run once per parameter change
template <int CLOSURE>
__global__ void k_prepare_params(const double *__restrict__ S_s, const double *__restrict__ a, const double *__restrict__ b,
                                 const double *__restrict__ m, double *__restrict__ o_ss, double *__restrict__ o_a,
                                 double *__restrict__ o_b, double *__restrict__ o_m, int64_t n)
{
    const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    o_ss[k] = 1.0 / S_s[k];
    if (CLOSURE == kVanGenuchten) {
        o_a[k] = 1.0 / a[k];
        o_b[k] = 1.0 / b[k];
        o_m[k] = 1.0 / m[k];
    } else {
        o_a[k] = -1.0 / a[k];
        o_b[k] = b[k];
    }
}

#ifndef CLB_PAIR_BLOCK
#define CLB_PAIR_BLOCK 64
#endif

// Lane geometry of a column split over 2 * PARTS lanes (PARTS = 1: lane pair, 2: lane quad, 4: lane octet),
// Q cells per lane, NR = 2 * PARTS * Q level rows per slot (>= N; the pads sit at the top of the column).
//   lane = column-in-warp + CPW * (part * 2 + half);  part 0 touches the column boundary,
//   part PARTS-1 the seam.  xor CPW swaps the halves; the previous / next part of a half is 2*CPW lanes
//   down / up.  With the half in the low bit a half-warp (the unit of a 64-bit shared-memory access) holds the
//   level rows q' and NR-1 - q' of a slot: opposite parity, i.e. different banks for 8-column rows, so the
//   slot accesses of the quad are conflict-free (half-major order put rows q and q + 4 on the same banks).
template <int PARTS, int Q_ = kPairQ / PARTS>
struct LaneGeom {
    static_assert(PARTS == 1 || PARTS == 2 || PARTS == 4, "lane pair, quad or octet");
    static constexpr int LPC = 2 * PARTS;     // lanes per column
    static constexpr int CPW = 32 / LPC;      // columns per warp
    static constexpr int Q = Q_;              // cells per lane
    static constexpr int HR = PARTS * Q;      // cell slots per half column
    static constexpr int NR = 2 * HR;         // level rows of a slot
    static constexpr int SEAM = CPW;          // xor mask: the other half of the column
    static constexpr int PARTD = 2 * CPW;     // lane distance to the next part of this half
    static constexpr int kSlotBytes = NR * CPW * 8;  // one slot of a warp tile: NR level rows x CPW columns
    // level-fastest tiles: rows per column in shared memory = the TMA box's level extent, >= NR, a multiple of 2
    // (16-byte box rows) chosen so that the CPW columns of a row fall on different banks (NR = 56 -> 58: column
    // stride 116 words = 20 mod 32)
    static constexpr int kRowsLF = NR + 2;
    static constexpr int kSlotBytesLF = (kRowsLF * CPW * 8 + 127) / 128 * 128;  // TMA destinations are 128-byte aligned
};

// value of the previous (outer) / next (inner) part's lane of this half; own value where there is none
template <int DIST>
__device__ __forceinline__ double from_prev_part(double v) { return __shfl_up_sync(0xffffffffu, v, DIST); }
template <int DIST>
__device__ __forceinline__ double from_next_part(double v) { return __shfl_down_sync(0xffffffffu, v, DIST); }

// Values of the neighbours across lane boundaries: `inner` = neighbour of this lane's last cell
// (the other half's last cell at the seam, else the first cell of the next part), `outer` =
// neighbour of its first cell (last cell of the previous part; finite garbage for part 0, whose
// outer face is the column boundary and carries a zero coefficient).
template <class Gm>
__device__ __forceinline__ void nb_exchange(double first, double last, bool innermost, double &outer, double &inner)
{
    const double seam = xchg<Gm::SEAM>(last);
    if (Gm::LPC == 2) {
        inner = seam;
        outer = first;
        // (The quad's two part-crossing shuffles as ONE xor-2CPW exchange of `innermost ? first : last` were measured:
        // 48.3 against 48.8 us for EnergyHydrology, 31.2 against 31.0 us for Richards, and nothing on top of the I2F
        // change of log_tab -- not kept.)
    } else {
        const double nb_first = from_next_part<Gm::PARTD>(first), nb_last = from_prev_part<Gm::PARTD>(last);
        inner = innermost ? seam : nb_first;
        outer = nb_last;
    }
}

// Dynamic shared memory of a block: the log / exp tables, then NBUF NS-slot tiles and NBUF mbarriers per warp.
template <int PARTS, int NS, int NBUF, int BLOCK, int Q = kPairQ / PARTS, bool LF = false>
__host__ __device__ constexpr size_t pair_smem_bytes()
{
    using Gm = LaneGeom<PARTS, Q>;
    return (size_t)fmv::kMathTabBytes + (size_t)(BLOCK / 32) * NBUF * (NS * (LF ? Gm::kSlotBytesLF : Gm::kSlotBytes) + 8);
}

// Per-column scalars of a tile, fetched one tile ahead into registers.
struct ColScalars {
    double Rss, hg, top_w, bot_w, Ress, top_h, bot_h, intF_w, intF_e;
    double theta_bc;  // BCL: the boundary value of this lane's half (MoistureStateBC), else unused
};
template <int MODEL, bool BCL = false>
__device__ __forceinline__ ColScalars load_col_scalars(const DevView &P, int64_t cs, int half = 0)
{
    ColScalars v;
    v.theta_bc = 0.0;
    // fetched a tile ahead with the other per-column scalars: read where it is used (the set-up), it was one exposed
    // HBM round trip per tile
    if (BCL) v.theta_bc = half ? P.theta_bc_top[cs] : ((P.bottom_bc == 2) ? P.theta_bc_bot[cs] : 0.0);
    v.Rss = P.R_ss[cs];
    v.hg = P.h_grad[cs];
    v.top_w = P.top_bc_w[cs];
    v.bot_w = P.bot_bc_w[cs];
    v.intF_w = P.Y_intF_w[cs];
    v.Ress = v.top_h = v.bot_h = v.intF_e = 0.0;
    if (MODEL == 1) {
        v.Ress = P.R_ess[cs];
        v.top_h = P.top_bc_h[cs];
        v.bot_h = P.bot_bc_h[cs];
        v.intF_e = P.Y_intF_e[cs];
    }
    return v;
}

// PERSISTENT: every warp walks over tiles (CPW columns each) t = warp, warp + nwarps, ...; with
// NBUF = 2 the TMA boxes and the per-column scalars of the next tile are requested before the
// current tile is touched, so HBM latency hides behind a whole tile of FP64 work.
// NT: the level count as a template parameter (the N = 15 / 16 / 50 instantiations: every `level < N` and the position
// of the top boundary fold at compile time), or 0: N is read from G.nlev (NR - 8 < N <= NR, any parity).
// Phase clocks (tuning builds only, -DCLB_PHASE_CLOCKS; tools/phase_clocks.py): lane 0 of every warp adds the SM
// clock ticks it spent in each phase of a tile; read back through clb_debug_phase_clocks.
#ifdef CLB_PHASE_CLOCKS
__device__ unsigned long long g_phase_clk[16];
#define CLB_PC_DECL unsigned pc_[16] = {}; unsigned pc_t_ = (unsigned)clock();
#define CLB_PC(k) { const unsigned now_ = (unsigned)clock(); pc_[k] += now_ - pc_t_; pc_t_ = now_; }
#define CLB_PC_FLUSH if (lane == 0) { _Pragma("unroll") for (int k_ = 0; k_ < 16; ++k_) atomicAdd(&g_phase_clk[k_], (unsigned long long)pc_[k_]); }
#else
#define CLB_PC_DECL
#define CLB_PC(k)
#define CLB_PC_FLUSH
#endif
#if defined(CLB_PHASE_CLOCKS) && defined(CLB_PC_SETUP)
#define CLB_PCS(k) CLB_PC(k)  // sub-phases of the set-up (slots 12 .. 15; the rest of it stays in slot 2)
#else
#define CLB_PCS(k)
#endif
// BCL (RichardsModel only): the cache holds dfluxBCdY, i.e. the top boundary is a MoistureStateBC (rre.jl:460-468): the
// boundary fluxes follow the iterate -- top: -K_N ((psi_bc + dz_top) - psi_N) / dz_top with dfluxBCdY = K_N dpsi_N / dz_top
// on the top cell's diagonal (boundary_conditions.jl:227-267, 375-411, rre.jl:434-449); bottom: FreeDrainage -K_1, a
// MoistureStateBC, or the lagged flux value -- the flux integral's Newton recurrence runs with them iteration by
// iteration, and the last evaluation is left in p.soil.top_bc / bottom_bc.
template <int CLOSURE, int MODEL, int NT, int PARTS, int NS, int NBUF, int BLOCK, int MINB, int QC = kPairQ / PARTS,
          bool LF = false, bool BCL = false>
__global__ void __launch_bounds__(BLOCK, MINB)
    k_step_lanes(const DevView P, const PairGridT<PARTS * QC> G, const __grid_constant__ PairMaps M, double dtg,
                 int max_iters)
{
    using Gm = LaneGeom<PARTS, QC>;
    constexpr int Q = Gm::Q, CPW = Gm::CPW, NR = Gm::NR;
    // cells evaluated together: a divisor of Q close to kPairW (Q = 7 -> one group of 7 would not fit: 4 + 3)
    constexpr int WFULL = (Q <= kPairW) ? Q : kPairW;
    static_assert(NT <= NR, "N <= NR level rows");
    const int N = (NT > 0) ? NT : G.nlev;
    // The NR - N pad rows sit at the top of the column, i.e. at the outer end of the top half: its first real slot is
    // slot qT of part pT (all lanes of parts < pT hold pads only: identity rows, zero face coefficients).
    const int Q0T = NR - N, pT = Q0T / Q, qT = Q0T % Q;
    static_assert(NBUF == 1 || NBUF == 2, "single or double buffered");
    constexpr int NTOT = PairSlots<MODEL>::kConst, NRAW = PairSlots<MODEL>::kRaw;
    static_assert(NS >= NRAW && NS <= NTOT, "the raw fields are staged in the shared-memory slots");
    extern __shared__ __align__(128) unsigned char pair_sm_all[];
    static_assert(fmv::kMathTabBytes % 128 == 0 && BLOCK >= mtab::kLogN, "table block keeps the TMA tiles 128-byte aligned");

    const int tid = threadIdx.x, lane = tid & 31, wib = tid >> 5;
    const fmv::MathTab MT = fmv::math_tab_fill(pair_sm_all, tid);
    unsigned char *const pair_sm = pair_sm_all + fmv::kMathTabBytes;
    const int idx = lane / CPW, half = idx & 1, part = idx >> 1;
    const int r0 = part * Q;
    const bool innermost = (part == PARTS - 1), outermost = (part == 0);
    // warp w of block b starts at tile w * gridDim + b: when the tiles do not divide evenly the extra ones go to
    // one warp of every SM first instead of to every warp of the first SMs
    const int64_t warp0 = (int64_t)wib * gridDim.x + blockIdx.x;
    const int64_t nwarps = (int64_t)gridDim.x * (BLOCK / 32);
    const int64_t ntiles = (P.ncol + CPW - 1) / CPW;
    const EarthConst &E = P.earth;
    const double C1 = E.cp_l * E.rho_l, C2 = E.cp_i * E.rho_i, T_ref = E.T_ref;

    constexpr int kSlotB = LF ? Gm::kSlotBytesLF : Gm::kSlotBytes;  // one field of a tile
    constexpr int RS = LF ? Gm::kRowsLF : 0;
    constexpr size_t kTileBytes = (size_t)NS * kSlotB;
    unsigned char *const tiles = pair_sm + (size_t)wib * NBUF * kTileBytes;
    const unsigned bar0 = smem_u32(pair_sm + (size_t)(BLOCK / 32) * NBUF * kTileBytes + (size_t)wib * NBUF * 8);
    PairStore<NS, NTOT, Q, CPW, NR, RS, kSlotB / 8> S;
    S.half = half;
    S.r0 = r0;
    auto level_of = [&](int q) { return half ? NR - 1 - (r0 + q) : r0 + q; };
    auto col_of = [&](int64_t t) { return t * CPW + (lane % CPW); };
    auto col_clamped = [&](int64_t t) { const int64_t c_ = col_of(t); return c_ < P.ncol ? c_ : P.ncol - 1; };

    // every per-cell field of a tile: HBM -> shared memory, one TMA box per field.  `which`: 1 = the
    // time-invariant parameter fields (safe to fetch while the previous launch of the stream still runs),
    // 2 = the stage inputs (state, lagged cache), 3 = all
    auto is_param = [&](int j) {
        return (MODEL == 1) ? (j <= 5 || j == 9) : (j <= 6);
    };
    auto request_tile = [&](int64_t t, int buf, int which) {
        const unsigned bar = bar0 + buf * 8;
        // One lane issues the boxes back to back.  (One box per lane -- lane j fetching field j -- costs ~1 000 cycles
        // per tile: UTMALDG takes its operands from uniform registers, so the divergent form becomes a loop over the
        // active lanes, ELECT + R2UR + UTMALDG each time; tools/phase_clocks.py.)
        if (!LF && NBUF == 2 && M.arena) {
            // Two boxes instead of one per field: a UTMALDG blocks the issuing warp for ~75 cycles (measured,
            // tools/phase_clocks.py: 1 050 ticks per tile for 14 boxes).  Double-buffered kernels only: a box is
            // fetched row by row, and where the warp waits for the tile it has just requested (NBUF = 1) the many
            // small boxes land sooner than two large ones (measured: the run-time-N octet 8-20 % slower with two).
            if (lane == 0) {
                if (which & 1) mbar_expect_tx(bar, (unsigned)NRAW * kSlotB);
                const unsigned dst0 = smem_u32(tiles + buf * kTileBytes);
                const int cc_ = G.col0 + (int)(t * CPW);
                if (which & 1) tma_load_3d(dst0, &M.a, cc_, 0, 0, bar);
                if (which & 2) tma_load_3d(dst0 + M.fa * kSlotB, &M.b, cc_, 0, M.fa, bar);
            }
        } else if (lane == 0) {
            // bytes of a box: out-of-bounds rows of a level-fastest box (levels >= N, zero-filled) count as well
            if (which & 1)
                mbar_expect_tx(bar, (unsigned)(NRAW - ((CLOSURE == kVanGenuchten) ? 0 : 1)) * (LF ? Gm::kRowsLF : N) * CPW * 8);
            const unsigned dst0 = smem_u32(tiles + buf * kTileBytes);
            const int cc_ = G.col0 + (int)(t * CPW);
#pragma unroll
            for (int j = 0; j < NRAW; ++j) {
                const bool skip = (CLOSURE != kVanGenuchten) && j == ((MODEL == 1) ? 5 : 6);  // no m field for Brooks-Corey
                const bool mine = is_param(j) ? (which & 1) : (which & 2);
                // coordinates in the order of the tensor's dimensions: {column, level} column-fastest, {level, column} level-fastest
                if (!skip && mine) tma_load_2d(dst0 + j * kSlotB, &M.m[j], LF ? 0 : cc_, LF ? cc_ : 0, bar);
            }
        }
    };

    if (lane == 0) {
#pragma unroll
        for (int b_ = 0; b_ < NBUF; ++b_) mbar_init(bar0 + b_ * 8, 1);
    }
    __syncwarp();
    const bool has_tile = warp0 < ntiles;
    if (has_tile) request_tile(warp0, 0, 1);
    __syncthreads();  // the tables are in place; the only block-wide barrier: everything below is warp-private
    // Programmatic dependent launch (both are no-ops without the launch attribute): up to here this launch
    // only read time-invariant parameter fields, so it may overlap the tail of the stream's previous kernel;
    // everything below waits for that kernel to complete.  The next launch is released only now, so its
    // prologue can overlap this kernel alone, never one further back.
    asm volatile("griddepcontrol.wait;" ::: "memory");
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    if (!has_tile) return;  // whole warp
    request_tile(warp0, 0, 2);
    if (NBUF == 2 && warp0 + nwarps < ntiles) request_tile(warp0 + nwarps, 1, 3);  // untouched buffer: no proxy fence needed
    ColScalars nxt = load_col_scalars<MODEL, BCL>(P, col_clamped(warp0), half);

    double dx2_acc = 0.0, bad = 0.0;
    CLB_PC_DECL
    unsigned phase = 0;  // bit b: parity the next wait on buffer b expects
    int buf = 0;
#pragma unroll 1
    for (int64_t tile_id = warp0; tile_id < ntiles; tile_id += nwarps) {
    const int64_t c = col_of(tile_id);
    const bool col_ok = c < P.ncol;
    const ColScalars cur = nxt;
    if (NBUF == 1) {
        // single buffer: the next tile cannot land in shared memory yet; start it towards L2 so that the
        // TMA issued after this tile's store is an L2 hit
        const int64_t tn = tile_id + nwarps;
        if (tn < ntiles) {
            const bool skip = (CLOSURE != kVanGenuchten) && lane == ((MODEL == 1) ? 5 : 6);
            if (lane < NRAW && !skip)
                tma_prefetch_l2_2d(&M.m[lane], LF ? 0 : G.col0 + (int)(tn * CPW), LF ? G.col0 + (int)(tn * CPW) : 0);
            if (lane >= 16 && lane < 16 + ((MODEL == 1) ? 9 : 5)) {  // the CPW per-column scalars of an array share a line
                const int j_ = lane - 16;
                const double *a_ = P.R_ss;
                a_ = (j_ == 1) ? P.h_grad : a_;
                a_ = (j_ == 2) ? P.top_bc_w : a_;
                a_ = (j_ == 3) ? P.bot_bc_w : a_;
                a_ = (j_ == 4) ? P.Y_intF_w : a_;
                a_ = (j_ == 5) ? P.R_ess : a_;
                a_ = (j_ == 6) ? P.top_bc_h : a_;
                a_ = (j_ == 7) ? P.bot_bc_h : a_;
                a_ = (j_ == 8) ? P.Y_intF_e : a_;
                prefetch_l2(a_ + tn * CPW);
            }
        }
    }
    S.base = reinterpret_cast<double *>(tiles + buf * kTileBytes) + (lane % CPW) * (LF ? RS : 1);

    const double ld_Rss = cur.Rss, ld_hg = cur.hg;
    const double top_w = cur.top_w, bot_w = cur.bot_w;
    const double ld_Ress = cur.Ress, top_h = cur.top_h, bot_h = cur.bot_h;
    const double inv_hg = fm::rcp(fmv::max_nn(ld_hg, kEps));
    const double src_w = ld_Rss * inv_hg, src_e = ld_Ress * inv_hg;
    // boundary flux of this half in the inward convention; it enters at face 0 of part 0 (bottom
    // half, top half when N == 16) or at face Q0T of part 0 (top half with pads)
    const double bin_w = half ? -top_w : bot_w, bin_e = half ? -top_h : bot_h;
    const bool top_lane = half && part == pT;  // the lane of the top half that holds the top cell
    const bool at0 = half ? (top_lane && qT == 0) : outermost, atT = top_lane && qT > 0;
    const double b0_w = at0 ? bin_w : 0.0, b0_e = at0 ? bin_e : 0.0;
    const double bT_w = atT ? bin_w : 0.0, bT_e = atT ? bin_e : 0.0;  // enters at the lane's face qT > 0

    CLB_PC(0)  // tile prologue (column scalars, flux integrals)
    mbar_wait(bar0 + buf * 8, (phase >> buf) & 1u);
    phase ^= 1u << buf;
    CLB_PC(1)  // waiting for the tile

    // flux integrals (W = -I, lagged boundary fluxes): their Newton recurrence does not depend on
    // the iterate, so it runs here, off the hot loop (one lane per column stores it); placed after the wait so that
    // its dependent chain interleaves with the set-up below
    static_assert(!BCL || MODEL == 0, "only a RichardsModel cache holds dfluxBCdY");
    double dx2_int = 0.0;
    if (!BCL) {
        // The recurrence U <- U - (-(t + dtg T - U)) started at U = t, written out: U_1 = t + ((t + dtg T) - t), and
        // from the second iteration on U = a := t + dtg T exactly (a - U_1 is exactly representable and U_1 + (a - U_1)
        // = a), with dx = -(a - U_1) in the second iteration and -0 afterwards.  Straight-line code: as a loop with a
        // run-time trip count it was a serial chain of ~4 dependent DADDs per iteration in basic blocks of its own.
        const double tiw = cur.intF_w, tie = cur.intF_e;
        const double Tiw = -(top_w - bot_w) - ld_Rss, Tie = -(top_h - bot_h) - ld_Ress;
        const double aw = tiw + dtg * Tiw, ae = tie + dtg * Tie;
        const double dw1 = -(aw - tiw), de1 = -(ae - tie);
        const double Uw1 = tiw - dw1, Ue1 = tie - de1;
        const double dw2 = -(aw - Uw1), de2 = -(ae - Ue1);
        const double Uw = (max_iters >= 2) ? Uw1 - dw2 : Uw1, Ue = (max_iters >= 2) ? Ue1 - de2 : Ue1;
        const double dxw = (max_iters == 1) ? dw1 : ((max_iters == 2) ? dw2 : 0.0);
        const double dxe = (max_iters == 1) ? de1 : ((max_iters == 2) ? de2 : 0.0);
        if (idx == 0 && col_ok) {
            dx2_int = dxw * dxw + dxe * dxe;
            P.out_intF_w[c] = Uw;
            if (MODEL == 1) P.out_intF_e[c] = Ue;
        }
    }


    // ---- set-up: transform the raw fields in place into the stage constants ------------------
    double U1[Q], U2[Q];
    double aK8 = 0.0, aC8 = 0.0, r22 = 0.0, c22_last = 0.0;  // inner face of the last cell; seam of W22
    double psi_bc = 0.0;                                     // BCL: pressure head of this half's boundary value
    double live_top = top_w, live_bot = bot_w, Uiw = cur.intF_w, dxw_last = 0.0;  // BCL: fluxes at the iterate, flux integral
    // The raw values of ALL of the lane's cells are loaded before the first constant is stored: the stores go to the
    // same slots (the transform is in place), so with loads and stores alternating cell by cell a load may not move
    // above the stores before it and the cells' reciprocal chains run one after the other.  (Phase clocks: set-up
    // 2 810 -> 2 580 ticks per tile for EnergyHydrology, 1 280 -> 1 120 for Richards; the launch time moves < 1 %.)
    if (MODEL == 0) {
        double nu[Q], theta_r[Q], K_sat[Q], iSs[Q], pa[Q], pb[Q], pm[Q], theta[Q], sat[Q];
#pragma unroll
        for (int q = 0; q < Q; ++q) {
            const bool real = level_of(q) < N;
            nu[q] = S.template get<0>(q); theta_r[q] = S.template get<1>(q); K_sat[q] = S.template get<2>(q);
            iSs[q] = S.template get<3>(q); pa[q] = S.template get<4>(q); pb[q] = S.template get<5>(q);
            pm[q] = (CLOSURE == kVanGenuchten) ? S.template get<6>(q) : 0.0;
            theta[q] = S.template get<7>(q); sat[q] = S.template get<8>(q);
            if (!real) {  // pad slot: benign parameters, identity rows (dti = 0, zero face coefficients)
                nu[q] = 0.5; theta_r[q] = 0.1; K_sat[q] = 0.0; iSs[q] = 1e3; pb[q] = (CLOSURE == kVanGenuchten) ? 0.5 : 2.0;
                pa[q] = (CLOSURE == kVanGenuchten) ? 1.0 : -1.0; pm[q] = 2.0;
                theta[q] = 0.3; sat[q] = 0.0;
            }
        }
        ClosureConst cc[Q];
#pragma unroll
        for (int q = 0; q < Q; ++q) cc[q] = pair_prepare<CLOSURE, true>(iSs[q], pa[q], pb[q], pm[q], theta_r[q], nu[q]);
        if (BCL) {
            // pressure head of the boundary value with the boundary CELL's parameters (boundary_conditions.jl:240-250): the
            // top cell is slot qT of the top lane, the bottom cell slot 0 of the bottom half's outer lane
            const int qb = half ? qT : 0;
            double th_[1], thr_[1] = {0.0}, nu_[1] = {0.0}, irg_[1] = {0.0}, ca_[1] = {0.0}, ca2_[1] = {0.0}, cb_[1] = {0.0},
                   cc_[1] = {0.0}, cd_[1] = {0.0}, iss_[1] = {0.0}, ks_[1] = {0.0}, K_[1], psi_[1], dp_[1];
#pragma unroll
            for (int q = 0; q < Q; ++q)
                if (q == qb) {
                    thr_[0] = theta_r[q]; nu_[0] = nu[q]; irg_[0] = cc[q].inv_range; ca_[0] = cc[q].ca; ca2_[0] = cc[q].ca2;
                    cb_[0] = cc[q].cb; cc_[0] = cc[q].cc; cd_[0] = cc[q].cd; iss_[0] = cc[q].inv_Ss; ks_[0] = K_sat[q];
                }
            const double tb = (half || P.bottom_bc == 2) ? cur.theta_bc : nu_[0];
            th_[0] = tb;
            fmv::closure<CLOSURE, false, 1, true>(MT, th_, thr_, nu_, irg_, ca_, ca2_, cb_, cc_, cd_, iss_, ks_, K_, psi_, dp_);
            psi_bc = psi_[0];
        }
#pragma unroll
        for (int q = 0; q < Q; ++q) {
            U1[q] = theta[q];
            S.template put<R_THETA_R>(q, theta_r[q]);
            S.template put<R_NU>(q, nu[q]);
            S.template put<R_CA>(q, cc[q].ca);
            S.template put<R_CA2>(q, cc[q].ca2);
            S.template put<R_CB>(q, cc[q].cb);
            S.template put<R_INV_SS>(q, cc[q].inv_Ss);
            S.template put<R_CC>(q, cc[q].cc);
            S.template put<R_CD>(q, cc[q].cd);
            S.template put<R_KSAT>(q, K_sat[q]);
            S.template put<R_T1>(q, fma(-dtg, src_w * sat[q], theta[q]));
            S.template put<R_IRANGE>(q, cc[q].inv_range);
        }
    } else {
        // lagged fields that couple neighbours: K, kappa and 1/rho_c_s at the LAGGED theta_l
        // (the Jacobian's choice, energy_hydrology.jl:561-566)
        double Kl[Q], kap[Q], rc[Q];
#pragma unroll
        for (int q = 0; q < Q; ++q) {
            const bool real = level_of(q) < N;
            Kl[q] = real ? S.template get<10>(q) : 0.0;
            kap[q] = real ? S.template get<11>(q) : 0.0;
            rc[q] = fm::rcp(volumetric_heat_capacity(real ? S.template get<12>(q) : 0.0, real ? S.template get<8>(q) : 0.0,
                                                     real ? S.template get<9>(q) : 1e6, E));
        }
        CLB_PCS(12)  // lagged loads, 1/rho_c
        double K_out, K_in, kap_out, kap_in, rc_out, rc_in;
        nb_exchange<Gm>(Kl[0], Kl[Q - 1], innermost, K_out, K_in);
        nb_exchange<Gm>(kap[0], kap[Q - 1], innermost, kap_out, kap_in);
        nb_exchange<Gm>(rc[0], rc[Q - 1], innermost, rc_out, rc_in);
        aK8 = (Kl[Q - 1] + K_in) * G.hidzf[half][r0 + Q];
        aC8 = (kap[Q - 1] + kap_in) * G.hidzf[half][r0 + Q];
        CLB_PCS(13)  // exchange of the lagged fields
        double aCo[Q], o22[Q], d22[Q], i22[Q];
        // CH cells at a time: all Q of them up to 4 cells per lane; one by one beyond (11 raw values per cell: with
        // Q = 7 the N = 50 kernel, already at 254 registers, spills -- measured 498 against 456 us)
        constexpr int CH = (Q <= 4) ? Q : 1;
#pragma unroll
        for (int q0 = 0; q0 < Q; q0 += CH) {
            double nu[CH], theta_r[CH], iSs[CH], pa[CH], pb[CH], pm[CH], theta[CH], sat[CH], theta_i[CH], rcds[CH], rho_e[CH];
#pragma unroll
            for (int j = 0; j < CH; ++j) {
                const int q = q0 + j;
                const bool real = level_of(q) < N;
                nu[j] = S.template get<0>(q); theta_r[j] = S.template get<1>(q);
                iSs[j] = S.template get<2>(q); pa[j] = S.template get<3>(q); pb[j] = S.template get<4>(q);
                pm[j] = (CLOSURE == kVanGenuchten) ? S.template get<5>(q) : 0.0;
                theta[j] = S.template get<6>(q); sat[j] = S.template get<7>(q); theta_i[j] = S.template get<8>(q);
                rcds[j] = S.template get<9>(q); rho_e[j] = S.template get<13>(q);
                if (!real) {
                    nu[j] = 0.5; theta_r[j] = 0.1; iSs[j] = 1e3; pb[j] = (CLOSURE == kVanGenuchten) ? 0.5 : 2.0;
                    pa[j] = (CLOSURE == kVanGenuchten) ? 1.0 : -1.0; pm[j] = 2.0;
                    theta[j] = 0.3; sat[j] = 0.0; theta_i[j] = 0.0; rcds[j] = 1e6; rho_e[j] = 0.0;
                }
            }
            ClosureConst cc[CH];
#pragma unroll
            for (int j = 0; j < CH; ++j) cc[j] = pair_prepare<CLOSURE, false>(iSs[j], pa[j], pb[j], pm[j], theta_r[j], nu[j] - theta_i[j]);
#pragma unroll
            for (int j = 0; j < CH; ++j) {
                const int q = q0 + j;
                const double nu_eff = nu[j] - theta_i[j];
                U1[q] = theta[j];
                U2[q] = rho_e[j];
                // lagged coefficients of the cell's outer face (zero table entry at the column boundary)
                const double hid_o = G.hidzf[half][r0 + q];
                const double aK_o = (Kl[q] + ((q == 0) ? K_out : Kl[q - 1])) * hid_o;
                aCo[q] = (kap[q] + ((q == 0) ? kap_out : kap[q - 1])) * hid_o;
                S.template put<E_THETA_R>(q, theta_r[j]);
                S.template put<E_NU_EFF>(q, nu_eff);
                S.template put<E_ICE>(q, theta_i[j] * E.rho_i * E.LH_f0);
                S.template put<E_RCBASE>(q, fma(theta_i[j], C2, rcds[j]));
                S.template put<E_CA>(q, cc[j].ca);
                S.template put<E_KC>(q, Kl[q] * C1);
                S.template put<E_CB>(q, cc[j].cb);
                S.template put<E_INV_SS>(q, cc[j].inv_Ss);
                S.template put<E_CC>(q, cc[j].cc);
                S.template put<E_CD>(q, cc[j].cd);
                S.template put<E_AK>(q, aK_o);
                S.template put<E_AC>(q, aCo[q]);
                S.template put<E_T1>(q, fma(-dtg, src_w * sat[j], theta[j]));
                S.template put<E_T2>(q, fma(-dtg, src_e * sat[j], rho_e[j]));
                S.template put<E_IRANGE>(q, cc[j].inv_range);
            }
        }
        CLB_PCS(14)  // constants
        // rows of W22 = dtgamma d(T_rho_e)/d(rho_e) - I and their elimination, boundary -> seam
#pragma unroll
        for (int q = 0; q < Q; ++q) {
            const double dti = G.dti[half][r0 + q];
            const double aC_in = (q < Q - 1) ? aCo[q + 1] : aC8;
            o22[q] = (aCo[q] * ((q == 0) ? rc_out : rc[q - 1])) * dti;
            i22[q] = (aC_in * ((q < Q - 1) ? rc[q + 1] : rc_in)) * dti;
            d22[q] = fma(-((aC_in + aCo[q]) * rc[q]), dti, -1.0);
        }
        // Elimination boundary -> seam WITHOUT a reciprocal in the dependency chain: the leading principal
        // minors D_q = d_q D_{q-1} - (o_q i_{q-1}) D_{q-2} run as one fma per cell; the pivots' reciprocals
        // D_{q-1}/D_q are then Q independent divisions.  (|d| >= 1 and the rows are diagonally dominant, so D
        // only grows: <= |d|^16, far inside the double range.)  An inner part continues the recurrence of
        // the outer one: (D_{Q-1}, i_{Q-1} D_{Q-2}) cross lanes between the passes.
        {
            double Din = 1.0, iDin = 0.0, D[Q], iD[Q];
#pragma unroll
            for (int pass = 0; pass < PARTS; ++pass) {
                double Dp = Din, iDp = iDin;
#pragma unroll
                for (int q = 0; q < Q; ++q) {
                    iD[q] = i22[q] * Dp;
                    D[q] = fma(d22[q], Dp, -(o22[q] * iDp));
                    iDp = iD[q];
                    Dp = D[q];
                }
                if (pass + 1 < PARTS) {
                    const double rD_ = from_prev_part<Gm::PARTD>(D[Q - 1]), riD_ = from_prev_part<Gm::PARTD>(iD[Q - 1]);
                    Din = outermost ? 1.0 : rD_;
                    iDin = outermost ? 0.0 : riD_;
                }
            }
#pragma unroll
            for (int q = 0; q < Q; ++q) {
                const double rD = fm::rcp(D[q]);
                const double den = ((q == 0) ? Din : D[q - 1]) * rD;
                const double cp = iD[q] * rD;
                S.template put<E_DEN22>(q, den);
                S.template put<E_OD22>(q, o22[q] * den);
                S.template put<E_C22>(q, cp);
                if (q == Q - 1) c22_last = cp;
            }
        }
        r22 = fm::rcp(fma(-c22_last, xchg<Gm::SEAM>(c22_last), 1.0));
        CLB_PCS(15)  // W22 rows, factorisation
    }

    // The next tile's per-column scalars are fetched only now, after everything that consumes this tile's has
    // been issued: fetched at the top of the tile, their loads shared a scoreboard with `cur` and the first use of
    // `cur` waited for them (a full HBM latency per tile); here they have the whole Newton loop to land.  (The
    // next tile's boxes were requested when the tile before this one finished its Newton loop, below.)
    if (NBUF == 2) {
        const int64_t tn = tile_id + nwarps;
        if (tn < ntiles) nxt = load_col_scalars<MODEL, BCL>(P, col_clamped(tn), half);
    }

    CLB_PC(2)  // set-up
    // ---- Newton iterations -----------------------------------------------------------------
    double dx2 = 0.0;
    // right-hand side of the pending (rho_e, rho_e) solve, already scaled by the pivots' reciprocals (b2 den22)
    double b2p[Q];
#pragma unroll
    for (int q = 0; q < Q; ++q) b2p[q] = 0.0;
    // forward / backward substitution with the factors of the set-up, then the update of rho_e
    auto w22_solve = [&](const bool is_last) {
        if constexpr (MODEL == 1) {
        double g2[Q], gin = 0.0;
#pragma unroll
        for (int pass = 0; pass < PARTS; ++pass) {
            double gp = gin;
#pragma unroll
            for (int q = 0; q < Q; ++q) {
                gp = fma(-S.template get<E_OD22>(q), gp, b2p[q]);
                g2[q] = gp;
            }
            if (pass + 1 < PARTS) {
                const double rg_ = from_prev_part<Gm::PARTD>(gp);
                gin = outermost ? 0.0 : rg_;
            }
        }
        const double xs = fma(-c22_last, xchg<Gm::SEAM>(g2[Q - 1]), g2[Q - 1]) * r22;
        double xn = xs, x2[Q];
#pragma unroll
        for (int pass = 0; pass < PARTS; ++pass) {
            x2[Q - 1] = (pass == 0) ? xs : (innermost ? xs : fma(-c22_last, xn, g2[Q - 1]));
#pragma unroll
            for (int q = Q - 2; q >= 0; --q) x2[q] = fma(-S.template get<E_C22>(q), x2[q + 1], g2[q]);
            if (pass + 1 < PARTS) xn = from_next_part<Gm::PARTD>(x2[0]);
        }
#pragma unroll
        for (int q = 0; q < Q; ++q) {
            U2[q] -= x2[q];
            if (is_last) dx2 = fma(x2[q], x2[q], dx2);
        }
        }
    };
#ifndef CLB_NEWTON_UNROLL
#define CLB_NEWTON_UNROLL 1
#endif
    constexpr int kNewtonUnroll = CLB_NEWTON_UNROLL;
#pragma unroll kNewtonUnroll
    for (int it = 0; it < max_iters; ++it) {
        // cache_imp!: closures (and temperature) at the iterate, W cells at a time
        double h[Q], dps[Q], Kc[Q], Td[Q], eK[Q], rcsv[Q];
        // The (rho_e, rho_e) solve of the PREVIOUS iteration runs here, in the same basic block as the closures of
        // this one: the closures only need theta (K, kappa are lagged, so the water equation never reads rho_e), and
        // the substitution sweeps of W22 are a pure latency chain (4-deep fma chains, three lane crossings) that
        // fills the closures' issue slots instead of standing alone.  First iteration: b2 = 0, so x2 = 0 exactly.
        if (MODEL == 1) w22_solve(false);
        // one group of WW cells (WW independent dependency chains); Q = WG full groups of W and a tail
        auto closure_group = [&](auto wtag, const int g) {
            constexpr int W = decltype(wtag)::value;
            double th[W], thr[W], nue[W], ca[W], ca2[W], cb[W], iSs[W], ccc[W], cd[W], Ksat[W], irg[W];
            double K[W], psi[W], dp[W];
#pragma unroll
            for (int j = 0; j < W; ++j) {
                th[j] = U1[g + j];
                thr[j] = S.template get<0>(g + j);  // theta_r / nu(_eff) are slots 0 / 1 of both models
                nue[j] = S.template get<1>(g + j);
                ca2[j] = 0.0;
                Ksat[j] = 0.0;
            }
            if (MODEL == 1) {
                // volumetric heat capacity at the iterate (theta_l clipped to the pore space left by ice); the
                // temperature itself follows the closures, below
#pragma unroll
                for (int j = 0; j < W; ++j) rcsv[g + j] = fma(fmv::min_nn(nue[j], th[j]), C1, S.template get<E_RCBASE>(g + j));
#pragma unroll
                for (int j = 0; j < W; ++j) {
                    ca[j] = S.template get<E_CA>(g + j);
                    cb[j] = S.template get<E_CB>(g + j);
                    iSs[j] = S.template get<E_INV_SS>(g + j);
                    ccc[j] = S.template get<E_CC>(g + j);
                    cd[j] = S.template get<E_CD>(g + j);
                    irg[j] = S.template get<E_IRANGE>(g + j);
                }
            } else {
#pragma unroll
                for (int j = 0; j < W; ++j) {
                    ca[j] = S.template get<R_CA>(g + j);
                    ca2[j] = S.template get<R_CA2>(g + j);
                    cb[j] = S.template get<R_CB>(g + j);
                    iSs[j] = S.template get<R_INV_SS>(g + j);
                    ccc[j] = S.template get<R_CC>(g + j);
                    cd[j] = S.template get<R_CD>(g + j);
                    Ksat[j] = S.template get<R_KSAT>(g + j);
                    irg[j] = S.template get<R_IRANGE>(g + j);
                }
            }
            fmv::closure<CLOSURE, MODEL == 0, W, true>(MT, th, thr, nue, irg, ca, ca2, cb, ccc, cd, iSs, Ksat, K, psi, dp);
#pragma unroll
            for (int j = 0; j < W; ++j) {
                h[g + j] = psi[j] + G.z[half][r0 + g + j];
                dps[g + j] = dp[j];
                if (MODEL == 0) Kc[g + j] = K[j];
            }
                };
#pragma unroll
        for (int g = 0; g + WFULL <= Q; g += WFULL) closure_group(std::integral_constant<int, WFULL>{}, g);
        if constexpr (Q % WFULL != 0) closure_group(std::integral_constant<int, Q % WFULL>{}, Q - Q % WFULL);
        if (MODEL == 1) {
            // update_implicit_aux (energy_hydrology.jl:427-445): T from (theta_l, rho_e_int, theta_i); after the
            // closures because it reads the rho_e the lagged solve above has just updated
            double num[Q], Tq[Q];
#pragma unroll
            for (int q = 0; q < Q; ++q) num[q] = U2[q] + S.template get<E_ICE>(q);
            fmv::div<Q>(num, rcsv, Tq);
#pragma unroll
            for (int q = 0; q < Q; ++q) {
                const double T = T_ref + Tq[q];
                Td[q] = T;
                eK[q] = (T - T_ref) * S.template get<E_KC>(q);
            }
        }
        CLB_PC(3)  // closures + T
        double h_out, h_in, dps_out, dps_in, K_out = 0.0, K_in = 0.0, T_out = 0.0, T_in = 0.0, eK_out = 0.0, eK_in = 0.0;
        nb_exchange<Gm>(h[0], h[Q - 1], innermost, h_out, h_in);
        nb_exchange<Gm>(dps[0], dps[Q - 1], innermost, dps_out, dps_in);
        if (MODEL == 0) nb_exchange<Gm>(Kc[0], Kc[Q - 1], innermost, K_out, K_in);
        if (MODEL == 1) {
            nb_exchange<Gm>(Td[0], Td[Q - 1], innermost, T_out, T_in);
            nb_exchange<Gm>(eK[0], eK[Q - 1], innermost, eK_out, eK_in);
        }

        CLB_PC(4)  // neighbour exchange
        // T_imp! and Wfact: face fluxes, residuals and the rows of W11 = dtgamma dT/dtheta - I
        double o1[Q], d1[Q], i1[Q], f1[Q], f2[Q], aE[Q + 1];
        double Fw_o, Fe_o = 0.0, aK_o;
        double b0_wi = b0_w, bT_wi = bT_w, top_dflux = 0.0;  // boundary terms of this iteration
        if (BCL) {
            // update_implicit_boundary_fluxes (rre.jl:460-468): the boundary cell of this half -- slot qT of the top
            // lane, slot 0 of the bottom half's outer lane
            const int qb = half ? qT : 0;
            double Kb = 0.0, psib = 0.0, dpb = 0.0;
#pragma unroll
            for (int q = 0; q < Q; ++q)
                if (q == qb) { Kb = Kc[q]; psib = h[q] - G.z[half][r0 + q]; dpb = dps[q]; }
            if (half) {
                const double fl = fm::div_by(-Kb * ((psi_bc + P.dz_top) - psib), P.dz_top, P.inv_dz_top);
                if (top_lane) {
                    live_top = fl;
                    top_dflux = fm::div_by(Kb * dpb, P.dz_top, P.inv_dz_top);
                    if (qT == 0) b0_wi = -fl;
                    else bT_wi = -fl;
                }
            } else if (outermost) {
                if (P.bottom_bc == 1) live_bot = -1 * Kb;
                else if (P.bottom_bc == 2) live_bot = fm::div_by(-Kb * ((psib + P.dz_bot) - psi_bc), P.dz_bot, P.inv_dz_bot);
                b0_wi = live_bot;
            }
            // the flux integral (W = -I) with the fluxes of this iterate, in the lane that stores it (the bottom
            // half's outer lane, which owns bottom_bc; top_bc comes from the top lane of the column)
            const double tw_ = __shfl_sync(0xffffffffu, live_top, (lane % CPW) + CPW * (pT * 2 + 1));
            if (idx == 0) {
                const double Tiw = -(tw_ - live_bot) - ld_Rss;
                dxw_last = -(cur.intF_w + dtg * Tiw - Uiw);
                Uiw -= dxw_last;
                live_top = tw_;
            }
        }
        {
            // outer face of the first cell: the column boundary (zero coefficient, boundary flux)
            // or, for an inner part, the face shared with the previous part
            const double hid_o = G.hidzf[half][r0];
            const double dh0 = h[0] - h_out;
            aK_o = (MODEL == 1) ? S.template get<E_AK>(0) : (Kc[0] + K_out) * hid_o;
            Fw_o = b0_wi - aK_o * dh0;
            aE[0] = 0.0;
            if (MODEL == 1) {
                aE[0] = (eK[0] + eK_out) * hid_o;
                Fe_o = fma(-S.template get<E_AC>(0), Td[0] - T_out, b0_e - aE[0] * dh0);
            }
        }
#pragma unroll
        for (int q = 0; q < Q; ++q) {
            const double dti = G.dti[half][r0 + q];
            const double hid_in = G.hidzf[half][r0 + q + 1];
            const double hn = (q < Q - 1) ? h[q + 1] : h_in;
            const double dpn = (q < Q - 1) ? dps[q + 1] : dps_in;
            const double dh = hn - h[q];
            double aK_in, aC_in = 0.0;  // coefficients of the inner face
            if (MODEL == 1) {
                aK_in = (q < Q - 1) ? S.template get<E_AK>(q + 1) : aK8;
                aC_in = (q < Q - 1) ? S.template get<E_AC>(q + 1) : aC8;
            } else {
                aK_in = (Kc[q] + ((q < Q - 1) ? Kc[q + 1] : K_in)) * hid_in;
            }
            double Fw_in = -(aK_in * dh);
            if (q + 1 == qT) Fw_in += bT_wi;  // bT is zero in every lane but the top cell's
            double t1;
            if (MODEL == 0) t1 = S.template get<R_T1>(q);
            else t1 = S.template get<E_T1>(q);
            f1[q] = fma(Fw_o - Fw_in, dti, t1) - U1[q];
            o1[q] = (aK_o * ((q == 0) ? dps_out : dps[q - 1])) * dti;
            i1[q] = (aK_in * dpn) * dti;
            if (BCL) d1[q] = fma(-((aK_in + aK_o) * dps[q] + ((q == qT) ? top_dflux : 0.0)), dti, -1.0);
            else d1[q] = fma(-((aK_in + aK_o) * dps[q]), dti, -1.0);
            if (MODEL == 1) {
                const double Tn = (q < Q - 1) ? Td[q + 1] : T_in;
                const double eKn = (q < Q - 1) ? eK[q + 1] : eK_in;
                const double aE_in = (eK[q] + eKn) * hid_in;
                aE[q + 1] = aE_in;
                double Fe_in = fma(-aC_in, Tn - Td[q], -(aE_in * dh));
                if (q + 1 == qT) Fe_in += bT_e;
                f2[q] = fma(Fe_o - Fe_in, dti, S.template get<E_T2>(q)) - U2[q];
                Fe_o = Fe_in;
            }
            Fw_o = Fw_in;
            aK_o = aK_in;
        }

        CLB_PC(5)  // faces, residuals, rows
        // ldiv! of W11: twisted Thomas, eliminated boundary -> seam (once per part, the carry crossing
        // lanes in between), 2x2 seam system, back substitution seam -> boundary
        double c1[Q], g1[Q], x1[Q], y[Q];
        {
            // forward elimination by the minors' recurrence (see the W22 set-up): D_q as there,
            // gamma_q = f_q D_{q-1} - o_q gamma_{q-1}; then c_q = i_q D_{q-1} / D_q and g_q = gamma_q / D_q
            double Din = 1.0, iDin = 0.0, gin = 0.0, D[Q], iD[Q], gam[Q];
#pragma unroll
            for (int pass = 0; pass < PARTS; ++pass) {
                double Dp = Din, iDp = iDin, gp = gin;
#pragma unroll
                for (int q = 0; q < Q; ++q) {
                    iD[q] = i1[q] * Dp;
                    gam[q] = fma(-o1[q], gp, f1[q] * Dp);
                    D[q] = fma(d1[q], Dp, -(o1[q] * iDp));
                    iDp = iD[q];
                    gp = gam[q];
                    Dp = D[q];
                }
                if (pass + 1 < PARTS) {
                    const double rD_ = from_prev_part<Gm::PARTD>(D[Q - 1]), riD_ = from_prev_part<Gm::PARTD>(iD[Q - 1]),
                                 rg_ = from_prev_part<Gm::PARTD>(gam[Q - 1]);
                    Din = outermost ? 1.0 : rD_;
                    iDin = outermost ? 0.0 : riD_;
                    gin = outermost ? 0.0 : rg_;
                }
            }
#pragma unroll
            for (int q = 0; q < Q; ++q) {
                const double rD = fm::rcp(D[q]);
                c1[q] = iD[q] * rD;
                g1[q] = gam[q] * rD;
            }
            const double cs_ = xchg<Gm::SEAM>(c1[Q - 1]), gs_ = xchg<Gm::SEAM>(g1[Q - 1]);
            const double xs = fma(-c1[Q - 1], gs_, g1[Q - 1]) * fm::rcp(fma(-c1[Q - 1], cs_, 1.0));
            double xn = xs;  // value of the inner neighbour's unknown, seam value for the innermost part
#pragma unroll
            for (int pass = 0; pass < PARTS; ++pass) {
                x1[Q - 1] = (pass == 0) ? xs : (innermost ? xs : fma(-c1[Q - 1], xn, g1[Q - 1]));
#pragma unroll
                for (int q = Q - 2; q >= 0; --q) x1[q] = fma(-c1[q], x1[q + 1], g1[q]);
                if (pass + 1 < PARTS) xn = from_next_part<Gm::PARTD>(x1[0]);
            }
        }
        CLB_PC(6)  // W11 solve
        const bool last = (it == max_iters - 1);
        if (last) {
            dx2 = 0.0;
#pragma unroll
            for (int q = 0; q < Q; ++q) dx2 = fma(x1[q], x1[q], dx2);
        }
#pragma unroll
        for (int q = 0; q < Q; ++q) {
            U1[q] -= x1[q];
            y[q] = dps[q] * x1[q];
        }
        CLB_PC(7)  // update of theta
        if (MODEL == 1) {
            // ldiv!: BlockLowerTriangularSolve(theta_l): b2 = f2 - W21 x1 with
            // W21 = -dtgamma (D . Diag(interp(-e_l K)) . G . Diag(dpsi)) - I  (energy_hydrology.jl:545-556),
            // then the pre-factored W22
            double y_out, y_in;
            nb_exchange<Gm>(y[0], y[Q - 1], innermost, y_out, y_in);
#pragma unroll
            for (int q = 0; q < Q; ++q) {
                const double yn = (q < Q - 1) ? y[q + 1] : y_in;
                const double yo = (q == 0) ? y_out : y[q - 1];
                const double acc = fma(aE[q], yo - y[q], aE[q + 1] * (yn - y[q]));
                const double s = fma(G.dti[half][r0 + q], acc, -x1[q]);
                b2p[q] = (f2[q] - s) * S.template get<E_DEN22>(q);  // solved at the top of the next iteration / after the loop
            }
        }
        CLB_PC(8)  // W21 x1, W22 solve, update of rho_e
    }
    if (MODEL == 1) w22_solve(true);  // the last iteration's

    // ---- this tile's constants are dead: hand its buffer to the TMA ------------------------------------
    // BEFORE the global stores below: fence.proxy.async carries a MEMBAR.ALL.CTA, which would otherwise wait
    // for those stores to be acknowledged (~900 cycles per tile when the request sat at the top of the loop).
    {
        const int64_t tr = tile_id + NBUF * nwarps;  // double-buffered: the tile after next lands in this buffer
        if (tr < ntiles) {
            __syncwarp();
            // generic-proxy accesses of this buffer are ordered before the async-proxy writes of the TMA
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            CLB_PC(11)  // of the request: __syncwarp + proxy fence
            request_tile(tr, buf, 3);
            if (NBUF == 1) nxt = load_col_scalars<MODEL, BCL>(P, col_clamped(tr), half);
        }
    }

    CLB_PC(9)  // next tile's request
    // ---- write the new state -----------------------------------------------------------------
#pragma unroll
    for (int q = 0; q < Q; ++q) {
        const int level = level_of(q);
        if (level < N && col_ok) {
            const int64_t k = P.at(level, c);
            P.out_theta_l[k] = U1[q];
            if (MODEL == 1) P.out_rho_e[k] = U2[q];
            if (P.stats) {  // the NaN count is only taken when the caller asked for clb_stats
                if (!isfinite(U1[q])) bad += 1.0;
                if (MODEL == 1 && !isfinite(U2[q])) bad += 1.0;
            }
        }
    }
    if (BCL && idx == 0 && col_ok) {
        dx2_int = dxw_last * dxw_last;
        P.out_intF_w[c] = Uiw;
        P.top_bc_w[c] = live_top;  // the last evaluation stays in the cache, as update_implicit_boundary_fluxes leaves it
        P.bot_bc_w[c] = live_bot;
    }
    if (col_ok) dx2_acc += dx2 + dx2_int;
    if (NBUF == 2) buf ^= 1;
    else __syncwarp();
    CLB_PC(10)  // stores
    }  // tiles
    CLB_PC_FLUSH
    if (P.stats) accumulate_stats(P, dx2_acc, bad);  // only when the caller asked for clb_stats
}

}  // namespace clb
