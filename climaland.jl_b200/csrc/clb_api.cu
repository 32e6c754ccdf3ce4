// clb_api.cu -- implementation of the C ABI declared in include/climaland_b200.h.
// Host-side plumbing only: handle, device mirrors, layout transfers, kernel dispatch,
// optional NCCL reductions.  All arithmetic is in the kernels (soil_*.cuh).
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <nvtx3/nvToolsExt.h>

#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <cstdlib>
#include <new>
#include <string>
#include <vector>

#include "../../include/climaland_b200.h"
#include "layout_kernels.cuh"
#include "soil_fused.cuh"
#include "soil_hooks.cuh"
#include "soil_warp.cuh"
#include "soil_pair.cuh"
#include "soil_explicit.cuh"
#include "soil_co2.cuh"
#include "soil_co2_lanes.cuh"

namespace {

thread_local std::string g_err;

int fail(int code, const char *fmt, ...)
{
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    g_err = buf;
    return code;
}

#define CUDA_TRY(expr)                                                                               \
    do {                                                                                             \
        cudaError_t e__ = (expr);                                                                    \
        if (e__ != cudaSuccess)                                                                      \
            return fail(CLB_ERR_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e__), __FILE__, __LINE__); \
    } while (0)

#define TRY(expr)                  \
    do {                           \
        int rc__ = (expr);         \
        if (rc__ != CLB_OK) return rc__; \
    } while (0)

// ---- NCCL, loaded at run time (no link dependency; torch's or the system's copy) ----
struct NcclUniqueId {
    char internal[128];
};
typedef void *NcclComm;
struct NcclApi {
    void *lib = nullptr;
    int (*GetUniqueId)(NcclUniqueId *) = nullptr;
    int (*CommInitRank)(NcclComm *, int, NcclUniqueId, int) = nullptr;
    int (*AllReduce)(const void *, void *, size_t, int, int, NcclComm, cudaStream_t) = nullptr;
    int (*CommDestroy)(NcclComm) = nullptr;
    const char *(*GetErrorString)(int) = nullptr;
};
NcclApi g_nccl;
constexpr int kNcclFloat64 = 8, kNcclSum = 0;

int load_nccl()
{
    if (g_nccl.lib) return CLB_OK;
    const char *names[] = {"libnccl.so.2", "libnccl.so"};
    void *lib = nullptr;
    for (const char *n : names) {
        lib = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
        if (lib) break;
    }
    if (!lib) return fail(CLB_ERR_NCCL, "cannot dlopen libnccl.so.2: %s", dlerror());
    NcclApi a;
    a.lib = lib;
    a.GetUniqueId = (decltype(a.GetUniqueId))dlsym(lib, "ncclGetUniqueId");
    a.CommInitRank = (decltype(a.CommInitRank))dlsym(lib, "ncclCommInitRank");
    a.AllReduce = (decltype(a.AllReduce))dlsym(lib, "ncclAllReduce");
    a.CommDestroy = (decltype(a.CommDestroy))dlsym(lib, "ncclCommDestroy");
    a.GetErrorString = (decltype(a.GetErrorString))dlsym(lib, "ncclGetErrorString");
    if (!a.GetUniqueId || !a.CommInitRank || !a.AllReduce || !a.CommDestroy || !a.GetErrorString)
        return fail(CLB_ERR_NCCL, "libnccl is missing a required symbol");
    g_nccl = a;
    return CLB_OK;
}

#define NCCL_TRY(expr)                                                                             \
    do {                                                                                           \
        int r__ = (expr);                                                                          \
        if (r__ != 0) return fail(CLB_ERR_NCCL, "%s failed: %s", #expr, g_nccl.GetErrorString(r__)); \
    } while (0)

bool is_cell_field(int f) { return f >= 0 && f < CLB_F_NUM_CELL; }
bool is_col_field(int f) { return f >= CLB_F_NUM_CELL && f < CLB_F_NUM; }

constexpr int kBlock = 128;
constexpr int kMaxLevels = 512;
constexpr int kMaxDevices = 64;  // device ordinals with a cached per-device kernel configuration

}  // namespace

struct clb_handle_s {
    clb_config cfg;
    cudaStream_t stream = nullptr;
    int64_t ld = 0;  // ncol rounded up to 32 doubles: length of per-column mirrors
    int64_t sl = 0, sc = 0;    // strides of the per-cell mirrors: element (i, c) at i*sl + c*sc
    size_t cell_elems = 0;     // allocation size of a per-cell mirror
    bool out_of_place = false; // fused stage writes the U fields and leaves Y untouched
    int last_variant = 0;      // kernel variant the last clb_implicit_step launched
    clb::PairMaps pair_maps;   // TMA descriptors of the lane-pair kernel's inputs and the mirrors they describe
    const double *pair_map_src[14] = {};
    int pair_map_box = 0;      // columns per TMA box those descriptors were encoded for
    const double *arena_map_src = nullptr;  // the same for the two 3-D descriptors (box columns * 1000 + level rows)
    int arena_map_box = 0;
    // prepared (reciprocal) mirrors of S_s / a / b / m read by the lane-quad kernels (soil_pair.cuh: k_prepare_params)
    double *prep[4] = {};
    // The raw per-cell inputs of the lane kernels live at a uniform stride in ONE allocation, in the kernels' slot
    // order (arena_slot), so that a tile of all of them is two TMA boxes of a 3-D tensor {columns, levels, fields}
    // instead of one box per field (soil_pair.cuh: request_tile).  Column-fastest mirrors only.
    double *arena = nullptr;
    int arena_fields = 0;
    bool prep_dirty = true;     // a parameter field changed since they were written
    bool prep_volatile = false; // the caller holds a device pointer to a parameter mirror: re-prepare every stage
    // this handle's last enqueued operation wrote a time-invariant parameter mirror: the next stage kernel must
    // not start reading parameters before it completes (no programmatic dependent launch for that one)
    bool param_write_pending = true;
    double *field[CLB_F_NUM] = {};
    bool field_set[CLB_F_NUM] = {};
    // grid
    bool grid_set = false;
    std::vector<double> z_c, z_f, dz_c, inv_dz_c, inv_dz_f;
    double *d_grid = nullptr;  // z_c | dz_c | inv_dz_c | inv_dz_f, N each
    // mask compaction
    int64_t *d_idx = nullptr;
    int64_t idx_max = -1;
    // staging for host transfers
    double *d_stage = nullptr;
    size_t stage_bytes = 0;
    // scratch
    double *work[6] = {};
    double *carry = nullptr;
    double *zeros_cell = nullptr, *zeros_col = nullptr;  // stand-ins for fields the configuration does not use
    double *d_stats = nullptr;  // [0] dx^2, [1] non-finite count, [2] norm of the tolerance path, [3..6] balance
    int32_t *d_flags = nullptr; // [0] converged, [1] iterations of the tolerance path
    // pipelined host-buffer stage (clb_implicit_step_host): copy-in / copy-out streams, per-chunk events, staging
    cudaStream_t s_in = nullptr, s_out = nullptr;
    cudaEvent_t ev_start = nullptr, ev_done = nullptr, ev_in[16] = {}, ev_out[16] = {};
    double *d_stage_in = nullptr, *d_stage_out = nullptr;
    size_t stage_in_bytes = 0, stage_out_bytes = 0;
    // explicit stage of EnergyHydrology (soil_explicit.cuh)
    clb::ExplicitConst explicit_k = {};
    bool explicit_set = false;
    clb_runoff_params runoff_k = {};
    bool runoff_set = false;
    bool co2_top_state[2] = {false, false};  // SoilCO2Model: AtmosCO2StateBC / AtmosO2StateBC at the top
    int host_route = 0, host_chunks = 0, tile_boxes = 0;  // CLB_OPT_HOST_ROUTE / _HOST_CHUNKS / _TILE_BOXES
    int explicit_kernel = 0;                              // CLB_OPT_EXPLICIT_KERNEL
    int runoff_model = CLB_RUNOFF_TOPMODEL;               // CLB_OPT_RUNOFF_MODEL
    bool top_atmos = false, bottom_ewfd = false;          // CLB_OPT_TOP_ATMOS_DRIVEN, CLB_OPT_BOTTOM_EWFD
    // multi-GPU
    NcclComm comm = nullptr;
    int32_t n_ranks = 1, rank = 0;
};

namespace {

struct DeviceGuard {
    int prev = -1;
    explicit DeviceGuard(int dev)
    {
        cudaGetDevice(&prev);
        if (prev != dev) cudaSetDevice(dev);
    }
    ~DeviceGuard()
    {
        int cur = -1;
        cudaGetDevice(&cur);
        if (prev >= 0 && cur != prev) cudaSetDevice(prev);
    }
};

// slot of a field (prep = false) or of prepared mirror k (prep = true) in the arena, -1: not an arena field.  The order
// is the slot order of make_pair_maps / soil_pair.cuh.
int arena_slot(clb_handle h, int f, bool prep)
{
    if (h->cfg.model == CLB_ENERGY_HYDROLOGY) {
        if (prep) return 2 + f;
        switch (f) {
        case CLB_F_NU: return 0; case CLB_F_THETA_R: return 1; case CLB_F_Y_THETA_L: return 6;
        case CLB_F_IS_SATURATED: return 7; case CLB_F_Y_THETA_I: return 8; case CLB_F_RHO_C_DS: return 9;
        case CLB_F_K_LAG: return 10; case CLB_F_KAPPA_LAG: return 11; case CLB_F_THETA_L_LAG: return 12;
        case CLB_F_Y_RHO_E_INT: return 13; default: return -1;
        }
    }
    if (prep) return 3 + f;
    switch (f) {
    case CLB_F_NU: return 0; case CLB_F_THETA_R: return 1; case CLB_F_K_SAT: return 2; case CLB_F_Y_THETA_L: return 7;
    case CLB_F_IS_SATURATED: return 8; default: return -1;
    }
}

int arena_pointer(clb_handle h, int slot, double **out)
{
    *out = nullptr;
    if (slot < 0 || h->sc != 1) return CLB_OK;
    if (!h->arena) {
        h->arena_fields = (h->cfg.model == CLB_ENERGY_HYDROLOGY) ? 14 : 9;
        const size_t bytes = (size_t)h->arena_fields * h->cell_elems * sizeof(double);
        CUDA_TRY(cudaMalloc(&h->arena, bytes));
        CUDA_TRY(cudaMemsetAsync(h->arena, 0, bytes, h->stream));
    }
    *out = h->arena + (size_t)slot * h->cell_elems;
    return CLB_OK;
}

inline bool in_arena(clb_handle h, const double *p)
{
    return h->arena && p >= h->arena && p < h->arena + (size_t)h->arena_fields * h->cell_elems;
}

int ensure_field(clb_handle h, int f)
{
    if (h->field[f]) return CLB_OK;
    const size_t n = is_cell_field(f) ? h->cell_elems : (size_t)h->ld;
    if (is_cell_field(f)) {
        TRY(arena_pointer(h, arena_slot(h, f, false), &h->field[f]));
        if (h->field[f]) return CLB_OK;  // zeroed with the arena
    }
    CUDA_TRY(cudaMalloc(&h->field[f], n * sizeof(double)));
    CUDA_TRY(cudaMemsetAsync(h->field[f], 0, n * sizeof(double), h->stream));
    return CLB_OK;
}

int ensure_work(clb_handle h, int count)
{
    const size_t n = h->cell_elems;
    for (int w = 0; w < count; ++w)
        if (!h->work[w]) CUDA_TRY(cudaMalloc(&h->work[w], n * sizeof(double)));
    if (!h->carry) CUDA_TRY(cudaMalloc(&h->carry, 4 * (size_t)h->ld * sizeof(double)));
    return CLB_OK;
}

int ensure_stage(clb_handle h, size_t bytes)
{
    if (h->stage_bytes >= bytes) return CLB_OK;
    if (h->d_stage) {
        CUDA_TRY(cudaStreamSynchronize(h->stream));
        CUDA_TRY(cudaFree(h->d_stage));
        h->d_stage = nullptr;
        h->stage_bytes = 0;
    }
    CUDA_TRY(cudaMalloc(&h->d_stage, bytes));
    h->stage_bytes = bytes;
    return CLB_OK;
}

int require(clb_handle h, std::initializer_list<int> fields, const char *who)
{
    for (int f : fields)
        if (!h->field_set[f]) return fail(CLB_ERR_UNSET, "%s: field %d was never set", who, f);
    return CLB_OK;
}

int alloc_fields(clb_handle h, std::initializer_list<int> fields)
{
    for (int f : fields) {
        TRY(ensure_field(h, f));
        h->field_set[f] = true;
    }
    return CLB_OK;
}

clb::DevView make_view(clb_handle h)
{
    clb::DevView P;
    std::memset(&P, 0, sizeof P);
    const clb_config &c = h->cfg;
    P.model = c.model; P.closure = c.closure; P.top_bc = c.top_bc; P.bottom_bc = c.bottom_bc;
    P.topmodel = c.has_topmodel_source;
    P.N = c.n_levels; P.ncol = c.n_columns; P.ld = h->ld;
    P.sl = h->sl; P.sc = h->sc;
    P.earth = {c.rho_l, c.rho_i, c.cp_l, c.cp_i, c.T_ref, c.LH_f0};
    const int N = c.n_levels;
    P.z_c = h->d_grid; P.dz_c = h->d_grid + N; P.inv_dz_c = h->d_grid + 2 * N; P.inv_dz_f = h->d_grid + 3 * N;
    if (h->grid_set) {
        P.dz_top = h->dz_c[N - 1] / 2.0;
        P.dz_bot = h->dz_c[0] / 2.0;
        P.inv_dz_top = 1.0 / P.dz_top;
        P.inv_dz_bot = 1.0 / P.dz_bot;
    }
    double *const *F = h->field;
    P.nu = F[CLB_F_NU]; P.theta_r = F[CLB_F_THETA_R]; P.K_sat = F[CLB_F_K_SAT]; P.S_s = F[CLB_F_S_S];
    P.hcm_a = F[CLB_F_HCM_A]; P.hcm_b = F[CLB_F_HCM_B]; P.hcm_m = F[CLB_F_HCM_M]; P.rho_c_ds = F[CLB_F_RHO_C_DS];
    P.K_lag = F[CLB_F_K_LAG]; P.kappa_lag = F[CLB_F_KAPPA_LAG]; P.theta_l_lag = F[CLB_F_THETA_L_LAG];
    P.is_sat = F[CLB_F_IS_SATURATED];
    P.R_ss = F[CLB_F_R_SS]; P.R_ess = F[CLB_F_R_ESS]; P.h_grad = F[CLB_F_H_GRAD];
    if (!c.has_topmodel_source) {  // the lane-per-cell kernel loads these unconditionally
        P.is_sat = h->zeros_cell;
        // the arena's own (never written, zero) slot keeps the lane kernels' inputs equally spaced: two TMA boxes per tile
        if (h->arena && !F[CLB_F_IS_SATURATED])
            P.is_sat = h->arena + (size_t)arena_slot(h, CLB_F_IS_SATURATED, false) * h->cell_elems;
        P.R_ss = P.R_ess = P.h_grad = h->zeros_col;
    } else if (!P.R_ess) {
        P.R_ess = h->zeros_col;
    }
    P.theta_bc_top = F[CLB_F_THETA_BC_TOP]; P.theta_bc_bot = F[CLB_F_THETA_BC_BOT];
    P.Y_theta_l = F[CLB_F_Y_THETA_L]; P.Y_rho_e = F[CLB_F_Y_RHO_E_INT]; P.Y_theta_i = F[CLB_F_Y_THETA_I];
    P.Y_intF_w = F[CLB_F_Y_INTF_W]; P.Y_intF_e = F[CLB_F_Y_INTF_E];
    if (h->out_of_place) {
        P.out_theta_l = F[CLB_F_U_THETA_L]; P.out_rho_e = F[CLB_F_U_RHO_E_INT];
        P.out_intF_w = F[CLB_F_U_INTF_W]; P.out_intF_e = F[CLB_F_U_INTF_E];
    } else {
        P.out_theta_l = P.Y_theta_l; P.out_rho_e = P.Y_rho_e; P.out_intF_w = P.Y_intF_w; P.out_intF_e = P.Y_intF_e;
    }
    P.p_K = F[CLB_F_P_K]; P.p_psi = F[CLB_F_P_PSI]; P.p_T = F[CLB_F_P_T];
    P.top_bc_w = F[CLB_F_TOP_BC_W]; P.bot_bc_w = F[CLB_F_BOT_BC_W];
    P.top_bc_h = F[CLB_F_TOP_BC_H]; P.bot_bc_h = F[CLB_F_BOT_BC_H];
    P.dfluxBCdY = F[CLB_F_DFLUXBCDY]; P.total_water = F[CLB_F_TOTAL_WATER];
    P.dY_theta_l = F[CLB_F_DY_THETA_L]; P.dY_rho_e = F[CLB_F_DY_RHO_E_INT]; P.dY_theta_i = F[CLB_F_DY_THETA_I];
    P.dY_intF_w = F[CLB_F_DY_INTF_W]; P.dY_intF_e = F[CLB_F_DY_INTF_E];
    P.w11_lo = F[CLB_F_W11_LO]; P.w11_di = F[CLB_F_W11_DI]; P.w11_up = F[CLB_F_W11_UP];
    P.w21_lo = F[CLB_F_W21_LO]; P.w21_di = F[CLB_F_W21_DI]; P.w21_up = F[CLB_F_W21_UP];
    P.w22_lo = F[CLB_F_W22_LO]; P.w22_di = F[CLB_F_W22_DI]; P.w22_up = F[CLB_F_W22_UP];
    P.b_theta_l = F[CLB_F_B_THETA_L]; P.b_rho_e = F[CLB_F_B_RHO_E_INT]; P.b_theta_i = F[CLB_F_B_THETA_I];
    P.b_intF_w = F[CLB_F_B_INTF_W]; P.b_intF_e = F[CLB_F_B_INTF_E];
    P.x_theta_l = F[CLB_F_X_THETA_L]; P.x_rho_e = F[CLB_F_X_RHO_E_INT]; P.x_theta_i = F[CLB_F_X_THETA_I];
    P.x_intF_w = F[CLB_F_X_INTF_W]; P.x_intF_e = F[CLB_F_X_INTF_E];
    for (int w = 0; w < 6; ++w) P.work[w] = h->work[w];
    P.carry = h->carry;
    P.stats = h->d_stats;
    P.converged = nullptr;
    return P;
}

inline unsigned grid_for(int64_t ncol) { return (unsigned)((ncol + kBlock - 1) / kBlock); }

// closure x math dispatch for kernels templated <CLOSURE, MATH>
#define DISPATCH_CM(h, KERNEL, grid, ...)                                                               \
    do {                                                                                                \
        const int cl__ = (h)->cfg.closure, ma__ = (h)->cfg.math_mode;                                   \
        if (cl__ == 0 && ma__ == 0) KERNEL<0, 0><<<grid, kBlock, 0, (h)->stream>>>(__VA_ARGS__);         \
        else if (cl__ == 0 && ma__ == 1) KERNEL<0, 1><<<grid, kBlock, 0, (h)->stream>>>(__VA_ARGS__);    \
        else if (cl__ == 1 && ma__ == 0) KERNEL<1, 0><<<grid, kBlock, 0, (h)->stream>>>(__VA_ARGS__);    \
        else KERNEL<1, 1><<<grid, kBlock, 0, (h)->stream>>>(__VA_ARGS__);                                \
    } while (0)

#define DISPATCH_CMN(h, KERNEL, NS, grid, ...)                                                             \
    do {                                                                                                   \
        const int cl__ = (h)->cfg.closure, ma__ = (h)->cfg.math_mode;                                      \
        if (cl__ == 0 && ma__ == 0) KERNEL<0, 0, NS><<<grid, kBlock, 0, (h)->stream>>>(__VA_ARGS__);        \
        else if (cl__ == 0 && ma__ == 1) KERNEL<0, 1, NS><<<grid, kBlock, 0, (h)->stream>>>(__VA_ARGS__);   \
        else if (cl__ == 1 && ma__ == 0) KERNEL<1, 0, NS><<<grid, kBlock, 0, (h)->stream>>>(__VA_ARGS__);   \
        else KERNEL<1, 1, NS><<<grid, kBlock, 0, (h)->stream>>>(__VA_ARGS__);                               \
    } while (0)

#define DISPATCH_CMN2(h, KERNEL, A, B, grid, ...)                                                             \
    do {                                                                                                      \
        const int cl__ = (h)->cfg.closure, ma__ = (h)->cfg.math_mode;                                         \
        if (cl__ == 0 && ma__ == 0) KERNEL<0, 0, A, B><<<grid, kBlock, 0, (h)->stream>>>(__VA_ARGS__);         \
        else if (cl__ == 0 && ma__ == 1) KERNEL<0, 1, A, B><<<grid, kBlock, 0, (h)->stream>>>(__VA_ARGS__);    \
        else if (cl__ == 1 && ma__ == 0) KERNEL<1, 0, A, B><<<grid, kBlock, 0, (h)->stream>>>(__VA_ARGS__);    \
        else KERNEL<1, 1, A, B><<<grid, kBlock, 0, (h)->stream>>>(__VA_ARGS__);                                \
    } while (0)

template <int NS>
clb::GridConst<NS> make_grid_const(clb_handle h)
{
    clb::GridConst<NS> g;
    for (int i = 0; i < NS; ++i) {
        g.z_c[i] = h->z_c[i];
        g.inv_dz_c[i] = h->inv_dz_c[i];
        g.inv_dz_f[i] = h->inv_dz_f[i];
    }
    return g;
}

// Grid constants of the lane-quad kernel in its lane-local orientation (soil_pair.cuh):
// bottom half slot q = level q, top half slot q = level 15 - q (pads for levels >= N).
// Grid constants of the lane kernels in the lane-local orientation: half 0 holds levels 0 .. HR-1 bottom -> seam,
// half 1 holds levels NR-1 .. HR top -> seam (NR = 2 HR rows, the rows >= N are pads).
template <int HR>
clb::PairGridT<HR> make_pair_grid(clb_handle h, double dtg)
{
    const int N = h->cfg.n_levels;
    constexpr int NR = 2 * HR;
    clb::PairGridT<HR> g;
    g.col0 = 0;
    g.nlev = N;
    for (int half = 0; half < 2; ++half) {
        for (int q = 0; q < HR; ++q) {
            const int level = half ? NR - 1 - q : q;
            const bool real = level < N;
            g.z[half][q] = real ? h->z_c[level] : 0.0;
            g.dti[half][q] = real ? dtg * h->inv_dz_c[level] : 0.0;
        }
        for (int f = 0; f <= HR; ++f) {
            // face f lies between slots f-1 and f; inv_dz_f[i] is the face between levels i-1 and i
            double v = 0.0;
            if (f == HR) v = h->inv_dz_f[HR];
            else if (!half && f >= 1) v = h->inv_dz_f[f];
            else if (half && f >= 1 && NR - f <= N - 1) v = h->inv_dz_f[NR - f];
            g.hidzf[half][f] = v / 2.0;
        }
    }
    return g;
}

// cuTensorMapEncodeTiled through the runtime's driver entry point (no link-time libcuda dependency)
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                  const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
int encode_tiled_fn(EncodeTiledFn *out)
{
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void *p = nullptr;
        cudaDriverEntryPointQueryResult q;
        CUDA_TRY(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q));
        if (!p || q != cudaDriverEntryPointSuccess)
            return fail(CLB_ERR_CUDA, "cuTensorMapEncodeTiled is not available in this driver");
        fn = (EncodeTiledFn)p;
    }
    *out = fn;
    return CLB_OK;
}

// TMA descriptor of one column-fastest per-cell mirror: dims {ncol, N} (columns beyond ncol read as 0),
// row pitch ld doubles, box {columns of one warp, N levels}.
int encode_field_map(clb_handle h, const double *ptr, int box_columns, int box_levels_lf, CUtensorMap *map)
{
    EncodeTiledFn enc;
    TRY(encode_tiled_fn(&enc));
    const cuuint32_t estr[2] = {1, 1};
    CUresult r;
    if (h->sc == 1) {  // column-fastest mirror: dims {columns, levels}, box {box_columns, N}
        const cuuint64_t dims[2] = {(cuuint64_t)h->cfg.n_columns, (cuuint64_t)h->cfg.n_levels};
        const cuuint64_t strides[1] = {(cuuint64_t)h->ld * sizeof(double)};
        const cuuint32_t box[2] = {(cuuint32_t)box_columns, (cuuint32_t)h->cfg.n_levels};
        r = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, (void *)ptr, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    } else {  // level-fastest mirror: dims {levels, columns}, box {box_levels_lf >= N (the rest is zero-filled), box_columns}
        const cuuint64_t dims[2] = {(cuuint64_t)h->cfg.n_levels, (cuuint64_t)h->cfg.n_columns};
        const cuuint64_t strides[1] = {(cuuint64_t)h->cfg.n_levels * sizeof(double)};
        const cuuint32_t box[2] = {(cuuint32_t)box_levels_lf, (cuuint32_t)box_columns};
        r = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, (void *)ptr, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    }
    if (r != CUDA_SUCCESS) return fail(CLB_ERR_CUDA, "cuTensorMapEncodeTiled failed (%d)", (int)r);
    return CLB_OK;
}

inline bool is_closure_param(int field)
{
    return field == CLB_F_S_S || field == CLB_F_HCM_A || field == CLB_F_HCM_B || field == CLB_F_HCM_M;
}
// fields the lane-quad kernels fetch before griddepcontrol.wait (soil_pair.cuh: is_param)
inline bool is_invariant_param(int field)
{
    return is_closure_param(field) || field == CLB_F_NU || field == CLB_F_THETA_R || field == CLB_F_K_SAT ||
           field == CLB_F_RHO_C_DS;
}

// (Re)writes the prepared parameter mirrors when a parameter changed since the last stage.
int ensure_prepared(clb_handle h, const clb::DevView &P)
{
    if (!h->prep_dirty && !h->prep_volatile) return CLB_OK;
    for (int k = 0; k < 4; ++k) {
        double *&p = h->prep[k];
        if (!p) TRY(arena_pointer(h, arena_slot(h, k, true), &p));
        if (!p) CUDA_TRY(cudaMalloc(&p, h->cell_elems * sizeof(double)));
    }
    const int64_t n = (int64_t)h->cell_elems;
    const unsigned grid = (unsigned)((n + 255) / 256);
    if (h->cfg.closure == CLB_VAN_GENUCHTEN)
        clb::k_prepare_params<0><<<grid, 256, 0, h->stream>>>(P.S_s, P.hcm_a, P.hcm_b, P.hcm_m, h->prep[0], h->prep[1],
                                                              h->prep[2], h->prep[3], n);
    else
        clb::k_prepare_params<1><<<grid, 256, 0, h->stream>>>(P.S_s, P.hcm_a, P.hcm_b, nullptr, h->prep[0], h->prep[1],
                                                              h->prep[2], h->prep[3], n);
    CUDA_TRY(cudaGetLastError());
    h->prep_dirty = false;
    h->param_write_pending = true;
    return CLB_OK;
}

// Fields of the lane-quad kernel in the order of its shared-memory slots (soil_pair.cuh); the closure
// parameters S_s / a / b / m come from the prepared mirrors
// 3-D tensor {columns, levels, fields} over `nf` mirrors `stride` doubles apart, box {box_columns, box_levels, box_fields}
int encode_arena_map(clb_handle h, const double *ptr, size_t stride, int nf, int box_columns, int box_levels, int box_fields,
                     CUtensorMap *map)
{
    EncodeTiledFn enc;
    TRY(encode_tiled_fn(&enc));
    const cuuint32_t estr[3] = {1, 1, 1};
    const cuuint64_t dims[3] = {(cuuint64_t)h->cfg.n_columns, (cuuint64_t)h->cfg.n_levels, (cuuint64_t)nf};
    const cuuint64_t strides[2] = {(cuuint64_t)h->ld * sizeof(double), (cuuint64_t)stride * sizeof(double)};
    const cuuint32_t box[3] = {(cuuint32_t)box_columns, (cuuint32_t)box_levels, (cuuint32_t)box_fields};
    const CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 3, (void *)ptr, dims, strides, box, estr,
                           CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                           CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(CLB_ERR_CUDA, "cuTensorMapEncodeTiled (3-D) failed (%d)", (int)r);
    return CLB_OK;
}

// box_levels_arena > 0: the kernel takes the two-box form when the mirrors are equally spaced (the handle's arena)
int make_pair_maps(clb_handle h, const clb::DevView &P, int box_columns, int box_levels_lf, clb::PairMaps *maps,
                   int box_levels_arena = 0)
{
    TRY(ensure_prepared(h, P));
    const double *const *Q = h->prep;
    const double *eh[14] = {P.nu, P.theta_r, Q[0], Q[1], Q[2], Q[3], P.Y_theta_l, P.is_sat, P.Y_theta_i,
                            P.rho_c_ds, P.K_lag, P.kappa_lag, P.theta_l_lag, P.Y_rho_e};
    const double *ri[9] = {P.nu, P.theta_r, P.K_sat, Q[0], Q[1], Q[2], Q[3], P.Y_theta_l, P.is_sat};
    const bool is_eh = h->cfg.model == CLB_ENERGY_HYDROLOGY;
    const int n = is_eh ? 14 : 9;
    const double **src = is_eh ? eh : ri;
    if (h->pair_map_box != box_columns) std::fill(h->pair_map_src, h->pair_map_src + 14, nullptr);
    h->pair_map_box = box_columns;
    for (int j = 0; j < n; ++j) {
        if (h->pair_map_src[j] != src[j]) {  // descriptors are cached per handle; mirrors rarely move
            if (src[j]) TRY(encode_field_map(h, src[j], box_columns, box_levels_lf, &h->pair_maps.m[j]));
            h->pair_map_src[j] = src[j];
        }
    }
    bool spaced = box_levels_arena > 0 && h->sc == 1 && h->tile_boxes == 0;
    for (int j = 1; j < n && spaced; ++j) spaced = src[j] == src[0] + (size_t)j * h->cell_elems;
    if (spaced && (h->arena_map_src != src[0] || h->arena_map_box != box_columns * 1000 + box_levels_arena)) {
        // the parameter fields come first in the box order of Richards (slots 0-6); EnergyHydrology's rho_c_ds (slot 9)
        // rides with the stage inputs
        const int fa = is_eh ? 6 : 7;
        TRY(encode_arena_map(h, src[0], h->cell_elems, n, box_columns, box_levels_arena, fa, &h->pair_maps.a));
        TRY(encode_arena_map(h, src[0], h->cell_elems, n, box_columns, box_levels_arena, n - fa, &h->pair_maps.b));
        h->pair_maps.fa = fa;
        h->arena_map_src = src[0];
        h->arena_map_box = box_columns * 1000 + box_levels_arena;
    }
    h->pair_maps.arena = spaced ? 1 : 0;
    *maps = h->pair_maps;
    return CLB_OK;
}

// Lane kernels (soil_pair.cuh), launch shapes:
//   quad, PIPELINED   4 lanes x 4 cells; persistent, 1 block of 256 threads per SM, each warp double-buffers its tiles:
//                     2 x 14 slots x 1 KB of shared memory per warp (4 of EnergyHydrology's 18 stage constants stay in
//                     registers); 8 x 28 KB of tiles + the 2.5 KB of tables fill the 227 KB
//   quad, plain       one tile per warp, all constants in shared memory (18 KB per warp), 3 blocks of 128 per SM
//   octet, N <= 16    8 lanes x 2 cells; persistent, 1 block of 512 threads per SM (4 warps per sub-partition at <= 128
//                     registers), double-buffered tiles of 4 columns: 2 x 14 slots x 512 B per warp
//   octet, N = 50     8 lanes x 7 cells (56 level rows); persistent, single-buffered tiles of 4 columns (1.75 KB per
//                     slot) with an L2 prefetch of the next tile; 8 (Richards) or 6 (EnergyHydrology) warps per SM
template <int CLOSURE, int MODEL, int N, int PARTS, int Q, int NS, int NBUF, int BLOCK, int MINB, bool PERSISTENT,
          bool LF = false, bool BCL = false>
int launch_lanes(clb_handle h, const clb::DevView &P, double dtg, int max_iters, int64_t col0)
{
    using Gm = clb::LaneGeom<PARTS, Q>;
    constexpr int CPW = Gm::CPW;
    auto kern = clb::k_step_lanes<CLOSURE, MODEL, N, PARTS, NS, NBUF, BLOCK, MINB, Q, LF, BCL>;
    const size_t smem = clb::pair_smem_bytes<PARTS, NS, NBUF, BLOCK, Q, LF>();
    // Function attributes are per DEVICE (and per instantiation): one flag per device ordinal, so that a process
    // holding handles on several GPUs opts every one of them in to the > 48 KB of dynamic shared memory.
    static std::atomic<bool> configured[kMaxDevices];
    const int dev = h->cfg.device;
    if (dev < 0 || dev >= kMaxDevices || !configured[dev].load(std::memory_order_acquire)) {
        CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
        if (dev >= 0 && dev < kMaxDevices) configured[dev].store(true, std::memory_order_release);
    }
    const int64_t tiles = (P.ncol + CPW - 1) / CPW;
    int64_t blocks = (tiles * 32 + BLOCK - 1) / BLOCK;
    if (PERSISTENT) {
        int sms = 0;
        CUDA_TRY(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, h->cfg.device));
        blocks = std::min<int64_t>(blocks, (int64_t)sms * MINB);
    }
    clb::PairGridT<Gm::HR> g = make_pair_grid<Gm::HR>(h, dtg);
    g.col0 = (int)col0;
    clb::PairMaps maps;
    // the descriptors always describe the whole mirrors (P may be a column sub-range with shifted pointers;
    // its offset travels as g.col0)
    TRY(make_pair_maps(h, col0 ? make_view(h) : P, CPW, Gm::kRowsLF, &maps, (LF || NBUF != 2) ? 0 : Gm::NR));
    // programmatic dependent launch: the kernel's prologue (tables, barriers, the first tile's parameter
    // fields) may overlap the tail of the stream's previous kernel unless that kernel may be writing this
    // handle's parameter mirrors
    cudaLaunchConfig_t lc = {};
    lc.gridDim = dim3((unsigned)blocks);
    lc.blockDim = dim3(BLOCK);
    lc.dynamicSmemBytes = smem;
    lc.stream = h->stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    lc.attrs = attr;
    lc.numAttrs = (h->param_write_pending || h->prep_volatile) ? 0 : 1;
    CUDA_TRY(cudaLaunchKernelEx(&lc, kern, P, g, maps, dtg, (int)max_iters));
    h->param_write_pending = false;
    return CLB_OK;
}

#ifndef CLB_QUAD_MINB
#define CLB_QUAD_MINB 3
#endif
#ifndef CLB_QUAD_NS
#define CLB_QUAD_NS 18
#endif
template <int CLOSURE, int MODEL, int N, bool PIPELINED>
int launch_quad(clb_handle h, const clb::DevView &P, double dtg, int max_iters, int64_t col0)
{
#ifndef CLB_QUAD_NBUF
#define CLB_QUAD_NBUF 2
#endif
#ifndef CLB_QUAD_NS_R
#define CLB_QUAD_NS_R 11  // RichardsModel: all 11 stage constants in shared memory (9 = the raw fields' slots only, 2 in registers)
#endif
    constexpr int NS = (MODEL == 1) ? ((PIPELINED && CLB_QUAD_NBUF == 2) ? 14 : CLB_QUAD_NS) : (PIPELINED ? CLB_QUAD_NS_R : 11);
    if constexpr (PIPELINED)
#ifndef CLB_QUAD_BLOCK_R
#define CLB_QUAD_BLOCK_R 256
#endif
#ifndef CLB_QUAD_BLOCK
#define CLB_QUAD_BLOCK 256  // 8 warps per SM; 128 = one warp per sub-partition (latency experiment)
#endif
    {
        // RichardsModel with a MoistureStateBC top: the boundary fluxes follow the iterate (BCL instantiation)
        if constexpr (MODEL == 0) {
            if (h->cfg.top_bc == CLB_TOP_MOISTURE_STATE)
                return launch_lanes<CLOSURE, MODEL, N, 2, 4, NS, CLB_QUAD_NBUF, CLB_QUAD_BLOCK_R, 1, true, false, true>(h, P, dtg, max_iters, col0);
        }
        return launch_lanes<CLOSURE, MODEL, N, 2, 4, NS, CLB_QUAD_NBUF, (MODEL == 0) ? CLB_QUAD_BLOCK_R : CLB_QUAD_BLOCK, 1, true>(h, P, dtg, max_iters,
                                                                                                                col0);
    }
    else
        return launch_lanes<CLOSURE, MODEL, N, 2, 4, NS, 1, 128, CLB_QUAD_MINB, false>(h, P, dtg, max_iters, col0);
}

template <int N, bool PIPELINED>
int launch_quad_n(clb_handle h, const clb::DevView &P, double dtg, int max_iters, int64_t col0 = 0)
{
    const bool eh = h->cfg.model == CLB_ENERGY_HYDROLOGY;
    const bool vg = h->cfg.closure == CLB_VAN_GENUCHTEN;
    if (eh)
        return vg ? launch_quad<0, 1, N, PIPELINED>(h, P, dtg, max_iters, col0)
                  : launch_quad<1, 1, N, PIPELINED>(h, P, dtg, max_iters, col0);
    return vg ? launch_quad<0, 0, N, PIPELINED>(h, P, dtg, max_iters, col0)
              : launch_quad<1, 0, N, PIPELINED>(h, P, dtg, max_iters, col0);
}

#ifndef CLB_OCTET_BLOCK_EH1
// threads per block of the single-buffered EnergyHydrology octets on column-fastest mirrors with up to 7 cells per lane
// (18 slots of 56 rows x 4 columns per warp: seven warps fill the 227 KB; the level-fastest tiles' padded rows and
// Q = 8 leave room for six): N = 49 / 56 at 1e5 columns 590 / 619 -> 559 / 585 us
#define CLB_OCTET_BLOCK_EH1 224
#endif
template <int CLOSURE, int MODEL, int N>
int launch_octet(clb_handle h, const clb::DevView &P, double dtg, int max_iters)
{
    if constexpr (N <= 16) {
        constexpr int NS = (MODEL == 1) ? 14 : 11;
        return launch_lanes<CLOSURE, MODEL, N, 4, 2, NS, 2, 512, 1, true>(h, P, dtg, max_iters, 0);
    } else if constexpr (MODEL == 1) {
        if (h->sc != 1) return launch_lanes<CLOSURE, MODEL, N, 4, 7, 18, 1, 192, 1, true, true>(h, P, dtg, max_iters, 0);
        return launch_lanes<CLOSURE, MODEL, N, 4, 7, 18, 1, CLB_OCTET_BLOCK_EH1, 1, true>(h, P, dtg, max_iters, 0);
    } else {
        if (h->sc != 1) return launch_lanes<CLOSURE, MODEL, N, 4, 7, 11, 1, 256, 1, true, true>(h, P, dtg, max_iters, 0);
        return launch_lanes<CLOSURE, MODEL, N, 4, 7, 11, 1, 256, 1, true>(h, P, dtg, max_iters, 0);
    }
}

// Octet with the level count taken at run time (template N = 0): Q cells per lane, NR = 8 Q level rows, for
// NR - 8 < N <= NR; persistent tiles of 4 columns cut from column-fastest mirrors.
template <int CLOSURE, int MODEL, int Q>
int launch_octet_rt(clb_handle h, const clb::DevView &P, double dtg, int max_iters)
{
    // double-buffered (and then two TMA boxes of the arena per tile) where two tiles per warp fit the 227 KB and, for
    // EnergyHydrology, 4 of the 18 slots fit the registers: N <= 40 (Richards), N <= 32 (EnergyHydrology).  Measured at
    // 1e5 columns against the single-buffered form: Richards N = 20 / 30 / 40: 95 / 122 / 133 us against 109 / 141 /
    // 182 us, EnergyHydrology N = 20: 190 against 244 us
    if constexpr (MODEL == 1 && Q <= 4) return launch_lanes<CLOSURE, MODEL, 0, 4, Q, 14, 2, 192, 1, true>(h, P, dtg, max_iters, 0);
    else if constexpr (MODEL == 0 && Q <= 5) return launch_lanes<CLOSURE, MODEL, 0, 4, Q, 11, 2, 256, 1, true>(h, P, dtg, max_iters, 0);
    else if constexpr (MODEL == 1) return launch_lanes<CLOSURE, MODEL, 0, 4, Q, 18, 1, (Q <= 7) ? CLB_OCTET_BLOCK_EH1 : 192, 1, true>(h, P, dtg, max_iters, 0);
    else return launch_lanes<CLOSURE, MODEL, 0, 4, Q, 11, 1, 256, 1, true>(h, P, dtg, max_iters, 0);
}

template <int Q>
int launch_octet_rt_q(clb_handle h, const clb::DevView &P, double dtg, int max_iters)
{
    const bool eh = h->cfg.model == CLB_ENERGY_HYDROLOGY;
    const bool vg = h->cfg.closure == CLB_VAN_GENUCHTEN;
    if (eh) return vg ? launch_octet_rt<0, 1, Q>(h, P, dtg, max_iters) : launch_octet_rt<1, 1, Q>(h, P, dtg, max_iters);
    return vg ? launch_octet_rt<0, 0, Q>(h, P, dtg, max_iters) : launch_octet_rt<1, 0, Q>(h, P, dtg, max_iters);
}

template <int N>
int launch_octet_n(clb_handle h, const clb::DevView &P, double dtg, int max_iters)
{
    const bool eh = h->cfg.model == CLB_ENERGY_HYDROLOGY;
    const bool vg = h->cfg.closure == CLB_VAN_GENUCHTEN;
    if (eh) return vg ? launch_octet<0, 1, N>(h, P, dtg, max_iters) : launch_octet<1, 1, N>(h, P, dtg, max_iters);
    return vg ? launch_octet<0, 0, N>(h, P, dtg, max_iters) : launch_octet<1, 0, N>(h, P, dtg, max_iters);
}

bool pair_variant_applies(clb_handle h, bool octet = false)
{
    const int N = h->cfg.n_levels;
    // octet: N = 15 / 16 / 50 as template instantiations, 17 .. 64 with the level count at run time
    if (N != 15 && N != 16 && !(octet && N >= 17 && N <= 64)) return false;
    if (h->cfg.math_mode != CLB_MATH_FAST) return false;
    // a MoistureStateBC top re-evaluates the boundary fluxes every iteration (rre.jl:460-468): the pipelined quad has an
    // instantiation for it (BCL); the octets do not, nor does the plain quad: lane-per-cell kernel
    if (h->cfg.model == CLB_RICHARDS && h->cfg.top_bc == 1 &&
        (octet || h->cfg.kernel_variant == CLB_VARIANT_LANE_QUAD || h->cfg.kernel_variant == CLB_VARIANT_LANE_OCTET))
        return false;
    // the TMA boxes of the N = 15 / 16 kernels are cut from column-fastest mirrors; the N = 50 octet also reads
    // level-fastest ones (a tile is then one contiguous piece of each field)
    if (h->cfg.layout == CLB_LAYOUT_LEVEL_FASTEST && !(octet && N == 50)) return false;
    return true;
}

int step_inputs_ready(clb_handle h)
{
    if (!h->grid_set) return fail(CLB_ERR_UNSET, "clb_set_grid was never called");
    TRY(require(h, {CLB_F_NU, CLB_F_THETA_R, CLB_F_K_SAT, CLB_F_S_S, CLB_F_HCM_A, CLB_F_HCM_B, CLB_F_Y_THETA_L},
                "implicit path"));
    if (h->cfg.closure == CLB_VAN_GENUCHTEN) TRY(require(h, {CLB_F_HCM_M}, "van Genuchten closure"));
    if (h->cfg.model == CLB_ENERGY_HYDROLOGY)
        TRY(require(h, {CLB_F_RHO_C_DS, CLB_F_K_LAG, CLB_F_KAPPA_LAG, CLB_F_THETA_L_LAG, CLB_F_Y_RHO_E_INT,
                        CLB_F_Y_THETA_I},
                    "EnergyHydrology"));
    if (h->cfg.has_topmodel_source) {
        TRY(require(h, {CLB_F_IS_SATURATED, CLB_F_R_SS, CLB_F_H_GRAD}, "TOPMODEL source"));
        if (h->cfg.model == CLB_ENERGY_HYDROLOGY) TRY(require(h, {CLB_F_R_ESS}, "TOPMODEL source"));
    }
    if (h->cfg.top_bc == CLB_TOP_MOISTURE_STATE) TRY(require(h, {CLB_F_THETA_BC_TOP}, "MoistureStateBC top"));
    if (h->cfg.bottom_bc == CLB_BOT_MOISTURE_STATE) TRY(require(h, {CLB_F_THETA_BC_BOT}, "MoistureStateBC bottom"));
    // boundary fluxes and flux integrals default to zero when the host never set them
    TRY(alloc_fields(h, {CLB_F_TOP_BC_W, CLB_F_BOT_BC_W, CLB_F_Y_INTF_W}));
    if (h->cfg.model == CLB_ENERGY_HYDROLOGY) TRY(alloc_fields(h, {CLB_F_TOP_BC_H, CLB_F_BOT_BC_H, CLB_F_Y_INTF_E}));
    if (h->cfg.model == CLB_RICHARDS && h->cfg.top_bc == CLB_TOP_MOISTURE_STATE) TRY(alloc_fields(h, {CLB_F_DFLUXBCDY}));
    return CLB_OK;
}

int check_handle(clb_handle h)
{
    if (!h) return fail(CLB_ERR_INVALID, "null handle");
    return CLB_OK;
}

int allreduce_doubles(clb_handle h, double *dptr, size_t n)
{
    if (!h->comm) return CLB_OK;
    NCCL_TRY(g_nccl.AllReduce(dptr, dptr, n, kNcclFloat64, kNcclSum, h->comm, h->stream));
    return CLB_OK;
}

// weighted per-column balance terms -> 4 atomically accumulated sums
__global__ void __launch_bounds__(128) k_balance(const clb::DevView P, const double *w, double *out4)
{
    const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    double v[4] = {0, 0, 0, 0};
    if (c < P.ncol) {
        const double wc = w ? w[c] : 1.0;
        double water = 0.0, energy = 0.0;
        for (int i = 0; i < P.N; ++i) {
            const int64_t k = P.at(i, c);
            double vol = P.Y_theta_l[k];
            if (P.model == 1) {
                vol += P.Y_theta_i[k] * P.earth.rho_i / P.earth.rho_l;
                energy += P.Y_rho_e[k] * P.dz_c[i];
            }
            water += vol * P.dz_c[i];
        }
        v[0] = wc * water;
        v[1] = wc * P.Y_intF_w[c];
        if (P.model == 1) {
            v[2] = wc * energy;
            v[3] = wc * P.Y_intF_e[c];
        }
    }
    for (int j = 0; j < 4; ++j) {
        double s = v[j];
        for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        if ((threadIdx.x & 31) == 0 && s != 0.0) atomicAdd(out4 + j, s);
    }
}

// device-side evaluation of the soil_math.cuh functions, for the accuracy tests
__global__ void k_test_math(int kind, const double *x, const double *y, double *out, int64_t n)
{
    __shared__ __align__(16) unsigned char tab_sm[clb::fmv::kMathTabBytes];
    const clb::fmv::MathTab MT = clb::fmv::math_tab_fill(tab_sm, threadIdx.x);
    __syncthreads();
    const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    double r = 0.0;
    const double xv[1] = {x[k]};
    double rv[1] = {0.0};
    switch (kind) {
    case 6: clb::fmv::log_tab<1>(MT, xv, rv); r = rv[0]; break;
    case 7: clb::fmv::exp_tab<1>(MT, xv, rv); r = rv[0]; break;
    case 0: r = clb::fm::rcp(x[k]); break;
    case 1: r = clb::fm::div(x[k], y[k]); break;
    case 2: r = clb::fm::log(x[k]); break;
    case 3: r = clb::fm::exp(x[k]); break;
    case 4: r = clb::fm::sqrt(x[k]); break;
    case 5: r = clb::fm::rcp_seed(x[k]); break;
    case 8: r = clb::fm::div_by(x[k], y[k], 1.0 / y[k]); break;  // the reciprocal divided as the host does (IEEE)
    default: r = NAN;
    }
    out[k] = r;
}

}  // namespace

// =============================================================================
#include "clb_api_explicit.inc"

namespace {
// the assembly kernels on columns [c0, c0 + n) (c0 = 0, n = ncol: all)
int launch_atmos_assembly(clb_handle h, const clb::DevView &P, int64_t c0)
{
    const bool eh = h->cfg.model == CLB_ENERGY_HYDROLOGY;
    double *const *F = h->field;
    clb::AtmosView A = {};
    A.infiltration = F[CLB_F_INFILTRATION] + c0;
    A.top_bc_w = F[CLB_F_TOP_BC_W] + c0;
    if (eh) {
        A.vapor_flux_liq = F[CLB_F_VAPOR_FLUX_LIQ] + c0; A.lhf = F[CLB_F_LHF] + c0; A.shf = F[CLB_F_SHF] + c0;
        A.R_n = F[CLB_F_R_N] + c0; A.T_air = F[CLB_F_T_AIR] + c0; A.top_bc_h = F[CLB_F_TOP_BC_H] + c0;
    }
    clb::k_atmos_driven_fluxes<<<grid_for(P.ncol), kBlock, 0, h->stream>>>(P, A);
    CUDA_TRY(cudaGetLastError());
    return CLB_OK;
}
int atmos_fields_ready(clb_handle h, const char *who)
{
    TRY(require(h, {CLB_F_PRECIP}, who));
    if (h->cfg.model == CLB_ENERGY_HYDROLOGY)
        TRY(require(h, {CLB_F_VAPOR_FLUX_LIQ, CLB_F_LHF, CLB_F_SHF, CLB_F_R_N, CLB_F_T_AIR}, who));
    TRY(alloc_fields(h, {CLB_F_INFILTRATION, CLB_F_TOP_BC_W}));
    if (h->cfg.model == CLB_ENERGY_HYDROLOGY) TRY(alloc_fields(h, {CLB_F_TOP_BC_H}));
    return CLB_OK;
}
}  // namespace

#include "clb_api_host_step.inc"

#include "clb_api_soilco2.inc"


extern "C" {

int clb_test_math(int32_t kind, const double *x, const double *y, double *out, int64_t n)
{
    if (!x || !out || n < 1) return fail(CLB_ERR_INVALID, "clb_test_math: bad arguments");
    double *dx = nullptr, *dy = nullptr, *dout = nullptr;
    CUDA_TRY(cudaMalloc(&dx, n * sizeof(double)));
    CUDA_TRY(cudaMalloc(&dy, n * sizeof(double)));
    CUDA_TRY(cudaMalloc(&dout, n * sizeof(double)));
    CUDA_TRY(cudaMemcpy(dx, x, n * sizeof(double), cudaMemcpyHostToDevice));
    CUDA_TRY(cudaMemcpy(dy, y ? y : x, n * sizeof(double), cudaMemcpyHostToDevice));
    k_test_math<<<(unsigned)((n + 255) / 256), 256>>>(kind, dx, dy, dout, n);
    CUDA_TRY(cudaGetLastError());
    CUDA_TRY(cudaMemcpy(out, dout, n * sizeof(double), cudaMemcpyDeviceToHost));
    cudaFree(dx); cudaFree(dy); cudaFree(dout);
    return CLB_OK;
}


int clb_abi_version(void) { return CLB_ABI_VERSION; }

const char *clb_last_error(void) { return g_err.c_str(); }

int clb_create(clb_handle *out, const clb_config *cfg)
{
    if (!out || !cfg) return fail(CLB_ERR_INVALID, "clb_create: null argument");
    *out = nullptr;
    if (cfg->abi_version != CLB_ABI_VERSION)
        return fail(CLB_ERR_INVALID, "clb_create: abi_version %d, library is %d", cfg->abi_version, CLB_ABI_VERSION);
    if (cfg->model != CLB_RICHARDS && cfg->model != CLB_ENERGY_HYDROLOGY)
        return fail(CLB_ERR_INVALID, "clb_create: unknown model %d", cfg->model);
    if (cfg->closure != CLB_VAN_GENUCHTEN && cfg->closure != CLB_BROOKS_COREY)
        return fail(CLB_ERR_INVALID, "clb_create: unknown closure %d", cfg->closure);
    if (cfg->top_bc < 0 || cfg->top_bc > 1 || cfg->bottom_bc < 0 || cfg->bottom_bc > 2)
        return fail(CLB_ERR_INVALID, "clb_create: unknown boundary condition kind");
    if (cfg->layout < CLB_LAYOUT_AUTO || cfg->layout > CLB_LAYOUT_LEVEL_FASTEST)
        return fail(CLB_ERR_INVALID, "clb_create: unknown layout %d", cfg->layout);
    if (cfg->kernel_variant < CLB_VARIANT_AUTO || cfg->kernel_variant > CLB_VARIANT_LANE_OCTET)
        return fail(CLB_ERR_INVALID, "clb_create: unknown kernel_variant %d", cfg->kernel_variant);
    if (cfg->math_mode != CLB_MATH_FAST && cfg->math_mode != CLB_MATH_LIBM)
        return fail(CLB_ERR_INVALID, "clb_create: unknown math_mode %d", cfg->math_mode);
    if (cfg->n_levels < 2 || cfg->n_levels > kMaxLevels)
        return fail(CLB_ERR_INVALID, "clb_create: n_levels must be in [2, %d]", kMaxLevels);
    if (cfg->n_columns < 1) return fail(CLB_ERR_INVALID, "clb_create: n_columns must be >= 1");
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
        cudaGetLastError();
        return fail(CLB_ERR_NO_DEVICE, "clb_create: no CUDA device available; this library has no CPU path");
    }
    if (cfg->device < 0 || cfg->device >= ndev)
        return fail(CLB_ERR_INVALID, "clb_create: device %d out of range (have %d)", cfg->device, ndev);
    clb_handle h = new (std::nothrow) clb_handle_s();
    if (!h) return fail(CLB_ERR_INVALID, "clb_create: out of host memory");
    h->cfg = *cfg;
    h->stream = (cudaStream_t)cfg->stream;
    h->ld = (cfg->n_columns + 31) / 32 * 32;
    int layout = cfg->layout;
    // lane-per-cell kernels read level-fastest mirrors, every column-per-thread(-pair) kernel column-fastest ones
    if (layout == CLB_LAYOUT_AUTO) {
        // (a MoistureStateBC top of RichardsModel: the pipelined quad only, see pair_variant_applies)
        const bool live_bc = cfg->model == CLB_RICHARDS && cfg->top_bc == 1;
        const bool quad_ok = (cfg->n_levels == 15 || cfg->n_levels == 16) && cfg->math_mode == CLB_MATH_FAST &&
                             (live_bc ? (cfg->kernel_variant == CLB_VARIANT_AUTO || cfg->kernel_variant == CLB_VARIANT_LANE_QUAD_PIPELINED)
                                      : (cfg->kernel_variant == CLB_VARIANT_AUTO || cfg->kernel_variant == CLB_VARIANT_LANE_QUAD ||
                                         cfg->kernel_variant == CLB_VARIANT_LANE_QUAD_PIPELINED ||
                                         cfg->kernel_variant == CLB_VARIANT_LANE_OCTET));
        const bool lane_per_cell = cfg->n_levels <= 31 && (cfg->kernel_variant == CLB_VARIANT_AUTO ||
                                                           cfg->kernel_variant == CLB_VARIANT_LANE_PER_CELL);
        // the N = 50 octet reads either layout; level-fastest (the reference's own) keeps a tile's 50 rows of a field
        // contiguous, so its rate does not depend on the number of columns (column-fastest: rows ld * 8 bytes apart)
        const bool octet50 = cfg->n_levels == 50 && cfg->math_mode == CLB_MATH_FAST &&
                             !(cfg->model == CLB_RICHARDS && cfg->top_bc == 1) &&
                             (cfg->kernel_variant == CLB_VARIANT_AUTO || cfg->kernel_variant == CLB_VARIANT_LANE_OCTET);
        // 17 <= N <= 64 (other than 50): the octet with the level count at run time reads column-fastest mirrors; it is the choice
        // while a field stays below 80 MB (beyond that its tiles fall out of the TLB, see clb_implicit_step).  Measured
        // at 1e5 columns (tools/time_other_n.py): faster than the lane-per-cell / generic kernels at every N
        const int64_t ld0 = (cfg->n_columns + 31) / 32 * 32;
        const bool octet_rt = cfg->n_levels >= 17 && cfg->n_levels <= 64 && cfg->n_levels != 50 && cfg->math_mode == CLB_MATH_FAST &&
                              !(cfg->model == CLB_RICHARDS && cfg->top_bc == 1) &&
                              (cfg->kernel_variant == CLB_VARIANT_LANE_OCTET ||
                               (cfg->kernel_variant == CLB_VARIANT_AUTO && ld0 * cfg->n_levels * 8 <= ((int64_t)80 << 20)));
        layout = (((!quad_ok && lane_per_cell) && !octet_rt) || octet50) ? CLB_LAYOUT_LEVEL_FASTEST : CLB_LAYOUT_COLUMN_FASTEST;
    }
    h->cfg.layout = layout;
    if (layout == CLB_LAYOUT_LEVEL_FASTEST) {
        h->sl = 1;
        h->sc = cfg->n_levels;
    } else {
        h->sl = h->ld;
        h->sc = 1;
    }
    h->cell_elems = (size_t)cfg->n_levels * (size_t)h->ld;
    DeviceGuard guard(cfg->device);
    int rc = CLB_OK;
    auto init = [&]() -> int {
        CUDA_TRY(cudaMalloc(&h->d_grid, 4 * (size_t)cfg->n_levels * sizeof(double)));
        CUDA_TRY(cudaMalloc(&h->zeros_cell, h->cell_elems * sizeof(double)));
        CUDA_TRY(cudaMemsetAsync(h->zeros_cell, 0, h->cell_elems * sizeof(double), h->stream));
        CUDA_TRY(cudaMalloc(&h->zeros_col, (size_t)h->ld * sizeof(double)));
        CUDA_TRY(cudaMemsetAsync(h->zeros_col, 0, (size_t)h->ld * sizeof(double), h->stream));
        CUDA_TRY(cudaMalloc(&h->d_stats, 8 * sizeof(double)));
        CUDA_TRY(cudaMemsetAsync(h->d_stats, 0, 8 * sizeof(double), h->stream));
        CUDA_TRY(cudaMalloc(&h->d_flags, 4 * sizeof(int32_t)));
        CUDA_TRY(cudaMemsetAsync(h->d_flags, 0, 4 * sizeof(int32_t), h->stream));
        return CLB_OK;
    };
    rc = init();
    if (rc != CLB_OK) {
        clb_destroy(h);
        return rc;
    }
    *out = h;
    return CLB_OK;
}

int clb_destroy(clb_handle h)
{
    if (!h) return CLB_OK;
    DeviceGuard guard(h->cfg.device);
    cudaStreamSynchronize(h->stream);
    if (h->comm && g_nccl.CommDestroy) g_nccl.CommDestroy(h->comm);
    for (auto &p : h->field) if (!in_arena(h, p)) cudaFree(p);
    for (auto &p : h->work) cudaFree(p);
    for (auto &p : h->prep) if (!in_arena(h, p)) cudaFree(p);
    cudaFree(h->arena);
    cudaFree(h->carry);
    cudaFree(h->zeros_cell);
    cudaFree(h->zeros_col);
    cudaFree(h->d_grid);
    cudaFree(h->d_idx);
    cudaFree(h->d_stage);
    cudaFree(h->d_stage_in);
    cudaFree(h->d_stage_out);
    if (h->s_in) cudaStreamDestroy(h->s_in);
    if (h->s_out) cudaStreamDestroy(h->s_out);
    if (h->ev_start) cudaEventDestroy(h->ev_start);
    if (h->ev_done) cudaEventDestroy(h->ev_done);
    for (auto &e : h->ev_in) if (e) cudaEventDestroy(e);
    for (auto &e : h->ev_out) if (e) cudaEventDestroy(e);
    cudaFree(h->d_stats);
    cudaFree(h->d_flags);
    delete h;
    return CLB_OK;
}

int clb_sync(clb_handle h)
{
    TRY(check_handle(h));
    DeviceGuard guard(h->cfg.device);
    CUDA_TRY(cudaStreamSynchronize(h->stream));
    return CLB_OK;
}

int clb_set_stream(clb_handle h, void *stream)
{
    TRY(check_handle(h));
    DeviceGuard guard(h->cfg.device);
    CUDA_TRY(cudaStreamSynchronize(h->stream));
    h->stream = (cudaStream_t)stream;
    return CLB_OK;
}

int clb_set_option(clb_handle h, int32_t option, int64_t value)
{
    TRY(check_handle(h));
    switch (option) {
    case CLB_OPT_OUT_OF_PLACE:
        h->out_of_place = value != 0;
        return CLB_OK;
    case CLB_OPT_CO2_TOP_STATE:
        h->co2_top_state[0] = value != 0;
        return CLB_OK;
    case CLB_OPT_O2_TOP_STATE:
        h->co2_top_state[1] = value != 0;
        return CLB_OK;
    case CLB_OPT_HOST_ROUTE:
        if (value < 0 || value > 3) return fail(CLB_ERR_INVALID, "clb_set_option: CLB_OPT_HOST_ROUTE takes 0 .. 3");
        h->host_route = (int)value;
        return CLB_OK;
    case CLB_OPT_HOST_CHUNKS:
        if (value < 0 || value > 16) return fail(CLB_ERR_INVALID, "clb_set_option: CLB_OPT_HOST_CHUNKS takes 0 .. 16");
        h->host_chunks = (int)value;
        return CLB_OK;
    case CLB_OPT_TILE_BOXES:
        h->tile_boxes = value != 0;
        return CLB_OK;
    case CLB_OPT_EXPLICIT_KERNEL:
        h->explicit_kernel = value != 0;
        return CLB_OK;
    case CLB_OPT_RUNOFF_MODEL:
        if (value < CLB_RUNOFF_NONE || value > CLB_RUNOFF_TOPMODEL) return fail(CLB_ERR_INVALID, "clb_set_option: CLB_OPT_RUNOFF_MODEL takes a CLB_RUNOFF_* value");
        h->runoff_model = (int)value;
        return CLB_OK;
    case CLB_OPT_TOP_ATMOS_DRIVEN:
        h->top_atmos = value != 0;
        return CLB_OK;
    case CLB_OPT_BOTTOM_EWFD:
        h->bottom_ewfd = value != 0;
        return CLB_OK;
    default:
        return fail(CLB_ERR_INVALID, "clb_set_option: unknown option %d", option);
    }
}

int clb_set_grid(clb_handle h, const double *z_c, const double *z_f)
{
    TRY(check_handle(h));
    if (!z_c || !z_f) return fail(CLB_ERR_INVALID, "clb_set_grid: null pointer");
    const int N = h->cfg.n_levels;
    for (int i = 0; i < N; ++i)
        if (!(z_f[i + 1] > z_f[i])) return fail(CLB_ERR_INVALID, "clb_set_grid: z_f must increase (bottom -> top)");
    for (int i = 1; i < N; ++i)
        if (!(z_c[i] > z_c[i - 1])) return fail(CLB_ERR_INVALID, "clb_set_grid: z_c must increase (bottom -> top)");
    h->z_c.assign(z_c, z_c + N);
    h->z_f.assign(z_f, z_f + N + 1);
    h->dz_c.resize(N);
    h->inv_dz_c.resize(N);
    h->inv_dz_f.assign(N, 0.0);
    for (int i = 0; i < N; ++i) {
        h->dz_c[i] = z_f[i + 1] - z_f[i];
        h->inv_dz_c[i] = 1.0 / h->dz_c[i];
        if (i > 0) h->inv_dz_f[i] = 1.0 / (z_c[i] - z_c[i - 1]);
    }
    std::vector<double> pack;
    pack.insert(pack.end(), h->z_c.begin(), h->z_c.end());
    pack.insert(pack.end(), h->dz_c.begin(), h->dz_c.end());
    pack.insert(pack.end(), h->inv_dz_c.begin(), h->inv_dz_c.end());
    pack.insert(pack.end(), h->inv_dz_f.begin(), h->inv_dz_f.end());
    DeviceGuard guard(h->cfg.device);
    CUDA_TRY(cudaStreamSynchronize(h->stream));
    CUDA_TRY(cudaMemcpy(h->d_grid, pack.data(), pack.size() * sizeof(double), cudaMemcpyHostToDevice));
    h->grid_set = true;
    return CLB_OK;
}

int clb_set_active_columns(clb_handle h, const int64_t *idx, int64_t n)
{
    TRY(check_handle(h));
    DeviceGuard guard(h->cfg.device);
    CUDA_TRY(cudaStreamSynchronize(h->stream));
    if (!idx) {
        cudaFree(h->d_idx);
        h->d_idx = nullptr;
        h->idx_max = -1;
        return CLB_OK;
    }
    if (n != h->cfg.n_columns)
        return fail(CLB_ERR_INVALID, "clb_set_active_columns: n = %lld but the handle holds %lld columns", (long long)n,
                    (long long)h->cfg.n_columns);
    int64_t mx = -1;
    for (int64_t j = 0; j < n; ++j) {
        if (idx[j] < 0) return fail(CLB_ERR_INVALID, "clb_set_active_columns: negative index");
        mx = std::max(mx, idx[j]);
    }
    if (!h->d_idx) CUDA_TRY(cudaMalloc(&h->d_idx, (size_t)n * sizeof(int64_t)));
    CUDA_TRY(cudaMemcpy(h->d_idx, idx, (size_t)n * sizeof(int64_t), cudaMemcpyHostToDevice));
    h->idx_max = mx;
    return CLB_OK;
}

static int transfer_geometry(clb_handle h, int32_t field, int64_t sl, int64_t sc, bool &cell, int64_t &extent,
                             bool &dense)
{
    if (!is_cell_field(field) && !is_col_field(field)) return fail(CLB_ERR_INVALID, "unknown field id %d", field);
    if (sl < 0 || sc < 0) return fail(CLB_ERR_INVALID, "strides must be >= 0");
    cell = is_cell_field(field);
    const int N = h->cfg.n_levels;
    const int64_t maxcol = h->d_idx ? h->idx_max : h->cfg.n_columns - 1;
    extent = maxcol * sc + (cell ? (int64_t)(N - 1) * sl : 0) + 1;
    dense = !h->d_idx && (cell ? (sl == 1 && sc == N) : (sc == 1));
    return CLB_OK;
}

int clb_set_field(clb_handle h, int32_t field, const double *src, int64_t stride_level, int64_t stride_column,
                  int32_t mem)
{
    TRY(check_handle(h));
    if (!src) return fail(CLB_ERR_INVALID, "clb_set_field: null source");
    bool cell, dense;
    int64_t extent;
    TRY(transfer_geometry(h, field, stride_level, stride_column, cell, extent, dense));
    DeviceGuard guard(h->cfg.device);
    TRY(ensure_field(h, field));
    const double *dsrc = src;
    if (mem == CLB_HOST) {
        TRY(ensure_stage(h, (size_t)extent * sizeof(double)));
        CUDA_TRY(cudaMemcpyAsync(h->d_stage, src, (size_t)extent * sizeof(double), cudaMemcpyHostToDevice, h->stream));
        // a pinned source makes that copy truly asynchronous: wait for it, so that the caller may reuse or free
        // `src` as soon as this call returns (the relayout below stays asynchronous)
        CUDA_TRY(cudaStreamSynchronize(h->stream));
        dsrc = h->d_stage;
    } else if (mem != CLB_DEVICE) {
        return fail(CLB_ERR_INVALID, "clb_set_field: mem must be CLB_HOST or CLB_DEVICE");
    }
    const int N = h->cfg.n_levels;
    const int64_t ncol = h->cfg.n_columns;
    if (cell) {
        const unsigned grid = (unsigned)((ncol + clb::kTileCols - 1) / clb::kTileCols);
        const size_t smem = (size_t)N * (clb::kTileCols + 1) * sizeof(double);
        if (smem > 48 * 1024)
            CUDA_TRY(cudaFuncSetAttribute(clb::k_relayout, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        clb::k_relayout<<<grid, 256, smem, h->stream>>>(h->field[field], h->sl, h->sc, nullptr, dsrc, stride_level,
                                                        stride_column, h->d_idx, N, ncol);
    } else {
        clb::k_gather_cols<<<(unsigned)((ncol + 255) / 256), 256, 0, h->stream>>>(h->field[field], dsrc, stride_column,
                                                                                  h->d_idx, ncol);
    }
    CUDA_TRY(cudaGetLastError());
    h->field_set[field] = true;
    if (is_closure_param(field)) h->prep_dirty = true;
    if (is_invariant_param(field)) h->param_write_pending = true;
    return CLB_OK;
}

int clb_get_field(clb_handle h, int32_t field, double *dst, int64_t stride_level, int64_t stride_column, int32_t mem)
{
    TRY(check_handle(h));
    if (!dst) return fail(CLB_ERR_INVALID, "clb_get_field: null destination");
    bool cell, dense;
    int64_t extent;
    TRY(transfer_geometry(h, field, stride_level, stride_column, cell, extent, dense));
    if (!h->field[field]) return fail(CLB_ERR_UNSET, "clb_get_field: field %d was never set or computed", field);
    DeviceGuard guard(h->cfg.device);
    double *ddst = dst;
    if (mem == CLB_HOST) {
        TRY(ensure_stage(h, (size_t)extent * sizeof(double)));
        // a strided / masked destination keeps the elements the library does not own
        if (!dense)
            CUDA_TRY(cudaMemcpyAsync(h->d_stage, dst, (size_t)extent * sizeof(double), cudaMemcpyHostToDevice, h->stream));
        ddst = h->d_stage;
    } else if (mem != CLB_DEVICE) {
        return fail(CLB_ERR_INVALID, "clb_get_field: mem must be CLB_HOST or CLB_DEVICE");
    }
    const int N = h->cfg.n_levels;
    const int64_t ncol = h->cfg.n_columns;
    if (cell) {
        const unsigned grid = (unsigned)((ncol + clb::kTileCols - 1) / clb::kTileCols);
        const size_t smem = (size_t)N * (clb::kTileCols + 1) * sizeof(double);
        if (smem > 48 * 1024)
            CUDA_TRY(cudaFuncSetAttribute(clb::k_relayout, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        clb::k_relayout<<<grid, 256, smem, h->stream>>>(ddst, stride_level, stride_column, h->d_idx, h->field[field],
                                                        h->sl, h->sc, nullptr, N, ncol);
    } else {
        clb::k_scatter_cols<<<(unsigned)((ncol + 255) / 256), 256, 0, h->stream>>>(h->field[field], ddst,
                                                                                   stride_column, h->d_idx, ncol);
    }
    CUDA_TRY(cudaGetLastError());
    if (mem == CLB_HOST) {
        CUDA_TRY(cudaMemcpyAsync(dst, h->d_stage, (size_t)extent * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
        CUDA_TRY(cudaStreamSynchronize(h->stream));
    }
    return CLB_OK;
}

static int field_axpby(clb_handle h, int32_t y, double a, int32_t x, double b, const char *who)
{
    TRY(check_handle(h));
    const bool cell = is_cell_field(y);
    if (!(cell || is_col_field(y)) || !(is_cell_field(x) || is_col_field(x)) || cell != is_cell_field(x))
        return fail(CLB_ERR_INVALID, "%s: two per-cell or two per-column field ids expected", who);
    if (!h->field[x]) return fail(CLB_ERR_UNSET, "%s: field %d was never set or computed", who, x);
    if (b != 0.0 && !h->field[y]) return fail(CLB_ERR_UNSET, "%s: field %d was never set or computed", who, y);
    DeviceGuard guard(h->cfg.device);
    TRY(ensure_field(h, y));
    const int64_t n = cell ? (int64_t)h->cell_elems : h->ld;
    clb::k_axpby<<<(unsigned)((n + 255) / 256), 256, 0, h->stream>>>(h->field[y], h->field[x], a, b, n);
    CUDA_TRY(cudaGetLastError());
    h->field_set[y] = true;
    if (is_closure_param(y)) h->prep_dirty = true;
    if (is_invariant_param(y)) h->param_write_pending = true;
    return CLB_OK;
}

int clb_ldiv_diagonal(clb_handle h, int32_t w_field, int32_t b_field, int32_t x_field)
{
    TRY(check_handle(h));
    const int f[3] = {w_field, b_field, x_field};
    for (int k = 0; k < 3; ++k)
        if (!(is_cell_field(f[k]) || is_col_field(f[k])) || is_cell_field(f[k]) != is_cell_field(f[0]))
            return fail(CLB_ERR_INVALID, "clb_ldiv_diagonal: three per-cell or three per-column field ids expected");
    if (!h->field_set[w_field] || !h->field_set[b_field])
        return fail(CLB_ERR_UNSET, "clb_ldiv_diagonal: the block (field %d) and the right-hand side (field %d) must be set", w_field, b_field);
    DeviceGuard guard(h->cfg.device);
    TRY(alloc_fields(h, {x_field}));
    const int64_t n = is_cell_field(w_field) ? (int64_t)h->cell_elems : h->ld;
    clb::k_div<<<(unsigned)((n + 255) / 256), 256, 0, h->stream>>>(h->field[x_field], h->field[b_field], h->field[w_field], n);
    CUDA_TRY(cudaGetLastError());
    return CLB_OK;
}

int clb_field_axpy(clb_handle h, int32_t y, double a, int32_t x) { return field_axpby(h, y, a, x, 1.0, "clb_field_axpy"); }
int clb_field_copy(clb_handle h, int32_t dst, int32_t src) { return field_axpby(h, dst, 1.0, src, 0.0, "clb_field_copy"); }

int clb_fill_field(clb_handle h, int32_t field, double value)
{
    TRY(check_handle(h));
    if (!is_cell_field(field) && !is_col_field(field)) return fail(CLB_ERR_INVALID, "unknown field id %d", field);
    DeviceGuard guard(h->cfg.device);
    TRY(ensure_field(h, field));
    const int64_t n = is_cell_field(field) ? (int64_t)h->cell_elems : h->ld;
    clb::k_fill<<<(unsigned)((n + 255) / 256), 256, 0, h->stream>>>(h->field[field], n, value);
    CUDA_TRY(cudaGetLastError());
    h->field_set[field] = true;
    if (is_closure_param(field)) h->prep_dirty = true;
    if (is_invariant_param(field)) h->param_write_pending = true;
    return CLB_OK;
}

int clb_field_device_ptr(clb_handle h, int32_t field, double **ptr, int64_t *stride_level, int64_t *stride_column)
{
    TRY(check_handle(h));
    if (!is_cell_field(field) && !is_col_field(field)) return fail(CLB_ERR_INVALID, "unknown field id %d", field);
    if (!ptr) return fail(CLB_ERR_INVALID, "clb_field_device_ptr: null output");
    DeviceGuard guard(h->cfg.device);
    TRY(ensure_field(h, field));
    h->field_set[field] = true;  // the caller fills it in place
    if (is_invariant_param(field)) h->prep_volatile = true;
    *ptr = h->field[field];
    if (stride_level) *stride_level = is_cell_field(field) ? h->sl : 0;
    if (stride_column) *stride_column = is_cell_field(field) ? h->sc : 1;
    return CLB_OK;
}

int clb_update_implicit_cache(clb_handle h)
{
    TRY(check_handle(h));
    DeviceGuard guard(h->cfg.device);
    TRY(step_inputs_ready(h));
    TRY(alloc_fields(h, {CLB_F_P_PSI}));
    if (h->cfg.model == CLB_RICHARDS)
        TRY(alloc_fields(h, {CLB_F_P_K, CLB_F_TOTAL_WATER}));
    else
        TRY(alloc_fields(h, {CLB_F_P_T}));
    nvtxRangePushA("update_implicit_cache!");
    const clb::DevView P = make_view(h);
    DISPATCH_CM(h, clb::k_update_implicit_cache, grid_for(P.ncol), P);
    nvtxRangePop();
    CUDA_TRY(cudaGetLastError());
    return CLB_OK;
}

int clb_set_explicit_params(clb_handle h, const clb_explicit_params *p)
{
    TRY(check_handle(h));
    if (!p) return fail(CLB_ERR_INVALID, "clb_set_explicit_params: null parameters");
    if (!(p->grav > 0.0) || !(p->T_freeze > 0.0))
        return fail(CLB_ERR_INVALID, "clb_set_explicit_params: grav and T_freeze must be positive");
    h->explicit_k = {p->Omega, p->gamma, p->gammaT_ref, p->alpha, p->beta, p->T_freeze, p->grav};
    h->explicit_set = true;
    return CLB_OK;
}

int clb_soilco2_update_boundary_fluxes(clb_handle h) { return co2_launch(h, 0, 0.0, 0, "soilco2 update_boundary_fluxes!"); }
int clb_soilco2_compute_imp_tendency(clb_handle h) { return co2_launch(h, 1, 0.0, 0, "soilco2 compute_imp_tendency!"); }
int clb_soilco2_compute_jacobian(clb_handle h, double dtgamma) { return co2_launch(h, 2, dtgamma, 0, "soilco2 compute_jacobian!"); }
int clb_soilco2_implicit_step(clb_handle h, double dtgamma, int32_t max_iters)
{
    if (max_iters < 1) return fail(CLB_ERR_INVALID, "clb_soilco2_implicit_step: max_iters must be >= 1");
    return co2_launch(h, 3, dtgamma, max_iters, "soilco2 implicit_step!");
}

int clb_set_runoff_params(clb_handle h, const clb_runoff_params *p)
{
    TRY(check_handle(h));
    if (!p) return fail(CLB_ERR_INVALID, "clb_set_runoff_params: null parameters");
    if (!(p->depth > 0.0)) return fail(CLB_ERR_INVALID, "clb_set_runoff_params: depth must be positive");
    h->runoff_k = *p;
    h->runoff_set = true;
    return CLB_OK;
}

int clb_update_runoff(clb_handle h)
{
    TRY(check_handle(h));
    DeviceGuard guard(h->cfg.device);
    const bool eh = h->cfg.model == CLB_ENERGY_HYDROLOGY;
    if (!h->grid_set) return fail(CLB_ERR_UNSET, "clb_set_grid was never called");
    if (!h->runoff_set) return fail(CLB_ERR_UNSET, "update_runoff: clb_set_runoff_params was never called");
    TRY(require(h, {CLB_F_NU, CLB_F_THETA_R, CLB_F_K_SAT, CLB_F_Y_THETA_L, CLB_F_F_MAX, CLB_F_PRECIP}, "update_runoff"));
    if (eh) {
        if (!h->explicit_set) return fail(CLB_ERR_UNSET, "update_runoff: clb_set_explicit_params was never called");
        TRY(require(h, {CLB_F_Y_THETA_I, CLB_F_THETA_L_LAG, CLB_F_P_T}, "update_runoff (call clb_update_aux first)"));
        TRY(alloc_fields(h, {CLB_F_R_ESS}));
    }
    TRY(alloc_fields(h, {CLB_F_IS_SATURATED, CLB_F_H_GRAD, CLB_F_R_SS, CLB_F_INFILTRATION, CLB_F_R_S}));
    const clb::DevView P = make_view(h);
    double *const *F = h->field;
    clb::RunoffView R;
    R.f_max = F[CLB_F_F_MAX]; R.precip = F[CLB_F_PRECIP];
    R.is_sat = F[CLB_F_IS_SATURATED]; R.h_grad = F[CLB_F_H_GRAD]; R.infiltration = F[CLB_F_INFILTRATION];
    R.R_s = F[CLB_F_R_S]; R.R_ss = F[CLB_F_R_SS]; R.R_ess = F[CLB_F_R_ESS];
    R.p_theta_l = F[CLB_F_THETA_L_LAG]; R.p_T = F[CLB_F_P_T];
    R.f_over = h->runoff_k.f_over; R.R_sb = h->runoff_k.R_sb; R.depth = h->runoff_k.depth;
    R.Omega = h->explicit_k.Omega; R.gamma = h->explicit_k.gamma; R.gammaT_ref = h->explicit_k.gammaT_ref;
    nvtxRangePushA("update_runoff!");
    if (h->cfg.math_mode == CLB_MATH_FAST) clb::k_update_runoff<0><<<grid_for(P.ncol), kBlock, 0, h->stream>>>(P, R);
    else clb::k_update_runoff<1><<<grid_for(P.ncol), kBlock, 0, h->stream>>>(P, R);
    nvtxRangePop();
    CUDA_TRY(cudaGetLastError());
    return CLB_OK;
}

int clb_update_atmos_driven_fluxes(clb_handle h, int32_t runoff_model)
{
    TRY(check_handle(h));
    const bool eh = h->cfg.model == CLB_ENERGY_HYDROLOGY;
    if (runoff_model < CLB_RUNOFF_NONE || runoff_model > CLB_RUNOFF_TOPMODEL)
        return fail(CLB_ERR_INVALID, "clb_update_atmos_driven_fluxes: runoff_model must be a CLB_RUNOFF_* value");
    if (!h->grid_set) return fail(CLB_ERR_UNSET, "clb_set_grid was never called");
    TRY(atmos_fields_ready(h, "update_atmos_driven_fluxes"));
    if (runoff_model == CLB_RUNOFF_TOPMODEL) {
        TRY(clb_update_runoff(h));
    } else {
        DeviceGuard guard(h->cfg.device);
        TRY(require(h, {CLB_F_NU, CLB_F_THETA_R, CLB_F_K_SAT, CLB_F_Y_THETA_L}, "update_atmos_driven_fluxes"));
        if (eh && runoff_model == CLB_RUNOFF_SURFACE) {
            if (!h->explicit_set) return fail(CLB_ERR_UNSET, "update_atmos_driven_fluxes: clb_set_explicit_params was never called");
            TRY(require(h, {CLB_F_Y_THETA_I, CLB_F_THETA_L_LAG, CLB_F_P_T}, "update_atmos_driven_fluxes (call clb_update_aux first)"));
        }
        if (runoff_model == CLB_RUNOFF_SURFACE) TRY(alloc_fields(h, {CLB_F_IS_SATURATED, CLB_F_R_S}));
        const clb::DevView P = make_view(h);
        double *const *F = h->field;
        clb::RunoffView R = {};
        R.model = runoff_model;
        R.precip = F[CLB_F_PRECIP];
        R.is_sat = F[CLB_F_IS_SATURATED]; R.infiltration = F[CLB_F_INFILTRATION]; R.R_s = F[CLB_F_R_S];
        R.p_theta_l = F[CLB_F_THETA_L_LAG]; R.p_T = F[CLB_F_P_T];
        R.Omega = h->explicit_k.Omega; R.gamma = h->explicit_k.gamma; R.gammaT_ref = h->explicit_k.gammaT_ref;
        if (eh && !R.p_T) R.p_T = F[CLB_F_Y_THETA_L];  // NoRunoff never reads it; keeps the sweep off its own-T path
        if (h->cfg.math_mode == CLB_MATH_FAST) clb::k_update_runoff<0><<<grid_for(P.ncol), kBlock, 0, h->stream>>>(P, R);
        else clb::k_update_runoff<1><<<grid_for(P.ncol), kBlock, 0, h->stream>>>(P, R);
        CUDA_TRY(cudaGetLastError());
    }
    DeviceGuard guard(h->cfg.device);
    return launch_atmos_assembly(h, make_view(h), 0);
}

int clb_update_energy_water_free_drainage(clb_handle h)
{
    TRY(check_handle(h));
    if (h->cfg.model != CLB_ENERGY_HYDROLOGY) return fail(CLB_ERR_INVALID, "clb_update_energy_water_free_drainage: EnergyHydrology only");
    TRY(require(h, {CLB_F_K_LAG, CLB_F_P_T}, "update_energy_water_free_drainage (call clb_update_aux first)"));
    DeviceGuard guard(h->cfg.device);
    TRY(alloc_fields(h, {CLB_F_BOT_BC_W, CLB_F_BOT_BC_H}));
    const clb::DevView P = make_view(h);
    clb::k_energy_water_free_drainage<<<grid_for(P.ncol), kBlock, 0, h->stream>>>(P, h->field[CLB_F_K_LAG], h->field[CLB_F_P_T],
                                                                                    h->field[CLB_F_BOT_BC_W], h->field[CLB_F_BOT_BC_H]);
    CUDA_TRY(cudaGetLastError());
    return CLB_OK;
}

int clb_update_aux(clb_handle h) { return launch_explicit<true, false>(h, "update_aux!"); }
int clb_phase_change_source(clb_handle h) { return launch_explicit<false, true>(h, "source!(PhaseChange)"); }
int clb_update_aux_and_phase_change(clb_handle h) { return launch_explicit<true, true>(h, "update_aux! + PhaseChange"); }

int clb_update_boundary_fluxes(clb_handle h)
{
    TRY(check_handle(h));
    DeviceGuard guard(h->cfg.device);
    TRY(step_inputs_ready(h));
    TRY(require(h, {CLB_F_P_PSI}, "update_boundary_fluxes"));
    if (h->cfg.model == CLB_RICHARDS) TRY(require(h, {CLB_F_P_K}, "update_boundary_fluxes"));
    const clb::DevView P = make_view(h);
    DISPATCH_CM(h, clb::k_update_boundary_fluxes, grid_for(P.ncol), P);
    CUDA_TRY(cudaGetLastError());
    return CLB_OK;
}

int clb_compute_imp_tendency(clb_handle h)
{
    TRY(check_handle(h));
    DeviceGuard guard(h->cfg.device);
    TRY(step_inputs_ready(h));
    TRY(require(h, {CLB_F_P_PSI}, "compute_imp_tendency (call update_implicit_cache first)"));
    TRY(alloc_fields(h, {CLB_F_DY_THETA_L, CLB_F_DY_INTF_W}));
    if (h->cfg.model == CLB_RICHARDS) {
        TRY(require(h, {CLB_F_P_K}, "compute_imp_tendency"));
    } else {
        TRY(require(h, {CLB_F_P_T}, "compute_imp_tendency"));
        TRY(alloc_fields(h, {CLB_F_DY_RHO_E_INT, CLB_F_DY_THETA_I, CLB_F_DY_INTF_E}));
    }
    nvtxRangePushA("compute_imp_tendency!");
    const clb::DevView P = make_view(h);
    clb::k_imp_tendency<<<grid_for(P.ncol), kBlock, 0, h->stream>>>(P);
    nvtxRangePop();
    CUDA_TRY(cudaGetLastError());
    return CLB_OK;
}

int clb_compute_jacobian(clb_handle h, double dtgamma)
{
    TRY(check_handle(h));
    DeviceGuard guard(h->cfg.device);
    TRY(step_inputs_ready(h));
    TRY(alloc_fields(h, {CLB_F_W11_LO, CLB_F_W11_DI, CLB_F_W11_UP}));
    if (h->cfg.model == CLB_RICHARDS) {
        TRY(require(h, {CLB_F_P_K}, "compute_jacobian (call update_implicit_cache first)"));
    } else {
        TRY(require(h, {CLB_F_P_T}, "compute_jacobian (call update_implicit_cache first)"));
        TRY(alloc_fields(h, {CLB_F_W21_LO, CLB_F_W21_DI, CLB_F_W21_UP, CLB_F_W22_LO, CLB_F_W22_DI, CLB_F_W22_UP}));
    }
    nvtxRangePushA("compute_jacobian!");
    const clb::DevView P = make_view(h);
    DISPATCH_CM(h, clb::k_jacobian, grid_for(P.ncol), P, dtgamma);
    nvtxRangePop();
    CUDA_TRY(cudaGetLastError());
    return CLB_OK;
}

int clb_ldiv(clb_handle h)
{
    TRY(check_handle(h));
    DeviceGuard guard(h->cfg.device);
    TRY(require(h, {CLB_F_W11_LO, CLB_F_W11_DI, CLB_F_W11_UP, CLB_F_B_THETA_L}, "ldiv"));
    TRY(alloc_fields(h, {CLB_F_X_THETA_L, CLB_F_X_INTF_W}));
    TRY(ensure_field(h, CLB_F_B_INTF_W));
    if (h->cfg.model == CLB_ENERGY_HYDROLOGY) {
        TRY(require(h, {CLB_F_W21_LO, CLB_F_W21_DI, CLB_F_W21_UP, CLB_F_W22_LO, CLB_F_W22_DI, CLB_F_W22_UP,
                        CLB_F_B_RHO_E_INT},
                    "ldiv"));
        TRY(alloc_fields(h, {CLB_F_X_RHO_E_INT, CLB_F_X_THETA_I, CLB_F_X_INTF_E}));
        TRY(ensure_field(h, CLB_F_B_THETA_I));
        TRY(ensure_field(h, CLB_F_B_INTF_E));
    }
    TRY(ensure_work(h, 2));
    nvtxRangePushA("ldiv!");
    const clb::DevView P = make_view(h);
    clb::k_ldiv<<<grid_for(P.ncol), kBlock, 0, h->stream>>>(P);
    nvtxRangePop();
    CUDA_TRY(cudaGetLastError());
    return CLB_OK;
}

int clb_ldiv_all(clb_handle h, uint32_t blocks)
{
    TRY(check_handle(h));
    if (blocks == 0 || (blocks & ~7u)) return fail(CLB_ERR_INVALID, "clb_ldiv_all: blocks must be a non-empty CLB_LDIV_* mask");
    DeviceGuard guard(h->cfg.device);
    if (blocks & CLB_LDIV_SOIL) {
        TRY(require(h, {CLB_F_W11_LO, CLB_F_W11_DI, CLB_F_W11_UP, CLB_F_B_THETA_L}, "ldiv_all"));
        TRY(alloc_fields(h, {CLB_F_X_THETA_L, CLB_F_X_INTF_W}));
        TRY(ensure_field(h, CLB_F_B_INTF_W));
        if (h->cfg.model == CLB_ENERGY_HYDROLOGY) {
            TRY(require(h, {CLB_F_W21_LO, CLB_F_W21_DI, CLB_F_W21_UP, CLB_F_W22_LO, CLB_F_W22_DI, CLB_F_W22_UP, CLB_F_B_RHO_E_INT}, "ldiv_all"));
            TRY(alloc_fields(h, {CLB_F_X_RHO_E_INT, CLB_F_X_THETA_I, CLB_F_X_INTF_E}));
            TRY(ensure_field(h, CLB_F_B_THETA_I));
            TRY(ensure_field(h, CLB_F_B_INTF_E));
        }
    }
    if (blocks & CLB_LDIV_SOILCO2) {
        TRY(require(h, {CLB_F_CO2_W_LO, CLB_F_CO2_W_DI, CLB_F_CO2_W_UP, CLB_F_O2_W_LO, CLB_F_O2_W_DI, CLB_F_O2_W_UP, CLB_F_CO2_B, CLB_F_O2_B},
                    "ldiv_all (clb_soilco2_compute_jacobian and the right-hand sides first)"));
        TRY(alloc_fields(h, {CLB_F_CO2_X, CLB_F_O2_X}));
    }
    if (blocks & CLB_LDIV_SURFACE) {
        TRY(require(h, {CLB_F_SFC_W_DI, CLB_F_SFC_B}, "ldiv_all"));
        TRY(alloc_fields(h, {CLB_F_SFC_X}));
    }
    TRY(ensure_work(h, 4));
    const clb::DevView P = make_view(h);
    double *const *F = h->field;
    clb::LdivAllView A = {};
    A.blocks = blocks;
    A.co2_lo[0] = F[CLB_F_CO2_W_LO]; A.co2_di[0] = F[CLB_F_CO2_W_DI]; A.co2_up[0] = F[CLB_F_CO2_W_UP];
    A.co2_lo[1] = F[CLB_F_O2_W_LO]; A.co2_di[1] = F[CLB_F_O2_W_DI]; A.co2_up[1] = F[CLB_F_O2_W_UP];
    A.co2_b[0] = F[CLB_F_CO2_B]; A.co2_b[1] = F[CLB_F_O2_B]; A.co2_x[0] = F[CLB_F_CO2_X]; A.co2_x[1] = F[CLB_F_O2_X];
    A.sfc_w = F[CLB_F_SFC_W_DI]; A.sfc_b = F[CLB_F_SFC_B]; A.sfc_x = F[CLB_F_SFC_X];
    nvtxRangePushA("ldiv! (all blocks)");
    clb::k_ldiv_all<<<dim3(grid_for(P.ncol), 4), kBlock, 0, h->stream>>>(P, A);
    nvtxRangePop();
    CUDA_TRY(cudaGetLastError());
    return CLB_OK;
}

int clb_implicit_step(clb_handle h, double dtgamma, int32_t max_iters, double tol, clb_stats *stats)
{
    TRY(check_handle(h));
    if (max_iters < 1) return fail(CLB_ERR_INVALID, "clb_implicit_step: max_iters must be >= 1");
    DeviceGuard guard(h->cfg.device);
    TRY(step_inputs_ready(h));
    const int N = h->cfg.n_levels;
    const bool eh = h->cfg.model == CLB_ENERGY_HYDROLOGY;
    const bool fixed = tol < 0.0;
    int variant = h->cfg.kernel_variant;
    const bool level_fast = h->cfg.layout == CLB_LAYOUT_LEVEL_FASTEST;
    if (!fixed) {
        variant = CLB_VARIANT_GENERIC;  // the tolerance path is one generic launch per iteration
    } else if (variant == CLB_VARIANT_AUTO) {
        if (pair_variant_applies(h))
            variant = CLB_VARIANT_LANE_QUAD_PIPELINED;
        else if (pair_variant_applies(h, true) && (level_fast || (int64_t)h->ld * N * 8 <= (int64_t)80 << 20))
            // N = 50 (and 17 .. 64 on column-fastest mirrors).  A tile of the octet touches all 50 level rows of every field at once.  In level-fastest mirrors
            // (what CLB_LAYOUT_AUTO picks for N = 50) they are contiguous.  In column-fastest mirrors they lie
            // ld * 8 bytes apart and, as the fields grow, the tile's ~550 pages fall out of the TLB (measured against
            // the generic kernel, tools/n50_crossover.py: 2.1x faster at 1e5 columns, equal at ~2.2e5 = 88 MB per
            // field, 1.7x slower at 1e6): there the generic kernel takes over above 80 MB per field
            variant = CLB_VARIANT_LANE_OCTET;
        else if (level_fast && N <= 31)
            variant = CLB_VARIANT_LANE_PER_CELL;
        else if (N == 15)
            variant = CLB_VARIANT_REGISTER_COLUMN;
        else
            variant = CLB_VARIANT_GENERIC;
    }
    if (variant == CLB_VARIANT_REGISTER_COLUMN && N != 15)
        return fail(CLB_ERR_INVALID, "clb_implicit_step: the register-column variant is built for N == 15");
    if ((variant == CLB_VARIANT_LANE_QUAD || variant == CLB_VARIANT_LANE_QUAD_PIPELINED) && !pair_variant_applies(h))
        return fail(CLB_ERR_INVALID,
                    "clb_implicit_step: the lane-quad variants need N = 15 or 16, CLB_MATH_FAST and column-fastest mirrors (a "
                    "MoistureStateBC top of RichardsModel: the pipelined quad only)");
    if (variant == CLB_VARIANT_LANE_OCTET && !pair_variant_applies(h, true))
        return fail(CLB_ERR_INVALID,
                    "clb_implicit_step: the lane-octet variant needs 15 <= N <= 64 on column-fastest mirrors or N = 50, CLB_MATH_FAST and "
                    "flux boundary conditions");
    if (variant == CLB_VARIANT_LANE_PER_CELL && N > 31)
        return fail(CLB_ERR_INVALID, "clb_implicit_step: the lane-per-cell variant needs N <= 31");
    if (variant == CLB_VARIANT_GENERIC) TRY(ensure_work(h, eh ? 6 : 3));
    h->last_variant = variant;
    if (h->out_of_place) {
        TRY(alloc_fields(h, {CLB_F_U_THETA_L, CLB_F_U_INTF_W}));
        if (eh) TRY(alloc_fields(h, {CLB_F_U_RHO_E_INT, CLB_F_U_INTF_E}));
    }
    clb::DevView P = make_view(h);
    const bool quad = variant == CLB_VARIANT_LANE_QUAD || variant == CLB_VARIANT_LANE_QUAD_PIPELINED ||
                      variant == CLB_VARIANT_LANE_OCTET;
    if (quad && fixed && !stats)
        P.stats = nullptr;  // nobody reads them: no memset between stages, no atomics in the kernel
    else
        CUDA_TRY(cudaMemsetAsync(h->d_stats, 0, 3 * sizeof(double), h->stream));
    const unsigned grid = grid_for(P.ncol);
    nvtxRangePushA("implicit_step!");
    int iters_done = max_iters;
    if (fixed) {
#ifdef CLB_DEV_FAST  // tuning builds (tools/build_variant.sh NAME -DCLB_DEV_FAST): only the pipelined N = 15 quad is compiled
        if (variant == CLB_VARIANT_LANE_QUAD_PIPELINED && N == 15) {
            TRY((launch_quad_n<15, true>(h, P, dtgamma, max_iters)));
        } else if (variant == CLB_VARIANT_LANE_QUAD || variant == CLB_VARIANT_LANE_QUAD_PIPELINED ||
                   variant == CLB_VARIANT_LANE_OCTET) {
            return fail(CLB_ERR_INVALID, "CLB_DEV_FAST build: only the pipelined N = 15 quad is compiled");
#else
        if (variant == CLB_VARIANT_LANE_QUAD) {
            if (N == 15) TRY((launch_quad_n<15, false>(h, P, dtgamma, max_iters)));
            else TRY((launch_quad_n<16, false>(h, P, dtgamma, max_iters)));
        } else if (variant == CLB_VARIANT_LANE_QUAD_PIPELINED) {
            if (N == 15) TRY((launch_quad_n<15, true>(h, P, dtgamma, max_iters)));
            else TRY((launch_quad_n<16, true>(h, P, dtgamma, max_iters)));
        } else if (variant == CLB_VARIANT_LANE_OCTET) {
            if (N == 15) TRY((launch_octet_n<15>(h, P, dtgamma, max_iters)));
            else if (N == 16) TRY((launch_octet_n<16>(h, P, dtgamma, max_iters)));
            else if (N == 50) TRY((launch_octet_n<50>(h, P, dtgamma, max_iters)));
            else if (N <= 24) TRY((launch_octet_rt_q<3>(h, P, dtgamma, max_iters)));
            else if (N <= 32) TRY((launch_octet_rt_q<4>(h, P, dtgamma, max_iters)));
            else if (N <= 40) TRY((launch_octet_rt_q<5>(h, P, dtgamma, max_iters)));
            else if (N <= 48) TRY((launch_octet_rt_q<6>(h, P, dtgamma, max_iters)));
            else if (N <= 56) TRY((launch_octet_rt_q<7>(h, P, dtgamma, max_iters)));
            else TRY((launch_octet_rt_q<8>(h, P, dtgamma, max_iters)));
#endif
        } else if (variant == CLB_VARIANT_LANE_PER_CELL) {
            const int cpw = (N <= 15) ? 2 : 1;  // columns per warp (one lane of each segment is a ghost)
            const int64_t warps = (P.ncol + cpw - 1) / cpw;
            const unsigned wgrid = (unsigned)((warps * 32 + kBlock - 1) / kBlock);
            if (N <= 15) {
                if (eh) DISPATCH_CMN2(h, clb::k_step_warp, 1, 16, wgrid, P, dtgamma, max_iters);
                else DISPATCH_CMN2(h, clb::k_step_warp, 0, 16, wgrid, P, dtgamma, max_iters);
            } else {
                if (eh) DISPATCH_CMN2(h, clb::k_step_warp, 1, 32, wgrid, P, dtgamma, max_iters);
                else DISPATCH_CMN2(h, clb::k_step_warp, 0, 32, wgrid, P, dtgamma, max_iters);
            }
        } else if (variant == CLB_VARIANT_REGISTER_COLUMN) {
            const clb::GridConst<15> gc = make_grid_const<15>(h);
            if (eh)
                DISPATCH_CMN(h, clb::k_eh_step_reg, 15, grid, P, gc, dtgamma, max_iters);
            else
                DISPATCH_CMN(h, clb::k_richards_step_reg, 15, grid, P, gc, dtgamma, max_iters);
        } else {
            if (eh)
                DISPATCH_CM(h, clb::k_eh_step_generic, grid, P, dtgamma, 0, max_iters);
            else
                DISPATCH_CM(h, clb::k_richards_step_generic, grid, P, dtgamma, 0, max_iters);
            clb::k_commit_state<<<grid, kBlock, 0, h->stream>>>(P);
        }
        CUDA_TRY(cudaGetLastError());
        if (stats) TRY(allreduce_doubles(h, h->d_stats, 2));
    } else {
        // tolerance path: one launch per iteration; converged flag and norm stay on the device
        CUDA_TRY(cudaMemsetAsync(h->d_flags, 0, 2 * sizeof(int32_t), h->stream));
        P.converged = h->d_flags;
        for (int it = 0; it < max_iters; ++it) {
            if (eh)
                DISPATCH_CM(h, clb::k_eh_step_generic, grid, P, dtgamma, it, it + 1);
            else
                DISPATCH_CM(h, clb::k_richards_step_generic, grid, P, dtgamma, it, it + 1);
            CUDA_TRY(cudaGetLastError());
            TRY(allreduce_doubles(h, h->d_stats, 1));
            clb::k_convergence_test<<<1, 32, 0, h->stream>>>(P, tol, h->d_stats + 2, h->d_flags + 1);
        }
        P.converged = nullptr;
        clb::k_commit_state<<<grid, kBlock, 0, h->stream>>>(P);
        CUDA_TRY(cudaGetLastError());
        if (stats) TRY(allreduce_doubles(h, h->d_stats + 1, 1));
    }
    nvtxRangePop();
    if (stats) {
        double hs[3];
        int32_t hf[2] = {0, 0};
        CUDA_TRY(cudaMemcpyAsync(hs, h->d_stats, sizeof hs, cudaMemcpyDeviceToHost, h->stream));
        if (!fixed) CUDA_TRY(cudaMemcpyAsync(hf, h->d_flags, sizeof hf, cudaMemcpyDeviceToHost, h->stream));
        CUDA_TRY(cudaStreamSynchronize(h->stream));
        stats->iterations = fixed ? iters_done : hf[1];
        stats->converged = fixed ? 0 : hf[0];
        stats->dx_norm = fixed ? std::sqrt(hs[0]) : hs[2];
        stats->nan_count = (int64_t)std::llround(hs[1]);
    }
    return CLB_OK;
}

int clb_implicit_step_host(clb_handle h, double dtgamma, int32_t max_iters, const int32_t *in_fields,
                           const double *const *in_ptrs, int32_t n_in, const int32_t *out_fields,
                           double *const *out_ptrs, int32_t n_out)
{
    TRY(check_handle(h));
    if (max_iters < 1) return fail(CLB_ERR_INVALID, "clb_implicit_step_host: max_iters must be >= 1");
    if ((n_in > 0 && (!in_fields || !in_ptrs)) || (n_out > 0 && (!out_fields || !out_ptrs)))
        return fail(CLB_ERR_INVALID, "clb_implicit_step_host: null field list");
    {
        DeviceGuard guard(h->cfg.device);
        const int rc = (h->host_route == 2) ? 0 : step_host_pipelined(h, dtgamma, max_iters, in_fields, in_ptrs, n_in, out_fields,
                                                                      out_ptrs, n_out);
        if (rc != 0) return rc < 0 ? rc : CLB_OK;
    }
    const int N = h->cfg.n_levels;
    for (int j = 0; j < n_in; ++j) {
        const bool cell = is_cell_field(in_fields[j]);
        TRY(clb_set_field(h, in_fields[j], in_ptrs[j], 1, cell ? N : 1, CLB_HOST));
    }
    TRY(clb_implicit_step(h, dtgamma, max_iters, -1.0, nullptr));
    for (int j = 0; j < n_out; ++j) {
        const bool cell = is_cell_field(out_fields[j]);
        TRY(clb_get_field(h, out_fields[j], out_ptrs[j], 1, cell ? N : 1, CLB_HOST));
    }
    return clb_sync(h);
}

// A whole soil step of EnergyHydrology from and to HOST arrays holding the state at t_n (the reference's step_u! of
// ARS111 for the soil, Simulations.jl:127-135 with the tendencies of energy_hydrology.jl:285-425, 722-906 and
// Runoff.jl:234-283): per column chunk, upload the given fields, update_aux! + PhaseChange, TOPMODEL runoff (whose
// infiltration becomes the top water flux), u + dt T_exp(u), the fused implicit stage, and the write-back of the
// chunk before -- so PCIe carries only the state and the per-column forcing, never the lagged cache.
int clb_soil_step_host(clb_handle h, double dt, int32_t max_iters, const int32_t *in_fields, const double *const *in_ptrs,
                       int32_t n_in, const int32_t *out_fields, double *const *out_ptrs, int32_t n_out)
{
    TRY(check_handle(h));
    if (max_iters < 1 || !(dt > 0.0)) return fail(CLB_ERR_INVALID, "clb_soil_step_host: dt > 0 and max_iters >= 1 expected");
    if ((n_in > 0 && (!in_fields || !in_ptrs)) || (n_out > 0 && (!out_fields || !out_ptrs)))
        return fail(CLB_ERR_INVALID, "clb_soil_step_host: null field list");
    if (h->cfg.model != CLB_ENERGY_HYDROLOGY) return fail(CLB_ERR_INVALID, "clb_soil_step_host: EnergyHydrology only");
    const int N = h->cfg.n_levels;
    {
        DeviceGuard guard(h->cfg.device);
        const int rc = (h->host_route == 2) ? 0 : step_host_pipelined(h, dt, max_iters, in_fields, in_ptrs, n_in, out_fields,
                                                                      out_ptrs, n_out, dt);
        if (rc != 0) return rc < 0 ? rc : CLB_OK;
    }
    for (int j = 0; j < n_in; ++j) {
        if (!is_cell_field(in_fields[j]) && !is_col_field(in_fields[j])) return fail(CLB_ERR_INVALID, "unknown field id %d", in_fields[j]);
        TRY(clb_set_field(h, in_fields[j], in_ptrs[j], 1, is_cell_field(in_fields[j]) ? N : 1, CLB_HOST));
    }
    {
        DeviceGuard guard(h->cfg.device);
        TRY(whole_step_ready(h));
        TRY(launch_explicit_chunk(h, make_view(h), 0, h->cfg.n_columns, dt));
    }
    TRY(clb_implicit_step(h, dt, max_iters, -1.0, nullptr));
    for (int j = 0; j < n_out; ++j) {
        if (!is_cell_field(out_fields[j]) && !is_col_field(out_fields[j])) return fail(CLB_ERR_INVALID, "unknown field id %d", out_fields[j]);
        TRY(clb_get_field(h, out_fields[j], out_ptrs[j], 1, is_cell_field(out_fields[j]) ? N : 1, CLB_HOST));
    }
    return clb_sync(h);
}

// the whole soil step on resident state: see clb_soil_step_host for the sequence
int clb_soil_step(clb_handle h, double dt, int32_t max_iters)
{
    TRY(check_handle(h));
    if (max_iters < 1 || !(dt > 0.0)) return fail(CLB_ERR_INVALID, "clb_soil_step: dt > 0 and max_iters >= 1 expected");
    if (h->cfg.model != CLB_ENERGY_HYDROLOGY) return fail(CLB_ERR_INVALID, "clb_soil_step: EnergyHydrology only");
    {
        DeviceGuard guard(h->cfg.device);
        TRY(whole_step_ready(h));
        TRY(launch_explicit_chunk(h, make_view(h), 0, h->cfg.n_columns, dt));
    }
    return clb_implicit_step(h, dt, max_iters, -1.0, nullptr);
}

#ifdef CLB_PHASE_CLOCKS
// tuning builds only: the lane kernels' phase clocks summed over warps since the last call (then reset)
__attribute__((visibility("default"))) int clb_debug_phase_clocks(unsigned long long *out16)
{
    cudaDeviceSynchronize();
    cudaMemcpyFromSymbol(out16, clb::g_phase_clk, sizeof(unsigned long long) * 16);
    unsigned long long zero[16] = {};
    cudaMemcpyToSymbol(clb::g_phase_clk, zero, sizeof(zero));
    return CLB_OK;
}
#endif

int clb_last_variant(clb_handle h, int32_t *variant)
{
    TRY(check_handle(h));
    if (!variant) return fail(CLB_ERR_INVALID, "clb_last_variant: null output");
    *variant = h->last_variant;
    return CLB_OK;
}

int clb_column_integral(clb_handle h, int32_t cell_field, int32_t col_field_out)
{
    TRY(check_handle(h));
    if (!is_cell_field(cell_field) || !is_col_field(col_field_out))
        return fail(CLB_ERR_INVALID, "clb_column_integral: (cell field, column field) expected");
    if (!h->grid_set) return fail(CLB_ERR_UNSET, "clb_set_grid was never called");
    if (!h->field[cell_field]) return fail(CLB_ERR_UNSET, "clb_column_integral: field %d was never set", cell_field);
    DeviceGuard guard(h->cfg.device);
    TRY(alloc_fields(h, {col_field_out}));
    const clb::DevView P = make_view(h);
    clb::k_column_integral<<<grid_for(P.ncol), kBlock, 0, h->stream>>>(P, h->field[cell_field], h->field[col_field_out]);
    CUDA_TRY(cudaGetLastError());
    return CLB_OK;
}

int clb_global_balance(clb_handle h, double *out4)
{
    TRY(check_handle(h));
    if (!out4) return fail(CLB_ERR_INVALID, "clb_global_balance: null output");
    if (!h->grid_set) return fail(CLB_ERR_UNSET, "clb_set_grid was never called");
    TRY(require(h, {CLB_F_Y_THETA_L}, "clb_global_balance"));
    DeviceGuard guard(h->cfg.device);
    TRY(alloc_fields(h, {CLB_F_Y_INTF_W}));
    if (h->cfg.model == CLB_ENERGY_HYDROLOGY) {
        TRY(require(h, {CLB_F_Y_RHO_E_INT, CLB_F_Y_THETA_I}, "clb_global_balance"));
        TRY(alloc_fields(h, {CLB_F_Y_INTF_E}));
    }
    double *acc = h->d_stats + 3;
    CUDA_TRY(cudaMemsetAsync(acc, 0, 4 * sizeof(double), h->stream));
    const clb::DevView P = make_view(h);
    k_balance<<<grid_for(P.ncol), kBlock, 0, h->stream>>>(P, h->field_set[CLB_F_AREA_WEIGHT] ? h->field[CLB_F_AREA_WEIGHT] : nullptr,
                                                        acc);
    CUDA_TRY(cudaGetLastError());
    TRY(allreduce_doubles(h, acc, 4));
    CUDA_TRY(cudaMemcpyAsync(out4, acc, 4 * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    CUDA_TRY(cudaStreamSynchronize(h->stream));
    return CLB_OK;
}

int clb_comm_unique_id(void *id128)
{
    if (!id128) return fail(CLB_ERR_INVALID, "clb_comm_unique_id: null output");
    TRY(load_nccl());
    NcclUniqueId id;
    NCCL_TRY(g_nccl.GetUniqueId(&id));
    std::memcpy(id128, &id, sizeof id);
    return CLB_OK;
}

int clb_comm_init(clb_handle h, const void *id128, int32_t n_ranks, int32_t rank)
{
    TRY(check_handle(h));
    if (!id128 || n_ranks < 1 || rank < 0 || rank >= n_ranks) return fail(CLB_ERR_INVALID, "clb_comm_init: bad arguments");
    TRY(load_nccl());
    DeviceGuard guard(h->cfg.device);
    NcclUniqueId id;
    std::memcpy(&id, id128, sizeof id);
    NCCL_TRY(g_nccl.CommInitRank(&h->comm, n_ranks, id, rank));
    h->n_ranks = n_ranks;
    h->rank = rank;
    return CLB_OK;
}

}  // extern "C"
