// layout_kernels.cuh -- strided copies between a caller array (any strides, usually
// ClimaCore's level-fastest `parent(field)`) and a library mirror (column-fastest or
// level-fastest).  A 32-column tile is staged through shared memory and each side is
// traversed in its own fastest direction, so both the reads and the writes coalesce.
// `idx` maps handle column -> caller column (land-sea mask compaction,
// test/standalone/Soil/mask_test.jl:53-61): caller columns outside it are never read or
// written.
#pragma once
#include <stdint.h>

namespace clb {

constexpr int kTileCols = 32;

// dst[i*dsl + dcol(c)*dsc] = src[i*ssl + scol(c)*ssc],  c = 0..ncol-1, i = 0..N-1
__global__ void __launch_bounds__(256) k_relayout(double *__restrict__ dst, int64_t dsl, int64_t dsc,
                                                  const int64_t *__restrict__ didx, const double *__restrict__ src,
                                                  int64_t ssl, int64_t ssc, const int64_t *__restrict__ sidx, int N,
                                                  int64_t ncol)
{
    extern __shared__ double tile[];  // [N][kTileCols + 1]
    const int64_t c0 = (int64_t)blockIdx.x * kTileCols;
    const int ncl = (int)min((int64_t)kTileCols, ncol - c0);
    const bool s_level_fast = ssl <= ssc, d_level_fast = dsl <= dsc;
    for (int e = threadIdx.x; e < kTileCols * N; e += blockDim.x) {
        int cl, i;
        if (s_level_fast) { cl = e / N; i = e - cl * N; } else { i = e / kTileCols; cl = e - i * kTileCols; }
        if (cl < ncl) {
            const int64_t col = sidx ? sidx[c0 + cl] : c0 + cl;
            tile[i * (kTileCols + 1) + cl] = src[(int64_t)i * ssl + col * ssc];
        }
    }
    __syncthreads();
    for (int e = threadIdx.x; e < kTileCols * N; e += blockDim.x) {
        int cl, i;
        if (d_level_fast) { cl = e / N; i = e - cl * N; } else { i = e / kTileCols; cl = e - i * kTileCols; }
        if (cl < ncl) {
            const int64_t col = didx ? didx[c0 + cl] : c0 + cl;
            dst[(int64_t)i * dsl + col * dsc] = tile[i * (kTileCols + 1) + cl];
        }
    }
}

// The same for up to kManyFields fields in one launch (blockIdx.y = field): the pipelined host-buffer stage
// moves a column chunk of every per-step field with one kernel.
constexpr int kManyFields = 10;
struct ManyFields {
    double *dst[kManyFields];
    const double *src[kManyFields];
};
__global__ void __launch_bounds__(256) k_relayout_many(ManyFields f, int64_t dsl, int64_t dsc, int64_t ssl, int64_t ssc,
                                                       int N, int64_t ncol)
{
    extern __shared__ double tile[];  // [N][kTileCols + 1]
    double *__restrict__ dst = f.dst[blockIdx.y];
    const double *__restrict__ src = f.src[blockIdx.y];
    const int64_t c0 = (int64_t)blockIdx.x * kTileCols;
    const int ncl = (int)min((int64_t)kTileCols, ncol - c0);
    const bool s_level_fast = ssl <= ssc, d_level_fast = dsl <= dsc;
    for (int e = threadIdx.x; e < kTileCols * N; e += blockDim.x) {
        int cl, i;
        if (s_level_fast) { cl = e / N; i = e - cl * N; } else { i = e / kTileCols; cl = e - i * kTileCols; }
        if (cl < ncl) tile[i * (kTileCols + 1) + cl] = src[(int64_t)i * ssl + (c0 + cl) * ssc];
    }
    __syncthreads();
    for (int e = threadIdx.x; e < kTileCols * N; e += blockDim.x) {
        int cl, i;
        if (d_level_fast) { cl = e / N; i = e - cl * N; } else { i = e / kTileCols; cl = e - i * kTileCols; }
        if (cl < ncl) dst[(int64_t)i * dsl + (c0 + cl) * dsc] = tile[i * (kTileCols + 1) + cl];
    }
}
// Both directions in one launch, for the zero-copy host route: the first n_out fields are written back to the host
// layout (level fastest, column stride N) for `ncol_out` columns, the remaining fields are read from it for
// `ncol_in` columns.  One launch keeps PCIe reads and writes in flight together (two kernels on two streams do not
// overlap reliably: the reading kernel's blocks fill the machine).  msl / msc: strides of the mirrors.
// idx_out / idx_in (nullable): handle column -> row of the host arrays for the chunk written / read (land-sea mask
// compaction: rows of inactive columns are never touched).
__global__ void __launch_bounds__(256) k_relayout_dual(ManyFields f, int n_out, int64_t msl, int64_t msc, int N,
                                                       int64_t ncol_out, int64_t ncol_in,
                                                       const int64_t *__restrict__ idx_out, const int64_t *__restrict__ idx_in)
{
    extern __shared__ double tile[];  // [N][kTileCols + 1]
    // the field is the FAST grid dimension: consecutive blocks alternate between the two directions
    const bool to_host = (int)blockIdx.x < n_out;
    const int64_t ncol = to_host ? ncol_out : ncol_in;
    const int64_t c0 = (int64_t)blockIdx.y * kTileCols;
    if (c0 >= ncol) return;  // whole block
    double *__restrict__ dst = f.dst[blockIdx.x];
    const double *__restrict__ src = f.src[blockIdx.x];
    const int64_t ssl = to_host ? msl : 1, ssc = to_host ? msc : N, dsl = to_host ? 1 : msl, dsc = to_host ? N : msc;
    const int64_t *__restrict__ sidx = to_host ? nullptr : idx_in, *__restrict__ didx = to_host ? idx_out : nullptr;
    const int ncl = (int)min((int64_t)kTileCols, ncol - c0);
    const bool s_level_fast = ssl <= ssc, d_level_fast = dsl <= dsc;
    for (int e = threadIdx.x; e < kTileCols * N; e += blockDim.x) {
        int cl, i;
        if (s_level_fast) { cl = e / N; i = e - cl * N; } else { i = e / kTileCols; cl = e - i * kTileCols; }
        if (cl < ncl) tile[i * (kTileCols + 1) + cl] = src[(int64_t)i * ssl + (sidx ? sidx[c0 + cl] : c0 + cl) * ssc];
    }
    __syncthreads();
    for (int e = threadIdx.x; e < kTileCols * N; e += blockDim.x) {
        int cl, i;
        if (d_level_fast) { cl = e / N; i = e - cl * N; } else { i = e / kTileCols; cl = e - i * kTileCols; }
        if (cl < ncl) dst[(int64_t)i * dsl + (didx ? didx[c0 + cl] : c0 + cl) * dsc] = tile[i * (kTileCols + 1) + cl];
    }
}
// per-column fields; sidx / didx (nullable): the caller-side index of handle column c on the source / destination side
__global__ void __launch_bounds__(256) k_copy_many(ManyFields f, int64_t n, const int64_t *__restrict__ sidx = nullptr,
                                                   const int64_t *__restrict__ didx = nullptr)
{
    const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (c < n) f.dst[blockIdx.y][didx ? didx[c] : c] = f.src[blockIdx.y][sidx ? sidx[c] : c];
}

__global__ void __launch_bounds__(256) k_gather_cols(double *__restrict__ mirror, const double *__restrict__ src,
                                                     int64_t sc, const int64_t *__restrict__ idx, int64_t ncol)
{
    const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (c < ncol) mirror[c] = src[(idx ? idx[c] : c) * sc];
}

__global__ void __launch_bounds__(256) k_scatter_cols(const double *__restrict__ mirror, double *__restrict__ dst,
                                                      int64_t sc, const int64_t *__restrict__ idx, int64_t ncol)
{
    const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (c < ncol) dst[(idx ? idx[c] : c) * sc] = mirror[c];
}

// y <- a x + b y over a whole mirror (pads included: they are never read back).  Product and sum are rounded
// separately, as Julia's broadcast `@. y + a * x` is (no muladd).
__global__ void __launch_bounds__(256) k_axpby(double *__restrict__ y, const double *__restrict__ x, double a, double b,
                                               int64_t n)
{
    const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k < n) y[k] = (b == 0.0) ? a * x[k] : __dadd_rn(__dmul_rn(b, y[k]), __dmul_rn(a, x[k]));
}

// x <- b / w entry by entry (a DiagonalMatrixRow block)
__global__ void __launch_bounds__(256) k_div(double *__restrict__ x, const double *__restrict__ b,
                                             const double *__restrict__ w, int64_t n)
{
    const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k < n) x[k] = b[k] / w[k];
}

__global__ void __launch_bounds__(256) k_fill(double *__restrict__ p, int64_t n, double v)
{
    const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k < n) p[k] = v;
}

}  // namespace clb
