// layout_kernels.cuh -- caller layout (any strides, usually ClimaCore's level-fastest
// `parent(field)`) <-> the library's column-fastest SoA mirrors.  A 32-column tile is
// staged through shared memory so both the caller-side and the mirror-side accesses are
// coalesced when the caller is level-fastest.  `idx` (may be null) maps handle column ->
// caller column (land-sea mask compaction, mask_test.jl:53-61): columns outside it are
// never read or written.
#pragma once
#include <stdint.h>

namespace clb {

constexpr int kTileCols = 32;

// mirror[i*ld + c] = src[i*sl + col(c)*sc]
__global__ void __launch_bounds__(256) k_gather_cells(double *__restrict__ mirror, int64_t ld,
                                                      const double *__restrict__ src, int64_t sl, int64_t sc,
                                                      const int64_t *__restrict__ idx, int N, int64_t ncol)
{
    extern __shared__ double tile[];  // [N][kTileCols + 1]
    const int64_t c0 = (int64_t)blockIdx.x * kTileCols;
    const int ncl = (int)min((int64_t)kTileCols, ncol - c0);
    for (int e = threadIdx.x; e < ncl * N; e += blockDim.x) {
        const int cl = e / N, i = e - cl * N;
        const int64_t col = idx ? idx[c0 + cl] : c0 + cl;
        tile[i * (kTileCols + 1) + cl] = src[(int64_t)i * sl + col * sc];
    }
    __syncthreads();
    for (int e = threadIdx.x; e < kTileCols * N; e += blockDim.x) {
        const int i = e / kTileCols, cl = e - i * kTileCols;
        if (cl < ncl) mirror[(int64_t)i * ld + c0 + cl] = tile[i * (kTileCols + 1) + cl];
    }
}

// dst[i*sl + col(c)*sc] = mirror[i*ld + c]
__global__ void __launch_bounds__(256) k_scatter_cells(const double *__restrict__ mirror, int64_t ld,
                                                       double *__restrict__ dst, int64_t sl, int64_t sc,
                                                       const int64_t *__restrict__ idx, int N, int64_t ncol)
{
    extern __shared__ double tile[];
    const int64_t c0 = (int64_t)blockIdx.x * kTileCols;
    const int ncl = (int)min((int64_t)kTileCols, ncol - c0);
    for (int e = threadIdx.x; e < kTileCols * N; e += blockDim.x) {
        const int i = e / kTileCols, cl = e - i * kTileCols;
        if (cl < ncl) tile[i * (kTileCols + 1) + cl] = mirror[(int64_t)i * ld + c0 + cl];
    }
    __syncthreads();
    for (int e = threadIdx.x; e < ncl * N; e += blockDim.x) {
        const int cl = e / N, i = e - cl * N;
        const int64_t col = idx ? idx[c0 + cl] : c0 + cl;
        dst[(int64_t)i * sl + col * sc] = tile[i * (kTileCols + 1) + cl];
    }
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

__global__ void __launch_bounds__(256) k_fill(double *__restrict__ p, int64_t n, double v)
{
    const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k < n) p[k] = v;
}

}  // namespace clb
