// FP64 pipe microbenchmark: cycles per DFMA warp-instruction per SM sub-partition as a function of
// resident warps per sub-partition (W) and independent chains per thread (ILP).
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp64_ilp fp64_ilp.cu
#include <cstdio>
#include <cuda_runtime.h>
template <int ILP>
__global__ void k(double *out, double a, double b, int iters, long long *cyc)
{
    double x[ILP];
#pragma unroll
    for (int j = 0; j < ILP; ++j) x[j] = threadIdx.x * 1e-3 + j;
    __syncthreads();
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int r = 0; r < 16; ++r) {
#pragma unroll
            for (int j = 0; j < ILP; ++j) x[j] = fma(x[j], a, b);
        }
    }
    long long t1 = clock64();
    double s = 0;
#pragma unroll
    for (int j = 0; j < ILP; ++j) s += x[j];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
template <int ILP>
void run(int warps_per_smsp, double *out, long long *dc)
{
    const int iters = 2000;
    const int threads = warps_per_smsp * 4 * 32;  // one block per SM, warps spread over the 4 sub-partitions
    k<ILP><<<148, threads>>>(out, 0.999999, 1e-7, iters, dc);
    cudaDeviceSynchronize();
    long long c;
    cudaMemcpy(&c, dc, sizeof c, cudaMemcpyDeviceToHost);
    const double inst = (double)iters * 16 * ILP * warps_per_smsp;  // DFMA warp-instructions per sub-partition
    printf("W=%d ILP=%d  cycles/DFMA/SMSP = %.2f   (chain latency if serial: %.1f)\n", warps_per_smsp, ILP, c / inst,
           (double)c / (iters * 16.0));
}
int main()
{
    double *out; long long *dc;
    cudaMalloc(&out, 148 * 1024 * sizeof(double));
    cudaMalloc(&dc, sizeof(long long));
    for (int w : {1, 2, 4, 8}) {
        run<1>(w, out, dc); run<2>(w, out, dc); run<3>(w, out, dc); run<4>(w, out, dc); run<6>(w, out, dc); run<8>(w, out, dc);
    }
    return 0;
}
