// Is the 64-bit register-operand bandwidth shared with the other pipes?  DFMA with three distinct register operands
// (3.0 cycles alone, fp64_ops.cu) accompanied by K FFMA with three distinct register operands (FMA pipe) or K LOP3
// with three register operands (ALU pipe): cycles per DFMA per sub-partition.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp64_rf fp64_rf.cu
#include <cstdio>
#include <cuda_runtime.h>
template <int ILP, int K, int KIND, int DREG>
__global__ void k(double *out, double a, double b, int iters, long long *cyc, int seed)
{
    double x[ILP], y[ILP], z[ILP];
    float f[ILP * (K > 0 ? K : 1)], g[ILP], h[ILP];
    unsigned u[ILP * (K > 0 ? K : 1)], v[ILP], w[ILP];
#pragma unroll
    for (int j = 0; j < ILP; ++j) {
        x[j] = 1.0 + threadIdx.x * 1e-6 + j * 1e-3; y[j] = 0.999 + j * 1e-7 + threadIdx.x * 1e-9; z[j] = 1e-7 * (1 + j);
        g[j] = 0.999f + j * 1e-4f + threadIdx.x * 1e-6f; h[j] = 0.25f + j; v[j] = seed * 3 + j + threadIdx.x; w[j] = seed + 7 * j;
    }
#pragma unroll
    for (int j = 0; j < ILP * (K > 0 ? K : 1); ++j) { f[j] = seed * 0.5f + j; u[j] = seed + j * 5 + threadIdx.x; }
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int r = 0; r < 8; ++r) {
#pragma unroll
            for (int j = 0; j < ILP; ++j) {
                x[j] = DREG == 3 ? fma(x[j], y[j], z[j]) : fma(x[j], a, b);
#pragma unroll
                for (int q = 0; q < K; ++q) {
                    const int s = j * K + q;
                    if (KIND == 0) f[s] = fmaf(f[s], g[j], h[j]);             // FFMA r, r, r
                    if (KIND == 1) u[s] = (u[s] & v[j]) ^ w[j];               // LOP3 r, r, r
                }
            }
        }
    }
    long long t1 = clock64();
    double s = 0;
#pragma unroll
    for (int j = 0; j < ILP; ++j) s += x[j] + y[j] + z[j] + g[j] + h[j] + v[j] + w[j];
#pragma unroll
    for (int j = 0; j < ILP * (K > 0 ? K : 1); ++j) s += f[j] + u[j];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
template <int ILP, int K, int KIND, int DREG>
void run(const char *name, int wps, double *out, long long *dc)
{
    const int iters = 1000;
    k<ILP, K, KIND, DREG><<<148, wps * 128>>>(out, 0.999999, 1e-7, iters, dc, 3);
    cudaDeviceSynchronize();
    long long c;
    cudaMemcpy(&c, dc, sizeof c, cudaMemcpyDeviceToHost);
    printf("%-40s W=%d  cycles per DFMA per sub-partition = %.2f\n", name, wps, c / ((double)iters * 8 * ILP * wps));
}
int main()
{
    double *out; long long *dc;
    cudaMalloc(&out, 148 * 1024 * 8); cudaMalloc(&dc, 8);
    for (int w = 1; w <= 2; ++w) {
        run<4, 0, 0, 1>("DFMA r,c,c", w, out, dc);
        run<4, 0, 0, 3>("DFMA r,r,r", w, out, dc);
        run<4, 1, 0, 1>("DFMA r,c,c + 1 FFMA r,r,r", w, out, dc);
        run<4, 1, 0, 3>("DFMA r,r,r + 1 FFMA r,r,r", w, out, dc);
        run<4, 2, 0, 3>("DFMA r,r,r + 2 FFMA r,r,r", w, out, dc);
        run<4, 1, 1, 3>("DFMA r,r,r + 1 LOP3 r,r,r", w, out, dc);
        run<4, 2, 1, 3>("DFMA r,r,r + 2 LOP3 r,r,r", w, out, dc);
        run<4, 2, 1, 1>("DFMA r,c,c + 2 LOP3 r,r,r", w, out, dc);
    }
    return 0;
}
