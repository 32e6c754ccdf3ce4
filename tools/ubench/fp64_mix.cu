// FP64 pipe vs issue slots: cycles per DFMA warp-instruction per SM sub-partition when every DFMA is
// accompanied by K independent non-FP64 instructions (integer LOP3/IADD3, FSEL-like selects, or LDS).
// Answers: does a DFMA hold the issue port for its 2 pipe cycles (then K = 1 costs 3 cycles / DFMA) or can
// other pipes issue in the shadow (then K = 1 stays at 2)?
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp64_mix fp64_mix.cu
#include <cstdio>
#include <cuda_runtime.h>
template <int ILP, int K, int KIND>
__global__ void k(double *out, double a, double b, int iters, long long *cyc, int seed)
{
    __shared__ double sm[1024];
    for (int i = threadIdx.x; i < 1024; i += blockDim.x) sm[i] = i * 0.5;
    double x[ILP];
    unsigned u[ILP * (K > 0 ? K : 1)];
    float f[ILP * (K > 0 ? K : 1)];
#pragma unroll
    for (int j = 0; j < ILP; ++j) x[j] = threadIdx.x * 1e-3 + j;
#pragma unroll
    for (int j = 0; j < ILP * (K > 0 ? K : 1); ++j) { u[j] = seed + j * 7 + threadIdx.x; f[j] = seed * 0.5f + j; }
    __syncthreads();
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int r = 0; r < 8; ++r) {
#pragma unroll
            for (int j = 0; j < ILP; ++j) {
                x[j] = fma(x[j], a, b);
#pragma unroll
                for (int q = 0; q < K; ++q) {
                    const int s = j * K + q;
                    if (KIND == 0) u[s] = (u[s] ^ (u[s] >> 3)) + 0x9e3779b9u;          // SHF + LOP3/IADD3 (ALU)
                    if (KIND == 1) f[s] = fmaf(f[s], 0.999f, 0.25f);                     // FFMA (FMA pipe)
                    if (KIND == 2) u[s] = __float_as_uint(sm[(u[s] + r) & 1023] > 3.0 ? 1.f : 2.f) + u[s];  // LDS.64 + DSETP ...
                }
            }
        }
    }
    long long t1 = clock64();
    double s = 0;
#pragma unroll
    for (int j = 0; j < ILP; ++j) s += x[j];
#pragma unroll
    for (int j = 0; j < ILP * (K > 0 ? K : 1); ++j) s += u[j] + f[j];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
template <int ILP, int K, int KIND>
void run(int wps, double *out, long long *dc)
{
    const int iters = 2000;
    k<ILP, K, KIND><<<148, wps * 128>>>(out, 0.999999, 1e-7, iters, dc, 3);
    cudaDeviceSynchronize();
    long long c;
    cudaMemcpy(&c, dc, sizeof c, cudaMemcpyDeviceToHost);
    const double inst = (double)iters * 8 * ILP * wps;
    printf("W=%d ILP=%d K=%d kind=%d  cycles/DFMA/SMSP = %.2f\n", wps, ILP, K, KIND, c / inst);
}
int main()
{
    double *out; long long *dc;
    cudaMalloc(&out, 148 * 1024 * 8); cudaMalloc(&dc, 8);
    for (int w = 1; w <= 4; ++w) {
        run<4, 0, 0>(w, out, dc);
        run<4, 1, 0>(w, out, dc);   // 2 ALU instrs (SHF/LOP3 + IADD) per DFMA
        run<4, 2, 0>(w, out, dc);
        run<4, 1, 1>(w, out, dc);   // 1 FFMA per DFMA
        run<4, 2, 1>(w, out, dc);
        run<4, 4, 1>(w, out, dc);
        run<8, 0, 0>(w, out, dc);
        run<8, 1, 1>(w, out, dc);
    }
    return 0;
}
