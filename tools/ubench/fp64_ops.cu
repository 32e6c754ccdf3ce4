// Throughput of the individual FP64 instructions on a B200 sub-partition: cycles per warp-instruction for DFMA,
// DMUL, DADD, DSETP(+FSEL), MUFU.RCP64H, and 64-bit SHFL / LDS, each as ILP independent chains per thread with W
// warps per sub-partition.  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp64_ops fp64_ops.cu
#include <cstdio>
#include <cuda_runtime.h>
template <int OP, int ILP>
__global__ void k(double *out, double a, double b, int iters, long long *cyc)
{
    __shared__ double sm[2048];
    for (int i = threadIdx.x; i < 2048; i += blockDim.x) sm[i] = 1.0 + i * 1e-6;
    double x[ILP], y[ILP], z[ILP];
#pragma unroll
    for (int j = 0; j < ILP; ++j) {
        x[j] = 1.0 + threadIdx.x * 1e-6 + j * 1e-3;
        y[j] = 0.999 + threadIdx.x * 1e-9 + j * 1e-7;   // distinct register operands (no constant bank, no reuse)
        z[j] = 1e-7 * (1 + j) + threadIdx.x * 1e-12;
    }
    __syncthreads();
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int r = 0; r < 8; ++r) {
#pragma unroll
            for (int j = 0; j < ILP; ++j) {
                if (OP == 0) x[j] = fma(x[j], a, b);
                if (OP == 1) x[j] = x[j] * a;
                if (OP == 2) x[j] = x[j] + b;
                if (OP == 3) x[j] = (x[j] > a) ? b : x[j] + 0.0 * b;                       // DSETP + select (+ DADD folded?)
                if (OP == 4) {                                                             // MUFU.RCP64H seed
                    int hi = __double2hiint(x[j]);
                    double s;
                    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(s) : "d"(x[j]));
                    x[j] = s + __hiloint2double(hi & 0, 0);
                }
                if (OP == 5) x[j] = __shfl_xor_sync(0xffffffffu, x[j], 8);                 // 2 SHFL
                if (OP == 6) x[j] = sm[(__double2loint(x[j]) + threadIdx.x) & 2047];       // LDS.64 (+ address)
                if (OP == 7) x[j] = fmax(x[j], a);                                         // DSETP + 2 FSEL (or DMNMX?)
                if (OP == 8) x[j] = fma(x[j], y[j], z[j]);                                 // DFMA, three register operands
                if (OP == 9) x[j] = fma(y[(j + 1) % ILP], z[(j + 3) % ILP], x[j]);         // ... from other chains' registers
                if (OP == 10) x[j] = fma(x[j], y[j], b);                                   // DFMA reg, reg, constant
                if (OP == 11) x[j] = __dmul_rn(x[j], y[j]);                                // DMUL reg, reg
                if (OP == 12) x[j] = __dadd_rn(x[j], z[j]);                                // DADD reg, reg
                if (OP == 13) x[j] = fma(x[j], x[j], x[j]);                                // DFMA, one register three times
            }
        }
    }
    long long t1 = clock64();
    double s = 0;
#pragma unroll
    for (int j = 0; j < ILP; ++j) s += x[j] + y[j] + z[j];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
template <int OP, int ILP>
void run(const char *name, int wps, double *out, long long *dc)
{
    const int iters = 1000;
    k<OP, ILP><<<148, wps * 128>>>(out, 0.999999, 1e-7, iters, dc);
    cudaDeviceSynchronize();
    long long c;
    cudaMemcpy(&c, dc, sizeof c, cudaMemcpyDeviceToHost);
    printf("%-22s W=%d ILP=%d  cycles per op per sub-partition = %.2f\n", name, wps, ILP, c / ((double)iters * 8 * ILP * wps));
}
int main()
{
    double *out; long long *dc;
    cudaMalloc(&out, 148 * 1024 * 8); cudaMalloc(&dc, 8);
    for (int w = 1; w <= 2; ++w) {
        run<0, 8>("DFMA", w, out, dc);
        run<8, 8>("DFMA reg,reg,reg", w, out, dc);
        run<9, 8>("DFMA reg,reg,reg (x)", w, out, dc);
        run<8, 4>("DFMA reg,reg,reg ILP4", w, out, dc);
        run<10, 8>("DFMA reg,reg,const", w, out, dc);
        run<11, 8>("DMUL reg,reg", w, out, dc);
        run<12, 8>("DADD reg,reg", w, out, dc);
        run<13, 8>("DFMA r,r,r same reg", w, out, dc);
        run<1, 8>("DMUL", w, out, dc);
        run<2, 8>("DADD", w, out, dc);
        run<3, 8>("DSETP+select", w, out, dc);
        run<7, 8>("fmax", w, out, dc);
        run<4, 8>("MUFU.RCP64H", w, out, dc);
        run<5, 8>("SHFL x2 (double)", w, out, dc);
        run<6, 8>("LDS.64", w, out, dc);
    }
    return 0;
}
