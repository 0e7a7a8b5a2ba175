// Microbenchmark: does FFMA2 (fma.rn.f32x2, sm_100a) take one issue slot for two FMAs per lane?
// Four kernels, 768 threads x 148 CTAs, ITER iterations of an unrolled body:
//   ffma        8 FFMA                      (8 accumulators)
//   ffma2       4 FFMA2                     (the same 8 accumulators as 4 register pairs)
//   ffma_alu    8 FFMA  + 4 LOP3 on independent integer registers (ALU pipe: half rate, 8 cycles)
//   ffma2_alu   4 FFMA2 + the same 4 integer operations
//   dfma        4 DFMA;   ffma_dfma   8 FFMA + 4 DFMA   (does an FP64 instruction hold the issue port for two cycles?)
// Prints cycles per iteration per SM sub-partition (6 warps each).  build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ unsigned long long pk(float a, float b)
{
    unsigned long long r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b));
    return r;
}
__device__ __forceinline__ void fma2(unsigned long long& acc, float x, float y)
{
    asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(acc) : "l"(pk(x, x)), "l"(pk(y, y)));
}
__device__ __forceinline__ void fma1(float& acc, float x, float y) { asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(acc) : "f"(x), "f"(y)); }
__device__ __forceinline__ void alu(uint32_t& b, uint32_t c) { asm volatile("lop3.b32 %0, %0, %1, %0, 0x96;" : "+r"(b) : "r"(c)); }

template <int MODE> __global__ void __launch_bounds__(768, 1) K(float* out, uint32_t* iout, int iters, float x, float y, uint32_t c, long long* cyc)
{
    float a0 = threadIdx.x, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
    unsigned long long p0 = pk(a0, a1), p1 = pk(a2, a3), p2 = pk(a4, a5), p3 = pk(a6, a7);
    uint32_t b0 = threadIdx.x, b1 = b0 + 1, b2 = b0 + 2, b3 = b0 + 3;
    const long long t0 = clock64();
    for (int i = 0; i < (MODE >= 4 ? 0 : iters); ++i) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            if (MODE == 0 || MODE == 2) {
                fma1(a0, x, y); fma1(a1, x, y); fma1(a2, x, y); fma1(a3, x, y);
                fma1(a4, x, y); fma1(a5, x, y); fma1(a6, x, y); fma1(a7, x, y);
            } else {
                fma2(p0, x, y); fma2(p1, x, y); fma2(p2, x, y); fma2(p3, x, y);
            }
            if (MODE >= 2) { alu(b0, c); alu(b1, c); alu(b2, c); alu(b3, c); }
        }
    }
    const long long t1 = clock64();
    double d0 = threadIdx.x, d1 = d0 + 1, d2 = d0 + 2, d3 = d0 + 3;
    const double dx = (double)x, dy = (double)y;
    const long long t2 = clock64();
    if (MODE >= 4) {
        for (int i = 0; i < iters; ++i) {
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                if (MODE == 5) {
                    fma1(a0, x, y); fma1(a1, x, y); fma1(a2, x, y); fma1(a3, x, y);
                    fma1(a4, x, y); fma1(a5, x, y); fma1(a6, x, y); fma1(a7, x, y);
                }
                asm volatile("fma.rn.f64 %0, %0, %1, %2;" : "+d"(d0) : "d"(dx), "d"(dy));
                asm volatile("fma.rn.f64 %0, %0, %1, %2;" : "+d"(d1) : "d"(dx), "d"(dy));
                asm volatile("fma.rn.f64 %0, %0, %1, %2;" : "+d"(d2) : "d"(dx), "d"(dy));
                asm volatile("fma.rn.f64 %0, %0, %1, %2;" : "+d"(d3) : "d"(dx), "d"(dy));
            }
        }
    }
    const long long t3 = clock64();
    a0 += (float)(d0 + d1 + d2 + d3);
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = MODE >= 4 ? t3 - t2 : t1 - t0;
    out[blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7 + (float)(p0 ^ p1 ^ p2 ^ p3);
    iout[blockIdx.x * blockDim.x + threadIdx.x] = b0 ^ b1 ^ b2 ^ b3;
}

int main()
{
    float* out; uint32_t* iout; long long* cyc; long long h;
    cudaMalloc(&out, 148 * 768 * 4); cudaMalloc(&iout, 148 * 768 * 4); cudaMalloc(&cyc, 8);
    const int iters = 20000;
    const char* names[6] = {"ffma (8 FFMA)", "ffma2 (4 FFMA2)", "ffma_alu (8 FFMA + 4 ALU)", "ffma2_alu (4 FFMA2 + 4 ALU)", "dfma (4 DFMA)", "ffma_dfma (8 FFMA + 4 DFMA)"};
    for (int m = 0; m < 6; ++m) {
        for (int rep = 0; rep < 2; ++rep) {
            if (m == 0) K<0><<<148, 768>>>(out, iout, iters, 1.0001f, 0.5f, 0x9e3779b9u, cyc);
            if (m == 1) K<1><<<148, 768>>>(out, iout, iters, 1.0001f, 0.5f, 0x9e3779b9u, cyc);
            if (m == 2) K<2><<<148, 768>>>(out, iout, iters, 1.0001f, 0.5f, 0x9e3779b9u, cyc);
            if (m == 3) K<3><<<148, 768>>>(out, iout, iters, 1.0001f, 0.5f, 0x9e3779b9u, cyc);
            if (m == 4) K<4><<<148, 768>>>(out, iout, iters, 1.0001f, 0.5f, 0x9e3779b9u, cyc);
            if (m == 5) K<5><<<148, 768>>>(out, iout, iters, 1.0001f, 0.5f, 0x9e3779b9u, cyc);
            cudaDeviceSynchronize();
        }
        cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
        // per SMSP: 6 warps; body per iteration = 8 unrolled groups
        printf("%-30s %8.2f cycles per unrolled group per SMSP (6 warps)  -> %.3f cycles per warp-group\n", names[m], (double)h / iters / 8.0,
               (double)h / iters / 8.0 / 6.0);
    }
    printf("err=%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
