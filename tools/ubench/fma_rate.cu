// FP32 issue-rate microbenchmark for sm_100a: scalar FFMA (3 registers), FFMA with a constant-bank operand, FADD+FFMA pairs,
// packed fma.rn.f32x2 / add.rn.f32x2. Prints FMA lanes per clock per SM for each variant.
#include <cstdio>
#include <cuda_runtime.h>
__constant__ float c_t[64];
template <int V>
__global__ void k(float* out, float a, float b, int iters, long long* cyc)
{
    float acc[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) acc[i] = threadIdx.x * 0.001f + i;
    float x0 = a + threadIdx.x, x1 = b + threadIdx.x;
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int r = 0; r < 8; ++r) {
            if (V == 0) {
#pragma unroll
                for (int i = 0; i < 16; ++i) acc[i] = __fmaf_rn(x0, x1, acc[i]);
            } else if (V == 1) {
#pragma unroll
                for (int i = 0; i < 16; ++i) acc[i] = __fmaf_rn(c_t[r * 4 + (i & 3)], x0, acc[i]);
            } else if (V == 2) { // pair: FADD then FFMA with a constant tap (the vertical pass)
#pragma unroll
                for (int i = 0; i < 16; ++i) acc[i] = __fmaf_rn(c_t[r], __fadd_rn(x0, acc[(i + 1) & 15]), acc[i]);
            } else if (V == 3) { // packed f32x2 FMA: 8 instructions = 16 FMAs
#pragma unroll
                for (int i = 0; i < 16; i += 2) {
                    unsigned long long A, B, C;
                    asm("mov.b64 %0, {%1, %2};" : "=l"(A) : "f"(x0), "f"(x1));
                    asm("mov.b64 %0, {%1, %2};" : "=l"(B) : "f"(x1), "f"(x0));
                    asm("mov.b64 %0, {%1, %2};" : "=l"(C) : "f"(acc[i]), "f"(acc[i + 1]));
                    asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(C) : "l"(A), "l"(B));
                    asm("mov.b64 {%0, %1}, %2;" : "=f"(acc[i]), "=f"(acc[i + 1]) : "l"(C));
                }
            }
        }
    }
    long long t1 = clock64();
    float s = 0;
#pragma unroll
    for (int i = 0; i < 16; ++i) s += acc[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
template <int V>
void run(const char* name, int fma_per_iter, int threads)
{
    float* out; long long* cyc;
    cudaMalloc(&out, 148 * 1024 * 4); cudaMalloc(&cyc, 148 * 8);
    const int iters = 2000;
    k<V><<<148, threads>>>(out, 1.0f, 2.0f, 10, cyc);
    k<V><<<148, threads>>>(out, 1.0f, 2.0f, iters, cyc);
    long long h[148]; cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
    double c = 0; for (int i = 0; i < 148; ++i) c += h[i]; c /= 148;
    printf("%-34s threads/SM %4d: %.1f FMA-lanes/clk/SM  (%.2f warp-instr/clk/SMSP)\n", name, threads,
           (double)fma_per_iter * iters * threads / c, (double)fma_per_iter * iters * threads / c / 128.0);
    cudaFree(out); cudaFree(cyc);
}
int main()
{
    float t[64]; for (int i = 0; i < 64; ++i) t[i] = 1.0f / (i + 3);
    cudaMemcpyToSymbol(c_t, t, sizeof(t));
    for (int th : {128, 256, 512, 1024}) {
        run<0>("FFMA reg,reg,reg", 128, th);
        run<1>("FFMA const,reg,reg", 128, th);
        run<2>("FADD + FFMA const (2 instr/tap pair)", 128, th);
        run<3>("fma.rn.f32x2 (per scalar FMA)", 128, th);
    }
    return 0;
}
