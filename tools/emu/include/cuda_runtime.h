// Host stand-in for <cuda_runtime.h> -- TEST TOOLING (tools/emu): lets g++ compile the plain per-pixel kernels of
// csrc/taxim_shadow_kernel.cu for the CPU so that their logic can be compared with the canonical restatement without a GPU.
// Every CUDA arithmetic intrinsic used there is a single IEEE-754 binary32 operation, which the host reproduces exactly
// (-ffp-contract=off, fmaf for the fused multiply-add). One "thread" runs at a time: atomics are plain updates.
#pragma once
#include <math.h>
#include <stddef.h>
#include <stdint.h>
#include <string.h>

#define TX_EMULATE 1
#define __global__
#define __device__
#define __host__
#define __constant__
#define __forceinline__ inline __attribute__((always_inline))
#define __launch_bounds__(...)
#define __restrict__
#define __align__(n) __attribute__((aligned(n)))

struct float2 { float x, y; };
struct float4 { float x, y, z, w; };
struct int4 { int x, y, z, w; };
struct uint3 { unsigned x, y, z; };
struct dim3 { unsigned x, y, z; dim3(unsigned a = 1, unsigned b = 1, unsigned c = 1) : x(a), y(b), z(c) {} };
static inline float4 make_float4(float a, float b, float c, float d) { return float4{a, b, c, d}; }
static inline float2 make_float2(float a, float b) { return float2{a, b}; }
static inline int4 make_int4(int a, int b, int c, int d) { return int4{a, b, c, d}; }
typedef void* cudaStream_t;
typedef int cudaError_t;
#define cudaSuccess 0
extern uint3 blockIdx, threadIdx;

static inline float __fadd_rn(float a, float b) { return a + b; }
static inline float __fmul_rn(float a, float b) { return a * b; }
static inline float __fdiv_rn(float a, float b) { return a / b; }
static inline float __fmaf_rn(float a, float b, float c) { return fmaf(a, b, c); }
static inline float __fsqrt_rn(float a) { return sqrtf(a); }
static inline int __float2int_rz(float a) { return (int)a; }
static inline int __float_as_int(float a) { int i; memcpy(&i, &a, 4); return i; }
static inline unsigned __float_as_uint(float a) { unsigned i; memcpy(&i, &a, 4); return i; }
static inline float __int_as_float(int i) { float a; memcpy(&a, &i, 4); return a; }
template <class T> static inline T __ldg(const T* p) { return *p; }
static inline int min(int a, int b) { return a < b ? a : b; }
static inline int max(int a, int b) { return a > b ? a : b; }
static inline int atomicMin(int* p, int v) { int o = *p; if (v < o) *p = v; return o; }
static inline unsigned atomicMax(unsigned* p, unsigned v) { unsigned o = *p; if (v > o) *p = v; return o; }
// referenced by helpers of tx_common.cuh that the emulated kernels never call
static inline size_t __cvta_generic_to_shared(const void* p) { return (size_t)p; }
static inline float __shfl_xor_sync(unsigned, float v, int) { return v; }
static inline unsigned __shfl_xor_sync(unsigned, unsigned v, int) { return v; }
