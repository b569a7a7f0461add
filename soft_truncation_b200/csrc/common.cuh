// Shared helpers for libst_b200 (sm_100a only).
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "st_b200.h"

#define ST_ERR_ARG 1
#define ST_ERR_CUDA 2
#define ST_ERR_UNSUPPORTED 3

void st_set_error(const char* fmt, ...);

#define ST_CHECK_ARG(cond, ...)          \
  do {                                   \
    if (!(cond)) {                       \
      st_set_error(__VA_ARGS__);         \
      return ST_ERR_ARG;                 \
    }                                    \
  } while (0)

// Launch errors (bad configuration, missing image) surface here; execution errors surface at the
// caller's next synchronisation, as with any CUDA stream work.
#define ST_CHECK_LAUNCH(name)                                            \
  do {                                                                   \
    cudaError_t e__ = cudaGetLastError();                                \
    if (e__ != cudaSuccess) {                                            \
      st_set_error("%s: launch failed: %s", name, cudaGetErrorString(e__)); \
      return ST_ERR_CUDA;                                                \
    }                                                                    \
  } while (0)

typedef __nv_bfloat16 bf16;

__device__ __forceinline__ float to_f(float v) { return v; }
__device__ __forceinline__ float to_f(bf16 v) { return __bfloat162float(v); }
template <typename T> __device__ __forceinline__ T from_f(float v);
template <> __device__ __forceinline__ float from_f<float>(float v) { return v; }
template <> __device__ __forceinline__ bf16 from_f<bf16>(float v) { return __float2bfloat16_rn(v); }

// 4 consecutive elements <-> 4 floats (pointer must be aligned to 4 elements)
__device__ __forceinline__ void load4(const float* p, float v[4]) {
  float4 t = *reinterpret_cast<const float4*>(p);
  v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
}
__device__ __forceinline__ void load4(const bf16* p, float v[4]) {
  uint2 t = *reinterpret_cast<const uint2*>(p);
  __nv_bfloat162 a = *reinterpret_cast<__nv_bfloat162*>(&t.x);
  __nv_bfloat162 b = *reinterpret_cast<__nv_bfloat162*>(&t.y);
  v[0] = __low2float(a); v[1] = __high2float(a); v[2] = __low2float(b); v[3] = __high2float(b);
}
__device__ __forceinline__ void store4(float* p, const float v[4]) {
  *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
}
__device__ __forceinline__ void store4(bf16* p, const float v[4]) {
  __nv_bfloat162 a = __floats2bfloat162_rn(v[0], v[1]);
  __nv_bfloat162 b = __floats2bfloat162_rn(v[2], v[3]);
  uint2 t;
  t.x = *reinterpret_cast<uint32_t*>(&a);
  t.y = *reinterpret_cast<uint32_t*>(&b);
  *reinterpret_cast<uint2*>(p) = t;
}

// MUFU-based sigmoid: FMUL, MUFU.EX2, FADD, MUFU.RCP (~2 ulp).  The streaming kernels are issue-bound; the .ftz forms
// drop the denormal fix-up code that __expf/__fdividef carry without -use_fast_math (x < -87 gives exactly 0, x > 87
// exactly 1 - both correct to fp32 precision).
__device__ __forceinline__ float sigmoid_f(float x) {
  float e, s;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(-1.4426950408889634f * x));
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(s) : "f"(1.f + e));
  return s;
}
__device__ __forceinline__ float silu_f(float x) { return x * sigmoid_f(x); }
// d/dx [x*sigmoid(x)] = s + x*s*(1-s)
__device__ __forceinline__ float silu_grad_f(float x) {
  float s = sigmoid_f(x);
  return fmaf(x, fmaf(-s, s, s), s);
}

// bf16 storage paths: one MUFU per element instead of two (the SiLU streaming kernels sit at 70 % MUFU-pipe
// utilisation with ex2 + rcp).  sigmoid(x) = 0.5 + 0.5*tanh(x/2); tanh.approx has ~2^-11 absolute error, i.e. below
// the 2^-9 relative rounding of the bf16 value the result is stored as.  fp32 parity paths keep sigmoid_f.
__device__ __forceinline__ float tanh_half_f(float x) {
  float t;
  asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(0.5f * x));
  return t;
}
template <typename T> __device__ __forceinline__ float silu_t(float x) {
  if constexpr (sizeof(T) == 2) { const float h = 0.5f * x; return fmaf(h, tanh_half_f(x), h); }
  else return silu_f(x);
}
template <typename T> __device__ __forceinline__ float silu_grad_t(float x) {
  if constexpr (sizeof(T) == 2) {
    const float t = tanh_half_f(x);
    const float s = fmaf(0.5f, t, 0.5f);                 // sigmoid
    const float q = fmaf(-0.5f * t, t, 0.5f);            // 2*s*(1-s)
    return fmaf(0.5f * x, q, s);
  } else return silu_grad_f(x);
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// Counter-based RNG (Philox-4x32-7): 4 uniform 32-bit words per (seed, counter).
__device__ __forceinline__ uint4 philox4(uint64_t seed, uint64_t ctr) {
  uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
  uint32_t c0 = (uint32_t)ctr, c1 = (uint32_t)(ctr >> 32), c2 = 0x5eed5eedu, c3 = 0x0b200b20u;
#pragma unroll
  for (int r = 0; r < 7; ++r) {      // Philox-4x32-7 (Crush-resistant per the Random123 paper; dropout needs no more)
    uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
    uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
    uint32_t n0 = hi1 ^ c1 ^ k0, n1 = lo1, n2 = hi0 ^ c3 ^ k1, n3 = lo0;
    c0 = n0; c1 = n1; c2 = n2; c3 = n3;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
  return make_uint4(c0, c1, c2, c3);
}
// keep-multiplier for 4 consecutive elements starting at element index e4*4
__device__ __forceinline__ void dropout4(uint64_t seed, uint64_t e4, float p, float keep[4]) {
  uint4 r = philox4(seed, e4);
  float inv = 1.f / (1.f - p);
  uint32_t thr = (uint32_t)(p * 4294967296.0);
  keep[0] = r.x >= thr ? inv : 0.f;
  keep[1] = r.y >= thr ? inv : 0.f;
  keep[2] = r.z >= thr ? inv : 0.f;
  keep[3] = r.w >= thr ? inv : 0.f;
}

// Programmatic dependent launch: a kernel launched through st_launch may be scheduled while its predecessor in the
// stream is still draining; it must execute pdl_wait() before its first global-memory access (the wait returns when
// the predecessor grid has completed and its writes are visible).  pdl_trigger() lets the NEXT kernel be scheduled
// once every CTA of this grid has started.  Both are no-ops for a normal launch.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
const uint64_t* st_seed_offset();        // device addend of in-kernel dropout seeds, or nullptr (lib.cu)
bool st_pdl_on(cudaStream_t stream);      // ST_PDL=0 disables; never used while the stream is being captured

template <typename... KArgs, typename... Args>
inline cudaError_t st_launch(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream, Args&&... args) {
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = st_pdl_on(stream) ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

static inline int st_num_sms() {
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    if (n <= 0) n = 148;
  }
  return n;
}

// dispatch on dtype code
#define ST_DISPATCH_DTYPE(dtype, T, ...)                      \
  do {                                                        \
    if ((dtype) == ST_F32) { typedef float T; __VA_ARGS__; }  \
    else if ((dtype) == ST_BF16) { typedef bf16 T; __VA_ARGS__; } \
    else { st_set_error("bad dtype %d", (int)(dtype)); return ST_ERR_ARG; } \
  } while (0)
