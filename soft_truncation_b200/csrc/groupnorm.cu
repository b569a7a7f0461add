// GroupNorm (+SiLU, +dropout) forward and backward on NHWC tensors whose channel axis may be the
// concatenation of two tensors (the U-Net skip concatenation is never materialised).
// Reference: nn.GroupNorm(min(C//4,32), C, eps=1e-6) -> act -> Dropout, models/layerspp.py:232,244-245,
// 258,275-278; models/ncsnpp.py:219-253.
//
// All kernels are HBM-bound streaming passes.  Thread mapping: one thread owns a fixed run of 8 channels
// (16-byte bf16 / 32-byte fp32 vectors) of one image and walks that image's pixels, so per-channel constants
// (gamma, beta, mean, rstd) sit in registers and the inner loop has no integer division; two pixels are in
// flight per iteration.  Reductions are deterministic (fixed-order shared-memory trees, no float atomics).
#include <stdlib.h>

#include <type_traits>

#include "common.cuh"

namespace {

__device__ __forceinline__ void load8(const float* p, float v[8]) {
  float4 a = reinterpret_cast<const float4*>(p)[0], b = reinterpret_cast<const float4*>(p)[1];
  v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
}
__device__ __forceinline__ void load8(const bf16* p, float v[8]) {
  uint4 t = *reinterpret_cast<const uint4*>(p);
  const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&t);
#pragma unroll
  for (int i = 0; i < 4; ++i) { v[2 * i] = __low2float(h[i]); v[2 * i + 1] = __high2float(h[i]); }
}
__device__ __forceinline__ void store8(float* p, const float v[8]) {
  reinterpret_cast<float4*>(p)[0] = make_float4(v[0], v[1], v[2], v[3]);
  reinterpret_cast<float4*>(p)[1] = make_float4(v[4], v[5], v[6], v[7]);
}
__device__ __forceinline__ void store8(bf16* p, const float v[8]) {
  uint4 t;
  __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&t);
#pragma unroll
  for (int i = 0; i < 4; ++i) h[i] = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]);
  *reinterpret_cast<uint4*>(p) = t;
}

// Thread-private cp.async software pipeline.  Every thread streams its own sequence of 8-element vectors (one per
// pixel and input stream) through DEPTH shared-memory slots: the bytes in flight per SM are set by the shared-memory
// ring (DEPTH x streams x 16 B x 256 threads x resident blocks), not by registers, which is what an HBM-bound pass
// needs to cover ~1-2 us of loaded-memory latency.  Slots are private to a thread (it only reads what it copied),
// so no block barrier is involved; slot addresses are strided by the block size -> conflict-free.
template <typename T, int NS, int DEPTH>
struct Pipe {
  static constexpr int PARTS = (int)sizeof(T) * 8 / 16;              // 16-byte pieces per vector (1 bf16, 2 fp32)
  static constexpr int VEC_BYTES = DEPTH * NS * PARTS * 256 * 16;
  static constexpr int BYTES = VEC_BYTES + DEPTH * 256 * 4;          // + one 4-byte side slot per stage (keep bits)
  uint32_t base;                                                     // shared address of this thread's first slot
  uint32_t side;                                                     // ... and of its first 4-byte side slot
  __device__ __forceinline__ explicit Pipe(uint8_t* smem) {
    base = (uint32_t)__cvta_generic_to_shared(smem) + threadIdx.x * 16;
    side = (uint32_t)__cvta_generic_to_shared(smem) + VEC_BYTES + threadIdx.x * 4;
  }
  // the byte `bits[idx]` travels inside its aligned 4-byte word
  __device__ __forceinline__ void issue_byte(int stage, const uint8_t* byte) const {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(side + (uint32_t)(stage * 256 * 4)),
                 "l"(reinterpret_cast<uintptr_t>(byte) & ~(uintptr_t)3)
                 : "memory");
  }
  // `idx` = index of that byte in its (4-byte aligned) array; only its low two bits are used
  __device__ __forceinline__ uint32_t read_byte(int stage, uint32_t idx) const {
    uint32_t w;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(w) : "r"(side + (uint32_t)(stage * 256 * 4)));
    return (w >> (8 * (idx & 3u))) & 0xFFu;
  }
  __device__ __forceinline__ uint32_t slot(int stage, int stream, int part) const {
    return base + (uint32_t)(((stage * NS + stream) * PARTS + part) * 256 * 16);
  }
  __device__ __forceinline__ void issue(int stage, int stream, const T* g) const {
#pragma unroll
    for (int part = 0; part < PARTS; ++part)
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(slot(stage, stream, part)),
                   "l"(reinterpret_cast<const uint8_t*>(g) + part * 16)
                   : "memory");
  }
  __device__ __forceinline__ void read(int stage, int stream, float v[8]) const {
    if constexpr (sizeof(T) == 2) {
      uint4 t;
      asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(t.x), "=r"(t.y), "=r"(t.z), "=r"(t.w) : "r"(slot(stage, stream, 0)));
      const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&t);
#pragma unroll
      for (int i = 0; i < 4; ++i) { v[2 * i] = __low2float(h[i]); v[2 * i + 1] = __high2float(h[i]); }
    } else {
#pragma unroll
      for (int part = 0; part < 2; ++part)
        asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];"
                     : "=f"(v[4 * part]), "=f"(v[4 * part + 1]), "=f"(v[4 * part + 2]), "=f"(v[4 * part + 3])
                     : "r"(slot(stage, stream, part)));
    }
  }
  static __device__ __forceinline__ void commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
  static __device__ __forceinline__ void wait() { asm volatile("cp.async.wait_group %0;" ::"n"(DEPTH - 1) : "memory"); }
};
constexpr int GN_DEPTH = 8;       // slots per thread (forward passes, backward reduction)
constexpr int GN_BWD_DEPTH = 4;   // backward apply streams up to 5 inputs per pixel

// keep-multipliers (0 or 1/(1-p)) of 8 consecutive elements: one Philox call, 16 random bits per element.
// Returns the 8 keep flags as a byte (bit i = element i kept) so that the backward pass can reload them
// (1 byte per 8 elements) instead of re-running the generator twice.
__device__ __forceinline__ uint32_t dropout8(uint64_t seed, uint64_t oct, float p, float keep[8]) {
  uint4 r = philox4(seed, oct);
  const float inv = 1.f / (1.f - p);
  const uint32_t thr = (uint32_t)(p * 65536.f);
  const uint32_t w[4] = {r.x, r.y, r.z, r.w};
  uint32_t bits = 0;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const bool k0 = (w[i] & 0xFFFFu) >= thr, k1 = (w[i] >> 16) >= thr;
    keep[2 * i] = k0 ? inv : 0.f;
    keep[2 * i + 1] = k1 ? inv : 0.f;
    bits |= (k0 ? 1u : 0u) << (2 * i) | (k1 ? 1u : 0u) << (2 * i + 1);
  }
  return bits;
}
__device__ __forceinline__ void keep_from_bits(uint32_t bits, float p, float keep[8]) {
  const float inv = 1.f / (1.f - p);
#pragma unroll
  for (int i = 0; i < 8; ++i) keep[i] = (bits >> i) & 1u ? inv : 0.f;
}

template <typename T>
struct Src2 {
  const T* x1;
  const T* x2;
  int C1, C2;
  // pointer to channel c0 (multiple of 8) of pixel row `row`
  __device__ __forceinline__ const T* at(long long row, int c0) const {
    return c0 < C1 ? x1 + row * C1 + c0 : x2 + row * C2 + (c0 - C1);
  }
};

// per-thread constants of the 8 channels it owns
struct ChanConst {
  float gam[8], bet[8], mu[2], r[2];
  int g[2];
};
__device__ __forceinline__ void load_consts(ChanConst& k, int n, int c0, int G, int cpg, const float* gamma, const float* beta,
                                            const float* mean, const float* rstd) {
  load8(gamma + c0, k.gam);
  load8(beta + c0, k.bet);
  k.g[0] = c0 / cpg;
  k.g[1] = (c0 + 4) / cpg;
  k.mu[0] = mean[n * G + k.g[0]]; k.r[0] = rstd[n * G + k.g[0]];
  k.mu[1] = mean[n * G + k.g[1]]; k.r[1] = rstd[n * G + k.g[1]];
}

// Software-pipeline driver shared by all streaming kernels: `issue(stage)` copies the next row of every input stream
// into `stage` and advances the stream pointers, `body(stage)` consumes the oldest row.  The ring is walked by a loop
// unrolled DEPTH times so that every shared-memory slot address is an immediate and no ring index is kept.
template <int DEPTH, typename Issue, typename Body>
__device__ __forceinline__ void run_pipeline(int n_it, Issue&& issue, Body&& body) {
#pragma unroll
  for (int d = 0; d < DEPTH; ++d) {
    if (d < n_it) issue(d);
    asm volatile("cp.async.commit_group;" ::: "memory");
  }
  int rem = n_it;
  while (rem > 0) {
#pragma unroll
    for (int st = 0; st < DEPTH; ++st) {
      if (rem <= 0) break;
      asm volatile("cp.async.wait_group %0;" ::"n"(DEPTH - 1) : "memory");
      body(st);
      if (rem > DEPTH) issue(st);
      asm volatile("cp.async.commit_group;" ::: "memory");
      --rem;
    }
  }
}

// thread -> (pixel lane, 8-channel vector) mapping and the walk of one thread over its pixels
struct Walk {
  int V, lanes, v, lane, c0, n_it;
  long long row0;                    // first pixel row (n*hw + p) of this thread
  __device__ __forceinline__ Walk(int Ct, int n, int hw, int part, int parts) {
    V = Ct / 8;
    lanes = 256 / V;
    v = threadIdx.x % V;
    lane = threadIdx.x / V;
    c0 = v * 8;
    const int per = (hw + parts - 1) / parts;
    const int p0 = part * per, p1 = min(hw, p0 + per);
    n_it = (lane < lanes && p1 > p0 + lane) ? (p1 - p0 - lane + lanes - 1) / lanes : 0;
    row0 = (long long)n * hw + p0 + lane;
  }
};

// one input stream of a thread: pointer to the next row to copy, advanced by `step` elements per row walked
template <typename T>
struct Stream {
  const T* p;
  int step;
  __device__ __forceinline__ const T* next() { const T* q = p; p += step; return q; }
};
template <typename T>
__device__ __forceinline__ Stream<T> stream_of(const Src2<T>& s, const Walk& w) {     // (possibly concatenated) input
  const bool first = w.c0 < s.C1;
  const int ld = first ? s.C1 : s.C2;
  const T* base = first ? s.x1 + w.c0 : s.x2 + (w.c0 - s.C1);
  return Stream<T>{base + w.row0 * ld, w.lanes * ld};
}
template <typename T>
__device__ __forceinline__ Stream<T> stream_of(const T* t, int Ct, const Walk& w) {    // plain [rows][Ct] tensor
  return Stream<T>{t + w.row0 * Ct + w.c0, w.lanes * Ct};
}

// ---------------------------------------------------------------- stats
// grid (n_img, splits): part[n][split][G][2] = (sum, sum of squares) over the split's pixels
template <typename T>
__global__ void __launch_bounds__(256, 4) gn_stats_kernel(Src2<T> s, int hw, int G, int splits, float* part) {
  extern __shared__ __align__(16) uint8_t gsm[];
  pdl_wait();
  pdl_trigger();
  using P = Pipe<T, 1, GN_DEPTH>;
  const P pipe(gsm);
  const int Ct = s.C1 + s.C2, cpg = Ct / G;
  const int n = blockIdx.x, sp = blockIdx.y;
  const Walk w(Ct, n, hw, sp, splits);
  const int V = w.V, lanes = w.lanes;
  float sum[2] = {0.f, 0.f}, sq[2] = {0.f, 0.f};
  Stream<T> xs = stream_of(s, w);
  run_pipeline<GN_DEPTH>(
      w.n_it, [&](int st) { pipe.issue(st, 0, xs.next()); },
      [&](int st) {
        float a[8];
        pipe.read(st, 0, a);
#pragma unroll
        for (int i = 0; i < 8; ++i) { sum[i >> 2] += a[i]; sq[i >> 2] = fmaf(a[i], a[i], sq[i >> 2]); }
      });
  // quad q = 2*v + half holds channels [4q, 4q+4): whole quads never straddle a group (cpg % 4 == 0)
  __shared__ float s_sum[512], s_sq[512];
  s_sum[2 * threadIdx.x] = sum[0]; s_sum[2 * threadIdx.x + 1] = sum[1];
  s_sq[2 * threadIdx.x] = sq[0]; s_sq[2 * threadIdx.x + 1] = sq[1];
  __syncthreads();
  if (threadIdx.x < G) {
    const int g = threadIdx.x, qpg = cpg / 4, Q = 2 * V;
    double a = 0., b = 0.;
    for (int l = 0; l < lanes; ++l)
      for (int q = 0; q < qpg; ++q) {
        a += (double)s_sum[l * Q + g * qpg + q];
        b += (double)s_sq[l * Q + g * qpg + q];
      }
    float* o = part + (((long long)n * splits + sp) * G + g) * 2;
    o[0] = (float)a;
    o[1] = (float)b;
  }
}

__global__ void gn_finalize_kernel(const float* part, int n_img, int splits, int G, double inv_count, float eps,
                                   float* mean, float* rstd) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_img * G) return;
  int n = i / G, g = i % G;
  double a = 0., b = 0.;
  for (int sp = 0; sp < splits; ++sp) {
    const float* o = part + (((long long)n * splits + sp) * G + g) * 2;
    a += (double)o[0];
    b += (double)o[1];
  }
  double mu = a * inv_count;
  double var = b * inv_count - mu * mu;
  if (var < 0.) var = 0.;
  mean[i] = (float)mu;
  rstd[i] = (float)(1.0 / sqrt(var + (double)eps));
}

// Dropout handling is a template parameter so that the streaming loops carry no run-time mode branches:
//   DROP_NONE  p == 0
//   DROP_FAST  forward: in-kernel Philox, keep bits written;  backward: keep bits read back through the pipeline
//   DROP_SLOW  injected mask (parity tests) or, in the backward passes, the generator re-run (no keep bits kept)
enum { DROP_NONE = 0, DROP_FAST = 1, DROP_SLOW = 2 };

// ---------------------------------------------------------------- apply
// grid (chunks, n_img); pipelined stream: x (an injected dropout mask - parity tests only - is read directly)
template <typename T, bool ACT, int DROP>
__global__ void __launch_bounds__(256, 4) gn_apply_kernel(Src2<T> s, int hw, int G, const float* __restrict__ gamma,
                                                          const float* __restrict__ beta, const float* __restrict__ mean,
                                                          const float* __restrict__ rstd, float p_drop, uint64_t seed,
                                                          const T* mask, uint8_t* keepbits, T* y,
                                                          const float* __restrict__ part, int splits, double inv_count,
                                                          float eps, float* mean_out, float* rstd_out) {
  extern __shared__ __align__(16) uint8_t gsm[];
  pdl_wait();
  pdl_trigger();
  using P = Pipe<T, 1, GN_DEPTH>;
  const P pipe(gsm);
  const int Ct = s.C1 + s.C2, cpg = Ct / G;
  const int n = blockIdx.y;
  const Walk w(Ct, n, hw, blockIdx.x, gridDim.x);
  ChanConst k;
  if (part) {
    // statistics straight from the partial sums of st_gn_stats (same arithmetic as gn_finalize_kernel); the first
    // chunk of every image publishes mean / rstd for the backward passes
    __shared__ float s_mean[64], s_rstd[64];
    if (threadIdx.x < G) {
      double a = 0., b = 0.;
      for (int sp = 0; sp < splits; ++sp) {
        const float* o = part + (((long long)n * splits + sp) * G + threadIdx.x) * 2;
        a += (double)o[0];
        b += (double)o[1];
      }
      const double mu = a * inv_count;
      double var = b * inv_count - mu * mu;
      if (var < 0.) var = 0.;
      s_mean[threadIdx.x] = (float)mu;
      s_rstd[threadIdx.x] = (float)(1.0 / sqrt(var + (double)eps));
      if (blockIdx.x == 0) {
        mean_out[n * G + threadIdx.x] = s_mean[threadIdx.x];
        rstd_out[n * G + threadIdx.x] = s_rstd[threadIdx.x];
      }
    }
    __syncthreads();
    if (w.lane >= w.lanes) return;
    load_consts(k, 0, w.c0, G, cpg, gamma, beta, s_mean, s_rstd);
  } else {
    if (w.lane >= w.lanes) return;
    load_consts(k, n, w.c0, G, cpg, gamma, beta, mean, rstd);
  }
  // y = x*A + B with A = rstd*gamma, B = beta - mean*rstd*gamma
  float A[8], Bc[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    A[i] = k.r[i >> 2] * k.gam[i];
    Bc[i] = fmaf(-k.mu[i >> 2], A[i], k.bet[i]);
  }
  Stream<T> xs = stream_of(s, w);
  long long oct = w.row0 * w.V + w.v;              // index of the 8-vector being produced
  const int octstep = w.lanes * w.V;
  run_pipeline<GN_DEPTH>(
      w.n_it, [&](int st) { pipe.issue(st, 0, xs.next()); },
      [&](int st) {
        float x[8], o[8];
        pipe.read(st, 0, x);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const float u = fmaf(x[i], A[i], Bc[i]);
          o[i] = ACT ? silu_t<T>(u) : u;
        }
        if constexpr (DROP == DROP_FAST) {
          float keep[8];
          const uint32_t bits = dropout8(seed, (uint64_t)oct, p_drop, keep);
          if (keepbits) keepbits[oct] = (uint8_t)bits;
#pragma unroll
          for (int i = 0; i < 8; ++i) o[i] *= keep[i];
        } else if constexpr (DROP == DROP_SLOW) {
          float mk[8];
          load8(mask + oct * 8, mk);
#pragma unroll
          for (int i = 0; i < 8; ++i) o[i] *= mk[i];
        }
        store8(y + oct * 8, o);
        oct += octstep;
      });
}

// per-thread constants of the backward passes: xhat = x*r + nmr, pre-activation u = x*rg + bc
struct BwdConst {
  float rg[8], bc[8], r[2], nmr[2];
  __device__ __forceinline__ explicit BwdConst(const ChanConst& k) {
#pragma unroll
    for (int h = 0; h < 2; ++h) { r[h] = k.r[h]; nmr[h] = -k.mu[h] * k.r[h]; }
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      rg[i] = k.r[i >> 2] * k.gam[i];
      bc[i] = fmaf(-k.mu[i >> 2], rg[i], k.bet[i]);
    }
  }
};

// dz and xhat of one 8-vector (shared by both backward passes)
template <typename T, bool ACT, int DROP>
__device__ __forceinline__ void gn_dz8(const float x[8], const float dyv[8], const BwdConst& k,
                                       float p_drop, float inv_keep, uint64_t seed, const float* mkv, uint32_t bits,
                                       long long oct, float xhat[8], float dz[8]) {
  float mk[8];
  if constexpr (DROP == DROP_SLOW) {
    if (mkv) {
#pragma unroll
      for (int i = 0; i < 8; ++i) mk[i] = mkv[i];
    } else {
      dropout8(seed, (uint64_t)oct, p_drop, mk);
    }
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    xhat[i] = fmaf(x[i], k.r[i >> 2], k.nmr[i >> 2]);
    float d = dyv[i];
    if constexpr (DROP == DROP_FAST) d = (bits >> i) & 1u ? d * inv_keep : 0.f;
    if constexpr (DROP == DROP_SLOW) d *= mk[i];
    if constexpr (ACT) d *= silu_grad_t<T>(fmaf(x[i], k.rg[i], k.bc[i]));
    dz[i] = d;
  }
}

// ---------------------------------------------------------------- backward pass 1
// grid (n_img, splits): red[n][split][c][2] = (sum dz, sum dz*xhat) over the split's pixels
// (body shared by the stand-alone kernel and the fused two-phase kernel; image n, pixel split sp of `splits`)
template <typename T, bool ACT, int DROP>
__device__ __forceinline__ void gn_bwd_reduce_body(const Src2<T>& s, const T* dy, int hw, int G, int splits,
                                                   const float* __restrict__ gamma, const float* __restrict__ beta,
                                                   const float* __restrict__ mean, const float* __restrict__ rstd,
                                                   float p_drop, uint64_t seed, const T* mask,
                                                   const uint8_t* __restrict__ keepbits, float* red, int n, int sp) {
  extern __shared__ __align__(16) uint8_t gsm[];
  const int Ct = s.C1 + s.C2, cpg = Ct / G;
  const Walk w(Ct, n, hw, sp, splits);
  const int V = w.V, lanes = w.lanes, v = w.v, lane = w.lane;
  float a[8], b[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) a[i] = b[i] = 0.f;
  using P = Pipe<T, 2, GN_DEPTH>;            // streams: x, dy (+ the keep-bits side stream)
  const P pipe(gsm);
  ChanConst kc;
  load_consts(kc, n, lane < lanes ? w.c0 : 0, G, cpg, gamma, beta, mean, rstd);
  const BwdConst k(kc);
  const float inv_keep = 1.f / (1.f - p_drop);
  Stream<T> xs = stream_of(s, w), ds = stream_of(dy, Ct, w);
  const int octstep = lanes * V;
  long long oct = w.row0 * V + v;                    // vector being consumed
  Stream<uint8_t> bs{keepbits + oct, octstep};       // its keep bits
  run_pipeline<GN_DEPTH>(
      w.n_it,
      [&](int st) {
        pipe.issue(st, 0, xs.next());
        pipe.issue(st, 1, ds.next());
        if constexpr (DROP == DROP_FAST) pipe.issue_byte(st, bs.next());
      },
      [&](int st) {
        float x0[8], d0[8], mk[8], xh[8], dz[8];
        pipe.read(st, 0, x0);
        pipe.read(st, 1, d0);
        uint32_t bits = 0;
        if constexpr (DROP == DROP_FAST) bits = pipe.read_byte(st, (uint32_t)oct);
        if constexpr (DROP == DROP_SLOW) { if (mask) load8(mask + oct * 8, mk); }
        gn_dz8<T, ACT, DROP>(x0, d0, k, p_drop, inv_keep, seed, mask ? mk : nullptr, bits, oct, xh, dz);
#pragma unroll
        for (int i = 0; i < 8; ++i) { a[i] += dz[i]; b[i] = fmaf(dz[i], xh[i], b[i]); }
        if constexpr (DROP != DROP_NONE) oct += octstep;
      });
  // reduce over pixel lanes: smem [lane][V][16] (reuses the pipeline's shared memory once it has drained)
  __syncthreads();
  float* s_red = reinterpret_cast<float*>(gsm);
  if (lane < lanes) {
    float* o = s_red + ((size_t)lane * V + v) * 16;
#pragma unroll
    for (int i = 0; i < 8; ++i) { o[2 * i] = a[i]; o[2 * i + 1] = b[i]; }
  }
  __syncthreads();
  // thread t sums one (channel, component) column: 16*V columns
  for (int col = threadIdx.x; col < 16 * V; col += 256) {
    float t = 0.f;
    for (int l = 0; l < lanes; ++l) t += s_red[(size_t)l * V * 16 + col];
    red[((long long)n * splits + sp) * Ct * 2 + col] = t;     // col = (v*8 + i)*2 + comp = c*2 + comp
  }
}

template <typename T, bool ACT, int DROP>
__global__ void __launch_bounds__(256, 3) gn_bwd_reduce_kernel(Src2<T> s, const T* dy, int hw, int G, int splits,
                                                            const float* __restrict__ gamma, const float* __restrict__ beta,
                                                            const float* __restrict__ mean, const float* __restrict__ rstd,
                                                            float p_drop, uint64_t seed, const T* mask,
                                                            const uint8_t* __restrict__ keepbits, float* red) {
  pdl_wait();
  pdl_trigger();
  gn_bwd_reduce_body<T, ACT, DROP>(s, dy, hw, G, splits, gamma, beta, mean, rstd, p_drop, seed, mask, keepbits, red,
                                   blockIdx.x, blockIdx.y);
}

// dgamma[c] += sum_rows red[row][c][1]; dbeta[c] += sum_rows red[row][c][0]
// block = 16 channels x 32 row lanes (float2 reads), fixed-order reduction; the input is tiny (rows x C x 8 bytes),
// the kernel is latency-bound, so the rows are spread over as many lanes as a block holds
__global__ void __launch_bounds__(512) gn_bwd_params_kernel(const float* __restrict__ red, int rows, int C, float* dgamma,
                                                            float* dbeta) {
  const int c = blockIdx.x * 16 + threadIdx.x % 16;
  const int rl = threadIdx.x / 16;
  float a = 0.f, b = 0.f;
  if (c < C) {
    for (int r = rl; r < rows; r += 32) {
      float2 v = *reinterpret_cast<const float2*>(red + ((long long)r * C + c) * 2);
      a += v.x;
      b += v.y;
    }
  }
  __shared__ float sa[512], sb[512];
  sa[threadIdx.x] = a;
  sb[threadIdx.x] = b;
  __syncthreads();
  if (rl == 0 && c < C) {
    double ta = 0., tb = 0.;
    for (int l = 0; l < 32; ++l) { ta += (double)sa[l * 16 + threadIdx.x]; tb += (double)sb[l * 16 + threadIdx.x]; }
    dbeta[c] += (float)ta;
    dgamma[c] += (float)tb;
  }
}

// ---------------------------------------------------------------- backward pass 2
// grid (chunks, n_img)
// (body shared with the fused kernel; image n of n_img, pixel chunk `chunk` of `chunks`; `red` is read with plain
// loads - in the fused kernel other CTAs of the cluster wrote it moments ago)
template <typename T, bool ACT, int DROP, bool CSUM>
__device__ __forceinline__ void gn_bwd_apply_body(const Src2<T>& s, const T* dy, int hw, int G, int splits,
                                                  const float* __restrict__ gamma, const float* __restrict__ beta,
                                                  const float* __restrict__ mean, const float* __restrict__ rstd,
                                                  float p_drop, uint64_t seed, const T* mask,
                                                  const uint8_t* __restrict__ keepbits,
                                                  const float* red, const T* extra, float extra_scale,
                                                  T* dx1, int accum1, T* dx2, int accum2, float* csum,
                                                  float* dgamma, float* dbeta, int n, int n_img, int chunk, int chunks) {
  extern __shared__ __align__(16) uint8_t gsm[];
  const int Ct = s.C1 + s.C2, cpg = Ct / G;
  __shared__ float sh1[64], sh2[64];
  if (dgamma && chunk == 0) {
    // parameter gradients (what st_gn_bwd_params computes), spread over the first chunk's blocks: block n reduces
    // channels [16n, 16n+16) over all rows of `red` with 16 row lanes, fixed order
    __shared__ float sa[256], sb[256];
    const int rows = n_img * splits;
    for (int cb = n; cb * 16 < Ct; cb += n_img) {
      const int c = cb * 16 + threadIdx.x % 16, rl = threadIdx.x / 16;
      float a = 0.f, b = 0.f;
      if (c < Ct) {
        for (int r = rl; r < rows; r += 16) {
          const float2 v = *reinterpret_cast<const float2*>(red + ((long long)r * Ct + c) * 2);
          a += v.x;
          b += v.y;
        }
      }
      sa[threadIdx.x] = a;
      sb[threadIdx.x] = b;
      __syncthreads();
      if (rl == 0 && c < Ct) {
        double ta = 0., tb = 0.;
        for (int l = 0; l < 16; ++l) { ta += (double)sa[l * 16 + threadIdx.x]; tb += (double)sb[l * 16 + threadIdx.x]; }
        dbeta[c] += (float)ta;
        dgamma[c] += (float)tb;
      }
      __syncthreads();
    }
  }
  if (threadIdx.x < G) {
    const int g = threadIdx.x;
    double a = 0., b = 0.;
    for (int sp = 0; sp < splits; ++sp)
      for (int c = g * cpg; c < (g + 1) * cpg; ++c) {
        const float* o = red + (((long long)n * splits + sp) * Ct + c) * 2;
        a += (double)gamma[c] * (double)__ldcg(o);
        b += (double)gamma[c] * (double)__ldcg(o + 1);
      }
    const double inv = 1.0 / ((double)hw * cpg);
    sh1[g] = (float)(a * inv);
    sh2[g] = (float)(b * inv);
  }
  __syncthreads();
  Walk w(Ct, n, hw, chunk, chunks);
  const int V = w.V, lanes = w.lanes, v = w.v, lane = w.lane;
  const bool active = lane < lanes;
  if (!active) w.c0 = 0;
  const int c0 = w.c0;
  ChanConst kc;
  load_consts(kc, n, c0, G, cpg, gamma, beta, mean, rstd);
  const BwdConst k(kc);
  // dx = rstd*gamma*dz - rstd*s1 - rstd*s2*xhat
  const float rs1[2] = {-kc.r[0] * sh1[kc.g[0]], -kc.r[1] * sh1[kc.g[1]]}, rs2[2] = {-kc.r[0] * sh2[kc.g[0]], -kc.r[1] * sh2[kc.g[1]]};
  const float inv_keep = 1.f / (1.f - p_drop);
  float cs[8];                               // column sums of this thread's contributions (optional output)
#pragma unroll
  for (int i = 0; i < 8; ++i) cs[i] = 0.f;
  // destination of this thread's channels (first or second tensor of the concatenation)
  const bool first = c0 < s.C1;
  const int dld = first ? s.C1 : s.C2;
  const int acc = first ? accum1 : accum2;
  T* dp = (first ? dx1 + c0 : dx2 + (c0 - s.C1)) + w.row0 * dld;          // destination row being produced
  const int dstep = lanes * dld;
  using P = Pipe<T, 4, GN_BWD_DEPTH>;        // streams: x, dy, extra, old destination (+ the keep-bits side stream)
  const P pipe(gsm);
  Stream<T> xs = stream_of(s, w), ds = stream_of(dy, Ct, w), es = stream_of(extra, Ct, w);
  Stream<T> os{dp, dstep};
  const int octstep = lanes * V;
  long long oct = w.row0 * V + v;
  Stream<uint8_t> bs{keepbits + oct, octstep};
  run_pipeline<GN_BWD_DEPTH>(
      w.n_it,
      [&](int st) {
        pipe.issue(st, 0, xs.next());
        pipe.issue(st, 1, ds.next());
        if (extra) pipe.issue(st, 2, es.next());
        if (acc) pipe.issue(st, 3, os.next());
        if constexpr (DROP == DROP_FAST) pipe.issue_byte(st, bs.next());
      },
      [&](int st) {
        float x0[8], d0[8], mk[8], xh[8], dz[8], o[8];
        pipe.read(st, 0, x0);
        pipe.read(st, 1, d0);
        uint32_t bits = 0;
        if constexpr (DROP == DROP_FAST) bits = pipe.read_byte(st, (uint32_t)oct);
        if constexpr (DROP == DROP_SLOW) { if (mask) load8(mask + oct * 8, mk); }
        gn_dz8<T, ACT, DROP>(x0, d0, k, p_drop, inv_keep, seed, mask ? mk : nullptr, bits, oct, xh, dz);
#pragma unroll
        for (int i = 0; i < 8; ++i) o[i] = fmaf(k.rg[i], dz[i], fmaf(xh[i], rs2[i >> 2], rs1[i >> 2]));
        if (extra) {
          float ex[8];
          pipe.read(st, 2, ex);
#pragma unroll
          for (int i = 0; i < 8; ++i) o[i] = fmaf(extra_scale, ex[i], o[i]);
        }
        if constexpr (CSUM) {
#pragma unroll
          for (int i = 0; i < 8; ++i) cs[i] += o[i];
        }
        if (acc) {
          float old[8];
          pipe.read(st, 3, old);
#pragma unroll
          for (int i = 0; i < 8; ++i) o[i] += old[i];
        }
        store8(dp, o);
        dp += dstep;
        if constexpr (DROP != DROP_NONE) oct += octstep;
      });
  if constexpr (CSUM) {
    // csum[n][chunk][c] = sum over this block's pixels of the gradient it contributed (fixed-order lane reduction
    // through the drained pipeline memory): the caller turns these into bias / time-embedding gradients without
    // another pass over the tensor
    __syncthreads();
    float* s_cs = reinterpret_cast<float*>(gsm);
    if (active) {
#pragma unroll
      for (int i = 0; i < 8; ++i) s_cs[(lane * V + v) * 8 + i] = cs[i];
    }
    __syncthreads();
    for (int col = threadIdx.x; col < Ct; col += 256) {
      float t = 0.f;
      for (int l = 0; l < lanes; ++l) t += s_cs[l * Ct + col];
      csum[((long long)n * chunks + chunk) * Ct + col] = t;
    }
  }
}

template <typename T, bool ACT, int DROP, bool CSUM>
__global__ void __launch_bounds__(256, 3) gn_bwd_apply_kernel(Src2<T> s, const T* dy, int hw, int G, int splits,
                                                           const float* __restrict__ gamma, const float* __restrict__ beta,
                                                           const float* __restrict__ mean, const float* __restrict__ rstd,
                                                           float p_drop, uint64_t seed, const T* mask,
                                                           const uint8_t* __restrict__ keepbits,
                                                           const float* red, const T* extra, float extra_scale,
                                                           T* dx1, int accum1, T* dx2, int accum2, float* csum,
                                                           float* dgamma, float* dbeta) {
  pdl_wait();
  pdl_trigger();
  gn_bwd_apply_body<T, ACT, DROP, CSUM>(s, dy, hw, G, splits, gamma, beta, mean, rstd, p_drop, seed, mask, keepbits, red,
                                        extra, extra_scale, dx1, accum1, dx2, accum2, csum, dgamma, dbeta, blockIdx.y,
                                        gridDim.y, blockIdx.x, gridDim.x);
}

// Both backward passes in one launch: a thread-block cluster of `chunks` CTAs owns one image; every CTA reduces its
// pixel chunk (phase 1, writes red[n][chunk]), the cluster synchronises, and every CTA produces dx for the same chunk
// (phase 2).  Saves a launch and the stream-ordered round trip of `red`; the second read of x and dy hits L2 only
// partly (see fused_chunks_for for the measured policy).  grid (chunks, n_img).
template <typename T, bool ACT, int DROP, bool CSUM>
__global__ void __launch_bounds__(256, 3) gn_bwd_fused_kernel(Src2<T> s, const T* dy, int hw, int G,
                                                           const float* __restrict__ gamma, const float* __restrict__ beta,
                                                           const float* __restrict__ mean, const float* __restrict__ rstd,
                                                           float p_drop, uint64_t seed, const T* mask,
                                                           const uint8_t* __restrict__ keepbits,
                                                           float* red, const T* extra, float extra_scale,
                                                           T* dx1, int accum1, T* dx2, int accum2, float* csum) {
  pdl_wait();
  pdl_trigger();
  const int n = blockIdx.y, chunk = blockIdx.x, chunks = gridDim.x;
  gn_bwd_reduce_body<T, ACT, DROP>(s, dy, hw, G, chunks, gamma, beta, mean, rstd, p_drop, seed, mask, keepbits, red, n, chunk);
  __threadfence();
  __syncthreads();
  if (chunks > 1) {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
  }
  gn_bwd_apply_body<T, ACT, DROP, CSUM>(s, dy, hw, G, chunks, gamma, beta, mean, rstd, p_drop, seed, mask, keepbits, red,
                                        extra, extra_scale, dx1, accum1, dx2, accum2, csum, nullptr, nullptr, n,
                                        (int)gridDim.y, chunk, chunks);
}

// opt a kernel in to more than 48 KB of dynamic shared memory (once per kernel instance)
template <typename K>
bool allow_smem(K kernel, int bytes) {
  if (bytes <= 48 * 1024) return true;
  cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
  if (e != cudaSuccess) { st_set_error("groupnorm: cudaFuncSetAttribute(%d bytes): %s", bytes, cudaGetErrorString(e)); return false; }
  return true;
}

int check_geom(int C1, int C2, int G) {
  int Ct = C1 + C2;
  ST_CHECK_ARG(C1 > 0 && C2 >= 0 && C1 % 8 == 0 && C2 % 8 == 0, "groupnorm: channel counts must be multiples of 8 (got %d,%d)", C1, C2);
  ST_CHECK_ARG(G > 0 && G <= 64 && Ct % G == 0 && (Ct / G) % 4 == 0, "groupnorm: group size must be a multiple of 4 (C=%d,G=%d)", Ct, G);
  ST_CHECK_ARG(Ct <= 2048, "groupnorm: C > 2048 unsupported");
  return 0;
}

// pixel chunks per image so that the grid is a few waves of 148 SMs x 8 resident blocks
int chunks_for(int n_img, int hw, int V) {
  int lanes = 256 / V;
  int max_chunks = (hw + GN_DEPTH * lanes - 1) / (GN_DEPTH * lanes);   // at least one full pipeline per block
  static const int waves = getenv("ST_GN_CHUNK_W") ? atoi(getenv("ST_GN_CHUNK_W")) : 4;   // tuning knob
  int want = (st_num_sms() * waves + n_img - 1) / n_img;
  int c = want < max_chunks ? want : max_chunks;
  if (c > 65535) c = 65535;
  return c < 1 ? 1 : c;
}

// run f(std::bool_constant<act>, std::integral_constant<int, drop>) for the run-time (act, drop) pair
template <typename F>
void dispatch_mode(int act, int drop, F&& f) {
  auto with_act = [&](auto A) {
    if (drop == DROP_NONE) f(A, std::integral_constant<int, DROP_NONE>{});
    else if (drop == DROP_FAST) f(A, std::integral_constant<int, DROP_FAST>{});
    else f(A, std::integral_constant<int, DROP_SLOW>{});
  };
  if (act) with_act(std::true_type{}); else with_act(std::false_type{});
}

}  // namespace

extern "C" __attribute__((visibility("default"))) int st_gn_chunks(int n_img, int hw, int C) {
  return chunks_for(n_img, hw, C / 8);
}

extern "C" __attribute__((visibility("default"))) int st_gn_stats(const void* x1, const void* x2, int dtype, int n_img, int hw, int C1, int C2, int G,
                           int splits, float* part, void* stream) {
  if (int e = check_geom(C1, C2, G)) return e;
  ST_CHECK_ARG(splits >= 1 && splits <= 65535, "st_gn_stats: bad splits");
  ST_DISPATCH_DTYPE(dtype, T, {
    Src2<T> s{(const T*)x1, (const T*)x2, C1, C2};
    constexpr int smem = Pipe<T, 1, GN_DEPTH>::BYTES;
    static bool smem_ok = false;
    if (!smem_ok) { if (!allow_smem(gn_stats_kernel<T>, smem)) return ST_ERR_CUDA; smem_ok = true; }
    st_launch(gn_stats_kernel<T>, dim3(n_img, splits), dim3(256), smem, (cudaStream_t)stream, s, hw, G, splits, part);
  });
  ST_CHECK_LAUNCH("st_gn_stats");
  return 0;
}

extern "C" __attribute__((visibility("default"))) int st_gn_finalize(const float* part, int n_img, int splits, int G, int64_t count, float eps,
                              float* mean, float* rstd, void* stream) {
  int n = n_img * G;
  gn_finalize_kernel<<<(n + 127) / 128, 128, 0, (cudaStream_t)stream>>>(part, n_img, splits, G, 1.0 / (double)count,
                                                                       eps, mean, rstd);
  ST_CHECK_LAUNCH("st_gn_finalize");
  return 0;
}

extern "C" __attribute__((visibility("default"))) int st_gn_apply(const void* x1, const void* x2, int dtype, int n_img, int hw, int C1, int C2, int G,
                           const float* gamma, const float* beta, float* mean, float* rstd, int act,
                           float p_drop, uint64_t seed, const void* mask, uint8_t* keepbits, void* y, const float* part, int splits,
                           int64_t count, float eps, void* stream) {
  if (int e = check_geom(C1, C2, G)) return e;
  ST_CHECK_ARG(n_img <= 65535, "st_gn_apply: more than 65535 images");
  ST_CHECK_ARG(!part || (splits >= 1 && count > 0 && mean && rstd), "st_gn_apply: partial sums need splits, count and mean/rstd outputs");
  const int V = (C1 + C2) / 8;
  const int drop = mask ? DROP_SLOW : (p_drop > 0.f ? DROP_FAST : DROP_NONE);
  int rc = 0;
  ST_DISPATCH_DTYPE(dtype, T, {
    Src2<T> s{(const T*)x1, (const T*)x2, C1, C2};
    constexpr int smem = Pipe<T, 1, GN_DEPTH>::BYTES;
    dispatch_mode(act, drop, [&](auto A, auto D) {
      constexpr bool ACT = decltype(A)::value;
      constexpr int DROP = decltype(D)::value;
      static bool smem_ok = false;
      if (!smem_ok) { if (!allow_smem(gn_apply_kernel<T, ACT, DROP>, smem)) { rc = ST_ERR_CUDA; return; } smem_ok = true; }
      st_launch(gn_apply_kernel<T, ACT, DROP>, dim3(chunks_for(n_img, hw, V), n_img), dim3(256), smem, (cudaStream_t)stream,
          s, hw, G, gamma, beta, mean, rstd, p_drop, seed, (const T*)mask, keepbits, (T*)y, part, splits,
          part ? 1.0 / (double)count : 0.0, eps, mean, rstd);
    });
  });
  if (rc) return rc;
  ST_CHECK_LAUNCH("st_gn_apply");
  return 0;
}

extern "C" __attribute__((visibility("default"))) int st_gn_bwd_reduce(const void* x1, const void* x2, const void* dy, int dtype, int n_img, int hw, int C1,
                                int C2, int G, const float* gamma, const float* beta, const float* mean,
                                const float* rstd, int act, float p_drop, uint64_t seed, const void* mask,
                                const uint8_t* keepbits, int splits, float* red, void* stream) {
  if (int e = check_geom(C1, C2, G)) return e;
  const int drop = mask ? DROP_SLOW : (p_drop > 0.f ? (keepbits ? DROP_FAST : DROP_SLOW) : DROP_NONE);
  int rc = 0;
  ST_DISPATCH_DTYPE(dtype, T, {
    Src2<T> s{(const T*)x1, (const T*)x2, C1, C2};
    constexpr int smem = Pipe<T, 2, GN_DEPTH>::BYTES;      // >= the 16 KB the final lane reduction reuses
    dispatch_mode(act, drop, [&](auto A, auto D) {
      constexpr bool ACT = decltype(A)::value;
      constexpr int DROP = decltype(D)::value;
      static bool smem_ok = false;
      if (!smem_ok) { if (!allow_smem(gn_bwd_reduce_kernel<T, ACT, DROP>, smem)) { rc = ST_ERR_CUDA; return; } smem_ok = true; }
      st_launch(gn_bwd_reduce_kernel<T, ACT, DROP>, dim3(n_img, splits), dim3(256), smem, (cudaStream_t)stream,
          s, (const T*)dy, hw, G, splits, gamma, beta, mean, rstd, p_drop, seed, (const T*)mask, keepbits, red);
    });
  });
  if (rc) return rc;
  ST_CHECK_LAUNCH("st_gn_bwd_reduce");
  return 0;
}

extern "C" __attribute__((visibility("default"))) int st_gn_bwd_params(const float* red, int rows, int C, float* dgamma, float* dbeta, void* stream) {
  gn_bwd_params_kernel<<<(C + 15) / 16, 512, 0, (cudaStream_t)stream>>>(red, rows, C, dgamma, dbeta);
  ST_CHECK_LAUNCH("st_gn_bwd_params");
  return 0;
}

extern "C" __attribute__((visibility("default"))) int st_gn_bwd_apply(const void* x1, const void* x2, const void* dy, int dtype, int n_img, int hw, int C1,
                               int C2, int G, const float* gamma, const float* beta, const float* mean,
                               const float* rstd, int act, float p_drop, uint64_t seed, const void* mask,
                               const uint8_t* keepbits, int splits, const float* red, const void* extra, float extra_scale, void* dx1,
                               int accum1, void* dx2, int accum2, int chunks, float* csum, float* dgamma, float* dbeta, void* stream) {
  if (int e = check_geom(C1, C2, G)) return e;
  ST_CHECK_ARG(n_img <= 65535, "st_gn_bwd_apply: more than 65535 images");
  const int V = (C1 + C2) / 8;
  ST_CHECK_ARG(!csum || chunks > 0, "st_gn_bwd_apply: csum needs an explicit chunk count");
  if (chunks <= 0) chunks = chunks_for(n_img, hw, V);
  ST_CHECK_ARG(chunks <= 65535, "st_gn_bwd_apply: too many chunks");
  ST_CHECK_ARG((dgamma == nullptr) == (dbeta == nullptr), "st_gn_bwd_apply: dgamma and dbeta go together");
  const int drop = mask ? DROP_SLOW : (p_drop > 0.f ? (keepbits ? DROP_FAST : DROP_SLOW) : DROP_NONE);
  int rc = 0;
  ST_DISPATCH_DTYPE(dtype, T, {
    Src2<T> s{(const T*)x1, (const T*)x2, C1, C2};
    constexpr int smem = Pipe<T, 4, GN_BWD_DEPTH>::BYTES;
    dispatch_mode(act, drop, [&](auto A, auto D) {
      constexpr bool ACT = decltype(A)::value;
      constexpr int DROP = decltype(D)::value;
      auto launch = [&](auto CS) {
        constexpr bool CSUM = decltype(CS)::value;
        static bool smem_ok = false;
        if (!smem_ok) { if (!allow_smem(gn_bwd_apply_kernel<T, ACT, DROP, CSUM>, smem)) { rc = ST_ERR_CUDA; return; } smem_ok = true; }
        st_launch(gn_bwd_apply_kernel<T, ACT, DROP, CSUM>, dim3(chunks, n_img), dim3(256), smem, (cudaStream_t)stream,
                  s, (const T*)dy, hw, G, splits, gamma, beta, mean, rstd, p_drop, seed, (const T*)mask, keepbits, red,
                  (const T*)extra, extra_scale, (T*)dx1, accum1, (T*)dx2, accum2, csum, dgamma, dbeta);
      };
      if (csum) launch(std::true_type{}); else launch(std::false_type{});
    });
  });
  if (rc) return rc;
  ST_CHECK_LAUNCH("st_gn_bwd_apply");
  return 0;
}

// ---------------------------------------------------------------- fused backward (one launch, cluster per image)
namespace {
// Pixel chunks per image (= cluster size) of the fused backward kernel, or 0 when the two-kernel path should run.
// Measured on B200 at B=512 (tools/gn_bench.py, profiles/r01_gn_fused.txt): the saving is the second launch and the
// reduction buffer round trip, not HBM traffic - a resident wave of CTAs touches more x/dy than L2 retains, and finer
// chunks (clusters of 4-16) are bound by the per-CTA latency chain (constants, pipeline fill, cluster barrier).
// One CTA per image wins up to 8x8, a pair at 16x16, and at >= 32x32 a pair only for the 2-stream form (x, dy); with
// `extra` / accumulate streams the two-kernel form stays ahead there.
int fused_chunks_for(int n_img, int hw, int Ct, int elem_bytes, int streams) {
  static const int mode = getenv("ST_GN_FUSED") ? atoi(getenv("ST_GN_FUSED")) : 1;
  static const int force = getenv("ST_GN_FUSED_CHUNKS") ? atoi(getenv("ST_GN_FUSED_CHUNKS")) : 0;
  (void)Ct; (void)elem_bytes;
  if (!mode) return 0;
  int c = hw <= 64 ? 1 : 2;
  if (hw >= 1024 && streams > 2) c = 0;
  if (force > 0) c = force > 16 ? 16 : force;
  if (c > hw) c = hw;
  if ((long long)n_img * c < st_num_sms()) return 0;                  // too few CTAs: the two-kernel path splits finer
  return c;
}
}  // namespace

extern "C" __attribute__((visibility("default"))) int st_gn_bwd_fused_chunks(int n_img, int hw, int C, int dtype, int streams) {
  return fused_chunks_for(n_img, hw, C, dtype == ST_BF16 ? 2 : 4, streams);
}

extern "C" __attribute__((visibility("default"))) int st_gn_bwd_fused(const void* x1, const void* x2, const void* dy, int dtype, int n_img, int hw, int C1,
                               int C2, int G, const float* gamma, const float* beta, const float* mean,
                               const float* rstd, int act, float p_drop, uint64_t seed, const void* mask,
                               const uint8_t* keepbits, int chunks, float* red, const void* extra, float extra_scale,
                               void* dx1, int accum1, void* dx2, int accum2, float* csum, void* stream) {
  if (int e = check_geom(C1, C2, G)) return e;
  ST_CHECK_ARG(n_img <= 65535, "st_gn_bwd_fused: more than 65535 images");
  ST_CHECK_ARG(chunks >= 1 && chunks <= 16, "st_gn_bwd_fused: chunks (cluster size) must be 1..16, got %d", chunks);
  const int drop = mask ? DROP_SLOW : (p_drop > 0.f ? (keepbits ? DROP_FAST : DROP_SLOW) : DROP_NONE);
  int rc = 0;
  ST_DISPATCH_DTYPE(dtype, T, {
    Src2<T> s{(const T*)x1, (const T*)x2, C1, C2};
    constexpr int smem_a = Pipe<T, 2, GN_DEPTH>::BYTES, smem_b = Pipe<T, 4, GN_BWD_DEPTH>::BYTES;
    constexpr int smem = smem_a > smem_b ? smem_a : smem_b;
    dispatch_mode(act, drop, [&](auto A, auto D) {
      constexpr bool ACT = decltype(A)::value;
      constexpr int DROP = decltype(D)::value;
      auto launch = [&](auto CS) {
        constexpr bool CSUM = decltype(CS)::value;
        auto kernel = gn_bwd_fused_kernel<T, ACT, DROP, CSUM>;
        static bool attr_ok = false;
        if (!attr_ok) {
          if (!allow_smem(kernel, smem)) { rc = ST_ERR_CUDA; return; }
          cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
          if (e != cudaSuccess) { st_set_error("st_gn_bwd_fused: cluster attribute: %s", cudaGetErrorString(e)); rc = ST_ERR_CUDA; return; }
          attr_ok = true;
        }
        cudaLaunchConfig_t cfg;
        memset(&cfg, 0, sizeof(cfg));
        cfg.gridDim = dim3(chunks, n_img);
        cfg.blockDim = dim3(256);
        cfg.dynamicSmemBytes = smem;
        cfg.stream = (cudaStream_t)stream;
        cudaLaunchAttribute attr[2];
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = chunks;
        attr[0].val.clusterDim.y = 1;
        attr[0].val.clusterDim.z = 1;
        attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        attr[1].val.programmaticStreamSerializationAllowed = 1;
        cfg.attrs = attr;
        cfg.numAttrs = st_pdl_on((cudaStream_t)stream) ? 2 : 1;
        cudaError_t e = cudaLaunchKernelEx(&cfg, kernel, s, (const T*)dy, hw, G, gamma, beta, mean, rstd, p_drop, seed,
                                           (const T*)mask, keepbits, red, (const T*)extra, extra_scale, (T*)dx1, accum1,
                                           (T*)dx2, accum2, csum);
        if (e != cudaSuccess) { st_set_error("st_gn_bwd_fused: launch: %s", cudaGetErrorString(e)); rc = ST_ERR_CUDA; }
      };
      if (csum) launch(std::true_type{}); else launch(std::false_type{});
    });
  });
  if (rc) return rc;
  ST_CHECK_LAUNCH("st_gn_bwd_fused");
  return 0;
}

// ---------------------------------------------------------------- fused forward (statistics + apply, one launch)
// A thread-block cluster of `chunks` CTAs owns one image and keeps it RESIDENT in shared memory: every thread copies
// all of its pixels (<= GN_RES 8-channel vectors) with cp.async up front - the whole chunk is in flight at once -,
// sums them, the CTAs exchange their per-group partial sums through distributed shared memory, and the normalised /
// activated output is produced from the resident copy.  HBM sees the tensor once in and once out (2 passes instead of
// the 3 of st_gn_stats + st_gn_apply) and one launch instead of two.
namespace {
constexpr int GN_RES = 16;        // resident 8-channel vectors per thread: 16 x 16 B x 256 threads = 64 KB (bf16)

template <int I, int N, typename F>
__device__ __forceinline__ void static_for(F&& f) {
  if constexpr (I < N) {
    f(std::integral_constant<int, I>{});
    static_for<I + 1, N>(f);
  }
}
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ float ld_dsmem_f32(const float* local_smem_ptr, uint32_t cta) {
  uint32_t a = (uint32_t)__cvta_generic_to_shared(local_smem_ptr), ra;
  float v;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(ra) : "r"(a), "r"(cta));
  asm volatile("ld.shared::cluster.f32 %0, [%1];" : "=f"(v) : "r"(ra) : "memory");
  return v;
}

template <typename T, bool ACT, int DROP>
__global__ void __launch_bounds__(256, 3) gn_fwd_fused_kernel(Src2<T> s, int hw, int G, const float* __restrict__ gamma,
                                                              const float* __restrict__ beta, double inv_count, float eps,
                                                              float p_drop, uint64_t seed, const T* mask, uint8_t* keepbits,
                                                              T* y, float* mean_out, float* rstd_out) {
  extern __shared__ __align__(16) uint8_t gsm[];
  __shared__ float s_sum[512], s_sq[512];
  __shared__ float s_part[128];               // this CTA's per-group (sum, sum of squares): read by the whole cluster
  __shared__ float s_mean[64], s_rstd[64];
  pdl_wait();
  pdl_trigger();
  using P = Pipe<T, 1, GN_RES>;
  const P pipe(gsm);
  const int Ct = s.C1 + s.C2, cpg = Ct / G;
  const int n = blockIdx.y, chunk = blockIdx.x, chunks = gridDim.x;
  const Walk w(Ct, n, hw, chunk, chunks);
  const int V = w.V, lanes = w.lanes;
  const bool active = w.lane < lanes;
  // ---- the whole chunk in flight
  {
    Stream<T> xs = stream_of(s, w);
#pragma unroll
    for (int st = 0; st < GN_RES; ++st) {          // one commit group per vector: the sums below start on arrival
      if (st < w.n_it) pipe.issue(st, 0, xs.next());
      asm volatile("cp.async.commit_group;" ::: "memory");
    }
  }
  float gam[8], bet[8];
  load8(gamma + (active ? w.c0 : 0), gam);
  load8(beta + (active ? w.c0 : 0), bet);
  // dropout keep flags of every resident vector, drawn while the copies are in flight (the generator is ~60 ALU
  // instructions per vector: in the store phase it would not overlap anything) and carried in registers
  uint32_t kb[GN_RES / 4];
#pragma unroll
  for (int i = 0; i < GN_RES / 4; ++i) kb[i] = 0u;
  if constexpr (DROP == DROP_FAST) {
    const long long oct0 = w.row0 * V + w.v;
    const int octstep = lanes * V;
#pragma unroll
    for (int st = 0; st < GN_RES; ++st)
      if (st < w.n_it) {
        float keep[8];
        const long long oct = oct0 + (long long)st * octstep;
        const uint32_t bits = dropout8(seed, (uint64_t)oct, p_drop, keep);
        if (keepbits) keepbits[oct] = (uint8_t)bits;
        kb[st >> 2] |= bits << (8 * (st & 3));
      }
  }
  // ---- statistics of the resident chunk
  float sum[2] = {0.f, 0.f}, sq[2] = {0.f, 0.f};
  static_for<0, GN_RES>([&](auto I) {
    constexpr int st = decltype(I)::value;
    if (st < w.n_it) {
      asm volatile("cp.async.wait_group %0;" ::"n"(GN_RES - 1 - st) : "memory");
      float a[8];
      pipe.read(st, 0, a);
#pragma unroll
      for (int i = 0; i < 8; ++i) { sum[i >> 2] += a[i]; sq[i >> 2] = fmaf(a[i], a[i], sq[i >> 2]); }
    }
  });
  s_sum[2 * threadIdx.x] = sum[0]; s_sum[2 * threadIdx.x + 1] = sum[1];
  s_sq[2 * threadIdx.x] = sq[0]; s_sq[2 * threadIdx.x + 1] = sq[1];
  __syncthreads();
  if (threadIdx.x < G) {
    const int g = threadIdx.x, qpg = cpg / 4, Q = 2 * V;      // quads never straddle a group (cpg % 4 == 0)
    double a = 0., b = 0.;
    for (int l = 0; l < lanes; ++l)
      for (int q = 0; q < qpg; ++q) {
        a += (double)s_sum[l * Q + g * qpg + q];
        b += (double)s_sq[l * Q + g * qpg + q];
      }
    s_part[2 * g] = (float)a;
    s_part[2 * g + 1] = (float)b;
  }
  // ---- exchange across the cluster (fixed rank order: every CTA derives bit-identical statistics)
  if (chunks > 1) {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
  } else {
    __syncthreads();
  }
  if (threadIdx.x < G) {
    const int g = threadIdx.x;
    double a = 0., b = 0.;
    if (chunks > 1) {
      for (int r = 0; r < chunks; ++r) {
        a += (double)ld_dsmem_f32(&s_part[2 * g], (uint32_t)r);
        b += (double)ld_dsmem_f32(&s_part[2 * g + 1], (uint32_t)r);
      }
    } else {
      a = (double)s_part[2 * g];
      b = (double)s_part[2 * g + 1];
    }
    const double mu = a * inv_count;
    double var = b * inv_count - mu * mu;
    if (var < 0.) var = 0.;
    s_mean[g] = (float)mu;
    s_rstd[g] = (float)(1.0 / sqrt(var + (double)eps));
    if (chunk == 0) {
      mean_out[n * G + g] = s_mean[g];
      rstd_out[n * G + g] = s_rstd[g];
    }
  }
  if (chunks > 1) asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");   // done reading peers' s_part
  __syncthreads();
  // ---- apply from the resident copy: y = x*A + B with A = rstd*gamma, B = beta - mean*rstd*gamma
  if (active) {
    const int c0 = w.c0;
    float A[8], Bc[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int g = (c0 + (i & 4)) / cpg;
      A[i] = s_rstd[g] * gam[i];
      Bc[i] = fmaf(-s_mean[g], A[i], bet[i]);
    }
    long long oct = w.row0 * V + w.v;              // index of the 8-vector being produced
    const int octstep = lanes * V;
#pragma unroll
    for (int st = 0; st < GN_RES; ++st)
      if (st < w.n_it) {
        float x[8], o[8];
        pipe.read(st, 0, x);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const float u = fmaf(x[i], A[i], Bc[i]);
          o[i] = ACT ? silu_t<T>(u) : u;
        }
        if constexpr (DROP == DROP_FAST) {
          float keep[8];
          keep_from_bits((kb[st >> 2] >> (8 * (st & 3))) & 0xFFu, p_drop, keep);
#pragma unroll
          for (int i = 0; i < 8; ++i) o[i] *= keep[i];
        } else if constexpr (DROP == DROP_SLOW) {
          float mk[8];
          load8(mask + oct * 8, mk);
#pragma unroll
          for (int i = 0; i < 8; ++i) o[i] *= mk[i];
        }
        store8(y + oct * 8, o);
        oct += octstep;
      }
  }
  // no CTA may exit while a peer can still read its s_part
  if (chunks > 1) asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}

// cluster size for the resident forward kernel, 0 = use st_gn_stats + st_gn_apply
int fwd_fused_chunks_for(int n_img, int hw, int Ct) {
  static const int mode = getenv("ST_GN_FWD_FUSED") ? atoi(getenv("ST_GN_FWD_FUSED")) : 1;
  static const int max_cluster = getenv("ST_GN_FWD_CLUSTER") ? atoi(getenv("ST_GN_FWD_CLUSTER")) : 16;
  if (!mode) return 0;
  const int V = Ct / 8, lanes = 256 / V;
  if (lanes < 1) return 0;
  const int per_cta = lanes * GN_RES;                          // pixels one CTA can hold
  const int c = (hw + per_cta - 1) / per_cta;
  if (c > max_cluster || c > 16) return 0;
  // every chunk must fit: chunk size = ceil(hw / c) pixels
  if ((hw + c - 1) / c > per_cta) return 0;
  if (mode == 1 && (long long)n_img * c < st_num_sms()) return 0;   // too few CTAs to fill the machine
  return c;
}
}  // namespace

extern "C" __attribute__((visibility("default"))) int st_gn_fwd_fused_chunks(int n_img, int hw, int C) {
  return fwd_fused_chunks_for(n_img, hw, C);
}

extern "C" __attribute__((visibility("default"))) int st_gn_fwd_fused(const void* x1, const void* x2, int dtype, int n_img, int hw, int C1, int C2, int G,
                               const float* gamma, const float* beta, float eps, int act, float p_drop, uint64_t seed,
                               const void* mask, uint8_t* keepbits, void* y, float* mean, float* rstd, int chunks,
                               void* stream) {
  if (int e = check_geom(C1, C2, G)) return e;
  const int Ct = C1 + C2, V = Ct / 8, lanes = 256 / V;
  ST_CHECK_ARG(n_img <= 65535, "st_gn_fwd_fused: more than 65535 images");
  ST_CHECK_ARG(chunks >= 1 && chunks <= 16, "st_gn_fwd_fused: chunks (cluster size) must be 1..16, got %d", chunks);
  ST_CHECK_ARG(lanes >= 1 && (hw + chunks - 1) / chunks <= lanes * GN_RES,
               "st_gn_fwd_fused: a chunk of %d pixels x %d channels does not fit the resident buffer", (hw + chunks - 1) / chunks, Ct);
  ST_CHECK_ARG(mean && rstd, "st_gn_fwd_fused: mean / rstd outputs are required");
  const int drop = mask ? DROP_SLOW : (p_drop > 0.f ? DROP_FAST : DROP_NONE);
  const double inv_count = 1.0 / ((double)hw * (Ct / G));
  int rc = 0;
  ST_DISPATCH_DTYPE(dtype, T, {
    Src2<T> s{(const T*)x1, (const T*)x2, C1, C2};
    constexpr int smem = Pipe<T, 1, GN_RES>::VEC_BYTES;
    dispatch_mode(act, drop, [&](auto A, auto D) {
      constexpr bool ACT = decltype(A)::value;
      constexpr int DROP = decltype(D)::value;
      auto kernel = gn_fwd_fused_kernel<T, ACT, DROP>;
      static bool attr_ok = false;
      if (!attr_ok) {
        if (!allow_smem(kernel, smem)) { rc = ST_ERR_CUDA; return; }
        cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
        if (e != cudaSuccess) { st_set_error("st_gn_fwd_fused: cluster attribute: %s", cudaGetErrorString(e)); rc = ST_ERR_CUDA; return; }
        attr_ok = true;
      }
      cudaLaunchConfig_t cfg;
      memset(&cfg, 0, sizeof(cfg));
      cfg.gridDim = dim3(chunks, n_img);
      cfg.blockDim = dim3(256);
      cfg.dynamicSmemBytes = smem;
      cfg.stream = (cudaStream_t)stream;
      cudaLaunchAttribute attr[2];
      attr[0].id = cudaLaunchAttributeClusterDimension;
      attr[0].val.clusterDim.x = chunks;
      attr[0].val.clusterDim.y = 1;
      attr[0].val.clusterDim.z = 1;
      attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
      attr[1].val.programmaticStreamSerializationAllowed = 1;
      cfg.attrs = attr;
      cfg.numAttrs = st_pdl_on((cudaStream_t)stream) ? 2 : 1;
      cudaError_t e = cudaLaunchKernelEx(&cfg, kernel, s, hw, G, gamma, beta, inv_count, eps, p_drop, seed, (const T*)mask,
                                         keepbits, (T*)y, mean, rstd);
      if (e != cudaSuccess) { st_set_error("st_gn_fwd_fused: launch: %s", cudaGetErrorString(e)); rc = ST_ERR_CUDA; }
    });
  });
  if (rc) return rc;
  ST_CHECK_LAUNCH("st_gn_fwd_fused");
  return 0;
}

// ---------------------------------------------------------------- resident backward (x and dy read ONCE)
// The backward without an accumulated destination (GroupNorm_1 of every res-block; GroupNorm_0 / attention norms with
// their `extra` shortcut gradient as a third resident stream where the image is small enough) with the cluster's
// image resident in shared memory: every thread copies its <= R pixels of x and dy (and extra) up front, phase 1 forms dz
// (written back over dy's slot) and the per-channel sums, the cluster exchanges the gamma-weighted per-group sums through
// distributed shared memory, phase 2 produces dx from the resident x and dz.  HBM traffic: x + dy in, dx out - 3 tensor
// passes instead of the 5 of st_gn_bwd_reduce + st_gn_bwd_apply.  The dropout keep bits ride in registers (8 byte loads
// issued with the copies).  `red` [n_img][chunks][C][2] is still written: the parameter gradients are its column sums.
namespace {
// resident pixels per thread: 8 x (x, dy) or 5 x (x, dy, extra) 16-byte vectors x 256 threads = 64 / 60 KB (bf16)
constexpr int bres_for(int streams) { return streams == 2 ? 8 : 5; }

template <typename T>
__device__ __forceinline__ void slot_write(uint32_t addr, const float v[8]) {
  if constexpr (sizeof(T) == 2) {
    uint4 t;
    __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&t);
#pragma unroll
    for (int i = 0; i < 4; ++i) h[i] = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]);
    asm volatile("st.shared.v4.u32 [%0], {%1,%2,%3,%4};" ::"r"(addr), "r"(t.x), "r"(t.y), "r"(t.z), "r"(t.w) : "memory");
  }
}

// sum the 8 per-thread values `val` over the pixel lanes; column (v*8+i) of the result lands in thread `col % 256`'s
// out[col / 256] (columns = channels, <= 512).  s_red: 256*8 floats.
__device__ __forceinline__ void lane_reduce8(const float val[8], bool active, int lane, int lanes, int V, int v, float* s_red,
                                             float out[2]) {
  __syncthreads();
  if (active) {
#pragma unroll
    for (int i = 0; i < 8; ++i) s_red[(lane * V + v) * 8 + i] = val[i];
  }
  __syncthreads();
#pragma unroll
  for (int j = 0; j < 2; ++j) {
    const int col = threadIdx.x + 256 * j;
    float t = 0.f;
    if (col < 8 * V)
      for (int l = 0; l < lanes; ++l) t += s_red[l * 8 * V + col];
    out[j] = t;
  }
}

template <typename T, bool ACT, int DROP, bool CSUM, int NS>
__global__ void __launch_bounds__(256, 3) gn_bwd_resident_kernel(Src2<T> s, const T* dy, const T* extra, float extra_scale, int hw, int G,
                                                                 const float* __restrict__ gamma, const float* __restrict__ beta,
                                                                 const float* __restrict__ mean, const float* __restrict__ rstd,
                                                                 float p_drop, uint64_t seed, const T* mask,
                                                                 const uint8_t* __restrict__ keepbits, float* red, T* dx1,
                                                                 T* dx2, float* csum) {
  extern __shared__ __align__(16) uint8_t gsm[];
  __shared__ float s_red[256 * 8];
  __shared__ float s_gpart[128];             // gamma-weighted per-group (sum dz, sum dz*xhat) of this CTA's chunk
  __shared__ float sh1[64], sh2[64];
  pdl_wait();
  pdl_trigger();
  constexpr int GN_BRES = bres_for(NS);
  using P = Pipe<T, NS, GN_BRES>;
  const P pipe(gsm);
  const int Ct = s.C1 + s.C2, cpg = Ct / G;
  const int n = blockIdx.y, chunk = blockIdx.x, chunks = gridDim.x;
  Walk w(Ct, n, hw, chunk, chunks);
  const int V = w.V, lanes = w.lanes, v = w.v, lane = w.lane;
  const bool active = lane < lanes;
  if (!active) w.c0 = 0;
  const int c0 = w.c0;
  const int octstep = lanes * V;
  const long long oct0 = w.row0 * V + v;
  // ---- everything in flight: x, dy (cp.async, one group per pixel) and the keep bits (plain byte loads)
  {
    Stream<T> xs = stream_of(s, w), ds = stream_of(dy, Ct, w), es = stream_of(extra, Ct, w);
#pragma unroll
    for (int st = 0; st < GN_BRES; ++st) {
      if (st < w.n_it) {
        pipe.issue(st, 0, xs.next());
        pipe.issue(st, 1, ds.next());
        if constexpr (NS == 3) pipe.issue(st, 2, es.next());
      }
      asm volatile("cp.async.commit_group;" ::: "memory");
    }
  }
  uint32_t kb[2] = {0u, 0u};
  if constexpr (DROP == DROP_FAST) {
#pragma unroll
    for (int st = 0; st < GN_BRES; ++st)
      if (st < w.n_it) kb[st >> 2] |= (uint32_t)__ldg(keepbits + oct0 + (long long)st * octstep) << (8 * (st & 3));
  }
  ChanConst kc;
  load_consts(kc, n, c0, G, cpg, gamma, beta, mean, rstd);
  const BwdConst k(kc);
  const float inv_keep = 1.f / (1.f - p_drop);
  // ---- phase 1: dz (kept in dy's slot) and the per-channel sums of this thread's pixels
  float a[8], b[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) a[i] = b[i] = 0.f;
  static_for<0, GN_BRES>([&](auto I) {
    constexpr int st = decltype(I)::value;
    if (st < w.n_it) {
      asm volatile("cp.async.wait_group %0;" ::"n"(GN_BRES - 1 - st) : "memory");
      float x0[8], d0[8], mk[8], xh[8], dz[8];
      pipe.read(st, 0, x0);
      pipe.read(st, 1, d0);
      const long long oct = oct0 + (long long)st * octstep;
      const uint32_t bits = (kb[st >> 2] >> (8 * (st & 3))) & 0xFFu;
      if constexpr (DROP == DROP_SLOW) { if (mask) load8(mask + oct * 8, mk); }
      gn_dz8<T, ACT, DROP>(x0, d0, k, p_drop, inv_keep, seed, mask ? mk : nullptr, bits, oct, xh, dz);
#pragma unroll
      for (int i = 0; i < 8; ++i) { a[i] += dz[i]; b[i] = fmaf(dz[i], xh[i], b[i]); }
      if constexpr (sizeof(T) == 2) {
        slot_write<T>(pipe.slot(st, 1, 0), dz);
      } else {
        asm volatile("st.shared.v4.f32 [%0], {%1,%2,%3,%4};" ::"r"(pipe.slot(st, 1, 0)), "f"(dz[0]), "f"(dz[1]), "f"(dz[2]), "f"(dz[3]) : "memory");
        asm volatile("st.shared.v4.f32 [%0], {%1,%2,%3,%4};" ::"r"(pipe.slot(st, 1, 1)), "f"(dz[4]), "f"(dz[5]), "f"(dz[6]), "f"(dz[7]) : "memory");
      }
    }
  });
  // ---- per-channel sums of the chunk -> red (parameter gradients), gamma-weighted per-group sums -> s_gpart
  float ta[2], tb[2];
  lane_reduce8(a, active, lane, lanes, V, v, s_red, ta);
  lane_reduce8(b, active, lane, lanes, V, v, s_red, tb);
  __syncthreads();
#pragma unroll
  for (int j = 0; j < 2; ++j) {
    const int c = threadIdx.x + 256 * j;
    if (c < Ct) {
      float* o = red + (((long long)n * chunks + chunk) * Ct + c) * 2;
      o[0] = ta[j];
      o[1] = tb[j];
      const float gm = gamma[c];
      s_red[c] = gm * ta[j];
      s_red[512 + c] = gm * tb[j];
    }
  }
  __syncthreads();
  if (threadIdx.x < G) {
    const int g = threadIdx.x;
    double A = 0., Bq = 0.;
    for (int c = g * cpg; c < (g + 1) * cpg; ++c) { A += (double)s_red[c]; Bq += (double)s_red[512 + c]; }
    s_gpart[2 * g] = (float)A;
    s_gpart[2 * g + 1] = (float)Bq;
  }
  if (chunks > 1) {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
  } else {
    __syncthreads();
  }
  if (threadIdx.x < G) {
    const int g = threadIdx.x;
    double A = 0., Bq = 0.;
    if (chunks > 1) {
      for (int r = 0; r < chunks; ++r) {
        A += (double)ld_dsmem_f32(&s_gpart[2 * g], (uint32_t)r);
        Bq += (double)ld_dsmem_f32(&s_gpart[2 * g + 1], (uint32_t)r);
      }
    } else {
      A = (double)s_gpart[2 * g];
      Bq = (double)s_gpart[2 * g + 1];
    }
    const double inv = 1.0 / ((double)hw * cpg);
    sh1[g] = (float)(A * inv);
    sh2[g] = (float)(Bq * inv);
  }
  if (chunks > 1) asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");   // done reading peers' s_gpart
  __syncthreads();
  // ---- phase 2: dx = rstd*gamma*dz - rstd*s1 - rstd*s2*xhat from the resident x and dz
  const float rs1[2] = {-kc.r[0] * sh1[kc.g[0]], -kc.r[1] * sh1[kc.g[1]]}, rs2[2] = {-kc.r[0] * sh2[kc.g[0]], -kc.r[1] * sh2[kc.g[1]]};
  float cs[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) cs[i] = 0.f;
  const bool first = c0 < s.C1;
  const int dld = first ? s.C1 : s.C2;
  T* dp = (first ? dx1 + c0 : dx2 + (c0 - s.C1)) + w.row0 * dld;
  const int dstep = lanes * dld;
#pragma unroll
  for (int st = 0; st < GN_BRES; ++st)
    if (st < w.n_it) {
      float x0[8], dz[8], o[8];
      pipe.read(st, 0, x0);
      pipe.read(st, 1, dz);
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const float xh = fmaf(x0[i], k.r[i >> 2], k.nmr[i >> 2]);
        o[i] = fmaf(k.rg[i], dz[i], fmaf(xh, rs2[i >> 2], rs1[i >> 2]));
      }
      if constexpr (NS == 3) {
        float ex[8];
        pipe.read(st, 2, ex);
#pragma unroll
        for (int i = 0; i < 8; ++i) o[i] = fmaf(extra_scale, ex[i], o[i]);
      }
      if constexpr (CSUM) {
#pragma unroll
        for (int i = 0; i < 8; ++i) cs[i] += o[i];
      }
      store8(dp, o);
      dp += dstep;
    }
  if constexpr (CSUM) {
    float tc[2];
    lane_reduce8(cs, active, lane, lanes, V, v, s_red, tc);
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      const int c = threadIdx.x + 256 * j;
      if (c < Ct) csum[((long long)n * chunks + chunk) * Ct + c] = tc[j];
    }
  }
  if (chunks > 1) asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}

// streams: 2 (x, dy) or 3 (+ extra); an accumulated destination is a further stream -> not applicable.
// Measured (tools/gn_bench.py, B=512): ahead of the two-kernel and fused-pair forms up to clusters of 8; a cluster of
// 16 (32x32x256) is slower (302 vs 257 us).
int bwd_resident_chunks_for(int n_img, int hw, int Ct, int streams) {
  static const int mode = getenv("ST_GN_BWD_RESIDENT") ? atoi(getenv("ST_GN_BWD_RESIDENT")) : 1;
  static const int max_cluster = getenv("ST_GN_BWD_CLUSTER") ? atoi(getenv("ST_GN_BWD_CLUSTER")) : 8;
  if (!mode || (streams != 2 && streams != 3) || Ct > 512) return 0;
  const int V = Ct / 8, lanes = 256 / V;
  if (lanes < 1) return 0;
  const int per_cta = lanes * bres_for(streams);
  const int c = (hw + per_cta - 1) / per_cta;
  if (c > 16 || c > max_cluster || (hw + c - 1) / c > per_cta) return 0;
  if (streams == 3 && c > 2) return 0;      // 5 pixels per thread: ahead only at 8x8x256 and 4x4 (89.6 vs 71.7 us at 16x16x256)
  if (mode == 1 && (long long)n_img * c < st_num_sms()) return 0;
  return c;
}
}  // namespace

extern "C" __attribute__((visibility("default"))) int st_gn_bwd_resident_chunks(int n_img, int hw, int C, int streams) {
  return bwd_resident_chunks_for(n_img, hw, C, streams);
}

extern "C" __attribute__((visibility("default"))) int st_gn_bwd_resident(const void* x1, const void* x2, const void* dy, int dtype, int n_img, int hw,
                               int C1, int C2, int G, const float* gamma, const float* beta, const float* mean,
                               const float* rstd, int act, float p_drop, uint64_t seed, const void* mask,
                               const uint8_t* keepbits, int chunks, float* red, const void* extra, float extra_scale,
                               void* dx1, void* dx2, float* csum, void* stream) {
  if (int e = check_geom(C1, C2, G)) return e;
  const int Ct = C1 + C2, V = Ct / 8, lanes = 256 / V;
  ST_CHECK_ARG(n_img <= 65535, "st_gn_bwd_resident: more than 65535 images");
  ST_CHECK_ARG(Ct <= 512, "st_gn_bwd_resident: C > 512 unsupported");
  ST_CHECK_ARG(chunks >= 1 && chunks <= 16, "st_gn_bwd_resident: chunks (cluster size) must be 1..16, got %d", chunks);
  const int ns = extra ? 3 : 2;
  ST_CHECK_ARG(lanes >= 1 && (hw + chunks - 1) / chunks <= lanes * bres_for(ns),
               "st_gn_bwd_resident: a chunk of %d pixels x %d channels does not fit the resident buffer", (hw + chunks - 1) / chunks, Ct);
  const int drop = mask ? DROP_SLOW : (p_drop > 0.f ? (keepbits ? DROP_FAST : DROP_SLOW) : DROP_NONE);
  int rc = 0;
  ST_DISPATCH_DTYPE(dtype, T, {
    Src2<T> s{(const T*)x1, (const T*)x2, C1, C2};
    dispatch_mode(act, drop, [&](auto A, auto D) {
      constexpr bool ACT = decltype(A)::value;
      constexpr int DROP = decltype(D)::value;
      auto launch = [&](auto CS, auto NSt) {
        constexpr bool CSUM = decltype(CS)::value;
        constexpr int NS = decltype(NSt)::value;
        constexpr int smem = Pipe<T, NS, bres_for(NS)>::VEC_BYTES;
        auto kernel = gn_bwd_resident_kernel<T, ACT, DROP, CSUM, NS>;
        static bool attr_ok = false;
        if (!attr_ok) {
          if (!allow_smem(kernel, smem)) { rc = ST_ERR_CUDA; return; }
          cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
          if (e != cudaSuccess) { st_set_error("st_gn_bwd_resident: cluster attribute: %s", cudaGetErrorString(e)); rc = ST_ERR_CUDA; return; }
          attr_ok = true;
        }
        cudaLaunchConfig_t cfg;
        memset(&cfg, 0, sizeof(cfg));
        cfg.gridDim = dim3(chunks, n_img);
        cfg.blockDim = dim3(256);
        cfg.dynamicSmemBytes = smem;
        cfg.stream = (cudaStream_t)stream;
        cudaLaunchAttribute attr[2];
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = chunks;
        attr[0].val.clusterDim.y = 1;
        attr[0].val.clusterDim.z = 1;
        attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        attr[1].val.programmaticStreamSerializationAllowed = 1;
        cfg.attrs = attr;
        cfg.numAttrs = st_pdl_on((cudaStream_t)stream) ? 2 : 1;
        cudaError_t e = cudaLaunchKernelEx(&cfg, kernel, s, (const T*)dy, (const T*)extra, extra_scale, hw, G, gamma, beta, mean,
                                           rstd, p_drop, seed, (const T*)mask, keepbits, red, (T*)dx1, (T*)dx2, csum);
        if (e != cudaSuccess) { st_set_error("st_gn_bwd_resident: launch: %s", cudaGetErrorString(e)); rc = ST_ERR_CUDA; }
      };
      auto with_ns = [&](auto CS) {
        if (extra) launch(CS, std::integral_constant<int, 3>{}); else launch(CS, std::integral_constant<int, 2>{});
      };
      if (csum) with_ns(std::true_type{}); else with_ns(std::false_type{});
    });
  });
  if (rc) return rc;
  ST_CHECK_LAUNCH("st_gn_bwd_resident");
  return 0;
}
