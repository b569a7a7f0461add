// Shared pieces of the GroupNorm kernels (groupnorm.cu: streaming passes; groupnorm_cluster.cu: thread-block-cluster
// forms): vector loads / stores, the thread-private cp.async ring, the thread -> (pixel lane, 8-channel vector) walk,
// dropout helpers, the backward math and the bodies of the two backward passes, launch helpers.  Everything has
// internal linkage (anonymous namespace) - each translation unit instantiates what it launches.
#pragma once
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

// Dropout handling is a template parameter so that the streaming loops carry no run-time mode branches:
//   DROP_NONE  p == 0
//   DROP_FAST  forward: in-kernel Philox, keep bits written;  backward: keep bits read back through the pipeline
//   DROP_SLOW  injected mask (parity tests) or, in the backward passes, the generator re-run (no keep bits kept)
enum { DROP_NONE = 0, DROP_FAST = 1, DROP_SLOW = 2 };

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
  {
    // gamma-weighted sums of `red` over the splits: per channel by the whole block (256 threads = (split lane, channel),
    // a short chain of independent loads each; through the not yet used pipeline memory), then per group
    float* s_ch = reinterpret_cast<float*>(gsm);           // [lanes_c][Ct][2]
    const int lanes_c = Ct <= 256 ? 256 / Ct : 1;
    auto channel = [&](int c, int ln) {
      float a = 0.f, b = 0.f;
      const float2* o = reinterpret_cast<const float2*>(red) + (long long)n * splits * Ct + c;
      for (int sp = ln; sp < splits; sp += lanes_c) {
        const float2 v = __ldcg(o + (long long)sp * Ct);
        a += v.x;
        b += v.y;
      }
      const float gm = gamma[c];
      s_ch[((size_t)ln * Ct + c) * 2] = gm * a;
      s_ch[((size_t)ln * Ct + c) * 2 + 1] = gm * b;
    };
    if (lanes_c > 1) {
      if ((int)threadIdx.x < lanes_c * Ct) channel(threadIdx.x % Ct, threadIdx.x / Ct);
    } else {
      for (int c = threadIdx.x; c < Ct; c += 256) channel(c, 0);
    }
    __syncthreads();
    if (threadIdx.x < G) {
      const int g = threadIdx.x;
      double a = 0., b = 0.;
      for (int l = 0; l < lanes_c; ++l)
        for (int c = g * cpg; c < (g + 1) * cpg; ++c) {
          a += (double)s_ch[((size_t)l * Ct + c) * 2];
          b += (double)s_ch[((size_t)l * Ct + c) * 2 + 1];
        }
      const double inv = 1.0 / ((double)hw * cpg);
      sh1[g] = (float)(a * inv);
      sh2[g] = (float)(b * inv);
    }
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

// Image order of a launch.  Consecutive kernels of the step stream the same activation tensors (134-268 MB at B=512,
// L2 = 126 MB): a kernel that walks the images in the order OPPOSITE to its producer starts on what the producer left
// in L2.  The GEMMs walk their row tiles upwards, so the GroupNorm passes that follow a GEMM walk the images downwards
// (ST_GN_ORDER bit 0: forward apply / resident forward, bit 1: single-launch backward forms, bit 2: backward reduction
// of the two-kernel form, whose apply pass then walks upwards again over what the reduction just read).  Measured at
// B=512 (10-step bench lines, one box): forward passes reversed 44.94 ms/step and sampler 26.52 ms/step against 45.43-45.63
// and 26.71 unreversed; reversing the backward forms as well measured 45.21-45.34 -> default 1.
inline int gn_order_bits() {
  static const int v = getenv("ST_GN_ORDER") ? atoi(getenv("ST_GN_ORDER")) : 1;
  return v;
}
__device__ __forceinline__ int img_of(int idx, int n_img, int rev) { return rev ? n_img - 1 - idx : idx; }

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
