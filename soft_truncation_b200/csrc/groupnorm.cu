// GroupNorm (+SiLU, +dropout) forward and backward on NHWC tensors whose channel axis may be the
// concatenation of two tensors (the U-Net skip concatenation is never materialised).
// Reference: nn.GroupNorm(min(C//4,32), C, eps=1e-6) -> act -> Dropout, models/layerspp.py:232,244-245,
// 258,275-278; models/ncsnpp.py:219-253.
//
// All kernels are HBM-bound streaming passes.  Thread mapping: one thread owns a fixed run of 8 channels
// (16-byte bf16 / 32-byte fp32 vectors) of one image and walks that image's pixels, so per-channel constants
// (gamma, beta, mean, rstd) sit in registers and the inner loop has no integer division; two pixels are in
// flight per iteration.  Reductions are deterministic (fixed-order shared-memory trees, no float atomics).
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
  __device__ __forceinline__ void issue_byte(int stage, const uint8_t* bits, long long idx) const {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(side + (uint32_t)(stage * 256 * 4)), "l"(bits + (idx & ~3LL))
                 : "memory");
  }
  __device__ __forceinline__ uint32_t read_byte(int stage, long long idx) const {
    uint32_t w;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(w) : "r"(side + (uint32_t)(stage * 256 * 4)));
    return (w >> (8 * (int)(idx & 3))) & 0xFFu;
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

// ---------------------------------------------------------------- stats
// grid (n_img, splits): part[n][split][G][2] = (sum, sum of squares) over the split's pixels
template <typename T>
__global__ void __launch_bounds__(256) gn_stats_kernel(Src2<T> s, int hw, int G, int splits, float* part) {
  extern __shared__ __align__(16) uint8_t gsm[];
  using P = Pipe<T, 1, GN_DEPTH>;
  const P pipe(gsm);
  const int Ct = s.C1 + s.C2, V = Ct / 8, cpg = Ct / G;
  const int n = blockIdx.x, sp = blockIdx.y;
  const int lanes = 256 / V;
  const int v = threadIdx.x % V, lane = threadIdx.x / V;
  const int per = (hw + splits - 1) / splits;
  const int p0 = sp * per, p1 = min(hw, p0 + per);
  float sum[2] = {0.f, 0.f}, sq[2] = {0.f, 0.f};
  const int c0 = v * 8;
  const int n_it = (lane < lanes && p1 > p0 + lane) ? (p1 - p0 - lane + lanes - 1) / lanes : 0;
  const long long row0 = (long long)n * hw + p0 + lane;
  for (int d = 0; d < GN_DEPTH; ++d) {
    if (d < n_it) pipe.issue(d, 0, s.at(row0 + (long long)d * lanes, c0));
    P::commit();
  }
  int stage = 0;
  for (int it = 0; it < n_it; ++it) {
    P::wait();
    float a[8];
    pipe.read(stage, 0, a);
#pragma unroll
    for (int i = 0; i < 8; ++i) { sum[i >> 2] += a[i]; sq[i >> 2] = fmaf(a[i], a[i], sq[i >> 2]); }
    if (it + GN_DEPTH < n_it) pipe.issue(stage, 0, s.at(row0 + (long long)(it + GN_DEPTH) * lanes, c0));
    P::commit();
    stage = stage + 1 == GN_DEPTH ? 0 : stage + 1;
  }
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

// ---------------------------------------------------------------- apply
// grid (chunks, n_img); pipelined stream: x (an injected dropout mask - parity tests only - is read directly)
template <typename T>
__global__ void __launch_bounds__(256) gn_apply_kernel(Src2<T> s, int hw, int G, const float* __restrict__ gamma,
                                                       const float* __restrict__ beta, const float* __restrict__ mean,
                                                       const float* __restrict__ rstd, int act, float p_drop, uint64_t seed,
                                                       const T* mask, uint8_t* keepbits, T* y) {
  extern __shared__ __align__(16) uint8_t gsm[];
  using P = Pipe<T, 1, GN_DEPTH>;
  const P pipe(gsm);
  const int Ct = s.C1 + s.C2, V = Ct / 8, cpg = Ct / G;
  const int n = blockIdx.y;
  const int lanes = 256 / V;
  const int v = threadIdx.x % V, lane = threadIdx.x / V;
  if (lane >= lanes) return;
  const int c0 = v * 8;
  ChanConst k;
  load_consts(k, n, c0, G, cpg, gamma, beta, mean, rstd);
  const int per = (hw + gridDim.x - 1) / gridDim.x;
  const int p0 = blockIdx.x * per, p1 = min(hw, p0 + per);
  const int n_it = p1 > p0 + lane ? (p1 - p0 - lane + lanes - 1) / lanes : 0;
  const long long row0 = (long long)n * hw + p0 + lane;
  auto issue = [&](int stage, int j) {
    const long long row = row0 + (long long)j * lanes;
    pipe.issue(stage, 0, s.at(row, c0));
  };
  for (int d = 0; d < GN_DEPTH; ++d) {
    if (d < n_it) issue(d, d);
    P::commit();
  }
  int stage = 0;
  for (int it = 0; it < n_it; ++it) {
    P::wait();
    const long long oct = (row0 + (long long)it * lanes) * V + v;
    float x[8], o[8];
    pipe.read(stage, 0, x);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      float u = fmaf((x[i] - k.mu[i >> 2]) * k.r[i >> 2], k.gam[i], k.bet[i]);
      o[i] = act ? silu_f(u) : u;
    }
    if (mask) {
      float mk[8];
      load8(mask + oct * 8, mk);
#pragma unroll
      for (int i = 0; i < 8; ++i) o[i] *= mk[i];
    } else if (p_drop > 0.f) {
      float keep[8];
      const uint32_t bits = dropout8(seed, (uint64_t)oct, p_drop, keep);
      if (keepbits) keepbits[oct] = (uint8_t)bits;
#pragma unroll
      for (int i = 0; i < 8; ++i) o[i] *= keep[i];
    }
    store8(y + oct * 8, o);
    if (it + GN_DEPTH < n_it) issue(stage, it + GN_DEPTH);
    P::commit();
    stage = stage + 1 == GN_DEPTH ? 0 : stage + 1;
  }
}

// dz and xhat of one 8-vector (shared by both backward passes)
// `mk` holds the injected mask values when has_mask, else it is filled here (in-kernel RNG or ones)
__device__ __forceinline__ void gn_dz8(const float x[8], const float dyv[8], const ChanConst& k, int act, float p_drop,
                                       uint64_t seed, bool has_mask, float mk[8], bool has_bits, uint32_t bits, long long oct,
                                       float xhat[8], float dz[8]) {
  if (!has_mask) {
    if (p_drop > 0.f && has_bits) keep_from_bits(bits, p_drop, mk);
    else if (p_drop > 0.f) dropout8(seed, (uint64_t)oct, p_drop, mk);
    else {
#pragma unroll
      for (int i = 0; i < 8; ++i) mk[i] = 1.f;
    }
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    xhat[i] = (x[i] - k.mu[i >> 2]) * k.r[i >> 2];
    float d = dyv[i] * mk[i];
    if (act) d *= silu_grad_f(fmaf(xhat[i], k.gam[i], k.bet[i]));
    dz[i] = d;
  }
}

// ---------------------------------------------------------------- backward pass 1
// grid (n_img, splits): red[n][split][c][2] = (sum dz, sum dz*xhat) over the split's pixels
template <typename T>
__global__ void __launch_bounds__(256, 3) gn_bwd_reduce_kernel(Src2<T> s, const T* dy, int hw, int G, int splits,
                                                            const float* __restrict__ gamma, const float* __restrict__ beta,
                                                            const float* __restrict__ mean, const float* __restrict__ rstd,
                                                            int act, float p_drop, uint64_t seed, const T* mask,
                                                            const uint8_t* __restrict__ keepbits, float* red) {
  const int Ct = s.C1 + s.C2, V = Ct / 8, cpg = Ct / G;
  const int n = blockIdx.x, sp = blockIdx.y;
  const int lanes = 256 / V;
  const int v = threadIdx.x % V, lane = threadIdx.x / V;
  extern __shared__ __align__(16) uint8_t gsm[];
  const int per = (hw + splits - 1) / splits;
  const int p0 = sp * per, p1 = min(hw, p0 + per);
  const int c0 = v * 8;
  float a[8], b[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) a[i] = b[i] = 0.f;
  using P = Pipe<T, 2, GN_DEPTH>;            // streams: x, dy (+ the keep-bits side stream)
  const bool use_bits = keepbits != nullptr && p_drop > 0.f && mask == nullptr;
  const P pipe(gsm);
  const int n_it = (lane < lanes && p1 > p0 + lane) ? (p1 - p0 - lane + lanes - 1) / lanes : 0;
  const long long row0 = (long long)n * hw + p0 + lane;
  ChanConst k;
  if (lane < lanes) load_consts(k, n, c0, G, cpg, gamma, beta, mean, rstd);
  auto issue = [&](int stage, int j) {
    const long long row = row0 + (long long)j * lanes;
    pipe.issue(stage, 0, s.at(row, c0));
    pipe.issue(stage, 1, dy + (row * V + v) * 8);
    if (use_bits) pipe.issue_byte(stage, keepbits, row * V + v);
  };
  for (int d = 0; d < GN_DEPTH; ++d) {
    if (d < n_it) issue(d, d);
    P::commit();
  }
  int stage = 0;
  for (int it = 0; it < n_it; ++it) {
    P::wait();
    const long long oct = (row0 + (long long)it * lanes) * V + v;
    float x0[8], d0[8], mk[8], xh[8], dz[8];
    pipe.read(stage, 0, x0);
    pipe.read(stage, 1, d0);
    if (mask) load8(mask + oct * 8, mk);
    gn_dz8(x0, d0, k, act, p_drop, seed, mask != nullptr, mk, use_bits, use_bits ? pipe.read_byte(stage, oct) : 0u, oct, xh, dz);
#pragma unroll
    for (int i = 0; i < 8; ++i) { a[i] += dz[i]; b[i] = fmaf(dz[i], xh[i], b[i]); }
    if (it + GN_DEPTH < n_it) issue(stage, it + GN_DEPTH);
    P::commit();
    stage = stage + 1 == GN_DEPTH ? 0 : stage + 1;
  }
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
template <typename T, bool CSUM>
__global__ void __launch_bounds__(256, CSUM ? 2 : 3) gn_bwd_apply_kernel(Src2<T> s, const T* dy, int hw, int G, int splits,
                                                           const float* __restrict__ gamma, const float* __restrict__ beta,
                                                           const float* __restrict__ mean, const float* __restrict__ rstd,
                                                           int act, float p_drop, uint64_t seed, const T* mask,
                                                           const uint8_t* __restrict__ keepbits,
                                                           const float* __restrict__ red, const T* extra, float extra_scale,
                                                           T* dx1, int accum1, T* dx2, int accum2, float* csum) {
  extern __shared__ __align__(16) uint8_t gsm[];
  const int Ct = s.C1 + s.C2, V = Ct / 8, cpg = Ct / G;
  const int n = blockIdx.y;
  __shared__ float sh1[64], sh2[64];
  if (threadIdx.x < G) {
    const int g = threadIdx.x;
    double a = 0., b = 0.;
    for (int sp = 0; sp < splits; ++sp)
      for (int c = g * cpg; c < (g + 1) * cpg; ++c) {
        const float* o = red + (((long long)n * splits + sp) * Ct + c) * 2;
        a += (double)gamma[c] * (double)o[0];
        b += (double)gamma[c] * (double)o[1];
      }
    const double inv = 1.0 / ((double)hw * cpg);
    sh1[g] = (float)(a * inv);
    sh2[g] = (float)(b * inv);
  }
  __syncthreads();
  const int lanes = 256 / V;
  const int v = threadIdx.x % V, lane = threadIdx.x / V;
  const bool active = lane < lanes;
  const int c0 = active ? v * 8 : 0;
  ChanConst k;
  load_consts(k, n, c0, G, cpg, gamma, beta, mean, rstd);
  const float s1[2] = {sh1[k.g[0]], sh1[k.g[1]]}, s2[2] = {sh2[k.g[0]], sh2[k.g[1]]};
  float cs[8];                               // column sums of this thread's contributions (optional output)
#pragma unroll
  for (int i = 0; i < 8; ++i) cs[i] = 0.f;
  // destination of this thread's channels (first or second tensor of the concatenation)
  T* const dbase = c0 < s.C1 ? dx1 + c0 : dx2 + (c0 - s.C1);
  const int dld = c0 < s.C1 ? s.C1 : s.C2;
  const int acc = c0 < s.C1 ? accum1 : accum2;
  const int per = (hw + gridDim.x - 1) / gridDim.x;
  const int p0 = blockIdx.x * per, p1 = min(hw, p0 + per);
  using P = Pipe<T, 4, GN_BWD_DEPTH>;        // streams: x, dy, extra, old destination (+ the keep-bits side stream)
  const bool use_bits = keepbits != nullptr && p_drop > 0.f && mask == nullptr;
  const P pipe(gsm);
  const int n_it = (active && p1 > p0 + lane) ? (p1 - p0 - lane + lanes - 1) / lanes : 0;
  const long long row0 = (long long)n * hw + p0 + lane;
  auto issue = [&](int stage, int j) {
    const long long row = row0 + (long long)j * lanes;
    pipe.issue(stage, 0, s.at(row, c0));
    pipe.issue(stage, 1, dy + (row * V + v) * 8);
    if (extra) pipe.issue(stage, 2, extra + (row * V + v) * 8);
    if (acc) pipe.issue(stage, 3, dbase + row * dld);
    if (use_bits) pipe.issue_byte(stage, keepbits, row * V + v);
  };
  for (int d = 0; d < GN_BWD_DEPTH; ++d) {
    if (d < n_it) issue(d, d);
    P::commit();
  }
  int stage = 0;
  for (int it = 0; it < n_it; ++it) {
    P::wait();
    const long long row = row0 + (long long)it * lanes;
    float x0[8], d0[8], mk[8], xh[8], dz[8], o[8];
    pipe.read(stage, 0, x0);
    pipe.read(stage, 1, d0);
    if (mask) load8(mask + (row * V + v) * 8, mk);
    gn_dz8(x0, d0, k, act, p_drop, seed, mask != nullptr, mk, use_bits, use_bits ? pipe.read_byte(stage, row * V + v) : 0u,
           row * V + v, xh, dz);
#pragma unroll
    for (int i = 0; i < 8; ++i) o[i] = k.r[i >> 2] * (k.gam[i] * dz[i] - s1[i >> 2] - xh[i] * s2[i >> 2]);
    if (extra) {
      float ex[8];
      pipe.read(stage, 2, ex);
#pragma unroll
      for (int i = 0; i < 8; ++i) o[i] = fmaf(extra_scale, ex[i], o[i]);
    }
    if constexpr (CSUM) {
#pragma unroll
      for (int i = 0; i < 8; ++i) cs[i] += o[i];
    }
    if (acc) {
      float old[8];
      pipe.read(stage, 3, old);
#pragma unroll
      for (int i = 0; i < 8; ++i) o[i] += old[i];
    }
    store8(dbase + row * dld, o);
    if (it + GN_BWD_DEPTH < n_it) issue(stage, it + GN_BWD_DEPTH);
    P::commit();
    stage = stage + 1 == GN_BWD_DEPTH ? 0 : stage + 1;
  }
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
      csum[((long long)n * gridDim.x + blockIdx.x) * Ct + col] = t;
    }
  }
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
  int want = (st_num_sms() * 8 + n_img - 1) / n_img;
  int c = want < max_chunks ? want : max_chunks;
  if (c > 65535) c = 65535;
  return c < 1 ? 1 : c;
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
    gn_stats_kernel<T><<<dim3(n_img, splits), 256, smem, (cudaStream_t)stream>>>(s, hw, G, splits, part);
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
                           const float* gamma, const float* beta, const float* mean, const float* rstd, int act,
                           float p_drop, uint64_t seed, const void* mask, uint8_t* keepbits, void* y, void* stream) {
  if (int e = check_geom(C1, C2, G)) return e;
  ST_CHECK_ARG(n_img <= 65535, "st_gn_apply: more than 65535 images");
  const int V = (C1 + C2) / 8;
  ST_DISPATCH_DTYPE(dtype, T, {
    Src2<T> s{(const T*)x1, (const T*)x2, C1, C2};
    constexpr int smem = Pipe<T, 1, GN_DEPTH>::BYTES;
    static bool smem_ok = false;
    if (!smem_ok) { if (!allow_smem(gn_apply_kernel<T>, smem)) return ST_ERR_CUDA; smem_ok = true; }
    gn_apply_kernel<T><<<dim3(chunks_for(n_img, hw, V), n_img), 256, smem, (cudaStream_t)stream>>>(
        s, hw, G, gamma, beta, mean, rstd, act, p_drop, seed, (const T*)mask, keepbits, (T*)y);
  });
  ST_CHECK_LAUNCH("st_gn_apply");
  return 0;
}

extern "C" __attribute__((visibility("default"))) int st_gn_bwd_reduce(const void* x1, const void* x2, const void* dy, int dtype, int n_img, int hw, int C1,
                                int C2, int G, const float* gamma, const float* beta, const float* mean,
                                const float* rstd, int act, float p_drop, uint64_t seed, const void* mask,
                                const uint8_t* keepbits, int splits, float* red, void* stream) {
  if (int e = check_geom(C1, C2, G)) return e;
  const int V = (C1 + C2) / 8;
  (void)V;
  ST_DISPATCH_DTYPE(dtype, T, {
    Src2<T> s{(const T*)x1, (const T*)x2, C1, C2};
    constexpr int smem = Pipe<T, 2, GN_DEPTH>::BYTES;      // >= the 16 KB the final lane reduction reuses
    static bool smem_ok = false;
    if (!smem_ok) { if (!allow_smem(gn_bwd_reduce_kernel<T>, smem)) return ST_ERR_CUDA; smem_ok = true; }
    gn_bwd_reduce_kernel<T><<<dim3(n_img, splits), 256, smem, (cudaStream_t)stream>>>(
        s, (const T*)dy, hw, G, splits, gamma, beta, mean, rstd, act, p_drop, seed, (const T*)mask, keepbits, red);
  });
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
                               int accum1, void* dx2, int accum2, int chunks, float* csum, void* stream) {
  if (int e = check_geom(C1, C2, G)) return e;
  ST_CHECK_ARG(n_img <= 65535, "st_gn_bwd_apply: more than 65535 images");
  const int V = (C1 + C2) / 8;
  ST_CHECK_ARG(!csum || chunks > 0, "st_gn_bwd_apply: csum needs an explicit chunk count");
  if (chunks <= 0) chunks = chunks_for(n_img, hw, V);
  ST_CHECK_ARG(chunks <= 65535, "st_gn_bwd_apply: too many chunks");
  ST_DISPATCH_DTYPE(dtype, T, {
    Src2<T> s{(const T*)x1, (const T*)x2, C1, C2};
    constexpr int smem = Pipe<T, 4, GN_BWD_DEPTH>::BYTES;
    static bool smem_ok = false;
    if (!smem_ok) {
      if (!allow_smem(gn_bwd_apply_kernel<T, false>, smem) || !allow_smem(gn_bwd_apply_kernel<T, true>, smem)) return ST_ERR_CUDA;
      smem_ok = true;
    }
    if (csum)
      gn_bwd_apply_kernel<T, true><<<dim3(chunks, n_img), 256, smem, (cudaStream_t)stream>>>(
          s, (const T*)dy, hw, G, splits, gamma, beta, mean, rstd, act, p_drop, seed, (const T*)mask, keepbits, red,
          (const T*)extra, extra_scale, (T*)dx1, accum1, (T*)dx2, accum2, csum);
    else
      gn_bwd_apply_kernel<T, false><<<dim3(chunks, n_img), 256, smem, (cudaStream_t)stream>>>(
          s, (const T*)dy, hw, G, splits, gamma, beta, mean, rstd, act, p_drop, seed, (const T*)mask, keepbits, red,
          (const T*)extra, extra_scale, (T*)dx1, accum1, (T*)dx2, accum2, csum);
  });
  ST_CHECK_LAUNCH("st_gn_bwd_apply");
  return 0;
}
