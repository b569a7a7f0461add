// HBM-bound elementwise / small-reduction kernels of the score-network hot path.
#include "common.cuh"

namespace {

inline int grid1d(long long items, int per_block) {
  long long b = (items + per_block - 1) / per_block;
  long long cap = (long long)st_num_sms() * 16;
  return (int)(b < 1 ? 1 : (b > cap ? cap : b));
}

#define GRID_STRIDE(i, n) \
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < (n); i += (long long)gridDim.x * blockDim.x)

// ---------------------------------------------------------------- cast / axpby / silu
template <typename TS, typename TD>
__global__ void cast_kernel(const TS* __restrict__ s, TD* __restrict__ d, long long n4, long long n) {
  GRID_STRIDE(i, n4) {
    float v[4];
    load4(s + i * 4, v);
    store4(d + i * 4, v);
  }
  if (blockIdx.x == 0 && threadIdx.x < (n & 3)) {
    long long j = (n & ~3LL) + threadIdx.x;
    d[j] = from_f<TD>(to_f(s[j]));
  }
}

template <typename T>
__global__ void axpby_kernel(const T* __restrict__ a, const T* __restrict__ b, T* __restrict__ o, float alpha,
                             float beta, long long n4, long long n) {
  GRID_STRIDE(i, n4) {
    float va[4], vb[4] = {0.f, 0.f, 0.f, 0.f}, r[4];
    load4(a + i * 4, va);
    if (b) load4(b + i * 4, vb);
#pragma unroll
    for (int k = 0; k < 4; ++k) r[k] = alpha * va[k] + beta * vb[k];
    store4(o + i * 4, r);
  }
  if (blockIdx.x == 0 && threadIdx.x < (n & 3)) {
    long long j = (n & ~3LL) + threadIdx.x;
    o[j] = from_f<T>(alpha * to_f(a[j]) + (b ? beta * to_f(b[j]) : 0.f));
  }
}

template <typename T>
__global__ void silu_kernel(const T* __restrict__ x, T* __restrict__ y, long long n) {
  GRID_STRIDE(i, n) y[i] = from_f<T>(silu_f(to_f(x[i])));
}
template <typename T>
__global__ void silu_bwd_kernel(const T* __restrict__ x, const T* __restrict__ dy, T* __restrict__ dx, long long n) {
  GRID_STRIDE(i, n) dx[i] = from_f<T>(to_f(dy[i]) * silu_grad_f(to_f(x[i])));
}

// ---------------------------------------------------------------- transposed conv weights for the data gradient
// dst[ci][taps-1-t][co] = src[co][t][ci] for every convolution of the table, one launch: the data gradient of a
// convolution then runs as a forward convolution over dY with a K-major weight operand (one 256-row TMA box per K
// block instead of four MN-major 64x64 boxes: +8 % at 16x16, +28 % at 8x8, tools/mn_probe.py).
// table[e] = {offset (elements), Co, taps, Ci}; tile_prefix[e] = first 32x32 tile of entry e.
template <typename T>
__global__ void __launch_bounds__(256) transpose_conv_weights_kernel(const T* __restrict__ src, T* __restrict__ dst,
                                                                     const long long* __restrict__ table,
                                                                     const long long* __restrict__ tile_prefix, int n_entries) {
  __shared__ T tile[32][33];
  const long long t = blockIdx.x;
  int lo = 0, hi = n_entries - 1;              // last entry whose first tile is <= t
  while (lo < hi) {
    const int mid = (lo + hi + 1) >> 1;
    if (tile_prefix[mid] <= t) lo = mid; else hi = mid - 1;
  }
  const long long off = table[4 * lo];
  const int Co = (int)table[4 * lo + 1], taps = (int)table[4 * lo + 2], Ci = (int)table[4 * lo + 3];
  const int tci = (Ci + 31) / 32, tco = (Co + 31) / 32;
  int r = (int)(t - tile_prefix[lo]);
  const int tap = r / (tco * tci);
  r -= tap * tco * tci;
  const int co0 = (r / tci) * 32, ci0 = (r % tci) * 32;
  const int tx = threadIdx.x % 32, ty = threadIdx.x / 32;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int co = co0 + ty + 8 * k, ci = ci0 + tx;
    if (co < Co && ci < Ci) tile[ty + 8 * k][tx] = src[off + ((long long)co * taps + tap) * Ci + ci];
  }
  __syncthreads();
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int ci = ci0 + ty + 8 * k, co = co0 + tx;
    if (co < Co && ci < Ci) dst[off + ((long long)ci * taps + (taps - 1 - tap)) * Co + co] = tile[tx][ty + 8 * k];
  }
}

// ---------------------------------------------------------------- batched column sums
// One launch for every small fp32 reduction a backward pass queues up (bias / time-embedding gradients from the
// column-sum partials of the GroupNorm backward kernels): ~150 launches of ~12 us each otherwise.
struct ColsumJob {                 // every field is 8 bytes: built as an int64 table on the host
  const float* part[4];            // kind 0: up to four [rows][ld] partial tables reduced into the same destination
  long long rows[4];
  long long ld[4];
  float* dst;
  long long n_parts;
  long long kind;                  // 0: dst[c] (+)= scale * sum_parts sum_rows part[r][c];  1: dst[g][c] = sum_k part0[g*K + k][c]
                                   // 2: GroupNorm dgamma (dst) / dbeta (part[1]) += column sums of red = part[0] [rows][C][2]
  long long C;
  long long groups, rows_per_group, ld_out;
  double scale;
  long long accumulate;
};
__global__ void __launch_bounds__(1024) colsum_batched_kernel(const ColsumJob* __restrict__ jobs) {
  pdl_wait();
  pdl_trigger();
  const ColsumJob& j = jobs[blockIdx.x];
  const int C = (int)j.C;
  if (j.kind == 0) {
    if ((int)blockIdx.y * 128 >= C) return;
    const int c0 = (blockIdx.y * 32 + threadIdx.x % 32) * 4;
    const int rl = threadIdx.x / 32;
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
    if (c0 < C) {
      for (int p = 0; p < (int)j.n_parts; ++p) {
        const float* base = j.part[p] + c0;
        const long long rows = j.rows[p], ld = j.ld[p];
        for (long long r = rl; r < rows; r += 32) {
          const float4 v = *reinterpret_cast<const float4*>(base + r * ld);
          acc[0] += v.x; acc[1] += v.y; acc[2] += v.z; acc[3] += v.w;
        }
      }
    }
    __shared__ float4 sm[1024];
    sm[threadIdx.x] = make_float4(acc[0], acc[1], acc[2], acc[3]);
    __syncthreads();
    if (rl == 0 && c0 < C) {
      float4 t = sm[threadIdx.x];
      for (int l = 1; l < 32; ++l) {
        const float4 u = sm[l * 32 + threadIdx.x];
        t.x += u.x; t.y += u.y; t.z += u.z; t.w += u.w;
      }
      const float sc = (float)j.scale;
      float* o = j.dst + c0;
      if (j.accumulate) { o[0] += sc * t.x; o[1] += sc * t.y; o[2] += sc * t.z; o[3] += sc * t.w; }
      else { o[0] = sc * t.x; o[1] = sc * t.y; o[2] = sc * t.z; o[3] = sc * t.w; }
    }
  } else if (j.kind == 2) {
    // GroupNorm parameter gradients from the backward reduction buffer red[rows][C][2] = (sum dz, sum dz*xhat):
    // dst[c] += sum_rows red[r][c][1] (dgamma), part[1][c] += sum_rows red[r][c][0] (dbeta); a float4 = 2 channels
    const int C2 = 2 * C;
    if ((int)blockIdx.y * 128 >= C2) return;
    const int c0 = (blockIdx.y * 32 + threadIdx.x % 32) * 4;
    const int rl = threadIdx.x / 32;
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
    if (c0 < C2) {
      const float* base = j.part[0] + c0;
      for (long long r = rl; r < j.rows[0]; r += 32) {
        const float4 v = *reinterpret_cast<const float4*>(base + r * C2);
        acc[0] += v.x; acc[1] += v.y; acc[2] += v.z; acc[3] += v.w;
      }
    }
    __shared__ float4 sm2[1024];
    sm2[threadIdx.x] = make_float4(acc[0], acc[1], acc[2], acc[3]);
    __syncthreads();
    if (rl == 0 && c0 < C2) {
      double t[4] = {0., 0., 0., 0.};
      for (int l = 0; l < 32; ++l) {
        const float4 u = sm2[l * 32 + threadIdx.x];
        t[0] += (double)u.x; t[1] += (double)u.y; t[2] += (double)u.z; t[3] += (double)u.w;
      }
      float* dgamma = j.dst + c0 / 2;
      float* dbeta = const_cast<float*>(j.part[1]) + c0 / 2;
      dbeta[0] += (float)t[0]; dgamma[0] += (float)t[1];
      dbeta[1] += (float)t[2]; dgamma[1] += (float)t[3];
    }
  } else {
    const int Q = C / 4;
    const long long idx = (long long)blockIdx.y * 1024 + threadIdx.x;
    if (idx >= j.groups * Q) return;
    const long long g = idx / Q;
    const int q = (int)(idx % Q);
    const float* base = j.part[0] + g * j.rows_per_group * j.ld[0] + q * 4;
    float4 t = make_float4(0.f, 0.f, 0.f, 0.f);
    for (long long k = 0; k < j.rows_per_group; ++k) {
      const float4 v = *reinterpret_cast<const float4*>(base + k * j.ld[0]);
      t.x += v.x; t.y += v.y; t.z += v.z; t.w += v.w;
    }
    const float sc = (float)j.scale;
    float* o = j.dst + g * j.ld_out + q * 4;
    if (j.accumulate) { o[0] += sc * t.x; o[1] += sc * t.y; o[2] += sc * t.z; o[3] += sc * t.w; }
    else { o[0] = sc * t.x; o[1] = sc * t.y; o[2] = sc * t.z; o[3] = sc * t.w; }
  }
}

// ---------------------------------------------------------------- 2x resampling
// dir=+1: y[n][2i+a][2j+b][c] = scale*x[n][i][j][c];  dir=-1: y[n][i][j][c] = scale*sum_ab x[n][2i+a][2j+b][c]
template <typename T>
__global__ void up2_kernel(const T* x1, const T* x2, T* y, long long total_quads, int H, int W, int C1, int C2, float scale) {
  const int Ct = C1 + C2, Q = Ct / 4;
  GRID_STRIDE(gq, total_quads) {   // over OUTPUT quads
    int quad = (int)(gq % Q);
    long long row = gq / Q;
    int ox = (int)(row % (2 * W));
    long long t = row / (2 * W);
    int oy = (int)(t % (2 * H));
    long long n = t / (2 * H);
    long long irow = (n * H + oy / 2) * W + ox / 2;
    int c0 = quad * 4;
    float v[4];
    load4(c0 < C1 ? x1 + irow * C1 + c0 : x2 + irow * C2 + (c0 - C1), v);
#pragma unroll
    for (int k = 0; k < 4; ++k) v[k] *= scale;
    store4(y + gq * 4, v);
  }
}
template <typename T>
__global__ void down2_kernel(const T* x1, const T* x2, T* y, long long total_quads, int H, int W, int C1, int C2, float scale) {
  const int Ct = C1 + C2, Q = Ct / 4;     // H, W are INPUT sizes
  const int Ho = H / 2, Wo = W / 2;
  GRID_STRIDE(gq, total_quads) {   // over OUTPUT quads
    int quad = (int)(gq % Q);
    long long row = gq / Q;
    int ox = (int)(row % Wo);
    long long t = row / Wo;
    int oy = (int)(t % Ho);
    long long n = t / Ho;
    int c0 = quad * 4;
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int a = 0; a < 2; ++a)
#pragma unroll
      for (int b = 0; b < 2; ++b) {
        long long irow = (n * H + 2 * oy + a) * W + 2 * ox + b;
        float v[4];
        load4(c0 < C1 ? x1 + irow * C1 + c0 : x2 + irow * C2 + (c0 - C1), v);
#pragma unroll
        for (int k = 0; k < 4; ++k) acc[k] += v[k];
      }
#pragma unroll
    for (int k = 0; k < 4; ++k) acc[k] *= scale;
    store4(y + gq * 4, acc);
  }
}

// 8-channel (16-byte bf16 / 32-byte fp32) versions with 32-bit index arithmetic: one thread per INPUT vector writes
// its four copies (up) / one thread per OUTPUT vector sums its four inputs (down).
__device__ __forceinline__ void ld8(const float* p, float v[8]) { load4(p, v); load4(p + 4, v + 4); }
__device__ __forceinline__ void ld8(const bf16* p, float v[8]) {
  uint4 t = *reinterpret_cast<const uint4*>(p);
  const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&t);
#pragma unroll
  for (int i = 0; i < 4; ++i) { v[2 * i] = __low2float(h[i]); v[2 * i + 1] = __high2float(h[i]); }
}
__device__ __forceinline__ void st8(float* p, const float v[8]) { store4(p, v); store4(p + 4, v + 4); }
__device__ __forceinline__ void st8(bf16* p, const float v[8]) {
  uint4 t;
  __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&t);
#pragma unroll
  for (int i = 0; i < 4; ++i) h[i] = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]);
  *reinterpret_cast<uint4*>(p) = t;
}
template <typename T>
__global__ void __launch_bounds__(256) up2v_kernel(const T* __restrict__ x1, const T* __restrict__ x2, T* __restrict__ y,
                                                   unsigned total, int H, int W, int C1, int C2, float scale) {
  const unsigned V = (unsigned)(C1 + C2) / 8;
  for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {   // INPUT vectors
    const unsigned v = i % V, row = i / V;
    const unsigned ix = row % (unsigned)W, t = row / (unsigned)W;
    const unsigned iy = t % (unsigned)H, n = t / (unsigned)H;
    const int c0 = (int)v * 8;
    float val[8];
    ld8(c0 < C1 ? x1 + (size_t)row * C1 + c0 : x2 + (size_t)row * C2 + (c0 - C1), val);
#pragma unroll
    for (int k = 0; k < 8; ++k) val[k] *= scale;
    const size_t orow = ((size_t)n * 2 * H + 2 * iy) * (2 * W) + 2 * ix;
    T* o = y + (orow * V + v) * 8;
    const size_t dn = (size_t)2 * W * V * 8;
    st8(o, val);
    st8(o + V * 8, val);
    st8(o + dn, val);
    st8(o + dn + V * 8, val);
  }
}
template <typename T>
__global__ void __launch_bounds__(256) down2v_kernel(const T* __restrict__ x1, const T* __restrict__ x2, T* __restrict__ y,
                                                     unsigned total, int H, int W, int C1, int C2, float scale) {
  const unsigned V = (unsigned)(C1 + C2) / 8;      // H, W are INPUT sizes
  const unsigned Ho = H / 2, Wo = W / 2;
  for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {   // OUTPUT vectors
    const unsigned v = i % V, row = i / V;
    const unsigned ox = row % Wo, t = row / Wo;
    const unsigned oy = t % Ho, n = t / Ho;
    const int c0 = (int)v * 8;
    const size_t irow = ((size_t)n * H + 2 * oy) * W + 2 * ox;
    const T* src;
    size_t ld;
    if (c0 < C1) { src = x1 + irow * C1 + c0; ld = C1; } else { src = x2 + irow * C2 + (c0 - C1); ld = C2; }
    float a[8], b[8], c[8], d[8];
    ld8(src, a);
    ld8(src + ld, b);
    ld8(src + (size_t)W * ld, c);
    ld8(src + (size_t)W * ld + ld, d);
#pragma unroll
    for (int k = 0; k < 8; ++k) a[k] = (((a[k] + b[k]) + c[k]) + d[k]) * scale;     // same order as the 4-wide kernel
    st8(y + (size_t)i * 8, a);
  }
}

// ---------------------------------------------------------------- column sums
// grid (groups, ceil(C/128)), block 1024 = 32 row-lanes x 32 quad-lanes
template <typename T>
__global__ void __launch_bounds__(1024) colsum_kernel(const T* __restrict__ x, long long rows_per_group, int C, long long ld,
                                                      float scale, float* out, int accumulate) {
  constexpr int RL = 32;
  pdl_wait();
  pdl_trigger();
  const long long g = blockIdx.x;
  const int c0 = (blockIdx.y * 32 + threadIdx.x % 32) * 4;
  const int rl = threadIdx.x / 32;
  float acc[4] = {0.f, 0.f, 0.f, 0.f};
  if (c0 < C) {
    const T* base = x + g * rows_per_group * ld + c0;
    for (long long r = rl; r < rows_per_group; r += RL) {
      float v[4];
      load4(base + r * ld, v);
#pragma unroll
      for (int k = 0; k < 4; ++k) acc[k] += v[k];
    }
  }
  __shared__ float4 sm[1024];
  sm[threadIdx.x] = make_float4(acc[0], acc[1], acc[2], acc[3]);
  __syncthreads();
  if (rl == 0 && c0 < C) {
    float4 t = sm[threadIdx.x];
    for (int l = 1; l < RL; ++l) {
      float4 u = sm[l * 32 + threadIdx.x];
      t.x += u.x; t.y += u.y; t.z += u.z; t.w += u.w;
    }
    float* o = out + g * C + c0;
    if (accumulate) { o[0] += scale * t.x; o[1] += scale * t.y; o[2] += scale * t.z; o[3] += scale * t.w; }
    else { o[0] = scale * t.x; o[1] = scale * t.y; o[2] = scale * t.z; o[3] = scale * t.w; }
  }
}

// ---------------------------------------------------------------- softmax (one warp per row)
template <typename T>
__global__ void __launch_bounds__(256) softmax_fwd_kernel(const float* __restrict__ logits, T* __restrict__ p, long long rows,
                                                          int L, float scale) {
  const int lane = threadIdx.x % 32;
  for (long long row = (long long)blockIdx.x * 8 + threadIdx.x / 32; row < rows; row += (long long)gridDim.x * 8) {
    const float* src = logits + row * L;
    float mx = -INFINITY;
    for (int j = lane; j < L; j += 32) mx = fmaxf(mx, src[j] * scale);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    float sum = 0.f;
    for (int j = lane; j < L; j += 32) sum += __expf(src[j] * scale - mx);
    sum = warp_sum(sum);
    const float inv = 1.f / sum;
    for (int j = lane; j < L; j += 32) p[row * L + j] = from_f<T>(__expf(src[j] * scale - mx) * inv);
  }
}
template <typename T>
__global__ void __launch_bounds__(256) softmax_bwd_kernel(const T* __restrict__ p, const float* __restrict__ dp,
                                                          T* __restrict__ ds, long long rows, int L, float scale) {
  const int lane = threadIdx.x % 32;
  for (long long row = (long long)blockIdx.x * 8 + threadIdx.x / 32; row < rows; row += (long long)gridDim.x * 8) {
    float dot = 0.f;
    for (int j = lane; j < L; j += 32) dot = fmaf(to_f(p[row * L + j]), dp[row * L + j], dot);
    dot = warp_sum(dot);
    for (int j = lane; j < L; j += 32)
      ds[row * L + j] = from_f<T>(scale * to_f(p[row * L + j]) * (dp[row * L + j] - dot));
  }
}

// ---------------------------------------------------------------- embeddings
__global__ void timestep_embedding_kernel(const float* labels, const float* freqs, float* out, int B, int dim) {
  // models/layers.py:515-529: emb = t[:,None] * freqs[None,:]; [sin | cos].  freqs = exp(-log(max_pos)/(half-1) * i)
  // is computed once on the host with the reference's own expression so that the arguments are bit-identical.
  int half = dim / 2;
  GRID_STRIDE(i, (long long)B * half) {
    int b = (int)(i / half), j = (int)(i % half);
    float arg = labels[b] * freqs[j];
    out[(long long)b * dim + j] = sinf(arg);
    out[(long long)b * dim + half + j] = cosf(arg);
  }
}
__global__ void fourier_embedding_kernel(const float* sigma, const float* W, float* out, int B, int nW) {
  // models/layerspp.py:52-54 on x = log(sigma): x[:,None]*W[None,:]*2*pi -> [sin | cos]
  GRID_STRIDE(i, (long long)B * nW) {
    int b = (int)(i / nW), j = (int)(i % nW);
    float proj = logf(sigma[b]) * W[j] * 2.f * 3.14159265358979323846f;
    out[(long long)b * 2 * nW + j] = sinf(proj);
    out[(long long)b * 2 * nW + nW + j] = cosf(proj);
  }
}

// ---------------------------------------------------------------- layout changes at the network boundary
template <typename T>
__global__ void nchw_to_nhwc_kernel(const float* __restrict__ x, T* __restrict__ y, long long total, int C, int HW,
                                    int Cpad, float alpha, float beta) {
  GRID_STRIDE(i, total) {           // i indexes the (channel-padded) NHWC output
    int c = (int)(i % Cpad);
    long long t = i / Cpad;
    int p = (int)(t % HW);
    long long n = t / HW;
    y[i] = from_f<T>(c < C ? alpha * x[(n * C + c) * HW + p] + beta : 0.f);
  }
}
// 8 output channels (one 16-byte bf16 / 32-byte fp32 store) per thread, 32-bit index arithmetic
template <typename T>
__global__ void __launch_bounds__(256) nchw_to_nhwc8_kernel(const float* __restrict__ x, T* __restrict__ y, unsigned total, int C,
                                                            int HW, int Cpad, float alpha, float beta) {
  const unsigned V = (unsigned)Cpad / 8;
  for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const unsigned v = i % V, t = i / V;
    const unsigned p = t % (unsigned)HW, n = t / (unsigned)HW;
    float val[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const int c = (int)v * 8 + k;
      val[k] = c < C ? fmaf(alpha, x[((size_t)n * C + c) * HW + p], beta) : 0.f;
    }
    st8(y + (size_t)i * 8, val);
  }
}
template <typename T>
__global__ void nhwc_to_nchw_kernel(const T* __restrict__ x, float* __restrict__ y, long long total, int C, int HW,
                                    int Cpad, const float* row_scale) {
  GRID_STRIDE(i, total) {           // i indexes NCHW output
    int p = (int)(i % HW);
    long long t = i / HW;
    int c = (int)(t % C);
    long long n = t / C;
    float v = to_f(x[(n * HW + p) * Cpad + c]);
    y[i] = row_scale ? v * row_scale[n] : v;
  }
}

// out[pixel][Kpad] = im2col of x (NHWC, small C), zero padded; k = tap*C + c
template <typename T>
__global__ void im2col_small_kernel(const T* __restrict__ x, bf16* __restrict__ out, long long total, int H, int W, int C,
                                    int kh, int kw, int Kpad) {
  GRID_STRIDE(i, total) {
    int k = (int)(i % Kpad);
    long long pix = i / Kpad;
    float v = 0.f;
    if (k < kh * kw * C) {
      int tap = k / C, c = k % C;
      int xx = (int)(pix % W) + tap % kw - (kw - 1) / 2;
      long long t = pix / W;
      int yy = (int)(t % H) + tap / kw - (kh - 1) / 2;
      long long n = t / H;
      if (xx >= 0 && xx < W && yy >= 0 && yy < H) v = to_f(x[((n * H + yy) * W + xx) * C + c]);
    }
    out[i] = __float2bfloat16_rn(v);
  }
}


// ---------------------------------------------------------------- strided im2col / col2im (pyramid down-convs)
// cols[(n,oy,ox)][tap][c] = x[n][oy*s + r - pad][ox*s + q - pad][c]  (0 outside), vectorised over channel quads
template <typename T>
__global__ void im2col_kernel(const T* __restrict__ x, T* __restrict__ cols, long long total_quads, int H, int W, int C,
                              int kh, int kw, int stride, int pad, int OH, int OW) {
  const int Q = C / 4, taps = kh * kw;
  GRID_STRIDE(gq, total_quads) {
    int quad = (int)(gq % Q);
    long long t = gq / Q;
    int tap = (int)(t % taps);
    t /= taps;
    int ox = (int)(t % OW);
    t /= OW;
    int oy = (int)(t % OH);
    long long n = t / OH;
    int y = oy * stride + tap / kw - pad, xx = ox * stride + tap % kw - pad;
    float v[4] = {0.f, 0.f, 0.f, 0.f};
    if (y >= 0 && y < H && xx >= 0 && xx < W) load4(x + ((n * H + y) * W + xx) * C + quad * 4, v);
    store4(cols + gq * 4, v);
  }
}
// dx[n][y][x][c] = sum over (oy,ox,tap) that read (y,x): dcols[(n,oy,ox)][tap][c]
template <typename T>
__global__ void col2im_kernel(const T* __restrict__ dcols, T* __restrict__ dx, long long total_quads, int H, int W, int C,
                              int kh, int kw, int stride, int pad, int OH, int OW) {
  const int Q = C / 4, taps = kh * kw;
  GRID_STRIDE(gq, total_quads) {
    int quad = (int)(gq % Q);
    long long t = gq / Q;
    int xx = (int)(t % W);
    t /= W;
    int y = (int)(t % H);
    long long n = t / H;
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
    for (int r = 0; r < kh; ++r) {
      int ty = y + pad - r;
      if (ty < 0 || ty % stride) continue;
      int oy = ty / stride;
      if (oy >= OH) continue;
      for (int q = 0; q < kw; ++q) {
        int tx = xx + pad - q;
        if (tx < 0 || tx % stride) continue;
        int ox = tx / stride;
        if (ox >= OW) continue;
        float v[4];
        load4(dcols + ((((n * OH + oy) * OW + ox) * taps + r * kw + q) * (long long)C) + quad * 4, v);
#pragma unroll
        for (int k = 0; k < 4; ++k) acc[k] += v[k];
      }
    }
    store4(dx + gq * 4, acc);
  }
}

// ---------------------------------------------------------------- fused_bias_act (reference op)
template <typename T>
__global__ void fused_bias_act_kernel(const T* __restrict__ x, const T* __restrict__ b, const T* __restrict__ ref,
                                      T* __restrict__ y, long long n, int size_b, int step_b, int act, int grad,
                                      float alpha, float scale) {
  GRID_STRIDE(i, n) {
    float v = to_f(x[i]);
    if (b) v += to_f(b[(i / step_b) % size_b]);
    float r = ref ? to_f(ref[i]) : 0.f;
    float o;
    if (act == 3) {
      if (grad == 0) o = v > 0.f ? v : v * alpha;
      else if (grad == 1) o = r > 0.f ? v : v * alpha;
      else o = 0.f;
    } else {
      o = grad < 2 ? v : 0.f;
    }
    y[i] = from_f<T>(o * scale);
  }
}

// ---------------------------------------------------------------- loss head
__global__ void dsm_perturb_kernel(const float* __restrict__ x0, const float* __restrict__ z,
                                   const float* __restrict__ mc, const float* __restrict__ sd, float* __restrict__ xt,
                                   long long total, long long D) {
  GRID_STRIDE(i, total) {
    long long n = i / D;
    xt[i] = fmaf(sd[n], z[i], mc[n] * x0[i]);
  }
}
// grid (parts, B): a thread-block cluster of `parts` CTAs per sample (1 for the 32x32 / 64x64 images, 8 for 256x256 where
// 16 single blocks would run alone on the GPU); the slices' sums meet in CTA 0 through distributed shared memory, in
// rank order (deterministic)
__global__ void __launch_bounds__(256) dsm_loss_kernel(const float* __restrict__ out, const float* __restrict__ z,
                                                       const float* __restrict__ a, const float* __restrict__ b,
                                                       const float* __restrict__ w, float* loss, float* dout,
                                                       const float* __restrict__ gvec, long long D, int reduce_mean) {
  const long long n = blockIdx.y;
  const int parts = gridDim.x, rank = blockIdx.x;
  const float an = a[n], bn = b[n], wn = w[n];
  const float gscale = (dout && gvec) ? gvec[n] : 1.f;
  const float red = reduce_mean ? 1.f / (float)D : 0.5f;
  const long long per = (D + parts - 1) / parts, i0 = rank * per, i1 = i0 + per < D ? i0 + per : D;
  float acc = 0.f;
  for (long long i = i0 + threadIdx.x; i < i1; i += 256) {
    float e = fmaf(an, out[n * D + i], bn * z[n * D + i]);
    acc = fmaf(e, e, acc);
    if (dout) dout[n * D + i] = gscale * wn * red * 2.f * e * an;
  }
  acc = warp_sum(acc);
  __shared__ float sm[8];
  __shared__ float s_tot;
  if (threadIdx.x % 32 == 0) sm[threadIdx.x / 32] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.;
    for (int i = 0; i < 8; ++i) t += (double)sm[i];
    s_tot = (float)t;
  }
  if (parts > 1) {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
  } else {
    __syncthreads();
  }
  if (threadIdx.x == 0 && rank == 0) {
    double t = 0.;
    for (int r = 0; r < parts; ++r) {
      float v;
      if (parts > 1) {
        uint32_t la = (uint32_t)__cvta_generic_to_shared(&s_tot), ra;
        asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(ra) : "r"(la), "r"((uint32_t)r));
        asm volatile("ld.shared::cluster.f32 %0, [%1];" : "=f"(v) : "r"(ra) : "memory");
      } else {
        v = s_tot;
      }
      t += (double)v;
    }
    loss[n] = wn * red * (float)t;
  }
  // no CTA may exit while CTA 0 can still read its s_tot
  if (parts > 1) {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
  }
}

// ---------------------------------------------------------------- optimizer
__global__ void __launch_bounds__(256) sumsq_kernel(const float* __restrict__ x, long long n4, long long n, float* acc) {
  float s = 0.f;
  GRID_STRIDE(i, n4) {
    float4 v = reinterpret_cast<const float4*>(x)[i];
    s = fmaf(v.x, v.x, s); s = fmaf(v.y, v.y, s); s = fmaf(v.z, v.z, s); s = fmaf(v.w, v.w, s);
  }
  if (blockIdx.x == 0 && threadIdx.x < (n & 3)) { float v = x[(n & ~3LL) + threadIdx.x]; s = fmaf(v, v, s); }
  s = warp_sum(s);
  __shared__ float sm[8];
  if (threadIdx.x % 32 == 0) sm[threadIdx.x / 32] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int i = 0; i < 8; ++i) t += sm[i];
    atomicAdd(acc, t);
  }
}

__global__ void __launch_bounds__(256) adam_ema_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                                                       float* __restrict__ v, float* __restrict__ ema,
                                                       const uint8_t* __restrict__ ema_mask, bf16* __restrict__ p16,
                                                       long long n, long long n4, const float* gnorm_sq, float clip, float lr, float b1,
                                                       float b2, float eps, float wd, float bc1, float bc2, float decay,
                                                       const float* __restrict__ dyn) {
  // step-dependent scalars from device memory (a captured CUDA graph replays with fresh values): {lr, bc1, bc2, decay}
  if (dyn) { lr = dyn[0]; bc1 = dyn[1]; bc2 = dyn[2]; decay = dyn[3]; }
  float coef = 1.f;
  if (gnorm_sq && clip >= 0.f) coef = fminf(1.f, clip / (sqrtf(*gnorm_sq) + 1e-6f));
  const float step_size = lr / bc1, sq_bc2 = sqrtf(bc2);
  auto update = [&](float& pi, float gi, float& mi, float& vi) {
    gi *= coef;
    if (wd != 0.f) gi = fmaf(wd, pi, gi);
    mi = fmaf(b1, mi, (1.f - b1) * gi);
    vi = fmaf(b2, vi, (1.f - b2) * gi * gi);
    const float denom = sqrtf(vi) / sq_bc2 + eps;
    pi -= step_size * (mi / denom);
  };
  // four elements per thread (all streams 16-byte aligned: checked by the caller through n4)
  GRID_STRIDE(i4, n4) {
    float4 P = reinterpret_cast<float4*>(p)[i4], M = reinterpret_cast<float4*>(m)[i4], V = reinterpret_cast<float4*>(v)[i4];
    const float4 G = reinterpret_cast<const float4*>(g)[i4];
    update(P.x, G.x, M.x, V.x);
    update(P.y, G.y, M.y, V.y);
    update(P.z, G.z, M.z, V.z);
    update(P.w, G.w, M.w, V.w);
    reinterpret_cast<float4*>(m)[i4] = M;
    reinterpret_cast<float4*>(v)[i4] = V;
    reinterpret_cast<float4*>(p)[i4] = P;
    if (ema) {
      float4 E = reinterpret_cast<float4*>(ema)[i4];
      uchar4 k = make_uchar4(1, 1, 1, 1);
      if (ema_mask) k = reinterpret_cast<const uchar4*>(ema_mask)[i4];
      if (k.x) E.x = E.x - (1.f - decay) * (E.x - P.x);
      if (k.y) E.y = E.y - (1.f - decay) * (E.y - P.y);
      if (k.z) E.z = E.z - (1.f - decay) * (E.z - P.z);
      if (k.w) E.w = E.w - (1.f - decay) * (E.w - P.w);
      reinterpret_cast<float4*>(ema)[i4] = E;
    }
    if (p16) {
      __nv_bfloat162 lo = __floats2bfloat162_rn(P.x, P.y), hi = __floats2bfloat162_rn(P.z, P.w);
      uint2 t;
      t.x = *reinterpret_cast<uint32_t*>(&lo);
      t.y = *reinterpret_cast<uint32_t*>(&hi);
      reinterpret_cast<uint2*>(p16)[i4] = t;
    }
  }
  for (long long i = n4 * 4 + (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    float pi = p[i], mi = m[i], vi = v[i];
    update(pi, g[i], mi, vi);
    m[i] = mi;
    v[i] = vi;
    p[i] = pi;
    if (ema && (!ema_mask || ema_mask[i])) {
      float e = ema[i];
      ema[i] = e - (1.f - decay) * (e - pi);
    }
    if (p16) p16[i] = __float2bfloat16_rn(pi);
  }
}

// ---------------------------------------------------------------- sampler
__global__ void pc_update_kernel(const float* __restrict__ x, const float* __restrict__ s, const float* __restrict__ noise,
                                 const float* __restrict__ ca, const float* __restrict__ cb, const float* __restrict__ cc,
                                 float* __restrict__ x_mean, float* __restrict__ x_new, long long total, long long D) {
  GRID_STRIDE(i, total) {
    long long n = i / D;
    float xm = fmaf(cb[n], s[i], ca[n] * x[i]);
    if (x_mean) x_mean[i] = xm;
    x_new[i] = noise ? fmaf(cc[n], noise[i], xm) : xm;
  }
}
// grid B blocks: per-sample L2 norms, then atomically accumulate the batch mean
__global__ void __launch_bounds__(256) batch_norms_kernel(const float* __restrict__ a, const float* __restrict__ b,
                                                          float* out, int B, long long D) {
  const long long n = blockIdx.x;
  float sa = 0.f, sb = 0.f;
  for (long long i = threadIdx.x; i < D; i += 256) {
    float u = a[n * D + i];
    sa = fmaf(u, u, sa);
    if (b) { float w = b[n * D + i]; sb = fmaf(w, w, sb); }
  }
  sa = warp_sum(sa);
  sb = warp_sum(sb);
  __shared__ float sm[16];
  if (threadIdx.x % 32 == 0) { sm[threadIdx.x / 32] = sa; sm[8 + threadIdx.x / 32] = sb; }
  __syncthreads();
  if (threadIdx.x == 0) {
    float ta = 0.f, tb = 0.f;
    for (int i = 0; i < 8; ++i) { ta += sm[i]; tb += sm[8 + i]; }
    atomicAdd(out + 0, sqrtf(ta) / (float)B);
    if (b) atomicAdd(out + 1, sqrtf(tb) / (float)B);
  }
}
__global__ void langevin_coeffs_kernel(const float* norms, const float* alpha, float snr, float* ca, float* cb, float* cc, int B) {
  int n = blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= B) return;
  float r = snr * norms[1] / norms[0];
  float step = r * r * 2.f * alpha[n];
  ca[n] = 1.f;
  cb[n] = step;
  cc[n] = sqrtf(step * 2.f);
}

}  // namespace

#define S ((cudaStream_t)stream)

extern "C" __attribute__((visibility("default"))) int st_cast(const void* src, int sdt, void* dst, int ddt, int64_t n, void* stream) {
  long long n4 = n / 4;
  int g = grid1d(n4, 256 * 2);
  if (sdt == ST_F32 && ddt == ST_BF16) cast_kernel<float, bf16><<<g, 256, 0, S>>>((const float*)src, (bf16*)dst, n4, n);
  else if (sdt == ST_BF16 && ddt == ST_F32) cast_kernel<bf16, float><<<g, 256, 0, S>>>((const bf16*)src, (float*)dst, n4, n);
  else if (sdt == ST_F32 && ddt == ST_F32) cast_kernel<float, float><<<g, 256, 0, S>>>((const float*)src, (float*)dst, n4, n);
  else if (sdt == ST_BF16 && ddt == ST_BF16) cast_kernel<bf16, bf16><<<g, 256, 0, S>>>((const bf16*)src, (bf16*)dst, n4, n);
  else { st_set_error("st_cast: bad dtypes"); return ST_ERR_ARG; }
  ST_CHECK_LAUNCH("st_cast");
  return 0;
}

extern "C" __attribute__((visibility("default"))) int st_axpby(const void* a, const void* b, void* out, int dtype, float alpha, float beta, int64_t n, void* stream) {
  long long n4 = n / 4;
  ST_DISPATCH_DTYPE(dtype, T, (axpby_kernel<T><<<grid1d(n4, 256 * 2), 256, 0, S>>>((const T*)a, (const T*)b, (T*)out, alpha, beta, n4, n)));
  ST_CHECK_LAUNCH("st_axpby");
  return 0;
}

extern "C" __attribute__((visibility("default"))) int st_silu(const void* x, void* y, int dtype, int64_t n, void* stream) {
  ST_DISPATCH_DTYPE(dtype, T, (silu_kernel<T><<<grid1d(n, 256 * 4), 256, 0, S>>>((const T*)x, (T*)y, n)));
  ST_CHECK_LAUNCH("st_silu");
  return 0;
}
extern "C" __attribute__((visibility("default"))) int st_silu_bwd(const void* x, const void* dy, void* dx, int dtype, int64_t n, void* stream) {
  ST_DISPATCH_DTYPE(dtype, T, (silu_bwd_kernel<T><<<grid1d(n, 256 * 4), 256, 0, S>>>((const T*)x, (const T*)dy, (T*)dx, n)));
  ST_CHECK_LAUNCH("st_silu_bwd");
  return 0;
}

extern "C" __attribute__((visibility("default"))) int st_resample2x(const void* x1, const void* x2, void* y, int dtype, int n_img, int H, int W, int C1, int C2,
                             int dir, float scale, void* stream) {
  ST_CHECK_ARG(C1 % 4 == 0 && C2 % 4 == 0, "st_resample2x: channels must be multiples of 4");
  ST_CHECK_ARG(dir == 1 || (dir == -1 && H % 2 == 0 && W % 2 == 0), "st_resample2x: bad dir/size");
  int Q = (C1 + C2) / 4;
  const bool vec8 = C1 % 8 == 0 && C2 % 8 == 0 && (long long)n_img * H * W * ((C1 + C2) / 8) < (1LL << 31) &&
                    ((reinterpret_cast<uintptr_t>(x1) | reinterpret_cast<uintptr_t>(x2) | reinterpret_cast<uintptr_t>(y)) & 31) == 0;
  if (vec8 && dir == 1) {
    const unsigned total = (unsigned)((long long)n_img * H * W * ((C1 + C2) / 8));
    ST_DISPATCH_DTYPE(dtype, T, (up2v_kernel<T><<<grid1d(total, 256), 256, 0, S>>>((const T*)x1, (const T*)x2, (T*)y, total, H, W, C1, C2, scale)));
  } else if (vec8) {
    const unsigned total = (unsigned)((long long)n_img * (H / 2) * (W / 2) * ((C1 + C2) / 8));
    ST_DISPATCH_DTYPE(dtype, T, (down2v_kernel<T><<<grid1d(total, 256), 256, 0, S>>>((const T*)x1, (const T*)x2, (T*)y, total, H, W, C1, C2, scale)));
  } else if (dir == 1) {
    long long total = (long long)n_img * (2 * H) * (2 * W) * Q;
    ST_DISPATCH_DTYPE(dtype, T, (up2_kernel<T><<<grid1d(total, 256 * 2), 256, 0, S>>>((const T*)x1, (const T*)x2, (T*)y, total, H, W, C1, C2, scale)));
  } else {
    long long total = (long long)n_img * (H / 2) * (W / 2) * Q;
    ST_DISPATCH_DTYPE(dtype, T, (down2_kernel<T><<<grid1d(total, 256 * 2), 256, 0, S>>>((const T*)x1, (const T*)x2, (T*)y, total, H, W, C1, C2, scale)));
  }
  ST_CHECK_LAUNCH("st_resample2x");
  return 0;
}

extern "C" __attribute__((visibility("default"))) int st_colsum(const void* x, int dtype, int64_t groups, int64_t rows_per_group, int C, int64_t ld, float scale,
                         float* out, int accumulate, void* stream) {
  ST_CHECK_ARG(C % 4 == 0 && ld % 4 == 0 && ld >= C, "st_colsum: C and ld must be multiples of 4, ld >= C");
  ST_CHECK_ARG(groups >= 1 && groups < (1LL << 31), "st_colsum: bad groups");
  dim3 grid((unsigned)groups, (C + 127) / 128);
  ST_DISPATCH_DTYPE(dtype, T, (st_launch(colsum_kernel<T>, grid, dim3(1024), 0, S, (const T*)x, rows_per_group, C, ld, scale, out, accumulate)));
  ST_CHECK_LAUNCH("st_colsum");
  return 0;
}

extern "C" __attribute__((visibility("default"))) int st_transpose_conv_weights(const void* src, void* dst, int dtype, const int64_t* table,
                                          const int64_t* tile_prefix, int n_entries, int64_t total_tiles, void* stream) {
  ST_CHECK_ARG(n_entries >= 0 && total_tiles >= 0 && total_tiles < (1LL << 31), "st_transpose_conv_weights: bad table");
  if (n_entries == 0 || total_tiles == 0) return 0;
  ST_DISPATCH_DTYPE(dtype, T, (transpose_conv_weights_kernel<T><<<(unsigned)total_tiles, 256, 0, S>>>(
                                  (const T*)src, (T*)dst, (const long long*)table, (const long long*)tile_prefix, n_entries)));
  ST_CHECK_LAUNCH("st_transpose_conv_weights");
  return 0;
}

extern "C" __attribute__((visibility("default"))) int st_colsum_batched(const void* jobs, int n_jobs, int blocks_y, void* stream) {
  ST_CHECK_ARG(n_jobs >= 0 && blocks_y >= 1 && blocks_y <= 65535, "st_colsum_batched: bad grid");
  static_assert(sizeof(ColsumJob) == 21 * 8, "ColsumJob is a table of 21 eight-byte fields");
  if (n_jobs == 0) return 0;
  st_launch(colsum_batched_kernel, dim3(n_jobs, blocks_y), dim3(1024), 0, S, (const ColsumJob*)jobs);
  ST_CHECK_LAUNCH("st_colsum_batched");
  return 0;
}

extern "C" __attribute__((visibility("default"))) int st_softmax_fwd(const float* logits, void* p, int dtype, int64_t rows, int L, float scale, void* stream) {
  ST_DISPATCH_DTYPE(dtype, T, (softmax_fwd_kernel<T><<<grid1d(rows, 8), 256, 0, S>>>(logits, (T*)p, rows, L, scale)));
  ST_CHECK_LAUNCH("st_softmax_fwd");
  return 0;
}
extern "C" __attribute__((visibility("default"))) int st_softmax_bwd(const void* p, const float* dp, void* ds, int dtype, int64_t rows, int L, float scale, void* stream) {
  ST_DISPATCH_DTYPE(dtype, T, (softmax_bwd_kernel<T><<<grid1d(rows, 8), 256, 0, S>>>((const T*)p, dp, (T*)ds, rows, L, scale)));
  ST_CHECK_LAUNCH("st_softmax_bwd");
  return 0;
}

extern "C" __attribute__((visibility("default"))) int st_timestep_embedding(const float* labels, const float* freqs, float* out, int B, int dim, void* stream) {
  ST_CHECK_ARG(dim % 2 == 0 && dim >= 4, "st_timestep_embedding: dim must be even");
  timestep_embedding_kernel<<<grid1d((long long)B * dim / 2, 256), 256, 0, S>>>(labels, freqs, out, B, dim);
  ST_CHECK_LAUNCH("st_timestep_embedding");
  return 0;
}
extern "C" __attribute__((visibility("default"))) int st_fourier_embedding(const float* sigma, const float* W, float* out, int B, int nW, void* stream) {
  fourier_embedding_kernel<<<grid1d((long long)B * nW, 256), 256, 0, S>>>(sigma, W, out, B, nW);
  ST_CHECK_LAUNCH("st_fourier_embedding");
  return 0;
}

extern "C" __attribute__((visibility("default"))) int st_nchw_to_nhwc(const float* x, void* y, int dtype, int n_img, int C, int H, int W, int Cpad, float alpha,
                               float beta, void* stream) {
  ST_CHECK_ARG(Cpad >= C, "st_nchw_to_nhwc: Cpad < C");
  long long total = (long long)n_img * Cpad * H * W;
  if (Cpad % 8 == 0 && total / 8 < (1LL << 31) && (reinterpret_cast<uintptr_t>(y) & 31) == 0) {
    const unsigned tv = (unsigned)(total / 8);
    ST_DISPATCH_DTYPE(dtype, T, (nchw_to_nhwc8_kernel<T><<<grid1d(tv, 256), 256, 0, S>>>(x, (T*)y, tv, C, H * W, Cpad, alpha, beta)));
  } else
  ST_DISPATCH_DTYPE(dtype, T, (nchw_to_nhwc_kernel<T><<<grid1d(total, 256 * 4), 256, 0, S>>>(x, (T*)y, total, C, H * W, Cpad, alpha, beta)));
  ST_CHECK_LAUNCH("st_nchw_to_nhwc");
  return 0;
}
extern "C" __attribute__((visibility("default"))) int st_nhwc_to_nchw(const void* x, int dtype, float* y, int n_img, int C, int H, int W, int Cpad,
                               const float* row_scale, void* stream) {
  ST_CHECK_ARG(Cpad >= C, "st_nhwc_to_nchw: Cpad < C");
  long long total = (long long)n_img * C * H * W;
  ST_DISPATCH_DTYPE(dtype, T, (nhwc_to_nchw_kernel<T><<<grid1d(total, 256 * 4), 256, 0, S>>>((const T*)x, y, total, C, H * W, Cpad, row_scale)));
  ST_CHECK_LAUNCH("st_nhwc_to_nchw");
  return 0;
}

extern "C" __attribute__((visibility("default"))) int st_im2col_small(const void* x, int dtype, void* out, int n_img, int H, int W, int C, int kh, int kw, int Kpad,
                               void* stream) {
  ST_CHECK_ARG(Kpad >= kh * kw * C, "st_im2col_small: Kpad too small");
  long long total = (long long)n_img * H * W * Kpad;
  ST_DISPATCH_DTYPE(dtype, T, (im2col_small_kernel<T><<<grid1d(total, 256 * 4), 256, 0, S>>>((const T*)x, (bf16*)out, total, H, W, C, kh, kw, Kpad)));
  ST_CHECK_LAUNCH("st_im2col_small");
  return 0;
}

extern "C" __attribute__((visibility("default"))) int st_fused_bias_act(const void* x, const void* b, const void* ref, void* y, int dtype, int64_t n, int size_b,
                                 int step_b, int act, int grad, float alpha, float scale, void* stream) {
  ST_CHECK_ARG(!b || (size_b > 0 && step_b > 0), "st_fused_bias_act: bad bias geometry");
  ST_DISPATCH_DTYPE(dtype, T, (fused_bias_act_kernel<T><<<grid1d(n, 256 * 4), 256, 0, S>>>((const T*)x, (const T*)b, (const T*)ref, (T*)y, n, size_b, step_b, act, grad, alpha, scale)));
  ST_CHECK_LAUNCH("st_fused_bias_act");
  return 0;
}

extern "C" __attribute__((visibility("default"))) int st_dsm_perturb(const float* x0, const float* z, const float* mean_coeff, const float* std, float* xt, int B,
                              int64_t D, void* stream) {
  long long total = (long long)B * D;
  dsm_perturb_kernel<<<grid1d(total, 256 * 4), 256, 0, S>>>(x0, z, mean_coeff, std, xt, total, D);
  ST_CHECK_LAUNCH("st_dsm_perturb");
  return 0;
}
extern "C" __attribute__((visibility("default"))) int st_dsm_loss(const float* out, const float* z, const float* a, const float* b, const float* w, float* loss,
                           float* dout, const float* gvec, int B, int64_t D, int reduce_mean, void* stream) {
  ST_CHECK_ARG(B > 0 && B <= 65535 && D > 0, "st_dsm_loss: bad shape");
  int parts = 1;
  if (D >= 32768) { parts = (int)(D / 16384); if (parts > 8) parts = 8; }
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3(parts, B);
  cfg.blockDim = dim3(256);
  cfg.stream = S;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = parts;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  cudaError_t e = cudaLaunchKernelEx(&cfg, dsm_loss_kernel, out, z, a, b, w, loss, dout, gvec, (long long)D, reduce_mean);
  if (e != cudaSuccess) { st_set_error("st_dsm_loss: launch: %s", cudaGetErrorString(e)); return ST_ERR_CUDA; }
  ST_CHECK_LAUNCH("st_dsm_loss");
  return 0;
}

extern "C" __attribute__((visibility("default"))) int st_sumsq(const float* x, int64_t n, float* acc, void* stream) {
  long long n4 = n / 4;
  int g = grid1d(n4, 256 * 8);
  sumsq_kernel<<<g, 256, 0, S>>>(x, n4, n, acc);
  ST_CHECK_LAUNCH("st_sumsq");
  return 0;
}
extern "C" __attribute__((visibility("default"))) int st_adam_ema(float* p, const float* grad, float* m, float* v, float* ema, const uint8_t* ema_mask, void* p16,
                           int64_t n, const float* gnorm_sq, float clip, float lr, float b1, float b2, float eps, float wd,
                           float bc1, float bc2, float ema_decay, const float* dyn, void* stream) {
  const uintptr_t al = reinterpret_cast<uintptr_t>(p) | reinterpret_cast<uintptr_t>(grad) | reinterpret_cast<uintptr_t>(m) |
                       reinterpret_cast<uintptr_t>(v) | reinterpret_cast<uintptr_t>(ema) | reinterpret_cast<uintptr_t>(p16);
  const long long n4 = ((al & 15) == 0 && (reinterpret_cast<uintptr_t>(ema_mask) & 3) == 0) ? n / 4 : 0;
  adam_ema_kernel<<<grid1d(n4 > 0 ? n4 : n, 256 * 2), 256, 0, S>>>(p, grad, m, v, ema, ema_mask, (bf16*)p16, n, n4, gnorm_sq, clip, lr,
                                                                    b1, b2, eps, wd, bc1, bc2, ema_decay, dyn);
  ST_CHECK_LAUNCH("st_adam_ema");
  return 0;
}

extern "C" __attribute__((visibility("default"))) int st_pc_update(const float* x, const float* s, const float* noise, const float* ca, const float* cb,
                            const float* cc, float* x_mean, float* x_new, int B, int64_t D, void* stream) {
  long long total = (long long)B * D;
  pc_update_kernel<<<grid1d(total, 256 * 4), 256, 0, S>>>(x, s, noise, ca, cb, cc, x_mean, x_new, total, D);
  ST_CHECK_LAUNCH("st_pc_update");
  return 0;
}
extern "C" __attribute__((visibility("default"))) int st_batch_norms(const float* a, const float* b, float* out, int B, int64_t D, void* stream) {
  cudaMemsetAsync(out, 0, 2 * sizeof(float), S);
  batch_norms_kernel<<<B, 256, 0, S>>>(a, b, out, B, D);
  ST_CHECK_LAUNCH("st_batch_norms");
  return 0;
}
extern "C" __attribute__((visibility("default"))) int st_langevin_coeffs(const float* norms, const float* alpha, float snr, float* ca, float* cb, float* cc, int B,
                                  void* stream) {
  langevin_coeffs_kernel<<<(B + 127) / 128, 128, 0, S>>>(norms, alpha, snr, ca, cb, cc, B);
  ST_CHECK_LAUNCH("st_langevin_coeffs");
  return 0;
}

extern "C" __attribute__((visibility("default"))) int st_im2col(const void* x, void* cols, int dtype, int n_img, int H, int W, int C, int kh, int kw, int stride,
                             int pad, int OH, int OW, void* stream) {
  ST_CHECK_ARG(C % 4 == 0 && stride > 0 && OH > 0 && OW > 0, "st_im2col: bad geometry");
  long long total = (long long)n_img * OH * OW * kh * kw * (C / 4);
  ST_DISPATCH_DTYPE(dtype, T, (im2col_kernel<T><<<grid1d(total, 256 * 2), 256, 0, S>>>((const T*)x, (T*)cols, total, H, W, C, kh, kw, stride, pad, OH, OW)));
  ST_CHECK_LAUNCH("st_im2col");
  return 0;
}
extern "C" __attribute__((visibility("default"))) int st_col2im(const void* dcols, void* dx, int dtype, int n_img, int H, int W, int C, int kh, int kw, int stride,
                             int pad, int OH, int OW, void* stream) {
  ST_CHECK_ARG(C % 4 == 0 && stride > 0 && OH > 0 && OW > 0, "st_col2im: bad geometry");
  long long total = (long long)n_img * H * W * (C / 4);
  ST_DISPATCH_DTYPE(dtype, T, (col2im_kernel<T><<<grid1d(total, 256 * 2), 256, 0, S>>>((const T*)dcols, (T*)dx, total, H, W, C, kh, kw, stride, pad, OH, OW)));
  ST_CHECK_LAUNCH("st_col2im");
  return 0;
}

// ---------------------------------------------------------------- training-batch preparation
// uint8 NHWC images -> fp32 NCHW network input in one pass: x = v/255 (tf.image.convert_image_dtype, datasets.py:117,
// 315-326), optional left-right flip per image (datasets.py:311,322), uniform dequantisation (255 x + u)/256
// (run_lib.py:73-74) and the data scaler a*x + b (datasets.py:56-62).  u comes from `u` (fp32 NCHW, injected) or from
// the counter-based generator keyed by (seed, output element).  One thread per output element; reads are 1 byte and
// strided by C (the whole uint8 batch is 1.5 MB at B=512 and sits in L2), writes are coalesced fp32.
__global__ void prep_batch_kernel(const uint8_t* __restrict__ src, const float* __restrict__ u,
                                  const uint8_t* __restrict__ flip, float* __restrict__ dst, long long total, int C, int H,
                                  int W, int dequant, uint64_t seed, float a, float b) {
  GRID_STRIDE(i, total) {
    const int x = (int)(i % W);
    long long t = i / W;
    const int y = (int)(t % H);
    t /= H;
    const int c = (int)(t % C);
    const long long n = t / C;
    const int sx = (flip && flip[n]) ? W - 1 - x : x;
    float v = (float)src[((n * H + y) * W + sx) * C + c] * (1.f / 255.f);
    if (dequant) {
      float r;
      if (u) {
        r = u[i];
      } else {
        const uint4 w4 = philox4(seed, (uint64_t)(i >> 2));
        const uint32_t w = (i & 3) == 0 ? w4.x : (i & 3) == 1 ? w4.y : (i & 3) == 2 ? w4.z : w4.w;
        r = (float)(w >> 8) * (1.f / 16777216.f);                 // [0, 1)
      }
      v = __fadd_rn(__fmul_rn(255.f, v), r) * (1.f / 256.f);      // unfused, as the reference's separate torch ops
    }
    dst[i] = __fadd_rn(__fmul_rn(a, v), b);
  }
}

extern "C" __attribute__((visibility("default"))) int st_prep_batch(const uint8_t* src, const float* u, const uint8_t* flip, float* dst, int n_img, int C,
                               int H, int W, int dequant, uint64_t seed, float a, float b, void* stream) {
  ST_CHECK_ARG(n_img >= 0 && C > 0 && H > 0 && W > 0, "st_prep_batch: bad shape");
  ST_CHECK_ARG(dequant == 0 || dequant == 1, "st_prep_batch: dequant must be 0 (none) or 1 (uniform)");
  const long long total = (long long)n_img * C * H * W;
  if (total == 0) return 0;
  prep_batch_kernel<<<grid1d(total, 256 * 4), 256, 0, S>>>(src, u, flip, dst, total, C, H, W, dequant, seed, a, b);
  ST_CHECK_LAUNCH("st_prep_batch");
  return 0;
}
