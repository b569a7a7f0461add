// GroupNorm (+SiLU, +dropout) forward and backward on NHWC tensors whose channel axis may be the
// concatenation of two tensors (the U-Net skip concatenation is never materialised).
// Reference: nn.GroupNorm(min(C//4,32), C, eps=1e-6) -> act -> Dropout, models/layerspp.py:232,244-245,
// 258,275-278; models/ncsnpp.py:219-253.
// All kernels are HBM-bound: 8/16-byte vector loads, fp32 math, deterministic reductions.
#include "common.cuh"

namespace {

template <typename T>
struct Src2 {
  const T* x1;
  const T* x2;
  int C1, C2;
  // pointer to channel c0 (multiple of 4) of pixel row `row`
  __device__ __forceinline__ const T* at(long long row, int c0) const {
    return c0 < C1 ? x1 + row * C1 + c0 : x2 + row * C2 + (c0 - C1);
  }
};

// ---------------------------------------------------------------- stats
// grid (n_img, splits), block 256.  Thread -> fixed channel quad, strided over pixels.
template <typename T>
__global__ void __launch_bounds__(256) gn_stats_kernel(Src2<T> s, int hw, int G, int splits, float* part) {
  const int Ct = s.C1 + s.C2, Q = Ct / 4, cpg = Ct / G;
  const int n = blockIdx.x, sp = blockIdx.y;
  const int ppb = 256 / Q;                 // pixels processed in parallel
  const int quad = threadIdx.x % Q, lane = threadIdx.x / Q;
  const int per = (hw + splits - 1) / splits;
  const int p0 = sp * per, p1 = min(hw, p0 + per);
  float sum = 0.f, sq = 0.f;
  if (lane < ppb) {
    for (int p = p0 + lane; p < p1; p += ppb) {
      float v[4];
      load4(s.at((long long)n * hw + p, quad * 4), v);
#pragma unroll
      for (int i = 0; i < 4; ++i) { sum += v[i]; sq = fmaf(v[i], v[i], sq); }
    }
  }
  __shared__ float s_sum[256], s_sq[256];
  s_sum[threadIdx.x] = sum;
  s_sq[threadIdx.x] = sq;
  __syncthreads();
  if (threadIdx.x < G) {
    const int g = threadIdx.x, qpg = cpg / 4;
    double a = 0., b = 0.;
    for (int l = 0; l < ppb; ++l)
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
template <typename T>
__global__ void __launch_bounds__(256) gn_apply_kernel(Src2<T> s, long long total_quads, int hw, int G,
                                                       const float* __restrict__ gamma, const float* __restrict__ beta,
                                                       const float* __restrict__ mean, const float* __restrict__ rstd,
                                                       int act, float p_drop, uint64_t seed, const T* mask, T* y) {
  const int Ct = s.C1 + s.C2, Q = Ct / 4, cpg = Ct / G;
  for (long long gq = (long long)blockIdx.x * blockDim.x + threadIdx.x; gq < total_quads;
       gq += (long long)gridDim.x * blockDim.x) {
    const int quad = (int)(gq % Q);
    const long long row = gq / Q;          // n*hw + pixel
    const int n = (int)(row / hw);
    const int c0 = quad * 4, g = c0 / cpg;
    float v[4], o[4];
    load4(s.at(row, c0), v);
    const float mu = mean[n * G + g], r = rstd[n * G + g];
    float4 ga = *reinterpret_cast<const float4*>(gamma + c0);
    float4 be = *reinterpret_cast<const float4*>(beta + c0);
    const float gam[4] = {ga.x, ga.y, ga.z, ga.w}, bet[4] = {be.x, be.y, be.z, be.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      float u = fmaf((v[i] - mu) * r, gam[i], bet[i]);
      o[i] = act ? silu_f(u) : u;
    }
    if (mask) {
      float mk[4];
      load4(mask + gq * 4, mk);
#pragma unroll
      for (int i = 0; i < 4; ++i) o[i] *= mk[i];
    } else if (p_drop > 0.f) {
      float keep[4];
      dropout4(seed, (uint64_t)gq, p_drop, keep);
#pragma unroll
      for (int i = 0; i < 4; ++i) o[i] *= keep[i];
    }
    store4(y + gq * 4, o);
  }
}

// dz for one quad (shared by both backward passes)
template <typename T>
__device__ __forceinline__ void gn_dz(const float v[4], const float dyv[4], float mu, float r, const float gam[4],
                                      const float bet[4], int act, float p_drop, uint64_t seed, const T* mask,
                                      long long gq, float xhat[4], float dz[4]) {
  float mk[4] = {1.f, 1.f, 1.f, 1.f};
  if (mask) load4(mask + gq * 4, mk);
  else if (p_drop > 0.f) dropout4(seed, (uint64_t)gq, p_drop, mk);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    xhat[i] = (v[i] - mu) * r;
    float d = dyv[i] * mk[i];
    if (act) d *= silu_grad_f(fmaf(xhat[i], gam[i], bet[i]));
    dz[i] = d;
  }
}

// ---------------------------------------------------------------- backward pass 1
// grid (n_img, splits): red[n][split][c][2] = (sum dz, sum dz*xhat) over the split's pixels
template <typename T>
__global__ void __launch_bounds__(256) gn_bwd_reduce_kernel(Src2<T> s, const T* dy, int hw, int G, int splits,
                                                            const float* __restrict__ gamma, const float* __restrict__ beta,
                                                            const float* __restrict__ mean, const float* __restrict__ rstd,
                                                            int act, float p_drop, uint64_t seed, const T* mask, float* red) {
  const int Ct = s.C1 + s.C2, Q = Ct / 4, cpg = Ct / G;
  const int n = blockIdx.x, sp = blockIdx.y;
  const int ppb = 256 / Q;
  const int quad = threadIdx.x % Q, lane = threadIdx.x / Q;
  const int per = (hw + splits - 1) / splits;
  const int p0 = sp * per, p1 = min(hw, p0 + per);
  const int c0 = quad * 4, g = c0 / cpg;
  float a[4] = {0.f, 0.f, 0.f, 0.f}, b[4] = {0.f, 0.f, 0.f, 0.f};
  if (lane < ppb) {
    const float mu = mean[n * G + g], r = rstd[n * G + g];
    float4 ga = *reinterpret_cast<const float4*>(gamma + c0);
    float4 be = *reinterpret_cast<const float4*>(beta + c0);
    const float gam[4] = {ga.x, ga.y, ga.z, ga.w}, bet[4] = {be.x, be.y, be.z, be.w};
    for (int p = p0 + lane; p < p1; p += ppb) {
      const long long row = (long long)n * hw + p;
      const long long gq = row * Q + quad;
      float v[4], dyv[4], xhat[4], dz[4];
      load4(s.at(row, c0), v);
      load4(dy + gq * 4, dyv);
      gn_dz<T>(v, dyv, mu, r, gam, bet, act, p_drop, seed, mask, gq, xhat, dz);
#pragma unroll
      for (int i = 0; i < 4; ++i) { a[i] += dz[i]; b[i] = fmaf(dz[i], xhat[i], b[i]); }
    }
  }
  __shared__ float4 s_a[256], s_b[256];
  s_a[threadIdx.x] = make_float4(a[0], a[1], a[2], a[3]);
  s_b[threadIdx.x] = make_float4(b[0], b[1], b[2], b[3]);
  __syncthreads();
  if (threadIdx.x < Q) {
    float4 ta = make_float4(0.f, 0.f, 0.f, 0.f), tb = ta;
    for (int l = 0; l < ppb; ++l) {
      float4 u = s_a[l * Q + threadIdx.x], w = s_b[l * Q + threadIdx.x];
      ta.x += u.x; ta.y += u.y; ta.z += u.z; ta.w += u.w;
      tb.x += w.x; tb.y += w.y; tb.z += w.z; tb.w += w.w;
    }
    float* o = red + (((long long)n * splits + sp) * Ct + threadIdx.x * 4) * 2;
    o[0] = ta.x; o[1] = tb.x; o[2] = ta.y; o[3] = tb.y; o[4] = ta.z; o[5] = tb.z; o[6] = ta.w; o[7] = tb.w;
  }
}

// dgamma[c] += sum_rows red[row][c][1]; dbeta[c] += sum_rows red[row][c][0]
// block = 32 channels x 8 row lanes (coalesced float2 reads), fixed-order reduction
__global__ void __launch_bounds__(256) gn_bwd_params_kernel(const float* __restrict__ red, int rows, int C, float* dgamma,
                                                            float* dbeta) {
  const int c = blockIdx.x * 32 + threadIdx.x % 32;
  const int rl = threadIdx.x / 32;
  float a = 0.f, b = 0.f;
  if (c < C) {
    for (int r = rl; r < rows; r += 8) {
      float2 v = *reinterpret_cast<const float2*>(red + ((long long)r * C + c) * 2);
      a += v.x;
      b += v.y;
    }
  }
  __shared__ float sa[256], sb[256];
  sa[threadIdx.x] = a;
  sb[threadIdx.x] = b;
  __syncthreads();
  if (rl == 0 && c < C) {
    double ta = 0., tb = 0.;
    for (int l = 0; l < 8; ++l) { ta += (double)sa[l * 32 + threadIdx.x]; tb += (double)sb[l * 32 + threadIdx.x]; }
    dbeta[c] += (float)ta;
    dgamma[c] += (float)tb;
  }
}

// ---------------------------------------------------------------- backward pass 2
// grid (n_img, chunks)
template <typename T>
__global__ void __launch_bounds__(256) gn_bwd_apply_kernel(Src2<T> s, const T* dy, int hw, int G, int splits,
                                                           const float* __restrict__ gamma, const float* __restrict__ beta,
                                                           const float* __restrict__ mean, const float* __restrict__ rstd,
                                                           int act, float p_drop, uint64_t seed, const T* mask,
                                                           const float* __restrict__ red, const T* extra, float extra_scale,
                                                           T* dx1, int accum1, T* dx2, int accum2) {
  const int Ct = s.C1 + s.C2, Q = Ct / 4, cpg = Ct / G;
  const int n = blockIdx.x;
  __shared__ float s1[64], s2[64];
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
    s1[g] = (float)(a * inv);
    s2[g] = (float)(b * inv);
  }
  __syncthreads();
  const long long quads_img = (long long)hw * Q;
  for (long long lq = (long long)blockIdx.y * blockDim.x + threadIdx.x; lq < quads_img;
       lq += (long long)gridDim.y * blockDim.x) {
    const int quad = (int)(lq % Q);
    const long long row = (long long)n * hw + lq / Q;
    const long long gq = row * Q + quad;
    const int c0 = quad * 4, g = c0 / cpg;
    const float mu = mean[n * G + g], r = rstd[n * G + g];
    float4 ga = *reinterpret_cast<const float4*>(gamma + c0);
    float4 be = *reinterpret_cast<const float4*>(beta + c0);
    const float gam[4] = {ga.x, ga.y, ga.z, ga.w}, bet[4] = {be.x, be.y, be.z, be.w};
    float v[4], dyv[4], xhat[4], dz[4], o[4];
    load4(s.at(row, c0), v);
    load4(dy + gq * 4, dyv);
    gn_dz<T>(v, dyv, mu, r, gam, bet, act, p_drop, seed, mask, gq, xhat, dz);
#pragma unroll
    for (int i = 0; i < 4; ++i) o[i] = r * (gam[i] * dz[i] - s1[g] - xhat[i] * s2[g]);
    if (extra) {
      float ex[4];
      load4(extra + gq * 4, ex);
#pragma unroll
      for (int i = 0; i < 4; ++i) o[i] = fmaf(extra_scale, ex[i], o[i]);
    }
    T* dst;
    int acc;
    if (c0 < s.C1) { dst = dx1 + row * s.C1 + c0; acc = accum1; }
    else { dst = dx2 + row * s.C2 + (c0 - s.C1); acc = accum2; }
    if (acc) {
      float old[4];
      load4(dst, old);
#pragma unroll
      for (int i = 0; i < 4; ++i) o[i] += old[i];
    }
    store4(dst, o);
  }
}

int check_geom(int C1, int C2, int G) {
  int Ct = C1 + C2;
  ST_CHECK_ARG(C1 > 0 && C2 >= 0 && C1 % 4 == 0 && C2 % 4 == 0, "groupnorm: channel counts must be multiples of 4 (got %d,%d)", C1, C2);
  ST_CHECK_ARG(G > 0 && G <= 64 && Ct % G == 0 && (Ct / G) % 4 == 0, "groupnorm: group size must be a multiple of 4 (C=%d,G=%d)", Ct, G);
  ST_CHECK_ARG(Ct <= 1024, "groupnorm: C > 1024 unsupported");
  return 0;
}

int grid_for(long long work_items, int per_block) {
  long long b = (work_items + per_block - 1) / per_block;
  long long cap = (long long)st_num_sms() * 16;
  return (int)(b < 1 ? 1 : (b > cap ? cap : b));
}

}  // namespace

extern "C" __attribute__((visibility("default"))) int st_gn_stats(const void* x1, const void* x2, int dtype, int n_img, int hw, int C1, int C2, int G,
                           int splits, float* part, void* stream) {
  if (int e = check_geom(C1, C2, G)) return e;
  ST_CHECK_ARG(splits >= 1 && splits <= 65535, "st_gn_stats: bad splits");
  ST_DISPATCH_DTYPE(dtype, T, {
    Src2<T> s{(const T*)x1, (const T*)x2, C1, C2};
    gn_stats_kernel<T><<<dim3(n_img, splits), 256, 0, (cudaStream_t)stream>>>(s, hw, G, splits, part);
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
                           float p_drop, uint64_t seed, const void* mask, void* y, void* stream) {
  if (int e = check_geom(C1, C2, G)) return e;
  long long total = (long long)n_img * hw * ((C1 + C2) / 4);
  ST_DISPATCH_DTYPE(dtype, T, {
    Src2<T> s{(const T*)x1, (const T*)x2, C1, C2};
    gn_apply_kernel<T><<<grid_for(total, 256 * 4), 256, 0, (cudaStream_t)stream>>>(
        s, total, hw, G, gamma, beta, mean, rstd, act, p_drop, seed, (const T*)mask, (T*)y);
  });
  ST_CHECK_LAUNCH("st_gn_apply");
  return 0;
}

extern "C" __attribute__((visibility("default"))) int st_gn_bwd_reduce(const void* x1, const void* x2, const void* dy, int dtype, int n_img, int hw, int C1,
                                int C2, int G, const float* gamma, const float* beta, const float* mean,
                                const float* rstd, int act, float p_drop, uint64_t seed, const void* mask,
                                int splits, float* red, void* stream) {
  if (int e = check_geom(C1, C2, G)) return e;
  ST_DISPATCH_DTYPE(dtype, T, {
    Src2<T> s{(const T*)x1, (const T*)x2, C1, C2};
    gn_bwd_reduce_kernel<T><<<dim3(n_img, splits), 256, 0, (cudaStream_t)stream>>>(
        s, (const T*)dy, hw, G, splits, gamma, beta, mean, rstd, act, p_drop, seed, (const T*)mask, red);
  });
  ST_CHECK_LAUNCH("st_gn_bwd_reduce");
  return 0;
}

extern "C" __attribute__((visibility("default"))) int st_gn_bwd_params(const float* red, int rows, int C, float* dgamma, float* dbeta, void* stream) {
  gn_bwd_params_kernel<<<(C + 31) / 32, 256, 0, (cudaStream_t)stream>>>(red, rows, C, dgamma, dbeta);
  ST_CHECK_LAUNCH("st_gn_bwd_params");
  return 0;
}

extern "C" __attribute__((visibility("default"))) int st_gn_bwd_apply(const void* x1, const void* x2, const void* dy, int dtype, int n_img, int hw, int C1,
                               int C2, int G, const float* gamma, const float* beta, const float* mean,
                               const float* rstd, int act, float p_drop, uint64_t seed, const void* mask,
                               int splits, const float* red, const void* extra, float extra_scale, void* dx1,
                               int accum1, void* dx2, int accum2, void* stream) {
  if (int e = check_geom(C1, C2, G)) return e;
  long long quads_img = (long long)hw * ((C1 + C2) / 4);
  int chunks = (int)((quads_img + 256 * 4 - 1) / (256 * 4));
  int cap = (st_num_sms() * 16 + n_img - 1) / n_img;
  if (chunks > cap) chunks = cap;
  if (chunks < 1) chunks = 1;
  ST_DISPATCH_DTYPE(dtype, T, {
    Src2<T> s{(const T*)x1, (const T*)x2, C1, C2};
    gn_bwd_apply_kernel<T><<<dim3(n_img, chunks), 256, 0, (cudaStream_t)stream>>>(
        s, (const T*)dy, hw, G, splits, gamma, beta, mean, rstd, act, p_drop, seed, (const T*)mask, red,
        (const T*)extra, extra_scale, (T*)dx1, accum1, (T*)dx2, accum2);
  });
  ST_CHECK_LAUNCH("st_gn_bwd_apply");
  return 0;
}
