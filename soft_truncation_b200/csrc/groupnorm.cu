// GroupNorm (+SiLU, +dropout) forward and backward on NHWC tensors whose channel axis may be the
// concatenation of two tensors (the U-Net skip concatenation is never materialised).
// Reference: nn.GroupNorm(min(C//4,32), C, eps=1e-6) -> act -> Dropout, models/layerspp.py:232,244-245,
// 258,275-278; models/ncsnpp.py:219-253.
//
// All kernels are HBM-bound streaming passes.  Thread mapping: one thread owns a fixed run of 8 channels
// (16-byte bf16 / 32-byte fp32 vectors) of one image and walks that image's pixels, so per-channel constants
// (gamma, beta, mean, rstd) sit in registers and the inner loop has no integer division; two pixels are in
// flight per iteration.  Reductions are deterministic (fixed-order shared-memory trees, no float atomics).
#include "groupnorm.cuh"

namespace {

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

// ---------------------------------------------------------------- apply
// grid (chunks, n_img); pipelined stream: x (an injected dropout mask - parity tests only - is read directly)
template <typename T, bool ACT, int DROP>
__global__ void __launch_bounds__(256, 4) gn_apply_kernel(Src2<T> s, int hw, int G, const float* __restrict__ gamma,
                                                          const float* __restrict__ beta, const float* __restrict__ mean,
                                                          const float* __restrict__ rstd, float p_drop, uint64_t seed,
                                                          const T* mask, uint8_t* keepbits, T* y,
                                                          const float* __restrict__ part, int splits, double inv_count,
                                                          float eps, float* mean_out, float* rstd_out,
                                                          const uint64_t* __restrict__ seed_off,
                                                          const float* __restrict__ q1, const float* __restrict__ q2, int qrows,
                                                          int rev) {
  extern __shared__ __align__(16) uint8_t gsm[];
  pdl_wait();
  pdl_trigger();
  if constexpr (DROP == DROP_FAST) { if (seed_off) seed += *seed_off; }
  using P = Pipe<T, 1, GN_DEPTH>;
  const P pipe(gsm);
  const int Ct = s.C1 + s.C2, cpg = Ct / G;
  const int n = img_of(blockIdx.y, gridDim.y, rev);
  const Walk w(Ct, n, hw, blockIdx.x, gridDim.x);
  ChanConst k;
  if (part || q1) {
    // statistics straight from the partial sums of st_gn_stats (same arithmetic as gn_finalize_kernel) or from the
    // per-(qrows rows, 4 channels) sums the producing GEMMs emitted (st_gemm gn_part; one buffer per concatenated
    // source); the first chunk of every image publishes mean / rstd for the backward passes
    __shared__ float s_mean[64], s_rstd[64];
    // The partial sums of one image are summed by the whole block - 256 threads = (row lane, column) over the
    // [rows][columns] partial array, then one thread per group adds its columns - so that the dependent-load chain in
    // front of the streaming loop stays a few loads long even when an image has hundreds of row blocks (256x256 images).
    __shared__ float s_pa[256], s_pb[256];
    const int NQ = q1 ? Ct >> 2 : G;                    // columns of the partial array: quads or groups
    const bool par = NQ <= 256;
    if (par) {
      const int rows = q1 ? hw / qrows : splits, lanes_q = 256 / NQ;
      const int col = threadIdx.x % NQ, ln = threadIdx.x / NQ;
      if (ln < lanes_q) {
        float a = 0.f, b = 0.f;
        const float2* src;
        int ld;
        if (q1) {
          const int c = col << 2;
          const bool first = c < s.C1;
          ld = (first ? s.C1 : s.C2) >> 2;
          src = reinterpret_cast<const float2*>(first ? q1 : q2) + (long long)n * rows * ld + ((first ? c : c - s.C1) >> 2);
        } else {
          ld = G;
          src = reinterpret_cast<const float2*>(part) + (long long)n * rows * G + col;
        }
        for (int r = ln; r < rows; r += lanes_q) {
          const float2 v = src[(long long)r * ld];
          a += v.x;
          b += v.y;
        }
        s_pa[ln * NQ + col] = a;
        s_pb[ln * NQ + col] = b;
      }
      __syncthreads();
    }
    if (threadIdx.x < G) {
      double a = 0., b = 0.;
      if (par) {
        const int lanes_q = 256 / NQ, per = q1 ? cpg >> 2 : 1;
        for (int l = 0; l < lanes_q; ++l)
          for (int k = 0; k < per; ++k) {
            a += (double)s_pa[l * NQ + threadIdx.x * per + k];
            b += (double)s_pb[l * NQ + threadIdx.x * per + k];
          }
      } else if (q1) {
        const int nch = hw / qrows;
        for (int c = threadIdx.x * cpg; c < (threadIdx.x + 1) * cpg; c += 4) {
          const bool first = c < s.C1;
          const float2* q = reinterpret_cast<const float2*>(first ? q1 : q2);
          const int qn = (first ? s.C1 : s.C2) >> 2, cc = (first ? c : c - s.C1) >> 2;
          for (int ch = 0; ch < nch; ++ch) {
            const float2 v = q[((long long)n * nch + ch) * qn + cc];
            a += (double)v.x;
            b += (double)v.y;
          }
        }
      } else
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

template <typename T, bool ACT, int DROP>
__global__ void __launch_bounds__(256, 3) gn_bwd_reduce_kernel(Src2<T> s, const T* dy, int hw, int G, int splits,
                                                            const float* __restrict__ gamma, const float* __restrict__ beta,
                                                            const float* __restrict__ mean, const float* __restrict__ rstd,
                                                            float p_drop, uint64_t seed, const T* mask,
                                                            const uint8_t* __restrict__ keepbits, float* red, int rev) {
  pdl_wait();
  pdl_trigger();
  gn_bwd_reduce_body<T, ACT, DROP>(s, dy, hw, G, splits, gamma, beta, mean, rstd, p_drop, seed, mask, keepbits, red,
                                   img_of(blockIdx.x, gridDim.x, rev), blockIdx.y);
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
                           int64_t count, float eps, const float* q1, const float* q2, int qrows, void* stream) {
  if (int e = check_geom(C1, C2, G)) return e;
  ST_CHECK_ARG(n_img <= 65535, "st_gn_apply: more than 65535 images");
  ST_CHECK_ARG(!part || (splits >= 1 && count > 0 && mean && rstd), "st_gn_apply: partial sums need splits, count and mean/rstd outputs");
  ST_CHECK_ARG(!q1 || (!part && qrows > 0 && hw % qrows == 0 && count > 0 && mean && rstd && (C2 == 0 || q2) && C1 % 4 == 0),
               "st_gn_apply: quad sums need qrows dividing hw, count, mean/rstd outputs and one buffer per source");
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
          (part || q1) ? 1.0 / (double)count : 0.0, eps, mean, rstd, st_seed_offset(), q1, q2, qrows, gn_order_bits() & 1);
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
          s, (const T*)dy, hw, G, splits, gamma, beta, mean, rstd, p_drop, seed, (const T*)mask, keepbits, red,
          (gn_order_bits() >> 2) & 1);
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
