// GroupNorm backward behind the data-gradient GEMM's dz epilogue (st_gemm_args::dz_x, gemm_tc.cu).
//
// The two-pass backward reads (x, dy) twice: once to reduce the per-group sums, once to apply them.  When the GEMM that
// produces dy already turned it into dz = dy * keep * act'(u) and emitted the per-(32 rows, 4 channels) sums
// (sum gamma*dz, sum gamma*dz*xhat) from its epilogue, what is left is ONE streaming pass without any activation or
// dropout arithmetic:
//     dx = rstd * (gamma*dz - mean_g(gamma*dz) - xhat * mean_g(gamma*dz*xhat)) [+ extra_scale*extra] [+ old dx]
// which also collects the per-channel sums (sum dz, sum dz*xhat) of its pixels for the parameter gradients and the
// column sums of what it wrote (bias / time-embedding gradients of the producer).
#include "groupnorm.cuh"

namespace {

// cst[n][c] = (rstd*gamma, beta - mean*rstd*gamma, gamma, beta): what the dz epilogue needs per (image, channel)
__global__ void gn_bwd_consts_kernel(const float* __restrict__ gamma, const float* __restrict__ beta,
                                     const float* __restrict__ mean, const float* __restrict__ rstd, int n_img, int C, int G,
                                     float4* cst) {
  pdl_wait();
  pdl_trigger();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_img * C) return;
  const int n = i / C, c = i - n * C, g = c / (C / G);
  const float r = rstd[n * G + g], mu = mean[n * G + g], gm = gamma[c], bt = beta[c];
  const float rg = r * gm;
  cst[i] = make_float4(rg, fmaf(-mu, rg, bt), gm, bt);
}

// grid (chunks, n_img).  qpart [n_img * hw / 32][Ct / 4][2]: the sums the GEMM epilogue emitted.
template <typename T, bool CSUM>
__global__ void __launch_bounds__(256, 3) gn_bwd_dz_apply_kernel(Src2<T> s, const T* dz, int hw, int G,
                                                                const float* __restrict__ gamma,
                                                                const float* __restrict__ mean, const float* __restrict__ rstd,
                                                                const float* __restrict__ qpart, const T* extra,
                                                                float extra_scale, T* dx1, int accum1, T* dx2, int accum2,
                                                                float* red, float* csum, int rev) {
  extern __shared__ __align__(16) uint8_t gsm[];
  __shared__ float sq[2][256];
  __shared__ float sh1[64], sh2[64];
  pdl_wait();
  pdl_trigger();
  const int Ct = s.C1 + s.C2, cpg = Ct / G;
  const int n = img_of(blockIdx.y, gridDim.y, rev), chunk = blockIdx.x, chunks = gridDim.x;
  // ---- per-group means of gamma*dz and gamma*dz*xhat from the quad sums: 256 threads = (row-block lane, quad)
  {
    const int NQ = Ct >> 2, nrb = hw >> 5;
    const int lanes_q = 256 / NQ;                       // NQ <= 256 (host-checked)
    const int qd = threadIdx.x % NQ, ln = threadIdx.x / NQ;
    float a = 0.f, b = 0.f;
    if (ln < lanes_q) {
      const float2* q = reinterpret_cast<const float2*>(qpart) + (long long)n * nrb * NQ + qd;
      for (int rb = ln; rb < nrb; rb += lanes_q) {
        const float2 v = __ldg(q + (long long)rb * NQ);
        a += v.x;
        b += v.y;
      }
      sq[0][ln * NQ + qd] = a;
      sq[1][ln * NQ + qd] = b;
    }
    __syncthreads();
    if (threadIdx.x < G) {
      const int g = threadIdx.x, qpg = cpg >> 2;
      double A = 0., B = 0.;
      for (int l = 0; l < lanes_q; ++l)
        for (int k = 0; k < qpg; ++k) {
          A += (double)sq[0][l * NQ + g * qpg + k];
          B += (double)sq[1][l * NQ + g * qpg + k];
        }
      const double inv = 1.0 / ((double)hw * cpg);
      sh1[g] = (float)(A * inv);
      sh2[g] = (float)(B * inv);
    }
    __syncthreads();
  }
  Walk w(Ct, n, hw, chunk, chunks);
  const int V = w.V, lanes = w.lanes, v = w.v, lane = w.lane;
  const bool active = lane < lanes;
  if (!active) w.c0 = 0;
  const int c0 = w.c0;
  // per-thread constants: xhat = x*r + nmr;  dx = rg*dz + xhat*rs2 + rs1
  float rg[8], r[2], nmr[2], rs1[2], rs2[2];
  {
    float gam[8];
    load8(gamma + c0, gam);
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int g = (c0 + 4 * h) / cpg;
      const float mu = mean[n * G + g];
      r[h] = rstd[n * G + g];
      nmr[h] = -mu * r[h];
      rs1[h] = -r[h] * sh1[g];
      rs2[h] = -r[h] * sh2[g];
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) rg[i] = r[i >> 2] * gam[i];
  }
  float a[8], b[8], cs[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) a[i] = b[i] = cs[i] = 0.f;
  const bool first = c0 < s.C1;
  const int dld = first ? s.C1 : s.C2;
  const int acc = first ? accum1 : accum2;
  T* dp = (first ? dx1 + c0 : dx2 + (c0 - s.C1)) + w.row0 * dld;
  const int dstep = lanes * dld;
  using P = Pipe<T, 4, GN_BWD_DEPTH>;          // streams: x, dz, extra, old destination
  const P pipe(gsm);
  Stream<T> xs = stream_of(s, w), ds = stream_of(dz, Ct, w), es = stream_of(extra, Ct, w);
  Stream<T> os{dp, dstep};
  run_pipeline<GN_BWD_DEPTH>(
      w.n_it,
      [&](int st) {
        pipe.issue(st, 0, xs.next());
        pipe.issue(st, 1, ds.next());
        if (extra) pipe.issue(st, 2, es.next());
        if (acc) pipe.issue(st, 3, os.next());
      },
      [&](int st) {
        float x0[8], d0[8], o[8];
        pipe.read(st, 0, x0);
        pipe.read(st, 1, d0);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const float xh = fmaf(x0[i], r[i >> 2], nmr[i >> 2]);
          a[i] += d0[i];
          b[i] = fmaf(d0[i], xh, b[i]);
          o[i] = fmaf(rg[i], d0[i], fmaf(xh, rs2[i >> 2], rs1[i >> 2]));
        }
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
      });
  // ---- per-channel sums of this block's pixels (fixed-order lane reduction through the drained pipeline memory):
  // red[n][chunk][c][2] = (sum dz, sum dz*xhat) -> parameter gradients;  csum[n][chunk][c] = column sums of the output
  __syncthreads();
  float* s_red = reinterpret_cast<float*>(gsm);
  if (active) {
    float* o = s_red + ((size_t)lane * V + v) * 16;
#pragma unroll
    for (int i = 0; i < 8; ++i) { o[2 * i] = a[i]; o[2 * i + 1] = b[i]; }
  }
  __syncthreads();
  if (red)
    for (int col = threadIdx.x; col < 16 * V; col += 256) {
      float t = 0.f;
      for (int l = 0; l < lanes; ++l) t += s_red[(size_t)l * V * 16 + col];
      red[((long long)n * chunks + chunk) * Ct * 2 + col] = t;
    }
  if constexpr (CSUM) {
    __syncthreads();
    if (active) {
#pragma unroll
      for (int i = 0; i < 8; ++i) s_red[(lane * V + v) * 8 + i] = cs[i];
    }
    __syncthreads();
    for (int col = threadIdx.x; col < Ct; col += 256) {
      float t = 0.f;
      for (int l = 0; l < lanes; ++l) t += s_red[l * Ct + col];
      csum[((long long)n * chunks + chunk) * Ct + col] = t;
    }
  }
}

}  // namespace

extern "C" __attribute__((visibility("default"))) int st_gn_bwd_consts(const float* gamma, const float* beta, const float* mean, const float* rstd,
                                                                     int n_img, int C, int G, float* cst, void* stream) {
  ST_CHECK_ARG(n_img > 0 && C > 0 && G > 0 && C % G == 0, "st_gn_bwd_consts: bad geometry");
  ST_CHECK_ARG((reinterpret_cast<uintptr_t>(cst) & 15) == 0, "st_gn_bwd_consts: cst must be 16-byte aligned");
  const int n = n_img * C;
  st_launch(gn_bwd_consts_kernel, dim3((n + 255) / 256), dim3(256), 0, (cudaStream_t)stream, gamma, beta, mean, rstd, n_img, C,
            G, reinterpret_cast<float4*>(cst));
  ST_CHECK_LAUNCH("st_gn_bwd_consts");
  return 0;
}

extern "C" __attribute__((visibility("default"))) int st_gn_bwd_dz_apply(const void* x1, const void* x2, const void* dz, int dtype, int n_img, int hw,
                                                                       int C1, int C2, int G, const float* gamma, const float* mean,
                                                                       const float* rstd, const float* qpart, const void* extra,
                                                                       float extra_scale, void* dx1, int accum1, void* dx2, int accum2,
                                                                       int chunks, float* red, float* csum, void* stream) {
  if (int e = check_geom(C1, C2, G)) return e;
  const int Ct = C1 + C2;
  ST_CHECK_ARG(n_img <= 65535, "st_gn_bwd_dz_apply: more than 65535 images");
  ST_CHECK_ARG(hw % 32 == 0 && Ct <= 1024, "st_gn_bwd_dz_apply: hw must be a multiple of 32 and C <= 1024 (got %d, %d)", hw, Ct);
  ST_CHECK_ARG(chunks >= 1 && chunks <= 65535, "st_gn_bwd_dz_apply: bad chunk count %d", chunks);
  ST_CHECK_ARG(qpart && dz, "st_gn_bwd_dz_apply: dz and its quad sums are required");
  int rc = 0;
  ST_DISPATCH_DTYPE(dtype, T, {
    Src2<T> s{(const T*)x1, (const T*)x2, C1, C2};
    constexpr int smem = Pipe<T, 4, GN_BWD_DEPTH>::BYTES;
    auto launch = [&](auto CS) {
      constexpr bool CSUM = decltype(CS)::value;
      static bool smem_ok = false;
      if (!smem_ok) { if (!allow_smem(gn_bwd_dz_apply_kernel<T, CSUM>, smem)) { rc = ST_ERR_CUDA; return; } smem_ok = true; }
      st_launch(gn_bwd_dz_apply_kernel<T, CSUM>, dim3(chunks, n_img), dim3(256), smem, (cudaStream_t)stream, s, (const T*)dz, hw,
                G, gamma, mean, rstd, qpart, (const T*)extra, extra_scale, (T*)dx1, accum1, (T*)dx2, accum2, red, csum,
                (gn_order_bits() >> 1) & 1);
    };
    if (csum) launch(std::true_type{}); else launch(std::false_type{});
  });
  if (rc) return rc;
  ST_CHECK_LAUNCH("st_gn_bwd_dz_apply");
  return 0;
}
