// GroupNorm forms built on thread-block clusters (sm_90+ / sm_100a): both backward passes in one launch, the forward
// with the image resident in shared memory (statistics + apply: 2 HBM passes instead of 3) and the backward with x / dy
// resident between its two phases (3 passes instead of 5).  Partial sums travel through distributed shared memory
// (mapa + ld.shared::cluster).  Measured policies: see *_chunks_for and DESIGN.md section 3.
#include "groupnorm.cuh"

namespace {

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
                                                           T* dx1, int accum1, T* dx2, int accum2, float* csum, int rev) {
  pdl_wait();
  pdl_trigger();
  const int n = img_of(blockIdx.y, gridDim.y, rev), chunk = blockIdx.x, chunks = gridDim.x;
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

}  // namespace

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
                                           (T*)dx2, accum2, csum, (gn_order_bits() >> 1) & 1);
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
                                                              T* y, float* mean_out, float* rstd_out,
                                                              const uint64_t* __restrict__ seed_off, int rev) {
  extern __shared__ __align__(16) uint8_t gsm[];
  __shared__ float s_sum[512], s_sq[512];
  __shared__ float s_part[128];               // this CTA's per-group (sum, sum of squares): read by the whole cluster
  __shared__ float s_mean[64], s_rstd[64];
  pdl_wait();
  pdl_trigger();
  if constexpr (DROP == DROP_FAST) { if (seed_off) seed += *seed_off; }
  using P = Pipe<T, 1, GN_RES>;
  const P pipe(gsm);
  const int Ct = s.C1 + s.C2, cpg = Ct / G;
  const int n = img_of(blockIdx.y, gridDim.y, rev), chunk = blockIdx.x, chunks = gridDim.x;
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
                                         keepbits, (T*)y, mean, rstd, st_seed_offset(), gn_order_bits() & 1);
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
                                                                 T* dx2, float* csum, int rev) {
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
  const int n = img_of(blockIdx.y, gridDim.y, rev), chunk = blockIdx.x, chunks = gridDim.x;
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
                                           rstd, p_drop, seed, (const T*)mask, keepbits, red, (T*)dx1, (T*)dx2, csum,
                                           (gn_order_bits() >> 1) & 1);
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
