// GroupNorm backward as ONE persistent launch whose two phases chase each other through L2.
//
// The two-pass backward needs the per-group sums of dz over a whole image before any dx of that image can be written.
// The cluster forms (groupnorm_cluster.cu) bridge the two phases by keeping the image in shared memory - 3 HBM passes, but
// every cluster is a load -> reduce -> barrier -> store chain that leaves the memory pipe idle half of the time - and the
// two-kernel form streams at copy speed but re-reads x and dy from HBM (5 passes).  Here the work items of both phases
// are handed out from ONE queue in an order that keeps the apply phase a fixed distance behind the reduction phase:
//
//     R(group 0)  R(group 1) A(group 0)  R(group 2) A(group 1)  ...  A(last group)
//
// with a group = as many images as fit a fraction of L2 (x + dy of two groups and dx of one stay cached).  A reduction
// item streams its pixels of (x, dy) from HBM and publishes red[n][chunk]; an apply item waits until every chunk of its
// image has been published (a per-image counter, release / acquire) and streams the SAME bytes again - now from L2 - while
// other SMs' reduction items keep HBM busy with the next group.  HBM sees x and dy once and dx once, every CTA is always
// streaming, nobody waits on a barrier.  Deadlock freedom: items are claimed in queue order by resident CTAs, an apply
// item only waits for reduction items that precede it in the queue (already claimed, and reduction items never wait).
#include "groupnorm.cuh"

namespace {

__device__ __forceinline__ int ld_acquire(const int* p) {
  int v;
  asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

// work[0] = next item, work[1 + n] = published reduction chunks of image n, work[1 + n_img] = CTAs that have left the
// loop.  All zero on entry; the last CTA to leave zeroes them again (the next launch in the stream finds them clean).
// Queue: segment k (0 <= k <= n_groups) holds the reduction items of group k (k < n_groups) followed by the apply items
// of group k - 1 (k >= 1), each part image-major with `chunks` items per image; the last group is padded with null items.
template <typename T, bool ACT, int DROP, bool CSUM>
__global__ void __launch_bounds__(256, 3) gn_bwd_wave_kernel(Src2<T> s, const T* dy, int n_img, int hw, int G,
                                                            const float* __restrict__ gamma, const float* __restrict__ beta,
                                                            const float* __restrict__ mean, const float* __restrict__ rstd,
                                                            float p_drop, uint64_t seed, const T* mask,
                                                            const uint8_t* __restrict__ keepbits, int chunks, int group,
                                                            float* red, const T* extra, float extra_scale, T* dx1, int accum1,
                                                            T* dx2, int accum2, float* csum, int* work) {
  __shared__ int s_item[2];
  pdl_wait();
  pdl_trigger();
  const int n_groups = (n_img + group - 1) / group;
  const int pg = group * chunks;                       // items of one (padded) group and phase
  const int total = 2 * n_groups * pg;
  if (threadIdx.x == 0) s_item[0] = atomicAdd(&work[0], 1);
  __syncthreads();
  for (int it = 0;; ++it) {
    const int item = s_item[it & 1];
    if (item >= total) break;
    if (threadIdx.x == 0) s_item[(it + 1) & 1] = atomicAdd(&work[0], 1);      // claim ahead: the latency hides under this item
    int phase, grp, idx;
    if (item < pg) {
      phase = 0; grp = 0; idx = item;
    } else {
      const int j = item - pg, k = 1 + j / (2 * pg), off = j % (2 * pg);
      if (k == n_groups) { phase = 1; grp = k - 1; idx = off; }
      else if (off < pg) { phase = 0; grp = k; idx = off; }
      else { phase = 1; grp = k - 1; idx = off - pg; }
    }
    const int n = grp * group + idx / chunks, chunk = idx % chunks;
    if (n < n_img) {
      if (phase == 0) {
        gn_bwd_reduce_body<T, ACT, DROP>(s, dy, hw, G, chunks, gamma, beta, mean, rstd, p_drop, seed, mask, keepbits, red, n, chunk);
        __threadfence();
        __syncthreads();
        if (threadIdx.x == 0) atomicAdd(&work[1 + n], 1);
      } else {
        if (threadIdx.x == 0) {
          while (ld_acquire(&work[1 + n]) < chunks) __nanosleep(64);
        }
        __syncthreads();
        gn_bwd_apply_body<T, ACT, DROP, CSUM>(s, dy, hw, G, chunks, gamma, beta, mean, rstd, p_drop, seed, mask, keepbits, red,
                                              extra, extra_scale, dx1, accum1, dx2, accum2, csum, nullptr, nullptr, n, n_img,
                                              chunk, chunks);
      }
    }
    __syncthreads();                                   // the bodies' shared memory is free again; the next item is visible
  }
  // ---- leave: the last CTA out resets the counters
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    s_item[0] = atomicAdd(&work[1 + n_img], 1) == (int)gridDim.x - 1;
  }
  __syncthreads();
  if (s_item[0]) {
    for (int i = threadIdx.x; i < n_img + 2; i += 256) work[i] = 0;
  }
}

}  // namespace

// Chunks per image and images per group of the wave form, or chunks = 0 when the other forms should run (tiny problems).
extern "C" __attribute__((visibility("default"))) int st_gn_bwd_wave_plan(int n_img, int hw, int C, int dtype, int* group_out) {
  // OFF by default (ST_GN_WAVE=1 turns it on): measured on B200 (B=512, bf16) the apply items' second read of x / dy comes
  // from HBM again whatever the group size (24 / 40 / 64 MB of x + dy per group: 49.1 / 48.1 / 47.1 ms per step against
  // 44.5 with the cluster / two-kernel forms; DRAM bytes per launch = two full reads) - the reduction items of the next
  // group and the dx writes push a group out of L2 before its apply items arrive.  Kept, parity-tested, as the record of
  // that measurement (profiles/r02_launches_train_step_gn_wave.md).
  static const int mode = getenv("ST_GN_WAVE") ? atoi(getenv("ST_GN_WAVE")) : 0;
  static const int l2_mb = getenv("ST_GN_WAVE_MB") ? atoi(getenv("ST_GN_WAVE_MB")) : 40;
  if (group_out) *group_out = 0;
  if (!mode) return 0;
  const int es = dtype == ST_BF16 ? 2 : 4;
  const long long img_bytes = (long long)hw * C * es * 2;               // x + dy of one image
  if ((long long)n_img * img_bytes < (8LL << 20)) return 0;             // small tensors: latency-bound, keep the cluster forms
  long long group = ((long long)l2_mb << 20) / img_bytes;
  if (group < 1) group = 1;
  if (group > n_img) group = n_img;
  // one group's reduction items should fill the resident CTAs (3 per SM) about once
  const int V = C / 8, lanes = 256 / (V < 256 ? V : 256);
  const int max_chunks = (hw + GN_DEPTH * lanes - 1) / (GN_DEPTH * lanes);      // at least one full pipeline per item
  long long want = (3LL * st_num_sms() + group - 1) / group;
  int chunks = (int)(want < 1 ? 1 : want);
  if (chunks > max_chunks) chunks = max_chunks;
  if (chunks > 1024) chunks = 1024;
  if (chunks < 1) chunks = 1;
  if (group_out) *group_out = (int)group;
  return chunks;
}

extern "C" __attribute__((visibility("default"))) int st_gn_bwd_wave(const void* x1, const void* x2, const void* dy, int dtype, int n_img, int hw, int C1,
                                                                   int C2, int G, const float* gamma, const float* beta, const float* mean,
                                                                   const float* rstd, int act, float p_drop, uint64_t seed, const void* mask,
                                                                   const uint8_t* keepbits, int chunks, int group, float* red, const void* extra,
                                                                   float extra_scale, void* dx1, int accum1, void* dx2, int accum2, float* csum,
                                                                   int* work, int reset, void* stream) {
  if (int e = check_geom(C1, C2, G)) return e;
  ST_CHECK_ARG(n_img >= 1 && chunks >= 1 && group >= 1 && group <= n_img, "st_gn_bwd_wave: bad plan (n_img %d chunks %d group %d)", n_img, chunks, group);
  ST_CHECK_ARG(((long long)n_img + group) * chunks < (1LL << 29), "st_gn_bwd_wave: too many work items");
  ST_CHECK_ARG(work && red, "st_gn_bwd_wave: work counters and red are required");
  const int drop = mask ? DROP_SLOW : (p_drop > 0.f ? (keepbits ? DROP_FAST : DROP_SLOW) : DROP_NONE);
  if (reset) {      // (the kernel leaves the counters zeroed; a caller that cannot vouch for them asks for this)
    cudaError_t me = cudaMemsetAsync(work, 0, sizeof(int) * (size_t)(n_img + 2), (cudaStream_t)stream);
    if (me != cudaSuccess) { st_set_error("st_gn_bwd_wave: memset: %s", cudaGetErrorString(me)); return ST_ERR_CUDA; }
  }
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
        auto kernel = gn_bwd_wave_kernel<T, ACT, DROP, CSUM>;
        static int resident = 0;
        if (!resident) {
          if (!allow_smem(kernel, smem)) { rc = ST_ERR_CUDA; return; }
          int per_sm = 0;
          if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, 256, smem) != cudaSuccess || per_sm < 1) per_sm = 1;
          resident = per_sm * st_num_sms();
        }
        // (apply items only wait for reduction items that running CTAs have claimed: no co-residency requirement, but
        // more CTAs than fit would only queue up behind the others)
        const long long total = 2LL * ((n_img + group - 1) / group) * group * chunks;
        const int grid = (int)(total < resident ? total : resident);
        cudaError_t e = st_launch(kernel, dim3(grid), dim3(256), smem, (cudaStream_t)stream, s, (const T*)dy, n_img, hw, G, gamma,
                                  beta, mean, rstd, p_drop, seed, (const T*)mask, keepbits, chunks, group, red, (const T*)extra,
                                  extra_scale, (T*)dx1, accum1, (T*)dx2, accum2, csum, work);
        if (e != cudaSuccess) { st_set_error("st_gn_bwd_wave: launch: %s", cudaGetErrorString(e)); rc = ST_ERR_CUDA; }
      };
      if (csum) launch(std::true_type{}); else launch(std::false_type{});
    });
  });
  if (rc) return rc;
  ST_CHECK_LAUNCH("st_gn_bwd_wave");
  return 0;
}
