// Fused attention core of AttnBlockpp (reference models/layerspp.py:95-99) for the shape every shipped config uses at
// its 16x16 level: L = 256 keys / queries per image, C = 256 channels, one head.
//
//   w = softmax_j( sum_c q[i,c] k[j,c] * C^-1/2 )        o[i,c] = sum_j w[i,j] v[j,c]
//
// The unfused path (st_gemm -> st_softmax_fwd -> st_gemm) materialises the fp32 logits (B, 256, 256) in HBM, reads them
// back for the softmax and writes / re-reads the probabilities: ~470 MB of traffic per layer at B = 512 for 67 GFLOP.
// Here one CTA owns 128 queries of one image and nothing but q, k, v (read) and o (written) touches HBM:
//
//   warp 0     TMA producer: Q (128 x 256) and K (256 x 256) as SWIZZLE_128B K-major boxes, then - into K's buffer, as
//              soon as the S MMAs have retired - V as MN-major 64 x 64 boxes
//   warp 1     tcgen05.mma issuer: S = Q K^T into TMEM columns 0..255, later O = P V into columns 256..511
//   warps 2-5  one thread per query row: two passes over the row of S with tcgen05.ld (max; exp2 + sum), P written
//              to shared memory as the bf16 K-major A operand of the second GEMM (over Q's buffer), then the epilogue:
//              O * 1/sum -> bf16 -> shared memory (over P) -> TMA store
//
// SAVE_P (training): the rows of P are normalised in shared memory before the second GEMM and P (bf16, B x 256 x 256) is
// also stored with TMA: it is what the backward pass needs (dV = P^T dO, dS = P (dP - rowsum(P dP))).
// Persistent: CTA b walks tiles b, b + grid, ...; tiles 2n and 2n+1 are the two query halves of image n, so K and V
// of an image are fetched from HBM once and from L2 the second time.
#include <cuda.h>
#include <cudaTypedefs.h>
#include <string.h>

#include "common.cuh"
#include "tc_ptx.cuh"

bool st_tc_encode_2d(CUtensorMap* map, const void* base, uint64_t inner, uint64_t rows, uint64_t row_stride_bytes,
                     uint32_t box_inner, uint32_t box_rows);      // gemm_tc.cu (bf16, SWIZZLE_128B)

namespace {

constexpr int AT_L = 256, AT_C = 256, AT_ROWS = 128;
constexpr int AT_THREADS = 192;
constexpr int RQ_BYTES = AT_ROWS * AT_C * 2;        // 64 KB: Q, then P, then the O staging boxes
constexpr int RKV_BYTES = AT_L * AT_C * 2;          // 128 KB: K, then V
constexpr int AT_SMEM = RQ_BYTES + RKV_BYTES + 64 + 16 + 1024;

struct AttnParams {
  int n_tiles;            // 2 * images
  float k2;               // C^-1/2 * log2(e): softmax(x * scale) = exp2((x - max) * k2) / sum
};

__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
               ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1)
               : "memory");
}
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* map, uint32_t src, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(map), "r"(src), "r"(c0), "r"(c1)
               : "memory");
}
__device__ __forceinline__ float ex2f(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ uint32_t pack_bf16(float a, float b) {
  __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&h);
}

template <bool SAVE_P>
__global__ void __launch_bounds__(AT_THREADS, 1) attn_fwd_kernel(const __grid_constant__ CUtensorMap mapQ,
                                                                 const __grid_constant__ CUtensorMap mapK,
                                                                 const __grid_constant__ CUtensorMap mapV,
                                                                 const __grid_constant__ CUtensorMap mapO,
                                                                 const __grid_constant__ CUtensorMap mapP, const AttnParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + RQ_BYTES + RKV_BYTES);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 8);
  const uint32_t rq = smem_u32(smem), rkv = rq + RQ_BYTES;
  const uint32_t b0 = smem_u32(bars);
  const uint32_t q_full = b0, k_full = b0 + 8, s_full = b0 + 16, v_full = b0 + 24, p_ready = b0 + 32, o_full = b0 + 40,
                 q_free = b0 + 48;
  const int warp = threadIdx.x / 32, lane = threadIdx.x % 32;

  if (threadIdx.x == 32) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&mapQ) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&mapK) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&mapV) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&mapO) : "memory");
  }
  if (threadIdx.x == 0) {
    mbar_init(q_full, 1);
    mbar_init(k_full, 1);
    mbar_init(s_full, 1);
    mbar_init(v_full, 1);
    mbar_init(p_ready, 128);
    mbar_init(o_full, 1);
    mbar_init(q_free, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();
  pdl_trigger();

  if (warp == 0) {
    // ===================================================== TMA producer
    int it = 0;
    for (int t = blockIdx.x; t < p.n_tiles; t += gridDim.x, ++it) {
      const int img = t >> 1, row0 = img * AT_L + (t & 1) * AT_ROWS;
      const uint32_t par = it & 1, prev = (it - 1) & 1;
      if (it > 0) mbar_wait(o_full, prev);                    // the previous tile's P V MMAs have read V: K may land
      if (elect_one()) {
        mbar_expect_tx(k_full, RKV_BYTES);
#pragma unroll
        for (int kb = 0; kb < 4; ++kb) tma_load_2d(rkv + kb * 32768, &mapK, k_full, AT_C + kb * 64, img * AT_L);
      }
      __syncwarp();
      if (it > 0) mbar_wait(q_free, prev);                    // the previous tile's O store has read its staging boxes
      if (elect_one()) {
        mbar_expect_tx(q_full, RQ_BYTES);
#pragma unroll
        for (int kb = 0; kb < 4; ++kb) tma_load_2d(rq + kb * 16384, &mapQ, q_full, kb * 64, row0);
      }
      __syncwarp();
      mbar_wait(s_full, par);                                 // S complete: K's buffer is free for V
      if (elect_one()) {
        mbar_expect_tx(v_full, RKV_BYTES);
#pragma unroll
        for (int kb = 0; kb < 4; ++kb)
#pragma unroll
          for (int j = 0; j < 4; ++j)
            tma_load_2d(rkv + kb * 32768 + j * 8192, &mapV, v_full, 2 * AT_C + j * 64, img * AT_L + kb * 64);
      }
      __syncwarp();
    }
  } else if (warp == 1) {
    // ===================================================== MMA issuer
    // instruction descriptor: D = f32, A = B = bf16, N >> 3 = 32, M >> 4 = 8; bit 16: B is MN-major (V)
    const uint32_t idesc_s = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(AT_L >> 3) << 17) | ((uint32_t)(AT_ROWS >> 4) << 24);
    const uint32_t idesc_o = idesc_s | (1u << 16);
    const uint32_t desc_hi = (1024u >> 4) | (1u << 14) | (2u << 29);      // SBO 1024, version 1, SWIZZLE_128B
    int it = 0;
    for (int t = blockIdx.x; t < p.n_tiles; t += gridDim.x, ++it) {
      const uint32_t par = it & 1;
      mbar_wait(q_full, par);
      mbar_wait(k_full, par);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      if (elect_one()) {
#pragma unroll
        for (int kb = 0; kb < 4; ++kb)
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const uint64_t ad = ((uint64_t)desc_hi << 32) | (uint64_t)(((rq + kb * 16384 + j * 32) >> 4) | (1u << 16));
            const uint64_t bd = ((uint64_t)desc_hi << 32) | (uint64_t)(((rkv + kb * 32768 + j * 32) >> 4) | (1u << 16));
            umma_f16(tmem_base, ad, bd, idesc_s, (kb | j) != 0 ? 1u : 0u);
          }
        umma_commit(s_full);
      }
      __syncwarp();
      mbar_wait(v_full, par);
      mbar_wait(p_ready, par);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      if (elect_one()) {
#pragma unroll
        for (int kb = 0; kb < 4; ++kb)
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const uint64_t ad = ((uint64_t)desc_hi << 32) | (uint64_t)(((rq + kb * 16384 + j * 32) >> 4) | (1u << 16));
            const uint64_t bd = ((uint64_t)desc_hi << 32) | (uint64_t)(((rkv + kb * 32768 + j * 2048) >> 4) | ((8192u >> 4) << 16));
            umma_f16(tmem_base + 256u, ad, bd, idesc_o, (kb | j) != 0 ? 1u : 0u);
          }
        umma_commit(o_full);
      }
      __syncwarp();
    }
  } else {
    // ===================================================== softmax + epilogue: one thread per query row
    const int q = warp % 4;                         // TMEM lane quarter this warp may access
    const int row = q * 32 + lane;
    const uint32_t lane_addr = tmem_base + ((uint32_t)(q * 32) << 16);
    const uint32_t rowaddr = rq + (uint32_t)(row * 128);
    const uint32_t sw = (uint32_t)(row & 7);
    const bool leader = (warp == 2 && lane == 0);
    int it = 0;
    for (int t = blockIdx.x; t < p.n_tiles; t += gridDim.x, ++it) {
      const int img = t >> 1, row0 = img * AT_L + (t & 1) * AT_ROWS;
      const uint32_t par = it & 1;
      mbar_wait(s_full, par);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      // ---- pass 1: row maximum
      float mx = -3.0e38f;
#pragma unroll 1
      for (int c = 0; c < 8; ++c) {
        uint32_t r[32];
        tmem_ld32(lane_addr + (uint32_t)(c * 32), r);
#pragma unroll
        for (int i = 0; i < 32; ++i) mx = fmaxf(mx, __uint_as_float(r[i]));
      }
      // ---- pass 2: e = exp2((s - max) * k2), row sum, bf16 P into the K-major SWIZZLE_128B boxes over Q
      const float nmk = -mx * p.k2;
      float sum = 0.f;
#pragma unroll 1
      for (int c = 0; c < 8; ++c) {
        uint32_t r[32];
        tmem_ld32(lane_addr + (uint32_t)(c * 32), r);
        float e[32];
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          e[i] = ex2f(fmaf(__uint_as_float(r[i]), p.k2, nmk));
          sum += e[i];
        }
        const uint32_t base = rowaddr + (uint32_t)((c >> 1) * 16384);
        const uint32_t pb = (uint32_t)((c & 1) * 4);
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          const uint32_t addr = base + (((pb + g) ^ sw) << 4);
          asm volatile("st.shared.v4.u32 [%0], {%1,%2,%3,%4};" ::"r"(addr), "r"(pack_bf16(e[g * 8], e[g * 8 + 1])),
                       "r"(pack_bf16(e[g * 8 + 2], e[g * 8 + 3])), "r"(pack_bf16(e[g * 8 + 4], e[g * 8 + 5])),
                       "r"(pack_bf16(e[g * 8 + 6], e[g * 8 + 7]))
                       : "memory");
        }
      }
      const float inv = 1.f / sum;
      if constexpr (SAVE_P) {
        // normalise this thread's row in place (its own 512 bytes: no other thread touches them)
#pragma unroll 1
        for (int kb = 0; kb < 4; ++kb)
#pragma unroll
          for (int g = 0; g < 8; ++g) {
            const uint32_t addr = rowaddr + (uint32_t)(kb * 16384) + ((((uint32_t)g) ^ sw) << 4);
            uint4 v;
            asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr));
            __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&v);
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              const float2 f = __bfloat1622float2(h[k]);
              h[k] = __floats2bfloat162_rn(f.x * inv, f.y * inv);
            }
            asm volatile("st.shared.v4.u32 [%0], {%1,%2,%3,%4};" ::"r"(addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
          }
      }
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      fence_async_smem();                           // P is read by the tensor core / TMA (async proxy)
      mbar_arrive(p_ready);
      if constexpr (SAVE_P) {
        named_bar(1, 128);
        if (leader) {
#pragma unroll
          for (int bx = 0; bx < 4; ++bx) tma_store_2d(&mapP, rq + bx * 16384, bx * 64, row0);
          bulk_commit();
        }
      }
      // ---- epilogue: O = (P V) [* 1/sum] -> bf16 -> staging boxes (over P, once the MMAs and the P store have read it)
      mbar_wait(o_full, par);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      if constexpr (SAVE_P) {
        if (leader) bulk_wait_read();
        named_bar(1, 128);
      }
      const float osc = SAVE_P ? 1.f : inv;
#pragma unroll 1
      for (int c = 0; c < 8; ++c) {
        uint32_t r[32];
        tmem_ld32(lane_addr + 256u + (uint32_t)(c * 32), r);
        const uint32_t base = rowaddr + (uint32_t)((c >> 1) * 16384);
        const uint32_t pb = (uint32_t)((c & 1) * 4);
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          const uint32_t addr = base + (((pb + g) ^ sw) << 4);
          asm volatile("st.shared.v4.u32 [%0], {%1,%2,%3,%4};" ::"r"(addr),
                       "r"(pack_bf16(__uint_as_float(r[g * 8]) * osc, __uint_as_float(r[g * 8 + 1]) * osc)),
                       "r"(pack_bf16(__uint_as_float(r[g * 8 + 2]) * osc, __uint_as_float(r[g * 8 + 3]) * osc)),
                       "r"(pack_bf16(__uint_as_float(r[g * 8 + 4]) * osc, __uint_as_float(r[g * 8 + 5]) * osc)),
                       "r"(pack_bf16(__uint_as_float(r[g * 8 + 6]) * osc, __uint_as_float(r[g * 8 + 7]) * osc))
                       : "memory");
        }
      }
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      fence_async_smem();
      named_bar(1, 128);
      if (leader) {
#pragma unroll
        for (int bx = 0; bx < 4; ++bx) tma_store_2d(&mapO, rq + bx * 16384, bx * 64, row0);
        bulk_commit();
        bulk_wait_read();                           // the boxes are free: the next tile's Q may land
        mbar_arrive(q_free);
      }
    }
    if (leader) bulk_wait_all();                    // global writes complete before the CTA retires
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512));
  }
}

}  // namespace

int st_tc_available(void);

/* see st_b200.h */
extern "C" __attribute__((visibility("default"))) int st_attn_fwd_supported(int L, int C, int dtype) {
  return st_tc_available() && L == AT_L && C == AT_C && dtype == ST_BF16;
}

extern "C" __attribute__((visibility("default"))) int st_attn_fwd(const void* qkv, void* o, void* p_out, int n_img, int L, int C,
                                                                  float scale, void* stream) {
  ST_CHECK_ARG(st_attn_fwd_supported(L, C, ST_BF16), "st_attn_fwd: needs sm_100, L = 256, C = 256, bf16 (got L=%d C=%d)", L, C);
  ST_CHECK_ARG(qkv && o && n_img > 0, "st_attn_fwd: null operand");
  ST_CHECK_ARG(((reinterpret_cast<uintptr_t>(qkv) | reinterpret_cast<uintptr_t>(o) | reinterpret_cast<uintptr_t>(p_out)) & 15) == 0,
               "st_attn_fwd: operands must be 16-byte aligned");
  CUtensorMap maps[5];
  memset(maps, 0, sizeof(maps));
  const uint64_t rows = (uint64_t)n_img * L;
  if (!st_tc_encode_2d(&maps[0], qkv, 3 * C, rows, (uint64_t)3 * C * 2, 64, AT_ROWS)) return ST_ERR_CUDA;
  if (!st_tc_encode_2d(&maps[1], qkv, 3 * C, rows, (uint64_t)3 * C * 2, 64, AT_L)) return ST_ERR_CUDA;
  if (!st_tc_encode_2d(&maps[2], qkv, 3 * C, rows, (uint64_t)3 * C * 2, 64, 64)) return ST_ERR_CUDA;
  if (!st_tc_encode_2d(&maps[3], o, C, rows, (uint64_t)C * 2, 64, AT_ROWS)) return ST_ERR_CUDA;
  if (p_out && !st_tc_encode_2d(&maps[4], p_out, L, rows, (uint64_t)L * 2, 64, AT_ROWS)) return ST_ERR_CUDA;
  AttnParams prm;
  prm.n_tiles = 2 * n_img;
  prm.k2 = scale * 1.4426950408889634f;
  static bool configured = false;
  if (!configured) {
    cudaError_t e1 = cudaFuncSetAttribute(attn_fwd_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, AT_SMEM);
    cudaError_t e2 = cudaFuncSetAttribute(attn_fwd_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, AT_SMEM);
    if (e1 != cudaSuccess || e2 != cudaSuccess) { st_set_error("st_attn_fwd: cudaFuncSetAttribute failed"); return ST_ERR_CUDA; }
    configured = true;
  }
  const int grid = prm.n_tiles < st_num_sms() ? prm.n_tiles : st_num_sms();
  cudaError_t e;
  if (p_out) e = st_launch(attn_fwd_kernel<true>, dim3(grid), dim3(AT_THREADS), AT_SMEM, (cudaStream_t)stream, maps[0], maps[1], maps[2], maps[3], maps[4], prm);
  else e = st_launch(attn_fwd_kernel<false>, dim3(grid), dim3(AT_THREADS), AT_SMEM, (cudaStream_t)stream, maps[0], maps[1], maps[2], maps[3], maps[4], prm);
  if (e != cudaSuccess) { st_set_error("st_attn_fwd: launch failed: %s", cudaGetErrorString(e)); return ST_ERR_CUDA; }
  return 0;
}
