// tcgen05 + TMA implicit-GEMM backend of st_gemm (bf16 operands, fp32 accumulation in TMEM).
//
// One CTA computes a 128 x BN output tile.  Warp roles (192 threads):
//   warp 0    TMA producer: one elected thread walks the K blocks (64 bf16 = one 128-byte swizzle row),
//             issuing cp.async.bulk.tensor loads into a STAGES-deep shared-memory ring (mbarrier full/empty).
//   warp 1    TMEM allocator + MMA issuer: one thread issues 4 x tcgen05.mma (K=16 each) per K block,
//             tcgen05.commit releases the smem slot / signals the epilogue.
//   warps 2-5 epilogue: tcgen05.ld 32 columns at a time, bias / per-image bias / residual / scale, store
//             (or fp32 red.add for split-K weight gradients).
//
// Operand addressing is entirely in the TMA descriptors (built on the host per call):
//   * K-major strided          2-D/3-D map, box 64(k) x rows
//   * MN-major strided         box 64(m|n) x 64(k); UMMA descriptors flagged MN-major
//   * implicit im2col (conv)   4-D map over NHWC (C, W, H, N); a 3x3 tap is a box shifted by (dx, dy) whose
//                              out-of-bounds elements TMA zero-fills -> no im2col buffer, no halo code.
//                              The channel axis may be the concatenation of two tensors (two maps).
//   * conv weights for dgrad   3-D map (Ci, taps, Co), MN-major, taps walked in reverse.
#include <cuda.h>
#include <cudaTypedefs.h>

#include <stdlib.h>
#include <string.h>

#include <mutex>

#include "common.cuh"
#include "tc_ptx.cuh"

namespace {

constexpr int BM = 128;
constexpr int BK = 64;                 // bf16 elements per K block = 128 bytes
constexpr int A_STAGE_BYTES = BM * 128;
constexpr int NUM_THREADS = 192;

enum OpKind { KMAJOR = 0, MNMAJOR = 1, GATHER_K = 2, GATHER_MN = 3, DGRADW = 4 };
// GATHER_MN is valid for either operand: as B it is the v1 weight-gradient form, as A it is the transposed
// weight-gradient form of the persistent kernel (M = taps*Cin, N = Cout, output stored transposed).

struct TcParams {
  int M, N, batch, split_k;
  int nk;                 // K blocks in total
  int a_kind, b_kind;
  // gather geometry
  int H, W, kh, kw, cblocks, c1blocks, ntaps;
  int Ct, C1;
  int logW, logH;         // H, W are powers of two (checked on the host)
  // epilogue
  void* C;
  long long ldc, sCb;
  int out_bf16, accumulate;
  const float* bias;
  const float* rowbias;
  int rows_per_rb;
  long long ld_rb;
  const bf16* residual;
  long long ldr, sRb;
  float alpha;
  // persistent kernel only
  int cluster;            // CTAs per cluster along M (B tiles are multicast across them)
  int trans_out;          // store C[n][m] instead of C[m][n] (accumulate mode)
  int m_tiles, n_tiles;
  int epi_tma;            // epilogue writes the tile through shared memory with TMA stores (mapC / mapR are valid)
  int rb_rows;            // distinct rowbias rows (images) one tile spans (1..4) when epi_tma
  // GroupNorm by-product (TMA epilogue, bf16 output, 256-column accumulators): per (gn_rows output rows, 4 channels) the
  // sum and sum of squares of the STORED values -> gn_part[row / gn_rows][N / 4][2]; nullptr = off
  float* gn_part;
  int gn_rows;            // 16, 64 or 128: min(pixels per image, 128)
  // GroupNorm backward phase 1 in the epilogue (st_gemm_args::dz_x; TMA epilogue, bf16 output, 256-column accumulators):
  // the tile of x arrives like a residual tile (mapR), dz replaces dy in the staged tile, and every epilogue warp emits
  // the (sum gamma*dz, sum dz*(u-beta)) of its 32 rows per 4-channel quad -> gn_part[row / 32][N / 4][2]
  const float4* dz_cst;   // nullptr = off
  const uint8_t* dz_keep;
  float dz_ks;            // alpha / (1 - p_drop)
  int dz_act, dz_loghw;
};

// 16 per-lane values -> their sums over the 32 lanes of the warp, value i left in lanes 2i and 2i+1 (recursive halving:
// 8 + 4 + 2 + 1 + 1 shuffles)
__device__ __forceinline__ float warp_reduce16(float (&v)[16], int lane) {
  float a[8], b[4], c[2];
  {
    const bool hi = lane & 16;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const float send = hi ? v[k] : v[8 + k], keep = hi ? v[8 + k] : v[k];
      a[k] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
    }
  }
  {
    const bool hi = lane & 8;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float send = hi ? a[k] : a[4 + k], keep = hi ? a[4 + k] : a[k];
      b[k] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
    }
  }
  {
    const bool hi = lane & 4;
#pragma unroll
    for (int k = 0; k < 2; ++k) {
      const float send = hi ? b[k] : b[2 + k], keep = hi ? b[2 + k] : b[k];
      c[k] = keep + __shfl_xor_sync(0xffffffffu, send, 4);
    }
  }
  const bool hi = lane & 2;
  const float send = hi ? c[0] : c[1], keep = hi ? c[1] : c[0];
  float d = keep + __shfl_xor_sync(0xffffffffu, send, 2);
  d += __shfl_xor_sync(0xffffffffu, d, 1);
  return d;
}

// ------------------------------------------------------------------------------------ kernel
template <int BN, int STAGES>
__global__ void __launch_bounds__(NUM_THREADS) gemm_tc_kernel(const __grid_constant__ CUtensorMap mapA0,
                                                              const __grid_constant__ CUtensorMap mapA1,
                                                              const __grid_constant__ CUtensorMap mapB0,
                                                              const __grid_constant__ CUtensorMap mapB1,
                                                              const TcParams p) {
  constexpr int B_STAGE_BYTES = BN * 128;
  constexpr int STAGE_BYTES = A_STAGE_BYTES + B_STAGE_BYTES;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  // carve: stages | barriers | tmem ptr  (the dynamic smem base is re-aligned to 1024 for SWIZZLE_128B)
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + STAGES * STAGE_BYTES);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * STAGES + 1);
  const uint32_t smem_base = smem_u32(smem);
  const uint32_t full0 = smem_u32(bars), empty0 = smem_u32(bars + STAGES), tfull = smem_u32(bars + 2 * STAGES);

  const int warp = threadIdx.x / 32, lane = threadIdx.x % 32;
  const int m0 = blockIdx.x * BM, n0 = blockIdx.y * BN;
  const int split = p.split_k > 1 ? p.split_k : 1;
  const int b = blockIdx.z / split, ks = blockIdx.z % split;
  const int kper = (p.nk + split - 1) / split;
  const int kt0 = ks * kper;
  const int kt1 = min(p.nk, kt0 + kper);
  const int nkt = kt1 - kt0;            // may be <= 0 for a trailing split: nothing to add

  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(full0 + 8 * s, 1);
      mbar_init(empty0 + 8 * s, 1);
    }
    mbar_init(tfull, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(BN));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = *tmem_slot;

  if (nkt > 0) {
    if (warp == 0) {
      // ===================================================== TMA producer
      if (lane == 0) {
        // tile-constant gather coordinates
        int gx0 = 0, gy0 = 0, gn0 = 0;
        if (p.a_kind == GATHER_K) {
          gx0 = m0 % p.W;
          gy0 = (m0 / p.W) % p.H;
          gn0 = m0 / (p.W * p.H);
        }
        for (int i = 0; i < nkt; ++i) {
          const int kt = kt0 + i;
          const int s = i % STAGES;
          if (i >= STAGES) mbar_wait(empty0 + 8 * s, ((i / STAGES) - 1) & 1);
          const uint32_t bar = full0 + 8 * s;
          mbar_expect_tx(bar, STAGE_BYTES);
          const uint32_t sa = smem_base + s * STAGE_BYTES, sb = sa + A_STAGE_BYTES;
          // ---- A
          if (p.a_kind == KMAJOR) {
            tma_load_3d(sa, &mapA0, bar, kt * BK, m0, b);
          } else if (p.a_kind == MNMAJOR) {
            tma_load_3d(sa, &mapA0, bar, m0, kt * BK, b);
            tma_load_3d(sa + 8192, &mapA0, bar, m0 + 64, kt * BK, b);
          } else {   // GATHER_K
            const int tap = kt / p.cblocks, cb = kt % p.cblocks;
            const int dy = tap / p.kw - (p.kh - 1) / 2, dx = tap % p.kw - (p.kw - 1) / 2;
            if (cb < p.c1blocks) tma_load_4d(sa, &mapA0, bar, cb * 64, gx0 + dx, gy0 + dy, gn0);
            else tma_load_4d(sa, &mapA1, bar, (cb - p.c1blocks) * 64, gx0 + dx, gy0 + dy, gn0);
          }
          // ---- B
          if (p.b_kind == KMAJOR) {
            tma_load_3d(sb, &mapB0, bar, kt * BK, n0, b);
          } else if (p.b_kind == MNMAJOR) {
#pragma unroll
            for (int j = 0; j < BN / 64; ++j) tma_load_3d(sb + j * 8192, &mapB0, bar, n0 + 64 * j, kt * BK, b);
          } else if (p.b_kind == DGRADW) {
            const int tap = kt / p.cblocks, cob = kt % p.cblocks;
#pragma unroll
            for (int j = 0; j < BN / 64; ++j)
              tma_load_3d(sb + j * 8192, &mapB0, bar, n0 + 64 * j, p.ntaps - 1 - tap, cob * 64);
          } else {   // GATHER_MN: k = pixel block, n = (tap, channel)
            const int p0 = kt * BK;
            const int x0 = p0 % p.W, y0 = (p0 / p.W) % p.H, i0 = p0 / (p.W * p.H);
#pragma unroll
            for (int j = 0; j < BN / 64; ++j) {
              const int nn = n0 + 64 * j;
              const int tap = nn / p.Ct, c = nn % p.Ct;
              const int dy = tap / p.kw - (p.kh - 1) / 2, dx = tap % p.kw - (p.kw - 1) / 2;
              if (tap >= p.ntaps) {
                // past the last tap (N not a multiple of BN): any in-range box will do, the columns are masked;
                // load tap 0 so that the expected byte count still arrives
                tma_load_4d(sb + j * 8192, &mapB0, bar, 0, x0, y0, i0);
              } else if (c < p.C1) {
                tma_load_4d(sb + j * 8192, &mapB0, bar, c, x0 + dx, y0 + dy, i0);
              } else {
                tma_load_4d(sb + j * 8192, &mapB1, bar, c - p.C1, x0 + dx, y0 + dy, i0);
              }
            }
          }
        }
      }
    } else if (warp == 1) {
      // ===================================================== MMA issuer
      if (lane == 0) {
        const bool a_mn = (p.a_kind == MNMAJOR);
        const bool b_mn = (p.b_kind != KMAJOR);
        // instruction descriptor: D=f32, A=B=bf16, majors, N>>3, M>>4
        const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((a_mn ? 1u : 0u) << 15) | ((b_mn ? 1u : 0u) << 16) |
                               ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
        for (int i = 0; i < nkt; ++i) {
          const int s = i % STAGES;
          mbar_wait(full0 + 8 * s, (i / STAGES) & 1);
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          const uint32_t sa = smem_base + s * STAGE_BYTES, sb = sa + A_STAGE_BYTES;
#pragma unroll
          for (int j = 0; j < BK / 16; ++j) {
            const uint64_t ad = a_mn ? make_desc(sa + j * 2048, 8192, 1024) : make_desc(sa + j * 32, 16, 1024);
            const uint64_t bd = b_mn ? make_desc(sb + j * 2048, 8192, 1024) : make_desc(sb + j * 32, 16, 1024);
            umma_f16(tmem_base, ad, bd, idesc, (i | j) != 0 ? 1u : 0u);
          }
          umma_commit(empty0 + 8 * s);      // frees the smem slot once these MMAs have read it
        }
        umma_commit(tfull);                 // accumulator complete
      }
    } else {
      // ===================================================== epilogue (warps 2..5)
      const int q = warp % 4;               // TMEM lane quarter this warp may access
      const int row = q * 32 + lane;
      const long long m = (long long)m0 + row;
      mbar_wait(tfull, 0);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const bool row_ok = m < p.M;
      const float* rb = (p.rowbias && row_ok) ? p.rowbias + (m / p.rows_per_rb) * p.ld_rb : nullptr;
      const bf16* res = (p.residual && row_ok) ? p.residual + (long long)b * p.sRb + m * p.ldr : nullptr;
      const long long crow = (long long)b * p.sCb + m * p.ldc;
#pragma unroll 1
      for (int c = 0; c < BN / 32; ++c) {
        uint32_t r[32];
        __syncwarp();                       // tcgen05.ld is warp-collective: reconverge after the masked `continue`s
        tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(c * 32), r);
        const int nb = n0 + c * 32;
        if (!row_ok || nb >= p.N) continue;
        float v[32];
#pragma unroll
        for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
        const bool full = (nb + 32 <= p.N);
        if (p.accumulate) {
          float* dst = reinterpret_cast<float*>(p.C) + crow + nb;
#pragma unroll
          for (int i = 0; i < 32; ++i)
            if (full || nb + i < p.N) atomicAdd(dst + i, p.alpha * v[i]);
          continue;
        }
        if (p.bias) {
#pragma unroll
          for (int i = 0; i < 32; ++i)
            if (full || nb + i < p.N) v[i] += __ldg(p.bias + nb + i);
        }
        if (rb) {
#pragma unroll
          for (int i = 0; i < 32; ++i)
            if (full || nb + i < p.N) v[i] += __ldg(rb + nb + i);
        }
        if (res) {
          if (full && ((p.ldr | p.sRb) % 8 == 0)) {
#pragma unroll
            for (int g = 0; g < 4; ++g) {
              uint4 t = *reinterpret_cast<const uint4*>(res + nb + g * 8);
              const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&t);
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                v[g * 8 + 2 * e] += __low2float(h[e]);
                v[g * 8 + 2 * e + 1] += __high2float(h[e]);
              }
            }
          } else {
#pragma unroll
            for (int i = 0; i < 32; ++i)
              if (nb + i < p.N) v[i] += __bfloat162float(res[nb + i]);
          }
        }
#pragma unroll
        for (int i = 0; i < 32; ++i) v[i] *= p.alpha;
        if (p.out_bf16) {
          bf16* dst = reinterpret_cast<bf16*>(p.C) + crow + nb;
          if (full && ((p.ldc | p.sCb) % 8 == 0)) {
#pragma unroll
            for (int g = 0; g < 4; ++g) {
              uint4 t;
              __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&t);
#pragma unroll
              for (int e = 0; e < 4; ++e) h[e] = __floats2bfloat162_rn(v[g * 8 + 2 * e], v[g * 8 + 2 * e + 1]);
              *reinterpret_cast<uint4*>(dst + g * 8) = t;
            }
          } else {
#pragma unroll
            for (int i = 0; i < 32; ++i)
              if (nb + i < p.N) dst[i] = __float2bfloat16_rn(v[i]);
          }
        } else {
          float* dst = reinterpret_cast<float*>(p.C) + crow + nb;
          if (full && ((p.ldc | p.sCb) % 4 == 0)) {
#pragma unroll
            for (int g = 0; g < 8; ++g)
              *reinterpret_cast<float4*>(dst + g * 4) = make_float4(v[g * 4], v[g * 4 + 1], v[g * 4 + 2], v[g * 4 + 3]);
          } else {
#pragma unroll
            for (int i = 0; i < 32; ++i)
              if (nb + i < p.N) dst[i] = v[i];
          }
        }
      }
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    }
  }
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(BN));
  }
}


// ------------------------------------------------------------------------------------ persistent kernel (v2)
// Same producer / MMA roles as gemm_tc_kernel, but
//   * one CTA per SM loops over output tiles (static round-robin over clusters), the smem ring keeps streaming
//     across tile boundaries,
//   * the accumulator is double-buffered in TMEM (2 x BN columns): the epilogue of tile i overlaps the MMAs of
//     tile i+1; EIGHT epilogue warps (two per TMEM lane quarter, each taking half of the columns),
//   * the residual tile of the epilogue is prefetched with cp.async into thread-private shared-memory slots
//     while the tile's MMAs are still running, so no global-load latency sits between tcgen05.ld and the store,
//   * `cluster` CTAs that work on M-adjacent tiles share the B tile: each loads 1/cluster of it and TMA-multicasts
//     its slice into every CTA of the cluster (L2 -> SM traffic for B drops by `cluster`),
//   * weight gradients run transposed (A = shifted NHWC boxes MN-major, B = dY MN-major, C stored transposed) so
//     that the big dimension taps*Cin is M and the dY tile is the multicast operand.
// GroupNorm by-product of the TMA epilogue: column sums over a half-group's staged 128 x 128 bf16 tile (two SWIZZLE_128B
// boxes of 64 columns), per 4-channel quad and per 128 / NP rows.  A warp owns 32 columns (4 swizzle chunks of one box);
// lane = (row within an 8-row block) x (chunk): one LDS.128 per lane covers 8 rows x 64 bytes = every bank once per
// wavefront (the swizzle spreads the 8 rows), 16 loads per lane for the tile, then three shuffles over the row lanes.
template <int NP>
__device__ __forceinline__ void gn_quad_pass(uint32_t region, int w4, int lane, float* part, long long qn, int gq, int grow0,
                                             int M, bool col_ok) {
  const uint32_t rs = (uint32_t)(lane >> 2);
  const uint32_t c16 = (uint32_t)((w4 & 1) * 4 + (lane & 3));
  const uint32_t addr0 = region + (uint32_t)((w4 >> 1) * 16384) + rs * 128u + ((c16 ^ rs) << 4);
  float sm[NP][2], sq[NP][2];
#pragma unroll
  for (int i = 0; i < NP; ++i) { sm[i][0] = sm[i][1] = sq[i][0] = sq[i][1] = 0.f; }
#pragma unroll
  for (int i = 0; i < 16; ++i) {
    uint4 t;
    asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(t.x), "=r"(t.y), "=r"(t.z), "=r"(t.w) : "r"(addr0 + (uint32_t)(i * 1024)));
    const float2 a = __bfloat1622float2(*reinterpret_cast<__nv_bfloat162*>(&t.x));
    const float2 b = __bfloat1622float2(*reinterpret_cast<__nv_bfloat162*>(&t.y));
    const float2 c = __bfloat1622float2(*reinterpret_cast<__nv_bfloat162*>(&t.z));
    const float2 d = __bfloat1622float2(*reinterpret_cast<__nv_bfloat162*>(&t.w));
    constexpr int PER = 16 / NP;
    const int k = i / PER;
    sm[k][0] += (a.x + a.y) + (b.x + b.y);
    sq[k][0] += fmaf(a.x, a.x, fmaf(a.y, a.y, fmaf(b.x, b.x, b.y * b.y)));
    sm[k][1] += (c.x + c.y) + (d.x + d.y);
    sq[k][1] += fmaf(c.x, c.x, fmaf(c.y, c.y, fmaf(d.x, d.x, d.y * d.y)));
  }
#pragma unroll
  for (int k = 0; k < NP; ++k)
#pragma unroll
    for (int h = 0; h < 2; ++h)
#pragma unroll
      for (int o = 4; o <= 16; o <<= 1) {
        sm[k][h] += __shfl_xor_sync(0xffffffffu, sm[k][h], o);
        sq[k][h] += __shfl_xor_sync(0xffffffffu, sq[k][h], o);
      }
  if (rs == 0 && col_ok) {
    constexpr int PR = 128 / NP;
#pragma unroll
    for (int k = 0; k < NP; ++k) {
      const int row = grow0 + k * PR;
      if (row < M)
        *reinterpret_cast<float4*>(part + ((long long)(row / PR) * qn + gq) * 2) = make_float4(sm[k][0], sq[k][0], sm[k][1], sq[k][1]);
    }
  }
}

constexpr int NUM_THREADS2 = 352;     // warp 0 TMA (A operand), warp 1 MMA, warps 2..9 epilogue, warp 10 TMA (B operand)

// MH = number of 128-row accumulators per CTA tile: the tile is (128*MH) x BN with MH*BN <= 256 TMEM columns per
// buffer.  MH = 2 (256 x 128) is the N <= 128 counterpart of the 128 x 256 tile: both stream 48 KB per K block
// for 4.2 MFLOP, i.e. the single TMA / MMA issuing threads have 512 tensor-core cycles per K block to hide behind.
//
// CG = 2 (CTA pair, tcgen05.mma.cta_group::2): the two CTAs of a cluster compute one (2*128*MH) x BN tile.  Each CTA
// stages its own 128*MH rows of A and HALF of the B tile (the tensor core reads the other half from the peer's shared
// memory), so a K block costs 16*MH + BN/16 KB per SM instead of 16*MH + BN/8 and a tensor-core instruction reads
// 4 + BN/64 KB of shared memory per SM instead of 4 + BN/32: deeper ring, less shared-memory bandwidth per flop.
// Protocol: both CTAs' producers arm and signal the LEADER's full barrier (4 arrivals + the bytes of both CTAs); the
// leader's MMA thread issues for the pair and commits with .multicast::cluster to the empty / accumulator-full
// barriers of both CTAs; every epilogue warp of the pair arrives on the leader's accumulator-empty barrier.
//
// HALO (3x3 convolutions whose row tile lies inside one image and spans >= 4 image rows; CTA pairs only): the three
// vertical taps (dy = -1, 0, +1) of one (dx, channel block) read the SAME pixels shifted by whole image rows, so the
// producer loads ONE box of (tile rows + 2) image rows and the tensor core reads the three taps from it through
// descriptor start offsets of dy * W * 128 bytes (a multiple of the 1024-byte swizzle period for W >= 8): the A operand
// costs (rows + 2) / (3 rows) of the bytes (1/2 at 32x32, 3/8 at 16x16) in shared-memory writes and L2 reads.  A stage
// is then one such box plus the three weight slabs and feeds 12 (x MH) tensor-core instructions.
template <int BN, int MH, int STAGES, int CG, bool HALO = false>
__global__ void __launch_bounds__(NUM_THREADS2, 1) gemm_tc2_kernel(const __grid_constant__ CUtensorMap mapA0,
                                                                   const __grid_constant__ CUtensorMap mapA1,
                                                                   const __grid_constant__ CUtensorMap mapB0,
                                                                   const __grid_constant__ CUtensorMap mapB1,
                                                                   const __grid_constant__ CUtensorMap mapC,
                                                                   const __grid_constant__ CUtensorMap mapR,
                                                                   const TcParams p) {
  static_assert(BN * MH <= 256, "accumulator buffer exceeds half of TMEM");
  static_assert(CG == 1 || (BN % 128 == 0), "a CTA pair splits the B tile in 64-wide halves");
  constexpr int BMT = BM * MH;                // tile rows of this CTA
  constexpr int A_BYTES = A_STAGE_BYTES * MH;
  constexpr int B_STAGE_BYTES = BN * 128 / CG;     // this CTA's share of the B tile
  constexpr int A_SLOT_BYTES = HALO ? A_BYTES * 3 / 2 : A_BYTES;     // halo box: at most 1.5 x the tile (>= 4 rows + 2)
  constexpr int STAGE_BYTES = A_SLOT_BYTES + (HALO ? 3 : 1) * B_STAGE_BYTES;
  constexpr int TC = BN * MH;                 // TMEM columns per accumulator buffer
  constexpr int TMEM_COLS = 2 * TC;
  constexpr int CH = TC / 64;                 // 32-column chunks per epilogue warp
  constexpr int RES_BYTES = TC * 256;         // 256 epilogue threads x CH chunks x 64 bytes
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  {
    uint32_t dyn;
    asm volatile("mov.u32 %0, %%dynamic_smem_size;" : "=r"(dyn));
    constexpr uint32_t NEED = STAGES * STAGE_BYTES + TC * 256 + (2 * STAGES + 6) * 8 + 16 + 4 * BN * 4;     // (see launch2)
    if ((uint32_t)(smem - smem_raw) + NEED > dyn) __trap();     // launched without re-alignment slack and misaligned
  }
  uint8_t* res_stage = smem + STAGES * STAGE_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(res_stage + RES_BYTES);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * STAGES + 6);
  float* sbias = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(tmem_slot) + 16);     // [<=4 rows][BN] (TMA epilogue)
  const uint32_t smem_base = smem_u32(smem);
  const uint32_t full0 = smem_u32(bars), empty0 = smem_u32(bars + STAGES);
  const uint32_t tfull0 = smem_u32(bars + 2 * STAGES), tempty0 = smem_u32(bars + 2 * STAGES + 2);
  const uint32_t rfull0 = smem_u32(bars + 2 * STAGES + 4);     // residual tile landed (one per epilogue half-group)

  const int warp = threadIdx.x / 32, lane = threadIdx.x % 32;
  const int cs = CG == 2 ? 2 : p.cluster;
  const int rank = cs > 1 ? (int)cluster_ctarank() : 0;
  const bool leader_cta = (CG == 1) || rank == 0;
  const uint16_t mc_mask = (uint16_t)((1u << cs) - 1u);
  const int cluster_id = blockIdx.x / cs, n_clusters = gridDim.x / cs;
  const int split = p.split_k > 1 ? p.split_k : 1;
  const int m_groups = (p.m_tiles + cs - 1) / cs;
  const int total = p.batch * split * p.n_tiles * m_groups;
  const int kper = (p.nk + split - 1) / split;

  if (threadIdx.x == 32) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&mapA0) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&mapB0) : "memory");
    if (p.epi_tma) asm volatile("prefetch.tensormap [%0];" ::"l"(&mapC) : "memory");
  }
  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(full0 + 8 * s, 2 * CG);     // one arrive.expect_tx from each producer warp (of both CTAs of a pair)
      mbar_init(empty0 + 8 * s, CG == 2 ? 1 : cs);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(tfull0 + 8 * b, 1);
      mbar_init(tempty0 + 8 * b, 8 * CG);   // every epilogue warp (of both CTAs of a pair)
      mbar_init(rfull0 + 8 * b, 1);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    if constexpr (CG == 2) {
      asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(TMEM_COLS));
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::);
    } else {
      asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(TMEM_COLS));
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (cs > 1) cluster_sync_all();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = *tmem_slot;
  // everything above (barrier init, TMEM allocation) may overlap the tail of the previous kernel in the stream
  pdl_wait();
  pdl_trigger();

  // decode a work item
  // Tile order: CTAs that run at the same time should share operand tiles in L2.  n fastest (all N tiles of one block
  // of rows are in flight together, the activation rows are fetched from DRAM once) except for weight gradients,
  // whose big operands are both indexed by k: there every (m, n) tile of one K split runs together.
  auto decode = [&](int st, int& m0, int& n0, int& b, int& kt0, int& nkt) {
    int mg, nt;
    if (p.trans_out) { mg = st % m_groups; nt = (st / m_groups) % p.n_tiles; }
    else { nt = st % p.n_tiles; mg = (st / p.n_tiles) % m_groups; }
    const int z = st / (m_groups * p.n_tiles);
    b = z / split;
    const int ks = z % split;
    m0 = (mg * cs + rank) * BMT;
    n0 = nt * BN;
    kt0 = ks * kper;
    nkt = min(p.nk, kt0 + kper) - kt0;
  };

  // Both single-thread loops below are latency-bound chains of dependent scalar instructions, so they are written
  // with incremental state only: no integer division, no modulo, descriptors advanced by adding constants.
  if (warp == 0 || warp == 10) {
    // ===================================================== TMA producers: warp 0 streams the A operand, warp 10 the B operand
    // (a single elected thread cannot issue the 4-6 boxes per K block of the MN-major forms fast enough);
    // the whole warp walks the loop, one elected lane issues
    const bool do_a = (warp == 0);
    if constexpr (HALO) {
      int s = 0;
      uint32_t par = 1;
      const int lw = p.logW, lhw = p.logW + p.logH, Hm = p.H - 1;
      const int nsb = p.nk / 3;                           // (dx, channel block) super-blocks of three taps each
      const uint32_t a_bytes = (uint32_t)(BMT + 2 * p.W) * 128u;
      for (int st = cluster_id; st < total; st += n_clusters) {
        int m0, n0, b, kt0, nkt;
        decode(st, m0, n0, b, kt0, nkt);
        const int gy = (m0 >> lw) & Hm, gn = m0 >> lhw;   // the tile starts at column 0 of image row gy
        int dxi = 0, cb = 0;
        for (int i = 0; i < nsb; ++i) {
          mbar_wait(empty0 + 8 * s, par);
          const uint32_t bar = CG == 2 ? ((full0 + 8 * s) & PEER_MASK) : (full0 + 8 * s);
          const uint32_t sa = smem_base + s * STAGE_BYTES, sb = sa + A_SLOT_BYTES;
          if (elect_one()) {
            const uint32_t bytes = do_a ? a_bytes : 3u * B_STAGE_BYTES;
            if constexpr (CG == 2) mbar_expect_tx_cluster(bar, bytes);
            else mbar_expect_tx(bar, bytes);
            if (do_a) {
              if (cb < p.c1blocks) ld4<CG>(sa, &mapA0, bar, cb * 64, dxi - 1, gy - 1, gn);
              else ld4<CG>(sa, &mapA1, bar, (cb - p.c1blocks) * 64, dxi - 1, gy - 1, gn);
            } else {
#pragma unroll
              for (int dyi = 0; dyi < 3; ++dyi)
                ld3<CG>(sb + dyi * B_STAGE_BYTES, &mapB0, bar, ((dyi * 3 + dxi) * p.cblocks + cb) * 64, n0 + rank * (BN / CG), b);
            }
          }
          __syncwarp();
          if (++cb == p.cblocks) { cb = 0; ++dxi; }
          if (++s == STAGES) { s = 0; par ^= 1; }
        }
      }
    } else {
      int s = 0;
      uint32_t par = 1;                 // parity to wait on empty[s]: a fresh barrier passes parity 1 immediately
      const int pw = (p.kw - 1) / 2, ph = (p.kh - 1) / 2;
      const int lw = p.logW, lhw = p.logW + p.logH;
      const int Wm = p.W - 1, Hm = p.H - 1;
      for (int st = cluster_id; st < total; st += n_clusters) {
        int m0, n0, b, kt0, nkt;
        decode(st, m0, n0, b, kt0, nkt);
        if (nkt <= 0) continue;
        // ---- per-tile constants
        int gx0[MH], gy0[MH], gn0[MH];
#pragma unroll
        for (int h = 0; h < MH; ++h) {
          const int mh = m0 + h * BM;
          gx0[h] = mh & Wm; gy0[h] = (mh >> lw) & Hm; gn0[h] = mh >> lhw;
        }
        int tap = 0, cb = 0, dy = -ph, dx = -pw;         // (tap, channel block) walk of conv-shaped K axes
        if (p.a_kind == GATHER_K) {
          tap = kt0 / p.cblocks; cb = kt0 - tap * p.cblocks;
          dy = tap / p.kw - ph; dx = tap - (tap / p.kw) * p.kw - pw;
        }
        int a_c[2 * MH], a_dx[2 * MH], a_dy[2 * MH], a_src[2 * MH];
#pragma unroll
        for (int j = 0; j < 2 * MH; ++j) { a_c[j] = 0; a_dx[j] = 0; a_dy[j] = 0; a_src[j] = 2; }
        if (p.a_kind == GATHER_MN) {
#pragma unroll
          for (int j = 0; j < 2 * MH; ++j) {
            const int mm = m0 + 64 * j;
            const int t = mm / p.Ct, c = mm - t * p.Ct;
            if (t >= p.ntaps) { a_src[j] = 2; }            // masked rows: any in-range box
            else {
              a_dy[j] = t / p.kw - ph; a_dx[j] = t - (t / p.kw) * p.kw - pw;
              a_src[j] = c < p.C1 ? 0 : 1; a_c[j] = c < p.C1 ? c : c - p.C1;
            }
          }
        }
        // gathered B operand (weight gradient with few output channels, not transposed): n = (tap, channel)
        // (a CTA of a pair only walks its own half of the 64-wide column chunks)
        constexpr int NBJ = BN / 64 / CG;
        const int bj0 = CG == 2 ? rank * NBJ : 0;
        int b_c[NBJ], b_dx[NBJ], b_dy[NBJ], b_src[NBJ];
#pragma unroll
        for (int j = 0; j < NBJ; ++j) { b_c[j] = 0; b_dx[j] = 0; b_dy[j] = 0; b_src[j] = 2; }
        if (p.b_kind == GATHER_MN) {
#pragma unroll
          for (int j = 0; j < NBJ; ++j) {
            const int nn = n0 + 64 * (bj0 + j);
            const int t = nn / p.Ct, c = nn - t * p.Ct;
            if (t < p.ntaps) {                              // else masked columns: any in-range box
              b_dy[j] = t / p.kw - ph; b_dx[j] = t - (t / p.kw) * p.kw - pw;
              b_src[j] = c < p.C1 ? 0 : 1; b_c[j] = c < p.C1 ? c : c - p.C1;
            }
          }
        }
        int kk = kt0 * BK;
        for (int i = 0; i < nkt; ++i) {
          mbar_wait(empty0 + 8 * s, par);
          // CG == 2: both CTAs arm and signal the pair leader's barrier (same offset, peer bit cleared)
          const uint32_t bar = CG == 2 ? ((full0 + 8 * s) & PEER_MASK) : (full0 + 8 * s);
          const uint32_t sa = smem_base + s * STAGE_BYTES, sb = sa + A_BYTES;
          if (elect_one()) {
          if constexpr (CG == 2) mbar_expect_tx_cluster(bar, do_a ? A_BYTES : B_STAGE_BYTES);
          else mbar_expect_tx(bar, do_a ? A_BYTES : B_STAGE_BYTES);
          if (do_a) {
          // ---- A (private to this CTA): MH blocks of 128 rows, 16 KB each
          if (p.a_kind == KMAJOR) {
#pragma unroll
            for (int h = 0; h < MH; ++h) ld3<CG>(sa + h * A_STAGE_BYTES, &mapA0, bar, kk, m0 + h * BM, b);
          } else if (p.a_kind == MNMAJOR) {
#pragma unroll
            for (int j = 0; j < 2 * MH; ++j) ld3<CG>(sa + j * 8192, &mapA0, bar, m0 + 64 * j, kk, b);
          } else if (p.a_kind == GATHER_K) {
#pragma unroll
            for (int h = 0; h < MH; ++h) {
              if (cb < p.c1blocks) ld4<CG>(sa + h * A_STAGE_BYTES, &mapA0, bar, cb * 64, gx0[h] + dx, gy0[h] + dy, gn0[h]);
              else ld4<CG>(sa + h * A_STAGE_BYTES, &mapA1, bar, (cb - p.c1blocks) * 64, gx0[h] + dx, gy0[h] + dy, gn0[h]);
            }
          } else {   // GATHER_MN as A: m = (tap, channel), k = pixel block
            const int x0 = kk & Wm, y0 = (kk >> lw) & Hm, i0 = kk >> lhw;
#pragma unroll
            for (int j = 0; j < 2 * MH; ++j) {
              if (a_src[j] == 2) ld4<CG>(sa + j * 8192, &mapA0, bar, 0, x0, y0, i0);
              else if (a_src[j] == 0) ld4<CG>(sa + j * 8192, &mapA0, bar, a_c[j], x0 + a_dx[j], y0 + a_dy[j], i0);
              else ld4<CG>(sa + j * 8192, &mapA1, bar, a_c[j], x0 + a_dx[j], y0 + a_dy[j], i0);
            }
          }
          } else if constexpr (CG == 2) {
          // ---- B, CTA pair: this CTA's half of the tile (BN/2 rows or BN/128 64-wide chunks) at local offset 0
          constexpr int HALF = BN / 2, PER = BN / 128;
          if (p.b_kind == KMAJOR) {
            ld3<2>(sb, &mapB0, bar, kk, n0 + rank * HALF, b);
          } else if (p.b_kind == GATHER_MN) {
            const int x0 = kk & Wm, y0 = (kk >> lw) & Hm, i0 = kk >> lhw;
#pragma unroll
            for (int jj = 0; jj < PER; ++jj) {
              if (b_src[jj] == 2) ld4<2>(sb + jj * 8192, &mapB0, bar, 0, x0, y0, i0);
              else if (b_src[jj] == 0) ld4<2>(sb + jj * 8192, &mapB0, bar, b_c[jj], x0 + b_dx[jj], y0 + b_dy[jj], i0);
              else ld4<2>(sb + jj * 8192, &mapB1, bar, b_c[jj], x0 + b_dx[jj], y0 + b_dy[jj], i0);
            }
          } else {
#pragma unroll
            for (int jj = 0; jj < PER; ++jj) {
              const int j = rank * PER + jj;
              if (p.b_kind == MNMAJOR) ld3<2>(sb + jj * 8192, &mapB0, bar, n0 + 64 * j, kk, b);
              else ld3<2>(sb + jj * 8192, &mapB0, bar, n0 + 64 * j, p.ntaps - 1 - tap, cb * 64);
            }
          }
          } else {
          // ---- B: this CTA's slice, multicast to the whole cluster
          if (p.b_kind == KMAJOR) {
            if (cs > 1) {
              const int rows = BN / cs;
              tma_load_3d_mc(sb + rank * rows * 128, &mapB0, bar, kk, n0 + rank * rows, b, mc_mask);
            } else {
              tma_load_3d(sb, &mapB0, bar, kk, n0, b);
            }
          } else if (p.b_kind == GATHER_MN) {
            const int x0 = kk & Wm, y0 = (kk >> lw) & Hm, i0 = kk >> lhw;
#pragma unroll
            for (int j = 0; j < BN / 64; ++j) {
              if (b_src[j] == 2) tma_load_4d(sb + j * 8192, &mapB0, bar, 0, x0, y0, i0);
              else if (b_src[j] == 0) tma_load_4d(sb + j * 8192, &mapB0, bar, b_c[j], x0 + b_dx[j], y0 + b_dy[j], i0);
              else tma_load_4d(sb + j * 8192, &mapB1, bar, b_c[j], x0 + b_dx[j], y0 + b_dy[j], i0);
            }
          } else {
            const int per = (BN / 64) / cs;       // host guarantees cs <= BN/64
            for (int jj = 0; jj < per; ++jj) {
              const int j = rank * per + jj;
              int c0, c1, c2;
              if (p.b_kind == MNMAJOR) { c0 = n0 + 64 * j; c1 = kk; c2 = b; }
              else { c0 = n0 + 64 * j; c1 = p.ntaps - 1 - tap; c2 = cb * 64; }
              if (cs > 1) tma_load_3d_mc(sb + j * 8192, &mapB0, bar, c0, c1, c2, mc_mask);
              else tma_load_3d(sb + j * 8192, &mapB0, bar, c0, c1, c2);
            }
          }
          }
          }
          __syncwarp();
          // ---- advance
          kk += BK;
          if (++cb == p.cblocks) { cb = 0; ++tap; if (++dx > pw) { dx = -pw; ++dy; } }
          if (++s == STAGES) { s = 0; par ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================================================== MMA issuer (whole warp walks the loop, one elected lane issues)
    // (CTA pair: the leader issues for both CTAs, the peer's warp 1 only owns its TMEM allocation)
    if (leader_cta) {
      const bool a_mn = (p.a_kind == MNMAJOR || p.a_kind == GATHER_MN);
      const bool b_mn = (p.b_kind != KMAJOR);
      // instruction descriptor: D = f32, A = B = bf16, operand majors, N >> 3, M >> 4 (M = 256 across a CTA pair)
      const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((a_mn ? 1u : 0u) << 15) | ((b_mn ? 1u : 0u) << 16) |
                             ((uint32_t)(BN >> 3) << 17) | ((uint32_t)((BM * CG) >> 4) << 24);
      // descriptor = hi32 (SBO=1024, version 1, SWIZZLE_128B) : lo32 (start>>4 | LBO>>4 << 16)
      const uint32_t desc_hi = (1024u >> 4) | (1u << 14) | (2u << 29);
      const uint32_t a_lo0 = (smem_base >> 4) | ((a_mn ? (8192u >> 4) : 1u) << 16);
      const uint32_t b_lo0 = ((smem_base + A_BYTES) >> 4) | ((b_mn ? (8192u >> 4) : 1u) << 16);
      const uint32_t a_step = a_mn ? (2048u >> 4) : (32u >> 4), b_step = b_mn ? (2048u >> 4) : (32u >> 4);
      int s = 0, tl = 0;
      uint32_t par = 0;
      uint32_t a_lo = a_lo0, b_lo = b_lo0;
      if constexpr (HALO) {
        // both operands K-major; per stage: three taps x four K=16 steps x MH row blocks
        const uint32_t hb_lo0 = ((smem_base + A_SLOT_BYTES) >> 4) | (1u << 16);
        const uint32_t dy_step = (uint32_t)(p.W * 128) >> 4;
        const int nsb = p.nk / 3;
        b_lo = hb_lo0;
        for (int st = cluster_id; st < total; st += n_clusters) {
          const int buf = tl & 1;
          if (tl >= 2) mbar_wait(tempty0 + 8 * buf, ((tl >> 1) - 1) & 1);
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          const uint32_t tacc = tmem_base + (uint32_t)(buf * TC);
          uint32_t accum = 0;
          for (int i = 0; i < nsb; ++i) {
            mbar_wait(full0 + 8 * s, par);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            if (elect_one()) {
#pragma unroll
              for (int dyi = 0; dyi < 3; ++dyi)
#pragma unroll
                for (int j = 0; j < BK / 16; ++j) {
                  const uint64_t bd = ((uint64_t)desc_hi << 32) | (uint64_t)(b_lo + dyi * (B_STAGE_BYTES >> 4) + j * 2);
#pragma unroll
                  for (int h = 0; h < MH; ++h) {
                    const uint64_t ad = ((uint64_t)desc_hi << 32) | (uint64_t)(a_lo + dyi * dy_step + h * (A_STAGE_BYTES >> 4) + j * 2);
                    if constexpr (CG == 2) umma_f16_2sm(tacc + (uint32_t)(h * BN), ad, bd, idesc, accum);
                    else umma_f16(tacc + (uint32_t)(h * BN), ad, bd, idesc, accum);
                  }
                  accum = 1;
                }
              if constexpr (CG == 2) umma_commit_2sm(empty0 + 8 * s, 3);
              else umma_commit(empty0 + 8 * s);
            }
            __syncwarp();
            accum = 1;
            a_lo += STAGE_BYTES >> 4;
            b_lo += STAGE_BYTES >> 4;
            if (++s == STAGES) { s = 0; par ^= 1; a_lo = a_lo0; b_lo = hb_lo0; }
          }
          if (elect_one()) {
            if constexpr (CG == 2) umma_commit_2sm(tfull0 + 8 * buf, 3);
            else umma_commit(tfull0 + 8 * buf);
          }
          __syncwarp();
          ++tl;
        }
      } else
      for (int st = cluster_id; st < total; st += n_clusters) {
        int m0, n0, b, kt0, nkt;
        decode(st, m0, n0, b, kt0, nkt);
        if (nkt <= 0) continue;
        const int buf = tl & 1;
        if (tl >= 2) mbar_wait(tempty0 + 8 * buf, ((tl >> 1) - 1) & 1);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t tacc = tmem_base + (uint32_t)(buf * TC);
        uint32_t accum = 0;
        for (int i = 0; i < nkt; ++i) {
          mbar_wait(full0 + 8 * s, par);
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          if (elect_one()) {
#pragma unroll
          for (int j = 0; j < BK / 16; ++j) {
            const uint64_t bd = ((uint64_t)desc_hi << 32) | (uint64_t)(b_lo + j * b_step);
#pragma unroll
            for (int h = 0; h < MH; ++h) {
              const uint64_t ad = ((uint64_t)desc_hi << 32) | (uint64_t)(a_lo + h * (A_STAGE_BYTES >> 4) + j * a_step);
              if constexpr (CG == 2) umma_f16_2sm(tacc + (uint32_t)(h * BN), ad, bd, idesc, accum);
              else umma_f16(tacc + (uint32_t)(h * BN), ad, bd, idesc, accum);
            }
            accum = 1;
          }
          if constexpr (CG == 2) umma_commit_2sm(empty0 + 8 * s, 3);      // frees the slot in both CTAs
          else if (cs > 1) umma_commit_mc(empty0 + 8 * s, mc_mask);
          else umma_commit(empty0 + 8 * s);
          }
          __syncwarp();
          accum = 1;
          a_lo += STAGE_BYTES >> 4;
          b_lo += STAGE_BYTES >> 4;
          if (++s == STAGES) { s = 0; par ^= 1; a_lo = a_lo0; b_lo = b_lo0; }
        }
        if (elect_one()) {
          if constexpr (CG == 2) umma_commit_2sm(tfull0 + 8 * buf, 3);        // accumulators of both CTAs complete
          else umma_commit(tfull0 + 8 * buf);
        }
        __syncwarp();
        ++tl;
      }
    }
  } else {
    // ===================================================== epilogue (warps 2..9)
    const int q = warp % 4;                   // TMEM lane quarter of this warp
    const int half = (warp - 2) / 4;          // MH == 1: which half of the tile's columns; MH == 2: which 128 rows
    const int et = (warp - 2) * 32 + lane;    // epilogue thread index (0..255): its private staging slots
    const int row = (MH == 2 ? half * BM : 0) + q * 32 + lane;
    const bool vec_res = p.residual && ((p.ldr | p.sRb) % 8 == 0) && ((reinterpret_cast<uintptr_t>(p.residual) & 15) == 0);
    const bool vec_bias = p.bias && ((reinterpret_cast<uintptr_t>(p.bias) & 15) == 0);
    uint8_t* my_stage = res_stage + (size_t)et * 16;
    int tl = 0;
    if (p.epi_tma) {
      // ---------------- TMA epilogue.  Each half-group (4 warps = 128 accumulator rows) owns half of the staging region:
      // its threads write their rows into 16 KB boxes (128 rows x 128 B, SWIZZLE_128B) and one leader thread stores the
      // boxes with cp.async.bulk.tensor (reduce-add for split-K) - full 128-byte lines instead of row-scattered 16-byte
      // stores, out-of-range rows / columns clipped by the tensor map.  The residual tile arrives in the same boxes by
      // TMA and is updated in place; bias + per-image bias of the tile are staged in shared memory one tile ahead.
      constexpr int HALF_BYTES = RES_BYTES / 2;
      constexpr int BOXES = HALF_BYTES >= 16384 ? HALF_BYTES / 16384 : 1;     // (the host never selects this path for TC < 128)
      const bool f32 = !p.out_bf16;
      const int slabs = f32 ? 2 : 1;
      const int cps = CH / slabs;                         // 32-column chunks per slab
      const uint32_t region = smem_u32(res_stage) + (uint32_t)(half * HALF_BYTES);
      const uint32_t rowaddr = region + (uint32_t)((q * 32 + lane) * 128);
      const uint32_t sw = (uint32_t)(lane & 7);
      const bool leader = ((warp - 2) % 4 == 0) && lane == 0;
      const uint32_t rfull = rfull0 + 8 * half;
      const int bar_id = 1 + half;
      const bool has_bias = p.bias || p.rowbias;
      const bool dz_mode = TC == 256 && p.dz_cst != nullptr;
      const bool res_mode = p.residual != nullptr || dz_mode;      // dz: the x tile arrives like a residual tile
      const int nb_vals = p.rb_rows * BN;                 // staged bias values per tile
      const int my_rb = (p.rowbias && p.rows_per_rb < BMT) ? row / p.rows_per_rb : 0;
      const int colbase = (MH == 2 ? 0 : half * (BN / 2));
      const int rowbase = (MH == 2 ? half * BM : 0);
      float pre[4] = {0.f, 0.f, 0.f, 0.f};
      auto load_bias = [&](int tm0, int tn0) {
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const int idx = et + k * 256;
          float v = 0.f;
          if (idx < nb_vals) {
            const int r = idx / BN, n = tn0 + (idx % BN);
            if (n < p.N) {
              if (p.bias) v = __ldg(p.bias + n);
              if (p.rowbias) {
                const long long img = (long long)(tm0 / p.rows_per_rb) + r;
                if (img * p.rows_per_rb < p.M) v += __ldg(p.rowbias + img * p.ld_rb + n);
              }
            }
          }
          pre[k] = v;
        }
      };
      auto issue_res = [&](int tm0, int tn0, int tb) {
        mbar_expect_tx(rfull, BOXES * 16384);
#pragma unroll
        for (int bx = 0; bx < BOXES; ++bx)
          tma_load_3d(region + bx * 16384, &mapR, rfull, tn0 + colbase + bx * 64, tm0 + rowbase, tb);
      };
      {
        int m0, n0, b, kt0, nkt;
        if (cluster_id < total) {
          decode(cluster_id, m0, n0, b, kt0, nkt);
          if (has_bias) load_bias(m0, n0);
          if (res_mode && leader) issue_res(m0, n0, b);
        }
      }
      for (int st = cluster_id; st < total; st += n_clusters) {
        int m0, n0, b, kt0, nkt;
        decode(st, m0, n0, b, kt0, nkt);
        if (nkt <= 0) continue;
        const int buf = tl & 1;
        if (has_bias) {
          named_bar(3, 256);                              // every warp has finished with the previous tile's values
#pragma unroll
          for (int k = 0; k < 4; ++k)
            if (et + k * 256 < nb_vals) sbias[et + k * 256] = pre[k];
          named_bar(3, 256);
          if (st + n_clusters < total) {
            int m2, n2, b2, k2, nk2;
            decode(st + n_clusters, m2, n2, b2, k2, nk2);
            load_bias(m2, n2);                            // lands while this tile is processed
          }
        }
        // dz: this row's dropout keep flags (1 bit per element, 16 bytes for its 128 columns) and its image's constants
        uint4 kw4 = make_uint4(0xffffffffu, 0xffffffffu, 0xffffffffu, 0xffffffffu);
        const float4* cst_row = nullptr;
        float* part_row = nullptr;
        if (dz_mode) {
          const long long grow = (long long)m0 + rowbase + q * 32 + lane;      // (the host guarantees M % 256 == 0, N % 128 == 0)
          const int ncol0 = n0 + colbase;
          const long long row0 = grow - lane;             // first row of this warp: one image per 32 rows (hw % 32 == 0)
          if (row0 < p.M) {                               // (the odd last tile of a CTA pair lies past M)
            if (p.dz_keep) kw4 = __ldg(reinterpret_cast<const uint4*>(p.dz_keep + (grow * p.N + ncol0) / 8));
            cst_row = p.dz_cst + (row0 >> p.dz_loghw) * p.N + ncol0;
            part_row = p.gn_part + ((row0 >> 5) * (p.N >> 2) + (ncol0 >> 2)) * 2;
          } else {
            cst_row = p.dz_cst + ncol0;
          }
        }
        if (res_mode) mbar_wait(rfull, tl & 1);
        mbar_wait(tfull0 + 8 * buf, (tl >> 1) & 1);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        for (int slab = 0; slab < slabs; ++slab) {
          if (!res_mode) {
            if (leader) bulk_wait_read();                 // the previous store has finished reading the boxes
            named_bar(bar_id, 128);
          }
#pragma unroll 1
          for (int cc = 0; cc < cps; ++cc) {
            const int c = slab * cps + cc;
            uint32_t r[32];
            __syncwarp();
            tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(buf * TC + half * (TC / 2) + c * 32), r);
            if (c == CH - 1) {                            // accumulator drained: hand the TMEM buffer back to the MMA warp
              asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
              __syncwarp();
              if (lane == 0) {
                if constexpr (CG == 2) mbar_arrive_cluster((tempty0 + 8 * buf) & PEER_MASK);
                else mbar_arrive(tempty0 + 8 * buf);
              }
            }
            float v[32];
#pragma unroll
            for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
            if (has_bias) {
              const float4* sb = reinterpret_cast<const float4*>(sbias + my_rb * BN + colbase + c * 32);
#pragma unroll
              for (int g = 0; g < 8; ++g) {
                const float4 t = sb[g];
                v[4 * g] += t.x; v[4 * g + 1] += t.y; v[4 * g + 2] += t.z; v[4 * g + 3] += t.w;
              }
            }
            if (!f32) {
              const uint32_t base = rowaddr + (uint32_t)((cc >> 1) * 16384);
              const uint32_t pb = (uint32_t)((cc & 1) * 4);
              if (dz_mode) {
                // ---- GroupNorm backward phase 1: dz = dy * keep * act'(u) replaces dy; quad sums of the stored values
                const uint32_t kword = cc == 0 ? kw4.x : cc == 1 ? kw4.y : cc == 2 ? kw4.z : kw4.w;
                float qs[16];
#pragma unroll
                for (int g = 0; g < 4; ++g) {
                  const uint32_t addr = base + (((pb + g) ^ sw) << 4);
                  uint4 t;
                  asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(t.x), "=r"(t.y), "=r"(t.z), "=r"(t.w) : "r"(addr));
                  const uint32_t xw[4] = {t.x, t.y, t.z, t.w};
                  const uint32_t kb = kword >> (8 * g);
                  const float4* ct = cst_row + cc * 32 + g * 8;
                  float dzv[8], gm[8], ub[8];
#pragma unroll
                  for (int e = 0; e < 8; ++e) {
                    const float4 k = __ldg(ct + e);
                    const float xf = __uint_as_float((e & 1) ? (xw[e >> 1] & 0xffff0000u) : (xw[e >> 1] << 16));
                    const float u = fmaf(xf, k.x, k.y);
                    float d = v[g * 8 + e] * (((kb >> e) & 1u) ? p.dz_ks : 0.f);
                    if (p.dz_act) d *= silu_grad_t<bf16>(u);
                    dzv[e] = d; gm[e] = k.z; ub[e] = u - k.w;
                  }
                  uint4 o;
                  __nv_bfloat162* ho = reinterpret_cast<__nv_bfloat162*>(&o);
#pragma unroll
                  for (int e = 0; e < 4; ++e) ho[e] = __floats2bfloat162_rn(dzv[2 * e], dzv[2 * e + 1]);
                  asm volatile("st.shared.v4.u32 [%0], {%1,%2,%3,%4};" ::"r"(addr), "r"(o.x), "r"(o.y), "r"(o.z), "r"(o.w) : "memory");
                  const uint32_t ow[4] = {o.x, o.y, o.z, o.w};
#pragma unroll
                  for (int h2 = 0; h2 < 2; ++h2) {
                    float s1 = 0.f, s2 = 0.f;
#pragma unroll
                    for (int e = 4 * h2; e < 4 * h2 + 4; ++e) {
                      const float r = __uint_as_float((e & 1) ? (ow[e >> 1] & 0xffff0000u) : (ow[e >> 1] << 16));
                      s1 = fmaf(gm[e], r, s1);
                      s2 = fmaf(r, ub[e], s2);
                    }
                    qs[(2 * g + h2) * 2] = s1;
                    qs[(2 * g + h2) * 2 + 1] = s2;
                  }
                }
                const float tot = warp_reduce16(qs, lane);
                if ((lane & 1) == 0 && part_row) part_row[cc * 16 + (lane >> 1)] = tot;
              } else
#pragma unroll
              for (int g = 0; g < 4; ++g) {
                const uint32_t addr = base + (((pb + g) ^ sw) << 4);
                if (res_mode) {
                  uint4 t;
                  asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(t.x), "=r"(t.y), "=r"(t.z), "=r"(t.w) : "r"(addr));
                  const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&t);
#pragma unroll
                  for (int e = 0; e < 4; ++e) {
                    v[g * 8 + 2 * e] += __low2float(h[e]);
                    v[g * 8 + 2 * e + 1] += __high2float(h[e]);
                  }
                }
                uint4 o;
                __nv_bfloat162* ho = reinterpret_cast<__nv_bfloat162*>(&o);
#pragma unroll
                for (int e = 0; e < 4; ++e)
                  ho[e] = __floats2bfloat162_rn(v[g * 8 + 2 * e] * p.alpha, v[g * 8 + 2 * e + 1] * p.alpha);
                asm volatile("st.shared.v4.u32 [%0], {%1,%2,%3,%4};" ::"r"(addr), "r"(o.x), "r"(o.y), "r"(o.z), "r"(o.w) : "memory");
              }
            } else {
              const uint32_t base = rowaddr + (uint32_t)(cc * 16384);
#pragma unroll
              for (int g = 0; g < 8; ++g) {
                const uint32_t addr = base + ((((uint32_t)g) ^ sw) << 4);
                asm volatile("st.shared.v4.f32 [%0], {%1,%2,%3,%4};" ::"r"(addr), "f"(v[4 * g] * p.alpha), "f"(v[4 * g + 1] * p.alpha),
                             "f"(v[4 * g + 2] * p.alpha), "f"(v[4 * g + 3] * p.alpha)
                             : "memory");
              }
            }
          }
          fence_async_smem();
          named_bar(bar_id, 128);
          if (leader) {
            const int cols_per_box = f32 ? 32 : 64;
            const int rowg = m0 + rowbase;
#pragma unroll
            for (int bx = 0; bx < BOXES; ++bx) {
              const int col = n0 + colbase + slab * cps * 32 + bx * cols_per_box;
              if (col < p.N && rowg < p.M) {
                if (p.accumulate) tma_reduce_add_3d(&mapC, region + bx * 16384, col, rowg, b);
                else tma_store_3d(&mapC, region + bx * 16384, col, rowg, b);
              }
            }
            bulk_commit();
          }
          if constexpr (TC == 256) {
            if (p.gn_part && !f32 && !dz_mode) {
              // GroupNorm statistics of the consumer as a by-product: sums over this half-group's 128 x 128 staged bf16
              // values (what the store writes), per 4-channel quad and per gn_rows rows.
              const int w4 = (warp - 2) & 3;
              const bool col_ok = n0 + colbase + w4 * 32 + (lane & 3) * 8 < p.N;
              const int gq = ((n0 + colbase) >> 2) + w4 * 8 + (lane & 3) * 2;
              if (p.gn_rows == 128) gn_quad_pass<1>(region, w4, lane, p.gn_part, p.N >> 2, gq, m0 + rowbase, p.M, col_ok);
              else if (p.gn_rows == 64) gn_quad_pass<2>(region, w4, lane, p.gn_part, p.N >> 2, gq, m0 + rowbase, p.M, col_ok);
              else gn_quad_pass<8>(region, w4, lane, p.gn_part, p.N >> 2, gq, m0 + rowbase, p.M, col_ok);
              named_bar(bar_id, 128);       // everyone has read the boxes before the leader lets them be refilled
            }
          }
        }
        if (res_mode && leader && st + n_clusters < total) {
          int m2, n2, b2, k2, nk2;
          decode(st + n_clusters, m2, n2, b2, k2, nk2);
          bulk_wait_read();                               // this tile's store has read the boxes: refill them
          issue_res(m2, n2, b2);
        }
        ++tl;
      }
      if (leader) bulk_wait_read();                       // the boxes must outlive the last store's reads of them
    } else
    for (int st = cluster_id; st < total; st += n_clusters) {
      int m0, n0, b, kt0, nkt;
      decode(st, m0, n0, b, kt0, nkt);
      if (nkt <= 0) continue;
      const int buf = tl & 1;
      const long long m = (long long)m0 + row;
      const bool row_ok = m < p.M;
      const int ncol0 = n0 + (MH == 2 ? 0 : half * (BN / 2));
      constexpr int NCOLS = TC / 2;             // columns handled by this warp
      const float* rb = (p.rowbias && row_ok) ? p.rowbias + (m / p.rows_per_rb) * p.ld_rb : nullptr;
      const bf16* res = (p.residual && row_ok) ? p.residual + (long long)b * p.sRb + m * p.ldr : nullptr;
      const long long crow = (long long)b * p.sCb + m * p.ldc;
      // ---- prefetch this tile's residual slice (overlaps the MMAs that are still filling the accumulator)
      bool staged = false;
      if (res && vec_res && ncol0 + NCOLS <= p.N) {
        staged = true;
#pragma unroll
        for (int j = 0; j < CH * 4; ++j) {
          const uint32_t dst = smem_u32(my_stage + (size_t)j * 256 * 16);
          asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(res + ncol0 + j * 8) : "memory");
        }
      }
      asm volatile("cp.async.commit_group;" ::: "memory");
      mbar_wait(tfull0 + 8 * buf, (tl >> 1) & 1);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      asm volatile("cp.async.wait_group 0;" ::: "memory");
#pragma unroll 1
      for (int c = 0; c < CH; ++c) {
        uint32_t r[32];
        __syncwarp();
        tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(buf * TC + half * (TC / 2) + c * 32), r);
        const int nb = ncol0 + c * 32;
        if (!row_ok || nb >= p.N) continue;
        float v[32];
#pragma unroll
        for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
        const bool full = (nb + 32 <= p.N);
        if (p.accumulate) {
          if (p.trans_out) {
            float* dst = reinterpret_cast<float*>(p.C) + (long long)b * p.sCb + m;     // C[n][m]: lanes are contiguous
#pragma unroll
            for (int i = 0; i < 32; ++i)
              if (full || nb + i < p.N) atomicAdd(dst + (long long)(nb + i) * p.ldc, p.alpha * v[i]);
          } else {
            float* dst = reinterpret_cast<float*>(p.C) + crow + nb;
#pragma unroll
            for (int i = 0; i < 32; ++i)
              if (full || nb + i < p.N) atomicAdd(dst + i, p.alpha * v[i]);
          }
          continue;
        }
        if (p.bias) {
          if (full && vec_bias) {
#pragma unroll
            for (int g = 0; g < 8; ++g) {
              const float4 t = __ldg(reinterpret_cast<const float4*>(p.bias + nb) + g);
              v[4 * g] += t.x; v[4 * g + 1] += t.y; v[4 * g + 2] += t.z; v[4 * g + 3] += t.w;
            }
          } else {
#pragma unroll
            for (int i = 0; i < 32; ++i)
              if (nb + i < p.N) v[i] += __ldg(p.bias + nb + i);
          }
        }
        if (rb) {
          if (full && ((reinterpret_cast<uintptr_t>(rb + nb) & 15) == 0)) {
#pragma unroll
            for (int g = 0; g < 8; ++g) {
              const float4 t = __ldg(reinterpret_cast<const float4*>(rb + nb) + g);
              v[4 * g] += t.x; v[4 * g + 1] += t.y; v[4 * g + 2] += t.z; v[4 * g + 3] += t.w;
            }
          } else {
#pragma unroll
            for (int i = 0; i < 32; ++i)
              if (nb + i < p.N) v[i] += __ldg(rb + nb + i);
          }
        }
        if (res) {
          if (staged) {
#pragma unroll
            for (int g = 0; g < 4; ++g) {
              const uint4 t = *reinterpret_cast<const uint4*>(my_stage + (size_t)(c * 4 + g) * 256 * 16);
              const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&t);
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                v[g * 8 + 2 * e] += __low2float(h[e]);
                v[g * 8 + 2 * e + 1] += __high2float(h[e]);
              }
            }
          } else {
#pragma unroll
            for (int i = 0; i < 32; ++i)
              if (nb + i < p.N) v[i] += __bfloat162float(res[nb + i]);
          }
        }
#pragma unroll
        for (int i = 0; i < 32; ++i) v[i] *= p.alpha;
        if (p.out_bf16) {
          bf16* dst = reinterpret_cast<bf16*>(p.C) + crow + nb;
          if (full && ((p.ldc | p.sCb) % 8 == 0)) {
#pragma unroll
            for (int g = 0; g < 4; ++g) {
              uint4 t;
              __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&t);
#pragma unroll
              for (int e = 0; e < 4; ++e) h[e] = __floats2bfloat162_rn(v[g * 8 + 2 * e], v[g * 8 + 2 * e + 1]);
              *reinterpret_cast<uint4*>(dst + g * 8) = t;
            }
          } else {
#pragma unroll
            for (int i = 0; i < 32; ++i)
              if (nb + i < p.N) dst[i] = __float2bfloat16_rn(v[i]);
          }
        } else {
          float* dst = reinterpret_cast<float*>(p.C) + crow + nb;
          if (full && ((p.ldc | p.sCb) % 4 == 0)) {
#pragma unroll
            for (int g = 0; g < 8; ++g)
              *reinterpret_cast<float4*>(dst + g * 4) = make_float4(v[g * 4], v[g * 4 + 1], v[g * 4 + 2], v[g * 4 + 3]);
          } else {
#pragma unroll
            for (int i = 0; i < 32; ++i)
              if (nb + i < p.N) dst[i] = v[i];
          }
        }
      }
      // this warp has drained its part of the accumulator buffer
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      __syncwarp();
      if (lane == 0) {
        if constexpr (CG == 2) mbar_arrive_cluster((tempty0 + 8 * buf) & PEER_MASK);
        else mbar_arrive(tempty0 + 8 * buf);
      }
      ++tl;
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (cs > 1) cluster_sync_all();     // no CTA may exit while peers can still multicast into it / read its operands
  if (warp == 1) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    if constexpr (CG == 2) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS));
    else asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS));
  }
}

// ------------------------------------------------------------------------------------ host side
PFN_cuTensorMapEncodeTiled_v12000 g_encode = nullptr;
int g_tc_state = -1;     // -1 unknown, 0 unavailable, 1 ok
std::mutex g_mu;

void tc_init() {
  std::lock_guard<std::mutex> lk(g_mu);
  if (g_tc_state >= 0) return;
  g_tc_state = 0;
  int dev = 0, major = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return;
  if (cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev) != cudaSuccess || major != 10) return;
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult qres;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres) != cudaSuccess || !fn ||
      qres != cudaDriverEntryPointSuccess)
    return;
  g_encode = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(fn);
  g_tc_state = 1;
}

// bf16 tensor map, SWIZZLE_128B, zero OOB fill.  dims/strides innermost first; strides[i] is the byte stride of dim i+1.
bool encode_map(CUtensorMap* map, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                const uint32_t* box, CUtensorMapDataType dtype = CU_TENSOR_MAP_DATA_TYPE_BFLOAT16) {
  cuuint64_t gd[5], gs[4];
  cuuint32_t bx[5], es[5];
  for (int i = 0; i < rank; ++i) { gd[i] = dims[i]; bx[i] = box[i]; es[i] = 1; }
  for (int i = 0; i + 1 < rank; ++i) gs[i] = strides_bytes[i];
  CUresult r = g_encode(map, dtype, (cuuint32_t)rank, const_cast<void*>(base), gd, gs, bx, es,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    st_set_error("cuTensorMapEncodeTiled failed (%d): rank %d dims %llu %llu %llu %llu box %u %u %u %u", (int)r, rank,
                 (unsigned long long)dims[0], (unsigned long long)(rank > 1 ? dims[1] : 0),
                 (unsigned long long)(rank > 2 ? dims[2] : 0), (unsigned long long)(rank > 3 ? dims[3] : 0), box[0],
                 rank > 1 ? box[1] : 0, rank > 2 ? box[2] : 0, rank > 3 ? box[3] : 0);
    return false;
  }
  return true;
}

bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }
bool pow2(int v) { return v > 0 && (v & (v - 1)) == 0; }

// pixel box (bw, bh, bn) covering `npix` consecutive pixels of an NHWC tensor
bool pixel_box(int npix, int H, int W, uint32_t* bw, uint32_t* bh, uint32_t* bn) {
  if (!pow2(H) || !pow2(W)) return false;
  int w = W < npix ? W : npix;
  int h = (npix / w) < H ? (npix / w) : H;
  int n = npix / (w * h);
  if (w * h * n != npix || n > 256) return false;
  *bw = w; *bh = h; *bn = n;
  return true;
}

int choose_bn(const st_gemm_args* a) {
  if (a->N <= 64) return 64;
  return 128;
}

template <int BN, int STAGES>
int launch(const CUtensorMap* maps, const TcParams& p, dim3 grid, cudaStream_t stream) {
  constexpr int smem = STAGES * (A_STAGE_BYTES + BN * 128) + (2 * STAGES + 1) * 8 + 16 + 1024;
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(gemm_tc_kernel<BN, STAGES>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) { st_set_error("st_gemm(tc): cudaFuncSetAttribute: %s", cudaGetErrorString(e)); return ST_ERR_CUDA; }
    configured = true;
  }
  gemm_tc_kernel<BN, STAGES><<<grid, NUM_THREADS, smem, stream>>>(maps[0], maps[1], maps[2], maps[3], p);
  ST_CHECK_LAUNCH("st_gemm(tcgen05)");
  return 0;
}

}  // namespace

extern "C" __attribute__((visibility("default"))) int st_tc_available(void) {
  tc_init();
  return g_tc_state == 1;
}

// 2-D bf16 SWIZZLE_128B tensor map over a row-major matrix (for the other tcgen05 kernels of the library: attn_tc.cu)
bool st_tc_encode_2d(CUtensorMap* map, const void* base, uint64_t inner, uint64_t rows, uint64_t row_stride_bytes,
                     uint32_t box_inner, uint32_t box_rows) {
  tc_init();
  if (g_tc_state != 1) { st_set_error("no sm_100 device / cuTensorMapEncodeTiled"); return false; }
  const uint64_t dims[2] = {inner, rows};
  const uint64_t str[1] = {row_stride_bytes};
  const uint32_t box[2] = {box_inner, box_rows};
  return encode_map(map, base, 2, dims, str, box);
}

// Why (or whether) the tcgen05 backend can run this problem.
int st_gemm_tc_supported(const st_gemm_args* a, const char** why) {
#define NO(msg) do { *why = msg; return 0; } while (0)
  if (!st_tc_available()) NO("no sm_100 device / cuTensorMapEncodeTiled");
  if (a->in_dtype != ST_BF16) NO("operands must be bf16");
  if (!aligned16(a->A) || !aligned16(a->B) || (a->A2 && !aligned16(a->A2)) || (a->B2 && !aligned16(a->B2))) NO("operand not 16-byte aligned");
  const int Ct = a->C1 + a->C2;
  if (a->a_mode == ST_OP_GATHER) {
    if (a->C1 % 64 || a->C2 % 64) NO("gather channels must be multiples of 64");
    uint32_t bw, bh, bn;
    if (!pixel_box(128, a->H, a->W, &bw, &bh, &bn)) NO("H/W must be powers of two");
    if (a->b_mode == ST_OP_STRIDED && (a->sBk != 1 || a->sBn % 8)) NO("conv weights must be K-major with ld % 8 == 0");
    if (a->b_mode == ST_OP_DGRADW && (a->N % 8)) NO("dgrad Cin must be a multiple of 8");
  } else {
    if (a->sAk == 1) { if (a->sAm % 8 || a->sAb % 8) NO("A leading dimension must be a multiple of 8"); }
    else if (a->sAm == 1) { if (a->sAk % 8 || a->sAb % 8) NO("A leading dimension must be a multiple of 8"); }
    else NO("A must be contiguous along m or k");
  }
  if (a->b_mode == ST_OP_GATHER) {
    if (a->C1 % 64 || a->C2 % 64) NO("gather channels must be multiples of 64");
    uint32_t bw, bh, bn;
    if (!pixel_box(64, a->H, a->W, &bw, &bh, &bn)) NO("H/W must be powers of two");
    if (a->sAm != 1) NO("wgrad A must be MN-major");
  } else if (a->b_mode == ST_OP_STRIDED) {
    if (a->sBk == 1) { if (a->sBn % 8 || a->sBb % 8) NO("B leading dimension must be a multiple of 8"); }
    else if (a->sBn == 1) { if (a->sBk % 8 || a->sBb % 8) NO("B leading dimension must be a multiple of 8"); }
    else NO("B must be contiguous along n or k");
  }
  (void)Ct;
  return 1;
#undef NO
}

// ---------------------------------------------------------------- persistent kernel: plan + launch
namespace {

int env_int(const char* name, int dflt) {
  const char* v = getenv(name);
  return v ? atoi(v) : dflt;
}

template <int BN, int MH, int STAGES, int CG, bool HALO = false>
int launch2(const CUtensorMap* maps, TcParams& p, int total_super, cudaStream_t stream) {
  constexpr int stage = HALO ? (A_STAGE_BYTES * MH * 3 / 2 + 3 * BN * 128 / CG) : (A_STAGE_BYTES * MH + BN * 128 / CG);
  constexpr int need = STAGES * stage + BN * MH * 256 + (2 * STAGES + 6) * 8 + 16 + 4 * BN * 4;
  // 1 KB of slack for re-aligning the dynamic shared-memory base to 1024 bytes (SWIZZLE_128B) - dropped when the
  // configuration only fits without it: the base is 1024-aligned in practice (the kernel traps if it is not)
  constexpr int smem = need + 1024 <= 232448 ? need + 1024 : need;
  static_assert(smem <= 232448, "shared memory budget");
  static bool configured = false;
  static int max_clusters[5] = {0, 0, 0, 0, 0};
  auto kern = gemm_tc2_kernel<BN, MH, STAGES, CG, HALO>;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) { st_set_error("st_gemm(tc2): cudaFuncSetAttribute: %s", cudaGetErrorString(e)); return ST_ERR_CUDA; }
    configured = true;
  }
  const int cs = CG == 2 ? 2 : p.cluster;
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.blockDim = dim3(NUM_THREADS2);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = cs;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  if (max_clusters[cs] == 0) {
    int n = 0;
    cfg.gridDim = dim3(cs * 64);
    cudaError_t e = cudaOccupancyMaxActiveClusters(&n, kern, &cfg);
    if (e != cudaSuccess || n <= 0) { cudaGetLastError(); n = st_num_sms() / cs; }
    max_clusters[cs] = n;
  }
  int clusters = total_super < max_clusters[cs] ? total_super : max_clusters[cs];
  cfg.gridDim = dim3(clusters * cs);
  const int pdl_pairs = env_int("ST_TC_PDL2", 1);
  cfg.numAttrs = ((cs == 1 || (CG == 2 && pdl_pairs)) && st_pdl_on(stream)) ? 2 : 1;
  cudaError_t e = cudaLaunchKernelEx(&cfg, kern, maps[0], maps[1], maps[2], maps[3], maps[4], maps[5], p);
  if (e != cudaSuccess) { st_set_error("st_gemm(tc2): launch failed: %s", cudaGetErrorString(e)); return ST_ERR_CUDA; }
  return 0;
}

bool gather_maps(CUtensorMap* m0, CUtensorMap* m1, const void* p1, const void* p2, int C1, int C2, int W, int H, int n_img,
                 int npix) {
  uint32_t bw, bh, bn;
  pixel_box(npix, H, W, &bw, &bh, &bn);
  const uint32_t box[4] = {64, bw, bh, bn};
  {
    const uint64_t dims[4] = {(uint64_t)C1, (uint64_t)W, (uint64_t)H, (uint64_t)n_img};
    const uint64_t str[3] = {(uint64_t)C1 * 2, (uint64_t)W * C1 * 2, (uint64_t)H * W * C1 * 2};
    if (!encode_map(m0, p1, 4, dims, str, box)) return false;
  }
  if (C2 > 0) {
    const uint64_t dims[4] = {(uint64_t)C2, (uint64_t)W, (uint64_t)H, (uint64_t)n_img};
    const uint64_t str[3] = {(uint64_t)C2 * 2, (uint64_t)W * C2 * 2, (uint64_t)H * W * C2 * 2};
    if (!encode_map(m1, p2, 4, dims, str, box)) return false;
  }
  return true;
}

int st_gemm_tc2(const st_gemm_args* a, cudaStream_t stream) {
  TcParams p;
  memset(&p, 0, sizeof(p));
  CUtensorMap maps[6];
  memset(maps, 0, sizeof(maps));
  const int Ct = a->C1 + a->C2;
  // Weight gradients run transposed (M' = taps*Cin, N' = Cout) unless the layer has <= 128 output channels: then the
  // transposed form would be stuck with 128-wide MMAs and 128x128 tiles at the shared-memory read limit, and the
  // direct form (A = dY MN-major, B = shifted NHWC boxes, 128x256 tiles, TMA reduce-add epilogue) is used instead.
  const bool wgrad_any = (a->b_mode == ST_OP_GATHER);
  const bool wgrad_nt = wgrad_any && a->M <= 128 && a->N >= 256 && a->accumulate && a->out_dtype == ST_F32 &&
                        env_int("ST_TC_WGRAD_NT", 1) == 1 && env_int("ST_TC_EPI", 1) == 1 &&
                        aligned16(a->C) && (a->sCm * 4) % 16 == 0;
  const bool wgrad = wgrad_any && !wgrad_nt;
  const int M = wgrad ? a->N : a->M, N = wgrad ? a->M : a->N;
  p.M = M; p.N = N; p.batch = a->batch; p.split_k = a->split_k > 1 ? a->split_k : 1;
  p.H = a->H; p.W = a->W; p.kh = a->kh; p.kw = a->kw; p.ntaps = a->kh * a->kw; p.Ct = Ct; p.C1 = a->C1;
  p.nk = (a->K + BK - 1) / BK;
  p.cblocks = Ct > 0 ? Ct / 64 : 1;
  p.c1blocks = a->C1 / 64;
  for (p.logW = 0; (1 << p.logW) < a->W; ++p.logW) {}
  for (p.logH = 0; (1 << p.logH) < a->H; ++p.logH) {}
  int BN = (N >= 256 && (N % 256 == 0 || wgrad_nt)) ? 256 : (N > 64 ? 128 : 64);
  if (BN == 256 && env_int("ST_TC_BN", 256) == 128) BN = 128;
  // too few 128x256 tiles to occupy the SMs (4x4 level of the U-Net): halve the tile instead of idling half the GPU
  // a dz request (GroupNorm backward phase 1 in the epilogue) needs 256-column accumulators: keep / form them
  const bool want_dz = a->dz_x && a->dz_cst && a->gn_part && !wgrad_any && a->N % 128 == 0 && a->M % 256 == 0 &&
                       env_int("ST_TC_GN_DZ", 1) == 1;
  if (BN == 256 && !wgrad && !want_dz && env_int("ST_TC_SMALL", 1) == 1) {
    const long long tiles256 = (long long)((M + BM - 1) / BM) * (N / 256) * p.batch * p.split_k;
    if (tiles256 * 10 < (long long)st_num_sms() * 7) BN = 128;
  }
  // 256 x 128 tiles (two accumulators sharing the B tile) when N fits 128-wide tiles and M is large
  int MH = 1;
  if (BN == 128 && env_int("ST_TC_MH", 2) == 2) {
    const long long tiles2 = (long long)((M + 255) / 256) * ((N + BN - 1) / BN) * p.batch * p.split_k;
    // weight gradients with a short M' = taps*Cin axis keep 128-row tiles (4 gathered A boxes per K block would
    // make the A producer the bottleneck, and 1152 rows quantise badly into 256-row tiles)
    if ((tiles2 >= 120 || want_dz) && !(wgrad && M < env_int("ST_TC_WGRAD_MH_MIN", 2048))) MH = 2;
  }
  p.m_tiles = (M + BM * MH - 1) / (BM * MH);
  p.n_tiles = (N + BN - 1) / BN;
  if (wgrad_nt && p.split_k > 1) {
    // the caller sized split-K for 128x128 tiles: re-derive it so that the work items fill two waves of the SMs
    const int tiles = p.m_tiles * p.n_tiles * p.batch;
    int sk = (2 * st_num_sms()) / (tiles > 0 ? tiles : 1);
    const int cap = p.nk / 16 > 0 ? p.nk / 16 : 1;
    p.split_k = sk < 1 ? 1 : (sk > cap ? cap : sk);
  }

  // B kind first (it bounds the cluster size)
  const bool b_kmajor = !wgrad && a->b_mode == ST_OP_STRIDED && a->sBk == 1;
  // (round 1's B-tile multicast across 2 / 4 single-MMA CTAs is superseded by the CTA-pair form and no longer maintained:
  // with the round-2 barrier counts it produces wrong tiles - the ST_TC_CLUSTER switch is therefore gone)
  int cs = 1;
  while (cs > 1 && cs > p.m_tiles) cs >>= 1;
  if (!b_kmajor) while (cs > 1 && cs > BN / 64) cs >>= 1;
  // CTA pairs (tcgen05.mma.cta_group::2): the big tiles, whenever there are at least two row blocks to pair
  // (ST_TC_CG=1 keeps every launch on single-CTA instructions; bit 0/1 of ST_TC_CG2_MASK select the 128x256 / 256x128 forms)
  int CG = 1;
  {
    const int want = env_int("ST_TC_CG", 2), mask = env_int("ST_TC_CG2_MASK", 3);
    // bit 2: transposed weight gradients of <= 128-output-channel layers as 256 x 128 pair tiles (with ST_TC_WGRAD_NT=0)
    const bool big = (BN == 256 && MH == 1 && (mask & 1)) || (BN == 128 && MH == 2 && (mask & 2)) ||
                     (BN == 128 && MH == 1 && wgrad && (mask & 4));
    if (want == 2 && big && p.m_tiles >= 2) { CG = 2; cs = 2; }
  }
  p.cluster = cs;
  bool halo = false;

  // ---------------- A
  if (wgrad) {
    p.a_kind = GATHER_MN;
    if (!gather_maps(&maps[0], &maps[1], a->B, a->B2, a->C1, a->C2, a->W, a->H, a->n_img, 64)) return ST_ERR_CUDA;
  } else if (a->a_mode == ST_OP_GATHER) {
    p.a_kind = GATHER_K;
    // halo form (see the kernel): 3x3, K-major weights, CTA pair, the row tile inside one image and >= 4 image rows tall
    const int BMT = BM * MH;
    halo = CG == 2 && !wgrad_any && a->kh == 3 && a->kw == 3 && b_kmajor && p.split_k == 1 && a->batch == 1 && a->W >= 8 &&
           a->W * 4 <= BMT && ((long long)a->H * a->W) % BMT == 0 && env_int("ST_TC_HALO", 1) == 1;
    if (halo) {
      const uint32_t box[4] = {64, (uint32_t)a->W, (uint32_t)(BMT / a->W + 2), 1};
      const void* src[2] = {a->A, a->A2};
      const int Cs[2] = {a->C1, a->C2};
      for (int k = 0; k < 2; ++k) {
        if (Cs[k] <= 0) continue;
        const uint64_t dims[4] = {(uint64_t)Cs[k], (uint64_t)a->W, (uint64_t)a->H, (uint64_t)a->n_img};
        const uint64_t str[3] = {(uint64_t)Cs[k] * 2, (uint64_t)a->W * Cs[k] * 2, (uint64_t)a->H * a->W * Cs[k] * 2};
        if (!encode_map(&maps[k], src[k], 4, dims, str, box)) return ST_ERR_CUDA;
      }
    } else if (!gather_maps(&maps[0], &maps[1], a->A, a->A2, a->C1, a->C2, a->W, a->H, a->n_img, 128)) return ST_ERR_CUDA;
  } else if (a->sAk == 1) {
    p.a_kind = KMAJOR;
    const uint64_t dims[3] = {(uint64_t)a->K, (uint64_t)a->M, (uint64_t)a->batch};
    const uint64_t str[2] = {(uint64_t)a->sAm * 2, (uint64_t)(a->batch > 1 ? a->sAb : (int64_t)a->M * a->sAm) * 2};
    const uint32_t box[3] = {64, 128, 1};
    if (!encode_map(&maps[0], a->A, 3, dims, str, box)) return ST_ERR_CUDA;
  } else {
    p.a_kind = MNMAJOR;
    const uint64_t dims[3] = {(uint64_t)a->M, (uint64_t)a->K, (uint64_t)a->batch};
    const uint64_t str[2] = {(uint64_t)a->sAk * 2, (uint64_t)(a->batch > 1 ? a->sAb : (int64_t)a->K * a->sAk) * 2};
    const uint32_t box[3] = {64, 64, 1};
    if (!encode_map(&maps[0], a->A, 3, dims, str, box)) return ST_ERR_CUDA;
  }
  // ---------------- B
  if (wgrad_nt) {
    p.b_kind = GATHER_MN;
    if (!gather_maps(&maps[2], &maps[3], a->B, a->B2, a->C1, a->C2, a->W, a->H, a->n_img, 64)) return ST_ERR_CUDA;
  } else if (wgrad) {
    p.b_kind = MNMAJOR;        // dY[pixel][co]: n' = co contiguous, k = pixel
    const uint64_t dims[3] = {(uint64_t)a->M, (uint64_t)a->K, 1};
    const uint64_t str[2] = {(uint64_t)a->sAk * 2, (uint64_t)a->K * a->sAk * 2};
    const uint32_t box[3] = {64, 64, 1};
    if (!encode_map(&maps[2], a->A, 3, dims, str, box)) return ST_ERR_CUDA;
  } else if (a->b_mode == ST_OP_DGRADW) {
    p.b_kind = DGRADW;
    const uint64_t dims[3] = {(uint64_t)a->N, (uint64_t)p.ntaps, (uint64_t)Ct};
    const uint64_t str[2] = {(uint64_t)a->N * 2, (uint64_t)p.ntaps * a->N * 2};
    const uint32_t box[3] = {64, 1, 64};
    if (!encode_map(&maps[2], a->B, 3, dims, str, box)) return ST_ERR_CUDA;
  } else if (a->sBk == 1) {
    p.b_kind = KMAJOR;
    const uint64_t dims[3] = {(uint64_t)a->K, (uint64_t)a->N, (uint64_t)a->batch};
    const uint64_t str[2] = {(uint64_t)a->sBn * 2, (uint64_t)(a->batch > 1 ? a->sBb : (int64_t)a->N * a->sBn) * 2};
    const uint32_t box[3] = {64, (uint32_t)(BN / cs), 1};
    if (!encode_map(&maps[2], a->B, 3, dims, str, box)) return ST_ERR_CUDA;
  } else {
    p.b_kind = MNMAJOR;
    const uint64_t dims[3] = {(uint64_t)a->N, (uint64_t)a->K, (uint64_t)a->batch};
    const uint64_t str[2] = {(uint64_t)a->sBk * 2, (uint64_t)(a->batch > 1 ? a->sBb : (int64_t)a->K * a->sBk) * 2};
    const uint32_t box[3] = {64, 64, 1};
    if (!encode_map(&maps[2], a->B, 3, dims, str, box)) return ST_ERR_CUDA;
  }

  p.C = a->C; p.ldc = a->sCm; p.sCb = a->sCb; p.out_bf16 = a->out_dtype == ST_BF16; p.accumulate = a->accumulate;
  p.trans_out = wgrad ? 1 : 0;
  p.bias = a->bias; p.rowbias = a->rowbias; p.rows_per_rb = a->rows_per_rb > 0 ? a->rows_per_rb : 1; p.ld_rb = a->ld_rb;
  p.residual = reinterpret_cast<const bf16*>(a->residual); p.ldr = a->sRm; p.sRb = a->sRb; p.alpha = a->alpha;

  // ---------------- epilogue through shared memory + TMA store (else: per-thread global stores / atomics)
  {
    const int BMT = BM * MH;
    const int es = p.out_bf16 ? 2 : 4;
    const bool has_vec = a->bias || a->rowbias;
    bool ok = env_int("ST_TC_EPI", 1) == 1 && !wgrad && BN * MH >= 128;
    ok = ok && aligned16(a->C) && (a->sCm * es) % 16 == 0 && (a->batch == 1 || (a->sCb * es) % 16 == 0);
    ok = ok && !(a->accumulate && p.out_bf16);
    ok = ok && !((a->residual || has_vec) && p.split_k > 1);
    if (a->residual)
      ok = ok && p.out_bf16 && aligned16(a->residual) && a->sRm % 8 == 0 && (a->batch == 1 || a->sRb % 8 == 0);
    // GroupNorm backward phase 1 in the epilogue (st_gemm_args::dz_x): needs the TMA epilogue over 256-column accumulators
    const bool dz = ok && a->dz_x && a->dz_cst && a->gn_part && p.out_bf16 && !a->residual && !has_vec && !a->accumulate &&
                    p.split_k == 1 && a->batch == 1 && BN * MH == 256 && a->N % 128 == 0 && a->M % 256 == 0 &&
                    a->gn_hw >= 32 && (a->gn_hw & (a->gn_hw - 1)) == 0 && a->M % a->gn_hw == 0 && aligned16(a->dz_x) &&
                    a->dz_ldx % 8 == 0 && aligned16(a->dz_cst) && (!a->dz_keep || aligned16(a->dz_keep)) &&
                    env_int("ST_TC_GN_DZ", 1) == 1;
    int rb_rows = 1;
    if (a->rowbias) {
      if (p.rows_per_rb % BMT == 0) rb_rows = 1;
      else if (BMT % p.rows_per_rb == 0 && BMT / p.rows_per_rb <= 4) rb_rows = BMT / p.rows_per_rb;
      else ok = false;
    }
    if (ok) {
      const uint64_t dims[3] = {(uint64_t)a->N, (uint64_t)a->M, (uint64_t)a->batch};
      const uint64_t str[2] = {(uint64_t)a->sCm * es, (uint64_t)(a->batch > 1 ? a->sCb : (int64_t)a->M * a->sCm) * es};
      const uint32_t box[3] = {(uint32_t)(p.out_bf16 ? 64 : 32), 128, 1};
      if (!encode_map(&maps[4], a->C, 3, dims, str, box, p.out_bf16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32))
        return ST_ERR_CUDA;
      if (a->residual) {
        const uint64_t rstr[2] = {(uint64_t)a->sRm * 2, (uint64_t)(a->batch > 1 ? a->sRb : (int64_t)a->M * a->sRm) * 2};
        const uint32_t rbox[3] = {64, 128, 1};
        if (!encode_map(&maps[5], a->residual, 3, dims, rstr, rbox)) return ST_ERR_CUDA;
      }
      if (dz) {
        const uint64_t rstr[2] = {(uint64_t)a->dz_ldx * 2, (uint64_t)a->M * a->dz_ldx * 2};
        const uint32_t rbox[3] = {64, 128, 1};
        if (!encode_map(&maps[5], a->dz_x, 3, dims, rstr, rbox)) return ST_ERR_CUDA;
        p.dz_cst = reinterpret_cast<const float4*>(a->dz_cst);
        p.dz_keep = a->dz_keep;
        p.dz_ks = a->alpha * (a->dz_inv_keep > 0.f ? a->dz_inv_keep : 1.f);
        p.dz_act = a->dz_act;
        for (p.dz_loghw = 0; (1 << p.dz_loghw) < a->gn_hw; ++p.dz_loghw) {}
        p.gn_part = a->gn_part;
        p.gn_rows = 32;
      }
      p.epi_tma = 1;
      p.rb_rows = rb_rows;
    }
  }
  // ---------------- GroupNorm partial sums as a by-product (see TcParams::gn_part)
  if (a->gn_rows_out) *a->gn_rows_out = 0;
  if (p.dz_cst) {
    if (a->gn_rows_out) *a->gn_rows_out = 32;
  } else if (!a->dz_x && a->gn_part && a->gn_hw >= 16 && p.epi_tma && p.out_bf16 && BN * MH == 256 && p.split_k == 1 && a->batch == 1 &&
      a->N % 128 == 0 && a->M % a->gn_hw == 0 && (a->gn_hw & (a->gn_hw - 1)) == 0 && !a->accumulate &&
      env_int("ST_TC_GN_STATS", 1) == 1) {
    p.gn_part = a->gn_part;
    p.gn_rows = a->gn_hw < 128 ? a->gn_hw : 128;
    if (p.gn_rows != 16 && p.gn_rows != 64 && p.gn_rows != 128) p.gn_part = nullptr;
    else if (a->gn_rows_out) *a->gn_rows_out = p.gn_rows;
  }

  const int m_groups = (p.m_tiles + cs - 1) / cs;
  const long long total = (long long)p.batch * p.split_k * p.n_tiles * m_groups;
  ST_CHECK_ARG(total < (1LL << 30), "st_gemm(tc2): too many tiles");
  if (halo && BN == 256) return launch2<256, 1, 2, 2, true>(maps, p, (int)total, stream);
  if (halo) return launch2<128, 2, 2, 2, true>(maps, p, (int)total, stream);
  if (CG == 2 && BN == 256) return launch2<256, 1, 4, 2>(maps, p, (int)total, stream);
  if (CG == 2 && MH == 2) return launch2<128, 2, 4, 2>(maps, p, (int)total, stream);
  if (CG == 2) return launch2<128, 1, 6, 2>(maps, p, (int)total, stream);
  if (BN == 256) return launch2<256, 1, 3, 1>(maps, p, (int)total, stream);
  if (BN == 128 && MH == 2) return launch2<128, 2, 3, 1>(maps, p, (int)total, stream);
  if (BN == 128) return launch2<128, 1, 5, 1>(maps, p, (int)total, stream);
  return launch2<64, 1, 8, 1>(maps, p, (int)total, stream);
}

}  // namespace

int st_gemm_tc1(const st_gemm_args* a, cudaStream_t stream);

int st_gemm_tc(const st_gemm_args* a, cudaStream_t stream) {
  // ST_TC_VARIANT=1 selects the first-generation kernel (one tile per CTA, two CTAs per SM), kept as an on-device
  // cross-check; the persistent kernel is faster on every shape of the workload (tools/gemm_bench.py, profiles/).
  if (env_int("ST_TC_VARIANT", 2) == 1) return st_gemm_tc1(a, stream);
  return st_gemm_tc2(a, stream);
}

int st_gemm_tc1(const st_gemm_args* a, cudaStream_t stream) {
  TcParams p;
  memset(&p, 0, sizeof(p));
  CUtensorMap maps[4];
  memset(maps, 0, sizeof(maps));
  const int BN = choose_bn(a);
  const int Ct = a->C1 + a->C2;
  p.M = a->M; p.N = a->N; p.batch = a->batch; p.split_k = a->split_k > 1 ? a->split_k : 1;
  p.H = a->H; p.W = a->W; p.kh = a->kh; p.kw = a->kw; p.ntaps = a->kh * a->kw; p.Ct = Ct; p.C1 = a->C1;
  p.nk = (a->K + BK - 1) / BK;

  // ---------------- A
  if (a->a_mode == ST_OP_GATHER) {
    p.a_kind = GATHER_K;
    p.cblocks = Ct / 64;
    p.c1blocks = a->C1 / 64;
    uint32_t bw, bh, bn;
    pixel_box(128, a->H, a->W, &bw, &bh, &bn);
    const uint32_t box[4] = {64, bw, bh, bn};
    {
      const uint64_t dims[4] = {(uint64_t)a->C1, (uint64_t)a->W, (uint64_t)a->H, (uint64_t)a->n_img};
      const uint64_t str[3] = {(uint64_t)a->C1 * 2, (uint64_t)a->W * a->C1 * 2, (uint64_t)a->H * a->W * a->C1 * 2};
      if (!encode_map(&maps[0], a->A, 4, dims, str, box)) return ST_ERR_CUDA;
    }
    if (a->C2 > 0) {
      const uint64_t dims[4] = {(uint64_t)a->C2, (uint64_t)a->W, (uint64_t)a->H, (uint64_t)a->n_img};
      const uint64_t str[3] = {(uint64_t)a->C2 * 2, (uint64_t)a->W * a->C2 * 2, (uint64_t)a->H * a->W * a->C2 * 2};
      if (!encode_map(&maps[1], a->A2, 4, dims, str, box)) return ST_ERR_CUDA;
    }
  } else if (a->sAk == 1) {
    p.a_kind = KMAJOR;
    const uint64_t dims[3] = {(uint64_t)a->K, (uint64_t)a->M, (uint64_t)a->batch};
    const uint64_t str[2] = {(uint64_t)a->sAm * 2, (uint64_t)(a->batch > 1 ? a->sAb : (int64_t)a->M * a->sAm) * 2};
    const uint32_t box[3] = {64, 128, 1};
    if (!encode_map(&maps[0], a->A, 3, dims, str, box)) return ST_ERR_CUDA;
  } else {
    p.a_kind = MNMAJOR;
    const uint64_t dims[3] = {(uint64_t)a->M, (uint64_t)a->K, (uint64_t)a->batch};
    const uint64_t str[2] = {(uint64_t)a->sAk * 2, (uint64_t)(a->batch > 1 ? a->sAb : (int64_t)a->K * a->sAk) * 2};
    const uint32_t box[3] = {64, 64, 1};
    if (!encode_map(&maps[0], a->A, 3, dims, str, box)) return ST_ERR_CUDA;
  }
  // ---------------- B
  if (a->b_mode == ST_OP_GATHER) {
    p.b_kind = GATHER_MN;
    uint32_t bw, bh, bn;
    pixel_box(64, a->H, a->W, &bw, &bh, &bn);
    const uint32_t box[4] = {64, bw, bh, bn};
    {
      const uint64_t dims[4] = {(uint64_t)a->C1, (uint64_t)a->W, (uint64_t)a->H, (uint64_t)a->n_img};
      const uint64_t str[3] = {(uint64_t)a->C1 * 2, (uint64_t)a->W * a->C1 * 2, (uint64_t)a->H * a->W * a->C1 * 2};
      if (!encode_map(&maps[2], a->B, 4, dims, str, box)) return ST_ERR_CUDA;
    }
    if (a->C2 > 0) {
      const uint64_t dims[4] = {(uint64_t)a->C2, (uint64_t)a->W, (uint64_t)a->H, (uint64_t)a->n_img};
      const uint64_t str[3] = {(uint64_t)a->C2 * 2, (uint64_t)a->W * a->C2 * 2, (uint64_t)a->H * a->W * a->C2 * 2};
      if (!encode_map(&maps[3], a->B2, 4, dims, str, box)) return ST_ERR_CUDA;
    }
  } else if (a->b_mode == ST_OP_DGRADW) {
    p.b_kind = DGRADW;      // W[co][tap][ci], ci = N contiguous
    const uint64_t dims[3] = {(uint64_t)a->N, (uint64_t)p.ntaps, (uint64_t)Ct};
    const uint64_t str[2] = {(uint64_t)a->N * 2, (uint64_t)p.ntaps * a->N * 2};
    const uint32_t box[3] = {64, 1, 64};
    if (!encode_map(&maps[2], a->B, 3, dims, str, box)) return ST_ERR_CUDA;
  } else if (a->sBk == 1) {
    p.b_kind = KMAJOR;
    const uint64_t dims[3] = {(uint64_t)a->K, (uint64_t)a->N, (uint64_t)a->batch};
    const uint64_t str[2] = {(uint64_t)a->sBn * 2, (uint64_t)(a->batch > 1 ? a->sBb : (int64_t)a->N * a->sBn) * 2};
    const uint32_t box[3] = {64, (uint32_t)BN, 1};
    if (!encode_map(&maps[2], a->B, 3, dims, str, box)) return ST_ERR_CUDA;
  } else {
    p.b_kind = MNMAJOR;
    const uint64_t dims[3] = {(uint64_t)a->N, (uint64_t)a->K, (uint64_t)a->batch};
    const uint64_t str[2] = {(uint64_t)a->sBk * 2, (uint64_t)(a->batch > 1 ? a->sBb : (int64_t)a->K * a->sBk) * 2};
    const uint32_t box[3] = {64, 64, 1};
    if (!encode_map(&maps[2], a->B, 3, dims, str, box)) return ST_ERR_CUDA;
  }
  if (a->a_mode == ST_OP_GATHER && a->b_mode == ST_OP_DGRADW) p.cblocks = Ct / 64;

  p.C = a->C; p.ldc = a->sCm; p.sCb = a->sCb; p.out_bf16 = a->out_dtype == ST_BF16; p.accumulate = a->accumulate;
  p.bias = a->bias; p.rowbias = a->rowbias; p.rows_per_rb = a->rows_per_rb > 0 ? a->rows_per_rb : 1; p.ld_rb = a->ld_rb;
  p.residual = reinterpret_cast<const bf16*>(a->residual); p.ldr = a->sRm; p.sRb = a->sRb; p.alpha = a->alpha;

  dim3 grid((a->M + BM - 1) / BM, (a->N + BN - 1) / BN, a->batch * p.split_k);
  ST_CHECK_ARG(grid.y <= 65535 && grid.z <= 65535, "st_gemm(tc): grid too large");
  if (BN == 64) return launch<64, 4>(maps, p, grid, stream);
  return launch<128, 3>(maps, p, grid, stream);
}
