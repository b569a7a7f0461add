// Generic fp32-accumulate SIMT implicit GEMM.  This is the "parity" backend (fp32 storage, exact
// fp32 FMA accumulation) and the on-device checker for the tcgen05 backend; it accepts every
// addressing mode of st_gemm_args.  It is not the fast path.
#include "common.cuh"

namespace {

template <typename T>
struct Gather {
  const T* x1;
  const T* x2;
  int H, W, C1, C2, kh, kw;
  __device__ __forceinline__ float load(long long pixel, int tap, int c, long long n_pix) const {
    if (pixel >= n_pix) return 0.f;
    int x = (int)(pixel % W);
    long long t = pixel / W;
    int y = (int)(t % H);
    long long img = t / H;
    int yy = y + tap / kw - (kh - 1) / 2;
    int xx = x + tap % kw - (kw - 1) / 2;
    if (yy < 0 || yy >= H || xx < 0 || xx >= W) return 0.f;
    long long p2 = (img * H + yy) * W + xx;
    return c < C1 ? to_f(x1[p2 * C1 + c]) : to_f(x2[p2 * C2 + (c - C1)]);
  }
};

constexpr int BM = 64, BN = 64, BK = 16;

template <typename T, typename TO>
__global__ void __launch_bounds__(256) gemm_simt_kernel(st_gemm_args a) {
  __shared__ float As[BK][BM + 4];
  __shared__ float Bs[BK][BN + 4];
  const int tid = threadIdx.x;
  const int split = a.split_k > 1 ? a.split_k : 1;
  const int b = blockIdx.z / split;
  const int ks = blockIdx.z % split;
  const int m0 = blockIdx.x * BM, n0 = blockIdx.y * BN;
  const int kper = (((a.K + split - 1) / split + BK - 1) / BK) * BK;
  const int kbeg = ks * kper;
  const int kend = min(a.K, kbeg + kper);

  const T* A = reinterpret_cast<const T*>(a.A);
  const T* B = reinterpret_cast<const T*>(a.B);
  const int Ct = a.C1 + a.C2;
  const long long n_pix = (long long)a.n_img * a.H * a.W;
  Gather<T> ga{reinterpret_cast<const T*>(a.A), reinterpret_cast<const T*>(a.A2), a.H, a.W, a.C1, a.C2, a.kh, a.kw};
  Gather<T> gb{reinterpret_cast<const T*>(a.B), reinterpret_cast<const T*>(a.B2), a.H, a.W, a.C1, a.C2, a.kh, a.kw};
  const int ntaps = a.kh * a.kw;

  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  const int tx = tid % 16, ty = tid / 16;   // thread computes rows ty*4.., cols tx*4..
  for (int k0 = kbeg; k0 < kend; k0 += BK) {
    // each thread loads 4 elements of each tile
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      int idx = tid + e * 256;          // 0..1023
      int kk, ii;
      // choose the fastest-varying index to follow the contiguous axis of the operand
      if (a.a_mode == ST_OP_GATHER || a.sAk == 1) { kk = idx % BK; ii = idx / BK; }
      else { ii = idx % BM; kk = idx / BM; }
      int m = m0 + ii, k = k0 + kk;
      float v = 0.f;
      if (m < a.M && k < kend) {
        if (a.a_mode == ST_OP_GATHER) v = ga.load(m, k / Ct, k % Ct, n_pix);
        else v = to_f(A[(long long)b * a.sAb + (long long)m * a.sAm + (long long)k * a.sAk]);
      }
      As[kk][ii] = v;
    }
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      int idx = tid + e * 256;
      int kk, ii;
      if (a.b_mode == ST_OP_STRIDED && a.sBk == 1) { kk = idx % BK; ii = idx / BK; }
      else { ii = idx % BN; kk = idx / BN; }
      int n = n0 + ii, k = k0 + kk;
      float v = 0.f;
      if (n < a.N && k < kend) {
        if (a.b_mode == ST_OP_GATHER) v = gb.load(k, n / Ct, n % Ct, n_pix);
        else if (a.b_mode == ST_OP_DGRADW) {
          int tap = k / Ct, co = k % Ct;
          v = to_f(B[((long long)co * ntaps + (ntaps - 1 - tap)) * a.N + n]);
        } else v = to_f(B[(long long)b * a.sBb + (long long)n * a.sBn + (long long)k * a.sBk]);
      }
      Bs[kk][ii] = v;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < BK; ++kk) {
      float av[4], bv[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) av[i] = As[kk][ty * 4 + i];
#pragma unroll
      for (int j = 0; j < 4; ++j) bv[j] = Bs[kk][tx * 4 + j];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
    __syncthreads();
  }

  TO* C = reinterpret_cast<TO*>(a.C);
  const T* R = reinterpret_cast<const T*>(a.residual);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    int m = m0 + ty * 4 + i;
    if (m >= a.M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      int n = n0 + tx * 4 + j;
      if (n >= a.N) continue;
      long long ci = (long long)b * a.sCb + (long long)m * a.sCm + n;
      float v = acc[i][j];
      if (a.accumulate) {
        if constexpr (sizeof(TO) == 4) {
          if (split > 1) atomicAdd(reinterpret_cast<float*>(C) + ci, a.alpha * v);
          else reinterpret_cast<float*>(C)[ci] += a.alpha * v;
        }
      } else {
        if (a.bias) v += a.bias[n];
        if (a.rowbias) v += a.rowbias[(long long)(m / a.rows_per_rb) * a.ld_rb + n];
        if (R) v += to_f(R[(long long)b * a.sRb + (long long)m * a.sRm + n]);
        C[ci] = from_f<TO>(a.alpha * v);
      }
    }
  }
}

}  // namespace

int st_gemm_simt(const st_gemm_args* a, cudaStream_t stream) {
  dim3 grid((a->M + BM - 1) / BM, (a->N + BN - 1) / BN, a->batch * (a->split_k > 1 ? a->split_k : 1));
  ST_CHECK_ARG(grid.y <= 65535 && grid.z <= 65535, "st_gemm(simt): grid too large");
  if (a->in_dtype == ST_F32 && a->out_dtype == ST_F32) gemm_simt_kernel<float, float><<<grid, 256, 0, stream>>>(*a);
  else if (a->in_dtype == ST_BF16 && a->out_dtype == ST_BF16) gemm_simt_kernel<bf16, bf16><<<grid, 256, 0, stream>>>(*a);
  else if (a->in_dtype == ST_BF16 && a->out_dtype == ST_F32) gemm_simt_kernel<bf16, float><<<grid, 256, 0, stream>>>(*a);
  else if (a->in_dtype == ST_F32 && a->out_dtype == ST_BF16) gemm_simt_kernel<float, bf16><<<grid, 256, 0, stream>>>(*a);
  else { st_set_error("st_gemm: bad dtypes"); return ST_ERR_ARG; }
  ST_CHECK_LAUNCH("st_gemm(simt)");
  return 0;
}
