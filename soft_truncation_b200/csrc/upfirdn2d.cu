// upfirdn2d: zero-insert upsample -> pad/crop -> 2-D FIR -> decimate, on [major][h][w][minor] tensors.
// Replaces reference op/upfirdn2d_kernel.cu:49-207 (index arithmetic :182-203) behind the same
// argument list as op/upfirdn2d.cpp:12-19.  Inside the network the tensors are NHWC, i.e.
// major = batch and minor = channels, so consecutive threads walk the contiguous channel axis.
#include <stdlib.h>

#include "common.cuh"

namespace {

constexpr int MAX_TAPS = 64;

__device__ __forceinline__ int floor_div(int a, int b) {
  int c = a / b;
  if (c * b > a) --c;
  return c;
}

struct UpfirdnP {
  int major, in_h, in_w, minor, kh, kw, up_x, up_y, down_x, down_y, pad_x0, pad_y0, out_h, out_w;
};

// VEC = elements of `minor` handled per thread (4 when minor % 4 == 0, else 1)
template <typename T, int VEC>
__global__ void __launch_bounds__(256) upfirdn2d_kernel(const T* __restrict__ x, T* __restrict__ y, const float* __restrict__ k,
                                                        UpfirdnP p, long long total) {
  __shared__ float sk[MAX_TAPS];   // flipped on load (reference :137)
  for (int i = threadIdx.x; i < p.kh * p.kw; i += blockDim.x) {
    int ky = i / p.kw, kx = i % p.kw;
    sk[i] = k[(p.kh - 1 - ky) * p.kw + (p.kw - 1 - kx)];
  }
  __syncthreads();
  const int mq = p.minor / VEC;
  const int ny = p.kh / p.up_y, nx = p.kw / p.up_x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int m = (int)(i % mq) * VEC;
    long long t = i / mq;
    const int ox = (int)(t % p.out_w);
    t /= p.out_w;
    const int oy = (int)(t % p.out_h);
    const long long mj = t / p.out_h;
    const int mid_y = oy * p.down_y + p.up_y - 1 - p.pad_y0;
    const int in_y = floor_div(mid_y, p.up_y);
    const int ky0 = (in_y + 1) * p.up_y - mid_y - 1;
    const int mid_x = ox * p.down_x + p.up_x - 1 - p.pad_x0;
    const int in_x = floor_div(mid_x, p.up_x);
    const int kx0 = (in_x + 1) * p.up_x - mid_x - 1;
    float acc[VEC];
#pragma unroll
    for (int v = 0; v < VEC; ++v) acc[v] = 0.f;
    for (int yy = 0; yy < ny; ++yy) {
      const int iy = in_y + yy;
      if (iy < 0 || iy >= p.in_h) continue;
      for (int xx = 0; xx < nx; ++xx) {
        const int ix = in_x + xx;
        if (ix < 0 || ix >= p.in_w) continue;
        const float w = sk[(ky0 + yy * p.up_y) * p.kw + kx0 + xx * p.up_x];
        const T* src = x + ((mj * p.in_h + iy) * p.in_w + ix) * p.minor + m;
        if constexpr (VEC == 4) {
          float v[4];
          load4(src, v);
#pragma unroll
          for (int q = 0; q < 4; ++q) acc[q] = fmaf(v[q], w, acc[q]);
        } else {
          acc[0] = fmaf(to_f(*src), w, acc[0]);
        }
      }
    }
    T* dst = y + ((mj * p.out_h + oy) * p.out_w + ox) * p.minor + m;
    if constexpr (VEC == 4) store4(dst, acc);
    else *dst = from_f<T>(acc[0]);
  }
}

// ---------------------------------------------------------------- fast path: the three FIR shapes the networks use
// (4x4 taps; upsample_2d = up 2 / pad0 2, downsample_2d = down 2 / pad0 1, conv_downsample_2d
// pre-filter = 1:1 / pad0 2), NHWC with the channel count a multiple of the 16-byte vector.  HBM-bound work: one thread
// produces a BY x BX block of output pixels for one 16-byte channel vector, loading every input vector of the block's
// footprint ONCE (12 loads for 8 outputs when up-sampling, 36 for 4 when down-sampling, 25 for 4 at 1:1 - instead of 4 /
// 16 / 16 per output) and applying it to every output whose window contains it; which tap that is, is a compile-time
// function of the unrolled loop indices, so the inner loop is loads and FMAs only.
__device__ __forceinline__ void ld8(const bf16* p, float v[8]) {
  uint4 t = __ldg(reinterpret_cast<const uint4*>(p));
  const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&t);
#pragma unroll
  for (int i = 0; i < 4; ++i) { v[2 * i] = __low2float(h[i]); v[2 * i + 1] = __high2float(h[i]); }
}
__device__ __forceinline__ void ld8(const float*, float*) {}
__device__ __forceinline__ void st8(bf16* p, const float v[8]) {
  uint4 t;
  __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&t);
#pragma unroll
  for (int i = 0; i < 4; ++i) h[i] = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]);
  *reinterpret_cast<uint4*>(p) = t;
}
__device__ __forceinline__ void st8(float*, const float*) {}

// packed fp32 FMA (sm_100 FFMA2): two accumulator lanes per issued instruction, same rounding as two FFMAs
__device__ __forceinline__ float2 ffma2(float2 a, float2 b, float2 c) {
  float2 r;
  asm("{\n\t.reg .b64 ra, rb, rc, rd;\n\tmov.b64 ra, {%2, %3};\n\tmov.b64 rb, {%4, %5};\n\tmov.b64 rc, {%6, %7};\n\t"
      "fma.rn.f32x2 rd, ra, rb, rc;\n\tmov.b64 {%0, %1}, rd;\n\t}"
      : "=f"(r.x), "=f"(r.y)
      : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y), "f"(c.x), "f"(c.y));
  return r;
}

template <int UP, int DOWN, int PAD0, int K>
struct FirGeom {
  // output index j (relative to a block origin that is a multiple of UP) -> first input index (relative) and first tap
  static __host__ __device__ constexpr int fdiv(int a, int b) { return (a >= 0) ? a / b : -((-a + b - 1) / b); }
  static __host__ __device__ constexpr int mid(int j) { return j * DOWN + UP - 1 - PAD0; }       // + origin * DOWN
  static __host__ __device__ constexpr int in0(int j) { return fdiv(mid(j), UP); }                 // + origin * DOWN / UP
  static __host__ __device__ constexpr int k0(int j) { return (in0(j) + 1) * UP - mid(j) - 1; }
  static constexpr int TAPS = K / UP;
};

template <typename T, int UP, int DOWN, int PAD0, int BY, int BX, int MINB = 1, bool F2 = false>
__global__ void __launch_bounds__(256, MINB) upfirdn2d_fast_kernel(const T* __restrict__ x, T* __restrict__ y, const float* __restrict__ k,
                                                             int in_h, int in_w, int minor, int out_h, int out_w, long long total) {
  constexpr int K = 4, V = 16 / (int)sizeof(T);
  using G = FirGeom<UP, DOWN, PAD0, K>;
  constexpr int R0 = G::in0(0), RY = G::in0(BY - 1) + G::TAPS - R0, RX = G::in0(BX - 1) + G::TAPS - R0;
  static_assert(BY % UP == 0 && BX % UP == 0, "block origin must keep the tap phase compile-time");
  pdl_wait();
  pdl_trigger();
  float w[K * K];                               // flipped taps (reference upfirdn2d_kernel.cu:137)
#pragma unroll
  for (int i = 0; i < K * K; ++i) w[i] = __ldg(k + (K * K - 1 - i));
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int mq = minor / V, bxn = (out_w + BX - 1) / BX, byn = (out_h + BY - 1) / BY;
  const int m = (int)(i % mq) * V;
  long long t = i / mq;
  const int ox0 = (int)(t % bxn) * BX;
  t /= bxn;
  const int oy0 = (int)(t % byn) * BY;
  const long long mj = t / byn;
  const int iy0 = oy0 * DOWN / UP + R0, ix0 = ox0 * DOWN / UP + R0;      // first input row / column of the footprint
  float acc[BY][BX][V];                         // (F2: read as V / 2 adjacent pairs by the packed FMA)
#pragma unroll
  for (int a = 0; a < BY; ++a)
#pragma unroll
    for (int b = 0; b < BX; ++b)
#pragma unroll
      for (int v = 0; v < V; ++v) acc[a][b][v] = 0.f;
  const T* src0 = x + (mj * in_h * (long long)in_w) * minor + m;
  // Branch-free footprint walk: out-of-image rows / columns are clamped to a valid address and their contribution is
  // zeroed through the weight, so that every load of a footprint row can be in flight at once.
  int colo[RX];
  float colm[RX];
#pragma unroll
  for (int c = 0; c < RX; ++c) {
    const int ix = ix0 + c;
    colm[c] = (ix >= 0 && ix < in_w) ? 1.f : 0.f;
    colo[c] = min(max(ix, 0), in_w - 1) * minor;
  }
#pragma unroll
  for (int r = 0; r < RY; ++r) {
    const int iy = iy0 + r;
    const float rowm = (iy >= 0 && iy < in_h) ? 1.f : 0.f;
    const T* row = src0 + (long long)min(max(iy, 0), in_h - 1) * in_w * minor;
    float v[RX][V];
#pragma unroll
    for (int c = 0; c < RX; ++c) {
      if constexpr (V == 8) ld8(row + colo[c], v[c]);
      else load4(row + colo[c], v[c]);
    }
#pragma unroll
    for (int c = 0; c < RX; ++c) {
      const float msk = rowm * colm[c];
#pragma unroll
      for (int a = 0; a < BY; ++a) {
        const int dy = r - (G::in0(a) - R0);
        if (dy < 0 || dy >= G::TAPS) continue;
#pragma unroll
        for (int b = 0; b < BX; ++b) {
          const int dx = c - (G::in0(b) - R0);
          if (dx < 0 || dx >= G::TAPS) continue;
          const float wt = w[(G::k0(a) + dy * UP) * K + G::k0(b) + dx * UP] * msk;
          if constexpr (F2) {
            const float2 w2 = make_float2(wt, wt);
#pragma unroll
            for (int e = 0; e < V; e += 2) {
              const float2 r = ffma2(make_float2(v[c][e], v[c][e + 1]), w2, make_float2(acc[a][b][e], acc[a][b][e + 1]));
              acc[a][b][e] = r.x;
              acc[a][b][e + 1] = r.y;
            }
          } else {
#pragma unroll
            for (int e = 0; e < V; ++e) acc[a][b][e] = fmaf(v[c][e], wt, acc[a][b][e]);
          }
        }
      }
    }
  }
#pragma unroll
  for (int a = 0; a < BY; ++a)
#pragma unroll
    for (int b = 0; b < BX; ++b) {
      if (oy0 + a >= out_h || ox0 + b >= out_w) continue;
      T* dst = y + ((mj * out_h + oy0 + a) * (long long)out_w + ox0 + b) * minor + m;
      if constexpr (V == 8) st8(dst, acc[a][b]);
      else store4(dst, acc[a][b]);
    }
}

template <typename T, int UP, int DOWN, int PAD0, int BY, int BX, int MINB = 1, bool F2 = false>
int launch_fast(const void* x, void* y, const float* k, int major, int in_h, int in_w, int minor, int out_h, int out_w,
                cudaStream_t stream) {
  constexpr int V = 16 / (int)sizeof(T);
  const long long total = (long long)major * ((out_h + BY - 1) / BY) * ((out_w + BX - 1) / BX) * (minor / V);
  const long long blocks = (total + 255) / 256;
  if (blocks > 0x7fffffffLL) return -1;
  cudaError_t e = st_launch(upfirdn2d_fast_kernel<T, UP, DOWN, PAD0, BY, BX, MINB, F2>, dim3((unsigned)blocks), dim3(256), 0, stream,
                            (const T*)x, (T*)y, k, in_h, in_w, minor, out_h, out_w, total);
  return e == cudaSuccess ? 0 : -1;
}

}  // namespace

extern "C" __attribute__((visibility("default"))) int st_upfirdn2d(const void* x, void* y, int dtype, const float* k, int major, int in_h, int in_w, int minor,
                            int kh, int kw, int up_x, int up_y, int down_x, int down_y, int pad_x0, int pad_x1, int pad_y0,
                            int pad_y1, void* stream) {
  ST_CHECK_ARG(kh * kw <= MAX_TAPS && kh > 0 && kw > 0, "st_upfirdn2d: FIR larger than %d taps", MAX_TAPS);
  ST_CHECK_ARG(up_x > 0 && up_y > 0 && down_x > 0 && down_y > 0, "st_upfirdn2d: bad up/down factors");
  UpfirdnP p;
  p.major = major; p.in_h = in_h; p.in_w = in_w; p.minor = minor; p.kh = kh; p.kw = kw;
  p.up_x = up_x; p.up_y = up_y; p.down_x = down_x; p.down_y = down_y; p.pad_x0 = pad_x0; p.pad_y0 = pad_y0;
  p.out_h = (in_h * up_y + pad_y0 + pad_y1 - kh + down_y) / down_y;   // op/upfirdn2d_kernel.cu:237-240
  p.out_w = (in_w * up_x + pad_x0 + pad_x1 - kw + down_x) / down_x;
  ST_CHECK_ARG(p.out_h > 0 && p.out_w > 0, "st_upfirdn2d: empty output");
  // ---- fast path for the networks' FIR shapes
  {
    const int V = dtype == ST_BF16 ? 8 : 4;
    const bool al = ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(y)) & 15) == 0;
    const char* env = getenv("ST_UPFIRDN_FAST");
    const bool on = !env || atoi(env) != 0;
    if (on && al && kh == 4 && kw == 4 && up_x == up_y && down_x == down_y && pad_x0 == pad_y0 && minor % V == 0 &&
        (dtype == ST_BF16 || dtype == ST_F32)) {
      int rc = 1;
      // Registers, not bytes, decide these kernels: uncapped, the down-sampling / 1:1 kernels take 170-220 registers
      // (ONE CTA of 8 warps per SM).  Measured per shape (profiles/r02_upfirdn_variants.txt): up-sampling is fastest
      // uncapped (128 registers, 2 CTAs per SM), down-sampling capped at 2 CTAs per SM (-20 %), the 1:1 pre-filter at 3 (-22 %).
      auto args = [&](auto fn) { return fn(x, y, k, major, in_h, in_w, minor, p.out_h, p.out_w, (cudaStream_t)stream); };
      // packed fp32 FMAs (FFMA2) in the tap loop: 3-10 % faster on every shape (profiles/r02_upfirdn_variants.txt); =0: scalar FFMAs
      static const int f2 = getenv("ST_FIR_FFMA2") ? atoi(getenv("ST_FIR_FFMA2")) : 1;
      ST_DISPATCH_DTYPE(dtype, T, {
        if (up_x == 2 && down_x == 1 && pad_x0 == 2)
          rc = f2 ? args(launch_fast<T, 2, 1, 2, 2, 4, 1, true>) : args(launch_fast<T, 2, 1, 2, 2, 4, 1>);
        else if (up_x == 1 && down_x == 2 && pad_x0 == 1)
          rc = f2 ? args(launch_fast<T, 1, 2, 1, 2, 2, 2, true>) : args(launch_fast<T, 1, 2, 1, 2, 2, 2>);
        else if (up_x == 1 && down_x == 1 && pad_x0 == 2)
          rc = f2 ? args(launch_fast<T, 1, 1, 2, 2, 2, 3, true>) : args(launch_fast<T, 1, 1, 2, 2, 2, 3>);
      });
      if (rc == 0) { ST_CHECK_LAUNCH("st_upfirdn2d"); return 0; }
      if (rc < 0) { cudaGetLastError(); }
    }
  }
  const bool vec = (minor % 4 == 0);
  long long total = (long long)major * p.out_h * p.out_w * (vec ? minor / 4 : minor);
  long long blocks = (total + 255) / 256;
  long long cap = (long long)st_num_sms() * 16;
  int grid = (int)(blocks > cap ? cap : (blocks < 1 ? 1 : blocks));
  ST_DISPATCH_DTYPE(dtype, T, {
    if (vec) upfirdn2d_kernel<T, 4><<<grid, 256, 0, (cudaStream_t)stream>>>((const T*)x, (T*)y, k, p, total);
    else upfirdn2d_kernel<T, 1><<<grid, 256, 0, (cudaStream_t)stream>>>((const T*)x, (T*)y, k, p, total);
  });
  ST_CHECK_LAUNCH("st_upfirdn2d");
  return 0;
}
