// upfirdn2d: zero-insert upsample -> pad/crop -> 2-D FIR -> decimate, on [major][h][w][minor] tensors.
// Replaces reference op/upfirdn2d_kernel.cu:49-207 (index arithmetic :182-203) behind the same
// argument list as op/upfirdn2d.cpp:12-19.  Inside the network the tensors are NHWC, i.e.
// major = batch and minor = channels, so consecutive threads walk the contiguous channel axis.
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
