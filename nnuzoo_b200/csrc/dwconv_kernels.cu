// dwconv_kernels.cu -- SS2D's depthwise 3x3 convolution fused with its SiLU, forward and backward.
//
// Reference: `x = self.act(self.conv2d(x))` with conv2d = Conv2d(d_inner, d_inner, 3, padding=1, groups=d_inner)
// (nnunetv2/nets/m2net.py:69-77, :214-215), NCHW planes.  PyTorch runs it as three native depthwise kernels plus two SiLU
// kernels; their weight-gradient kernel alone costs 26 ms of an M2Net step (profiles/r01_train_profile_v5_*.txt).
// The op is HBM-bound: forward reads x and writes y once; backward reads x and dy and writes dx once.
//
// Forward: a thread owns one column and kRows consecutive rows of a plane and slides a 3x3 window down
// (3 loads per new row, neighbours' loads hit L1).  Backward: a CTA owns a kTH x kTW tile; it stages x with a two-pixel
// ring in shared memory, recomputes the pre-activation on the tile plus a one-pixel ring and stores
// dpre = dy * silu'(pre) in shared memory, then every thread forms dx (the transposed stencil over dpre) and its share of
// dweight[3][3] / dbias, which are reduced over the CTA (warp shuffles + shared memory) and flushed with 10 atomicAdds
// per CTA (caller zeroes them).
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/nnuzoo_b200.h"

namespace nz {
void count_launch(int n);
void set_error(const char* fmt, ...);

template <typename T>
__device__ __forceinline__ float ld_f32(const T* p);
template <>
__device__ __forceinline__ float ld_f32<float>(const float* p) { return *p; }
template <>
__device__ __forceinline__ float ld_f32<__nv_bfloat16>(const __nv_bfloat16* p) { return __bfloat162float(*p); }
template <>
__device__ __forceinline__ float ld_f32<__half>(const __half* p) { return __half2float(*p); }
template <typename T>
__device__ __forceinline__ void st_f32(T* p, float v);
template <>
__device__ __forceinline__ void st_f32<float>(float* p, float v) { *p = v; }
template <>
__device__ __forceinline__ void st_f32<__nv_bfloat16>(__nv_bfloat16* p, float v) { *p = __float2bfloat16_rn(v); }
template <>
__device__ __forceinline__ void st_f32<__half>(__half* p, float v) { *p = __float2half_rn(v); }

__device__ __forceinline__ float silu_f(float p) { return p / (1.f + __expf(-p)); }
__device__ __forceinline__ float dsilu_f(float p) {
  const float s = 1.f / (1.f + __expf(-p));
  return s * (1.f + p * (1.f - s));
}

constexpr int kRows = 8;  // forward: rows per thread

template <typename T>
__global__ void __launch_bounds__(256) dwconv3x3_silu_fwd_kernel(const T* __restrict__ x, const float* __restrict__ wgt,
                                                                 const float* __restrict__ bias, T* __restrict__ y,
                                                                 long planes, int dim, int H, int W, int silu) {
  const int strips = (H + kRows - 1) / kRows;
  const long total = planes * strips * W;
  for (long i = blockIdx.x * 256L + threadIdx.x; i < total; i += (long)gridDim.x * 256) {
    const int w = (int)(i % W);
    const int strip = (int)((i / W) % strips);
    const long plane = i / ((long)W * strips);
    const int d = (int)(plane % dim);
    float k[9];
#pragma unroll
    for (int q = 0; q < 9; ++q) k[q] = wgt[d * 9 + q];
    const float b = bias ? bias[d] : 0.f;
    const T* xp = x + plane * (long)H * W;
    T* yp = y + plane * (long)H * W;
    const int h0 = strip * kRows;
    const bool wl = w > 0, wr = w + 1 < W;
    float r0[3], r1[3], r2[3];
    auto load_row = [&](int h, float (&r)[3]) {
      if (h < 0 || h >= H) {
        r[0] = r[1] = r[2] = 0.f;
      } else {
        const T* p = xp + (long)h * W + w;
        r[0] = wl ? ld_f32<T>(p - 1) : 0.f;
        r[1] = ld_f32<T>(p);
        r[2] = wr ? ld_f32<T>(p + 1) : 0.f;
      }
    };
    load_row(h0 - 1, r0);
    load_row(h0, r1);
#pragma unroll
    for (int j = 0; j < kRows; ++j) {
      const int h = h0 + j;
      if (h >= H) break;
      load_row(h + 1, r2);
      float acc = b;
#pragma unroll
      for (int q = 0; q < 3; ++q) acc = fmaf(k[q], r0[q], fmaf(k[3 + q], r1[q], fmaf(k[6 + q], r2[q], acc)));
      st_f32<T>(yp + (long)h * W + w, silu ? silu_f(acc) : acc);
#pragma unroll
      for (int q = 0; q < 3; ++q) r0[q] = r1[q], r1[q] = r2[q];
    }
  }
}


// 8 consecutive elements -> 8 floats / back (one 16-byte access for 16-bit types, two for fp32)
template <typename T>
__device__ __forceinline__ void ld8g(const T* __restrict__ p, float* f) {
  if constexpr (sizeof(T) == 4) {
    const float4 a = *reinterpret_cast<const float4*>(p), b = *reinterpret_cast<const float4*>(p + 4);
    f[0] = a.x, f[1] = a.y, f[2] = a.z, f[3] = a.w, f[4] = b.x, f[5] = b.y, f[6] = b.z, f[7] = b.w;
  } else {
    const uint4 v = *reinterpret_cast<const uint4*>(p);
    const T* e = reinterpret_cast<const T*>(&v);
#pragma unroll
    for (int i = 0; i < 8; ++i) f[i] = ld_f32<T>(e + i);
  }
}
template <typename T>
__device__ __forceinline__ void st8g(T* __restrict__ p, const float* f) {
  if constexpr (sizeof(T) == 4) {
    *reinterpret_cast<float4*>(p) = make_float4(f[0], f[1], f[2], f[3]);
    *reinterpret_cast<float4*>(p + 4) = make_float4(f[4], f[5], f[6], f[7]);
  } else {
    uint4 v;
    T* e = reinterpret_cast<T*>(&v);
#pragma unroll
    for (int i = 0; i < 8; ++i) st_f32<T>(e + i, f[i]);
    *reinterpret_cast<uint4*>(p) = v;
  }
}

// forward, W % 8 == 0 and 16-byte aligned planes: a thread owns 8 columns x kRows rows; per row one vector load plus
// the two neighbours' edge elements
template <typename T>
__global__ void __launch_bounds__(256) dwconv3x3_silu_fwd_vec_kernel(const T* __restrict__ x,
                                                                     const float* __restrict__ wgt,
                                                                     const float* __restrict__ bias, T* __restrict__ y,
                                                                     long planes, int dim, int H, int W, int silu) {
  const int strips = (H + kRows - 1) / kRows, W8 = W / 8;
  const long total = planes * strips * W8;
  for (long i = blockIdx.x * 256L + threadIdx.x; i < total; i += (long)gridDim.x * 256) {
    const int w0 = (int)(i % W8) * 8;
    const int strip = (int)((i / W8) % strips);
    const long plane = i / ((long)W8 * strips);
    const int d = (int)(plane % dim);
    float k[9];
#pragma unroll
    for (int q = 0; q < 9; ++q) k[q] = wgt[d * 9 + q];
    const float b = bias ? bias[d] : 0.f;
    const T* xp = x + plane * (long)H * W;
    T* yp = y + plane * (long)H * W;
    const int h0 = strip * kRows;
    float r0[10], r1[10], r2[10];
    auto load_row = [&](int h, float* r) {
      if (h < 0 || h >= H) {
#pragma unroll
        for (int c = 0; c < 10; ++c) r[c] = 0.f;
      } else {
        const T* p = xp + (long)h * W + w0;
        ld8g<T>(p, r + 1);
        r[0] = w0 > 0 ? ld_f32<T>(p - 1) : 0.f;
        r[9] = w0 + 8 < W ? ld_f32<T>(p + 8) : 0.f;
      }
    };
    load_row(h0 - 1, r0);
    load_row(h0, r1);
#pragma unroll
    for (int j = 0; j < kRows; ++j) {
      const int h = h0 + j;
      if (h >= H) break;
      load_row(h + 1, r2);
      float o[8];
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        float acc = b;
#pragma unroll
        for (int q = 0; q < 3; ++q) acc = fmaf(k[q], r0[c + q], fmaf(k[3 + q], r1[c + q], fmaf(k[6 + q], r2[c + q], acc)));
        o[c] = silu ? silu_f(acc) : acc;
      }
      st8g<T>(yp + (long)h * W + w0, o);
#pragma unroll
      for (int c = 0; c < 10; ++c) r0[c] = r1[c], r1[c] = r2[c];
    }
  }
}

constexpr int kTH = 32, kTW = 64;  // backward tile (interior); 256 threads, 8 interior pixels per thread between barriers

template <typename T>
__global__ void __launch_bounds__(256) dwconv3x3_silu_bwd_kernel(const T* __restrict__ x, const T* __restrict__ dy,
                                                                 const float* __restrict__ wgt,
                                                                 const float* __restrict__ bias, T* __restrict__ dx,
                                                                 float* __restrict__ dwgt, float* __restrict__ dbias,
                                                                 long planes, int dim, int H, int W, int silu) {
  __shared__ float xt[kTH + 4][kTW + 4];    // x on the tile + a two-pixel ring (zero outside the image)
  __shared__ float dpre[kTH + 2][kTW + 2];
  __shared__ float red[8][10];
  const int tiles_w = (W + kTW - 1) / kTW, tiles_h = (H + kTH - 1) / kTH;
  const long ntile = planes * tiles_w * tiles_h;
  const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
  for (long tile = blockIdx.x; tile < ntile; tile += gridDim.x) {
    const long plane = tile / ((long)tiles_w * tiles_h);
    const int tr = (int)(tile % ((long)tiles_w * tiles_h));
    const int h0 = (tr / tiles_w) * kTH, w0 = (tr % tiles_w) * kTW;
    const int d = (int)(plane % dim);
    float k[9];
#pragma unroll
    for (int q = 0; q < 9; ++q) k[q] = wgt[d * 9 + q];
    const float b = bias ? bias[d] : 0.f;
    const T* xp = x + plane * (long)H * W;
    const T* gp = dy + plane * (long)H * W;
    for (int i = t; i < (kTH + 4) * (kTW + 4); i += 256) {
      const int rr = i / (kTW + 4), cc = i % (kTW + 4);
      const int h = h0 + rr - 2, w = w0 + cc - 2;
      xt[rr][cc] = (h >= 0 && h < H && w >= 0 && w < W) ? ld_f32<T>(xp + (long)h * W + w) : 0.f;
    }
    __syncthreads();
    auto xat = [&](int h, int w) -> float { return xt[h - h0 + 2][w - w0 + 2]; };
    // phase 1: dpre on the tile + ring (zero outside the image: those outputs do not exist)
    for (int i = t; i < (kTH + 2) * (kTW + 2); i += 256) {
      const int rr = i / (kTW + 2), cc = i % (kTW + 2);
      const int h = h0 + rr - 1, w = w0 + cc - 1;
      float v = 0.f;
      if (h >= 0 && h < H && w >= 0 && w < W) {
        const float g = ld_f32<T>(gp + (long)h * W + w);
        if (silu) {
          float acc = b;
#pragma unroll
          for (int a = 0; a < 3; ++a)
#pragma unroll
            for (int c = 0; c < 3; ++c) acc = fmaf(k[a * 3 + c], xat(h + a - 1, w + c - 1), acc);
          v = g * dsilu_f(acc);
        } else {
          v = g;
        }
      }
      dpre[rr][cc] = v;
    }
    __syncthreads();
    // phase 2: dx and the parameter gradients of the interior pixels
    float gw[9], gb = 0.f;
#pragma unroll
    for (int q = 0; q < 9; ++q) gw[q] = 0.f;
#pragma unroll
    for (int j = 0; j < (kTH * kTW) / 256; ++j) {
      const int p = t + j * 256;
      const int rr = p / kTW, cc = p % kTW;
      const int h = h0 + rr, w = w0 + cc;
      if (h < H && w < W) {
        // dx[h, w] = sum_{a, c} k[a][c] * dpre[h - a + 1, w - c + 1]
        float acc = 0.f;
#pragma unroll
        for (int a = 0; a < 3; ++a)
#pragma unroll
          for (int c = 0; c < 3; ++c) acc = fmaf(k[a * 3 + c], dpre[rr + 1 - a + 1][cc + 1 - c + 1], acc);
        st_f32<T>(dx + plane * (long)H * W + (long)h * W + w, acc);
        const float g = dpre[rr + 1][cc + 1];
        gb += g;
#pragma unroll
        for (int a = 0; a < 3; ++a)
#pragma unroll
          for (int c = 0; c < 3; ++c) gw[a * 3 + c] = fmaf(g, xat(h + a - 1, w + c - 1), gw[a * 3 + c]);
      }
    }
    // CTA reduction of the 10 parameter gradients
#pragma unroll
    for (int q = 0; q < 9; ++q)
      for (int o = 16; o > 0; o >>= 1) gw[q] += __shfl_xor_sync(0xffffffffu, gw[q], o);
    for (int o = 16; o > 0; o >>= 1) gb += __shfl_xor_sync(0xffffffffu, gb, o);
    if (lane == 0) {
#pragma unroll
      for (int q = 0; q < 9; ++q) red[warp][q] = gw[q];
      red[warp][9] = gb;
    }
    __syncthreads();
    if (t < 10) {
      float s = 0.f;
#pragma unroll
      for (int wq = 0; wq < 8; ++wq) s += red[wq][t];
      if (t < 9) {
        if (dwgt) atomicAdd(dwgt + d * 9 + t, s);
      } else if (dbias) {
        atomicAdd(dbias + d, s);
      }
    }
    __syncthreads();
  }
}

static int grid_for(long work_items) {
  long blocks = (work_items + 255) / 256;
  const long cap = 148L * 16;
  if (blocks > cap) blocks = cap;
  return (int)(blocks < 1 ? 1 : blocks);
}

}  // namespace nz

extern "C" int nz_dwconv3x3_fwd(const void* x, const float* weight, const float* bias, void* y, int32_t dtype,
                                int32_t batch, int32_t dim, int32_t H, int32_t W, int32_t silu, void* stream) {
  using namespace nz;
  if (!x || !weight || !y || batch < 1 || dim < 1 || H < 1 || W < 1) {
    set_error("nz_dwconv3x3_fwd: null pointer or empty shape");
    return NZ_EINVAL;
  }
  const long planes = (long)batch * dim;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (W % 8 == 0 && (reinterpret_cast<uintptr_t>(x) & 15) == 0 && (reinterpret_cast<uintptr_t>(y) & 15) == 0 &&
      (dtype == NZ_F32 || dtype == NZ_BF16 || dtype == NZ_F16)) {
    const int gridv = grid_for(planes * ((H + kRows - 1) / kRows) * (W / 8));
    if (dtype == NZ_F32)
      dwconv3x3_silu_fwd_vec_kernel<float><<<gridv, 256, 0, st>>>(static_cast<const float*>(x), weight, bias,
                                                                  static_cast<float*>(y), planes, dim, H, W, silu);
    else if (dtype == NZ_BF16)
      dwconv3x3_silu_fwd_vec_kernel<__nv_bfloat16><<<gridv, 256, 0, st>>>(
          static_cast<const __nv_bfloat16*>(x), weight, bias, static_cast<__nv_bfloat16*>(y), planes, dim, H, W, silu);
    else
      dwconv3x3_silu_fwd_vec_kernel<__half><<<gridv, 256, 0, st>>>(static_cast<const __half*>(x), weight, bias,
                                                                   static_cast<__half*>(y), planes, dim, H, W, silu);
    count_launch(1);
    return cudaGetLastError() == cudaSuccess ? NZ_OK : NZ_ECUDA;
  }
  const int grid = grid_for(planes * ((H + kRows - 1) / kRows) * W);
  if (dtype == NZ_F32)
    dwconv3x3_silu_fwd_kernel<float><<<grid, 256, 0, st>>>(static_cast<const float*>(x), weight, bias,
                                                           static_cast<float*>(y), planes, dim, H, W, silu);
  else if (dtype == NZ_BF16)
    dwconv3x3_silu_fwd_kernel<__nv_bfloat16><<<grid, 256, 0, st>>>(static_cast<const __nv_bfloat16*>(x), weight, bias,
                                                                   static_cast<__nv_bfloat16*>(y), planes, dim, H, W, silu);
  else if (dtype == NZ_F16)
    dwconv3x3_silu_fwd_kernel<__half><<<grid, 256, 0, st>>>(static_cast<const __half*>(x), weight, bias,
                                                            static_cast<__half*>(y), planes, dim, H, W, silu);
  else {
    set_error("nz_dwconv3x3_fwd: unsupported dtype %d", dtype);
    return NZ_EINVAL;
  }
  count_launch(1);
  return cudaGetLastError() == cudaSuccess ? NZ_OK : NZ_ECUDA;
}

extern "C" int nz_dwconv3x3_bwd(const void* x, const void* dy, const float* weight, const float* bias, void* dx,
                                float* dweight, float* dbias, int32_t dtype, int32_t batch, int32_t dim, int32_t H,
                                int32_t W, int32_t silu, void* stream) {
  using namespace nz;
  if (!x || !dy || !weight || !dx || batch < 1 || dim < 1 || H < 1 || W < 1) {
    set_error("nz_dwconv3x3_bwd: null pointer or empty shape");
    return NZ_EINVAL;
  }
  const long planes = (long)batch * dim;
  const long ntile = planes * ((W + kTW - 1) / kTW) * ((H + kTH - 1) / kTH);
  const int grid = (int)(ntile < 148L * 16 ? ntile : 148L * 16);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (dtype == NZ_F32)
    dwconv3x3_silu_bwd_kernel<float><<<grid, 256, 0, st>>>(static_cast<const float*>(x), static_cast<const float*>(dy),
                                                           weight, bias, static_cast<float*>(dx), dweight, dbias, planes,
                                                           dim, H, W, silu);
  else if (dtype == NZ_BF16)
    dwconv3x3_silu_bwd_kernel<__nv_bfloat16><<<grid, 256, 0, st>>>(
        static_cast<const __nv_bfloat16*>(x), static_cast<const __nv_bfloat16*>(dy), weight, bias,
        static_cast<__nv_bfloat16*>(dx), dweight, dbias, planes, dim, H, W, silu);
  else if (dtype == NZ_F16)
    dwconv3x3_silu_bwd_kernel<__half><<<grid, 256, 0, st>>>(static_cast<const __half*>(x),
                                                            static_cast<const __half*>(dy), weight, bias,
                                                            static_cast<__half*>(dx), dweight, dbias, planes, dim, H, W,
                                                            silu);
  else {
    set_error("nz_dwconv3x3_bwd: unsupported dtype %d", dtype);
    return NZ_EINVAL;
  }
  count_launch(1);
  return cudaGetLastError() == cudaSuccess ? NZ_OK : NZ_ECUDA;
}
