// epilogue_kernels.cu -- SS2D's post-scan chain as one kernel each way (2-D, 4 directions):
//   y   = CrossMerge(out_y)                                  nnunetv2/nets/m2net.py:202-206 + :218
//   y   = y.transpose(1, 2).view(B, H, W, D)                 :219
//   y   = LayerNorm_D(y)                                     :220
//   out = y * SiLU(z)                                        :221
// The reference (and our un-fused modules) run this as a merge, a transpose copy, a LayerNorm, a SiLU, a multiply and
// the cast in front of out_proj: seven passes over (B, D, L)-sized tensors each way.  Fused: read the four scan outputs
// and z once, write the gated result once (+ the merged y in channels-last order and the row statistics the backward
// needs); backward: read dout, y, z once, write dz and the four permuted copies of dy once.
//
// A CTA owns a TH x TW tile of spatial positions and all D channels of it in shared memory ([position][D + 1] fp32, the
// odd pitch makes both access patterns conflict-free).  The row-major directions (k0, k2) are read / written in runs
// of TW elements, the column-major ones (k1, k3) in runs of TH, so every direction moves whole sectors; the four values
// are added in the reference's order ((y0 + flip y2) + T y1) + T flip y3, fp32, so the merged y is bit-identical to
// nz_cross_merge.  LayerNorm: one warp per position, lanes strided over channels, fp32 two-pass statistics.
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/nnuzoo_b200.h"

namespace nz {
void count_launch(int n);
void set_error(const char* fmt, ...);

template <typename T>
__device__ __forceinline__ float e_ld(const T* p);
template <>
__device__ __forceinline__ float e_ld<float>(const float* p) { return *p; }
template <>
__device__ __forceinline__ float e_ld<__nv_bfloat16>(const __nv_bfloat16* p) { return __bfloat162float(*p); }
template <>
__device__ __forceinline__ float e_ld<__half>(const __half* p) { return __half2float(*p); }
template <typename T>
__device__ __forceinline__ void e_st(T* p, float v);
template <>
__device__ __forceinline__ void e_st<float>(float* p, float v) { *p = v; }
template <>
__device__ __forceinline__ void e_st<__nv_bfloat16>(__nv_bfloat16* p, float v) { *p = __float2bfloat16_rn(v); }
template <>
__device__ __forceinline__ void e_st<__half>(__half* p, float v) { *p = __float2half_rn(v); }

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

constexpr int kEpThreads = 256;
constexpr int kEpMaxDPL = 8;  // channels per lane: D <= 256

struct EpArgs {
  // forward
  const float* out_y;  // (B, 4, D, L) fp32
  const void* z;       // (B, L, D) rows of D contiguous elements, row stride z_ls, batch stride z_bs (elements)
  long z_bs, z_ls;
  const float *gamma, *beta;
  void* out;     // (B, L, D) contiguous, TO
  float* ym;     // (B, L, D) fp32 merged y (saved for the backward)
  float *mean, *rstd;  // (B * L)
  // backward
  const void* dout;  // (B, L, D) contiguous, TO
  void* d_out_y;     // (B, 4, D, L), TG
  void* dz;          // (B, L, D) contiguous, TZ
  float *dgamma, *dbeta;
  int B, D, H, W, TH, TW;
  float eps;
};

// tile decode shared by both kernels
struct EpTile {
  long b;
  int h0, w0;
};
__device__ __forceinline__ bool ep_tile(const EpArgs& a, long tile, EpTile* t) {
  const int tw = (a.W + a.TW - 1) / a.TW, th = (a.H + a.TH - 1) / a.TH;
  const long per = (long)tw * th;
  if (tile >= per * a.B) return false;
  t->b = tile / per;
  const int r = (int)(tile % per);
  t->h0 = (r / tw) * a.TH;
  t->w0 = (r % tw) * a.TW;
  return true;
}

template <typename TZ, typename TO>
__global__ void __launch_bounds__(kEpThreads) ss2d_epilogue_fwd_kernel(EpArgs a) {
  extern __shared__ float sm[];  // [P][D + 1]
  const int D = a.D, P = a.TH * a.TW, pitch = D + 1;
  const long L = (long)a.H * a.W;
  const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
  const int dpl = D / 32;
  float gm[kEpMaxDPL], bt[kEpMaxDPL];
#pragma unroll
  for (int i = 0; i < kEpMaxDPL; ++i)
    if (i < dpl) gm[i] = a.gamma ? a.gamma[lane + 32 * i] : 1.f, bt[i] = a.beta ? a.beta[lane + 32 * i] : 0.f;
  EpTile tl;
  for (long tile = blockIdx.x; ep_tile(a, tile, &tl); tile += gridDim.x) {
    const float* oy = a.out_y + tl.b * 4 * D * L;
    const long ks = (long)D * L;
    // row-major directions: y0[p] + y2[L-1-p]
    for (int i = t; i < P * D; i += kEpThreads) {
      const int p = i % P, d = i / P;
      const int h = tl.h0 + p / a.TW, w = tl.w0 + p % a.TW;
      float v = 0.f;
      if (h < a.H && w < a.W) {
        const long pos = (long)h * a.W + w;
        v = oy[d * L + pos];
        v = v + oy[2 * ks + d * L + (L - 1 - pos)];
      }
      sm[p * pitch + d] = v;
    }
    __syncthreads();
    // column-major directions: + y1[w*H + h] + y3[L-1-(w*H + h)]
    for (int i = t; i < P * D; i += kEpThreads) {
      const int q = i % P, d = i / P;
      const int qh = q % a.TH, qw = q / a.TH;
      const int h = tl.h0 + qh, w = tl.w0 + qw;
      if (h < a.H && w < a.W) {
        const long j = (long)w * a.H + h;
        float* s = sm + (qh * a.TW + qw) * pitch + d;
        float v = *s;
        v = v + oy[ks + d * L + j];
        v = v + oy[3 * ks + d * L + (L - 1 - j)];
        *s = v;
      }
    }
    __syncthreads();
    // LayerNorm over D and the SiLU(z) gate, one warp per position
    for (int p = warp; p < P; p += kEpThreads / 32) {
      const int h = tl.h0 + p / a.TW, w = tl.w0 + p % a.TW;
      if (h >= a.H || w >= a.W) continue;  // warp-uniform
      const long pos = (long)h * a.W + w;
      float v[kEpMaxDPL];
      float s = 0.f;
#pragma unroll
      for (int i = 0; i < kEpMaxDPL; ++i)
        if (i < dpl) v[i] = sm[p * pitch + lane + 32 * i], s += v[i];
      const float mu = warp_sum(s) / (float)D;
      float q = 0.f;
#pragma unroll
      for (int i = 0; i < kEpMaxDPL; ++i)
        if (i < dpl) q = fmaf(v[i] - mu, v[i] - mu, q);
      const float rs = rsqrtf(warp_sum(q) / (float)D + a.eps);
      const long row = tl.b * L + pos;
      const TZ* zr = static_cast<const TZ*>(a.z) + tl.b * a.z_bs + pos * a.z_ls;
#pragma unroll
      for (int i = 0; i < kEpMaxDPL; ++i)
        if (i < dpl) {
          const int d = lane + 32 * i;
          const float zz = e_ld<TZ>(zr + d);
          const float ln = fmaf((v[i] - mu) * rs, gm[i], bt[i]);
          e_st<TO>(static_cast<TO*>(a.out) + row * D + d, ln * (zz / (1.f + __expf(-zz))));
          a.ym[row * D + d] = v[i];
        }
      if (lane == 0) a.mean[row] = mu, a.rstd[row] = rs;
    }
    __syncthreads();
  }
}

template <typename TZ, typename TO, typename TG>
__global__ void __launch_bounds__(kEpThreads) ss2d_epilogue_bwd_kernel(EpArgs a) {
  extern __shared__ float sm[];  // [P][D + 1] dy, then [2][D] parameter-gradient fold
  const int D = a.D, P = a.TH * a.TW, pitch = D + 1;
  float* red = sm + P * pitch;
  const long L = (long)a.H * a.W;
  const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
  const int dpl = D / 32;
  float gm[kEpMaxDPL], bt[kEpMaxDPL], ag[kEpMaxDPL], ab[kEpMaxDPL];
#pragma unroll
  for (int i = 0; i < kEpMaxDPL; ++i) {
    ag[i] = ab[i] = 0.f;
    if (i < dpl) gm[i] = a.gamma ? a.gamma[lane + 32 * i] : 1.f, bt[i] = a.beta ? a.beta[lane + 32 * i] : 0.f;
  }
  for (int i = t; i < 2 * D; i += kEpThreads) red[i] = 0.f;
  EpTile tl;
  for (long tile = blockIdx.x; ep_tile(a, tile, &tl); tile += gridDim.x) {
    __syncthreads();
    for (int p = warp; p < P; p += kEpThreads / 32) {
      const int h = tl.h0 + p / a.TW, w = tl.w0 + p % a.TW;
      if (h >= a.H || w >= a.W) continue;  // warp-uniform
      const long pos = (long)h * a.W + w;
      const long row = tl.b * L + pos;
      const float mu = a.mean[row], rs = a.rstd[row];
      const TZ* zr = static_cast<const TZ*>(a.z) + tl.b * a.z_bs + pos * a.z_ls;
      float xh[kEpMaxDPL], gg[kEpMaxDPL];
      float s1 = 0.f, s2 = 0.f;
#pragma unroll
      for (int i = 0; i < kEpMaxDPL; ++i)
        if (i < dpl) {
          const int d = lane + 32 * i;
          const float g = e_ld<TO>(static_cast<const TO*>(a.dout) + row * D + d);
          const float zz = e_ld<TZ>(zr + d);
          const float sg = 1.f / (1.f + __expf(-zz));
          xh[i] = (a.ym[row * D + d] - mu) * rs;
          const float ln = fmaf(xh[i], gm[i], bt[i]);
          e_st<TZ>(static_cast<TZ*>(a.dz) + row * D + d, g * ln * sg * (1.f + zz * (1.f - sg)));
          const float dln = g * zz * sg;
          ag[i] = fmaf(dln, xh[i], ag[i]);
          ab[i] += dln;
          gg[i] = dln * gm[i];
          s1 += gg[i];
          s2 = fmaf(gg[i], xh[i], s2);
        }
      const float m1 = warp_sum(s1) / (float)D, m2 = warp_sum(s2) / (float)D;
#pragma unroll
      for (int i = 0; i < kEpMaxDPL; ++i)
        if (i < dpl) sm[p * pitch + lane + 32 * i] = rs * (gg[i] - m1 - xh[i] * m2);
    }
    __syncthreads();
    TG* go = static_cast<TG*>(a.d_out_y) + tl.b * 4 * D * L;
    const long ks = (long)D * L;
    for (int i = t; i < P * D; i += kEpThreads) {  // row-major copies
      const int p = i % P, d = i / P;
      const int h = tl.h0 + p / a.TW, w = tl.w0 + p % a.TW;
      if (h < a.H && w < a.W) {
        const long pos = (long)h * a.W + w;
        const float v = sm[p * pitch + d];
        e_st<TG>(go + d * L + pos, v);
        e_st<TG>(go + 2 * ks + d * L + (L - 1 - pos), v);
      }
    }
    for (int i = t; i < P * D; i += kEpThreads) {  // column-major copies
      const int q = i % P, d = i / P;
      const int qh = q % a.TH, qw = q / a.TH;
      const int h = tl.h0 + qh, w = tl.w0 + qw;
      if (h < a.H && w < a.W) {
        const long j = (long)w * a.H + h;
        const float v = sm[(qh * a.TW + qw) * pitch + d];
        e_st<TG>(go + ks + d * L + j, v);
        e_st<TG>(go + 3 * ks + d * L + (L - 1 - j), v);
      }
    }
  }
  __syncthreads();
#pragma unroll
  for (int i = 0; i < kEpMaxDPL; ++i)
    if (i < dpl) {
      atomicAdd(red + lane + 32 * i, ag[i]);
      atomicAdd(red + D + lane + 32 * i, ab[i]);
    }
  __syncthreads();
  for (int i = t; i < D; i += kEpThreads) {
    if (a.dgamma) atomicAdd(a.dgamma + i, red[i]);
    if (a.dbeta) atomicAdd(a.dbeta + i, red[D + i]);
  }
}

static bool ep_config(EpArgs& a, size_t* smem, int* grid, bool bwd) {
  if (a.D % 32 || a.D < 32 || a.D > 32 * kEpMaxDPL) return false;
  // positions per tile so that the tile stays near 32 KB: 256 / 128 / 64 / 32 for D = 32 / 64 / 128 / 256
  int P = 8192 / a.D;
  if (P > 256) P = 256;
  a.TW = P >= 256 ? 16 : (P >= 128 ? 16 : 8);
  a.TH = P / a.TW;
  *smem = ((size_t)P * (a.D + 1) + (bwd ? 2 * a.D : 0)) * sizeof(float);
  const long tiles = (long)a.B * ((a.W + a.TW - 1) / a.TW) * ((a.H + a.TH - 1) / a.TH);
  *grid = (int)(tiles < 148L * 8 ? tiles : 148L * 8);
  return true;
}

}  // namespace nz

extern "C" int nz_ss2d_epilogue_supported(int32_t D) { return (D % 32 == 0 && D >= 32 && D <= 32 * nz::kEpMaxDPL) ? 1 : 0; }

extern "C" int nz_ss2d_epilogue_fwd(const float* out_y, const void* z, const int64_t* z_stride, const float* gamma,
                                    const float* beta, void* out, float* y_merged, float* mean, float* rstd,
                                    int32_t z_dtype, int32_t out_dtype, int32_t batch, int32_t D, int32_t H, int32_t W,
                                    float eps, void* stream) {
  using namespace nz;
  if (!out_y || !z || !z_stride || !out || !y_merged || !mean || !rstd || batch < 1 || H < 1 || W < 1) {
    set_error("nz_ss2d_epilogue_fwd: null pointer or empty shape");
    return NZ_EINVAL;
  }
  EpArgs a{};
  a.out_y = out_y, a.z = z, a.z_bs = z_stride[0], a.z_ls = z_stride[1], a.gamma = gamma, a.beta = beta, a.out = out;
  a.ym = y_merged, a.mean = mean, a.rstd = rstd, a.B = batch, a.D = D, a.H = H, a.W = W, a.eps = eps;
  size_t smem;
  int grid;
  if (!ep_config(a, &smem, &grid, false)) {
    set_error("nz_ss2d_epilogue_fwd: D = %d unsupported (multiple of 32, 32..256)", D);
    return NZ_EUNSUPPORTED;
  }
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
#define NZ_EP_FWD(TZ, TO)                                                                                    \
  do {                                                                                                       \
    auto kern = ss2d_epilogue_fwd_kernel<TZ, TO>;                                                            \
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);                      \
    kern<<<grid, kEpThreads, smem, st>>>(a);                                                                 \
  } while (0)
  const int key = z_dtype * 4 + out_dtype;
  switch (key) {
    case NZ_F32 * 4 + NZ_F32: NZ_EP_FWD(float, float); break;
    case NZ_BF16 * 4 + NZ_BF16: NZ_EP_FWD(__nv_bfloat16, __nv_bfloat16); break;
    case NZ_BF16 * 4 + NZ_F32: NZ_EP_FWD(__nv_bfloat16, float); break;
    case NZ_F16 * 4 + NZ_F16: NZ_EP_FWD(__half, __half); break;
    case NZ_F16 * 4 + NZ_F32: NZ_EP_FWD(__half, float); break;
    default:
      set_error("nz_ss2d_epilogue_fwd: unsupported dtypes (z %d, out %d)", z_dtype, out_dtype);
      return NZ_EINVAL;
  }
#undef NZ_EP_FWD
  count_launch(1);
  return cudaGetLastError() == cudaSuccess ? NZ_OK : NZ_ECUDA;
}

extern "C" int nz_ss2d_epilogue_bwd(const void* dout, const float* y_merged, const float* mean, const float* rstd,
                                    const void* z, const int64_t* z_stride, const float* gamma, const float* beta,
                                    void* d_out_y, void* dz, float* dgamma, float* dbeta, int32_t z_dtype,
                                    int32_t out_dtype, int32_t grad_dtype, int32_t batch, int32_t D, int32_t H, int32_t W,
                                    void* stream) {
  using namespace nz;
  if (!dout || !y_merged || !mean || !rstd || !z || !z_stride || !d_out_y || !dz || batch < 1 || H < 1 || W < 1) {
    set_error("nz_ss2d_epilogue_bwd: null pointer or empty shape");
    return NZ_EINVAL;
  }
  EpArgs a{};
  a.dout = dout, a.ym = const_cast<float*>(y_merged), a.mean = const_cast<float*>(mean), a.rstd = const_cast<float*>(rstd);
  a.z = z, a.z_bs = z_stride[0], a.z_ls = z_stride[1], a.gamma = gamma, a.beta = beta, a.d_out_y = d_out_y, a.dz = dz;
  a.dgamma = dgamma, a.dbeta = dbeta, a.B = batch, a.D = D, a.H = H, a.W = W;
  size_t smem;
  int grid;
  if (!ep_config(a, &smem, &grid, true)) {
    set_error("nz_ss2d_epilogue_bwd: D = %d unsupported (multiple of 32, 32..256)", D);
    return NZ_EUNSUPPORTED;
  }
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
#define NZ_EP_BWD(TZ, TO, TG)                                                                                \
  do {                                                                                                       \
    auto kern = ss2d_epilogue_bwd_kernel<TZ, TO, TG>;                                                        \
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);                      \
    kern<<<grid, kEpThreads, smem, st>>>(a);                                                                 \
  } while (0)
  const int key = (z_dtype * 4 + out_dtype) * 4 + grad_dtype;
  switch (key) {
    case (NZ_F32 * 4 + NZ_F32) * 4 + NZ_F32: NZ_EP_BWD(float, float, float); break;
    case (NZ_BF16 * 4 + NZ_BF16) * 4 + NZ_BF16: NZ_EP_BWD(__nv_bfloat16, __nv_bfloat16, __nv_bfloat16); break;
    case (NZ_BF16 * 4 + NZ_BF16) * 4 + NZ_F32: NZ_EP_BWD(__nv_bfloat16, __nv_bfloat16, float); break;
    case (NZ_BF16 * 4 + NZ_F32) * 4 + NZ_F32: NZ_EP_BWD(__nv_bfloat16, float, float); break;
    case (NZ_BF16 * 4 + NZ_F32) * 4 + NZ_BF16: NZ_EP_BWD(__nv_bfloat16, float, __nv_bfloat16); break;
    case (NZ_F16 * 4 + NZ_F16) * 4 + NZ_F16: NZ_EP_BWD(__half, __half, __half); break;
    case (NZ_F16 * 4 + NZ_F16) * 4 + NZ_F32: NZ_EP_BWD(__half, __half, float); break;
    case (NZ_F16 * 4 + NZ_F32) * 4 + NZ_F32: NZ_EP_BWD(__half, float, float); break;
    case (NZ_F16 * 4 + NZ_F32) * 4 + NZ_F16: NZ_EP_BWD(__half, float, __half); break;
    default:
      set_error("nz_ss2d_epilogue_bwd: unsupported dtypes (z %d, out %d, grad %d)", z_dtype, out_dtype, grad_dtype);
      return NZ_EINVAL;
  }
#undef NZ_EP_BWD
  count_launch(1);
  return cudaGetLastError() == cudaSuccess ? NZ_OK : NZ_ECUDA;
}
