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
// A CTA owns a TH x TW tile of spatial positions and all D channels of it in shared memory ([channel][TH x (TW + 1)]
// fp32 with an odd plane pitch: walking w, walking h and walking channels are all (nearly) conflict-free).
// The row-major directions (k0, k2) are read / written in runs
// of TW elements, the column-major ones (k1, k3) in runs of TH, so every direction moves whole sectors; the four values
// are added in the reference's order ((y0 + flip y2) + T y1) + T flip y3, fp32, so the merged y is bit-identical to
// nz_cross_merge.  LayerNorm / gate: D / 8 adjacent lanes per position, 8 channels (one 16-byte vector) per lane, so a
// warp works on 32 / (D / 8) positions at once and every warp makes exactly four passes per tile (TH * TW * D = 8192);
// fp32 two-pass statistics.
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <type_traits>

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

// 8 consecutive elements <-> 8 floats (16 bytes of a 16-bit type, 32 bytes of fp32)
template <typename T>
__device__ __forceinline__ void ld8(const T* __restrict__ p, float (&f)[8]) {
  if constexpr (sizeof(T) == 4) {
    const float4 a = *reinterpret_cast<const float4*>(p), b = *reinterpret_cast<const float4*>(p + 4);
    f[0] = a.x, f[1] = a.y, f[2] = a.z, f[3] = a.w, f[4] = b.x, f[5] = b.y, f[6] = b.z, f[7] = b.w;
  } else {
    const uint4 v = *reinterpret_cast<const uint4*>(p);
    const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      if constexpr (sizeof(T) == 2 && std::is_same<T, __nv_bfloat16>::value) {
        f[2 * i] = __uint_as_float(w[i] << 16), f[2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u);
      } else {
        const float2 q = __half22float2(*reinterpret_cast<const __half2*>(&w[i]));
        f[2 * i] = q.x, f[2 * i + 1] = q.y;
      }
    }
  }
}
template <typename T>
__device__ __forceinline__ void st8(T* __restrict__ p, const float (&f)[8]) {
  if constexpr (sizeof(T) == 4) {
    *reinterpret_cast<float4*>(p) = make_float4(f[0], f[1], f[2], f[3]);
    *reinterpret_cast<float4*>(p + 4) = make_float4(f[4], f[5], f[6], f[7]);
  } else {
    uint32_t w[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      if constexpr (std::is_same<T, __nv_bfloat16>::value) {
        const __nv_bfloat162 q = __floats2bfloat162_rn(f[2 * i], f[2 * i + 1]);
        w[i] = *reinterpret_cast<const uint32_t*>(&q);
      } else {
        const __half2 q = __floats2half2_rn(f[2 * i], f[2 * i + 1]);
        w[i] = *reinterpret_cast<const uint32_t*>(&q);
      }
    }
    *reinterpret_cast<uint4*>(p) = make_uint4(w[0], w[1], w[2], w[3]);
  }
}

__device__ __forceinline__ float group_sum(float v, int G) {
  for (int o = G >> 1; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

constexpr int kEpThreads = 256;
constexpr int kEpCPL = 8;     // channels per lane in the per-position phase (one 16-byte vector of a 16-bit type)
constexpr int kEpPasses = 4;  // P * D = 8192: every warp makes 4 passes over (32 / G) positions, G = D / 8 lanes each

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
  int B, D, H, W, TH, TW;  // TH, TW powers of two, TH * TW * D = 8192
  int lTH, lTW;            // their log2
  int folded;              // 1: out_y / d_out_y hold the four directions UN-flipped, ordered {row-major forward, row-major
                           // backward, column-major forward, column-major backward} as the scan's rev_mask path writes them
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
__global__ void __launch_bounds__(kEpThreads, 4) ss2d_epilogue_fwd_kernel(EpArgs a) {
  extern __shared__ float sm[];  // [D][S], S = TH * (TW + 1) made odd: position (ph, pw) sits at ph * (TW + 1) + pw
  const int D = a.D, P = a.TH * a.TW, lP = a.lTH + a.lTW, S = (a.TH * (a.TW + 1)) | 1, TW1 = a.TW + 1;
  const long L = (long)a.H * a.W;
  const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
  const int G = D / kEpCPL;           // lanes per position (4 .. 32)
  const int ppw = 32 / G;             // positions a warp handles per pass
  const int lg = lane % G, lp = lane / G;
  const int c0 = lg * kEpCPL;
  float gm[kEpCPL], bt[kEpCPL];
#pragma unroll
  for (int i = 0; i < kEpCPL; ++i) gm[i] = a.gamma ? a.gamma[c0 + i] : 1.f, bt[i] = a.beta ? a.beta[c0 + i] : 0.f;
  const TZ* __restrict__ zp = static_cast<const TZ*>(a.z);
  TO* __restrict__ outp = static_cast<TO*>(a.out);
  float* __restrict__ ymp = a.ym;
  const float invD = 1.f / (float)D;
  EpTile tl;
  for (long tile = blockIdx.x; ep_tile(a, tile, &tl); tile += gridDim.x) {
    const float* __restrict__ oy = a.out_y + tl.b * 4 * D * L;
    const long ks = (long)D * L;
    // row-major directions: y0[p] + y2[L-1-p]      (P * D / 256 = 32 trips)
#pragma unroll 8
    for (int i = t; i < P * D; i += kEpThreads) {
      const int p = i & (P - 1), d = i >> lP;
      const int h = tl.h0 + (p >> a.lTW), w = tl.w0 + (p & (a.TW - 1));
      float v = 0.f;
      if (h < a.H && w < a.W) {
        const long pos = (long)h * a.W + w;
        v = oy[d * L + pos];
        v = v + (a.folded ? oy[ks + d * L + pos] : oy[2 * ks + d * L + (L - 1 - pos)]);
      }
      sm[d * S + (p >> a.lTW) * TW1 + (p & (a.TW - 1))] = v;
    }
    __syncthreads();
    // column-major directions: + y1[w*H + h] + y3[L-1-(w*H + h)]
#pragma unroll 8
    for (int i = t; i < P * D; i += kEpThreads) {
      const int q = i & (P - 1), d = i >> lP;
      const int qh = q & (a.TH - 1), qw = q >> a.lTH;
      const int h = tl.h0 + qh, w = tl.w0 + qw;
      if (h < a.H && w < a.W) {
        const long j = (long)w * a.H + h;
        const float v1 = a.folded ? oy[2 * ks + d * L + j] : oy[ks + d * L + j];
        const float v3 = oy[3 * ks + d * L + (a.folded ? j : L - 1 - j)];
        float* s = sm + d * S + qh * TW1 + qw;
        *s = (*s + v1) + v3;
      }
    }
    __syncthreads();
    // LayerNorm over D and the SiLU(z) gate: G lanes per position, 8 channels per lane; the z vectors of two passes
    // are requested before the first of them is used
#pragma unroll
    for (int k0 = 0; k0 < kEpPasses; k0 += 2) {
      float zv[2][kEpCPL];
      long rowv[2];
#pragma unroll
      for (int kk = 0; kk < 2; ++kk) {
        const int p = ((k0 + kk) * (kEpThreads / 32) + warp) * ppw + lp;
        const int h = tl.h0 + (p >> a.lTW), w = tl.w0 + (p & (a.TW - 1));
        const bool ok = h < a.H && w < a.W;
        const long pos = ok ? (long)h * a.W + w : 0;
        rowv[kk] = ok ? tl.b * L + pos : -1;
        ld8<TZ>(zp + tl.b * a.z_bs + pos * a.z_ls + c0, zv[kk]);
      }
#pragma unroll
      for (int kk = 0; kk < 2; ++kk) {
        const int p = ((k0 + kk) * (kEpThreads / 32) + warp) * ppw + lp;
        float v[kEpCPL];
        float s = 0.f;
        const int slot = (p >> a.lTW) * TW1 + (p & (a.TW - 1));
#pragma unroll
        for (int i = 0; i < kEpCPL; ++i) v[i] = sm[(c0 + i) * S + slot], s += v[i];
        const float mu = group_sum(s, G) * invD;
        float q = 0.f;
#pragma unroll
        for (int i = 0; i < kEpCPL; ++i) q = fmaf(v[i] - mu, v[i] - mu, q);
        const float rs = rsqrtf(group_sum(q, G) * invD + a.eps);
        if (rowv[kk] >= 0) {
          float o[kEpCPL];
#pragma unroll
          for (int i = 0; i < kEpCPL; ++i) {
            const float zz = zv[kk][i];
            o[i] = fmaf((v[i] - mu) * rs, gm[i], bt[i]) * (zz / (1.f + __expf(-zz)));
          }
          st8<TO>(outp + rowv[kk] * D + c0, o);
          st8<float>(ymp + rowv[kk] * D + c0, v);
          if (lg == 0) a.mean[rowv[kk]] = mu, a.rstd[rowv[kk]] = rs;
        }
      }
    }
    __syncthreads();
  }
}

template <typename TZ, typename TO, typename TG>
__global__ void __launch_bounds__(kEpThreads, 3) ss2d_epilogue_bwd_kernel(EpArgs a) {
  extern __shared__ float sm[];  // [D][S] dy (layout as in the forward), then [2][D] parameter-gradient fold
  const int D = a.D, P = a.TH * a.TW, lP = a.lTH + a.lTW, S = (a.TH * (a.TW + 1)) | 1, TW1 = a.TW + 1;
  float* red = sm + D * S;
  const long L = (long)a.H * a.W;
  const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
  const int G = D / kEpCPL, ppw = 32 / G;
  const int lg = lane % G, lp = lane / G;
  const int c0 = lg * kEpCPL;
  float gm[kEpCPL], bt[kEpCPL], ag[kEpCPL], ab[kEpCPL];
#pragma unroll
  for (int i = 0; i < kEpCPL; ++i) {
    ag[i] = ab[i] = 0.f;
    gm[i] = a.gamma ? a.gamma[c0 + i] : 1.f, bt[i] = a.beta ? a.beta[c0 + i] : 0.f;
  }
  for (int i = t; i < 2 * D; i += kEpThreads) red[i] = 0.f;
  const TZ* __restrict__ zp = static_cast<const TZ*>(a.z);
  const TO* __restrict__ gp = static_cast<const TO*>(a.dout);
  const float* __restrict__ ymp = a.ym;
  TZ* __restrict__ dzp = static_cast<TZ*>(a.dz);
  const float invD = 1.f / (float)D;
  EpTile tl;
  for (long tile = blockIdx.x; ep_tile(a, tile, &tl); tile += gridDim.x) {
    __syncthreads();
#pragma unroll 2
    for (int k = 0; k < kEpPasses; ++k) {
      const int p = (k * (kEpThreads / 32) + warp) * ppw + lp;
      const int h = tl.h0 + (p >> a.lTW), w = tl.w0 + (p & (a.TW - 1));
      const bool ok = p < P && h < a.H && w < a.W;
      const long pos = ok ? (long)h * a.W + w : 0;
      const long row = tl.b * L + pos;
      float g[kEpCPL], zz[kEpCPL], y[kEpCPL];
      ld8<TO>(gp + row * D + c0, g);
      ld8<TZ>(zp + tl.b * a.z_bs + pos * a.z_ls + c0, zz);
      ld8<float>(ymp + row * D + c0, y);
      const float mu = a.mean[row], rs = a.rstd[row];
      float xh[kEpCPL], gg[kEpCPL], dzv[kEpCPL];
      float s1 = 0.f, s2 = 0.f;
#pragma unroll
      for (int i = 0; i < kEpCPL; ++i) {
        const float sg = 1.f / (1.f + __expf(-zz[i]));
        xh[i] = (y[i] - mu) * rs;
        const float ln = fmaf(xh[i], gm[i], bt[i]);
        dzv[i] = g[i] * ln * sg * (1.f + zz[i] * (1.f - sg));
        const float dln = ok ? g[i] * zz[i] * sg : 0.f;
        ag[i] = fmaf(dln, xh[i], ag[i]);
        ab[i] += dln;
        gg[i] = dln * gm[i];
        s1 += gg[i];
        s2 = fmaf(gg[i], xh[i], s2);
      }
      const float m1 = group_sum(s1, G) * invD, m2 = group_sum(s2, G) * invD;
      if (ok) st8<TZ>(dzp + row * D + c0, dzv);
      {
        const int slot = (p >> a.lTW) * TW1 + (p & (a.TW - 1));
#pragma unroll
        for (int i = 0; i < kEpCPL; ++i) sm[(c0 + i) * S + slot] = rs * (gg[i] - m1 - xh[i] * m2);
      }
    }
    __syncthreads();
    TG* __restrict__ go = static_cast<TG*>(a.d_out_y) + tl.b * 4 * D * L;
    const long ks = (long)D * L;
#pragma unroll 8
    for (int i = t; i < P * D; i += kEpThreads) {  // row-major copies
      const int p = i & (P - 1), d = i >> lP;
      const int h = tl.h0 + (p >> a.lTW), w = tl.w0 + (p & (a.TW - 1));
      if (h < a.H && w < a.W) {
        const long pos = (long)h * a.W + w;
        const float v = sm[d * S + (p >> a.lTW) * TW1 + (p & (a.TW - 1))];
        e_st<TG>(go + d * L + pos, v);
        if (a.folded)
          e_st<TG>(go + ks + d * L + pos, v);
        else
          e_st<TG>(go + 2 * ks + d * L + (L - 1 - pos), v);
      }
    }
#pragma unroll 8
    for (int i = t; i < P * D; i += kEpThreads) {  // column-major copies
      const int q = i & (P - 1), d = i >> lP;
      const int qh = q & (a.TH - 1), qw = q >> a.lTH;
      const int h = tl.h0 + qh, w = tl.w0 + qw;
      if (h < a.H && w < a.W) {
        const long j = (long)w * a.H + h;
        const float v = sm[d * S + qh * TW1 + qw];
        e_st<TG>(go + (a.folded ? 2 : 1) * ks + d * L + j, v);
        e_st<TG>(go + 3 * ks + d * L + (a.folded ? j : L - 1 - j), v);
      }
    }
  }
  __syncthreads();
#pragma unroll
  for (int i = 0; i < kEpCPL; ++i) {
    atomicAdd(red + c0 + i, ag[i]);
    atomicAdd(red + D + c0 + i, ab[i]);
  }
  __syncthreads();
  for (int i = t; i < D; i += kEpThreads) {
    if (a.dgamma) atomicAdd(a.dgamma + i, red[i]);
    if (a.dbeta) atomicAdd(a.dbeta + i, red[D + i]);
  }
}

static bool ep_supported(int D) { return D == 32 || D == 64 || D == 128 || D == 256; }

static bool ep_config(EpArgs& a, size_t* smem, int* grid, bool bwd) {
  if (!ep_supported(a.D)) return false;
  // positions per tile so that the tile stays near 32 KB: 256 / 128 / 64 / 32 for D = 32 / 64 / 128 / 256
  int lP = 8;
  while ((1 << lP) * a.D > 8192) --lP;  // 256 positions at D = 32 ... 32 at D = 256
  const int P = 1 << lP;
  a.lTW = lP >= 7 ? 4 : 3;
  a.lTH = lP - a.lTW;
  a.TW = 1 << a.lTW, a.TH = 1 << a.lTH;
  (void)P;
  *smem = ((size_t)a.D * ((a.TH * (a.TW + 1)) | 1) + (bwd ? 2 * a.D : 0)) * sizeof(float);
  const long tiles = (long)a.B * ((a.W + a.TW - 1) / a.TW) * ((a.H + a.TH - 1) / a.TH);
  *grid = (int)(tiles < 148L * 8 ? tiles : 148L * 8);
  return true;
}

}  // namespace nz

extern "C" int nz_ss2d_epilogue_supported(int32_t D) { return nz::ep_supported(D) ? 1 : 0; }

static int ep_fwd_impl(const float* out_y, const void* z, const int64_t* z_stride, const float* gamma, const float* beta,
                       void* out, float* y_merged, float* mean, float* rstd, int32_t z_dtype, int32_t out_dtype,
                       int32_t batch, int32_t D, int32_t H, int32_t W, float eps, void* stream, int folded) {
  using namespace nz;
  if (!out_y || !z || !z_stride || !out || !y_merged || !mean || !rstd || batch < 1 || H < 1 || W < 1) {
    set_error("nz_ss2d_epilogue_fwd: null pointer or empty shape");
    return NZ_EINVAL;
  }
  EpArgs a{};
  a.out_y = out_y, a.z = z, a.z_bs = z_stride[0], a.z_ls = z_stride[1], a.gamma = gamma, a.beta = beta, a.out = out;
  a.ym = y_merged, a.mean = mean, a.rstd = rstd, a.B = batch, a.D = D, a.H = H, a.W = W, a.eps = eps;
  a.folded = folded;
  size_t smem;
  int grid;
  if (!ep_config(a, &smem, &grid, false)) {
    set_error("nz_ss2d_epilogue_fwd: D = %d unsupported (32, 64, 128 or 256)", D);
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

static int ep_bwd_impl(const void* dout, const float* y_merged, const float* mean, const float* rstd, const void* z,
                       const int64_t* z_stride, const float* gamma, const float* beta, void* d_out_y, void* dz,
                       float* dgamma, float* dbeta, int32_t z_dtype, int32_t out_dtype, int32_t grad_dtype, int32_t batch,
                       int32_t D, int32_t H, int32_t W, void* stream, int folded) {
  using namespace nz;
  if (!dout || !y_merged || !mean || !rstd || !z || !z_stride || !d_out_y || !dz || batch < 1 || H < 1 || W < 1) {
    set_error("nz_ss2d_epilogue_bwd: null pointer or empty shape");
    return NZ_EINVAL;
  }
  EpArgs a{};
  a.dout = dout, a.ym = const_cast<float*>(y_merged), a.mean = const_cast<float*>(mean), a.rstd = const_cast<float*>(rstd);
  a.z = z, a.z_bs = z_stride[0], a.z_ls = z_stride[1], a.gamma = gamma, a.beta = beta, a.d_out_y = d_out_y, a.dz = dz;
  a.dgamma = dgamma, a.dbeta = dbeta, a.B = batch, a.D = D, a.H = H, a.W = W;
  a.folded = folded;
  size_t smem;
  int grid;
  if (!ep_config(a, &smem, &grid, true)) {
    set_error("nz_ss2d_epilogue_bwd: D = %d unsupported (32, 64, 128 or 256)", D);
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

extern "C" int nz_ss2d_epilogue_fwd(const float* out_y, const void* z, const int64_t* z_stride, const float* gamma,
                                    const float* beta, void* out, float* y_merged, float* mean, float* rstd,
                                    int32_t z_dtype, int32_t out_dtype, int32_t batch, int32_t D, int32_t H, int32_t W,
                                    float eps, void* stream) {
  return ep_fwd_impl(out_y, z, z_stride, gamma, beta, out, y_merged, mean, rstd, z_dtype, out_dtype, batch, D, H, W, eps,
                     stream, 0);
}
extern "C" int nz_ss2d_epilogue_bwd(const void* dout, const float* y_merged, const float* mean, const float* rstd,
                                    const void* z, const int64_t* z_stride, const float* gamma, const float* beta,
                                    void* d_out_y, void* dz, float* dgamma, float* dbeta, int32_t z_dtype,
                                    int32_t out_dtype, int32_t grad_dtype, int32_t batch, int32_t D, int32_t H, int32_t W,
                                    void* stream) {
  return ep_bwd_impl(dout, y_merged, mean, rstd, z, z_stride, gamma, beta, d_out_y, dz, dgamma, dbeta, z_dtype, out_dtype,
                     grad_dtype, batch, D, H, W, stream, 0);
}
// Same kernels on the folded direction layout (nz_scan_fwd / nz_scan_bwd with rev_mask = 0b1010, u_gdiv = 2):
// out_y / d_out_y = (batch, {row-major fwd, row-major bwd, column-major fwd, column-major bwd}, D, L), nothing flipped.
extern "C" int nz_ss2d_epilogue_fwd_folded(const float* out_y, const void* z, const int64_t* z_stride, const float* gamma,
                                           const float* beta, void* out, float* y_merged, float* mean, float* rstd,
                                           int32_t z_dtype, int32_t out_dtype, int32_t batch, int32_t D, int32_t H,
                                           int32_t W, float eps, void* stream) {
  return ep_fwd_impl(out_y, z, z_stride, gamma, beta, out, y_merged, mean, rstd, z_dtype, out_dtype, batch, D, H, W, eps,
                     stream, 1);
}
extern "C" int nz_ss2d_epilogue_bwd_folded(const void* dout, const float* y_merged, const float* mean, const float* rstd,
                                           const void* z, const int64_t* z_stride, const float* gamma, const float* beta,
                                           void* d_out_y, void* dz, float* dgamma, float* dbeta, int32_t z_dtype,
                                           int32_t out_dtype, int32_t grad_dtype, int32_t batch, int32_t D, int32_t H,
                                           int32_t W, void* stream) {
  return ep_bwd_impl(dout, y_merged, mean, rstd, z, z_stride, gamma, beta, d_out_y, dz, dgamma, dbeta, z_dtype, out_dtype,
                     grad_dtype, batch, D, H, W, stream, 1);
}
