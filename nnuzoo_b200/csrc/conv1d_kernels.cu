// conv1d_kernels.cu -- depthwise causal conv1d (+ SiLU) forward / backward for the 1-D Mamba path.
//
// Reference statement: the Mamba block's short convolution,
//   x = act(conv1d(x)[..., :L])            nnunetv2/nets/seg_mamba/mamba_simple.py:316-317
// (nn.Conv1d(d, d, kernel_size=W, groups=d, padding=W-1): out[b,d,l] = bias[d] + sum_k w[d,k] * x[b,d,l-(W-1)+k],
// x = 0 left of the sequence), which the reference's fast path gets from the un-vendored
// causal_conv1d_cuda.causal_conv1d_fwd / _bwd (selective_scan_interface.py:177, :247-252).
//
// HBM-bound byte work: forward reads x once and writes out once (halo re-reads hit L1/L2), backward
// reads x and dout once and writes dx once; dw / dbias are reduced per CTA and added with fp32 atomics.
// One thread = VEC = 4 consecutive steps of one (batch, channel) row; 128-bit coalesced accesses when
// the rows are 16-byte aligned, scalar accesses otherwise.  W <= 4 (nnUZoo uses d_conv = 4; SS2D's 2-D
// conv is a different operator and stays with cuDNN).
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/nnuzoo_b200.h"
#include "nz_common.cuh"

namespace nz {
void count_launch(int n);

constexpr int kConvVec = 4;     // steps per thread
constexpr int kConvMaxW = 4;    // taps
constexpr int kConvThreads = 256;

struct ConvArgs {
  const void *x, *dout;
  const float *w, *bias;  // (dim, W), (dim)
  void *out, *dx;
  float *dw, *dbias;
  long L, x_bs, x_ds, o_bs, o_ds, do_bs, do_ds, dx_bs, dx_ds;
  int batch, dim, width, silu;
  int vec;  // rows aligned for 4-element vector accesses and L % 4 == 0
};

template <typename T>
__device__ __forceinline__ float ld1(const T* p, long i, long L) {
  return (i >= 0 && i < L) ? Elem<T>::to_f(p[i]) : 0.f;
}

// four consecutive elements with one 16-byte (fp32) / 8-byte (16-bit) access
template <typename T>
__device__ __forceinline__ void load4(const T* p, float* v);
template <>
__device__ __forceinline__ void load4<float>(const float* p, float* v) {
  const float4 q = *reinterpret_cast<const float4*>(p);
  v[0] = q.x, v[1] = q.y, v[2] = q.z, v[3] = q.w;
}
template <>
__device__ __forceinline__ void load4<__nv_bfloat16>(const __nv_bfloat16* p, float* v) {
  const uint2 q = *reinterpret_cast<const uint2*>(p);
  v[0] = __uint_as_float(q.x << 16), v[1] = __uint_as_float(q.x & 0xffff0000u);
  v[2] = __uint_as_float(q.y << 16), v[3] = __uint_as_float(q.y & 0xffff0000u);
}
template <>
__device__ __forceinline__ void load4<__half>(const __half* p, float* v) {
  const uint2 q = *reinterpret_cast<const uint2*>(p);
  const float2 a = __half22float2(*reinterpret_cast<const __half2*>(&q.x));
  const float2 b = __half22float2(*reinterpret_cast<const __half2*>(&q.y));
  v[0] = a.x, v[1] = a.y, v[2] = b.x, v[3] = b.y;
}
template <typename T>
__device__ __forceinline__ void store4(T* p, const float* v);
template <>
__device__ __forceinline__ void store4<float>(float* p, const float* v) {
  *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
}
template <>
__device__ __forceinline__ void store4<__nv_bfloat16>(__nv_bfloat16* p, const float* v) {
  const __nv_bfloat162 a = __floats2bfloat162_rn(v[0], v[1]), b = __floats2bfloat162_rn(v[2], v[3]);
  *reinterpret_cast<uint2*>(p) = make_uint2(*reinterpret_cast<const uint32_t*>(&a), *reinterpret_cast<const uint32_t*>(&b));
}
template <>
__device__ __forceinline__ void store4<__half>(__half* p, const float* v) {
  const __half2 a = __floats2half2_rn(v[0], v[1]), b = __floats2half2_rn(v[2], v[3]);
  *reinterpret_cast<uint2*>(p) = make_uint2(*reinterpret_cast<const uint32_t*>(&a), *reinterpret_cast<const uint32_t*>(&b));
}

// load steps [l0 - 4*CL, l0 + 4*(1 + CR)) of one row into v[] (zero outside [0, L)); `vec`: the row is
// aligned for 4-element accesses and L % 4 == 0, so every 4-chunk is either fully inside or fully outside
template <typename T, int CL, int CR>
__device__ __forceinline__ void load_window(const T* row, long l0, long L, bool vec, float (&v)[4 * (CL + 1 + CR)]) {
#pragma unroll
  for (int cch = 0; cch < CL + 1 + CR; ++cch) {
    const long p0 = l0 + 4 * (cch - CL);
    if (vec) {
      if (p0 >= 0 && p0 < L) {
        load4<T>(row + p0, &v[4 * cch]);
      } else {
#pragma unroll
        for (int i = 0; i < 4; ++i) v[4 * cch + i] = 0.f;
      }
    } else {
#pragma unroll
      for (int i = 0; i < 4; ++i) v[4 * cch + i] = ld1(row, p0 + i, L);
    }
  }
}

__device__ __forceinline__ float silu_f(float p) { return p * sigmoid_f(p); }
__device__ __forceinline__ float dsilu_f(float p) {
  const float s = sigmoid_f(p);
  return s * (1.f + p * (1.f - s));
}

// Each thread owns CH consecutive 4-step chunks of one row (CH * 16 bytes of fp32, CH * 8 bytes of 16-bit data
// per access stream) and slides the halo through registers, so every element is loaded once per thread.
template <typename T>
struct ConvCh {
  static constexpr int CH = sizeof(T) == 4 ? 2 : 4;
  static constexpr int STEPS = CH * kConvVec;  // steps per thread
};

// REV = the same convolution run over the L-flipped sequence without materialising the flip (the reversed Mamba
// directions: mamba_simple.py:250-262 `xz.flip([-1])`, mamba_nd2net.py:638-641 `hidden_states.flip(1)`):
//   out[l] = act(bias + sum_k w[k] x[l + (W-1) - k]),  x = 0 right of the sequence.
// In the thread's mirrored step index i' = STEPS-1-i this IS the causal convolution, so the window is loaded from the
// other side and reversed in registers (static indices, free), the arithmetic below is shared, and the results are
// written back un-mirrored.
template <int N>
__device__ __forceinline__ void reverse_regs(float (&v)[N]) {
#pragma unroll
  for (int j = 0; j < N / 2; ++j) {
    const float t = v[j];
    v[j] = v[N - 1 - j];
    v[N - 1 - j] = t;
  }
}

template <typename T, bool REV>
__global__ void __launch_bounds__(kConvThreads) conv1d_fwd_kernel(const ConvArgs a) {
  constexpr int CH = ConvCh<T>::CH, STEPS = ConvCh<T>::STEPS;
  const long nblk_l = (a.L + kConvThreads * STEPS - 1) / (kConvThreads * STEPS);
  const long row = blockIdx.x / nblk_l;  // (batch, channel)
  const long l0 = ((blockIdx.x % nblk_l) * kConvThreads + threadIdx.x) * STEPS;
  if (l0 >= a.L) return;
  const int b = (int)(row / a.dim), d = (int)(row % a.dim);
  const T* xr = reinterpret_cast<const T*>(a.x) + (long)b * a.x_bs + (long)d * a.x_ds;
  T* orow = reinterpret_cast<T*>(a.out) + (long)b * a.o_bs + (long)d * a.o_ds;
  float w[kConvMaxW];
#pragma unroll
  for (int k = 0; k < kConvMaxW; ++k) w[k] = k < a.width ? __ldg(a.w + (long)d * a.width + k) : 0.f;
  const float bias = a.bias ? __ldg(a.bias + d) : 0.f;
  float v[4 * (CH + 1)];  // x[l0 - 4 .. l0 + STEPS); REV: x[l0 .. l0 + STEPS + 4) mirrored
  if (!REV) {
    load_window<T, 1, CH - 1>(xr, l0, a.L, a.vec != 0, v);
  } else {
    load_window<T, 0, CH>(xr, l0, a.L, a.vec != 0, v);
    reverse_regs(v);
  }
  const int sh = kConvMaxW - a.width;  // taps are right-aligned: tap k multiplies x[l - (W-1) + k]
  float o[STEPS];
#pragma unroll
  for (int i = 0; i < STEPS; ++i) {
    float p = bias;
#pragma unroll
    for (int k = 0; k < kConvMaxW; ++k) {
      const int kk = k - sh;  // index into w when width < kConvMaxW
      if (kk >= 0) p = fmaf(w[kk], v[1 + i + k], p);
    }
    o[i] = a.silu ? silu_f(p) : p;
  }
  if (REV) reverse_regs(o);
#pragma unroll
  for (int cch = 0; cch < CH; ++cch) {
    const long lc = l0 + 4 * cch;
    if (a.vec) {
      if (lc < a.L) store4<T>(orow + lc, &o[4 * cch]);
    } else {
#pragma unroll
      for (int i = 0; i < kConvVec; ++i)
        if (lc + i < a.L) orow[lc + i] = Elem<T>::from_f(o[4 * cch + i]);
    }
  }
}

template <typename T, bool REV>
__global__ void __launch_bounds__(kConvThreads) conv1d_bwd_kernel(const ConvArgs a) {
  constexpr int CH = ConvCh<T>::CH, STEPS = ConvCh<T>::STEPS;
  __shared__ float red[kConvThreads / 32][kConvMaxW + 1];
  const long nblk_l = (a.L + kConvThreads * STEPS - 1) / (kConvThreads * STEPS);
  const long row = blockIdx.x / nblk_l;
  const long l0 = ((blockIdx.x % nblk_l) * kConvThreads + threadIdx.x) * STEPS;
  const int b = (int)(row / a.dim), d = (int)(row % a.dim);
  const T* xr = reinterpret_cast<const T*>(a.x) + (long)b * a.x_bs + (long)d * a.x_ds;
  const T* gr = reinterpret_cast<const T*>(a.dout) + (long)b * a.do_bs + (long)d * a.do_ds;
  T* dxr = reinterpret_cast<T*>(a.dx) + (long)b * a.dx_bs + (long)d * a.dx_ds;
  float w[kConvMaxW];
#pragma unroll
  for (int k = 0; k < kConvMaxW; ++k) w[k] = k < a.width ? __ldg(a.w + (long)d * a.width + k) : 0.f;
  const float bias = a.bias ? __ldg(a.bias + d) : 0.f;
  const int sh = kConvMaxW - a.width;
  constexpr int HR = kConvMaxW - 1;                      // dx[l] needs dy[l .. l + W-1]
  float dwacc[kConvMaxW] = {0.f, 0.f, 0.f, 0.f}, dbacc = 0.f;
  if (l0 < a.L) {
    float xw[4 * (CH + 2)];                                // x[l0 - 4 .. l0 + STEPS + 4)
    load_window<T, 1, CH>(xr, l0, a.L, a.vec != 0, xw);
    if (REV) reverse_regs(xw);                             // mirrored step index: the same causal arithmetic below
    const float* xv = xw + 1;                              // xv[j] = x[l0 - (W_max-1) + j]
    float gw[4 * (CH + 1)];                                // dout[l0 .. l0 + STEPS + 4); REV: [l0 - 4 .. l0 + STEPS) mirrored
    if (!REV) {
      load_window<T, 0, CH>(gr, l0, a.L, a.vec != 0, gw);
    } else {
      load_window<T, 1, CH - 1>(gr, l0, a.L, a.vec != 0, gw);
      reverse_regs(gw);
    }
    float dy[STEPS + HR];                                  // dy[l0 .. l0 + STEPS + W-1)
#pragma unroll
    for (int i = 0; i < STEPS + HR; ++i) {
      float g = gw[i];
      if (a.silu) {  // recompute the pre-activation at step l0 + i
        float p = bias;
#pragma unroll
        for (int k = 0; k < kConvMaxW; ++k) {
          const int kk = k - sh;
          if (kk >= 0) p = fmaf(w[kk], xv[i + k], p);
        }
        g *= dsilu_f(p);
      }
      dy[i] = g;
    }
    float dxo[STEPS];
#pragma unroll
    for (int i = 0; i < STEPS; ++i) {
      // dx[l] = sum_k w[k] * dy[l + (W-1) - k]
      float acc = 0.f;
#pragma unroll
      for (int k = 0; k < kConvMaxW; ++k) {
        const int kk = k - sh;
        if (kk >= 0) acc = fmaf(w[kk], dy[i + (kConvMaxW - 1) - k], acc);
      }
      dxo[i] = acc;
      // dw[k] += dy[l] * x[l - (W-1) + k],  dbias += dy[l]   (this thread owns steps l0 .. l0 + STEPS)
      dbacc += dy[i];
#pragma unroll
      for (int k = 0; k < kConvMaxW; ++k) dwacc[k] = fmaf(dy[i], xv[i + k], dwacc[k]);
    }
    if (REV) reverse_regs(dxo);
#pragma unroll
    for (int cch = 0; cch < CH; ++cch) {
      const long lc = l0 + 4 * cch;
      if (a.vec) {
        if (lc < a.L) store4<T>(dxr + lc, &dxo[4 * cch]);
      } else {
#pragma unroll
        for (int ii = 0; ii < kConvVec; ++ii)
          if (lc + ii < a.L) dxr[lc + ii] = Elem<T>::from_f(dxo[4 * cch + ii]);
      }
    }
  }
  // CTA reduction of dw / dbias, then one atomic per (channel, tap) per CTA
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) {
    dbacc += __shfl_xor_sync(0xffffffffu, dbacc, off);
#pragma unroll
    for (int k = 0; k < kConvMaxW; ++k) dwacc[k] += __shfl_xor_sync(0xffffffffu, dwacc[k], off);
  }
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane == 0) {
#pragma unroll
    for (int k = 0; k < kConvMaxW; ++k) red[warp][k] = dwacc[k];
    red[warp][kConvMaxW] = dbacc;
  }
  __syncthreads();
  if (threadIdx.x <= kConvMaxW) {
    float s = 0.f;
#pragma unroll
    for (int wv = 0; wv < kConvThreads / 32; ++wv) s += red[wv][threadIdx.x];
    if (threadIdx.x == kConvMaxW) {
      if (a.dbias) atomicAdd(a.dbias + d, s);
    } else {
      const int kk = (int)threadIdx.x - sh;
      if (kk >= 0 && a.dw) atomicAdd(a.dw + (long)d * a.width + kk, s);
    }
  }
}

static int conv_validate(const NzConv1dDesc* c, bool bwd) {
  if (!c || c->batch < 1 || c->dim < 1 || c->seqlen < 1 || c->width < 1 || c->width > kConvMaxW) return NZ_EINVAL;
  if (c->dtype != NZ_F32 && c->dtype != NZ_BF16 && c->dtype != NZ_F16) return NZ_EINVAL;
  if (!c->x || !c->weight) return NZ_EINVAL;
  if (!bwd && !c->out) return NZ_EINVAL;
  if (bwd && (!c->dout || !c->dx)) return NZ_EINVAL;
  if ((long long)c->batch * c->dim * ((c->seqlen + 1023) / 1024) > 2000000000LL) return NZ_EINVAL;
  return NZ_OK;
}

static void conv_fill(const NzConv1dDesc* c, ConvArgs& a) {
  a.x = c->x; a.dout = c->dout; a.w = c->weight; a.bias = c->bias; a.out = c->out; a.dx = c->dx;
  a.dw = c->dweight; a.dbias = c->dbias;
  a.L = c->seqlen;
  a.x_bs = c->x_stride[0]; a.x_ds = c->x_stride[1];
  a.o_bs = c->out_stride[0]; a.o_ds = c->out_stride[1];
  a.do_bs = c->dout_stride[0]; a.do_ds = c->dout_stride[1];
  a.dx_bs = (long)c->dim * c->seqlen; a.dx_ds = c->seqlen;
  a.batch = c->batch; a.dim = c->dim; a.width = c->width; a.silu = c->silu;
  const size_t es = c->dtype == NZ_F32 ? 4 : 2, vb = 4 * es;
  auto ok = [&](const void* p, long bs, long ds) {
    return !p || ((reinterpret_cast<uintptr_t>(p) % vb) == 0 && (bs % 4) == 0 && (ds % 4) == 0);
  };
  a.vec = (a.L % 4 == 0) && ok(a.x, a.x_bs, a.x_ds) && ok(a.out, a.o_bs, a.o_ds) && ok(a.dout, a.do_bs, a.do_ds) &&
          ok(a.dx, a.dx_bs, a.dx_ds);
}

}  // namespace nz

extern "C" {

int64_t nz_sizeof_conv1d_desc(void) { return (int64_t)sizeof(NzConv1dDesc); }

int nz_causal_conv1d_fwd(const NzConv1dDesc* c, void* stream) {
  if (int rc = nz::conv_validate(c, false)) return rc;
  nz::ConvArgs a;
  nz::conv_fill(c, a);
  const int steps = c->dtype == NZ_F32 ? nz::ConvCh<float>::STEPS : nz::ConvCh<__half>::STEPS;
  const long nblk_l = (a.L + nz::kConvThreads * steps - 1) / (nz::kConvThreads * steps);
  const unsigned grid = (unsigned)((long)a.batch * a.dim * nblk_l);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (c->reverse) {
    if (c->dtype == NZ_F32) nz::conv1d_fwd_kernel<float, true><<<grid, nz::kConvThreads, 0, st>>>(a);
    else if (c->dtype == NZ_BF16) nz::conv1d_fwd_kernel<__nv_bfloat16, true><<<grid, nz::kConvThreads, 0, st>>>(a);
    else nz::conv1d_fwd_kernel<__half, true><<<grid, nz::kConvThreads, 0, st>>>(a);
  } else {
    if (c->dtype == NZ_F32) nz::conv1d_fwd_kernel<float, false><<<grid, nz::kConvThreads, 0, st>>>(a);
    else if (c->dtype == NZ_BF16) nz::conv1d_fwd_kernel<__nv_bfloat16, false><<<grid, nz::kConvThreads, 0, st>>>(a);
    else nz::conv1d_fwd_kernel<__half, false><<<grid, nz::kConvThreads, 0, st>>>(a);
  }
  nz::count_launch(1);
  return cudaGetLastError() == cudaSuccess ? NZ_OK : NZ_ECUDA;
}

int nz_causal_conv1d_bwd(const NzConv1dDesc* c, void* stream) {
  if (int rc = nz::conv_validate(c, true)) return rc;
  nz::ConvArgs a;
  nz::conv_fill(c, a);
  const int steps = c->dtype == NZ_F32 ? nz::ConvCh<float>::STEPS : nz::ConvCh<__half>::STEPS;
  const long nblk_l = (a.L + nz::kConvThreads * steps - 1) / (nz::kConvThreads * steps);
  const unsigned grid = (unsigned)((long)a.batch * a.dim * nblk_l);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (c->reverse) {
    if (c->dtype == NZ_F32) nz::conv1d_bwd_kernel<float, true><<<grid, nz::kConvThreads, 0, st>>>(a);
    else if (c->dtype == NZ_BF16) nz::conv1d_bwd_kernel<__nv_bfloat16, true><<<grid, nz::kConvThreads, 0, st>>>(a);
    else nz::conv1d_bwd_kernel<__half, true><<<grid, nz::kConvThreads, 0, st>>>(a);
  } else {
    if (c->dtype == NZ_F32) nz::conv1d_bwd_kernel<float, false><<<grid, nz::kConvThreads, 0, st>>>(a);
    else if (c->dtype == NZ_BF16) nz::conv1d_bwd_kernel<__nv_bfloat16, false><<<grid, nz::kConvThreads, 0, st>>>(a);
    else nz::conv1d_bwd_kernel<__half, false><<<grid, nz::kConvThreads, 0, st>>>(a);
  }
  nz::count_launch(1);
  return cudaGetLastError() == cudaSuccess ? NZ_OK : NZ_ECUDA;
}

}  // extern "C"
