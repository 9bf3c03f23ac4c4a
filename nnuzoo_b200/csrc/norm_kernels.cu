// norm_kernels.cu -- channels-last LayerNorm for the short rows of the SS2D nets.
//
// Every VSS block normalises twice (ln_1 over d_model, out_norm over d_inner: nnunetv2/nets/m2net.py:524, :220) and
// every patch merge / expand once (:241, :286-290); in M2Net that is 240 LayerNorms per step over rows of 16..1024
// channels, up to 3.1 M rows each.  Rows that short leave the library kernel (one CTA-wide reduction per row, fp32
// I/O under autocast plus the casts around it) at 24 % of a training step (profiles/r01_train_profile_*.txt).  The op is
// pure HBM streaming: read x once, write y once (+ 8 bytes of statistics per row).
//
// Layout: a row of C channels is owned by G = min(32, C / VEC) adjacent lanes (VEC = 16 bytes of elements), each lane
// holding VPL 16-byte vectors strided by G, so a warp reads whole 128-byte lines and a warp of 32 / G rows needs
// log2(G) shuffle rounds for the two moments.  fp32 statistics, two-pass variance on register-resident data.
// Backward: dx = rstd * (g - mean(g) - xhat * mean(g * xhat)), g = dy * gamma; dgamma / dbeta are accumulated per lane
// over the grid-stride row loop (a lane always owns the same channels), folded across the CTA in shared memory and
// flushed with one atomicAdd per channel per CTA (caller zeroes them).
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/nnuzoo_b200.h"

namespace nz {
void count_launch(int n);
void set_error(const char* fmt, ...);

constexpr int kNormThreads = 256;

// E elements per lane-vector: 16 bytes of the WIDER of the two element types (4 when fp32 is involved, else 8)
template <typename TI, typename TO>
struct VecE {
  static constexpr int N = 16 / (sizeof(TI) > sizeof(TO) ? sizeof(TI) : sizeof(TO));
};

template <typename T>
__device__ __forceinline__ float2 unpack2(uint32_t w);
template <>
__device__ __forceinline__ float2 unpack2<__nv_bfloat16>(uint32_t w) {
  return make_float2(__uint_as_float(w << 16), __uint_as_float(w & 0xffff0000u));
}
template <>
__device__ __forceinline__ float2 unpack2<__half>(uint32_t w) {
  return __half22float2(*reinterpret_cast<const __half2*>(&w));
}
template <typename T>
__device__ __forceinline__ uint32_t pack2(float a, float b);
template <>
__device__ __forceinline__ uint32_t pack2<__nv_bfloat16>(float a, float b) {
  const __nv_bfloat162 t = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<const uint32_t*>(&t);
}
template <>
__device__ __forceinline__ uint32_t pack2<__half>(float a, float b) {
  const __half2 t = __floats2half2_rn(a, b);
  return *reinterpret_cast<const uint32_t*>(&t);
}

template <typename T, int E>
__device__ __forceinline__ void load_vec(const T* p, float (&f)[E]) {
  if constexpr (sizeof(T) == 4) {
    static_assert(E == 4, "fp32 vectors are 4 wide");
    const float4 v = *reinterpret_cast<const float4*>(p);
    f[0] = v.x, f[1] = v.y, f[2] = v.z, f[3] = v.w;
  } else if constexpr (E == 8) {
    const uint4 v = *reinterpret_cast<const uint4*>(p);
    const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float2 t = unpack2<T>(w[i]);
      f[2 * i] = t.x, f[2 * i + 1] = t.y;
    }
  } else {
    const uint2 v = *reinterpret_cast<const uint2*>(p);
    const float2 a = unpack2<T>(v.x), b = unpack2<T>(v.y);
    f[0] = a.x, f[1] = a.y, f[2] = b.x, f[3] = b.y;
  }
}

template <typename T, int E>
__device__ __forceinline__ void store_vec(T* p, const float (&f)[E]) {
  if constexpr (sizeof(T) == 4) {
    *reinterpret_cast<float4*>(p) = make_float4(f[0], f[1], f[2], f[3]);
  } else if constexpr (E == 8) {
    *reinterpret_cast<uint4*>(p) = make_uint4(pack2<T>(f[0], f[1]), pack2<T>(f[2], f[3]), pack2<T>(f[4], f[5]),
                                              pack2<T>(f[6], f[7]));
  } else {
    *reinterpret_cast<uint2*>(p) = make_uint2(pack2<T>(f[0], f[1]), pack2<T>(f[2], f[3]));
  }
}

__device__ __forceinline__ float group_sum(float v, int G) {
  for (int o = G >> 1; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// ---------------------------------------------------------------------------------------------------------------
template <typename TI, typename TO, int VPL>
__global__ void __launch_bounds__(kNormThreads) layernorm_fwd_kernel(const TI* __restrict__ x,
                                                                     const float* __restrict__ gamma,
                                                                     const float* __restrict__ beta, TO* __restrict__ y,
                                                                     float* __restrict__ mean, float* __restrict__ rstd,
                                                                     long rows, int C, int G, float eps) {
  constexpr int VEC = VecE<TI, TO>::N;
  const int lane_g = threadIdx.x % G;                      // lane within the row group
  const int rows_per_cta = kNormThreads / G;
  const long row0 = (long)blockIdx.x * rows_per_cta + threadIdx.x / G;
  const long stride = (long)gridDim.x * rows_per_cta;
  float gm[VPL][VEC], bt[VPL][VEC];
#pragma unroll
  for (int v = 0; v < VPL; ++v) {
    const int c = (v * G + lane_g) * VEC;
#pragma unroll
    for (int i = 0; i < VEC; ++i) gm[v][i] = gamma ? gamma[c + i] : 1.f, bt[v][i] = beta ? beta[c + i] : 0.f;
  }
  const float invC = 1.f / (float)C;
  const int rsub = threadIdx.x / G;
  // the trip count is uniform over the CTA (full-mask shuffles); rows beyond the end are clamped and not stored
  for (long r = row0; r - rsub < rows; r += stride) {
    const bool live = r < rows;
    const long rr = live ? r : rows - 1;
    float xv[VPL][VEC];
    float s = 0.f;
#pragma unroll
    for (int v = 0; v < VPL; ++v) {
      load_vec<TI, VEC>(x + rr * C + (v * G + lane_g) * VEC, xv[v]);
#pragma unroll
      for (int i = 0; i < VEC; ++i) s += xv[v][i];
    }
    const float mu = group_sum(s, G) * invC;
    float q = 0.f;
#pragma unroll
    for (int v = 0; v < VPL; ++v)
#pragma unroll
      for (int i = 0; i < VEC; ++i) {
        const float d = xv[v][i] - mu;
        q = fmaf(d, d, q);
      }
    const float rs = rsqrtf(group_sum(q, G) * invC + eps);
    if (live) {
#pragma unroll
      for (int v = 0; v < VPL; ++v) {
        float o[VEC];
#pragma unroll
        for (int i = 0; i < VEC; ++i) o[i] = fmaf((xv[v][i] - mu) * rs, gm[v][i], bt[v][i]);
        store_vec<TO, VEC>(y + r * C + (v * G + lane_g) * VEC, o);
      }
      if (lane_g == 0) mean[r] = mu, rstd[r] = rs;
    }
  }
}

template <typename TI, typename TO, int VPL>
__global__ void __launch_bounds__(kNormThreads) layernorm_bwd_kernel(const TO* __restrict__ dy, const TI* __restrict__ x,
                                                                     const float* __restrict__ mean,
                                                                     const float* __restrict__ rstd,
                                                                     const float* __restrict__ gamma, TI* __restrict__ dx,
                                                                     float* __restrict__ dgamma,
                                                                     float* __restrict__ dbeta, long rows, int C, int G) {
  constexpr int VEC = VecE<TI, TO>::N;
  extern __shared__ float red[];  // [2][C] fold of the per-lane dgamma / dbeta partials
  const int lane_g = threadIdx.x % G;
  const int rows_per_cta = kNormThreads / G;
  const int rsub = threadIdx.x / G;
  const long row0 = (long)blockIdx.x * rows_per_cta + rsub;
  const long stride = (long)gridDim.x * rows_per_cta;
  float gm[VPL][VEC], ag[VPL][VEC], ab[VPL][VEC];
#pragma unroll
  for (int v = 0; v < VPL; ++v) {
    const int c = (v * G + lane_g) * VEC;
#pragma unroll
    for (int i = 0; i < VEC; ++i) gm[v][i] = gamma ? gamma[c + i] : 1.f, ag[v][i] = 0.f, ab[v][i] = 0.f;
  }
  for (int i = threadIdx.x; i < 2 * C; i += kNormThreads) red[i] = 0.f;
  __syncthreads();
  const float invC = 1.f / (float)C;
  for (long r = row0; r - rsub < rows; r += stride) {
    const bool live = r < rows;
    const long rr = live ? r : rows - 1;
    const float mu = mean[rr], rs = rstd[rr];
    float xh[VPL][VEC], g[VPL][VEC];
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int v = 0; v < VPL; ++v) {
      float xv[VEC], dv[VEC];
      load_vec<TI, VEC>(x + rr * C + (v * G + lane_g) * VEC, xv);
      load_vec<TO, VEC>(dy + rr * C + (v * G + lane_g) * VEC, dv);
#pragma unroll
      for (int i = 0; i < VEC; ++i) {
        xh[v][i] = (xv[i] - mu) * rs;
        g[v][i] = dv[i] * gm[v][i];
        s1 += g[v][i];
        s2 = fmaf(g[v][i], xh[v][i], s2);
        if (live) {
          ag[v][i] = fmaf(dv[i], xh[v][i], ag[v][i]);
          ab[v][i] += dv[i];
        }
      }
    }
    const float m1 = group_sum(s1, G) * invC, m2 = group_sum(s2, G) * invC;
    if (live) {
#pragma unroll
      for (int v = 0; v < VPL; ++v) {
        float o[VEC];
#pragma unroll
        for (int i = 0; i < VEC; ++i) o[i] = rs * (g[v][i] - m1 - xh[v][i] * m2);
        store_vec<TI, VEC>(dx + r * C + (v * G + lane_g) * VEC, o);
      }
    }
  }
  if (dgamma || dbeta) {
#pragma unroll
    for (int v = 0; v < VPL; ++v) {
      const int c = (v * G + lane_g) * VEC;
#pragma unroll
      for (int i = 0; i < VEC; ++i) {
        atomicAdd(red + c + i, ag[v][i]);
        atomicAdd(red + C + c + i, ab[v][i]);
      }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < C; i += kNormThreads) {
      if (dgamma) atomicAdd(dgamma + i, red[i]);
      if (dbeta) atomicAdd(dbeta + i, red[C + i]);
    }
  }
}

static bool norm_shape(int C, int vec, int* G, int* VPL) {
  if (C < vec || C % vec) return false;
  const int nv = C / vec;
  if (nv & (nv - 1)) return false;  // power-of-two vector count keeps the lane groups warp-aligned
  *G = nv < 32 ? nv : 32;
  *VPL = nv / *G;
  return *VPL * vec <= 32;  // at most 32 elements per lane in registers
}

static int norm_grid(long rows, int G) {
  const int rows_per_cta = kNormThreads / G;
  long blocks = (rows + rows_per_cta - 1) / rows_per_cta;
  const long cap = 148L * 8;
  if (blocks > cap) blocks = cap;
  return (int)(blocks < 1 ? 1 : blocks);
}

struct NormArgs {
  const void *x, *dy;
  const float *gamma, *beta;
  void *y, *dx;
  float *mean, *rstd, *dgamma, *dbeta;
  long rows;
  int C;
  float eps;
};

template <typename TI, typename TO, int VPL>
static void launch_one(const NormArgs& a, bool bwd, int grid, int G, cudaStream_t st) {
  if constexpr (VPL * VecE<TI, TO>::N <= 32) {
    if (!bwd)
      layernorm_fwd_kernel<TI, TO, VPL><<<grid, kNormThreads, 0, st>>>(
          static_cast<const TI*>(a.x), a.gamma, a.beta, static_cast<TO*>(a.y), a.mean, a.rstd, a.rows, a.C, G, a.eps);
    else
      layernorm_bwd_kernel<TI, TO, VPL><<<grid, kNormThreads, 2 * (size_t)a.C * sizeof(float), st>>>(
          static_cast<const TO*>(a.dy), static_cast<const TI*>(a.x), a.mean, a.rstd, a.gamma, static_cast<TI*>(a.dx),
          a.dgamma, a.dbeta, a.rows, a.C, G);
  }
}

template <typename TI, typename TO>
static int run_norm(const NormArgs& a, bool bwd, cudaStream_t st) {
  int G, VPL;
  if (!norm_shape(a.C, VecE<TI, TO>::N, &G, &VPL)) return NZ_EUNSUPPORTED;
  const int grid = norm_grid(a.rows, G);
  switch (VPL) {
    case 1: launch_one<TI, TO, 1>(a, bwd, grid, G, st); break;
    case 2: launch_one<TI, TO, 2>(a, bwd, grid, G, st); break;
    case 4: launch_one<TI, TO, 4>(a, bwd, grid, G, st); break;
    default: launch_one<TI, TO, 8>(a, bwd, grid, G, st); break;
  }
  return NZ_OK;
}

// supported (input, output) element types: equal, or one of them fp32
static int dispatch_norm(const NormArgs& a, int in_dtype, int out_dtype, bool bwd, cudaStream_t st) {
  const int key = in_dtype * 4 + out_dtype;
  switch (key) {
    case NZ_F32 * 4 + NZ_F32: return run_norm<float, float>(a, bwd, st);
    case NZ_BF16 * 4 + NZ_BF16: return run_norm<__nv_bfloat16, __nv_bfloat16>(a, bwd, st);
    case NZ_F16 * 4 + NZ_F16: return run_norm<__half, __half>(a, bwd, st);
    case NZ_BF16 * 4 + NZ_F32: return run_norm<__nv_bfloat16, float>(a, bwd, st);
    case NZ_F16 * 4 + NZ_F32: return run_norm<__half, float>(a, bwd, st);
    case NZ_F32 * 4 + NZ_BF16: return run_norm<float, __nv_bfloat16>(a, bwd, st);
    case NZ_F32 * 4 + NZ_F16: return run_norm<float, __half>(a, bwd, st);
    default: return NZ_EUNSUPPORTED;
  }
}

static int norm_vec(int in_dtype, int out_dtype) { return (in_dtype == NZ_F32 || out_dtype == NZ_F32) ? 4 : 8; }

}  // namespace nz

extern "C" int nz_layernorm_supported(int32_t C, int32_t in_dtype, int32_t out_dtype) {
  int G, VPL;
  if (in_dtype != out_dtype && in_dtype != NZ_F32 && out_dtype != NZ_F32) return 0;
  return nz::norm_shape(C, nz::norm_vec(in_dtype, out_dtype), &G, &VPL) ? 1 : 0;
}

extern "C" int nz_layernorm_fwd(const void* x, const float* gamma, const float* beta, void* y, float* mean, float* rstd,
                                int64_t rows, int32_t C, int32_t in_dtype, int32_t out_dtype, float eps, void* stream) {
  using namespace nz;
  if (!x || !y || !mean || !rstd || rows < 1 || C < 1) {
    set_error("nz_layernorm_fwd: null pointer or empty shape");
    return NZ_EINVAL;
  }
  NormArgs a{};
  a.x = x, a.gamma = gamma, a.beta = beta, a.y = y, a.mean = mean, a.rstd = rstd, a.rows = rows, a.C = C, a.eps = eps;
  const int rc = dispatch_norm(a, in_dtype, out_dtype, false, reinterpret_cast<cudaStream_t>(stream));
  if (rc != NZ_OK) {
    set_error("nz_layernorm_fwd: unsupported C = %d / dtypes (%d -> %d)", C, in_dtype, out_dtype);
    return rc;
  }
  count_launch(1);
  return cudaGetLastError() == cudaSuccess ? NZ_OK : NZ_ECUDA;
}

extern "C" int nz_layernorm_bwd(const void* dy, const void* x, const float* mean, const float* rstd, const float* gamma,
                                void* dx, float* dgamma, float* dbeta, int64_t rows, int32_t C, int32_t in_dtype,
                                int32_t out_dtype, void* stream) {
  using namespace nz;
  if (!dy || !x || !mean || !rstd || !dx || rows < 1 || C < 1) {
    set_error("nz_layernorm_bwd: null pointer or empty shape");
    return NZ_EINVAL;
  }
  NormArgs a{};
  a.dy = dy, a.x = x, a.mean = const_cast<float*>(mean), a.rstd = const_cast<float*>(rstd), a.gamma = gamma, a.dx = dx;
  a.dgamma = dgamma, a.dbeta = dbeta, a.rows = rows, a.C = C;
  const int rc = dispatch_norm(a, in_dtype, out_dtype, true, reinterpret_cast<cudaStream_t>(stream));
  if (rc != NZ_OK) {
    set_error("nz_layernorm_bwd: unsupported C = %d / dtypes (%d -> %d)", C, in_dtype, out_dtype);
    return rc;
  }
  count_launch(1);
  return cudaGetLastError() == cudaSuccess ? NZ_OK : NZ_ECUDA;
}
