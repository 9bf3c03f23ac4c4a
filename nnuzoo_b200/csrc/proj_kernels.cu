// proj_kernels.cu -- weight gradients of the per-direction SS2D projections.
//
// SS2D feeds the scan through two grouped 1x1 projections written as einsums in the reference:
//   x_dbl = einsum("b k d l, k c d -> b k c l", xs, x_proj_weight)        nnunetv2/nets/m2net.py:179
//   dts   = einsum("b k r l, k d r -> b k d l", dts, dt_projs_weight)     nnunetv2/nets/m2net.py:182
// Their weight gradients are reductions over (b, l) with tiny outputs,
//   dW[k, m, n] = sum_{b, l} G[b, k, m, l] * X[b, k, n, l]      (m, n) = (R+2N, D) resp. (D, R),
// i.e. a batched "NT" GEMM whose contraction runs over B*L = 3.1 M positions while M x N is 33 x 32.  The library
// GEMM picked for that shape runs ONE 64x64 tile per direction (4 CTAs on 148 SMs): measured 38 ms per call and
// 48 % of a whole M2Net training step (profiles/r01_train_profile_before_wgrad.txt).  The op is HBM-bound by
// rights: every G and X element is read once.
//
// Kernel: the (b, l) range is cut into chunks, one persistent CTA per (b, k, chunk).  A CTA stages 64 positions of all
// M + N rows in shared memory (fp32, rows padded to 68 words so 16-byte reads of 32 different rows are
// conflict-free per quarter warp), every thread keeps up to kMaxOut outputs (m, n) in registers and walks the 64
// positions with 16-byte shared loads (G is a warp broadcast when N >= 32).  Partial sums leave through one fp32
// atomicAdd per output per CTA (the caller zeroes dW).  fp32 accumulation whatever the I/O type.
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdlib.h>

#include "../../include/nnuzoo_b200.h"

namespace nz {
void count_launch(int n);
void set_error(const char* fmt, ...);

constexpr int kProjTL = 64;    // positions per shared-memory stage
constexpr int kProjTLP = 68;   // padded row length (words)
constexpr int kProjThreads = 256;
constexpr int kProjMaxOut = 40;  // outputs per thread: M*N <= 256 * 40

template <typename T>
__device__ __forceinline__ float to_f32(T v);
template <>
__device__ __forceinline__ float to_f32<float>(float v) { return v; }
template <>
__device__ __forceinline__ float to_f32<__nv_bfloat16>(__nv_bfloat16 v) { return __bfloat162float(v); }
template <>
__device__ __forceinline__ float to_f32<__half>(__half v) { return __half2float(v); }

struct ProjWgradArgs {
  const void* G;
  const void* X;
  float* dW;
  int B, K, M, N;
  long L;
  long gs_b, gs_k, gs_m;  // element strides of G (innermost 1)
  long xs_b, xs_k, xs_n;  // element strides of X
  int chunks;             // CTAs per (b, k)
  long tiles_per_chunk;
};

template <typename TG, typename TX, int NOUT>
__global__ void __launch_bounds__(kProjThreads) proj_wgrad_kernel(ProjWgradArgs a) {
  extern __shared__ float sm[];
  float* Gs = sm;                        // [M][kProjTLP]
  float* Xs = sm + (long)a.M * kProjTLP;  // [N][kProjTLP]
  const int t = threadIdx.x;
  const int chunk = blockIdx.x;
  const int bk = blockIdx.y;
  const int b = bk / a.K, k = bk % a.K;
  const TG* G = static_cast<const TG*>(a.G) + b * a.gs_b + k * a.gs_k;
  const TX* X = static_cast<const TX*>(a.X) + b * a.xs_b + k * a.xs_k;
  const int MN = a.M * a.N;

  float acc[NOUT];
  uint32_t off[NOUT];  // shared-memory word offsets of the two rows, 16 bits each (<= 296 * 68)
#pragma unroll
  for (int j = 0; j < NOUT; ++j) {
    acc[j] = 0.f;
    int idx = t + j * kProjThreads;
    if (idx >= MN) idx = MN - 1;  // clamped duplicates are discarded at the end
    off[j] = (uint32_t)((idx / a.N) * kProjTLP) | ((uint32_t)((idx % a.N) * kProjTLP) << 16);
  }

  const long tile0 = chunk * a.tiles_per_chunk;
  const long ntiles = (a.L + kProjTL - 1) / kProjTL;
  long tile1 = tile0 + a.tiles_per_chunk;
  if (tile1 > ntiles) tile1 = ntiles;
  const int rows = a.M + a.N;
  for (long tile = tile0; tile < tile1; ++tile) {
    const long l0 = tile * kProjTL;
    // stage (M + N) rows x 64 positions, zero-filled past L
    for (int i = t; i < rows * kProjTL; i += kProjThreads) {
      const int r = i / kProjTL, c = i % kProjTL;
      const long l = l0 + c;
      float v = 0.f;
      if (l < a.L) v = r < a.M ? to_f32<TG>(G[r * a.gs_m + l]) : to_f32<TX>(X[(r - a.M) * a.xs_n + l]);
      (r < a.M ? Gs + r * kProjTLP : Xs + (r - a.M) * kProjTLP)[c] = v;
    }
    __syncthreads();
#pragma unroll
    for (int j = 0; j < NOUT; ++j) {
      if (j * kProjThreads < MN) {  // uniform over the CTA
        const float4* g4 = reinterpret_cast<const float4*>(Gs + (off[j] & 0xffffu));
        const float4* x4 = reinterpret_cast<const float4*>(Xs + (off[j] >> 16));
        float s0 = 0.f, s1 = 0.f;
#pragma unroll
        for (int q = 0; q < kProjTL / 4; q += 2) {
          const float4 g0 = g4[q], x0 = x4[q], g1 = g4[q + 1], x1 = x4[q + 1];
          s0 = fmaf(g0.x, x0.x, s0);
          s0 = fmaf(g0.y, x0.y, s0);
          s0 = fmaf(g0.z, x0.z, s0);
          s0 = fmaf(g0.w, x0.w, s0);
          s1 = fmaf(g1.x, x1.x, s1);
          s1 = fmaf(g1.y, x1.y, s1);
          s1 = fmaf(g1.z, x1.z, s1);
          s1 = fmaf(g1.w, x1.w, s1);
        }
        acc[j] += s0 + s1;
      }
    }
    __syncthreads();
  }
  float* out = a.dW + (long)k * MN;
#pragma unroll
  for (int j = 0; j < NOUT; ++j) {
    const int idx = t + j * kProjThreads;
    if (idx < MN) atomicAdd(out + idx, acc[j]);
  }
}

// ---------------------------------------------------------------------------------------------------------------
// bf16 x bf16 path: the same reduction on the tensor cores (mma.sync m16n8k16, fp32 accumulate).  The op stays
// HBM-bound -- per 64 positions a CTA moves (M + N) * 128 bytes and issues MT * NT * 4 MMAs -- so the kernel is a
// copy pipeline: 16-byte cp.async of whole bf16 row segments into a 3-stage ring (rows XOR-swizzled by 16-byte chunk
// so ldmatrix is conflict-free), ldmatrix + mma on the oldest stage.  MMA tiles are dealt round-robin to the warps.
constexpr int kTcStages = 3;
constexpr int kTcRowB = kProjTL * 2;  // bytes of one staged row (64 bf16)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

template <int TPW, int WARPS>
__global__ void __launch_bounds__(WARPS * 32) proj_wgrad_tc_kernel(ProjWgradArgs a, int MT, int NT) {
  extern __shared__ __align__(128) unsigned char smraw[];
  const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
  const int Mp = MT * 16, Np = NT * 8, rows = Mp + Np;
  const int stageB = rows * kTcRowB;
  const int bk = blockIdx.y, b = bk / a.K, k = bk % a.K;
  const __nv_bfloat16* G = static_cast<const __nv_bfloat16*>(a.G) + b * a.gs_b + k * a.gs_k;
  const __nv_bfloat16* X = static_cast<const __nv_bfloat16*>(a.X) + b * a.xs_b + k * a.xs_k;
  // padding rows are never written by the copies: zero the whole ring once
  for (int i = t; i < kTcStages * stageB / 16; i += WARPS * 32) reinterpret_cast<uint4*>(smraw)[i] = make_uint4(0, 0, 0, 0);
  __syncthreads();

  const long ntiles = a.L / kProjTL;
  const long tile0 = blockIdx.x * a.tiles_per_chunk;
  long tile1 = tile0 + a.tiles_per_chunk;
  if (tile1 > ntiles) tile1 = ntiles;
  const int nt_local = (int)(tile1 - tile0);

  auto issue = [&](int it) {  // stage tile `it` of this CTA into ring slot it % kTcStages
    if (it < nt_local) {
      unsigned char* st = smraw + (it % kTcStages) * stageB;
      const long l0 = (tile0 + it) * kProjTL;
      for (int i = t; i < (a.M + a.N) * 8; i += WARPS * 32) {
        const int r = i >> 3, c = i & 7;
        const bool isg = r < a.M;
        const int srow = isg ? r : Mp + (r - a.M);
        const __nv_bfloat16* src = isg ? G + r * a.gs_m + l0 + c * 8 : X + (r - a.M) * a.xs_n + l0 + c * 8;
        const uint32_t dst = smem_u32(st + srow * kTcRowB + ((c ^ (srow & 7)) << 4));
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src));
      }
    }
    asm volatile("cp.async.commit_group;");
  };

  float acc[TPW][4];
#pragma unroll
  for (int j = 0; j < TPW; ++j) acc[j][0] = acc[j][1] = acc[j][2] = acc[j][3] = 0.f;
  const int ntile_mma = MT * NT;

  issue(0);
  issue(1);
  for (int it = 0; it < nt_local; ++it) {
    issue(it + 2);
    asm volatile("cp.async.wait_group 2;");
    __syncthreads();
    const unsigned char* st = smraw + (it % kTcStages) * stageB;
#pragma unroll
    for (int j = 0; j < TPW; ++j) {
      const int tid = warp + j * WARPS;
      if (tid < ntile_mma) {
        const int mt = tid % MT, nt = tid / MT;
        const int arow = mt * 16 + (lane & 7) + ((lane >> 3) & 1) * 8;
        const int brow = Mp + nt * 8 + (lane & 7);
#pragma unroll
        for (int ks = 0; ks < kProjTL / 16; ++ks) {
          uint32_t a0, a1, a2, a3, b0, b1;
          const int ac = ks * 2 + (lane >> 4);
          const int bc = ks * 2 + ((lane >> 3) & 1);
          asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
                       : "=r"(a0), "=r"(a1), "=r"(a2), "=r"(a3)
                       : "r"(smem_u32(st + arow * kTcRowB + ((ac ^ (arow & 7)) << 4))));
          asm volatile("ldmatrix.sync.aligned.m8n8.x2.shared.b16 {%0,%1}, [%2];"
                       : "=r"(b0), "=r"(b1)
                       : "r"(smem_u32(st + brow * kTcRowB + ((bc ^ (brow & 7)) << 4))));
          asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                       : "+f"(acc[j][0]), "+f"(acc[j][1]), "+f"(acc[j][2]), "+f"(acc[j][3])
                       : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
        }
      }
    }
    __syncthreads();  // the slot is refilled by the next iteration's issue()
  }
  asm volatile("cp.async.wait_group 0;");
  float* out = a.dW + (long)k * a.M * a.N;
#pragma unroll
  for (int j = 0; j < TPW; ++j) {
    const int tid = warp + j * WARPS;
    if (tid < ntile_mma) {
      const int mt = tid % MT, nt = tid / MT;
      const int m0 = mt * 16 + (lane >> 2), n0 = nt * 8 + (lane & 3) * 2;
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const int m = m0 + (q >> 1) * 8, n = n0 + (q & 1);
        if (m < a.M && n < a.N) atomicAdd(out + m * a.N + n, acc[j][q]);
      }
    }
  }
}

static bool tc_eligible(const ProjWgradArgs& a) {
  auto al = [](const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; };
  return a.L % kProjTL == 0 && al(a.G) && al(a.X) && a.gs_b % 8 == 0 && a.gs_k % 8 == 0 && a.gs_m % 8 == 0 &&
         a.xs_b % 8 == 0 && a.xs_k % 8 == 0 && a.xs_n % 8 == 0;
}

static int launch_wgrad_tc(const ProjWgradArgs& a, cudaStream_t st) {
  const int MT = (a.M + 15) / 16, NT = (a.N + 7) / 8;
  const int tiles = MT * NT;
  const size_t smem = (size_t)kTcStages * (MT * 16 + NT * 8) * kTcRowB;
  dim3 grid(a.chunks, a.B * a.K);
#define NZ_TC_LAUNCH(TPW, WARPS)                                                                                  \
  do {                                                                                                            \
    auto kern = proj_wgrad_tc_kernel<TPW, WARPS>;                                                                 \
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) {      \
      set_error("proj_wgrad_tc: cannot reserve %zu bytes of shared memory", smem);                                \
      return NZ_ECUDA;                                                                                            \
    }                                                                                                             \
    kern<<<grid, WARPS * 32, smem, st>>>(a, MT, NT);                                                              \
  } while (0)
  if (tiles <= 4)
    NZ_TC_LAUNCH(1, 4);
  else if (tiles <= 12)
    NZ_TC_LAUNCH(3, 4);
  else if (tiles <= 24)
    NZ_TC_LAUNCH(3, 8);
  else if (tiles <= 48)
    NZ_TC_LAUNCH(6, 8);
  else if (tiles <= 96)
    NZ_TC_LAUNCH(12, 8);
  else
    return NZ_EUNSUPPORTED;
#undef NZ_TC_LAUNCH
  count_launch(1);
  if (cudaGetLastError() != cudaSuccess) {
    set_error("proj_wgrad_tc launch failed");
    return NZ_ECUDA;
  }
  return NZ_OK;
}

template <typename TG, typename TX>
static int launch_wgrad(const ProjWgradArgs& a, cudaStream_t st) {
  const size_t smem = (size_t)(a.M + a.N) * kProjTLP * sizeof(float);
  const int MN = a.M * a.N;
  dim3 grid(a.chunks, a.B * a.K);
#define NZ_PROJ_LAUNCH(NOUT)                                                                                      \
  do {                                                                                                            \
    auto kern = proj_wgrad_kernel<TG, TX, NOUT>;                                                                  \
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) {      \
      set_error("proj_wgrad: cannot reserve %zu bytes of shared memory", smem);                                   \
      return NZ_ECUDA;                                                                                            \
    }                                                                                                             \
    kern<<<grid, kProjThreads, smem, st>>>(a);                                                                    \
  } while (0)
  if (MN <= 4 * kProjThreads)
    NZ_PROJ_LAUNCH(4);
  else if (MN <= 10 * kProjThreads)
    NZ_PROJ_LAUNCH(10);
  else if (MN <= 20 * kProjThreads)
    NZ_PROJ_LAUNCH(20);
  else
    NZ_PROJ_LAUNCH(kProjMaxOut);
#undef NZ_PROJ_LAUNCH
  count_launch(1);
  if (cudaGetLastError() != cudaSuccess) {
    set_error("proj_wgrad launch failed");
    return NZ_ECUDA;
  }
  return NZ_OK;
}

}  // namespace nz

extern "C" int nz_proj_wgrad(const void* G, const void* X, float* dW, int32_t g_dtype, int32_t x_dtype, int32_t batch,
                             int32_t K, int32_t M, int32_t N, int64_t L, const int64_t* g_stride,
                             const int64_t* x_stride, void* stream) {
  using namespace nz;
  if (!G || !X || !dW || !g_stride || !x_stride || batch < 1 || K < 1 || M < 1 || N < 1 || L < 1) {
    set_error("nz_proj_wgrad: null pointer or non-positive size");
    return NZ_EINVAL;
  }
  if ((long)M * N > (long)kProjThreads * kProjMaxOut || (size_t)(M + N) * kProjTLP * 4 > 200 * 1024) {
    set_error("nz_proj_wgrad: M*N = %ld exceeds %d (SS2D needs at most 40 x 256)", (long)M * N,
              kProjThreads * kProjMaxOut);
    return NZ_EUNSUPPORTED;
  }
  ProjWgradArgs a;
  a.G = G, a.X = X, a.dW = dW, a.B = batch, a.K = K, a.M = M, a.N = N, a.L = L;
  a.gs_b = g_stride[0], a.gs_k = g_stride[1], a.gs_m = g_stride[2];
  a.xs_b = x_stride[0], a.xs_k = x_stride[1], a.xs_n = x_stride[2];
  const long ntiles = (L + kProjTL - 1) / kProjTL;
  long want = (4L * 148 + (long)batch * K - 1) / ((long)batch * K);  // ~4 CTAs per SM over the whole grid
  if (want > ntiles) want = ntiles;
  if (want < 1) want = 1;
  a.tiles_per_chunk = (ntiles + want - 1) / want;
  a.chunks = (int)((ntiles + a.tiles_per_chunk - 1) / a.tiles_per_chunk);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (g_dtype == NZ_BF16 && x_dtype == NZ_BF16 && tc_eligible(a) && !getenv("NZ_PROJ_NO_TC")) {
    const int rc = launch_wgrad_tc(a, st);
    if (rc != NZ_EUNSUPPORTED) return rc;
  }
#define NZ_DISPATCH_X(TG)                                                    \
  switch (x_dtype) {                                                         \
    case NZ_F32: return launch_wgrad<TG, float>(a, st);                      \
    case NZ_BF16: return launch_wgrad<TG, __nv_bfloat16>(a, st);             \
    case NZ_F16: return launch_wgrad<TG, __half>(a, st);                     \
    default: break;                                                          \
  }
  switch (g_dtype) {
    case NZ_F32: NZ_DISPATCH_X(float) break;
    case NZ_BF16: NZ_DISPATCH_X(__nv_bfloat16) break;
    case NZ_F16: NZ_DISPATCH_X(__half) break;
    default: break;
  }
#undef NZ_DISPATCH_X
  set_error("nz_proj_wgrad: unsupported dtype (%d, %d)", g_dtype, x_dtype);
  return NZ_EINVAL;
}
