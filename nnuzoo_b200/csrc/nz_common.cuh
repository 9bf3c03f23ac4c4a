// nz_common.cuh -- device-side helpers shared by the sm_100a kernels of nnuzoo_b200.
//
// Everything here is written for Blackwell (sm_100a) only: TMA (cp.async.bulk[.tensor]) with
// mbarrier completion, 128-byte-swizzled shared-memory tiles, ex2.approx, packed fp32x2 math.
#pragma once

#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace nz {

constexpr float kLog2e = 1.4426950408889634f;
constexpr float kLn2 = 0.6931471805599453f;
constexpr int kMaxState = 16;  // d_state <= 16 (nnUZoo uses 16 everywhere: m2net.py:43)

// ----------------------------------------------------------------------------------------------
// 128-byte swizzle (CU_TENSOR_MAP_SWIZZLE_128B): byte-offset bits [4,7) ^= bits [7,10).
// `off` is relative to a 1024-byte-aligned tile base.  A thread that owns a 16..64-byte segment
// of a dense row reads it with LDS.128 conflict-free under this mapping (DESIGN.md "smem layout").
// ----------------------------------------------------------------------------------------------
__host__ __device__ __forceinline__ uint32_t swz128(uint32_t off) { return off ^ (((off >> 7) & 7u) << 4); }

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

// ------------------------------------- mbarrier ------------------------------------------------
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async_smem() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
               : "memory");
}
// plain arrive (release.cta): used for the CTA-internal producer/consumer hand-offs
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// suspend-time hint of try_wait: a waiting warp sleeps in hardware instead of spinning through the
// issue slots its CTA-mates need (the slab hand-off of the backward waited ~12 polls per state)
#ifndef NZ_MBAR_SUSPEND_NS
#define NZ_MBAR_SUSPEND_NS 2000
#endif
constexpr uint32_t kMbarSuspendNs = NZ_MBAR_SUSPEND_NS;
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity), "r"(kMbarSuspendNs)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  while (!mbar_try_wait(bar, parity)) {
  }
}

// "Last warp out issues the refill": every warp calls this once it no longer needs a shared-memory
// buffer; it returns true (warp-uniformly) in the warp that arrived last, which then re-arms the
// counter and may overwrite the buffer (e.g. issue the next TMA load into it).  No thread waits.
__device__ __forceinline__ bool warp_last_arrival(unsigned* cnt, unsigned nwarps, int lane) {
  __syncwarp();
  unsigned old = 0;
  if (lane == 0) {
    __threadfence_block();  // this warp's reads of the buffer are performed before the arrival
    old = atomicAdd(cnt, 1u);
    if (old == nwarps - 1) {
      atomicExch(cnt, 0u);
      __threadfence_block();
    }
  }
  old = __shfl_sync(0xffffffffu, old, 0);
  return old == nwarps - 1;
}

// ------------------------------------- TMA ------------------------------------------------------
// Tiled tensor loads, global -> shared, completion counted in bytes on an mbarrier.
__device__ __forceinline__ void tma_load_4d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1,
                                            int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1),
      "r"(c2), "r"(c3)
      : "memory");
}
__device__ __forceinline__ void tma_load_5d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1,
                                            int c2, int c3, int c4) {
  asm volatile(
      "cp.async.bulk.tensor.5d.shared::cluster.global.tile.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4, %5, %6, %7}], [%2];"
      ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1),
      "r"(c2), "r"(c3), "r"(c4)
      : "memory");
}
// Plain (non-tensor) bulk copy of `bytes` (multiple of 16, both addresses 16-byte aligned).
__device__ __forceinline__ void bulk_load_1d(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
      ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(src)), "r"(bytes), "r"(smem_u32(bar))
      : "memory");
}
__device__ __forceinline__ void prefetch_tmap(const CUtensorMap* map) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(map)) : "memory");
}

// ------------------------------------- math -----------------------------------------------------
__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float lg2_approx(float x) {
  float y;
  asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
// F.softplus with threshold 20 (selective_scan_interface.py:107), branch-free:
//   t = e^x;  log1p(t) = t - t^2/2 + t^3/3 for t < 0.01 (|err| < t^4/4), else ln2 * lg2(1 + t).
// Relative error ~2e-5 at worst (lg2.approx near 1), far inside the 1e-3 budget.
__device__ __forceinline__ float softplus_f(float x) {
  const float t = ex2_approx(fminf(x, 20.f) * kLog2e);
  const float series = t * fmaf(t, fmaf(t, 0.33333334f, -0.5f), 1.f);
  const float viaLog = kLn2 * lg2_approx(1.f + t);
  const float sp = t < 0.01f ? series : viaLog;
  return x > 20.f ? x : sp;
}
__device__ __forceinline__ float sigmoid_f(float x) { return __fdividef(1.f, 1.f + ex2_approx(-x * kLog2e)); }
// sigmoid(x) expressed through dl = softplus(x):  1 - e^{-dl}  (series below 0.02 to avoid cancellation)
__device__ __forceinline__ float sigmoid_from_softplus(float dl) {
  const float e = ex2_approx(-dl * kLog2e);
  const float series = dl * fmaf(dl, fmaf(dl, 0.16666667f, -0.5f), 1.f);
  return dl < 0.02f ? series : 1.f - e;
}

// ---- Kogge-Stone steps on (P, H) pairs of the affine recurrence h <- P*h + H ----------------------
// The shuffle's own predicate output ("source lane in range") guards the combine, so no lane-index
// compares or selects are needed.  kClampUp/Down encode the sub-warp width LPR for shfl.sync.
template <int LPR>
struct ShflClamp {
  static constexpr unsigned kUp = (unsigned)(32 - LPR) << 8;
  static constexpr unsigned kDown = ((unsigned)(32 - LPR) << 8) | 0x1fu;
};
// inclusive scan towards higher lanes: (P,H)[s] <- (P,H)[s] o (P,H)[s-off]
template <int LPR>
__device__ __forceinline__ void ks_up(float& P, float& H, int off) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t.reg .f32 pp, hp;\n\t"
      "shfl.sync.up.b32 pp|p, %0, %2, %3, 0xffffffff;\n\t"
      "shfl.sync.up.b32 hp, %1, %2, %3, 0xffffffff;\n\t"
      "@p fma.rn.f32 %1, %0, hp, %1;\n\t"
      "@p mul.f32 %0, %0, pp;\n\t}"
      : "+f"(P), "+f"(H)
      : "r"(off), "n"(ShflClamp<LPR>::kUp));
}
// inclusive scan towards lower lanes (suffix): (Q,G)[s] <- (Q,G)[s] o (Q,G)[s+off]
template <int LPR>
__device__ __forceinline__ void ks_down(float& Q, float& G, int off) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t.reg .f32 qq, gg;\n\t"
      "shfl.sync.down.b32 qq|p, %0, %2, %3, 0xffffffff;\n\t"
      "shfl.sync.down.b32 gg, %1, %2, %3, 0xffffffff;\n\t"
      "@p fma.rn.f32 %1, %0, gg, %1;\n\t"
      "@p mul.f32 %0, %0, qq;\n\t}"
      : "+f"(Q), "+f"(G)
      : "r"(off), "n"(ShflClamp<LPR>::kDown));
}
// value entering this lane's segment: carry (first lane) or P[s-1]*carry + H[s-1]
template <int LPR>
__device__ __forceinline__ float ks_enter_up(float P, float H, float carry) {
  float h;
  asm volatile(
      "{\n\t.reg .pred p;\n\t.reg .f32 pp, hp;\n\t"
      "shfl.sync.up.b32 pp|p, %1, 1, %4, 0xffffffff;\n\t"
      "shfl.sync.up.b32 hp, %2, 1, %4, 0xffffffff;\n\t"
      "mov.f32 %0, %3;\n\t"
      "@p fma.rn.f32 %0, pp, %3, hp;\n\t}"
      : "=f"(h)
      : "f"(P), "f"(H), "f"(carry), "n"(ShflClamp<LPR>::kUp));
  return h;
}
template <int LPR>
__device__ __forceinline__ float ks_enter_down(float Q, float G, float carry) {
  float h;
  asm volatile(
      "{\n\t.reg .pred p;\n\t.reg .f32 qq, gg;\n\t"
      "shfl.sync.down.b32 qq|p, %1, 1, %4, 0xffffffff;\n\t"
      "shfl.sync.down.b32 gg, %2, 1, %4, 0xffffffff;\n\t"
      "mov.f32 %0, %3;\n\t"
      "@p fma.rn.f32 %0, qq, %3, gg;\n\t}"
      : "=f"(h)
      : "f"(Q), "f"(G), "f"(carry), "n"(ShflClamp<LPR>::kDown));
  return h;
}

// ------------------------------------- element types --------------------------------------------
template <typename T>
struct Elem;
template <>
struct Elem<float> {
  static __device__ __forceinline__ float to_f(float v) { return v; }
  static __device__ __forceinline__ float from_f(float v) { return v; }
};
template <>
struct Elem<__nv_bfloat16> {
  static __device__ __forceinline__ float to_f(__nv_bfloat16 v) { return __bfloat162float(v); }
  static __device__ __forceinline__ __nv_bfloat16 from_f(float v) { return __float2bfloat16_rn(v); }
};
template <>
struct Elem<__half> {
  static __device__ __forceinline__ float to_f(__half v) { return __half2float(v); }
  static __device__ __forceinline__ __half from_f(float v) { return __float2half_rn(v); }
};

// Unpack one 16-byte vector of T into floats (4 for fp32, 8 for 16-bit types).
template <typename T>
__device__ __forceinline__ void unpack16(const uint4& q, float* v);
template <>
__device__ __forceinline__ void unpack16<float>(const uint4& q, float* v) {
  v[0] = __uint_as_float(q.x);
  v[1] = __uint_as_float(q.y);
  v[2] = __uint_as_float(q.z);
  v[3] = __uint_as_float(q.w);
}
template <>
__device__ __forceinline__ void unpack16<__nv_bfloat16>(const uint4& q, float* v) {
  const uint32_t w[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    v[2 * k] = __uint_as_float(w[k] << 16);
    v[2 * k + 1] = __uint_as_float(w[k] & 0xffff0000u);
  }
}
template <>
__device__ __forceinline__ void unpack16<__half>(const uint4& q, float* v) {
  const uint32_t w[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const __half2 h = *reinterpret_cast<const __half2*>(&w[k]);
    const float2 f = __half22float2(h);
    v[2 * k] = f.x;
    v[2 * k + 1] = f.y;
  }
}
template <typename T>
__device__ __forceinline__ uint4 pack16(const float* v);
template <>
__device__ __forceinline__ uint4 pack16<float>(const float* v) {
  return make_uint4(__float_as_uint(v[0]), __float_as_uint(v[1]), __float_as_uint(v[2]), __float_as_uint(v[3]));
}
template <>
__device__ __forceinline__ uint4 pack16<__nv_bfloat16>(const float* v) {
  uint32_t w[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const __nv_bfloat162 h = __floats2bfloat162_rn(v[2 * k], v[2 * k + 1]);
    w[k] = *reinterpret_cast<const uint32_t*>(&h);
  }
  return make_uint4(w[0], w[1], w[2], w[3]);
}
template <>
__device__ __forceinline__ uint4 pack16<__half>(const float* v) {
  uint32_t w[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const __half2 h = __floats2half2_rn(v[2 * k], v[2 * k + 1]);
    w[k] = *reinterpret_cast<const uint32_t*>(&h);
  }
  return make_uint4(w[0], w[1], w[2], w[3]);
}

// Read the M consecutive items a thread owns out of a swizzled dense-row tile.
//   tile      : 1024-byte aligned tile base (shared)
//   row_off   : byte offset of the row inside the tile (row * TL * sizeof(T))
//   seg_off   : byte offset of the thread's segment inside the row (segment * M * sizeof(T))
template <typename T, int M>
__device__ __forceinline__ void lds_items(const uint8_t* tile, uint32_t row_off, uint32_t seg_off, float (&v)[M]) {
  constexpr int kPer = 16 / sizeof(T);
  constexpr int kVec = M / kPer;
  static_assert(M % kPer == 0, "a thread's segment must be a whole number of 16-byte vectors");
#pragma unroll
  for (int j = 0; j < kVec; ++j) {
    const uint4 q = *reinterpret_cast<const uint4*>(tile + swz128(row_off + seg_off + 16u * j));
    unpack16<T>(q, &v[j * kPer]);
  }
}

// Store M consecutive items to global memory (blocked arrangement), masking the sequence tail.
template <typename T, int M>
__device__ __forceinline__ void stg_items(T* dst, const float (&v)[M], long t0, long L, bool vec_ok) {
  constexpr int kPer = 16 / sizeof(T);
  if (vec_ok && t0 + M <= L) {
#pragma unroll
    for (int j = 0; j < M / kPer; ++j) {
      *reinterpret_cast<uint4*>(dst + t0 + j * kPer) = pack16<T>(&v[j * kPer]);
    }
  } else {
#pragma unroll
    for (int i = 0; i < M; ++i) {
      if (t0 + i < L) dst[t0 + i] = Elem<T>::from_f(v[i]);
    }
  }
}

}  // namespace nz
