// scan_inst.cuh -- host-side launch wrappers; included by one .cu per element type so the
// (heavy) kernel instantiations compile in parallel.
#pragma once

#include "scan_kernels.cuh"

namespace nz {

constexpr int kM = 8;     // time steps per lane
constexpr int kLPR = 32;  // lanes per row  (kM * kLPR == NZ_CHUNK)
// forward fast path (TMA, 16 rows per CTA): 16 steps per lane, two rows per warp -> the two half-warps
// read the same B/C words (broadcast), 4 scan rounds instead of 5
#ifndef NZ_FWD_M16
#define NZ_FWD_M16 0  // measured: 2.81 vs 2.86 clk/elt/SM when rows are plentiful, 5.4 vs 4.1 when they are not (96 CTAs)
#endif

template <typename T, int WARPS, bool kTMA, bool kHasZ, int M = kM, int LPR = kLPR, int NQ = 2>
static cudaError_t launch_fwd_one(const ScanKArgs& a, cudaStream_t st) {
  using Cfg = ScanCfg<T, M, LPR, WARPS, kHasZ, false>;
  auto kern = scan_fwd_kernel<T, M, LPR, WARPS, NQ, kTMA, kHasZ>;
  const size_t smem = Cfg::smem_bytes(kTMA);
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  const unsigned grid = (unsigned)a.batch * a.ngroups * (a.dpg / Cfg::R);
  kern<<<grid, WARPS * 32, smem, st>>>(a);
  return cudaGetLastError();
}

template <typename T, int WARPS, bool kTMA, bool kHasZ>
static cudaError_t launch_bwd_one(const ScanKArgs& a, cudaStream_t st) {
  using Cfg = ScanCfg<T, kM, kLPR, WARPS, kHasZ, true>;
  auto kern = scan_bwd_kernel<T, kM, kLPR, WARPS, kTMA, kHasZ>;
  const size_t smem = Cfg::smem_bytes(kTMA);
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  const unsigned grid = (unsigned)a.batch * a.ngroups * (a.dpg / Cfg::R);
  kern<<<grid, WARPS * 32, smem, st>>>(a);
  return cudaGetLastError();
}

#define NZ_DISPATCH16(T)                                                                     \
  if (rows_per_cta == 16 && tma)                                                             \
    return has_z ? launch_fwd_one<T, 8, true, true, 16, 16, 1>(a, stream)                    \
                 : launch_fwd_one<T, 8, true, false, 16, 16, 1>(a, stream);

#define NZ_DISPATCH(FN, T)                                                                   \
  if (rows_per_cta == 8) {                                                                   \
    if (tma) return has_z ? FN<T, 8, true, true>(a, stream) : FN<T, 8, true, false>(a, stream);   \
    return has_z ? FN<T, 8, false, true>(a, stream) : FN<T, 8, false, false>(a, stream);     \
  }                                                                                          \
  if (rows_per_cta == 1 && !tma)                                                             \
    return has_z ? FN<T, 1, false, true>(a, stream) : FN<T, 1, false, false>(a, stream);     \
  return cudaErrorInvalidConfiguration;

#define NZ_INSTANTIATE_SCAN(T)                                                               \
  template <>                                                                                \
  cudaError_t launch_scan_fwd<T>(const ScanKArgs& a, bool tma, bool has_z, int rows_per_cta, \
                                 cudaStream_t stream) {                                      \
    NZ_DISPATCH16(T)                                                                         \
    NZ_DISPATCH(launch_fwd_one, T)                                                           \
  }                                                                                          \
  template <>                                                                                \
  cudaError_t launch_scan_bwd<T>(const ScanKArgs& a, bool tma, bool has_z, int rows_per_cta, \
                                 cudaStream_t stream) {                                      \
    NZ_DISPATCH(launch_bwd_one, T)                                                           \
  }

}  // namespace nz
