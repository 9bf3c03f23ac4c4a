// scan_inst.cuh -- host-side launch wrappers; included by one .cu per element type so the
// (heavy) kernel instantiations compile in parallel.
#pragma once

#include "scan_kernels.cuh"

namespace nz {

// Tile shapes (time steps per lane M, lane segments per row LPR; a warp owns 32/LPR rows):
//   forward : 16 x 16 -> tiles of 16 rows x 256 steps (two checkpoints per tile)
//   backward:  8 x 16 -> tiles of 16 rows x 128 steps (one checkpoint interval)
#ifndef NZ_FWD_M
#define NZ_FWD_M 16
#endif
#ifndef NZ_FWD_LPR
#define NZ_FWD_LPR 16
#endif
#ifndef NZ_FWD_NQ
#define NZ_FWD_NQ 1
#endif
#ifndef NZ_BWD_M
#define NZ_BWD_M 8
#endif
#ifndef NZ_BWD_LPR
#define NZ_BWD_LPR 16
#endif
constexpr int kWarps = 8;
constexpr int kFwdRows = (32 / NZ_FWD_LPR) * kWarps, kFwdTL = NZ_FWD_M * NZ_FWD_LPR;
constexpr int kBwdRows = (32 / NZ_BWD_LPR) * kWarps, kBwdTL = NZ_BWD_M * NZ_BWD_LPR;

// Persistent grid: as many CTAs as fit on the device (or as there are tiles).  The per-kernel facts behind it -- the
// dynamic shared-memory opt-in and the resident CTAs per SM -- are asked of the driver once per kernel instantiation and
// device, not on every call (BASELINE configs[3] makes hundreds of ~10 us scans per step).
constexpr int kMaxDevices = 64;
template <typename K>
static cudaError_t persistent_grid(K kern, int threads, size_t smem, int ntiles, unsigned* grid, int* cache) {
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return e;
  int* slot = (dev >= 0 && dev < kMaxDevices) ? &cache[dev] : nullptr;
  int cap = slot ? *slot : 0;
  if (cap <= 0) {
    int sms = 0, per_sm = 0;
    e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    e = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if (e != cudaSuccess) return e;
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, threads, smem);
    if (e != cudaSuccess) return e;
    if (per_sm < 1) return cudaErrorInvalidConfiguration;
    cap = sms * per_sm;
    if (slot) *slot = cap;
  }
  *grid = (unsigned)(ntiles < cap ? ntiles : cap);
  return cudaSuccess;
}

template <typename T, bool kTMA, bool kHasZ, bool kFineCk = false>
static cudaError_t launch_fwd_one(const ScanKArgs& a, cudaStream_t st) {
  if constexpr (kTMA && !kFineCk) {
    if (a.xf && a.cp_mode != 1) return launch_fwd_one<T, kTMA, kHasZ, true>(a, st);
  }
  using Cfg = ScanCfg<T, NZ_FWD_M, NZ_FWD_LPR, kWarps, kHasZ, false>;
  auto kern = scan_fwd_kernel<T, NZ_FWD_M, NZ_FWD_LPR, kWarps, NZ_FWD_NQ, kTMA, kHasZ, kFineCk>;
  const size_t smem = Cfg::smem_bytes(kTMA, kFineCk);
  static int cache[kMaxDevices] = {};
  unsigned grid = 0;
  cudaError_t e = persistent_grid(kern, kWarps * 32, smem, a.ntiles, &grid, cache);
  if (e != cudaSuccess) return e;
  kern<<<grid, kWarps * 32, smem, st>>>(a);
  return cudaGetLastError();
}

template <typename T, bool kTMA, bool kHasZ>
static cudaError_t launch_bwd_one(const ScanKArgs& a, cudaStream_t st) {
  using Cfg = ScanCfg<T, NZ_BWD_M, NZ_BWD_LPR, kWarps, kHasZ, true>;
  auto kern = scan_bwd_kernel<T, NZ_BWD_M, NZ_BWD_LPR, kWarps, kTMA, kHasZ>;
  const size_t smem = Cfg::smem_bytes(kTMA);
  static int cache[kMaxDevices] = {};
  unsigned grid = 0;
  cudaError_t e = persistent_grid(kern, kWarps * 32, smem, a.ntiles, &grid, cache);
  if (e != cudaSuccess) return e;
  kern<<<grid, kWarps * 32, smem, st>>>(a);
  return cudaGetLastError();
}

#define NZ_DISPATCH(FN, T)                                                                  \
  if (tma) return has_z ? FN<T, true, true>(a, stream) : FN<T, true, false>(a, stream);      \
  return has_z ? FN<T, false, true>(a, stream) : FN<T, false, false>(a, stream);

#define NZ_INSTANTIATE_SCAN(T)                                                               \
  template <>                                                                                \
  cudaError_t launch_scan_fwd<T>(const ScanKArgs& a, bool tma, bool has_z, cudaStream_t stream) { \
    NZ_DISPATCH(launch_fwd_one, T)                                                           \
  }                                                                                          \
  template <>                                                                                \
  cudaError_t launch_scan_bwd<T>(const ScanKArgs& a, bool tma, bool has_z, cudaStream_t stream) { \
    NZ_DISPATCH(launch_bwd_one, T)                                                           \
  }

}  // namespace nz
