// scan_rl_inst.cuh -- host-side launch wrapper of the row-per-lane backward; included by one .cu per element type.
#pragma once

#include "scan_rl_kernels.cuh"

namespace nz {

// Row-per-lane backward: aggregate pass + combine (only when the launch is split into chunks along L) + main pass.
// One warp per CTA; shared memory, not registers, bounds residency, so the carve-out is set to the maximum once.
template <typename F>
static cudaError_t rl_max_carveout(F f) {
  return cudaFuncSetAttribute(f, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
}

template <typename T, bool kHasZ, bool kRevCap>
static cudaError_t launch_rl_one(const RlArgs& a, cudaStream_t st) {
  auto agg = scan_rl_agg_kernel<T, kHasZ, false, kRevCap>;
  auto maink1 = scan_bwd_rl_kernel<T, kHasZ, true>;   // one warp per group: plain dB / dC stores
  auto maink0 = scan_bwd_rl_kernel<T, kHasZ, false>;  // several warps per group: RED
  constexpr size_t agg_smem = 1024 + 2 * ((kHasZ ? 3 : 2) * RlCfg<T>::ROWT + RlCfg<T>::BCT) + 64;
  constexpr size_t main_smem = RlMainSmem<T, kHasZ>::bytes();
  auto main2k1 = scan_bwd_rl2_kernel<T, kHasZ, true, kRevCap>;
  auto main2k0 = scan_bwd_rl2_kernel<T, kHasZ, false, kRevCap>;
  constexpr size_t main2_smem = RlMain2Smem<T, kHasZ>::bytes();
  static bool configured = false;
  if (!configured) {
    cudaError_t e = rl_max_carveout(agg);
    if (e == cudaSuccess) e = rl_max_carveout(maink0);
    if (e == cudaSuccess) e = rl_max_carveout(maink1);
    if (e == cudaSuccess) e = rl_max_carveout(main2k0);
    if (e == cudaSuccess) e = rl_max_carveout(main2k1);
    if (e != cudaSuccess) return e;
    configured = true;
  }
  const long rbt = (long)a.batch * a.ngroups * a.nrb;
  if (a.nchunks > 1) {
    agg<<<(unsigned)(rbt * (a.nchunks - 1)), 32, agg_smem, st>>>(a);
    const long nrows = (long)a.batch * a.dim;
    scan_rl_combine_kernel<<<(unsigned)((nrows * kMaxState + 127) / 128), 128, 0, st>>>(
        a.aggG, a.aggQ, a.Rin, nrows, a.nchunks, 0, a.dim, a.dpg, kRevCap ? a.rev_mask : 0);
  }
  if (a.v2 || kRevCap) {
    if (a.single)
      main2k1<<<(unsigned)(rbt * a.nchunks), 32, main2_smem, st>>>(a);
    else
      main2k0<<<(unsigned)(rbt * a.nchunks), 32, main2_smem, st>>>(a);
  } else if (a.single)
    maink1<<<(unsigned)(rbt * a.nchunks), 32, main_smem, st>>>(a);
  else
    maink0<<<(unsigned)(rbt * a.nchunks), 32, main_smem, st>>>(a);
  return cudaGetLastError();
}

// Row-per-lane forward: aggregate pass + combine (chunked launches only) + main pass.
template <typename T, bool kHasZ, bool kRevCap>
static cudaError_t launch_rl_fwd_one(const RlArgs& a, cudaStream_t st) {
  auto agg = scan_rl_agg_kernel<T, false, true, kRevCap>;
  auto maink = scan_fwd_rl_kernel<T, kHasZ, kRevCap>;
  auto maink2 = scan_fwd_rl2_kernel<T, kHasZ>;
  constexpr size_t agg_smem = 1024 + 2 * (2 * RlCfg<T>::ROWT + RlCfg<T>::BCT) + 64;
  constexpr size_t main_smem = RlFwdSmem<T, kHasZ>::bytes();
  static bool configured = false;
  if (!configured) {
    cudaError_t e = rl_max_carveout(agg);
    if (e == cudaSuccess) e = rl_max_carveout(maink);
    if (e == cudaSuccess) e = rl_max_carveout(maink2);
    if (e != cudaSuccess) return e;
    configured = true;
  }
  const long rbt = (long)a.batch * a.ngroups * a.nrb;
  if (a.nchunks > 1) {
    agg<<<(unsigned)(rbt * (a.nchunks - 1)), 32, agg_smem, st>>>(a);
    const long nrows = (long)a.batch * a.dim;
    scan_rl_combine_kernel<<<(unsigned)((nrows * kMaxState + 127) / 128), 128, 0, st>>>(
        a.aggG, a.aggQ, a.Rin, nrows, a.nchunks, 1, a.dim, a.dpg, kRevCap ? a.rev_mask : 0);
  }
  if (a.v2f && !kRevCap)
    maink2<<<(unsigned)(rbt * a.nchunks), 32, main_smem, st>>>(a);
  else
    maink<<<(unsigned)(rbt * a.nchunks), 32, main_smem, st>>>(a);
  return cudaGetLastError();
}

// the kRevCap instantiations serve the folded SS2D path (reversed groups, shared u rows); everything else takes the plain ones
template <typename T>
static cudaError_t dispatch_rl_bwd(const RlArgs& a, bool has_z, cudaStream_t st) {
  const bool rc = a.rev_mask != 0 || a.u_gdiv > 1;
  if (rc) return has_z ? launch_rl_one<T, true, true>(a, st) : launch_rl_one<T, false, true>(a, st);
  return has_z ? launch_rl_one<T, true, false>(a, st) : launch_rl_one<T, false, false>(a, st);
}
template <typename T>
static cudaError_t dispatch_rl_fwd(const RlArgs& a, bool has_z, cudaStream_t st) {
  const bool rc = a.rev_mask != 0 || a.u_gdiv > 1;
  if (rc) return has_z ? launch_rl_fwd_one<T, true, true>(a, st) : launch_rl_fwd_one<T, false, true>(a, st);
  return has_z ? launch_rl_fwd_one<T, true, false>(a, st) : launch_rl_fwd_one<T, false, false>(a, st);
}

#define NZ_INSTANTIATE_SCAN_RL(T)                                                            \
  template <>                                                                                \
  cudaError_t launch_scan_bwd_rl<T>(const RlArgs& a, bool has_z, cudaStream_t stream) {       \
    return dispatch_rl_bwd<T>(a, has_z, stream);                                             \
  }                                                                                          \
  template <>                                                                                \
  cudaError_t launch_scan_fwd_rl<T>(const RlArgs& a, bool has_z, cudaStream_t stream) {       \
    return dispatch_rl_fwd<T>(a, has_z, stream);                                             \
  }

}  // namespace nz
