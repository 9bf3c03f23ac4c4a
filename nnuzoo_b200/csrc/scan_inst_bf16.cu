// selective-scan kernels instantiated for __nv_bfloat16 I/O (fp32 parameters and accumulation)
#include "scan_inst.cuh"

namespace nz {
#ifdef NZ_F32_ONLY  // tuning builds (tools/tune_build.py): only the fp32 instantiation is compiled
template <>
cudaError_t launch_scan_fwd<__nv_bfloat16>(const ScanKArgs&, bool, bool, cudaStream_t) { return cudaErrorNotSupported; }
template <>
cudaError_t launch_scan_bwd<__nv_bfloat16>(const ScanKArgs&, bool, bool, cudaStream_t) { return cudaErrorNotSupported; }
#else
NZ_INSTANTIATE_SCAN(__nv_bfloat16)
#endif
}  // namespace nz
