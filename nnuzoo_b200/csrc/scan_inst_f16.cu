// selective-scan kernels instantiated for __half I/O (fp32 parameters and accumulation)
#include "scan_inst.cuh"

namespace nz {
#ifdef NZ_F32_ONLY  // tuning builds (tools/tune_build.py): only the fp32 instantiation is compiled
template <>
cudaError_t launch_scan_fwd<__half>(const ScanKArgs&, bool, bool, cudaStream_t) { return cudaErrorNotSupported; }
template <>
cudaError_t launch_scan_bwd<__half>(const ScanKArgs&, bool, bool, cudaStream_t) { return cudaErrorNotSupported; }
#else
NZ_INSTANTIATE_SCAN(__half)
#endif
}  // namespace nz
