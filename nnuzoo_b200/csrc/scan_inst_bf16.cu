// selective-scan kernels, element type __nv_bfloat16
#include "scan_inst.cuh"
namespace nz {
NZ_INSTANTIATE_SCAN(__nv_bfloat16)
}
