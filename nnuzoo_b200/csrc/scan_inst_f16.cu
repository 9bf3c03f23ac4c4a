// selective-scan kernels, element type __half
#include "scan_inst.cuh"
namespace nz {
NZ_INSTANTIATE_SCAN(__half)
}
