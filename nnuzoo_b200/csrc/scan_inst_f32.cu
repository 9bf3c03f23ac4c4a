// selective-scan kernels, element type float
#include "scan_inst.cuh"
namespace nz {
NZ_INSTANTIATE_SCAN(float)
}
