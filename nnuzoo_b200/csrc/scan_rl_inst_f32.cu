// row-per-lane selective-scan backward, element type float
#include "scan_rl_inst.cuh"
namespace nz {
NZ_INSTANTIATE_SCAN_RL(float)
}
