// row-per-lane selective-scan backward, element type __half
#include "scan_rl_inst.cuh"
namespace nz {
NZ_INSTANTIATE_SCAN_RL(__half)
}
