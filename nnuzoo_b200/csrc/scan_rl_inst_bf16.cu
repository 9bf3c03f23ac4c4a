// row-per-lane selective-scan backward, element type __nv_bfloat16
#include "scan_rl_inst.cuh"
namespace nz {
NZ_INSTANTIATE_SCAN_RL(__nv_bfloat16)
}
