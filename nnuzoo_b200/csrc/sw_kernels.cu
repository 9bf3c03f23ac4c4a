// sw_kernels.cu -- the sliding-window predictor's per-tile arithmetic as one kernel: mirror average, gaussian weighting,
// multiply-accumulate into the logit and weight accumulators.
//
// Reference statements (nnunetv2/inference/predict_from_raw_data.py):
//   :549-565  prediction = net(x); for each mirror combination c: prediction += flip(net(flip(x, c)), c);
//             prediction /= (len(combinations) + 1)
//   :617-623  prediction = prediction.to(results_device); if use_gaussian: prediction *= gaussian
//             predicted_logits[sl] += prediction;  n_predictions[sl[1:]] += gaussian
// In PyTorch that is ~3 + 2 * mirrors element-wise launches per tile, each a full pass over the (heads, *tile) block.
// Here one thread owns one (tile, head, element): it reads the mirrored passes at their flipped positions, and performs
// the SAME sequence of roundings as the eager expressions (after every add, after the division, after the conversion
// to the accumulator dtype, after the gaussian multiply, after each accumulate), so the accumulators are bit-identical
// to the PyTorch path.  Byte work, HBM-bound: reads (mirrors + 1) x heads x tile, read-modify-writes heads x tile.
//
// Tiles of one launch must not overlap in the volume (different threads own different tiles' elements); the host
// launches overlapping tiles one by one (nnuzoo_b200/predict.py).
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/nnuzoo_b200.h"
#include "nz_common.cuh"

namespace nz {
void count_launch(int n);

struct SwArgs {
  const void* pred;      // (nmirror * ntiles, heads, t0, t1, t2) contiguous; pass m of tile t at index m * ntiles + t
  const void* gaussian;  // (t0, t1, t2) in the accumulator dtype
  void* logits;          // (heads, v0, v1, v2)
  void* npred;           // (v0, v1, v2)
  long tile[3], vol[3];
  long off[NZ_SW_MAX_TILES][3];
  int masks[NZ_SW_MAX_MIRRORS];  // bit a: pass m was computed on the tile flipped along axis a
  int nmirror, ntiles, heads;
  float inv_passes;
};

template <typename TP, typename TR>
__global__ void __launch_bounds__(256) sw_accumulate_kernel(const SwArgs a) {
  const long te = a.tile[0] * a.tile[1] * a.tile[2];
  const long total = (long)a.ntiles * a.heads * te;
  const TP* pred = reinterpret_cast<const TP*>(a.pred);
  const TR* gauss = reinterpret_cast<const TR*>(a.gaussian);
  TR* logits = reinterpret_cast<TR*>(a.logits);
  TR* npred = reinterpret_cast<TR*>(a.npred);
  const long ve = a.vol[0] * a.vol[1] * a.vol[2];
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    const long e = i % te;
    const int hd = (int)((i / te) % a.heads);
    const int t = (int)(i / (te * a.heads));
    const long e2 = e % a.tile[2], e1 = (e / a.tile[2]) % a.tile[1], e0 = e / (a.tile[2] * a.tile[1]);
    // mirror average, rounded to the prediction dtype after every step (:549-565)
    float acc = Elem<TP>::to_f(pred[((long)t * a.heads + hd) * te + e]);
    for (int m = 1; m < a.nmirror; ++m) {
      const int mk = a.masks[m];
      const long f0 = (mk & 1) ? a.tile[0] - 1 - e0 : e0;
      const long f1 = (mk & 2) ? a.tile[1] - 1 - e1 : e1;
      const long f2 = (mk & 4) ? a.tile[2] - 1 - e2 : e2;
      const float v = Elem<TP>::to_f(pred[(((long)m * a.ntiles + t) * a.heads + hd) * te + (f0 * a.tile[1] + f1) * a.tile[2] + f2]);
      acc = Elem<TP>::to_f(Elem<TP>::from_f(__fadd_rn(acc, v)));  // (_rn intrinsics: never contracted into an FMA)
    }
    if (a.nmirror > 1) acc = Elem<TP>::to_f(Elem<TP>::from_f(__fmul_rn(acc, a.inv_passes)));  // 1 / 2^k: exact reciprocal
    // to the accumulator dtype, gaussian weight, accumulate (:617-623)
    const float g = Elem<TR>::to_f(gauss[e]);
    float p = Elem<TR>::to_f(Elem<TR>::from_f(acc));
    p = Elem<TR>::to_f(Elem<TR>::from_f(__fmul_rn(p, g)));
    const long v = ((a.off[t][0] + e0) * a.vol[1] + a.off[t][1] + e1) * a.vol[2] + a.off[t][2] + e2;
    TR* lp = logits + (long)hd * ve + v;
    *lp = Elem<TR>::from_f(__fadd_rn(Elem<TR>::to_f(*lp), p));
    if (hd == 0) npred[v] = Elem<TR>::from_f(__fadd_rn(Elem<TR>::to_f(npred[v]), g));
  }
}

template <typename TP>
static void sw_launch_r(const SwArgs& a, int res_dtype, int grid, cudaStream_t st) {
  if (res_dtype == NZ_F32) sw_accumulate_kernel<TP, float><<<grid, 256, 0, st>>>(a);
  else if (res_dtype == NZ_BF16) sw_accumulate_kernel<TP, __nv_bfloat16><<<grid, 256, 0, st>>>(a);
  else sw_accumulate_kernel<TP, __half><<<grid, 256, 0, st>>>(a);
}

}  // namespace nz

extern "C" {

int nz_sw_accumulate(const void* pred, int32_t pred_dtype, int32_t nmirror, const int32_t* mirror_masks, int32_t ntiles,
                     int32_t heads, const int64_t* tile, const void* gaussian, int32_t res_dtype, void* logits,
                     void* n_pred, const int64_t* vol, const int64_t* offsets, void* stream) {
  if (!pred || !gaussian || !logits || !n_pred || !tile || !vol || !offsets) return NZ_EINVAL;
  if (nmirror < 1 || nmirror > NZ_SW_MAX_MIRRORS || ntiles < 1 || ntiles > NZ_SW_MAX_TILES || heads < 1) return NZ_EINVAL;
  if (nmirror & (nmirror - 1)) return NZ_EINVAL;  // 2^k passes (all subsets of the mirror axes): the division is exact
  if (nmirror > 1 && !mirror_masks) return NZ_EINVAL;
  auto okdt = [](int d) { return d == NZ_F32 || d == NZ_BF16 || d == NZ_F16; };
  if (!okdt(pred_dtype) || !okdt(res_dtype)) return NZ_EINVAL;
  nz::SwArgs a;
  a.pred = pred; a.gaussian = gaussian; a.logits = logits; a.npred = n_pred;
  for (int k = 0; k < 3; ++k) {
    a.tile[k] = tile[k]; a.vol[k] = vol[k];
    if (tile[k] < 1 || vol[k] < tile[k]) return NZ_EINVAL;
  }
  for (int t = 0; t < ntiles; ++t)
    for (int k = 0; k < 3; ++k) {
      a.off[t][k] = offsets[3 * t + k];
      if (a.off[t][k] < 0 || a.off[t][k] + tile[k] > vol[k]) return NZ_EINVAL;
    }
  for (int m = 0; m < nmirror; ++m) a.masks[m] = m == 0 ? 0 : mirror_masks[m];
  a.nmirror = nmirror; a.ntiles = ntiles; a.heads = heads;
  a.inv_passes = 1.0f / (float)nmirror;
  const long total = (long)ntiles * heads * tile[0] * tile[1] * tile[2];
  long blocks = (total + 255) / 256;
  const long cap = 148L * 16;
  const int grid = (int)(blocks < cap ? blocks : cap);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (pred_dtype == NZ_F32) nz::sw_launch_r<float>(a, res_dtype, grid, st);
  else if (pred_dtype == NZ_BF16) nz::sw_launch_r<__nv_bfloat16>(a, res_dtype, grid, st);
  else nz::sw_launch_r<__half>(a, res_dtype, grid, st);
  nz::count_launch(1);
  return cudaGetLastError() == cudaSuccess ? NZ_OK : NZ_ECUDA;
}

}  // extern "C"
