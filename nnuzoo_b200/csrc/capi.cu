// capi.cu -- the extern "C" boundary declared in include/nnuzoo_b200.h.
//
// Validation mirrors the constraints of the reference wrapper
// (nnunetv2/nets/seg_mamba/selective_scan_interface.py:19-36: unit innermost stride, grouped B/C),
// builds the TMA tensor maps from the caller's pointers + strides (no .contiguous() copies: the
// SS2D callers pass B/C as split views, SURVEY.md hard part 5) and launches on the caller's stream.
#include <cuda.h>
#include <cuda_runtime.h>

#include <atomic>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "../../include/nnuzoo_b200.h"
#include "scan_inst.cuh"
#include "scan_rl_kernels.cuh"

namespace nz {

static thread_local char g_err[512] = "";
static std::atomic<long long> g_launches{0};

static int fail(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}

void set_error(const char* fmt, ...) {  // for the other translation units
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

void count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

static size_t esize(int dtype) { return dtype == NZ_F32 ? 4 : 2; }

constexpr size_t kWsHeader = 256;  // ticket counter (+ padding) in front of the hand-off slots

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

// Tensor map over a (…, rows, L) tensor viewed as (128-byte line, lines per row, outer dims…):
// the smem image of a box is then dense rows of NZ_CHUNK elements, 128B-swizzled.
static bool make_map(CUtensorMap* m, int dtype, const void* base, int nouter, const int64_t* outer_dim,
                     const int64_t* outer_stride_elems, int64_t L, const int* outer_box, int tile_len) {
  EncodeTiledFn enc = get_encode();
  if (!enc) return false;
  const size_t es = esize(dtype);
  const cuuint32_t inner = (cuuint32_t)(128 / es);
  cuuint64_t dims[5];
  cuuint64_t strides[4];
  cuuint32_t box[5], estr[5] = {1, 1, 1, 1, 1};
  dims[0] = inner;
  dims[1] = (cuuint64_t)(L * es / 128);
  strides[0] = 128;
  box[0] = inner;
  box[1] = (cuuint32_t)(tile_len * es / 128);
  const cuuint64_t safe = (cuuint64_t)L * es;  // stride used for extent-1 dims (any 16B multiple works)
  for (int i = 0; i < nouter; ++i) {
    dims[2 + i] = (cuuint64_t)outer_dim[i];
    cuuint64_t sb = (cuuint64_t)outer_stride_elems[i] * es;
    if (outer_dim[i] == 1 && (sb == 0 || sb % 16 != 0)) sb = safe;
    strides[1 + i] = sb;
    box[2 + i] = (cuuint32_t)outer_box[i];
  }
  const CUtensorMapDataType dt = dtype == NZ_F32    ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32
                                 : dtype == NZ_BF16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16
                                                    : CU_TENSOR_MAP_DATA_TYPE_FLOAT16;
  CUresult r = enc(m, dt, (cuuint32_t)(2 + nouter), const_cast<void*>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS;
}

// Tensor map over a (..., rows, L) tensor viewed as (8-step block, blocks per row, outer dims...): a box is then one
// block of `outer_box` rows -- dense rows of 8 elements (fp32: 32 bytes under SWIZZLE_32B, 16-bit: 16 bytes).  With
// inner_elems = 16 and fp32 it views the fine checkpoints (..., rows, L / 8, 16) as 64-byte rows under SWIZZLE_64B.
static bool make_map_blk(CUtensorMap* m, int dtype, const void* base, int inner_elems, int64_t nblocks, int nouter,
                         const int64_t* outer_dim, const int64_t* outer_stride_elems, const int* outer_box) {
  EncodeTiledFn enc = get_encode();
  if (!enc) return false;
  const size_t es = esize(dtype);
  cuuint64_t dims[5];
  cuuint64_t strides[4];
  cuuint32_t box[5], estr[5] = {1, 1, 1, 1, 1};
  dims[0] = (cuuint64_t)inner_elems;
  dims[1] = (cuuint64_t)nblocks;
  strides[0] = (cuuint64_t)inner_elems * es;
  box[0] = (cuuint32_t)inner_elems;
  box[1] = 1;
  const cuuint64_t safe = (cuuint64_t)nblocks * inner_elems * es;
  for (int i = 0; i < nouter; ++i) {
    dims[2 + i] = (cuuint64_t)outer_dim[i];
    cuuint64_t sb = (cuuint64_t)outer_stride_elems[i] * es;
    if (outer_dim[i] == 1 && (sb == 0 || sb % 16 != 0)) sb = safe;
    strides[1 + i] = sb;
    box[2 + i] = (cuuint32_t)outer_box[i];
  }
  const CUtensorMapDataType dt = dtype == NZ_F32    ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32
                                 : dtype == NZ_BF16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16
                                                    : CU_TENSOR_MAP_DATA_TYPE_FLOAT16;
  const size_t row_bytes = (size_t)inner_elems * es;
  const CUtensorMapSwizzle sw = row_bytes == 64   ? CU_TENSOR_MAP_SWIZZLE_64B
                                : row_bytes == 32 ? CU_TENSOR_MAP_SWIZZLE_32B
                                                  : CU_TENSOR_MAP_SWIZZLE_NONE;
  CUresult r = enc(m, dt, (cuuint32_t)(2 + nouter), const_cast<void*>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS;
}

static bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

// TMA can express a (batch, rows, L) operand iff every stride is a 16-byte multiple, the base is
// 16-byte aligned and a row is a whole number of 128-byte lines.
static bool tma_ok_rows(const void* p, int dtype, int64_t L, const int64_t* dims, const int64_t* strides, int n) {
  const size_t es = esize(dtype);
  if (!p || !aligned16(p) || (L * es) % 128 != 0) return false;
  for (int i = 0; i < n; ++i)
    if (dims[i] > 1 && ((strides[i] * (int64_t)es) % 16 != 0 || strides[i] <= 0)) return false;
  return true;
}

// SM count of the current device, asked once per device
static int sm_count() {
  static int cached[64] = {};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
  if (cached[dev] <= 0) {
    int sms = 148;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    cached[dev] = sms;
  }
  return cached[dev];
}

static inline bool folded(const NzScanDesc* d) { return d->rev_mask != 0 || d->u_gdiv > 1; }
static inline int u_gdiv(const NzScanDesc* d) { return d->u_gdiv > 1 ? d->u_gdiv : 1; }

static int validate(const NzScanDesc* d, bool bwd) {
  if (!d) return fail(NZ_EINVAL, "null descriptor");
  if (d->batch < 1 || d->dim < 1 || d->dstate < 1 || d->ngroups < 1 || d->seqlen < 1)
    return fail(NZ_EINVAL, "batch/dim/dstate/ngroups/seqlen must be >= 1 (got %d %d %d %d %lld)", d->batch, d->dim,
                d->dstate, d->ngroups, (long long)d->seqlen);
  if (d->dstate > NZ_MAX_DSTATE)
    return fail(NZ_EUNSUPPORTED, "d_state %d > %d is not implemented (nnUZoo uses 16)", d->dstate, NZ_MAX_DSTATE);
  if (d->dim % d->ngroups) return fail(NZ_EINVAL, "dim %d is not a multiple of ngroups %d", d->dim, d->ngroups);
  if (d->dtype != NZ_F32 && d->dtype != NZ_BF16 && d->dtype != NZ_F16) return fail(NZ_EINVAL, "bad dtype %d", d->dtype);
  if (!d->u || !d->delta || !d->A || !d->B || !d->C) return fail(NZ_EINVAL, "u/delta/A/B/C must be non-null");
  if (!d->x) return fail(NZ_EINVAL, "checkpoint buffer x must be non-null");
  if (!d->workspace || d->workspace_bytes < nz_scan_workspace_bytes(d))
    return fail(NZ_EINVAL, "workspace must hold nz_scan_workspace_bytes() = %lld bytes (got %lld)",
                (long long)nz_scan_workspace_bytes(d), (long long)d->workspace_bytes);
  if (reinterpret_cast<uintptr_t>(d->workspace) & 255u) return fail(NZ_EINVAL, "workspace must be 256-byte aligned");
  if (!bwd && !d->out) return fail(NZ_EINVAL, "out must be non-null");
  if (bwd) {
    if (!d->dout || !d->du || !d->ddelta || !d->dA || !d->dB || !d->dC)
      return fail(NZ_EINVAL, "dout/du/ddelta/dA/dB/dC must be non-null for the backward");
    if ((d->z != nullptr) != (d->dz != nullptr)) return fail(NZ_EINVAL, "dz must be given iff z is given");
  }
  if ((long long)d->batch * d->ngroups * d->dim > 2000000000LL) return fail(NZ_EINVAL, "grid too large");
  if (d->u_gdiv > 1 && d->ngroups % d->u_gdiv)
    return fail(NZ_EINVAL, "ngroups %d is not a multiple of u_gdiv %d", d->ngroups, d->u_gdiv);
  if (d->ngroups < 32 && ((unsigned)d->rev_mask >> d->ngroups))
    return fail(NZ_EINVAL, "rev_mask 0x%x names groups beyond %d", d->rev_mask, d->ngroups);
  if (d->rev_mask && d->ngroups > 32) return fail(NZ_EUNSUPPORTED, "rev_mask needs ngroups <= 32");
  return NZ_OK;
}

static void* g_trace = nullptr;  // tools only: see nz_debug_set_trace

static void fill_args(const NzScanDesc* d, ScanKArgs& a, bool bwd) {
  memset(&a, 0, sizeof(a));
  a.trace = bwd ? nullptr : reinterpret_cast<unsigned long long*>(g_trace);
  a.u = d->u; a.delta = d->delta; a.z = d->z; a.dout = d->dout; a.B = d->B; a.C = d->C;
  a.A = d->A; a.D = d->D; a.bias = d->delta_bias;
  a.out = d->out; a.du = d->du; a.ddelta = d->ddelta; a.dz = d->dz;
  a.x = d->x; a.dA = d->dA; a.dB = d->dB; a.dC = d->dC; a.dD = d->dD; a.dbias = d->ddelta_bias;
  a.L = d->seqlen;
  a.u_bs = d->u_stride[0]; a.u_ds = d->u_stride[1];
  a.dl_bs = d->delta_stride[0]; a.dl_ds = d->delta_stride[1];
  a.z_bs = d->z_stride[0]; a.z_ds = d->z_stride[1];
  a.o_bs = d->out_stride[0]; a.o_ds = d->out_stride[1];
  a.do_bs = d->dout_stride[0]; a.do_ds = d->dout_stride[1];
  a.B_bs = d->B_stride[0]; a.B_gs = d->B_stride[1]; a.B_ns = d->B_stride[2];
  a.C_bs = d->C_stride[0]; a.C_gs = d->C_stride[1]; a.C_ns = d->C_stride[2];
  a.A_ds = d->A_stride;
  a.batch = d->batch; a.dim = d->dim; a.dstate = d->dstate; a.ngroups = d->ngroups;
  a.dpg = d->dim / d->ngroups;
  const int rows = bwd ? kBwdRows : kFwdRows, tl = bwd ? kBwdTL : kFwdTL;
  a.nrb = (a.dpg + rows - 1) / rows;
  a.nrb_total = d->batch * d->ngroups * a.nrb;
  a.nchunks = (int)((d->seqlen + tl - 1) / tl);
  a.nck = (int)nz_scan_num_chunks(d->seqlen);
  a.ntiles = a.nrb_total * a.nchunks;
  // optional start skew: a dependent tile whose predecessor is still running may wait until that one is
  // `skew` states ahead before it starts polling.  Measured (profiles/r01_kernel_tuning.md): every value
  // >= 0 loses to not waiting at all (-1), so it stays an experiment knob.
  a.skew = -1;
  // chain-limited launches (far fewer row blocks than the 2 resident CTAs per SM) claim tickets just in time
  {
    const int sms = sm_count();
    a.claim_late = a.nrb_total <= sms / 2 ? 1 : 0;  // measured: 1.3-1.7x at 8 row blocks, neutral at 96, -5 % at 192
  }
  if (const char* e = getenv("NZ_CLAIM_LATE")) a.claim_late = atoi(e);  // tuning override
  if (const char* e = getenv("NZ_SKEW")) a.skew = atoi(e);  // tuning override
  a.ticket = reinterpret_cast<unsigned*>(d->workspace);
  a.carry = reinterpret_cast<unsigned long long*>(reinterpret_cast<char*>(d->workspace) + kWsHeader);
  a.softplus = d->delta_softplus;
  const size_t es = esize(d->dtype);
  a.out_f32 = (d->out_f32 && d->dtype != NZ_F32) ? 1 : 0;
  const size_t eo = a.out_f32 ? 4 : es;
  a.vec_out = d->out && aligned16(d->out) && (d->out_stride[0] * eo) % 16 == 0 && (d->out_stride[1] * eo) % 16 == 0;
  a.xf = nullptr;  // set by run_scan when the fine checkpoints apply
  a.nbt = d->seqlen / NZ_FINE;
  a.vec_grad = (d->seqlen * es) % 16 == 0 && (d->seqlen % 4 == 0) && (!d->du || aligned16(d->du)) &&
               (!d->ddelta || aligned16(d->ddelta)) && (!d->dz || aligned16(d->dz)) && (!d->dB || aligned16(d->dB)) &&
               (!d->dC || aligned16(d->dC));
}

// Decide the path and build the tensor maps.  Returns true when the TMA path is usable.
static bool setup_tma(const NzScanDesc* d, ScanKArgs& a, bool bwd) {
  const int rows_per_cta = bwd ? kBwdRows : kFwdRows, tl = bwd ? kBwdTL : kFwdTL;
  if (d->force_generic || d->dstate != NZ_MAX_DSTATE || a.dpg % rows_per_cta != 0) return false;
  const int64_t rd[2] = {d->dim, d->batch};
  const int64_t us[2] = {d->u_stride[1], d->u_stride[0]};
  const int64_t ds[2] = {d->delta_stride[1], d->delta_stride[0]};
  const int64_t zs[2] = {d->z_stride[1], d->z_stride[0]};
  const int64_t os[2] = {d->dout_stride[1], d->dout_stride[0]};
  const int64_t bd[3] = {d->dstate, d->ngroups, d->batch};
  const int64_t bs[3] = {d->B_stride[2], d->B_stride[1], d->B_stride[0]};
  const int64_t cs[3] = {d->C_stride[2], d->C_stride[1], d->C_stride[0]};
  const int64_t L = d->seqlen;
  if (!tma_ok_rows(d->u, d->dtype, L, rd, us, 2) || !tma_ok_rows(d->delta, d->dtype, L, rd, ds, 2) ||
      !tma_ok_rows(d->B, d->dtype, L, bd, bs, 3) || !tma_ok_rows(d->C, d->dtype, L, bd, cs, 3))
    return false;
  if (d->z && !tma_ok_rows(d->z, d->dtype, L, rd, zs, 2)) return false;
  if (bwd && !tma_ok_rows(d->dout, d->dtype, L, rd, os, 2)) return false;
  const int rbox[2] = {rows_per_cta, 1};
  const int bbox[3] = {NZ_MAX_DSTATE, 1, 1};
  bool ok = make_map(&a.tm_u, d->dtype, d->u, 2, rd, us, L, rbox, tl) &&
            make_map(&a.tm_delta, d->dtype, d->delta, 2, rd, ds, L, rbox, tl) &&
            make_map(&a.tm_B, d->dtype, d->B, 3, bd, bs, L, bbox, tl) && make_map(&a.tm_C, d->dtype, d->C, 3, bd, cs, L, bbox, tl);
  if (ok && d->z) ok = make_map(&a.tm_z, d->dtype, d->z, 2, rd, zs, L, rbox, tl);
  if (ok && bwd) ok = make_map(&a.tm_dout, d->dtype, d->dout, 2, rd, os, L, rbox, tl);
  return ok;
}

// ---- chunk-parallel forward (few rows, long L) --------------------------------------------------------------
// The chained hand-off serialises the chunks of a row block; with fewer row blocks than SMs that chain is all there
// is (BASELINE configs[2]: 128 rows x 2 M steps = 8 row blocks x 8192 links, 1.9 % of the HBM peak).  When the caller
// provides the larger workspace of nz_scan_workspace_bytes_cp(), the forward runs as three launches instead:
// aggregate pass (all tiles at once, h = 0 in, (prod a, h_end) out), a combine kernel that walks the chunks of every
// (row, state) -- h_in[c+1] = fma(P_c, h_in[c], H_c), the very operation the chained kernel performs, so the results are
// bit-identical -- and the final pass with every tile on the fast path.
static bool cp_eligible(const NzScanDesc* d) {
  const int sms = sm_count();
  const int dpg = d->dim / (d->ngroups > 0 ? d->ngroups : 1);
  const long nrb_total = (long)d->batch * d->ngroups * ((dpg + kFwdRows - 1) / kFwdRows);
  const long nchunks = (d->seqlen + kFwdTL - 1) / kFwdTL;
  if (getenv("NZ_NO_CP")) return false;
  return nrb_total <= sms / 2 && nchunks >= 16;
}
static int64_t cp_extra_bytes(const NzScanDesc* d) {
  const int64_t nchunks = (d->seqlen + kFwdTL - 1) / kFwdTL;
  return 3 * (int64_t)d->batch * d->dim * NZ_MAX_DSTATE * nchunks * (int64_t)sizeof(float);
}

__global__ void __launch_bounds__(128) cp_combine_kernel(const float* __restrict__ P, const float* __restrict__ H,
                                                         float* __restrict__ hin, long nseq, int nchunks) {
  const long s = blockIdx.x * 128L + threadIdx.x;  // one (row, state) sequence of chunk aggregates per thread
  if (s >= nseq) return;
  const float* p = P + s * nchunks;
  const float* g = H + s * nchunks;
  float* o = hin + s * nchunks;
  float h = 0.f;
  int c = 0;
  // There are only rows x 16 such chains, far fewer than the machine has lanes, so the walk is bound by load latency:
  // 32 links per trip as 16-byte loads, the next trip's 16 loads already in flight while this trip's chain runs.
  if ((nchunks & 3) == 0) {
    constexpr int V = 8;  // float4 per array per trip
    float4 pa[V], ga[V], pb[V], gb[V];
    const float4* p4 = reinterpret_cast<const float4*>(p);
    const float4* g4 = reinterpret_cast<const float4*>(g);
    float4* o4 = reinterpret_cast<float4*>(o);
    const int ntrip = nchunks / (4 * V);
    if (ntrip > 0) {
#pragma unroll
      for (int i = 0; i < V; ++i) pa[i] = p4[i], ga[i] = g4[i];
    }
    for (int t = 0; t < ntrip; ++t) {
      if (t + 1 < ntrip) {
#pragma unroll
        for (int i = 0; i < V; ++i) pb[i] = p4[(t + 1) * V + i], gb[i] = g4[(t + 1) * V + i];
      }
#pragma unroll
      for (int i = 0; i < V; ++i) {
        float4 r;
        r.x = h, h = fmaf(pa[i].x, h, ga[i].x);
        r.y = h, h = fmaf(pa[i].y, h, ga[i].y);
        r.z = h, h = fmaf(pa[i].z, h, ga[i].z);
        r.w = h, h = fmaf(pa[i].w, h, ga[i].w);
        o4[t * V + i] = r;
      }
#pragma unroll
      for (int i = 0; i < V; ++i) pa[i] = pb[i], ga[i] = gb[i];
    }
    c = ntrip * 4 * V;
  }
  for (; c < nchunks; ++c) {
    o[c] = h;
    h = fmaf(p[c], h, g[c]);
  }
}

template <typename T>
static cudaError_t run_fwd_cp(ScanKArgs a, const NzScanDesc* d, bool tma, bool has_z, cudaStream_t st) {
  float* base = reinterpret_cast<float*>(reinterpret_cast<char*>(d->workspace) + nz_scan_workspace_bytes(d));
  const long nseq = (long)d->batch * d->dim * NZ_MAX_DSTATE;
  a.cp_P = base;
  a.cp_H = base + nseq * a.nchunks;
  float* hin = base + 2 * nseq * a.nchunks;
  a.cp_hin = hin;
  a.claim_late = 0;  // the tiles are independent in both passes
  a.cp_mode = 1;
  cudaError_t e = launch_scan_fwd<T>(a, tma, has_z, st);
  if (e != cudaSuccess) return e;
  cp_combine_kernel<<<(unsigned)((nseq + 127) / 128), 128, 0, st>>>(a.cp_P, a.cp_H, hin, nseq, a.nchunks);
  e = cudaMemsetAsync(d->workspace, 0, kWsHeader, st);  // the ticket counter starts over
  if (e != cudaSuccess) return e;
  a.cp_mode = 2;
  e = launch_scan_fwd<T>(a, tma, has_z, st);
  if (e == cudaSuccess) count_launch(2);
  return e;
}

// ---- row-per-lane backward (csrc/scan_rl_kernels.cuh) -------------------------------------------------------------
// Applies when a warp can own 32 whole rows of one group and every operand is TMA-expressible.
static bool rl_shape_ok(const NzScanDesc* d) {
  if (!d || d->batch < 1 || d->dim < 1 || d->ngroups < 1 || d->seqlen < 1 || d->dim % d->ngroups) return false;
  if (d->force_generic || d->dstate != NZ_MAX_DSTATE || (d->dim / d->ngroups) % 32 != 0) return false;
  if (getenv("NZ_NO_RL")) return false;
  // Measured old (warp-scan, chained) vs row-per-lane over the M2Net shapes (profiles/r02_rl_table.log, butterfly
  // backward + 32-byte stores): the row-per-lane backward wins from about 12 M elements up (12 x 256 x 4096: 0.30 vs
  // 0.34 ms, 12 x 128 x 65536: 1.82 vs 2.73 ms) and is the only chunk-parallel backward (batch-1 inference shape
  // 3.5 -> 1.1 ms); below that its three launches and the aggregate pass cost more than they save (12 x 128 x 4096:
  // 0.23 vs 0.19 ms).
  {
    long min_elts = 12L << 20;
    if (const char* e = getenv("NZ_RL_MIN_ELTS")) min_elts = atol(e);  // tests / tuning
    if (!folded(d) && (long)d->batch * d->dim * d->seqlen < min_elts) return false;  // folded calls have no other path
  }
  const int64_t ud[2] = {d->dim / u_gdiv(d), d->batch};
  const int64_t rd[2] = {d->dim, d->batch};
  const int64_t us[2] = {d->u_stride[1], d->u_stride[0]};
  const int64_t ds[2] = {d->delta_stride[1], d->delta_stride[0]};
  const int64_t zs[2] = {d->z_stride[1], d->z_stride[0]};
  const int64_t bd[3] = {d->dstate, d->ngroups, d->batch};
  const int64_t bs[3] = {d->B_stride[2], d->B_stride[1], d->B_stride[0]};
  const int64_t cs[3] = {d->C_stride[2], d->C_stride[1], d->C_stride[0]};
  const int64_t L = d->seqlen;
  if (!tma_ok_rows(d->u, d->dtype, L, ud, us, 2) || !tma_ok_rows(d->delta, d->dtype, L, rd, ds, 2) ||
      !tma_ok_rows(d->B, d->dtype, L, bd, bs, 3) || !tma_ok_rows(d->C, d->dtype, L, bd, cs, 3))
    return false;
  if (d->z && !tma_ok_rows(d->z, d->dtype, L, rd, zs, 2)) return false;
  return true;
}

// Chunks along L: (row block, chunk) work items are one warp each and 8 warps are resident per SM; all items cost the
// same, so the launch is cut into about 6 waves of them (two full waves plus 32 stragglers measured 2.14 ms where 1.43
// would do: profiles/r02_kernel_tuning.md), each chunk a whole number of 128-byte tiles.  `resident` = warps per SM of
// the main pass (backward 12, forward 16).
static void rl_plan(const NzScanDesc* d, int* nchunks, int* tpc, int resident = 12) {
  const int sms = sm_count();
  const long rbt = (long)d->batch * (d->dim / 32);
  const long ntl = d->seqlen * (long)esize(d->dtype) / 128;
  long target = 6L * sms * resident;
  if (const char* e = getenv("NZ_RL_ITEMS")) target = atol(e);  // tuning override
  long nc = (target + rbt - 1) / rbt;
  if (nc < 1) nc = 1;
  if (nc > ntl) nc = ntl;
  const long t = (ntl + nc - 1) / nc;
  *tpc = (int)t;
  *nchunks = (int)((ntl + t - 1) / t);
}

static int64_t rl_extra_bytes(const NzScanDesc* d) {
  int nc = 1, tpc = 1;
  rl_plan(d, &nc, &tpc, 16);  // the forward's plan has the most chunks
  if (nc <= 1) return 0;
  const int64_t one = (((int64_t)d->batch * d->dim * nc * NZ_MAX_DSTATE * 4) + 255) & ~(int64_t)255;
  return 3 * one;
}

static int bwd_rl_v2() {
  int v2 = 1;
  if (const char* e = getenv("NZ_RL_BWD2")) v2 = atoi(e) != 0;  // A/B against the slab version
  return v2;
}

template <typename T>
static cudaError_t run_bwd_rl(const NzScanDesc* d, cudaStream_t st) {
  RlArgs r;
  memset(&r, 0, sizeof(r));
  const int64_t rd[2] = {d->dim, d->batch};
  const int64_t us[2] = {d->u_stride[1], d->u_stride[0]};
  const int64_t ds[2] = {d->delta_stride[1], d->delta_stride[0]};
  const int64_t zs[2] = {d->z_stride[1], d->z_stride[0]};
  const int64_t os[2] = {d->dout_stride[1], d->dout_stride[0]};
  const int64_t bd[3] = {d->dstate, d->ngroups, d->batch};
  const int64_t bs[3] = {d->B_stride[2], d->B_stride[1], d->B_stride[0]};
  const int64_t cs[3] = {d->C_stride[2], d->C_stride[1], d->C_stride[0]};
  const int64_t L = d->seqlen;
  const int rbox[2] = {32, 1};
  const int bbox[3] = {NZ_MAX_DSTATE, 1, 1};
  const int tl = 128 / (int)esize(d->dtype);
  const int64_t nbt = L / NZ_FINE;
  const int64_t xs[2] = {nbt * NZ_MAX_DSTATE, (int64_t)d->dim * nbt * NZ_MAX_DSTATE};
  // main pass: per-block boxes
  const int64_t ud[2] = {d->dim / u_gdiv(d), d->batch};
  bool ok = make_map_blk(&r.tm_u, d->dtype, d->u, NZ_FINE, nbt, 2, ud, us, rbox) &&
            make_map_blk(&r.tm_delta, d->dtype, d->delta, NZ_FINE, nbt, 2, rd, ds, rbox) &&
            make_map_blk(&r.tm_dout, d->dtype, d->dout, NZ_FINE, nbt, 2, rd, os, rbox) &&
            make_map_blk(&r.tm_B, d->dtype, d->B, NZ_FINE, nbt, 3, bd, bs, bbox) &&
            make_map_blk(&r.tm_C, d->dtype, d->C, NZ_FINE, nbt, 3, bd, cs, bbox) &&
            make_map_blk(&r.tm_xf, NZ_F32, d->xf, NZ_MAX_DSTATE, nbt, 2, rd, xs, rbox);
  if (ok && d->z) ok = make_map_blk(&r.tm_z, d->dtype, d->z, NZ_FINE, nbt, 2, rd, zs, rbox);
  // aggregate pass: per-tile boxes (128 bytes of a row)
  ok = ok && make_map(&r.g_delta, d->dtype, d->delta, 2, rd, ds, L, rbox, tl) &&
       make_map(&r.g_row1, d->dtype, d->dout, 2, rd, os, L, rbox, tl) && make_map(&r.g_bc, d->dtype, d->C, 3, bd, cs, L, bbox, tl);
  if (ok && d->z) ok = make_map(&r.g_z, d->dtype, d->z, 2, rd, zs, L, rbox, tl);
  if (!ok) return cudaErrorInvalidValue;
  r.A = d->A; r.D = d->D; r.bias = d->delta_bias; r.xf = d->xf;
  r.du = d->du; r.ddelta = d->ddelta; r.dz = d->dz;
  r.dA = d->dA; r.dB = d->dB; r.dC = d->dC; r.dD = d->dD; r.dbias = d->ddelta_bias;
  r.L = L; r.A_ds = d->A_stride;
  r.batch = d->batch; r.dim = d->dim; r.ngroups = d->ngroups; r.dpg = d->dim / d->ngroups;
  r.nrb = r.dpg / 32;
  r.ntl = (int)(L * (int64_t)esize(d->dtype) / 128);
  {
    auto a32 = [](const void* p) { return (reinterpret_cast<uintptr_t>(p) & 31u) == 0; };
    r.wide = a32(d->du) && a32(d->ddelta) && (!d->dz || a32(d->dz)) && (L * (int64_t)esize(d->dtype)) % 32 == 0;
    if (getenv("NZ_RL_NOWIDE")) r.wide = 0;
  }
  r.v2 = bwd_rl_v2();
  rl_plan(d, &r.nchunks, &r.tpc, r.v2 ? NZ_RL_BWD2_MINB : 12);
  r.softplus = d->delta_softplus;
  r.rev_mask = d->rev_mask;
  r.u_gdiv = u_gdiv(d);
  r.single = r.nrb == 1;
  r.zero_dbc = (!r.single && r.nchunks > 1 && !getenv("NZ_RL_NO_ZERO_DBC")) ? 1 : 0;  // nz_scan_bwd_overwrites_dbc() says so
  if (r.nchunks > 1) {
    const int64_t one = (((int64_t)d->batch * d->dim * r.nchunks * NZ_MAX_DSTATE * 4) + 255) & ~(int64_t)255;
    char* base = reinterpret_cast<char*>(d->workspace) + nz_scan_workspace_bytes(d);
    r.aggG = reinterpret_cast<float*>(base);
    r.aggQ = reinterpret_cast<float*>(base + one);
    r.Rin = reinterpret_cast<float*>(base + 2 * one);
  }
  cudaError_t e = launch_scan_bwd_rl<T>(r, d->z != nullptr, st);
  if (e == cudaSuccess) count_launch(r.nchunks > 1 ? 3 : 1);
  return e;
}

// Row-per-lane forward: same eligibility; also writes the fine checkpoints when d->xf is set.
template <typename T>
static cudaError_t run_fwd_rl(const NzScanDesc* d, cudaStream_t st) {
  RlArgs r;
  memset(&r, 0, sizeof(r));
  const int64_t rd[2] = {d->dim, d->batch};
  const int64_t us[2] = {d->u_stride[1], d->u_stride[0]};
  const int64_t ds[2] = {d->delta_stride[1], d->delta_stride[0]};
  const int64_t zs[2] = {d->z_stride[1], d->z_stride[0]};
  const int64_t bd[3] = {d->dstate, d->ngroups, d->batch};
  const int64_t bs[3] = {d->B_stride[2], d->B_stride[1], d->B_stride[0]};
  const int64_t cs[3] = {d->C_stride[2], d->C_stride[1], d->C_stride[0]};
  const int64_t L = d->seqlen;
  const int rbox[2] = {32, 1};
  const int bbox[3] = {NZ_MAX_DSTATE, 1, 1};
  const int tl = 128 / (int)esize(d->dtype);
  const int64_t nbt = L / NZ_FINE;
  const int64_t ud[2] = {d->dim / u_gdiv(d), d->batch};
  bool ok = make_map_blk(&r.tm_u, d->dtype, d->u, NZ_FINE, nbt, 2, ud, us, rbox) &&
            make_map_blk(&r.tm_delta, d->dtype, d->delta, NZ_FINE, nbt, 2, rd, ds, rbox) &&
            make_map_blk(&r.tm_B, d->dtype, d->B, NZ_FINE, nbt, 3, bd, bs, bbox) &&
            make_map_blk(&r.tm_C, d->dtype, d->C, NZ_FINE, nbt, 3, bd, cs, bbox);
  if (ok && d->z) ok = make_map_blk(&r.tm_z, d->dtype, d->z, NZ_FINE, nbt, 2, rd, zs, rbox);
  ok = ok && make_map(&r.g_delta, d->dtype, d->delta, 2, rd, ds, L, rbox, tl) &&
       make_map(&r.g_row1, d->dtype, d->u, 2, ud, us, L, rbox, tl) && make_map(&r.g_bc, d->dtype, d->B, 3, bd, bs, L, bbox, tl);
  if (!ok) return cudaErrorInvalidValue;
  r.A = d->A; r.D = d->D; r.bias = d->delta_bias;
  r.out = d->out; r.x = d->x; r.xfw = d->xf;
  r.o_bs = d->out_stride[0]; r.o_ds = d->out_stride[1];
  r.nck = (int)nz_scan_num_chunks(L);
  r.out_f32 = (d->out_f32 && d->dtype != NZ_F32) ? 1 : 0;
  r.L = L; r.A_ds = d->A_stride;
  r.batch = d->batch; r.dim = d->dim; r.ngroups = d->ngroups; r.dpg = d->dim / d->ngroups;
  r.nrb = r.dpg / 32;
  r.ntl = (int)(L * (int64_t)esize(d->dtype) / 128);
  {
    auto a32 = [](const void* p) { return (reinterpret_cast<uintptr_t>(p) & 31u) == 0; };
    const int64_t eo = r.out_f32 ? 4 : (int64_t)esize(d->dtype);
    r.wide = a32(d->out) && (d->out_stride[0] * eo) % 32 == 0 && (d->out_stride[1] * eo) % 32 == 0 && (!d->xf || a32(d->xf));
    if (getenv("NZ_RL_NOWIDE")) r.wide = 0;
  }
  r.v2f = 0;
  if (const char* e = getenv("NZ_RL_FWD2")) r.v2f = atoi(e) != 0;  // A/B against the first version
  rl_plan(d, &r.nchunks, &r.tpc, 16);
  r.softplus = d->delta_softplus;
  r.rev_mask = d->rev_mask;
  r.u_gdiv = u_gdiv(d);
  if (r.nchunks > 1) {
    const int64_t one = (((int64_t)d->batch * d->dim * r.nchunks * NZ_MAX_DSTATE * 4) + 255) & ~(int64_t)255;
    char* base = reinterpret_cast<char*>(d->workspace) + nz_scan_workspace_bytes(d);
    r.aggG = reinterpret_cast<float*>(base);
    r.aggQ = reinterpret_cast<float*>(base + one);
    r.Rin = reinterpret_cast<float*>(base + 2 * one);
  }
  cudaError_t e = launch_scan_fwd_rl<T>(r, d->z != nullptr, st);
  if (e == cudaSuccess) count_launch(r.nchunks > 1 ? 3 : 1);
  return e;
}

static bool rl_fwd_usable(const NzScanDesc* d) {
  // Measured (profiles/r02_rl_table.log): without fine checkpoints the warp-scan forward is as fast or faster on most
  // shapes; with them (training) the row-per-lane forward wins from about 24 M elements up (12 x 128 x 262144: 4.68 vs
  // 5.29 ms, 12 x 128 x 65536: 1.07 vs 1.33, 12 x 128 x 16384: 0.35 vs 0.36).  NZ_RL_FWD=0 / 1 forces the choice.
  if (!rl_shape_ok(d)) return false;
  if (folded(d)) {
    // reversed groups / shared u rows exist only in these kernels
  } else if (const char* on = getenv("NZ_RL_FWD")) {
    if (atoi(on) == 0) return false;
  } else if (!d->xf || (long)d->batch * d->dim * d->seqlen < (24L << 20)) {
    return false;
  }
  const size_t eo = (d->out_f32 && d->dtype != NZ_F32) ? 4 : esize(d->dtype);
  return aligned16(d->out) && (d->out_stride[0] * eo) % 16 == 0 && (d->out_stride[1] * eo) % 16 == 0 &&
         (!d->xf || aligned16(d->xf)) && aligned16(d->x) &&
         d->workspace_bytes >= nz_scan_workspace_bytes(d) + rl_extra_bytes(d);
}

static bool rl_bwd_usable(const NzScanDesc* d) {
  const size_t es = esize(d->dtype);
  const int64_t rd[2] = {d->dim, d->batch};
  const int64_t os[2] = {d->dout_stride[1], d->dout_stride[0]};
  return d->xf && rl_shape_ok(d) && tma_ok_rows(d->dout, d->dtype, d->seqlen, rd, os, 2) && aligned16(d->du) &&
         aligned16(d->ddelta) && (!d->dz || aligned16(d->dz)) && aligned16(d->dB) && aligned16(d->dC) &&
         aligned16(d->xf) && (d->seqlen * es) % 128 == 0 && d->workspace_bytes >= nz_scan_workspace_bytes_bwd(d);
}

static int run_scan(const NzScanDesc* d, void* stream, bool bwd) {
  int rc = validate(d, bwd);
  if (rc) return rc;
  if (folded(d) && !(bwd ? rl_bwd_usable(d) : rl_fwd_usable(d)))
    return fail(NZ_EUNSUPPORTED,
                "rev_mask / u_gdiv need the row-per-lane kernels: d_state 16, groups of a multiple of 32 rows, rows a "
                "whole number of 128-byte lines, 16-byte aligned operands, the nz_scan_workspace_bytes_bwd() scratch%s",
                bwd ? " and xf" : "");
  ScanKArgs a;
  fill_args(d, a, bwd);
  if ((long long)a.nrb_total * a.nchunks > 2000000000LL) return fail(NZ_EINVAL, "too many tiles");
  const bool tma = setup_tma(d, a, bwd);
  const bool has_z = d->z != nullptr;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  // (an xf the problem does not qualify for -- nz_scan_fine_bytes() == 0 -- is ignored: the forward does not write it and
  // the backward takes the warp-scan kernels, which only need x)
  const bool use_xf = d->xf && rl_shape_ok(d);
  if (!bwd && rl_fwd_usable(d)) {
    cudaError_t e = d->dtype == NZ_F32    ? run_fwd_rl<float>(d, st)
                    : d->dtype == NZ_BF16 ? run_fwd_rl<__nv_bfloat16>(d, st)
                                          : run_fwd_rl<__half>(d, st);
    if (e != cudaSuccess) return fail(NZ_ECUDA, "scan_fwd (row-per-lane) launch failed: %s", cudaGetErrorString(e));
    return NZ_OK;
  }
  if (!bwd && use_xf) {
    if (!tma) return fail(NZ_EINVAL, "xf given but the forward cannot take the TMA path for this problem");
    a.xf = d->xf;
  }
  if (bwd && rl_bwd_usable(d)) {
    cudaError_t e = d->dtype == NZ_F32    ? run_bwd_rl<float>(d, st)
                    : d->dtype == NZ_BF16 ? run_bwd_rl<__nv_bfloat16>(d, st)
                                          : run_bwd_rl<__half>(d, st);
    if (e != cudaSuccess) return fail(NZ_ECUDA, "scan_bwd (row-per-lane) launch failed: %s", cudaGetErrorString(e));
    return NZ_OK;
  }
  cudaError_t e = cudaMemsetAsync(d->workspace, 0, (size_t)nz_scan_workspace_bytes(d), st);
  if (e != cudaSuccess) return fail(NZ_ECUDA, "workspace memset failed: %s", cudaGetErrorString(e));
  const bool cp = !bwd && cp_eligible(d) && d->workspace_bytes >= nz_scan_workspace_bytes_cp(d);
  if (cp) {
    if (d->dtype == NZ_F32) e = run_fwd_cp<float>(a, d, tma, has_z, st);
    else if (d->dtype == NZ_BF16) e = run_fwd_cp<__nv_bfloat16>(a, d, tma, has_z, st);
    else e = run_fwd_cp<__half>(a, d, tma, has_z, st);
  } else if (d->dtype == NZ_F32)
    e = bwd ? launch_scan_bwd<float>(a, tma, has_z, st) : launch_scan_fwd<float>(a, tma, has_z, st);
  else if (d->dtype == NZ_BF16)
    e = bwd ? launch_scan_bwd<__nv_bfloat16>(a, tma, has_z, st) : launch_scan_fwd<__nv_bfloat16>(a, tma, has_z, st);
  else
    e = bwd ? launch_scan_bwd<__half>(a, tma, has_z, st) : launch_scan_fwd<__half>(a, tma, has_z, st);
  if (e != cudaSuccess)
    return fail(NZ_ECUDA, "%s launch failed: %s (tma=%d)", bwd ? "scan_bwd" : "scan_fwd", cudaGetErrorString(e), (int)tma);
  count_launch(1);  // (the workspace memset is a driver memset node, not one of our kernels)
  return NZ_OK;
}

}  // namespace nz

extern "C" {

int64_t nz_scan_num_chunks(int64_t seqlen) { return (seqlen + NZ_CHUNK - 1) / NZ_CHUNK; }

int64_t nz_scan_workspace_bytes(const NzScanDesc* d) {
  if (!d || d->batch < 1 || d->dim < 1) return 0;
  // ticket header + per (batch, dim) row: 2 ring slots x 16 states x {fp32 value, u32 tag}
  return (int64_t)nz::kWsHeader + (int64_t)d->batch * d->dim * 2 * NZ_MAX_DSTATE * 8;
}

int64_t nz_scan_workspace_bytes_cp(const NzScanDesc* d) {
  const int64_t base = nz_scan_workspace_bytes(d);
  if (base == 0 || d->ngroups < 1 || d->seqlen < 1 || d->dim % d->ngroups) return base;
  int64_t best = nz::cp_eligible(d) ? base + nz::cp_extra_bytes(d) : base;
  if (nz::rl_shape_ok(d)) {  // the row-per-lane forward keeps its chunk aggregates there
    const int64_t rl = base + nz::rl_extra_bytes(d);
    if (rl > best) best = rl;
  }
  return best;
}

int64_t nz_scan_fine_bytes(const NzScanDesc* d) {
  if (!nz::rl_shape_ok(d) || d->seqlen % NZ_FINE) return 0;
  return (int64_t)d->batch * d->dim * (d->seqlen / NZ_FINE) * NZ_MAX_DSTATE * (int64_t)sizeof(float);
}

int64_t nz_scan_workspace_bytes_bwd(const NzScanDesc* d) {
  const int64_t base = nz_scan_workspace_bytes(d);
  if (base == 0 || !d->xf || !nz::rl_shape_ok(d)) return base;
  return base + nz::rl_extra_bytes(d);
}

int nz_scan_bwd_overwrites_dbc(const NzScanDesc* d) {
  if (!d || d->ngroups < 1 || d->dim % d->ngroups) return 0;
  const int dpg = d->dim / d->ngroups;
  if (d->xf && nz::rl_shape_ok(d)) {
    // one warp owns every dB / dC element (32 rows per group), or the backward's aggregate pass zero-fills them on its
    // way (several row blocks per group and more than one chunk: RlArgs::zero_dbc)
    if (dpg == 32) return 1;
    int nc = 1, tpc = 1;
    nz::rl_plan(d, &nc, &tpc, nz::bwd_rl_v2() ? NZ_RL_BWD2_MINB : 12);
    return (nc > 1 && !getenv("NZ_RL_NO_ZERO_DBC")) ? 1 : 0;
  }
  return dpg <= nz::kBwdRows ? 1 : 0;
}

int nz_scan_fwd(const NzScanDesc* desc, void* stream) { return nz::run_scan(desc, stream, false); }

int nz_scan_bwd(const NzScanDesc* desc, void* stream) { return nz::run_scan(desc, stream, true); }

const char* nz_last_error(void) { return nz::g_err; }

int nz_abi_version(void) { return NZ_ABI_VERSION; }

int64_t nz_sizeof_scan_desc(void) { return (int64_t)sizeof(NzScanDesc); }

int nz_set_device(int device) {
  cudaError_t e = cudaSetDevice(device);
  return e == cudaSuccess ? NZ_OK : nz::fail(NZ_ECUDA, "cudaSetDevice(%d): %s", device, cudaGetErrorString(e));
}

int64_t nz_launch_count(void) { return (int64_t)nz::g_launches.load(); }

void nz_debug_set_trace(void* device_buffer) { nz::g_trace = device_buffer; }

// ------------------------------------------------------------------------------------------------
// Host-buffer entry point: stage H2D, run forward (+ backward), copy results D2H.
// ------------------------------------------------------------------------------------------------
#define NZ_CUDA(call)                                                                   \
  do {                                                                                  \
    cudaError_t e_ = (call);                                                            \
    if (e_ != cudaSuccess) {                                                            \
      rc = nz::fail(NZ_ECUDA, "%s failed: %s", #call, cudaGetErrorString(e_));          \
      goto done;                                                                        \
    }                                                                                   \
  } while (0)

int nz_scan_fwd_bwd_host(const NzScanDesc* h, void* stream) {
  if (!h) return nz::fail(NZ_EINVAL, "null descriptor");
  NzScanDesc probe = *h;
  if (!probe.x) probe.x = reinterpret_cast<float*>(16);  // x is optional on the host side
  probe.workspace = reinterpret_cast<void*>(256);        // scratch is allocated here
  probe.workspace_bytes = nz_scan_workspace_bytes(h);
  int rc = nz::validate(&probe, h->dout != nullptr);
  if (rc) return rc;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const size_t es = nz::esize(h->dtype);
  const int64_t Bt = h->batch, Dm = h->dim, N = h->dstate, G = h->ngroups, L = h->seqlen;
  const int64_t nch = nz_scan_num_chunks(L);
  const size_t row_bytes = (size_t)Bt * Dm * L * es, bc_bytes = (size_t)Bt * G * N * L * es;
  const size_t bc_f32 = (size_t)Bt * G * N * L * 4;
  const bool bwd = h->dout != nullptr;
  // the host tensors must be dense for the staged copies
  if (h->u_stride[1] != L || h->u_stride[0] != Dm * L || h->delta_stride[1] != L || h->delta_stride[0] != Dm * L ||
      h->B_stride[2] != L || h->B_stride[1] != N * L || h->B_stride[0] != G * N * L || h->C_stride[2] != L ||
      h->C_stride[1] != N * L || h->C_stride[0] != G * N * L || h->out_stride[1] != L || h->out_stride[0] != Dm * L)
    return nz::fail(NZ_EINVAL, "host entry point needs dense (contiguous) host tensors");
  {
    // keep the stream-ordered pool's memory across calls (default threshold 0 would hand it back
    // to the driver at every synchronisation and re-allocate GBs per call)
    static thread_local int pool_dev = -1;
    int dev = 0;
    cudaGetDevice(&dev);
    if (pool_dev != dev) {
      cudaMemPool_t mp;
      if (cudaDeviceGetDefaultMemPool(&mp, dev) == cudaSuccess) {
        unsigned long long thr = ~0ull;
        cudaMemPoolSetAttribute(mp, cudaMemPoolAttrReleaseThreshold, &thr);
      }
      pool_dev = dev;
    }
  }
  // Pipeline over batch slices on three streams: H2D of slice i+1 overlaps the kernels of slice i and the
  // D2H of slice i-1 (PCIe is full duplex; the staged copies, not the kernels, bound this entry point).
  static thread_local cudaStream_t s_in = nullptr, s_out = nullptr;
  static thread_local cudaEvent_t ev_in[64], ev_cmp[64], ev_start = nullptr, ev_done = nullptr;
  static thread_local bool ev_ready = false;
  char* pool = nullptr;
  size_t off = 0;
  auto carve = [&](size_t bytes) {
    char* p = pool + off;
    off += (bytes + 255) & ~(size_t)255;
    return p;
  };
  if (!ev_ready) {
    NZ_CUDA(cudaStreamCreateWithFlags(&s_in, cudaStreamNonBlocking));
    NZ_CUDA(cudaStreamCreateWithFlags(&s_out, cudaStreamNonBlocking));
    for (int i = 0; i < 64; ++i) {
      NZ_CUDA(cudaEventCreateWithFlags(&ev_in[i], cudaEventDisableTiming));
      NZ_CUDA(cudaEventCreateWithFlags(&ev_cmp[i], cudaEventDisableTiming));
    }
    NZ_CUDA(cudaEventCreateWithFlags(&ev_start, cudaEventDisableTiming));
    NZ_CUDA(cudaEventCreateWithFlags(&ev_done, cudaEventDisableTiming));
    ev_ready = true;
  }
  {
    const int nsl = (int)(Bt < 64 ? Bt : 64);  // one slice per batch entry (at most 64 slices)
    // fine checkpoints + the larger backward scratch when the row-per-lane backward applies (device buffers are
    // 256-byte aligned and dense, so only the shape decides)
    size_t fine_bytes = 0, ws_bytes = (size_t)nz_scan_workspace_bytes(h);
    {
      NzScanDesc q = *h;
      q.u = q.delta = q.B = q.C = reinterpret_cast<void*>(256);
      if (h->z) { q.z = reinterpret_cast<void*>(256); q.z_stride[0] = Dm * L; q.z_stride[1] = L; }
      if (bwd) fine_bytes = (size_t)nz_scan_fine_bytes(&q);
      if (fine_bytes) q.xf = reinterpret_cast<float*>(256);
      for (int64_t nb : {Bt / nsl, (Bt + nsl - 1) / nsl}) {
        q.batch = (int32_t)nb;
        size_t w = (size_t)nz_scan_workspace_bytes_cp(&q);
        if (w > ws_bytes) ws_bytes = w;
        w = (size_t)nz_scan_workspace_bytes_bwd(&q);
        if (w > ws_bytes) ws_bytes = w;
      }
    }
    const size_t total = 8 * (row_bytes + 256) + 2 * (bc_bytes + 256) + 2 * (bc_f32 + 256) +
                         (size_t)Bt * Dm * nch * N * 4 + 6 * ((size_t)Dm * N * 4 + 256) + 4096 +
                         ws_bytes + 256 + fine_bytes + 256;
    NZ_CUDA(cudaMallocAsync(reinterpret_cast<void**>(&pool), total, st));
    char* du_ = carve(row_bytes); char* dd_ = carve(row_bytes); char* dz_ = carve(row_bytes);
    char* u_ = carve(row_bytes); char* dl_ = carve(row_bytes); char* z_ = carve(row_bytes);
    char* out_ = carve(row_bytes); char* go_ = carve(row_bytes);
    char* B_ = carve(bc_bytes); char* C_ = carve(bc_bytes);
    char* dB_ = carve(bc_f32); char* dC_ = carve(bc_f32);
    char* x_ = carve((size_t)Bt * Dm * nch * N * 4);
    char* A_ = carve((size_t)Dm * N * 4); char* dA_ = carve((size_t)Dm * N * 4);
    char* D_ = carve((size_t)Dm * 4); char* bias_ = carve((size_t)Dm * 4);
    char* dD_ = carve((size_t)Dm * 4); char* db_ = carve((size_t)Dm * 4);
    char* ws_ = carve(ws_bytes);
    char* xf_ = fine_bytes ? carve(fine_bytes) : nullptr;
    // the side streams start after the allocation (and whatever precedes this call on `stream`)
    NZ_CUDA(cudaEventRecord(ev_start, st));
    NZ_CUDA(cudaStreamWaitEvent(s_in, ev_start, 0));
    NZ_CUDA(cudaStreamWaitEvent(s_out, ev_start, 0));
    // parameters + zeroed accumulators first
    NZ_CUDA(cudaMemcpyAsync(A_, h->A, (size_t)Dm * N * 4, cudaMemcpyHostToDevice, s_in));
    if (h->D) NZ_CUDA(cudaMemcpyAsync(D_, h->D, (size_t)Dm * 4, cudaMemcpyHostToDevice, s_in));
    if (h->delta_bias) NZ_CUDA(cudaMemcpyAsync(bias_, h->delta_bias, (size_t)Dm * 4, cudaMemcpyHostToDevice, s_in));
    if (bwd) {
      NZ_CUDA(cudaMemsetAsync(dA_, 0, (size_t)Dm * N * 4, s_in));
      NZ_CUDA(cudaMemsetAsync(dB_, 0, bc_f32, s_in));
      NZ_CUDA(cudaMemsetAsync(dC_, 0, bc_f32, s_in));
      NZ_CUDA(cudaMemsetAsync(dD_, 0, (size_t)Dm * 4, s_in));
      NZ_CUDA(cudaMemsetAsync(db_, 0, (size_t)Dm * 4, s_in));
    }
    for (int i = 0; i < nsl; ++i) {
      const int64_t b0 = Bt * i / nsl, b1 = Bt * (i + 1) / nsl, nb = b1 - b0;
      const size_t ro = (size_t)b0 * Dm * L * es, rb = (size_t)nb * Dm * L * es;
      const size_t bo = (size_t)b0 * G * N * L * es, bb = (size_t)nb * G * N * L * es;
      const size_t fo = (size_t)b0 * G * N * L * 4, fb = (size_t)nb * G * N * L * 4;
      const size_t xo = (size_t)b0 * Dm * nch * N * 4, xb = (size_t)nb * Dm * nch * N * 4;
      auto hp = [](const void* p, size_t o) { return static_cast<const char*>(p) + o; };
      auto hq = [](void* p, size_t o) { return static_cast<char*>(p) + o; };
      // ---- host -> device ----
      NZ_CUDA(cudaMemcpyAsync(u_ + ro, hp(h->u, ro), rb, cudaMemcpyHostToDevice, s_in));
      NZ_CUDA(cudaMemcpyAsync(dl_ + ro, hp(h->delta, ro), rb, cudaMemcpyHostToDevice, s_in));
      NZ_CUDA(cudaMemcpyAsync(B_ + bo, hp(h->B, bo), bb, cudaMemcpyHostToDevice, s_in));
      NZ_CUDA(cudaMemcpyAsync(C_ + bo, hp(h->C, bo), bb, cudaMemcpyHostToDevice, s_in));
      if (h->z) NZ_CUDA(cudaMemcpyAsync(z_ + ro, hp(h->z, ro), rb, cudaMemcpyHostToDevice, s_in));
      if (bwd) NZ_CUDA(cudaMemcpyAsync(go_ + ro, hp(h->dout, ro), rb, cudaMemcpyHostToDevice, s_in));
      NZ_CUDA(cudaEventRecord(ev_in[i], s_in));
      // ---- kernels on the caller's stream ----
      NZ_CUDA(cudaStreamWaitEvent(st, ev_in[i], 0));
      NzScanDesc d = *h;
      d.batch = (int32_t)nb;
      d.u = u_ + ro; d.delta = dl_ + ro; d.B = B_ + bo; d.C = C_ + bo;
      d.A = reinterpret_cast<float*>(A_); d.A_stride = N;
      d.D = h->D ? reinterpret_cast<float*>(D_) : nullptr;
      d.delta_bias = h->delta_bias ? reinterpret_cast<float*>(bias_) : nullptr;
      d.z = h->z ? z_ + ro : nullptr; d.z_stride[0] = Dm * L; d.z_stride[1] = L;
      d.out = out_ + ro; d.x = reinterpret_cast<float*>(x_ + xo);
      d.workspace = ws_; d.workspace_bytes = (int64_t)ws_bytes;
      d.xf = xf_ ? reinterpret_cast<float*>(xf_ + (size_t)b0 * Dm * (L / NZ_FINE) * N * 4) : nullptr;
      rc = nz_scan_fwd(&d, stream);
      if (rc) goto done;
      if (bwd) {
        d.dout = go_ + ro; d.dout_stride[0] = Dm * L; d.dout_stride[1] = L;
        d.du = du_ + ro; d.ddelta = dd_ + ro; d.dz = h->z ? dz_ + ro : nullptr;
        d.dA = reinterpret_cast<float*>(dA_);
        d.dB = reinterpret_cast<float*>(dB_ + fo); d.dC = reinterpret_cast<float*>(dC_ + fo);
        d.dD = h->dD ? reinterpret_cast<float*>(dD_) : nullptr;
        d.ddelta_bias = h->ddelta_bias ? reinterpret_cast<float*>(db_) : nullptr;
        rc = nz_scan_bwd(&d, stream);
        if (rc) goto done;
      }
      NZ_CUDA(cudaEventRecord(ev_cmp[i], st));
      // ---- device -> host ----
      NZ_CUDA(cudaStreamWaitEvent(s_out, ev_cmp[i], 0));
      NZ_CUDA(cudaMemcpyAsync(hq(h->out, ro), out_ + ro, rb, cudaMemcpyDeviceToHost, s_out));
      if (h->x) NZ_CUDA(cudaMemcpyAsync(hq(h->x, xo), x_ + xo, xb, cudaMemcpyDeviceToHost, s_out));
      if (bwd) {
        NZ_CUDA(cudaMemcpyAsync(hq(h->du, ro), du_ + ro, rb, cudaMemcpyDeviceToHost, s_out));
        NZ_CUDA(cudaMemcpyAsync(hq(h->ddelta, ro), dd_ + ro, rb, cudaMemcpyDeviceToHost, s_out));
        if (h->z) NZ_CUDA(cudaMemcpyAsync(hq(h->dz, ro), dz_ + ro, rb, cudaMemcpyDeviceToHost, s_out));
        NZ_CUDA(cudaMemcpyAsync(hq(h->dB, fo), dB_ + fo, fb, cudaMemcpyDeviceToHost, s_out));
        NZ_CUDA(cudaMemcpyAsync(hq(h->dC, fo), dC_ + fo, fb, cudaMemcpyDeviceToHost, s_out));
      }
    }
    if (bwd) {  // the (dim)-shaped sums are complete after the last slice
      NZ_CUDA(cudaMemcpyAsync(h->dA, dA_, (size_t)Dm * N * 4, cudaMemcpyDeviceToHost, s_out));
      if (h->dD) NZ_CUDA(cudaMemcpyAsync(h->dD, dD_, (size_t)Dm * 4, cudaMemcpyDeviceToHost, s_out));
      if (h->ddelta_bias) NZ_CUDA(cudaMemcpyAsync(h->ddelta_bias, db_, (size_t)Dm * 4, cudaMemcpyDeviceToHost, s_out));
    }
    // join the side streams back into the caller's stream
    NZ_CUDA(cudaEventRecord(ev_done, s_out));
    NZ_CUDA(cudaStreamWaitEvent(st, ev_done, 0));
  }
done:
  if (ev_ready) {  // nothing of this call may still be in flight when the pool memory is released
    cudaStreamSynchronize(s_in);
    cudaStreamSynchronize(s_out);
  }
  if (pool) cudaFreeAsync(pool, st);
  {
    cudaError_t e2 = cudaStreamSynchronize(st);
    if (!rc && e2 != cudaSuccess) rc = nz::fail(NZ_ECUDA, "stream sync failed: %s", cudaGetErrorString(e2));
  }
  return rc;
}

}  // extern "C"
