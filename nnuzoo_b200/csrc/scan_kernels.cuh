// scan_kernels.cuh -- selective-scan forward / backward for sm_100a.
//
// What is computed (reference statement: nnunetv2/nets/seg_mamba/selective_scan_interface.py:86-152):
//   dl_t   = softplus(delta_t + bias)                      (:103-107)
//   a_t[n] = exp(dl_t * A[n])                              (:121)
//   h_t[n] = a_t[n] * h_{t-1}[n] + dl_t * B_t[n] * u_t     (:129,:134)
//   y_t    = sum_n C_t[n] * h_t[n] + D * u_t               (:141,:148)
//   out_t  = y_t * silu(z_t)                               (:150)
// and its reverse-time adjoint for the backward.
//
// Parallel decomposition (DESIGN.md "Kernels"):
//   * one CTA = R rows (b, d) of ONE (batch, group), so the B_t/C_t tile is staged once in shared
//     memory and shared by all R rows (the upstream one-row-per-CTA scheme re-reads B/C per row);
//   * one row = LPR lanes of a warp, each lane owns M consecutive time steps of the current chunk
//     (chunk length TL = LPR * M = NZ_CHUNK); per state n a lane folds its M steps sequentially, the
//     LPR lane aggregates are combined with a Kogge-Stone scan over warp shuffles, and the carry
//     across chunks lives in shared memory (one float per (row, n));
//   * the product of a over a lane's segment is exp2(A2 * sum(dl)) -- one MUFU instead of M
//     multiplies; every element-wise product runs as packed fp32x2 (FFMA2/FMUL2) over time pairs,
//     only the four sequential recurrences stay scalar FFMA chains;
//   * tiles are moved with TMA (cp.async.bulk.tensor, 128B swizzle): the per-row tiles are
//     single-buffered and refilled for the next chunk as soon as every lane holds its values in
//     registers, the B/C tiles are double-buffered, each on its own mbarrier; a generic cooperative
//     loader with the same smem layout serves shapes TMA cannot express;
//   * the forward writes h at the end of every chunk (the reference ABI's `x`); the backward walks
//     the chunks last-to-first, recomputes h inside a chunk from that checkpoint, runs the reverse
//     scan for dh interleaved with the forward one, accumulates dA / dD / d(bias) in fp32 and reduces
//     dB / dC over the R rows of the CTA in shared memory before touching global memory.
#pragma once

#include "nz_common.cuh"

// tuning switches (tools/tune_build.py builds variants)
#ifndef NZ_FWD_UNROLL
#define NZ_FWD_UNROLL 1
#endif
#ifndef NZ_BWD_UNROLL
#define NZ_BWD_UNROLL 1
#endif
#ifndef NZ_BWD_SPLITBAR
#define NZ_BWD_SPLITBAR 1
#endif
#ifndef NZ_NOBAR
#define NZ_NOBAR 0  // 1: no CTA-wide barrier in the chunk loop (last warp out issues the TMA refill);
                    // measured on B200: no gain fwd, -8% bwd (profiles/r01_kernel_tuning.md)
#endif
// (tried and removed: issuing the next state pair's exps behind the current pair's fold -- measured
//  -8%: MUFU and SHFL share the in-order MIO queue, nothing overlaps; profiles/r01_kernel_tuning.md)
#ifndef NZ_BWD_KEEPB
#define NZ_BWD_KEEPB 1  // 1: keep B_t[n] in registers instead of re-reading the tile for sum_n dh*B
#endif
#define NZ_PRAGMA_(x) _Pragma(#x)
#define NZ_UNROLL(n) NZ_PRAGMA_(unroll n)

namespace nz {

struct alignas(64) ScanKArgs {
  CUtensorMap tm_u, tm_delta, tm_z, tm_dout, tm_B, tm_C;
  const void *u, *delta, *z, *dout, *B, *C;
  const float *A, *D, *bias;
  void *out, *du, *ddelta, *dz;
  float *x, *dA, *dB, *dC, *dD, *dbias;
  long L;
  long u_bs, u_ds, dl_bs, dl_ds, z_bs, z_ds, o_bs, o_ds, do_bs, do_ds;
  long B_bs, B_gs, B_ns, C_bs, C_gs, C_ns, A_ds;
  int batch, dim, dstate, ngroups, dpg, nchunks, softplus;
  int vec_out;   // out rows 16-byte aligned -> vector stores
  int vec_grad;  // du/ddelta/dz and dB/dC rows 16-byte aligned
};

template <typename T, int M, int LPR, int WARPS, bool kHasZ, bool kBwd>
struct ScanCfg {
  static constexpr int RPW = 32 / LPR;            // rows per warp
  static constexpr int R = RPW * WARPS;           // rows per CTA
  static constexpr int TL = LPR * M;              // chunk length
  static constexpr int ES = sizeof(T);
  static constexpr int ROWB = TL * ES;            // bytes of one tile row
  static constexpr int SEGB = M * ES;             // bytes of one lane's segment
  static constexpr int INNER = 128 / ES;          // elements in one 128-byte TMA box row
  static constexpr int ROWTILE = R * ROWB;
  static constexpr int BCTILE = kMaxState * ROWB;
  static constexpr int NROWT = (kBwd ? 3 : 2) + (kHasZ ? 1 : 0);  // u, delta, [dout], [z]
  static constexpr int CKB = kBwd ? R * kMaxState * 4 : 0;        // checkpoint carry-in landing (bwd)
  static constexpr int ROWS_TX = NROWT * ROWTILE;                 // bytes TMA delivers per chunk (rows)
  static constexpr int ROWS_REGION = ((ROWS_TX + CKB + 1023) / 1024) * 1024;
  static constexpr int BC_TX = 2 * BCTILE;                        // bytes per B/C stage
  static constexpr int SLAB = kBwd ? R * TL * 4 : 0;              // fp32 [R][TL]
  static constexpr int SMALL = 64 + (kBwd ? 5 : 2) * R * kMaxState * 4;
  static_assert(TL == 256, "chunk length must equal NZ_CHUNK");
  // TMA (hardware swizzle keyed on absolute smem address bits) is only used with R >= 8 rows
  static_assert(R < 8 || ROWTILE % 1024 == 0, "row tiles must keep the 1024-byte swizzle alignment");
  static_assert(SEGB >= 16 && SEGB % 16 == 0, "a lane's segment must be whole 16-byte vectors");
  static_assert(M % 2 == 0, "packed fp32x2 math works on time pairs");
  static constexpr size_t smem_bytes(bool tma) {
    return 1024 + (size_t)ROWS_REGION + (size_t)(tma ? 2 : 1) * BC_TX + 2 * (size_t)SLAB + SMALL;
  }
};

// ---- packed fp32x2 helpers (FFMA2 / FMUL2 / FADD2 on sm_100) ----
__device__ __forceinline__ float2 f2(float a, float b) { return make_float2(a, b); }
__device__ __forceinline__ float2 mul2(float2 a, float2 b) { return __fmul2_rn(a, b); }
__device__ __forceinline__ float2 fma2(float2 a, float2 b, float2 c) { return __ffma2_rn(a, b, c); }
__device__ __forceinline__ float2 sub2(float2 a, float2 b) { return __fadd2_rn(a, f2(-b.x, -b.y)); }

// Hand a shared-memory buffer back to the TMA producer: returns true in the one thread that must
// issue the refill.
template <int WARPS>
__device__ __forceinline__ bool release_buffer(unsigned* cnt, int lane, int tid) {
#if NZ_NOBAR
  return warp_last_arrival(cnt, WARPS, lane) && lane == 0;
#else
  __syncthreads();
  return tid == 0;
#endif
}

// Cooperative (non-TMA) fill of one swizzled tile: nrows_tile rows of TL elements, zero padded
// beyond `nrows_valid` rows and beyond L.
template <typename T, int TL>
__device__ __forceinline__ void coop_fill(uint8_t* tile, const T* base, long row_stride, int nrows_valid,
                                          int nrows_tile, long tstart, long L) {
  constexpr int ROWB = TL * (int)sizeof(T);
  for (int idx = threadIdx.x; idx < nrows_tile * TL; idx += blockDim.x) {
    const int r = idx / TL, t = idx - r * TL;
    const long gl = tstart + t;
    T v = Elem<T>::from_f(0.f);
    if (r < nrows_valid && gl < L) v = base[(long)r * row_stride + gl];
    *reinterpret_cast<T*>(tile + swz128((uint32_t)(r * ROWB + t * (int)sizeof(T)))) = v;
  }
}

// A lane's M items out of a swizzled dense-row tile.  When a tile row is a multiple of 1024 bytes the
// swizzle term does not depend on the row, so the per-thread in-row offsets `pre` are loop invariant.
template <typename T, int M, int ROWB>
__device__ __forceinline__ void lds_seg(const uint8_t* tile, int row, uint32_t segoff, const uint32_t (&pre)[4],
                                        float (&v)[M]) {
  constexpr int kPer = 16 / (int)sizeof(T);
  constexpr int kVec = M / kPer;
  static_assert(kVec <= 4, "at most four 16-byte vectors per lane segment");
#pragma unroll
  for (int j = 0; j < kVec; ++j) {
    uint32_t off;
    if constexpr (ROWB % 1024 == 0)
      off = (uint32_t)row * ROWB + pre[j];
    else
      off = swz128((uint32_t)row * ROWB + segoff + 16u * j);
    unpack16<T>(*reinterpret_cast<const uint4*>(tile + off), &v[j * kPer]);
  }
}

// ================================================================================================
// Forward
// ================================================================================================
template <typename T, int M, int LPR, int WARPS, int NQ, bool kTMA, bool kHasZ>
__global__ void __launch_bounds__(WARPS * 32, WARPS >= 8 ? 2 : 1) scan_fwd_kernel(const __grid_constant__ ScanKArgs a) {
  using Cfg = ScanCfg<T, M, LPR, WARPS, kHasZ, false>;
  constexpr int R = Cfg::R, TL = Cfg::TL, ROWB = Cfg::ROWB, SEGB = Cfg::SEGB;
  constexpr int ROWTILE = Cfg::ROWTILE, BCTILE = Cfg::BCTILE;
  [[maybe_unused]] constexpr int NROWT = Cfg::NROWT;
  constexpr int NT = WARPS * 32;
  constexpr int H2 = M / 2;

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* rows = smem;
  uint8_t* bcs = smem + Cfg::ROWS_REGION;
  uint8_t* tail = bcs + (kTMA ? 2 : 1) * Cfg::BC_TX;
  uint64_t* bars = reinterpret_cast<uint64_t*>(tail);  // [0] rows, [1],[2] B/C stages
  unsigned* cnts = reinterpret_cast<unsigned*>(tail + 40);  // buffer-release counters: rows, B/C 0, B/C 1
  float* sm_A2 = reinterpret_cast<float*>(tail + 64);
  float* sm_h = sm_A2 + R * kMaxState;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int sl = lane % LPR;                      // segment (time) index inside the row
  const int rloc = warp * Cfg::RPW + lane / LPR;  // row inside the CTA
  const int bpg = a.dpg / R;
  const int rb = blockIdx.x % bpg;
  const int g = (blockIdx.x / bpg) % a.ngroups;
  const int b = blockIdx.x / (bpg * a.ngroups);
  const int d0 = g * a.dpg + rb * R;
  const int d = d0 + rloc;
  const int N = a.dstate;

  for (int i = tid; i < R * kMaxState; i += NT) {
    const int r = i / kMaxState, n = i % kMaxState;
    sm_A2[i] = n < N ? a.A[(long)(d0 + r) * a.A_ds + n] * kLog2e : 0.f;
    sm_h[i] = 0.f;
  }

  auto issue_rows = [&](int c) {
    mbar_arrive_expect_tx(&bars[0], Cfg::ROWS_TX);
    const int c1 = c * (TL / Cfg::INNER);
    tma_load_4d(rows, &a.tm_u, &bars[0], 0, c1, d0, b);
    tma_load_4d(rows + ROWTILE, &a.tm_delta, &bars[0], 0, c1, d0, b);
    if (kHasZ) tma_load_4d(rows + 2 * ROWTILE, &a.tm_z, &bars[0], 0, c1, d0, b);
  };
  auto issue_bc = [&](int c, int s) {
    uint8_t* st = bcs + s * Cfg::BC_TX;
    mbar_arrive_expect_tx(&bars[1 + s], Cfg::BC_TX);
    const int c1 = c * (TL / Cfg::INNER);
    tma_load_5d(st, &a.tm_B, &bars[1 + s], 0, c1, 0, g, b);
    tma_load_5d(st + BCTILE, &a.tm_C, &bars[1 + s], 0, c1, 0, g, b);
  };

  if (tid == 0) {
    mbar_init(&bars[0], 1);
    mbar_init(&bars[1], 1);
    mbar_init(&bars[2], 1);
    cnts[0] = cnts[1] = cnts[2] = 0u;
    fence_mbar_init();
  }
  __syncthreads();
  if (kTMA && tid == 0) {
    issue_rows(0);
    issue_bc(0, 0);
    if (a.nchunks > 1) issue_bc(1, 1);
  }

  const float Dv = a.D ? a.D[d] : 0.f;
  const float bias = a.bias ? a.bias[d] : 0.f;
  T* outrow = reinterpret_cast<T*>(a.out) + (long)b * a.o_bs + (long)d * a.o_ds;
  float* xrow = a.x + ((long)b * a.dim + d) * (long)a.nchunks * N;
  const uint32_t segoff = sl * SEGB;
  const uint32_t pre[4] = {swz128(segoff), swz128(segoff + 16u), swz128(segoff + 32u), swz128(segoff + 48u)};
  const int NP = (N + NQ - 1) / NQ * NQ;  // states are processed NQ at a time; rows >= N of the tiles are zero

  for (int c = 0; c < a.nchunks; ++c) {
    const int s = kTMA ? (c & 1) : 0;
    if (kTMA) {
      mbar_wait(&bars[0], c & 1);
    } else {
      __syncthreads();
      const long ts = (long)c * TL;
      coop_fill<T, TL>(rows, reinterpret_cast<const T*>(a.u) + (long)b * a.u_bs + (long)d0 * a.u_ds, a.u_ds, R, R, ts, a.L);
      coop_fill<T, TL>(rows + ROWTILE, reinterpret_cast<const T*>(a.delta) + (long)b * a.dl_bs + (long)d0 * a.dl_ds,
                       a.dl_ds, R, R, ts, a.L);
      if (kHasZ)
        coop_fill<T, TL>(rows + 2 * ROWTILE, reinterpret_cast<const T*>(a.z) + (long)b * a.z_bs + (long)d0 * a.z_ds,
                         a.z_ds, R, R, ts, a.L);
      coop_fill<T, TL>(bcs, reinterpret_cast<const T*>(a.B) + (long)b * a.B_bs + (long)g * a.B_gs, a.B_ns, N,
                       kMaxState, ts, a.L);
      coop_fill<T, TL>(bcs + BCTILE, reinterpret_cast<const T*>(a.C) + (long)b * a.C_bs + (long)g * a.C_gs, a.C_ns, N,
                       kMaxState, ts, a.L);
      __syncthreads();
    }
    const long t0 = (long)c * TL + sl * M;

    float dlu[M], dl[M], y[M], zz[kHasZ ? M : 1];
    lds_seg<T, M, ROWB>(rows, rloc, segoff, pre, dlu);
    lds_seg<T, M, ROWB>(rows + ROWTILE, rloc, segoff, pre, dl);
    if constexpr (kHasZ) lds_seg<T, M, ROWB>(rows + 2 * ROWTILE, rloc, segoff, pre, zz);
    float dlsum = 0.f;
#pragma unroll
    for (int i = 0; i < M; ++i) {
      float x = dl[i] + bias;
      if (a.softplus) x = softplus_f(x);
      if (t0 + i >= a.L) x = 0.f;  // beyond the sequence: a = 1, b = 0 -> state passes through
      dl[i] = x;
      dlsum += x;
      y[i] = Dv * dlu[i];
      dlu[i] = x * dlu[i];
    }
    if (kTMA) {
      // the row values now live in registers; the last warp to get here refills the row tiles
      // (no CTA-wide barrier in this loop: warps drift apart and overlap each other's latencies)
      if (release_buffer<WARPS>(&cnts[0], lane, tid) && c + 1 < a.nchunks) issue_rows(c + 1);
      mbar_wait(&bars[1 + s], (c >> 1) & 1);
    }
    const uint8_t* tB = bcs + s * Cfg::BC_TX;
    const uint8_t* tC = tB + BCTILE;

#pragma unroll 1
    for (int n = 0; n < NP; n += NQ) {  // NQ independent states per trip (instruction-level parallelism)
      float hc[NQ], P[NQ], H[NQ];
      float av[NQ][M], bv[NQ][M];
#pragma unroll
      for (int q = 0; q < NQ; ++q) {
        const float A2 = sm_A2[rloc * kMaxState + n + q];
        hc[q] = sm_h[rloc * kMaxState + n + q];
        P[q] = ex2_approx(A2 * dlsum);  // product of a over this lane's segment
        lds_seg<T, M, ROWB>(tB, n + q, segoff, pre, bv[q]);
#pragma unroll
        for (int k = 0; k < H2; ++k) {
          const float2 x2 = mul2(f2(dl[2 * k], dl[2 * k + 1]), f2(A2, A2));
          av[q][2 * k] = ex2_approx(x2.x);
          av[q][2 * k + 1] = ex2_approx(x2.y);
          const float2 b2 = mul2(f2(dlu[2 * k], dlu[2 * k + 1]), f2(bv[q][2 * k], bv[q][2 * k + 1]));
          bv[q][2 * k] = b2.x;
          bv[q][2 * k + 1] = b2.y;
        }
        H[q] = 0.f;
      }
#pragma unroll
      for (int i = 0; i < M; ++i) {
#pragma unroll
        for (int q = 0; q < NQ; ++q) H[q] = fmaf(av[q][i], H[q], bv[q][i]);
      }
#pragma unroll
      for (int off = 1; off < LPR; off <<= 1) {
#pragma unroll
        for (int q = 0; q < NQ; ++q) ks_up<LPR>(P[q], H[q], off);
      }
      float h[NQ], hnew[NQ];
#pragma unroll
      for (int q = 0; q < NQ; ++q) {
        h[q] = ks_enter_up<LPR>(P[q], H[q], hc[q]);
        hnew[q] = fmaf(P[q], hc[q], H[q]);  // state at the end of this lane's segment
      }
#pragma unroll
      for (int i = 0; i < M; ++i) {  // replay with the true carry-in; bv is overwritten by h_t
#pragma unroll
        for (int q = 0; q < NQ; ++q) {
          h[q] = fmaf(av[q][i], h[q], bv[q][i]);
          bv[q][i] = h[q];
        }
      }
#pragma unroll
      for (int q = 0; q < NQ; ++q) {
        float cv[M];
        lds_seg<T, M, ROWB>(tC, n + q, segoff, pre, cv);
#pragma unroll
        for (int k = 0; k < H2; ++k) {
          float2 y2 = f2(y[2 * k], y[2 * k + 1]);
          y2 = fma2(f2(cv[2 * k], cv[2 * k + 1]), f2(bv[q][2 * k], bv[q][2 * k + 1]), y2);
          y[2 * k] = y2.x;
          y[2 * k + 1] = y2.y;
        }
      }
      __syncwarp();
      if (sl == LPR - 1) {
#pragma unroll
        for (int q = 0; q < NQ; ++q) {
          sm_h[rloc * kMaxState + n + q] = hnew[q];
          if (n + q < N) xrow[(long)c * N + n + q] = hnew[q];
        }
      }
    }

    if constexpr (kHasZ) {
#pragma unroll
      for (int i = 0; i < M; ++i) y[i] *= zz[i] * sigmoid_f(zz[i]);
    }
    stg_items<T, M>(outrow, y, t0, a.L, a.vec_out != 0);

    if (kTMA) {
      // last warp done with B/C stage s refills it for chunk c + 2
      if (release_buffer<WARPS>(&cnts[1 + s], lane, tid) && c + 2 < a.nchunks) issue_bc(c + 2, s);
    }
  }
}

// ================================================================================================
// Backward
// ================================================================================================
template <typename T, int M, int LPR, int WARPS, bool kTMA, bool kHasZ>
__global__ void __launch_bounds__(WARPS * 32, (WARPS >= 8 && !kHasZ) ? 2 : 1)
    scan_bwd_kernel(const __grid_constant__ ScanKArgs a) {
  using Cfg = ScanCfg<T, M, LPR, WARPS, kHasZ, true>;
  constexpr int R = Cfg::R, TL = Cfg::TL, ROWB = Cfg::ROWB, SEGB = Cfg::SEGB;
  constexpr int ROWTILE = Cfg::ROWTILE, BCTILE = Cfg::BCTILE;
  [[maybe_unused]] constexpr int NROWT = Cfg::NROWT;
  constexpr int NT = WARPS * 32;
  constexpr int SLROW = TL * 4;  // bytes of one fp32 slab row
  constexpr int H2 = M / 2;

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* rows = smem;
  float* ckl = reinterpret_cast<float*>(smem + Cfg::ROWS_TX);  // TMA landing of the checkpoint carry-in
  uint8_t* bcs = smem + Cfg::ROWS_REGION;
  uint8_t* slabB = bcs + (kTMA ? 2 : 1) * Cfg::BC_TX;
  uint8_t* slabC = slabB + Cfg::SLAB;
  uint8_t* tail = slabC + Cfg::SLAB;
  uint64_t* bars = reinterpret_cast<uint64_t*>(tail);  // [0] rows, [1],[2] B/C stages, [3] slab full, [4] slab free
  unsigned* cnts = reinterpret_cast<unsigned*>(tail + 40);  // buffer-release counters: rows, B/C 0, B/C 1
  float* sm_A2 = reinterpret_cast<float*>(tail + 64);
  float* sm_hc = sm_A2 + R * kMaxState;   // h carried into the current chunk (from the checkpoints)
  float* sm_dhc = sm_hc + R * kMaxState;  // dh carried in from the (later) chunk processed before
  float* sm_dA = sm_dhc + R * kMaxState;  // dA accumulator over chunks
  float* sm_dln = sm_dA + R * kMaxState;  // dl of the first step of that later chunk, per row

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int sl = lane % LPR;
  const int rloc = warp * Cfg::RPW + lane / LPR;
  const int bpg = a.dpg / R;
  const int rb = blockIdx.x % bpg;
  const int g = (blockIdx.x / bpg) % a.ngroups;
  const int b = blockIdx.x / (bpg * a.ngroups);
  const int d0 = g * a.dpg + rb * R;
  const int d = d0 + rloc;
  const int N = a.dstate;

  for (int i = tid; i < R * kMaxState; i += NT) {
    const int r = i / kMaxState, n = i % kMaxState;
    sm_A2[i] = n < N ? a.A[(long)(d0 + r) * a.A_ds + n] * kLog2e : 0.f;
    sm_hc[i] = 0.f;
    sm_dhc[i] = 0.f;
    sm_dA[i] = 0.f;
    sm_dln[i] = 0.f;
  }

  auto issue_rows = [&](int c) {
    mbar_arrive_expect_tx(&bars[0], Cfg::ROWS_TX + (c > 0 ? R * kMaxState * 4 : 0));
    const int c1 = c * (TL / Cfg::INNER);
    tma_load_4d(rows, &a.tm_u, &bars[0], 0, c1, d0, b);
    tma_load_4d(rows + ROWTILE, &a.tm_delta, &bars[0], 0, c1, d0, b);
    tma_load_4d(rows + 2 * ROWTILE, &a.tm_dout, &bars[0], 0, c1, d0, b);
    if (kHasZ) tma_load_4d(rows + 3 * ROWTILE, &a.tm_z, &bars[0], 0, c1, d0, b);
    if (c > 0) {  // state at the end of chunk c-1 == carry into chunk c (TMA path has N == 16)
      for (int r = 0; r < R; ++r)
        bulk_load_1d(ckl + r * kMaxState, a.x + (((long)b * a.dim + d0 + r) * a.nchunks + (c - 1)) * kMaxState,
                     kMaxState * 4, &bars[0]);
    }
  };
  auto issue_bc = [&](int c, int s) {
    uint8_t* st = bcs + s * Cfg::BC_TX;
    mbar_arrive_expect_tx(&bars[1 + s], Cfg::BC_TX);
    const int c1 = c * (TL / Cfg::INNER);
    tma_load_5d(st, &a.tm_B, &bars[1 + s], 0, c1, 0, g, b);
    tma_load_5d(st + BCTILE, &a.tm_C, &bars[1 + s], 0, c1, 0, g, b);
  };

  if (tid == 0) {
    mbar_init(&bars[0], 1);
    mbar_init(&bars[1], 1);
    mbar_init(&bars[2], 1);
    mbar_init(&bars[3], NT);
    mbar_init(&bars[4], NT);
    cnts[0] = cnts[1] = cnts[2] = 0u;
    fence_mbar_init();
  }
  __syncthreads();
  if (kTMA && tid == 0) {
    issue_rows(a.nchunks - 1);
    issue_bc(a.nchunks - 1, 0);
    if (a.nchunks > 1) issue_bc(a.nchunks - 2, 1);
  }

  const float Dv = a.D ? a.D[d] : 0.f;
  const float bias = a.bias ? a.bias[d] : 0.f;
  const long rowlin = ((long)b * a.dim + d) * a.L;  // gradients are contiguous (batch, dim, L)
  T* durow = reinterpret_cast<T*>(a.du) + rowlin;
  T* ddrow = reinterpret_cast<T*>(a.ddelta) + rowlin;
  T* dzrow = kHasZ ? reinterpret_cast<T*>(a.dz) + rowlin : nullptr;
  const uint32_t segoff = sl * SEGB;
  const uint32_t pre[4] = {swz128(segoff), swz128(segoff + 16u), swz128(segoff + 32u), swz128(segoff + 48u)};
  const uint32_t slab_off = rloc * SLROW + sl * (M * 4);
  float dD_acc = 0.f, db_acc = 0.f;
  [[maybe_unused]] unsigned gtrip = 0;  // states processed so far (parity of the slab hand-off barriers)

  for (int it = 0; it < a.nchunks; ++it) {
    const int c = a.nchunks - 1 - it;
    const int s = kTMA ? (it & 1) : 0;
    if (kTMA) {
      mbar_wait(&bars[0], it & 1);
    } else {
      __syncthreads();
      const long ts = (long)c * TL;
      coop_fill<T, TL>(rows, reinterpret_cast<const T*>(a.u) + (long)b * a.u_bs + (long)d0 * a.u_ds, a.u_ds, R, R, ts, a.L);
      coop_fill<T, TL>(rows + ROWTILE, reinterpret_cast<const T*>(a.delta) + (long)b * a.dl_bs + (long)d0 * a.dl_ds,
                       a.dl_ds, R, R, ts, a.L);
      coop_fill<T, TL>(rows + 2 * ROWTILE, reinterpret_cast<const T*>(a.dout) + (long)b * a.do_bs + (long)d0 * a.do_ds,
                       a.do_ds, R, R, ts, a.L);
      if (kHasZ)
        coop_fill<T, TL>(rows + 3 * ROWTILE, reinterpret_cast<const T*>(a.z) + (long)b * a.z_bs + (long)d0 * a.z_ds,
                         a.z_ds, R, R, ts, a.L);
      coop_fill<T, TL>(bcs, reinterpret_cast<const T*>(a.B) + (long)b * a.B_bs + (long)g * a.B_gs, a.B_ns, N,
                       kMaxState, ts, a.L);
      coop_fill<T, TL>(bcs + BCTILE, reinterpret_cast<const T*>(a.C) + (long)b * a.C_bs + (long)g * a.C_gs, a.C_ns, N,
                       kMaxState, ts, a.L);
      for (int i = tid; i < R * kMaxState; i += NT) {
        const int r = i / kMaxState, n = i % kMaxState;
        ckl[i] = (c > 0 && n < N) ? a.x[(((long)b * a.dim + d0 + r) * a.nchunks + (c - 1)) * N + n] : 0.f;
      }
      __syncthreads();
    }
    const long t0 = (long)c * TL + sl * M;

    float uu[M], dl[M], dy[M], dlu[M], sB[M], ddl[M];
    float yv[kHasZ ? M : 1], dzf[kHasZ ? M : 1];
    lds_seg<T, M, ROWB>(rows, rloc, segoff, pre, uu);
    lds_seg<T, M, ROWB>(rows + ROWTILE, rloc, segoff, pre, dl);
    lds_seg<T, M, ROWB>(rows + 2 * ROWTILE, rloc, segoff, pre, dy);
    if constexpr (kHasZ) {
      float zz[M];
      lds_seg<T, M, ROWB>(rows + 3 * ROWTILE, rloc, segoff, pre, zz);
#pragma unroll
      for (int i = 0; i < M; ++i) {
        const float sg = sigmoid_f(zz[i]);
        dzf[i] = dy[i] * sg * (1.f + zz[i] * (1.f - sg));  // dout * d silu(z)/dz
        dy[i] = dy[i] * zz[i] * sg;                         // dout * silu(z)
        yv[i] = Dv * uu[i];
      }
    }
    static_assert(LPR >= kMaxState, "each row's lanes copy that row's carry-in");
    if (sl < kMaxState) sm_hc[rloc * kMaxState + sl] = c > 0 ? ckl[rloc * kMaxState + sl] : 0.f;
    __syncwarp();

    float dlsum = 0.f;
#pragma unroll
    for (int i = 0; i < M; ++i) {
      float x = dl[i] + bias;
      if (a.softplus) x = softplus_f(x);
      if (t0 + i >= a.L) {
        x = 0.f;
        dy[i] = 0.f;
      }
      dl[i] = x;
      dlsum += x;
      dlu[i] = x * uu[i];
      sB[i] = 0.f;
      ddl[i] = 0.f;
    }
    // dl of the step right after this lane's segment (next lane's first step, or the first step
    // of the later chunk for the last lane): a_{t+1} of the reverse recurrence at the segment end
    float dlnext = __shfl_down_sync(0xffffffffu, dl[0], 1, LPR);
    if (sl == LPR - 1) dlnext = sm_dln[rloc];
    const float qsum = dlsum - dl[0] + dlnext;  // sum of dl over (segment shifted by one step)

    if (kTMA) {
      // row values are in registers, the carry-in is in sm_hc: the last warp here refills the row
      // tiles (no CTA-wide barrier in the chunk loop)
      if (release_buffer<WARPS>(&cnts[0], lane, tid) && it + 1 < a.nchunks) issue_rows(c - 1);
      mbar_wait(&bars[1 + s], (it >> 1) & 1);
    }
    const uint8_t* tB = bcs + s * Cfg::BC_TX;
    const uint8_t* tC = tB + BCTILE;
    // slab reduction work item of this thread: one float2 (two time steps) of dB or dC
    constexpr int NV = TL / 2;
    constexpr bool kOneItem = NT >= 2 * NV;  // 256-thread CTAs: exactly one item per thread
    const bool red_active = tid < 2 * NV;
    const int red_arr = tid / NV, red_t2 = tid - red_arr * NV;
    const long red_tg = (long)c * TL + red_t2 * 2;
    const uint8_t* red_slab = (red_arr ? slabC : slabB) + swz128((uint32_t)red_t2 * 8u);
    float* red_dst = (red_arr ? a.dC : a.dB) + (((long)b * a.ngroups + g) * N) * a.L + red_tg;
    const bool red_vec = a.vec_grad && red_tg + 2 <= a.L;

    auto reduce_item = [&](const uint8_t* slab, float* dst, long tg, bool vec) {
      float2 acc = f2(0.f, 0.f);
#pragma unroll
      for (int r = 0; r < R; ++r) acc = __fadd2_rn(acc, *reinterpret_cast<const float2*>(slab + r * SLROW));
      if (vec) {
        if (bpg == 1)
          *reinterpret_cast<float2*>(dst) = acc;
        else
          atomicAdd(reinterpret_cast<float2*>(dst), acc);
      } else {
        if (tg < a.L) {
          if (bpg == 1) dst[0] = acc.x; else atomicAdd(dst, acc.x);
        }
        if (tg + 1 < a.L) {
          if (bpg == 1) dst[1] = acc.y; else atomicAdd(dst + 1, acc.y);
        }
      }
    };
    auto reduce_all = [&](int n) {  // small CTAs: several items per thread
      for (int k = tid; k < 2 * NV; k += NT) {
        const int arr = k / NV, t2 = k - arr * NV;
        const long tg = (long)c * TL + t2 * 2;
        reduce_item((arr ? slabC : slabB) + swz128((uint32_t)t2 * 8u),
                    (arr ? a.dC : a.dB) + (((long)b * a.ngroups + g) * N + n) * a.L + tg, tg,
                    a.vec_grad && tg + 2 <= a.L);
      }
    };

    NZ_UNROLL(NZ_BWD_UNROLL)
    for (int n = 0; n < N; ++n) {
      const float A2 = sm_A2[rloc * kMaxState + n];
      const float An = A2 * kLn2;
      const float hc = sm_hc[rloc * kMaxState + n];
      const float dhc = sm_dhc[rloc * kMaxState + n];
      float av[M], bv[M], cv[M], hh[M];
      [[maybe_unused]] float Bk[M];
      lds_seg<T, M, ROWB>(tB, n, segoff, pre, bv);
      lds_seg<T, M, ROWB>(tC, n, segoff, pre, cv);
      float P = ex2_approx(A2 * dlsum);    // prod a over the segment
      float Q = ex2_approx(A2 * qsum);     // prod a over the segment shifted by one step
      const float anl = ex2_approx(A2 * dlnext);
      float cdy[M];
#pragma unroll
      for (int k = 0; k < H2; ++k) {
        const float2 x2 = mul2(f2(dl[2 * k], dl[2 * k + 1]), f2(A2, A2));
        av[2 * k] = ex2_approx(x2.x);
        av[2 * k + 1] = ex2_approx(x2.y);
        const float2 b2 = mul2(f2(dlu[2 * k], dlu[2 * k + 1]), f2(bv[2 * k], bv[2 * k + 1]));
        hh[2 * k] = b2.x;  // hh holds b_t until the replay overwrites it with h_t
        hh[2 * k + 1] = b2.y;
#if NZ_BWD_KEEPB
        Bk[2 * k] = bv[2 * k];
        Bk[2 * k + 1] = bv[2 * k + 1];
#endif
        const float2 c2 = mul2(f2(cv[2 * k], cv[2 * k + 1]), f2(dy[2 * k], dy[2 * k + 1]));
        cdy[2 * k] = c2.x;
        cdy[2 * k + 1] = c2.y;
      }
      // ---- both folds (independent chains): forward for h, reverse for dh ----
      float H = 0.f, G = 0.f;
#pragma unroll
      for (int i = 0; i < M; ++i) {
        H = fmaf(av[i], H, hh[i]);
        const int j = M - 1 - i;
        G = fmaf(j == M - 1 ? anl : av[j + 1], G, cdy[j]);
      }
      // ---- both Kogge-Stone scans across the lanes of the row, interleaved ----
#pragma unroll
      for (int off = 1; off < LPR; off <<= 1) {
        ks_up<LPR>(P, H, off);
        ks_down<LPR>(Q, G, off);
      }
      float h = ks_enter_up<LPR>(P, H, hc);
      float dh = ks_enter_down<LPR>(Q, G, dhc);
      const float dhnew = fmaf(Q, dhc, G);  // dh leaving this lane's segment towards earlier steps
#if NZ_BWD_SPLITBAR
      if (n > 0) {  // reduce the previous state's slab while this state's scans are in flight
        mbar_wait(&bars[3], (gtrip - 1) & 1);
        if constexpr (kOneItem) {
          if (red_active) reduce_item(red_slab, red_dst + (long)(n - 1) * a.L, red_tg, red_vec);
        } else {
          reduce_all(n - 1);
        }
        mbar_arrive(&bars[4]);
      }
#endif
      // ---- replay both recurrences with the true carries ----
      float dd[M];
#pragma unroll
      for (int i = 0; i < M; ++i) {
        h = fmaf(av[i], h, hh[i]);  // h_t
        const float bi = hh[i];
        hh[i] = h;
        bv[i] = bi;                 // keep b_t: a_t h_{t-1} = h_t - b_t
        const int j = M - 1 - i;
        dh = fmaf(j == M - 1 ? anl : av[j + 1], dh, cdy[j]);
        dd[j] = dh;
      }
      // ---- element-wise products, packed over time pairs ----
      float2 gs2 = f2(0.f, 0.f);
      float vC[M], vB[M];
#pragma unroll
      for (int k = 0; k < H2; ++k) {
        const float2 h2 = f2(hh[2 * k], hh[2 * k + 1]);
        const float2 d2 = f2(dd[2 * k], dd[2 * k + 1]);
        const float2 dy2 = f2(dy[2 * k], dy[2 * k + 1]);
        const float2 dlu2 = f2(dlu[2 * k], dlu[2 * k + 1]);
        const float2 c2 = mul2(dy2, h2);  // dC_t[n] contribution of this row
        vC[2 * k] = c2.x;
        vC[2 * k + 1] = c2.y;
        const float2 ah2 = sub2(h2, f2(bv[2 * k], bv[2 * k + 1]));  // a_t h_{t-1}
        const float2 gq2 = mul2(d2, ah2);
        float2 ddl2 = f2(ddl[2 * k], ddl[2 * k + 1]);
        ddl2 = fma2(f2(An, An), gq2, ddl2);
        ddl[2 * k] = ddl2.x;
        ddl[2 * k + 1] = ddl2.y;
        gs2 = fma2(f2(dl[2 * k], dl[2 * k + 1]), gq2, gs2);
        const float2 b2 = mul2(d2, dlu2);  // dB_t[n] contribution of this row
        vB[2 * k] = b2.x;
        vB[2 * k + 1] = b2.y;
      }
      {
#if NZ_BWD_KEEPB
        const float (&Bv)[M] = Bk;
#else
        float Bv[M];
        lds_seg<T, M, ROWB>(tB, n, segoff, pre, Bv);
#endif
#pragma unroll
        for (int k = 0; k < H2; ++k) {
          float2 s2 = f2(sB[2 * k], sB[2 * k + 1]);
          s2 = fma2(f2(dd[2 * k], dd[2 * k + 1]), f2(Bv[2 * k], Bv[2 * k + 1]), s2);
          sB[2 * k] = s2.x;
          sB[2 * k + 1] = s2.y;
        }
        if constexpr (kHasZ) {
          float Cv[M];
          lds_seg<T, M, ROWB>(tC, n, segoff, pre, Cv);
#pragma unroll
          for (int k = 0; k < H2; ++k) {
            float2 y2 = f2(yv[2 * k], yv[2 * k + 1]);
            y2 = fma2(f2(Cv[2 * k], Cv[2 * k + 1]), f2(hh[2 * k], hh[2 * k + 1]), y2);
            yv[2 * k] = y2.x;
            yv[2 * k + 1] = y2.y;
          }
        }
      }
      float gs = gs2.x + gs2.y;
#pragma unroll
      for (int off = LPR / 2; off > 0; off >>= 1) gs += __shfl_xor_sync(0xffffffffu, gs, off, LPR);

#if NZ_BWD_SPLITBAR
      // Hand-off without CTA-wide stalls: slab(n) is reduced by everybody during state n+1 (after
      // its scans), so both waits sit far behind the matching arrives.
      if (gtrip > 0) mbar_wait(&bars[4], (gtrip - 1) & 1);  // slab(n-1) has been consumed
#else
      __syncthreads();  // the previous state's slab reduction is finished: slabs may be rewritten
#endif
#pragma unroll
      for (int j = 0; j < M / 4; ++j) {
        *reinterpret_cast<float4*>(slabC + swz128(slab_off + 16 * j)) =
            make_float4(vC[4 * j], vC[4 * j + 1], vC[4 * j + 2], vC[4 * j + 3]);
        *reinterpret_cast<float4*>(slabB + swz128(slab_off + 16 * j)) =
            make_float4(vB[4 * j], vB[4 * j + 1], vB[4 * j + 2], vB[4 * j + 3]);
      }
      if (sl == 0) {
        sm_dA[rloc * kMaxState + n] += gs;
        sm_dhc[rloc * kMaxState + n] = dhnew;
      }
#if NZ_BWD_SPLITBAR
      mbar_arrive(&bars[3]);  // my part of slab(n) is written
      ++gtrip;
    }
    // the last state's slab of this chunk
    mbar_wait(&bars[3], (gtrip - 1) & 1);
    if constexpr (kOneItem) {
      if (red_active) reduce_item(red_slab, red_dst + (long)(N - 1) * a.L, red_tg, red_vec);
    } else {
      reduce_all(N - 1);
    }
    mbar_arrive(&bars[4]);
#else
      __syncthreads();  // slabs complete
      if constexpr (kOneItem) {
        if (red_active) reduce_item(red_slab, red_dst + (long)n * a.L, red_tg, red_vec);
      } else {
        reduce_all(n);
      }
    }
#endif

    // ---- per-(row, t) epilogue ----
    if (sl == 0) sm_dln[rloc] = dl[0];  // first dl of this chunk: read by the next (earlier) chunk
    float outv[M];
#pragma unroll
    for (int i = 0; i < M; ++i) outv[i] = fmaf(dl[i], sB[i], Dv * dy[i]);  // du
    stg_items<T, M>(durow, outv, t0, a.L, a.vec_grad != 0);
#pragma unroll
    for (int i = 0; i < M; ++i) {
      float gd = fmaf(uu[i], sB[i], ddl[i]);  // d loss / d dl
      // d softplus(x)/dx = sigmoid(x) = 1 - exp(-softplus(x))
      if (a.softplus) gd *= sigmoid_from_softplus(dl[i]);
      if (t0 + i >= a.L) gd = 0.f;
      outv[i] = gd;
      db_acc += gd;
      dD_acc = fmaf(dy[i], uu[i], dD_acc);
    }
    stg_items<T, M>(ddrow, outv, t0, a.L, a.vec_grad != 0);
    if constexpr (kHasZ) {
#pragma unroll
      for (int i = 0; i < M; ++i) outv[i] = dzf[i] * yv[i];
      stg_items<T, M>(dzrow, outv, t0, a.L, a.vec_grad != 0);
    }

    if (kTMA) {
      // last warp done with B/C stage s refills it for the chunk after next
      if (release_buffer<WARPS>(&cnts[1 + s], lane, tid) && it + 2 < a.nchunks) issue_bc(c - 2, s);
    }
  }

  // ---- (dim)-shaped reductions: over lanes here, over the batch with fp32 atomics ----
#pragma unroll
  for (int off = LPR / 2; off > 0; off >>= 1) {
    dD_acc += __shfl_xor_sync(0xffffffffu, dD_acc, off, LPR);
    db_acc += __shfl_xor_sync(0xffffffffu, db_acc, off, LPR);
  }
  if (sl == 0) {
    if (a.dD) atomicAdd(a.dD + d, dD_acc);
    if (a.dbias) atomicAdd(a.dbias + d, db_acc);
  }
  __syncthreads();
  for (int i = tid; i < R * kMaxState; i += NT) {
    const int r = i / kMaxState, n = i % kMaxState;
    if (n < N) atomicAdd(a.dA + (long)(d0 + r) * N + n, sm_dA[i]);
  }
}

// Launch wrappers instantiated per element type in scan_inst_*.cu
template <typename T>
cudaError_t launch_scan_fwd(const ScanKArgs& a, bool tma, bool has_z, int rows_per_cta, cudaStream_t stream);
template <typename T>
cudaError_t launch_scan_bwd(const ScanKArgs& a, bool tma, bool has_z, int rows_per_cta, cudaStream_t stream);

}  // namespace nz
