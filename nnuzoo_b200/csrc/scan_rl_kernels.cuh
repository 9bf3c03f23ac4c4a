// scan_rl_kernels.cuh -- "row-lane" selective-scan backward for sm_100a (round 2).
//
// Same math as scan_bwd_kernel (reference: nnunetv2/nets/seg_mamba/selective_scan_interface.py:50-74, the adjoint of
// :86-152), different decomposition.  The warp-scan backward of scan_kernels.cuh pays, per state, two Kogge-Stone scans
// over shuffles, a fold + replay of both recurrences and a CTA-wide slab hand-off for the dB / dC row reduction; ncu
// showed it bound by issue slots and the shared-memory pipe together at 25 % occupancy (profiles/r01_kernel_tuning.md).
// Here
//   * a LANE owns one row (b, d); a warp owns 32 consecutive rows of one (batch, group) and is an independent worker
//     (a CTA is one warp: no __syncthreads anywhere);
//   * time is walked sequentially in blocks of 8 steps, last block first.  The forward saved h at the end of every
//     block ("fine checkpoints", NzScanDesc::xf), so inside a block a lane runs ONE forward recurrence for h and ONE
//     reverse recurrence for dh per state: no intra-warp scan, no fold / replay, no shuffles in the recurrences;
//   * B_t[n] / C_t[n] reads are warp-wide broadcasts (every lane reads the same address);
//   * dB_t[n] / dC_t[n] = sums over the rows of the group: the 32 rows of the warp are reduced through a warp-private
//     slab (4 STS.128 per lane, 8 LDS.64 + adds + 2 shuffle rounds), one 32-byte store (or RED when the group has more
//     than 32 rows) per (state, block, array);
//   * parallelism along L comes from CHUNKS: an aggregate pass (scan_bwd_rl_agg_kernel, same lane mapping, all 16
//     states in registers, time-outer) computes for every chunk the pair (Q, G) of the reverse recurrence
//         R_in(c) = G(c+1) + Q(c+1) * R_in(c+1),      R := a_{t+1} * dh_{t+1} entering the chunk's last step,
//     a tiny combine kernel walks the chunks, and the main pass then runs every (row block, chunk) independently.
//     This is the chunk-parallel backward for few-rows / long-L launches as well (BASELINE configs[2]).
//
// Recurrence bookkeeping (R-form):  dh_t = C_t dy_t + R_t,  R_{t-1} = a_t dh_t.  A block keeps a_t of its 8 steps in
// registers, so inside a block dh_j = fma(a_{j+1}, dh_{j+1}, cdy_j); across blocks only R (one float per state) is
// carried, in registers (the state loop is fully unrolled).
#pragma once

#include "../../include/nnuzoo_b200.h"
#include "nz_common.cuh"
#include "scan_kernels.cuh"  // packed fp32x2 helpers, lds/sts helpers

namespace nz {

constexpr int kFine = NZ_FINE;  // steps per block = fine checkpoint interval
static_assert(kFine == 8, "the row-lane kernels are written for 8-step blocks");

struct alignas(64) RlArgs {
  CUtensorMap tm_u, tm_delta, tm_dout, tm_z, tm_B, tm_C;
  const float *A, *D, *bias;
  const float* xf;     // (batch, dim, L/8, 16): h at the end of every 8-step block
  void *du, *ddelta, *dz;
  float *dA, *dB, *dC, *dD, *dbias;
  float *aggG, *aggQ;  // [row][chunk][16] aggregates of the reverse recurrence (aggregate pass out)
  float* Rin;          // [row][chunk][16] R entering every chunk (combine out, main pass in)
  long L;
  long A_ds;
  int batch, dim, ngroups, dpg;
  int nrb;       // 32-row blocks per group
  int ntl;       // tiles along L (tile = 128 bytes of a row)
  int tpc;       // tiles per chunk
  int nchunks;   // chunks along L
  int softplus;
  int single;    // one warp owns each dB / dC element (dpg == 32): plain stores instead of RED
};

template <typename T>
struct RlCfg {
  static constexpr int ES = sizeof(T);
  static constexpr int TB = 128 / ES;         // steps per tile (one 128-byte line of a row)
  static constexpr int NBLK = TB / kFine;     // blocks per tile (4 fp32, 8 16-bit)
  static constexpr int VPB = kFine * ES / 16; // 16-byte vectors per block of a row (2 fp32, 1 16-bit)
  static constexpr int ROWT = 32 * 128;       // one row tile: 32 rows x 128 bytes
  static constexpr int BCT = kMaxState * 128; // one B or C tile
};

// ---- small helpers ----
__device__ __forceinline__ void stg64_or_red(float* p, float x, float y, bool single) {
  if (single) {
    *reinterpret_cast<float2*>(p) = make_float2(x, y);
  } else {
    asm volatile("red.relaxed.gpu.global.add.v2.f32 [%0], {%1, %2};" ::"l"(p), "f"(x), "f"(y) : "memory");
  }
}
__device__ __forceinline__ float2 lds64(uint32_t a) {
  float2 v;
  asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(v.x), "=f"(v.y) : "r"(a));
  return v;
}
// a lane's 8 items of block `blk` out of a swizzled [rows][128 B] tile; row = tile row (lane for row tiles, the state
// for B / C tiles -- then every lane reads the same address: a broadcast)
template <typename T>
__device__ __forceinline__ void lds_block(uint32_t tile_s, int row, int blk, float (&v)[8]) {
  constexpr int VPB = RlCfg<T>::VPB;
  constexpr int kPer = 16 / (int)sizeof(T);
  const uint32_t base = tile_s + (uint32_t)row * 128u;
  const uint32_t key = (uint32_t)(row & 7);
#pragma unroll
  for (int j = 0; j < VPB; ++j) {
    const uint32_t chunk = ((uint32_t)(blk * VPB + j)) ^ key;
    unpack16<T>(lds128(base + (chunk << 4)), &v[j * kPer]);
  }
}

// ================================================================================================
// Aggregate pass: (Q, G) of the reverse recurrence per (row, chunk, state)
// ================================================================================================
template <typename T, bool kHasZ>
__global__ void __launch_bounds__(32, 10) scan_bwd_rl_agg_kernel(const __grid_constant__ RlArgs a) {
  using Cfg = RlCfg<T>;
  constexpr int TB = Cfg::TB, NBLK = Cfg::NBLK, ROWT = Cfg::ROWT, BCT = Cfg::BCT;
  constexpr int NROW = kHasZ ? 3 : 2;               // delta, dout, [z]
  constexpr int STAGE = NROW * ROWT + BCT;          // bytes per pipeline stage
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + 2 * STAGE);
  const uint32_t smem_s = smem_u32(smem);
  const int lane = threadIdx.x;

  // work item: chunks 1 .. nchunks-1 (nobody needs the aggregate of the first chunk in time)
  const int nc1 = a.nchunks - 1;
  const int item = blockIdx.x;
  const int c = item % nc1 + 1;
  int w = item / nc1;
  const int rb = w % a.nrb;
  w /= a.nrb;
  const int g = w % a.ngroups, b = w / a.ngroups;
  const int d0 = g * a.dpg + rb * 32, d = d0 + lane;
  const long rowg = (long)b * a.dim + d;
  const int t_lo = c * a.tpc, t_hi = min(a.ntl, t_lo + a.tpc);

  auto issue = [&](int t, int s) {
    uint8_t* st = smem + s * STAGE;
    mbar_arrive_expect_tx(&bars[s], STAGE);
    tma_load_4d(st, &a.tm_delta, &bars[s], 0, t, d0, b);
    tma_load_4d(st + ROWT, &a.tm_dout, &bars[s], 0, t, d0, b);
    if (kHasZ) tma_load_4d(st + 2 * ROWT, &a.tm_z, &bars[s], 0, t, d0, b);
    tma_load_5d(st + NROW * ROWT, &a.tm_C, &bars[s], 0, t, 0, g, b);
  };
  if (lane == 0) {
    mbar_init(&bars[0], 1);
    mbar_init(&bars[1], 1);
    fence_mbar_init();
    issue(t_hi - 1, 0);
    if (t_hi - 2 >= t_lo) issue(t_hi - 2, 1);
  }
  __syncwarp();

  float A2[kMaxState], R[kMaxState];
#pragma unroll
  for (int n = 0; n < kMaxState; ++n) {
    A2[n] = __ldg(a.A + (long)d * a.A_ds + n) * kLog2e;
    R[n] = 0.f;
  }
  const float bias = a.bias ? __ldg(a.bias + d) : 0.f;
  float dlsum = 0.f;

  int k = 0;
  for (int t = t_hi - 1; t >= t_lo; --t, ++k) {
    const int s = k & 1;
    mbar_wait(&bars[s], (k >> 1) & 1);
    const uint32_t st = smem_s + s * STAGE;
#pragma unroll 1
    for (int blk = NBLK - 1; blk >= 0; --blk) {
      float dl[8], dy[8];
      lds_block<T>(st, lane, blk, dl);
      lds_block<T>(st + ROWT, lane, blk, dy);
      if constexpr (kHasZ) {
        float zz[8];
        lds_block<T>(st + 2 * ROWT, lane, blk, zz);
#pragma unroll
        for (int i = 0; i < 8; ++i) dy[i] = dy[i] * zz[i] * sigmoid_f(zz[i]);  // dout * silu(z)
      }
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        float x = dl[i] + bias;
        if (a.softplus) x = softplus_f(x);
        dl[i] = x;
        dlsum += x;
      }
#pragma unroll
      for (int n = 0; n < kMaxState; ++n) {
        float cv[8];
        lds_block<T>(st + NROW * ROWT, n, blk, cv);
        float r = R[n];
#pragma unroll
        for (int i = 7; i >= 0; --i) {
          const float av = ex2_approx(A2[n] * dl[i]);
          r = av * fmaf(cv[i], dy[i], r);  // R_{t-1} = a_t (C_t dy_t + R_t)
        }
        R[n] = r;
      }
    }
    __syncwarp();  // every lane is done with stage s
    if (lane == 0 && t - 2 >= t_lo) issue(t - 2, s);
  }
  float* G = a.aggG + (rowg * a.nchunks + c) * kMaxState;
  float* Q = a.aggQ + (rowg * a.nchunks + c) * kMaxState;
#pragma unroll
  for (int n = 0; n < kMaxState; n += 4) {
    *reinterpret_cast<float4*>(G + n) = make_float4(R[n], R[n + 1], R[n + 2], R[n + 3]);
    *reinterpret_cast<float4*>(Q + n) = make_float4(ex2_approx(A2[n] * dlsum), ex2_approx(A2[n + 1] * dlsum),
                                                    ex2_approx(A2[n + 2] * dlsum), ex2_approx(A2[n + 3] * dlsum));
  }
}

// R entering every chunk: one thread per (row, state) walks the chunks last to first
static __global__ void __launch_bounds__(128) scan_bwd_rl_combine_kernel(const float* __restrict__ G, const float* __restrict__ Q,
                                                                 float* __restrict__ Rin, long nrows, int nchunks) {
  const long i = blockIdx.x * 128L + threadIdx.x;
  if (i >= nrows * kMaxState) return;
  const long row = i / kMaxState;
  const int n = (int)(i % kMaxState);
  const long base = row * nchunks * kMaxState + n;
  float r = 0.f;
  Rin[base + (long)(nchunks - 1) * kMaxState] = 0.f;
  for (int c = nchunks - 2; c >= 0; --c) {
    r = fmaf(Q[base + (long)(c + 1) * kMaxState], r, G[base + (long)(c + 1) * kMaxState]);
    Rin[base + (long)c * kMaxState] = r;
  }
}

// ================================================================================================
// Main pass
// ================================================================================================
// Shared memory of one warp (bytes):
//   rows   NROWT x 4096   u, delta, dout, [z] tiles (single-buffered; the next tile is requested as soon as the last
//                         block of this one sits in registers)
//   bc     2 x 2 x 2048   B, C tiles, double-buffered
//   slab   2 x 1152       dB / dC products of one (state, block): 4 groups of 8 rows, 288-byte group pitch
// A warp is one CTA and shared memory caps residency at 8-9 warps per SM, so registers are free (up to 255): the state
// loop is fully unrolled and A, R (the carried reverse state) and the dA partial sums of all 16 states live in registers.
template <typename T, bool kHasZ>
struct RlMainSmem {
  static constexpr int NROWT = kHasZ ? 4 : 3;
  static constexpr int ROWS = NROWT * RlCfg<T>::ROWT;
  static constexpr int BC = 2 * 2 * RlCfg<T>::BCT;
  static constexpr int SLAB1 = 4 * 288;
  static constexpr int SLAB = 2 * SLAB1;
  static constexpr int OFF_BC = ROWS, OFF_SLAB = OFF_BC + BC, OFF_BARS = OFF_SLAB + ((SLAB + 127) / 128) * 128;
  static constexpr int TOTAL = OFF_BARS + 64;
  static constexpr size_t bytes() { return 1024 + TOTAL; }
};

template <typename T, bool kHasZ>
__global__ void __launch_bounds__(32, 8) scan_bwd_rl_kernel(const __grid_constant__ RlArgs a) {
  using Cfg = RlCfg<T>;
  using SM = RlMainSmem<T, kHasZ>;
  constexpr int NBLK = Cfg::NBLK, ROWT = Cfg::ROWT, BCT = Cfg::BCT;
  constexpr int ROWS_TX = SM::NROWT * ROWT;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + SM::OFF_BARS);  // [0] rows, [1],[2] B/C stages
  const uint32_t smem_s = keep(smem_u32(smem));
  const uint32_t bc_s = smem_s + SM::OFF_BC, slab_s = smem_s + SM::OFF_SLAB;
  const int lane = threadIdx.x;

  const int item = blockIdx.x;
  const int c = item % a.nchunks;
  int w = item / a.nchunks;
  const int rb = w % a.nrb;
  w /= a.nrb;
  const int g = w % a.ngroups, b = w / a.ngroups;
  const int d0 = g * a.dpg + rb * 32, d = d0 + lane;
  const long rowg = (long)b * a.dim + d;
  const int t_lo = c * a.tpc, t_hi = min(a.ntl, t_lo + a.tpc);
  const long nbt = a.L / kFine;  // fine blocks per row

  auto issue_rows = [&](int t) {
    mbar_arrive_expect_tx(&bars[0], ROWS_TX);
    tma_load_4d(smem, &a.tm_u, &bars[0], 0, t, d0, b);
    tma_load_4d(smem + ROWT, &a.tm_delta, &bars[0], 0, t, d0, b);
    tma_load_4d(smem + 2 * ROWT, &a.tm_dout, &bars[0], 0, t, d0, b);
    if (kHasZ) tma_load_4d(smem + 3 * ROWT, &a.tm_z, &bars[0], 0, t, d0, b);
  };
  auto issue_bc = [&](int t, int s) {
    uint8_t* st = smem + SM::OFF_BC + s * 2 * BCT;
    mbar_arrive_expect_tx(&bars[1 + s], 2 * BCT);
    tma_load_5d(st, &a.tm_B, &bars[1 + s], 0, t, 0, g, b);
    tma_load_5d(st + BCT, &a.tm_C, &bars[1 + s], 0, t, 0, g, b);
  };
  if (lane == 0) {
    mbar_init(&bars[0], 1);
    mbar_init(&bars[1], 1);
    mbar_init(&bars[2], 1);
    fence_mbar_init();
    issue_rows(t_hi - 1);
    issue_bc(t_hi - 1, 0);
    if (t_hi - 2 >= t_lo) issue_bc(t_hi - 2, 1);
  }
  __syncwarp();
  float A2[kMaxState], R[kMaxState], dAacc[kMaxState];
  {
    const float* rin = a.Rin + (rowg * a.nchunks + c) * kMaxState;
#pragma unroll
    for (int n = 0; n < kMaxState; ++n) {
      A2[n] = __ldg(a.A + (long)d * a.A_ds + n) * kLog2e;
      R[n] = a.nchunks > 1 ? __ldg(rin + n) : 0.f;  // a_{t+1} dh_{t+1} entering the chunk's last step
      dAacc[n] = 0.f;
    }
  }
  const float Dv = a.D ? __ldg(a.D + d) : 0.f;
  const float bias = a.bias ? __ldg(a.bias + d) : 0.f;
  float dD_acc = 0.f, db_acc = 0.f;

  // slab addressing.  Writer: row = lane -> group q = lane >> 3, i = lane & 7; the two 16-byte halves of the row's 8
  // floats swap places for i >= 4 (a quarter warp then covers all 32 banks).  Reader: lane -> (array, step pair p,
  // row group qq); rows i = 0..7 of the group sit 32 bytes apart, the halves swapped for i >= 4.
  const uint32_t slab_w = keep(slab_s + (uint32_t)((lane >> 3) * 288 + (lane & 7) * 32 + (((lane >> 2) & 1) << 4)));
  const int r_arr = lane >> 4, r_p = (lane >> 2) & 3, r_q = lane & 3;
  const uint32_t slab_r0 = keep(slab_s + (uint32_t)(r_arr * SM::SLAB1 + r_q * 288 + r_p * 8));
  const uint32_t slab_r1 = keep(slab_s + (uint32_t)(r_arr * SM::SLAB1 + r_q * 288 + ((r_p * 8) ^ 16)));
  float* dG0 = (r_arr ? a.dC : a.dB) + ((long)b * a.ngroups + g) * kMaxState * a.L + r_p * 2;
  const bool single = a.single != 0;

  int k = 0;
  for (int t = t_hi - 1; t >= t_lo; --t, ++k) {
    const int s = k & 1;
    mbar_wait(&bars[0], k & 1);
    mbar_wait(&bars[1 + s], (k >> 1) & 1);
    const uint32_t tB = bc_s + s * 2 * BCT, tC = tB + BCT;
#pragma unroll 1
    for (int blk = NBLK - 1; blk >= 0; --blk) {
      const long jb = (long)t * NBLK + blk;  // fine block index along L
      float dl[8], dlu[8], dy[8], uu[8], sB[8], ddl[8];
      float yv[kHasZ ? 8 : 1], dzf[kHasZ ? 8 : 1];
      lds_block<T>(smem_s, lane, blk, uu);
      lds_block<T>(smem_s + ROWT, lane, blk, dl);
      lds_block<T>(smem_s + 2 * ROWT, lane, blk, dy);
      if constexpr (kHasZ) {
        float zz[8];
        lds_block<T>(smem_s + 3 * ROWT, lane, blk, zz);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const float sg = sigmoid_f(zz[i]);
          dzf[i] = dy[i] * sg * (1.f + zz[i] * (1.f - sg));  // dout * d silu(z)/dz
          dy[i] = dy[i] * zz[i] * sg;                         // dout * silu(z)
          yv[i] = Dv * uu[i];
        }
      }
      // h entering the block: the forward's fine checkpoint of the previous block (zero at the sequence start)
      float4 hq[4];
      {
        const float4* xin = reinterpret_cast<const float4*>(a.xf + (rowg * nbt + (jb - 1)) * kMaxState);
#pragma unroll
        for (int q = 0; q < 4; ++q) hq[q] = jb > 0 ? __ldg(xin + q) : make_float4(0.f, 0.f, 0.f, 0.f);
      }
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        float x = dl[i] + bias;
        if (a.softplus) x = softplus_f(x);
        dl[i] = x;
        dlu[i] = x * uu[i];
        sB[i] = 0.f;
        ddl[i] = 0.f;
      }
      if (blk == 0) {  // the row tiles now live in registers: request the next tile
        __syncwarp();
        if (lane == 0 && t - 1 >= t_lo) issue_rows(t - 1);
      }
      float* dG = dG0 + jb * kFine;

#pragma unroll
      for (int n = 0; n < kMaxState; ++n) {
        const float An = A2[n] * kLn2;
        const float4 h4 = hq[n >> 2];
        const float hin = (n & 3) == 0 ? h4.x : (n & 3) == 1 ? h4.y : (n & 3) == 2 ? h4.z : h4.w;
        float av[8], bv[8], hh[8], cdy[8], dd[8];
        lds_block<T>(tB, n, blk, bv);
        lds_block<T>(tC, n, blk, cdy);
        [[maybe_unused]] float cz[kHasZ ? 8 : 1];
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) {
          const float2 x2 = mul2(f2(dl[2 * kk], dl[2 * kk + 1]), f2(A2[n], A2[n]));
          av[2 * kk] = ex2_approx(x2.x);
          av[2 * kk + 1] = ex2_approx(x2.y);
          const float2 b2 = mul2(f2(dlu[2 * kk], dlu[2 * kk + 1]), f2(bv[2 * kk], bv[2 * kk + 1]));
          hh[2 * kk] = b2.x;  // b_t until the recurrence overwrites it with h_t
          hh[2 * kk + 1] = b2.y;
          if constexpr (kHasZ) {
            cz[2 * kk] = cdy[2 * kk];
            cz[2 * kk + 1] = cdy[2 * kk + 1];
          }
          const float2 c2 = mul2(f2(cdy[2 * kk], cdy[2 * kk + 1]), f2(dy[2 * kk], dy[2 * kk + 1]));
          cdy[2 * kk] = c2.x;
          cdy[2 * kk + 1] = c2.y;
        }
        // forward recurrence for h, reverse recurrence for dh (independent chains)
        float bsave[8];
        float h = hin;
        float dh = cdy[7] + R[n];
        dd[7] = dh;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          bsave[i] = hh[i];
          h = fmaf(av[i], h, hh[i]);
          hh[i] = h;
          if (i < 7) {
            const int j = 6 - i;
            dh = fmaf(av[j + 1], dh, cdy[j]);
            dd[j] = dh;
          }
        }
        R[n] = av[0] * dd[0];  // R leaving the block
        // element-wise products, packed over time pairs
        float2 gs2 = f2(0.f, 0.f);
        float vC[8], vB[8];
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) {
          const float2 h2 = f2(hh[2 * kk], hh[2 * kk + 1]);
          const float2 d2 = f2(dd[2 * kk], dd[2 * kk + 1]);
          const float2 c2 = mul2(f2(dy[2 * kk], dy[2 * kk + 1]), h2);  // dC_t[n] of this row
          vC[2 * kk] = c2.x;
          vC[2 * kk + 1] = c2.y;
          const float2 ah2 = sub2(h2, f2(bsave[2 * kk], bsave[2 * kk + 1]));  // a_t h_{t-1}
          const float2 gq2 = mul2(d2, ah2);
          float2 ddl2 = f2(ddl[2 * kk], ddl[2 * kk + 1]);
          ddl2 = fma2(f2(An, An), gq2, ddl2);
          ddl[2 * kk] = ddl2.x;
          ddl[2 * kk + 1] = ddl2.y;
          gs2 = fma2(f2(dl[2 * kk], dl[2 * kk + 1]), gq2, gs2);
          const float2 b2 = mul2(d2, f2(dlu[2 * kk], dlu[2 * kk + 1]));  // dB_t[n] of this row
          vB[2 * kk] = b2.x;
          vB[2 * kk + 1] = b2.y;
          float2 s2 = f2(sB[2 * kk], sB[2 * kk + 1]);
          s2 = fma2(d2, f2(bv[2 * kk], bv[2 * kk + 1]), s2);
          sB[2 * kk] = s2.x;
          sB[2 * kk + 1] = s2.y;
          if constexpr (kHasZ) {
            float2 y2 = f2(yv[2 * kk], yv[2 * kk + 1]);
            y2 = fma2(f2(cz[2 * kk], cz[2 * kk + 1]), h2, y2);
            yv[2 * kk] = y2.x;
            yv[2 * kk + 1] = y2.y;
          }
        }
        dAacc[n] += gs2.x + gs2.y;  // this row's share of dA[n]
        // ---- dB / dC: reduce the warp's 32 rows through the slab ----
        sts128(slab_w, vB[0], vB[1], vB[2], vB[3]);
        sts128(slab_w ^ 16u, vB[4], vB[5], vB[6], vB[7]);
        sts128(slab_w + SM::SLAB1, vC[0], vC[1], vC[2], vC[3]);
        sts128((slab_w ^ 16u) + SM::SLAB1, vC[4], vC[5], vC[6], vC[7]);
        __syncwarp();
        float2 acc = lds64(slab_r0);
#pragma unroll
        for (int i = 1; i < 8; ++i) acc = __fadd2_rn(acc, lds64((i < 4 ? slab_r0 : slab_r1) + i * 32));
        __syncwarp();
        acc.x += __shfl_xor_sync(0xffffffffu, acc.x, 1);
        acc.y += __shfl_xor_sync(0xffffffffu, acc.y, 1);
        acc.x += __shfl_xor_sync(0xffffffffu, acc.x, 2);
        acc.y += __shfl_xor_sync(0xffffffffu, acc.y, 2);
        if (r_q == 0) stg64_or_red(dG + (long)n * a.L, acc.x, acc.y, single);
      }

      // ---- per-(row, t) epilogue of the block ----
      const long tpos = jb * kFine;
      float outv[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) outv[i] = fmaf(dl[i], sB[i], Dv * dy[i]);  // du
      stg_items<T, 8>(reinterpret_cast<T*>(a.du) + rowg * a.L, outv, tpos, a.L, true);
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        float gd = fmaf(uu[i], sB[i], ddl[i]);  // d loss / d dl
        if (a.softplus) gd *= sigmoid_from_softplus(dl[i]);
        outv[i] = gd;
        db_acc += gd;
        dD_acc = fmaf(dy[i], uu[i], dD_acc);
      }
      stg_items<T, 8>(reinterpret_cast<T*>(a.ddelta) + rowg * a.L, outv, tpos, a.L, true);
      if constexpr (kHasZ) {
#pragma unroll
        for (int i = 0; i < 8; ++i) outv[i] = dzf[i] * yv[i];
        stg_items<T, 8>(reinterpret_cast<T*>(a.dz) + rowg * a.L, outv, tpos, a.L, true);
      }
    }
    __syncwarp();  // every lane is done with B/C stage s
    if (lane == 0 && t - 2 >= t_lo) issue_bc(t - 2, s);
  }
  // (dim)-shaped sums: over batch and chunks with fp32 atomics
#pragma unroll
  for (int n = 0; n < kMaxState; ++n) atomicAdd(a.dA + (long)d * kMaxState + n, dAacc[n]);
  if (a.dD) atomicAdd(a.dD + d, dD_acc);
  if (a.dbias) atomicAdd(a.dbias + d, db_acc);
}

template <typename T>
cudaError_t launch_scan_bwd_rl(const RlArgs& a, bool has_z, cudaStream_t stream);

}  // namespace nz
