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

#include <type_traits>

namespace nz {

constexpr int kFine = NZ_FINE;  // steps per block = fine checkpoint interval
static_assert(kFine == 8, "the row-lane kernels are written for 8-step blocks");

struct alignas(64) RlArgs {
  // main pass: per-block boxes (8 steps).  aggregate pass: per-tile boxes (128 bytes of a row): the g_ maps
  CUtensorMap tm_u, tm_delta, tm_dout, tm_z, tm_B, tm_C, tm_xf;
  CUtensorMap g_delta, g_row1, g_z, g_bc;  // row1 / bc: dout and C (backward), u and B (forward)
  const float *A, *D, *bias;
  const float* xf;     // (batch, dim, L/8, 16): h at the end of every 8-step block
  void *du, *ddelta, *dz;
  float *dA, *dB, *dC, *dD, *dbias;
  float *aggG, *aggQ;  // [row][chunk][16] aggregates (G: value leaving the chunk from a zero start, Q: product of a)
  float* Rin;          // [row][chunk][16] state entering every chunk (combine out, main pass in): R for the backward,
                       // h for the forward
  void* out;           // forward: (batch, dim, L) result (T, or fp32 when out_f32)
  float* x;            // forward: coarse checkpoints (batch, dim, nck, 16)
  float* xfw;          // forward: fine checkpoints to write (may be NULL)
  long o_bs, o_ds;     // out strides (elements)
  int nck, out_f32;
  long L;
  long A_ds;
  int batch, dim, ngroups, dpg;
  int nrb;       // 32-row blocks per group
  int ntl;       // tiles along L (tile = 128 bytes of a row)
  int tpc;       // tiles per chunk
  int nchunks;   // chunks along L
  int softplus;
  int single;    // one warp owns each dB / dC element (dpg == 32): plain stores instead of RED
  unsigned zero; // always 0 (see order_after)
  int v2;        // backward main pass: 1 = butterfly version (scan_bwd_rl2_kernel), 0 = slab version
  int rev_mask;  // bit g set: the rows of group g walk the sequence BACKWARDS (kRevCap kernels only): the forward
                 // recurrence runs from t = L-1 down to 0 over the same, un-flipped arrays -- SS2D's flipped
                 // directions without a flipped copy (m2net.py:176)
  int u_gdiv;    // u has dim / u_gdiv rows per batch entry: group g reads the rows of group g / u_gdiv (the two
                 // directions that walk the same array forwards and backwards share it); 1 = plain
  int wide;      // every fp32 output row / checkpoint run is 32-byte aligned: one 32-byte store per block
  int v2f;       // forward main pass: 1 = pipelined version (scan_fwd_rl2_kernel)
  int zero_dbc;  // backward, several row blocks per group (dB / dC are reduced with RED): the aggregate pass zero-fills
                 // dB / dC on its way (its stores are free under the MUFU bound), so the caller need not
};

// ---- exp2 on the FMA pipe ----
// The aggregate passes and the forward main pass are bound by the MUFU pipe (16 ex2 per element at 16 lanes/clk/SM:
// 90 % busy, issue slots half idle).  A share of the states therefore computes a_t = 2^x with FMA-pipe instructions, two
// steps at a time: round-to-nearest split x = i + f (magic-number add), degree-5 polynomial for 2^f on [-0.5, 0.5]
// (max relative error 1.9e-7, the same order as ex2.approx), exponent add by integer arithmetic.  |x| <= 126 is the
// caller's business (the kernels clamp dl once per block so that |A * dl| * log2(e) <= 126; beyond that a_t is 0 to
// fp32 anyway).  MEASURED AND SWITCHED OFF (profiles/r02_kernel_tuning.md): with 5 of 16 states on the polynomial the
// aggregate pass went from 0.454 to 0.49 ms -- MUFU 90 -> 61 %, FMA 34 -> 53 %, issue 54 -> 59 %: at 2-3 resident warps
// per scheduler the pass is then bound by dependent-issue latency, not by either pipe.  Kept as a build switch.
#ifndef NZ_RL_POLY_AGG
#define NZ_RL_POLY_AGG 0   // aggregate passes: states with n % NZ_RL_POLY_AGG == 1 use the polynomial (0: none)
#endif
#ifndef NZ_RL_POLY_FWD
#define NZ_RL_POLY_FWD 0   // forward main pass: states with n % NZ_RL_POLY_FWD == 1
#endif
__host__ __device__ constexpr bool rl_poly_state(int n, int mod) { return mod > 0 && n % mod == 1; }
__device__ __forceinline__ float2 ex2_poly2(float2 x) {
  const float2 magic = make_float2(12582912.f, 12582912.f);
  const float2 t = __fadd2_rn(x, magic);
  const float2 i = __fadd2_rn(t, make_float2(-12582912.f, -12582912.f));
  const float2 f = __fadd2_rn(x, make_float2(-i.x, -i.y));
  float2 p = make_float2(0.001326472731307149f, 0.001326472731307149f);
  p = __ffma2_rn(p, f, make_float2(0.009671512991189957f, 0.009671512991189957f));
  p = __ffma2_rn(p, f, make_float2(0.05550733581185341f, 0.05550733581185341f));
  p = __ffma2_rn(p, f, make_float2(0.24022242426872253f, 0.24022242426872253f));
  p = __ffma2_rn(p, f, make_float2(0.6931470036506653f, 0.6931470036506653f));
  p = __ffma2_rn(p, f, make_float2(1.f, 1.f));
  float2 r;
  r.x = __int_as_float(__float_as_int(p.x) + (__float_as_int(t.x) << 23));
  r.y = __int_as_float(__float_as_int(p.y) + (__float_as_int(t.y) << 23));
  return r;
}
// a_t for a time pair: MUFU or polynomial (`poly` is a compile-time constant after the state loop is unrolled)
__device__ __forceinline__ float2 ex2_pair(bool poly, float2 x) {
  if (poly) return ex2_poly2(x);
  return make_float2(ex2_approx(x.x), ex2_approx(x.y));
}

#ifndef NZ_RL_BWD_MINB
#define NZ_RL_BWD_MINB 12  // resident warps per SM the backward main pass is compiled for (register cap 65536 / 32 / MINB)
#endif

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
// Scheduling fence: `x` may not be consumed before `dep` has been produced.  ptxas otherwise clumps a latency chain
// (LDS -> FADD2 -> SHFL -> FADD ...) right behind its first load and the in-order warp sits on the scoreboard with
// independent work queued behind it (profiles/r02_kernel_tuning.md: 45 % of the state loop's samples).  An empty asm
// only orders the PTX, which ptxas reschedules freely; this is one real LOP3, x |= dep & zero, with a zero ptxas cannot
// see through (RlArgs::zero, a kernel argument).
__device__ __forceinline__ void order_after(float& x, float dep, unsigned zero) {
  asm volatile("lop3.b32 %0, %0, %1, %2, 0xF8;" : "+f"(x) : "f"(dep), "r"(zero));
}
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
// A lane's 8 results of one block go out as ONE 32-byte store when they are fp32 (st.global.v8.f32, sm_100: STG.E.ENL2.256):
// two 16-byte stores write every 32-byte sector in halves, and the L2 then spends two partial-sector writes on it -- the
// forward main pass sat at 69 % L2 throughput with 2 sector writes per sector of data (profiles/r02_kernel_tuning.md).
__device__ __forceinline__ void stg256(float* p, const float* v) {
  asm volatile("st.global.v8.f32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(p), "f"(v[0]), "f"(v[1]), "f"(v[2]),
               "f"(v[3]), "f"(v[4]), "f"(v[5]), "f"(v[6]), "f"(v[7])
               : "memory");
}
template <typename T>
__device__ __forceinline__ void stg_blk8(T* row, const float (&v)[8], long t0, long L, bool wide) {
  if constexpr (sizeof(T) == 4) {
    if (wide) {
      stg256(reinterpret_cast<float*>(row) + t0, v);
      return;
    }
  }
  stg_items<T, 8>(row, v, t0, L, true);
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

// state n's 8 steps out of a per-block B / C tile [16 states][8 steps] (every lane reads the same address: broadcast)
template <typename T>
__device__ __forceinline__ void lds_bc(uint32_t tile_s, int n, float (&v)[8]) {
  if constexpr (sizeof(T) == 4) {
    const uint32_t base = tile_s + (uint32_t)n * 32u;
    const uint32_t sw = (uint32_t)((n >> 2) & 1) << 4;
    unpack16<T>(lds128(base + sw), &v[0]);
    unpack16<T>(lds128(base + (sw ^ 16u)), &v[4]);
  } else {
    unpack16<T>(lds128(tile_s + (uint32_t)n * 16u), &v[0]);
  }
}

// ================================================================================================
// Aggregate pass: (Q, G) of the reverse recurrence per (row, chunk, state)
// ================================================================================================
// kFwd = false: reverse recurrence R_{t-1} = a_t (C_t dy_t + R_t) over (delta, dout, [z], C);  chunks 1 .. nchunks-1
// kFwd = true : forward recurrence h_t = a_t h_{t-1} + dl_t u_t B_t over (delta, u, B);        chunks 0 .. nchunks-2
// kRevCap: compiled with support for reversed groups (RlArgs::rev_mask) and shared u rows (RlArgs::u_gdiv); the plain
// instantiation carries neither the branch nor the second copy of the state loop.
template <typename T, bool kHasZ, bool kFwd, bool kRevCap = false>
__global__ void __launch_bounds__(32, 10) scan_rl_agg_kernel(const __grid_constant__ RlArgs a) {
  using Cfg = RlCfg<T>;
  constexpr int TB = Cfg::TB, NBLK = Cfg::NBLK, ROWT = Cfg::ROWT, BCT = Cfg::BCT;
  constexpr int NROW = kHasZ ? 3 : 2;               // delta, dout, [z]
  constexpr int STAGE = NROW * ROWT + BCT;          // bytes per pipeline stage
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + 2 * STAGE);
  const uint32_t smem_s = smem_u32(smem);
  const int lane = threadIdx.x;

  // work item: nobody needs the aggregate of the first (backward) / last (forward) chunk in time
  const int nc1 = a.nchunks - 1;
  const int item = blockIdx.x;
  int w = item / nc1;
  const int rb = w % a.nrb;
  w /= a.nrb;
  const int g = w % a.ngroups, b = w / a.ngroups;
  const bool rev = kRevCap && ((a.rev_mask >> g) & 1);
  const bool down = kFwd ? rev : !rev;  // walking direction along the sequence
  const int c = item % nc1 + (down ? 1 : 0);  // the last chunk in walking order needs no aggregate
  const int d0 = g * a.dpg + rb * 32, d = d0 + lane;
  const int r1 = (kRevCap && kFwd) ? (g / a.u_gdiv) * a.dpg + rb * 32 : d0;  // rows of g_row1 (u when kFwd)
  const long rowg = (long)b * a.dim + d;
  const int t_lo = c * a.tpc, t_hi = min(a.ntl, t_lo + a.tpc);

  auto issue = [&](int t, int s) {
    uint8_t* st = smem + s * STAGE;
    mbar_arrive_expect_tx(&bars[s], STAGE);
    tma_load_4d(st, &a.g_delta, &bars[s], 0, t, d0, b);
    tma_load_4d(st + ROWT, &a.g_row1, &bars[s], 0, t, r1, b);
    if (kHasZ) tma_load_4d(st + 2 * ROWT, &a.g_z, &bars[s], 0, t, d0, b);
    tma_load_5d(st + NROW * ROWT, &a.g_bc, &bars[s], 0, t, 0, g, b);
  };
  const int nt = t_hi - t_lo;
  auto tile_of = [&](int k) { return down ? t_hi - 1 - k : t_lo + k; };  // k-th tile in walking order
  if (lane == 0) {
    mbar_init(&bars[0], 1);
    mbar_init(&bars[1], 1);
    fence_mbar_init();
    issue(tile_of(0), 0);
    if (nt > 1) issue(tile_of(1), 1);
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
  float amax = 0.f;
#pragma unroll
  for (int n = 0; n < kMaxState; ++n) amax = fmaxf(amax, fabsf(A2[n]));
  const float dlim = 126.f / fmaxf(amax, 1e-30f);  // |A2 * dl| <= 126 for the exponent arithmetic of ex2_poly2

  // dB / dC of this (batch, group), zero-filled tile by tile by the warp of row block 0 (kFwd == false only)
  auto zero_tile = [&](int t) {
    float* zb = a.dB + (((long)b * a.ngroups + g) * kMaxState) * a.L + (long)t * TB;
    float* zc = a.dC + (((long)b * a.ngroups + g) * kMaxState) * a.L + (long)t * TB;
#pragma unroll 4
    for (int n = 0; n < kMaxState; ++n)
#pragma unroll
      for (int j = 0; j < TB; j += 32) {
        zb[(long)n * a.L + j + lane] = 0.f;
        zc[(long)n * a.L + j + lane] = 0.f;
      }
  };
  const bool zero_dbc = !kFwd && a.zero_dbc && rb == 0;
  for (int k = 0; k < nt; ++k) {
    const int s = k & 1;
    if (zero_dbc) zero_tile(tile_of(k));
    mbar_wait(&bars[s], (k >> 1) & 1);
    const uint32_t st = smem_s + s * STAGE;
#pragma unroll 1
    for (int bi = 0; bi < NBLK; ++bi) {
      const int blk = down ? NBLK - 1 - bi : bi;
      float dl[8], dy[8];  // dy: dout (backward) / u (forward)
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
        dlsum += x;
        if (kFwd) dy[i] *= x;  // dl_t u_t
        dl[i] = NZ_RL_POLY_AGG > 0 ? fminf(fmaxf(x, -dlim), dlim) : x;
      }
      auto states = [&](auto down_tag) {
        constexpr bool kDown = decltype(down_tag)::value;
#pragma unroll
        for (int n = 0; n < kMaxState; ++n) {
          float cv[8];
          lds_block<T>(st + NROW * ROWT, n, blk, cv);
          float av[8];
#pragma unroll
          for (int kk = 0; kk < 4; ++kk) {
            const float2 e2 =
                ex2_pair(rl_poly_state(n, NZ_RL_POLY_AGG), mul2(f2(dl[2 * kk], dl[2 * kk + 1]), f2(A2[n], A2[n])));
            av[2 * kk] = e2.x;
            av[2 * kk + 1] = e2.y;
          }
          float r = R[n];
#pragma unroll
          for (int ii = 0; ii < 8; ++ii) {
            const int i = kDown ? 7 - ii : ii;
            if constexpr (kFwd)
              r = fmaf(av[i], r, cv[i] * dy[i]);  // h_t = a_t h_{t-1} + b_t
            else
              r = av[i] * fmaf(cv[i], dy[i], r);  // R_{t-1} = a_t (C_t dy_t + R_t)
          }
          R[n] = r;
        }
      };
      if constexpr (kRevCap) {
        if (down)
          states(std::true_type{});
        else
          states(std::false_type{});
      } else {
        states(std::integral_constant<bool, !kFwd>{});
      }
    }
    __syncwarp();  // every lane is done with stage s
    if (lane == 0 && k + 2 < nt) issue(tile_of(k + 2), s);
  }
  if (zero_dbc) {  // the chunk no work item aggregates (the last in walking order) is zero-filled by its neighbour
    const int cs = down ? 0 : a.nchunks - 1;
    if (c == (down ? 1 : a.nchunks - 2))
      for (int t = cs * a.tpc; t < min(a.ntl, (cs + 1) * a.tpc); ++t) zero_tile(t);
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

// State entering every chunk: one thread per (row, state) walks the chunks, last to first for the backward (R), first
// to last for the forward (h)
static __global__ void __launch_bounds__(128) scan_rl_combine_kernel(const float* __restrict__ G, const float* __restrict__ Q,
                                                                    float* __restrict__ Rin, long nrows, int nchunks,
                                                                    int fwd, int dim, int dpg, int rev_mask) {
  const long i = blockIdx.x * 128L + threadIdx.x;
  if (i >= nrows * kMaxState) return;
  const long row = i / kMaxState;
  const int n = (int)(i % kMaxState);
  fwd ^= (rev_mask >> ((int)(row % dim) / dpg)) & 1;  // reversed groups walk the chunks the other way
  const long base = row * nchunks * kMaxState + n;
  float r = 0.f;
  if (fwd) {
    Rin[base] = 0.f;
    for (int c = 1; c < nchunks; ++c) {
      r = fmaf(Q[base + (long)(c - 1) * kMaxState], r, G[base + (long)(c - 1) * kMaxState]);
      Rin[base + (long)c * kMaxState] = r;
    }
  } else {
    Rin[base + (long)(nchunks - 1) * kMaxState] = 0.f;
    for (int c = nchunks - 2; c >= 0; --c) {
      r = fmaf(Q[base + (long)(c + 1) * kMaxState], r, G[base + (long)(c + 1) * kMaxState]);
      Rin[base + (long)c * kMaxState] = r;
    }
  }
}

// ================================================================================================
// Main pass
// ================================================================================================
// Operands arrive per BLOCK (8 steps) through a 2-stage TMA ring: a block lasts ~5 us per warp, far longer than the
// DRAM latency, so one block of look-ahead is all the buffering needed and shared memory stays small enough for 12
// resident warps per SM (the first version buffered 32-step tiles: 27 KB per warp, 8 warps, issue slots 41 % busy).
//   rows   2 stages x {u, delta, dout, [z]} x 32 rows x (8 steps)       fp32: 32-byte rows, SWIZZLE_32B
//   xf     2 stages x 32 rows x 64 bytes (h entering the block)         SWIZZLE_64B
//   bc     2 stages x {B, C} x 16 states x (8 steps)
//   slab   2 x 2 x 1152   dB / dC products of one (state, block): 4 groups of 8 rows, 288-byte group pitch; double-
//                         buffered so that state n-1 is reduced while the recurrences of state n run
template <typename T, bool kHasZ>
struct RlMainSmem {
  static constexpr int NROWT = kHasZ ? 4 : 3;
  static constexpr int RB = 32 * kFine * (int)sizeof(T);        // one row array of one block (1024 / 512 bytes)
  static constexpr int XB = 32 * kMaxState * 4;                 // fine checkpoints of one block (2048 bytes)
  static constexpr int RSTAGE = ((NROWT * RB + 1023) / 1024) * 1024 + XB;  // xf sits 1024-aligned after the row arrays
  static constexpr int OFF_XF = RSTAGE - XB;
  static constexpr int BB = kMaxState * kFine * (int)sizeof(T); // one B or C array of one block (512 / 256 bytes)
  static constexpr int BSTAGE = 2 * BB;
  static constexpr int SLAB1 = 4 * 288;
  static constexpr int SLAB = 2 * SLAB1;
  static constexpr int OFF_BC = 2 * RSTAGE, OFF_SLAB = OFF_BC + 2 * BSTAGE;
  static constexpr int OFF_BARS = OFF_SLAB + ((2 * SLAB + 127) / 128) * 128;
  static constexpr int TOTAL = OFF_BARS + 64;
  static constexpr size_t bytes() { return 1024 + TOTAL; }
};

// a lane's 8 items of a per-block tile [rows][8 steps]: fp32 rows are 32 bytes under SWIZZLE_32B (the two 16-byte
// halves swap when address bit 7 is set, i.e. for rows 4-7 of every 8), 16-bit rows are 16 bytes, unswizzled
template <typename T>
__device__ __forceinline__ void lds_blockrow(uint32_t tile_s, int row, float (&v)[8]) {
  if constexpr (sizeof(T) == 4) {
    const uint32_t base = tile_s + (uint32_t)row * 32u;
    const uint32_t sw = (uint32_t)((row >> 2) & 1) << 4;
    unpack16<T>(lds128(base + sw), &v[0]);
    unpack16<T>(lds128(base + (sw ^ 16u)), &v[4]);
  } else {
    unpack16<T>(lds128(tile_s + (uint32_t)row * 16u), &v[0]);
  }
}

template <typename T, bool kHasZ, bool kSingle>
__global__ void __launch_bounds__(32, NZ_RL_BWD_MINB) scan_bwd_rl_kernel(const __grid_constant__ RlArgs a) {
  using Cfg = RlCfg<T>;
  using SM = RlMainSmem<T, kHasZ>;
  constexpr int NBLK = Cfg::NBLK, RB = SM::RB, BB = SM::BB;
  constexpr int ROWS_TX = SM::NROWT * RB + SM::XB;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + SM::OFF_BARS);  // [0],[1] row stages, [2],[3] B/C stages
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
  const int j_lo = t_lo * NBLK, j_hi = t_hi * NBLK;  // fine blocks [j_lo, j_hi), walked last to first

  auto issue_rows = [&](int j, int s) {
    uint8_t* st = smem + s * SM::RSTAGE;
    mbar_arrive_expect_tx(&bars[s], ROWS_TX);
    tma_load_4d(st, &a.tm_u, &bars[s], 0, j, d0, b);
    tma_load_4d(st + RB, &a.tm_delta, &bars[s], 0, j, d0, b);
    tma_load_4d(st + 2 * RB, &a.tm_dout, &bars[s], 0, j, d0, b);
    if (kHasZ) tma_load_4d(st + 3 * RB, &a.tm_z, &bars[s], 0, j, d0, b);
    // h entering block j = the forward's checkpoint at the end of block j - 1 (block -1 is out of bounds: zero fill)
    tma_load_4d(st + SM::OFF_XF, &a.tm_xf, &bars[s], 0, j - 1, d0, b);
  };
  auto issue_bc = [&](int j, int s) {
    uint8_t* st = smem + SM::OFF_BC + s * SM::BSTAGE;
    mbar_arrive_expect_tx(&bars[2 + s], SM::BSTAGE);
    tma_load_5d(st, &a.tm_B, &bars[2 + s], 0, j, 0, g, b);
    tma_load_5d(st + BB, &a.tm_C, &bars[2 + s], 0, j, 0, g, b);
  };
  if (lane == 0) {
#pragma unroll
    for (int i = 0; i < 4; ++i) mbar_init(&bars[i], 1);
    fence_mbar_init();
    issue_rows(j_hi - 1, 0);
    issue_bc(j_hi - 1, 0);
    if (j_hi - 2 >= j_lo) {
      issue_rows(j_hi - 2, 1);
      issue_bc(j_hi - 2, 1);
    }
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
  // h entering the block: lane's 64 bytes of the xf tile, 16-byte pieces swizzled with (row >> 1) & 3 (SWIZZLE_64B)
  const uint32_t xf_l = keep((uint32_t)SM::OFF_XF + (uint32_t)lane * 64u);
  const uint32_t xf_key = (uint32_t)((lane >> 1) & 3);
  const unsigned zero = a.zero;
  long Lb = a.L * 4;  // byte pitch of a dB / dC state row, kept in vector registers
  asm volatile("" : "+l"(Lb));

  int k = 0;
#pragma unroll 1
  for (int jb = j_hi - 1; jb >= j_lo; --jb, ++k) {
    const int s = k & 1;
    const uint32_t ph = (uint32_t)(k >> 1) & 1u;
    mbar_wait(&bars[s], ph);
    const uint32_t rst = smem_s + s * SM::RSTAGE;
    float dl[8], dlu[8], dy[8], uu[8], sB[8], ddl[8];
    float yv[kHasZ ? 8 : 1], dzf[kHasZ ? 8 : 1];
    lds_blockrow<T>(rst, lane, uu);
    lds_blockrow<T>(rst + RB, lane, dl);
    lds_blockrow<T>(rst + 2 * RB, lane, dy);
    if constexpr (kHasZ) {
      float zz[8];
      lds_blockrow<T>(rst + 3 * RB, lane, zz);
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const float sg = sigmoid_f(zz[i]);
        dzf[i] = dy[i] * sg * (1.f + zz[i] * (1.f - sg));  // dout * d silu(z)/dz
        dy[i] = dy[i] * zz[i] * sg;                         // dout * silu(z)
        yv[i] = Dv * uu[i];
      }
    }
    float4 hq[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const uint4 v = lds128(rst + xf_l + (((uint32_t)q ^ xf_key) << 4));
      hq[q] = make_float4(__uint_as_float(v.x), __uint_as_float(v.y), __uint_as_float(v.z), __uint_as_float(v.w));
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
    // this stage's row data now lives in registers: request the block after the next one into it
    __syncwarp();
    if (lane == 0 && jb - 2 >= j_lo) issue_rows(jb - 2, s);
    mbar_wait(&bars[2 + s], ph);
    const uint32_t tB = bc_s + s * SM::BSTAGE, tC = tB + BB;
    float* dG = dG0 + (long)jb * kFine;

    // The state loop is software-pipelined by hand (the shared-memory helpers are ordered asm, so the order written
    // here is the order issued): B/C of state n+1 are requested first, the slab of state n-1 is read back before
    // the recurrences of state n start and summed while they run, and the products of state n go to the other slab.
    float bcv[2][2][8];  // [buffer][B, C][step]
    lds_bc<T>(tB, 0, bcv[0][0]);
    lds_bc<T>(tC, 0, bcv[0][1]);
    float2 racc[8];
    char* dGn = reinterpret_cast<char*>(dG);  // dB / dC row of the state being reduced
    auto slab_load = [&](int buf) {
      __syncwarp();
#pragma unroll
      for (int i = 0; i < 8; ++i) racc[i] = lds64((i < 4 ? slab_r0 : slab_r1) + i * 32 + buf * SM::SLAB);
    };
    // the reduction of the previous state's slab is spread over the current state's work in three steps, each pinned
    // behind a value the current state produces at that point (order_after)
    float2 acc;
    auto slab_sum = [&](float dep) {
#pragma unroll
      for (int i = 0; i < 8; i += 2) order_after(racc[i].x, dep, zero);
      acc = __fadd2_rn(__fadd2_rn(__fadd2_rn(racc[0], racc[1]), __fadd2_rn(racc[2], racc[3])),
                       __fadd2_rn(__fadd2_rn(racc[4], racc[5]), __fadd2_rn(racc[6], racc[7])));
    };
    auto slab_xor = [&](int m, float dep) {
      order_after(acc.x, dep, zero);
      const float ox = __shfl_xor_sync(0xffffffffu, acc.x, m), oy = __shfl_xor_sync(0xffffffffu, acc.y, m);
      acc.x += ox;
      acc.y += oy;
    };
    auto slab_store = [&](float dep) {
      order_after(acc.x, dep, zero);
      if (r_q == 0) stg64_or_red(reinterpret_cast<float*>(dGn), acc.x, acc.y, kSingle);
      dGn += Lb;
    };
#pragma unroll
    for (int n = 0; n < kMaxState; ++n) {
      float (&bv)[8] = bcv[n & 1][0];
      float (&cdy)[8] = bcv[n & 1][1];
      if (n + 1 < kMaxState) {
        lds_bc<T>(tB, n + 1, bcv[(n + 1) & 1][0]);
        lds_bc<T>(tC, n + 1, bcv[(n + 1) & 1][1]);
      }
      const float An = A2[n] * kLn2;
      const float4 h4 = hq[n >> 2];
      const float hin = (n & 3) == 0 ? h4.x : (n & 3) == 1 ? h4.y : (n & 3) == 2 ? h4.z : h4.w;
      float av[8], hh[8], dd[8];
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
      if (n > 0) slab_load((n - 1) & 1);
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
        if (i == 4 && n > 0) slab_sum(h);
      }
      R[n] = av[0] * dd[0];  // R leaving the block
      if (n > 0) slab_xor(1, hh[7]);
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
        if (kk == 1 && n > 0) slab_xor(2, sB[2]);
      }
      if (n > 0) slab_store(sB[6]);
      dAacc[n] += gs2.x + gs2.y;  // this row's share of dA[n]
      // ---- dB / dC products of the warp's 32 rows go to slab buffer n & 1 (reduced during state n + 1) ----
      const uint32_t sw = slab_w + (n & 1) * SM::SLAB;
      sts128(sw, vB[0], vB[1], vB[2], vB[3]);
      sts128(sw ^ 16u, vB[4], vB[5], vB[6], vB[7]);
      sts128(sw + SM::SLAB1, vC[0], vC[1], vC[2], vC[3]);
      sts128((sw ^ 16u) + SM::SLAB1, vC[4], vC[5], vC[6], vC[7]);
    }
    slab_load((kMaxState - 1) & 1);
    slab_sum(0.f);
    slab_xor(1, 0.f);
    slab_xor(2, 0.f);
    slab_store(0.f);
    __syncwarp();  // the next block writes slab buffer 0 again; every lane is done with B/C stage s
    if (lane == 0 && jb - 2 >= j_lo) issue_bc(jb - 2, s);

    // ---- per-(row, t) epilogue of the block ----
    const long tpos = (long)jb * kFine;
    float outv[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) outv[i] = fmaf(dl[i], sB[i], Dv * dy[i]);  // du
    stg_blk8<T>(reinterpret_cast<T*>(a.du) + rowg * a.L, outv, tpos, a.L, a.wide);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      float gd = fmaf(uu[i], sB[i], ddl[i]);  // d loss / d dl
      if (a.softplus) gd *= sigmoid_from_softplus(dl[i]);
      outv[i] = gd;
      db_acc += gd;
      dD_acc = fmaf(dy[i], uu[i], dD_acc);
    }
    stg_blk8<T>(reinterpret_cast<T*>(a.ddelta) + rowg * a.L, outv, tpos, a.L, a.wide);
    if constexpr (kHasZ) {
#pragma unroll
      for (int i = 0; i < 8; ++i) outv[i] = dzf[i] * yv[i];
      stg_blk8<T>(reinterpret_cast<T*>(a.dz) + rowg * a.L, outv, tpos, a.L, a.wide);
    }
  }
  // (dim)-shaped sums: over batch and chunks with fp32 atomics
#pragma unroll
  for (int n = 0; n < kMaxState; ++n) atomicAdd(a.dA + (long)d * kMaxState + n, dAacc[n]);
  if (a.dD) atomicAdd(a.dD + d, dD_acc);
  if (a.dbias) atomicAdd(a.dbias + d, db_acc);
}

// ================================================================================================
// Main pass, butterfly version: the dB / dC row reduction stays in registers
// ================================================================================================
// The slab version above moves every dB / dC product through shared memory (256 bytes per element: its shared-memory
// floor of 2.7 clk/element is the largest pipe cost of the pass, and the STS -> LDS -> add -> SHFL chain needs the
// hand-made scheduling fences).  Here the 32 rows of the warp are summed by butterfly exchanges:
//   * lane l evaluates state n = s ^ (l & 3) in code slot s (A2 / R / dA / h_in are loaded with the permuted index, the
//     B / C broadcasts become 4 adjacent 32-byte rows = still one wavefront).  Lanes l and l ^ 1 then hold the SAME state
//     in slots s and s ^ 1, so "keep the even slot, send the odd slot" sums lane pairs without a single select; the
//     same holds for l ^ 2 and slots s, s ^ 2.  A quad of slots (64 products per lane) therefore collapses to 16 values
//     per lane, each the sum over 4 lanes, with 48 SHFL + 24 FADD2;
//   * the remaining three lane bits split those 16 values (array B / C, time half, time pair) with selects:
//     14 SHFL + 28 SEL + 7 FADD2, after which a lane owns one (array, state, time pair) of the quad = one 8-byte store.
// 248 shuffle wavefronts per block of 8 steps against ~640 for the slab, no __syncwarp in the state loop.
template <typename T, bool kHasZ>
struct RlMain2Smem {
  static constexpr int NROWT = kHasZ ? 4 : 3;
  static constexpr int RB = 32 * kFine * (int)sizeof(T);
  static constexpr int XB = 32 * kMaxState * 4;
  static constexpr int RSTAGE = ((NROWT * RB + 1023) / 1024) * 1024 + XB;
  static constexpr int OFF_XF = RSTAGE - XB;
  static constexpr int BB = kMaxState * kFine * (int)sizeof(T);
  static constexpr int BSTAGE = 2 * BB;
  static constexpr int OFF_BC = 2 * RSTAGE;
  static constexpr int OFF_BARS = OFF_BC + 2 * BSTAGE;
  static constexpr int TOTAL = OFF_BARS + 64;
  static constexpr size_t bytes() { return 1024 + TOTAL; }
};

// Resident warps per SM the butterfly pass is compiled for.  Its natural register demand is 249: at 12 warps (168
// registers) it spills 224 bytes per lane and runs 2.07 ms on 12 x 128 x 65536, at 8 warps (no spills) 1.88 ms, at 14 / 16
// warps (128 registers) 2.6 - 2.7 ms (profiles/r02_kernel_tuning.md).
#ifndef NZ_RL_BWD2_MINB
#define NZ_RL_BWD2_MINB 8
#endif

__device__ __forceinline__ float2 shfl_xor2(float2 v, int mask) {
  return make_float2(__shfl_xor_sync(0xffffffffu, v.x, mask), __shfl_xor_sync(0xffffffffu, v.y, mask));
}
__device__ __forceinline__ float2 sel2(bool p, float2 a, float2 b) { return make_float2(p ? a.x : b.x, p ? a.y : b.y); }

// state (4q + jj)'s 8 steps of a per-block B / C tile, jj = the lane's permuted low bits already folded into `base`
template <typename T>
__device__ __forceinline__ void lds_bc2(uint32_t base, int q, float (&v)[8]) {
  if constexpr (sizeof(T) == 4) {
    const uint32_t o = (uint32_t)q * 128u, sw = (uint32_t)(q & 1) << 4;
    unpack16<T>(lds128(base + o + sw), &v[0]);
    unpack16<T>(lds128(base + o + (sw ^ 16u)), &v[4]);
  } else {
    unpack16<T>(lds128(base + (uint32_t)q * 64u), &v[0]);
  }
}

#ifdef NZ_RL_BWD2_MAXREG  // tuning: an explicit register cap instead of a residency target
#define NZ_RL_BWD2_BOUNDS __maxnreg__(NZ_RL_BWD2_MAXREG)
#else
#define NZ_RL_BWD2_BOUNDS __launch_bounds__(32, NZ_RL_BWD2_MINB)
#endif
template <typename T, bool kHasZ, bool kSingle, bool kRevCap = false>
__global__ void NZ_RL_BWD2_BOUNDS scan_bwd_rl2_kernel(const __grid_constant__ RlArgs a) {
  using Cfg = RlCfg<T>;
  using SM = RlMain2Smem<T, kHasZ>;
  constexpr int NBLK = Cfg::NBLK, RB = SM::RB, BB = SM::BB;
  constexpr int ROWS_TX = SM::NROWT * RB + SM::XB;
  constexpr uint32_t RP = kFine * sizeof(T);  // pitch of a state row in a B / C tile
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + SM::OFF_BARS);  // [0],[1] row stages, [2],[3] B/C stages
  const uint32_t smem_s = keep(smem_u32(smem));
  const uint32_t bc_s = smem_s + SM::OFF_BC;
  const int lane = threadIdx.x;
  const int m = lane & 3;  // this lane's state permutation: slot s holds state s ^ m

  const int item = blockIdx.x;
  const int c = item % a.nchunks;
  int w = item / a.nchunks;
  const int rb = w % a.nrb;
  w /= a.nrb;
  const int g = w % a.ngroups, b = w / a.ngroups;
  const int d0 = g * a.dpg + rb * 32, d = d0 + lane;
  const long rowg = (long)b * a.dim + d;
  const int t_lo = c * a.tpc, t_hi = min(a.ntl, t_lo + a.tpc);
  const int j_lo = t_lo * NBLK, j_hi = t_hi * NBLK;  // fine blocks [j_lo, j_hi), walked last to first
  const bool rev = kRevCap && ((a.rev_mask >> g) & 1);              // the forward walked this group backwards
  const int ur = kRevCap ? (g / a.u_gdiv) * a.dpg + rb * 32 : d0;   // rows of u
  // Walking order of the blocks: last to first, or first to last for a reversed group.  (Block indices are stepped, not
  // derived from a trip counter: with `jb = f(k)` ptxas moved the counter into a uniform register and the kernel then
  // produced sparse, run-to-run different errors at 1 MB row pitch -- profiles/r02_kernel_tuning.md.)
  const int step = rev ? 1 : -1;
  const int j_first = rev ? j_lo : j_hi - 1;
  auto inside = [&](int j) { return j >= j_lo && j < j_hi; };

  auto issue_rows = [&](int j, int s) {
    uint8_t* st = smem + s * SM::RSTAGE;
    mbar_arrive_expect_tx(&bars[s], ROWS_TX);
    tma_load_4d(st, &a.tm_u, &bars[s], 0, j, ur, b);
    tma_load_4d(st + RB, &a.tm_delta, &bars[s], 0, j, d0, b);
    tma_load_4d(st + 2 * RB, &a.tm_dout, &bars[s], 0, j, d0, b);
    if (kHasZ) tma_load_4d(st + 3 * RB, &a.tm_z, &bars[s], 0, j, d0, b);
    tma_load_4d(st + SM::OFF_XF, &a.tm_xf, &bars[s], 0, j + step, d0, b);  // the block the FORWARD walked before j; out of bounds: zero fill
  };
  auto issue_bc = [&](int j, int s) {
    uint8_t* st = smem + SM::OFF_BC + s * SM::BSTAGE;
    mbar_arrive_expect_tx(&bars[2 + s], SM::BSTAGE);
    tma_load_5d(st, &a.tm_B, &bars[2 + s], 0, j, 0, g, b);
    tma_load_5d(st + BB, &a.tm_C, &bars[2 + s], 0, j, 0, g, b);
  };
  if (lane == 0) {
#pragma unroll
    for (int i = 0; i < 4; ++i) mbar_init(&bars[i], 1);
    fence_mbar_init();
    issue_rows(j_first, 0);
    issue_bc(j_first, 0);
    if (inside(j_first + step)) {
      issue_rows(j_first + step, 1);
      issue_bc(j_first + step, 1);
    }
  }
  __syncwarp();
  float A2[kMaxState], R[kMaxState], dAacc[kMaxState];
  {
    const float* rin = a.Rin + (rowg * a.nchunks + c) * kMaxState;
    const float* Ar = a.A + (long)d * a.A_ds;
#pragma unroll
    for (int s = 0; s < kMaxState; ++s) {
      A2[s] = __ldg(Ar + (s ^ m)) * kLog2e;
      R[s] = a.nchunks > 1 ? __ldg(rin + (s ^ m)) : 0.f;  // a_{t+1} dh_{t+1} entering the chunk's last step
      dAacc[s] = 0.f;
    }
  }
  const float Dv = a.D ? __ldg(a.D + d) : 0.f;
  const float bias = a.bias ? __ldg(a.bias + d) : 0.f;
  float dD_acc = 0.f, db_acc = 0.f;

  // B / C rows of the lane's four permuted low state bits, relative to a tile
  uint32_t bco[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) bco[j] = keep((uint32_t)(j ^ m) * RP);
  // h entering the block: lane's 64 bytes of the xf tile, 16-byte pieces swizzled with (row >> 1) & 3 (SWIZZLE_64B)
  const uint32_t xf_l = keep((uint32_t)SM::OFF_XF + (uint32_t)lane * 64u);
  const uint32_t xf_key = (uint32_t)((lane >> 1) & 3);
  const bool m0 = (lane & 1) != 0, m1 = (lane & 2) != 0;
  const bool p2 = (lane & 4) != 0, p3 = (lane & 8) != 0, p4 = (lane & 16) != 0;
  // where this lane's share of the reduced dB / dC goes: array p2, state 4q + m, steps 4 p3 + 2 p4 (+1) of the block
  char* dG0 = reinterpret_cast<char*>((p2 ? a.dC : a.dB) + (((long)b * a.ngroups + g) * kMaxState + m) * a.L +
                                      (p3 ? 4 : 0) + (p4 ? 2 : 0));
  long Lq = a.L * 16;  // byte pitch of four state rows of dB / dC
  asm volatile("" : "+l"(Lq));

  int k = 0;
#pragma unroll 1
  for (int jb = j_first; inside(jb); jb += step, ++k) {
    const int s = k & 1;
    const uint32_t ph = (uint32_t)(k >> 1) & 1u;
    mbar_wait(&bars[s], ph);
    const uint32_t rst = smem_s + s * SM::RSTAGE;
    float dl[8], dlu[8], dy[8], uu[8], sB[8], ddl[8];
    float yv[kHasZ ? 8 : 1], dzf[kHasZ ? 8 : 1];
    lds_blockrow<T>(rst, lane, uu);
    lds_blockrow<T>(rst + RB, lane, dl);
    lds_blockrow<T>(rst + 2 * RB, lane, dy);
    if constexpr (kHasZ) {
      float zz[8];
      lds_blockrow<T>(rst + 3 * RB, lane, zz);
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const float sg = sigmoid_f(zz[i]);
        dzf[i] = dy[i] * sg * (1.f + zz[i] * (1.f - sg));  // dout * d silu(z)/dz
        dy[i] = dy[i] * zz[i] * sg;                         // dout * silu(z)
        yv[i] = Dv * uu[i];
      }
    }
    float hp[kMaxState];  // h entering the block, in slot order
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const uint4 v = lds128(rst + xf_l + (((uint32_t)q ^ xf_key) << 4));
      const float x = __uint_as_float(v.x), y = __uint_as_float(v.y), z = __uint_as_float(v.z), ww = __uint_as_float(v.w);
      const float x1 = m0 ? y : x, y1 = m0 ? x : y, z1 = m0 ? ww : z, w1 = m0 ? z : ww;
      hp[4 * q + 0] = m1 ? z1 : x1;
      hp[4 * q + 1] = m1 ? w1 : y1;
      hp[4 * q + 2] = m1 ? x1 : z1;
      hp[4 * q + 3] = m1 ? y1 : w1;
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
    // this stage's row data now lives in registers: request the block after the next one into it
    __syncwarp();
    if (lane == 0 && inside(jb + 2 * step)) issue_rows(jb + 2 * step, s);
    mbar_wait(&bars[2 + s], ph);
    const uint32_t tB = bc_s + s * SM::BSTAGE;
    uint32_t tbj[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) tbj[j] = tB + bco[j];
    char* dGq = dG0 + (long)jb * (kFine * 4);

    float bcv[2][2][8];  // [buffer][B, C][step]: the next slot's rows are requested before the current slot's math
    lds_bc2<T>(tbj[0], 0, bcv[0][0]);
    lds_bc2<T>(tbj[0] + BB, 0, bcv[0][1]);
    float2 K0[8], K1[8];  // [0..3] dB, [4..7] dC of a slot pair, as time pairs
    auto states = [&](auto rev_tag) {
    constexpr bool kR = decltype(rev_tag)::value;
#pragma unroll
    for (int n = 0; n < kMaxState; ++n) {  // n = code slot; the state is n ^ m
      float (&bv)[8] = bcv[n & 1][0];
      float (&cdy)[8] = bcv[n & 1][1];
      if (n + 1 < kMaxState) {
        lds_bc2<T>(tbj[(n + 1) & 3], (n + 1) >> 2, bcv[(n + 1) & 1][0]);
        lds_bc2<T>(tbj[(n + 1) & 3] + BB, (n + 1) >> 2, bcv[(n + 1) & 1][1]);
      }
      const float An = A2[n] * kLn2;
      float av[8], hh[8], dd[8];
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
      // (a reversed group -- kR -- walked the block from step 7 down to 0, so both chains run the other way)
      float bsave[8];
      float h = hp[n];
      float dh = cdy[kR ? 0 : 7] + R[n];
      dd[kR ? 0 : 7] = dh;
#pragma unroll
      for (int ii = 0; ii < 8; ++ii) {
        const int i = kR ? 7 - ii : ii;
        bsave[i] = hh[i];
        h = fmaf(av[i], h, hh[i]);
        hh[i] = h;
        if (ii < 7) {
          const int j = kR ? ii + 1 : 6 - ii;          // step whose dh is next
          const int jn = kR ? ii : 7 - ii;             // the step after it in walking order
          dh = fmaf(av[jn], dh, cdy[j]);
          dd[j] = dh;
        }
      }
      R[n] = av[kR ? 7 : 0] * dd[kR ? 7 : 0];  // R leaving the block
      // element-wise products, packed over time pairs
      float2 gs2 = f2(0.f, 0.f);
      float2 V[8];  // this slot's dB (0..3) and dC (4..7) products
#pragma unroll
      for (int kk = 0; kk < 4; ++kk) {
        const float2 h2 = f2(hh[2 * kk], hh[2 * kk + 1]);
        const float2 d2 = f2(dd[2 * kk], dd[2 * kk + 1]);
        V[4 + kk] = mul2(f2(dy[2 * kk], dy[2 * kk + 1]), h2);  // dC_t[n] of this row
        const float2 ah2 = sub2(h2, f2(bsave[2 * kk], bsave[2 * kk + 1]));  // a_t h_{t-1}
        const float2 gq2 = mul2(d2, ah2);
        float2 ddl2 = f2(ddl[2 * kk], ddl[2 * kk + 1]);
        ddl2 = fma2(f2(An, An), gq2, ddl2);
        ddl[2 * kk] = ddl2.x;
        ddl[2 * kk + 1] = ddl2.y;
        gs2 = fma2(f2(dl[2 * kk], dl[2 * kk + 1]), gq2, gs2);
        V[kk] = mul2(d2, f2(dlu[2 * kk], dlu[2 * kk + 1]));  // dB_t[n] of this row
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
      dAacc[n] += gs2.x + gs2.y;  // this row's share of dA[n ^ m]
      // ---- row reduction: select-free butterflies over the permuted slots ----
      if ((n & 3) == 0) {
#pragma unroll
        for (int i = 0; i < 8; ++i) K0[i] = V[i];
      } else if ((n & 3) == 1) {
#pragma unroll
        for (int i = 0; i < 8; ++i) K0[i] = __fadd2_rn(K0[i], shfl_xor2(V[i], 1));
      } else if ((n & 3) == 2) {
#pragma unroll
        for (int i = 0; i < 8; ++i) K1[i] = V[i];
      } else {
#pragma unroll
        for (int i = 0; i < 8; ++i) K1[i] = __fadd2_rn(K1[i], shfl_xor2(V[i], 1));
#pragma unroll
        for (int i = 0; i < 8; ++i) K0[i] = __fadd2_rn(K0[i], shfl_xor2(K1[i], 2));
        // K0 = state 4q + m summed over the lane quad; split the 16 values over the other three lane bits
        float2 X[4], Y[2];
#pragma unroll
        for (int i = 0; i < 4; ++i)
          X[i] = __fadd2_rn(sel2(p2, K0[4 + i], K0[i]), shfl_xor2(sel2(p2, K0[i], K0[4 + i]), 4));
#pragma unroll
        for (int i = 0; i < 2; ++i)
          Y[i] = __fadd2_rn(sel2(p3, X[2 + i], X[i]), shfl_xor2(sel2(p3, X[i], X[2 + i]), 8));
        const float2 Z = __fadd2_rn(sel2(p4, Y[1], Y[0]), shfl_xor2(sel2(p4, Y[0], Y[1]), 16));
        stg64_or_red(reinterpret_cast<float*>(dGq), Z.x, Z.y, kSingle);
        dGq += Lq;
      }
    }
    };
    if constexpr (kRevCap) {
      if (rev)
        states(std::true_type{});
      else
        states(std::false_type{});
    } else {
      states(std::false_type{});
    }
    __syncwarp();  // every lane is done with B/C stage s
    if (lane == 0 && inside(jb + 2 * step)) issue_bc(jb + 2 * step, s);

    // ---- per-(row, t) epilogue of the block ----
    const long tpos = (long)jb * kFine;
    float outv[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) outv[i] = fmaf(dl[i], sB[i], Dv * dy[i]);  // du
    stg_blk8<T>(reinterpret_cast<T*>(a.du) + rowg * a.L, outv, tpos, a.L, a.wide);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      float gd = fmaf(uu[i], sB[i], ddl[i]);  // d loss / d dl
      if (a.softplus) gd *= sigmoid_from_softplus(dl[i]);
      outv[i] = gd;
      db_acc += gd;
      dD_acc = fmaf(dy[i], uu[i], dD_acc);
    }
    stg_blk8<T>(reinterpret_cast<T*>(a.ddelta) + rowg * a.L, outv, tpos, a.L, a.wide);
    if constexpr (kHasZ) {
#pragma unroll
      for (int i = 0; i < 8; ++i) outv[i] = dzf[i] * yv[i];
      stg_blk8<T>(reinterpret_cast<T*>(a.dz) + rowg * a.L, outv, tpos, a.L, a.wide);
    }
  }
  // (dim)-shaped sums: over batch and chunks with fp32 atomics
#pragma unroll
  for (int s = 0; s < kMaxState; ++s) atomicAdd(a.dA + (long)d * kMaxState + (s ^ m), dAacc[s]);
  if (a.dD) atomicAdd(a.dD + d, dD_acc);
  if (a.dbias) atomicAdd(a.dbias + d, db_acc);
}

// ================================================================================================
// Forward main pass (row-per-lane)
// ================================================================================================
// Same lane mapping and per-block TMA ring as the backward.  A lane walks its chunk first block to last with all 16
// states in registers: per state one forward recurrence over the block's 8 steps, y_t += C_t[n] h_t[n] packed over time
// pairs.  At every block end the 16 states leave as one 64-byte run of the fine checkpoints (free: they sit in
// registers), every 16th block also as the coarse checkpoint x the C ABI has always returned.
template <typename T, bool kHasZ>
struct RlFwdSmem {
  static constexpr int NROWT = kHasZ ? 3 : 2;                    // u, delta, [z]
  static constexpr int RB = 32 * kFine * (int)sizeof(T);
  static constexpr int RSTAGE = ((NROWT * RB + 1023) / 1024) * 1024;
  static constexpr int BB = kMaxState * kFine * (int)sizeof(T);
  static constexpr int BSTAGE = ((2 * BB + 1023) / 1024) * 1024;
  static constexpr int OFF_BC = 2 * RSTAGE, OFF_BARS = OFF_BC + 2 * BSTAGE;
  static constexpr int TOTAL = OFF_BARS + 64;
  static constexpr size_t bytes() { return 1024 + TOTAL; }
};

#ifndef NZ_RL_FWD_MINB
#define NZ_RL_FWD_MINB 16
#endif
template <typename T, bool kHasZ, bool kRevCap = false>
__global__ void __launch_bounds__(32, NZ_RL_FWD_MINB) scan_fwd_rl_kernel(const __grid_constant__ RlArgs a) {
  using SM = RlFwdSmem<T, kHasZ>;
  constexpr int NBLK = RlCfg<T>::NBLK, RB = SM::RB, BB = SM::BB;
  constexpr int BPC = NZ_CHUNK / kFine;  // fine blocks per coarse checkpoint
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + SM::OFF_BARS);  // [0],[1] row stages, [2],[3] B/C stages
  const uint32_t smem_s = keep(smem_u32(smem));
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
  const int j_lo = t_lo * NBLK, j_hi = t_hi * NBLK;
  const long nbt = a.L / kFine;
  const bool rev = kRevCap && ((a.rev_mask >> g) & 1);              // this group walks the sequence backwards
  const int ur = kRevCap ? (g / a.u_gdiv) * a.dpg + rb * 32 : d0;   // rows of u
  const int step = rev ? -1 : 1;  // walking order of the blocks (stepped block index, see scan_bwd_rl2_kernel)
  const int j_first = rev ? j_hi - 1 : j_lo;
  auto inside = [&](int j) { return j >= j_lo && j < j_hi; };

  auto issue = [&](int j, int s) {
    uint8_t* st = smem + s * SM::RSTAGE;
    mbar_arrive_expect_tx(&bars[s], SM::NROWT * RB);
    tma_load_4d(st, &a.tm_u, &bars[s], 0, j, ur, b);
    tma_load_4d(st + RB, &a.tm_delta, &bars[s], 0, j, d0, b);
    if (kHasZ) tma_load_4d(st + 2 * RB, &a.tm_z, &bars[s], 0, j, d0, b);
    uint8_t* sb = smem + SM::OFF_BC + s * SM::BSTAGE;
    mbar_arrive_expect_tx(&bars[2 + s], 2 * BB);
    tma_load_5d(sb, &a.tm_B, &bars[2 + s], 0, j, 0, g, b);
    tma_load_5d(sb + BB, &a.tm_C, &bars[2 + s], 0, j, 0, g, b);
  };
  if (lane == 0) {
#pragma unroll
    for (int i = 0; i < 4; ++i) mbar_init(&bars[i], 1);
    fence_mbar_init();
    issue(j_first, 0);
    if (inside(j_first + step)) issue(j_first + step, 1);
  }
  __syncwarp();
  float A2[kMaxState], h[kMaxState];
  {
    const float* hin = a.Rin + (rowg * a.nchunks + c) * kMaxState;
#pragma unroll
    for (int n = 0; n < kMaxState; ++n) {
      A2[n] = __ldg(a.A + (long)d * a.A_ds + n) * kLog2e;
      h[n] = a.nchunks > 1 ? __ldg(hin + n) : 0.f;
    }
  }
  const float Dv = a.D ? __ldg(a.D + d) : 0.f;
  const float bias = a.bias ? __ldg(a.bias + d) : 0.f;
  const long orow = (long)b * a.o_bs + (long)d * a.o_ds;
  float amax = 0.f;
#pragma unroll
  for (int n = 0; n < kMaxState; ++n) amax = fmaxf(amax, fabsf(A2[n]));
  const float dlim = 126.f / fmaxf(amax, 1e-30f);  // |A2 * dl| <= 126 for the exponent arithmetic of ex2_poly2

  int k = 0;
#pragma unroll 1
  for (int jb = j_first; inside(jb); jb += step, ++k) {
    const int s = k & 1;
    const uint32_t ph = (uint32_t)(k >> 1) & 1u;
    mbar_wait(&bars[s], ph);
    const uint32_t rst = smem_s + s * SM::RSTAGE;
    float dl[8], dlu[8], y[8];
    [[maybe_unused]] float zz[kHasZ ? 8 : 1];
    lds_blockrow<T>(rst, lane, dlu);  // u for now
    lds_blockrow<T>(rst + RB, lane, dl);
    if constexpr (kHasZ) lds_blockrow<T>(rst + 2 * RB, lane, zz);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      float x = dl[i] + bias;
      if (a.softplus) x = softplus_f(x);
      y[i] = Dv * dlu[i];
      dlu[i] = x * dlu[i];
      dl[i] = NZ_RL_POLY_FWD > 0 ? fminf(fmaxf(x, -dlim), dlim) : x;  // from here on dl only feeds the exponents
    }
    mbar_wait(&bars[2 + s], ph);
    const uint32_t tB = smem_s + SM::OFF_BC + s * SM::BSTAGE, tC = tB + BB;
    auto states = [&](auto rev_tag) {
      constexpr bool kR = decltype(rev_tag)::value;
#pragma unroll
      for (int n = 0; n < kMaxState; ++n) {
        float bv[8], cv[8], hh[8];
        lds_bc<T>(tB, n, bv);
        lds_bc<T>(tC, n, cv);
        float hc = h[n];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const int kk = kR ? 3 - q : q;
          const float2 e2 =
              ex2_pair(rl_poly_state(n, NZ_RL_POLY_FWD), mul2(f2(dl[2 * kk], dl[2 * kk + 1]), f2(A2[n], A2[n])));
          const float2 b2 = mul2(f2(dlu[2 * kk], dlu[2 * kk + 1]), f2(bv[2 * kk], bv[2 * kk + 1]));
          if constexpr (kR) {
            hc = fmaf(e2.y, hc, b2.y);
            hh[2 * kk + 1] = hc;
            hc = fmaf(e2.x, hc, b2.x);
            hh[2 * kk] = hc;
          } else {
            hc = fmaf(e2.x, hc, b2.x);
            hh[2 * kk] = hc;
            hc = fmaf(e2.y, hc, b2.y);
            hh[2 * kk + 1] = hc;
          }
        }
        h[n] = hc;
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) {
          float2 y2 = f2(y[2 * kk], y[2 * kk + 1]);
          y2 = fma2(f2(cv[2 * kk], cv[2 * kk + 1]), f2(hh[2 * kk], hh[2 * kk + 1]), y2);
          y[2 * kk] = y2.x;
          y[2 * kk + 1] = y2.y;
        }
      }
    };
    if constexpr (kRevCap) {
      if (rev)
        states(std::true_type{});
      else
        states(std::false_type{});
    } else {
      states(std::false_type{});
    }
    // every lane is done with this block's stages: request the block after the next one
    __syncwarp();
    if (lane == 0 && inside(jb + 2 * step)) issue(jb + 2 * step, s);
    if constexpr (kHasZ) {
#pragma unroll
      for (int i = 0; i < 8; ++i) y[i] *= zz[i] * sigmoid_f(zz[i]);
    }
    const long tpos = (long)jb * kFine;
    if (sizeof(T) == 2 && a.out_f32)
      stg_blk8<float>(reinterpret_cast<float*>(a.out) + orow, y, tpos, a.L, a.wide);
    else
      stg_blk8<T>(reinterpret_cast<T*>(a.out) + orow, y, tpos, a.L, a.wide);
    // checkpoints: h at the end of the block
    if (a.xfw) {
      float* xo = a.xfw + (rowg * nbt + jb) * kMaxState;
      if (a.wide) {
        stg256(xo, &h[0]);
        stg256(xo + 8, &h[8]);
      } else {
#pragma unroll
        for (int q = 0; q < 4; ++q)
          reinterpret_cast<float4*>(xo)[q] = make_float4(h[4 * q], h[4 * q + 1], h[4 * q + 2], h[4 * q + 3]);
      }
    }
    // coarse checkpoints: h after the last block of every 128-step chunk in walking order
    if (rev ? jb % BPC == 0 : ((jb + 1) % BPC == 0 || jb + 1 == nbt)) {
      float4* xo = reinterpret_cast<float4*>(a.x + (rowg * a.nck + jb / BPC) * kMaxState);
#pragma unroll
      for (int q = 0; q < 4; ++q) xo[q] = make_float4(h[4 * q], h[4 * q + 1], h[4 * q + 2], h[4 * q + 3]);
    }
  }
}

// ================================================================================================
// Forward main pass, pipelined version
// ================================================================================================
// Same decomposition as scan_fwd_rl_kernel; two changes aimed at the MUFU pipe's idle time (61 % busy there, top stall
// samples on the first use of every shared-memory load: profiles/r02_kernel_tuning.md):
//   * B / C rows of state n + 1 are requested before the math of state n (double-buffered registers);
//   * the row operands of block j + 1 are read, softplus'ed and multiplied in the middle of block j's state loop, so a
//     block no longer starts with a serial mbarrier wait -> LDS -> ex2 -> lg2 chain; the row ring therefore runs two
//     blocks ahead of the B / C ring.
#ifndef NZ_RL_FWD2_MINB
#define NZ_RL_FWD2_MINB 12
#endif
template <typename T, bool kHasZ>
__global__ void __launch_bounds__(32, NZ_RL_FWD2_MINB) scan_fwd_rl2_kernel(const __grid_constant__ RlArgs a) {
  using SM = RlFwdSmem<T, kHasZ>;
  constexpr int NBLK = RlCfg<T>::NBLK, RB = SM::RB, BB = SM::BB;
  constexpr int BPC = NZ_CHUNK / kFine;  // fine blocks per coarse checkpoint
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + SM::OFF_BARS);  // [0],[1] row stages, [2],[3] B/C stages
  const uint32_t smem_s = keep(smem_u32(smem));
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
  const int j_lo = t_lo * NBLK, j_hi = t_hi * NBLK;
  const long nbt = a.L / kFine;

  auto issue_rows = [&](int j, int s) {
    uint8_t* st = smem + s * SM::RSTAGE;
    mbar_arrive_expect_tx(&bars[s], SM::NROWT * RB);
    tma_load_4d(st, &a.tm_u, &bars[s], 0, j, d0, b);
    tma_load_4d(st + RB, &a.tm_delta, &bars[s], 0, j, d0, b);
    if (kHasZ) tma_load_4d(st + 2 * RB, &a.tm_z, &bars[s], 0, j, d0, b);
  };
  auto issue_bc = [&](int j, int s) {
    uint8_t* sb = smem + SM::OFF_BC + s * SM::BSTAGE;
    mbar_arrive_expect_tx(&bars[2 + s], 2 * BB);
    tma_load_5d(sb, &a.tm_B, &bars[2 + s], 0, j, 0, g, b);
    tma_load_5d(sb + BB, &a.tm_C, &bars[2 + s], 0, j, 0, g, b);
  };
  if (lane == 0) {
#pragma unroll
    for (int i = 0; i < 4; ++i) mbar_init(&bars[i], 1);
    fence_mbar_init();
    issue_rows(j_lo, 0);
    issue_bc(j_lo, 0);
    if (j_lo + 1 < j_hi) {
      issue_rows(j_lo + 1, 1);
      issue_bc(j_lo + 1, 1);
    }
  }
  __syncwarp();
  float A2[kMaxState], h[kMaxState];
  {
    const float* hin = a.Rin + (rowg * a.nchunks + c) * kMaxState;
#pragma unroll
    for (int n = 0; n < kMaxState; ++n) {
      A2[n] = __ldg(a.A + (long)d * a.A_ds + n) * kLog2e;
      h[n] = a.nchunks > 1 ? __ldg(hin + n) : 0.f;
    }
  }
  const float Dv = a.D ? __ldg(a.D + d) : 0.f;
  const float bias = a.bias ? __ldg(a.bias + d) : 0.f;
  const long orow = (long)b * a.o_bs + (long)d * a.o_ds;

  float dl[8], dlu[8], y[8];
  [[maybe_unused]] float zz[kHasZ ? 8 : 1];
  // row operands of the block whose rows sit in stage s -> (dl, dl u, D u, z)
  auto prologue = [&](int s, float (&pdl)[8], float (&pdlu)[8], float (&py)[8], float (&pz)[kHasZ ? 8 : 1]) {
    const uint32_t rst = smem_s + s * SM::RSTAGE;
    lds_blockrow<T>(rst, lane, pdlu);  // u for now
    lds_blockrow<T>(rst + RB, lane, pdl);
    if constexpr (kHasZ) lds_blockrow<T>(rst + 2 * RB, lane, pz);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      float x = pdl[i] + bias;
      if (a.softplus) x = softplus_f(x);
      py[i] = Dv * pdlu[i];
      pdlu[i] = x * pdlu[i];
      pdl[i] = x;
    }
  };
  mbar_wait(&bars[0], 0);
  prologue(0, dl, dlu, y, zz);
  __syncwarp();
  if (lane == 0 && j_lo + 2 < j_hi) issue_rows(j_lo + 2, 0);

  int k = 0;
#pragma unroll 1
  for (int jb = j_lo; jb < j_hi; ++jb, ++k) {
    const int s = k & 1;
    const uint32_t ph = (uint32_t)(k >> 1) & 1u;
    mbar_wait(&bars[2 + s], ph);
    const uint32_t tB = smem_s + SM::OFF_BC + s * SM::BSTAGE, tC = tB + BB;
    float ndl[8], ndlu[8], ny[8];
    [[maybe_unused]] float nz[kHasZ ? 8 : 1];
    float bcv[2][2][8];
    lds_bc<T>(tB, 0, bcv[0][0]);
    lds_bc<T>(tC, 0, bcv[0][1]);
#pragma unroll
    for (int n = 0; n < kMaxState; ++n) {
      float (&bv)[8] = bcv[n & 1][0];
      float (&cv)[8] = bcv[n & 1][1];
      if (n + 1 < kMaxState) {
        lds_bc<T>(tB, n + 1, bcv[(n + 1) & 1][0]);
        lds_bc<T>(tC, n + 1, bcv[(n + 1) & 1][1]);
      }
      if (n == kMaxState / 2 && jb + 1 < j_hi) {
        // next block's row operands (stage s ^ 1 of the row ring, phase of block k + 1)
        mbar_wait(&bars[s ^ 1], (uint32_t)((k + 1) >> 1) & 1u);
        prologue(s ^ 1, ndl, ndlu, ny, nz);
        __syncwarp();
        if (lane == 0 && jb + 3 < j_hi) issue_rows(jb + 3, s ^ 1);
      }
      float hh[8];
      float hc = h[n];
#pragma unroll
      for (int kk = 0; kk < 4; ++kk) {
        const float2 x2 = mul2(f2(dl[2 * kk], dl[2 * kk + 1]), f2(A2[n], A2[n]));
        const float2 b2 = mul2(f2(dlu[2 * kk], dlu[2 * kk + 1]), f2(bv[2 * kk], bv[2 * kk + 1]));
        hc = fmaf(ex2_approx(x2.x), hc, b2.x);
        hh[2 * kk] = hc;
        hc = fmaf(ex2_approx(x2.y), hc, b2.y);
        hh[2 * kk + 1] = hc;
      }
      h[n] = hc;
#pragma unroll
      for (int kk = 0; kk < 4; ++kk) {
        float2 y2 = f2(y[2 * kk], y[2 * kk + 1]);
        y2 = fma2(f2(cv[2 * kk], cv[2 * kk + 1]), f2(hh[2 * kk], hh[2 * kk + 1]), y2);
        y[2 * kk] = y2.x;
        y[2 * kk + 1] = y2.y;
      }
    }
    // every lane is done with this block's B / C stage: request the block after the next one
    __syncwarp();
    if (lane == 0 && jb + 2 < j_hi) issue_bc(jb + 2, s);
    if constexpr (kHasZ) {
#pragma unroll
      for (int i = 0; i < 8; ++i) y[i] *= zz[i] * sigmoid_f(zz[i]);
    }
    const long tpos = (long)jb * kFine;
    if (sizeof(T) == 2 && a.out_f32)
      stg_blk8<float>(reinterpret_cast<float*>(a.out) + orow, y, tpos, a.L, a.wide);
    else
      stg_blk8<T>(reinterpret_cast<T*>(a.out) + orow, y, tpos, a.L, a.wide);
    // checkpoints: h at the end of the block
    if (a.xfw) {
      float* xo = a.xfw + (rowg * nbt + jb) * kMaxState;
      if (a.wide) {
        stg256(xo, &h[0]);
        stg256(xo + 8, &h[8]);
      } else {
#pragma unroll
        for (int q = 0; q < 4; ++q)
          reinterpret_cast<float4*>(xo)[q] = make_float4(h[4 * q], h[4 * q + 1], h[4 * q + 2], h[4 * q + 3]);
      }
    }
    if ((jb + 1) % BPC == 0 || jb + 1 == nbt) {
      float4* xo = reinterpret_cast<float4*>(a.x + (rowg * a.nck + jb / BPC) * kMaxState);
#pragma unroll
      for (int q = 0; q < 4; ++q) xo[q] = make_float4(h[4 * q], h[4 * q + 1], h[4 * q + 2], h[4 * q + 3]);
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      dl[i] = ndl[i];
      dlu[i] = ndlu[i];
      y[i] = ny[i];
      if constexpr (kHasZ) zz[i] = nz[i];
    }
  }
}

template <typename T>
cudaError_t launch_scan_bwd_rl(const RlArgs& a, bool has_z, cudaStream_t stream);
template <typename T>
cudaError_t launch_scan_fwd_rl(const RlArgs& a, bool has_z, cudaStream_t stream);

}  // namespace nz
