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
//   * tiles are moved with TMA (cp.async.bulk.tensor, 128B swizzle) through a 2-stage mbarrier
//     pipeline; a generic cooperative loader with the same smem layout serves shapes TMA cannot
//     express (L not a multiple of 128 bytes, unaligned views);
//   * the forward writes h at the end of every chunk (the reference ABI's `x`); the backward walks
//     the chunks last-to-first, recomputes h inside a chunk from that checkpoint, runs the reverse
//     scan for dh, accumulates dA / dD / d(bias) in fp32 and reduces dB / dC over the R rows of the
//     CTA in shared memory before touching global memory.
#pragma once

#include "nz_common.cuh"

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
  static constexpr int CKB = kBwd ? R * kMaxState * 4 : 0;        // checkpoint carry-in (bwd)
  static constexpr int STAGE_TX = NROWT * ROWTILE + 2 * BCTILE;   // bytes TMA delivers (w/o ckpt)
  static constexpr int STAGE = ((STAGE_TX + CKB + 1023) / 1024) * 1024;
  static constexpr int SLAB = kBwd ? R * TL * 4 : 0;              // fp32 [R][TL]
  static constexpr int SMALL = 64 + (kBwd ? 5 : 2) * R * kMaxState * 4;
  static_assert(TL == 256, "chunk length must equal NZ_CHUNK");
  // TMA (hardware swizzle keyed on absolute smem address bits) is only used with R >= 8 rows
  static_assert(R < 8 || ROWTILE % 1024 == 0, "row tiles must keep the 1024-byte swizzle alignment");
  static_assert(SEGB >= 16 && SEGB % 16 == 0, "a lane's segment must be whole 16-byte vectors");
  static constexpr size_t smem_bytes(bool tma) {
    return 1024 + (size_t)(tma ? 2 : 1) * STAGE + 2 * (size_t)SLAB + SMALL;
  }
};

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

// ================================================================================================
// Forward
// ================================================================================================
template <typename T, int M, int LPR, int WARPS, bool kTMA, bool kHasZ>
__global__ void __launch_bounds__(WARPS * 32) scan_fwd_kernel(const __grid_constant__ ScanKArgs a) {
  using Cfg = ScanCfg<T, M, LPR, WARPS, kHasZ, false>;
  constexpr int R = Cfg::R, TL = Cfg::TL, ROWB = Cfg::ROWB, SEGB = Cfg::SEGB;
  constexpr int ROWTILE = Cfg::ROWTILE, BCTILE = Cfg::BCTILE, NROWT = Cfg::NROWT, STAGE = Cfg::STAGE;
  constexpr int NSTAGE = kTMA ? 2 : 1;

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* tail = smem + NSTAGE * STAGE;
  uint64_t* bars = reinterpret_cast<uint64_t*>(tail);
  float* sm_A2 = reinterpret_cast<float*>(tail + 64);
  float* sm_h = sm_A2 + R * kMaxState;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int sl = lane % LPR;                     // segment (time) index inside the row
  const int rloc = warp * Cfg::RPW + lane / LPR;  // row inside the CTA
  const int bpg = a.dpg / R;
  const int rb = blockIdx.x % bpg;
  const int g = (blockIdx.x / bpg) % a.ngroups;
  const int b = blockIdx.x / (bpg * a.ngroups);
  const int d0 = g * a.dpg + rb * R;
  const int d = d0 + rloc;
  const int N = a.dstate;

  for (int i = tid; i < R * kMaxState; i += WARPS * 32) {
    const int r = i / kMaxState, n = i % kMaxState;
    sm_A2[i] = n < N ? a.A[(long)(d0 + r) * a.A_ds + n] * kLog2e : 0.f;
    sm_h[i] = 0.f;
  }

  auto issue = [&](int c, int s) {
    uint8_t* st = smem + s * STAGE;
    mbar_arrive_expect_tx(&bars[s], Cfg::STAGE_TX);
    const int c1 = c * (TL / Cfg::INNER);
    tma_load_4d(st, &a.tm_u, &bars[s], 0, c1, d0, b);
    tma_load_4d(st + ROWTILE, &a.tm_delta, &bars[s], 0, c1, d0, b);
    if (kHasZ) tma_load_4d(st + 2 * ROWTILE, &a.tm_z, &bars[s], 0, c1, d0, b);
    tma_load_5d(st + NROWT * ROWTILE, &a.tm_B, &bars[s], 0, c1, 0, g, b);
    tma_load_5d(st + NROWT * ROWTILE + BCTILE, &a.tm_C, &bars[s], 0, c1, 0, g, b);
  };

  if (kTMA && tid == 0) {
    mbar_init(&bars[0], 1);
    mbar_init(&bars[1], 1);
    fence_mbar_init();
  }
  __syncthreads();
  if (kTMA && tid == 0) {
    issue(0, 0);
    if (a.nchunks > 1) issue(1, 1);
  }

  const float Dv = a.D ? a.D[d] : 0.f;
  const float bias = a.bias ? a.bias[d] : 0.f;
  T* outrow = reinterpret_cast<T*>(a.out) + (long)b * a.o_bs + (long)d * a.o_ds;
  float* xrow = a.x + ((long)b * a.dim + d) * (long)a.nchunks * N;
  const uint32_t rowoff = rloc * ROWB, segoff = sl * SEGB;

  for (int c = 0; c < a.nchunks; ++c) {
    const int s = kTMA ? (c & 1) : 0;
    uint8_t* st = smem + s * STAGE;
    if (kTMA) {
      mbar_wait(&bars[s], (c >> 1) & 1);
    } else {
      __syncthreads();
      const long ts = (long)c * TL;
      coop_fill<T, TL>(st, reinterpret_cast<const T*>(a.u) + (long)b * a.u_bs + (long)d0 * a.u_ds, a.u_ds, R, R, ts, a.L);
      coop_fill<T, TL>(st + ROWTILE, reinterpret_cast<const T*>(a.delta) + (long)b * a.dl_bs + (long)d0 * a.dl_ds,
                       a.dl_ds, R, R, ts, a.L);
      if (kHasZ)
        coop_fill<T, TL>(st + 2 * ROWTILE, reinterpret_cast<const T*>(a.z) + (long)b * a.z_bs + (long)d0 * a.z_ds,
                         a.z_ds, R, R, ts, a.L);
      coop_fill<T, TL>(st + NROWT * ROWTILE, reinterpret_cast<const T*>(a.B) + (long)b * a.B_bs + (long)g * a.B_gs,
                       a.B_ns, N, kMaxState, ts, a.L);
      coop_fill<T, TL>(st + NROWT * ROWTILE + BCTILE,
                       reinterpret_cast<const T*>(a.C) + (long)b * a.C_bs + (long)g * a.C_gs, a.C_ns, N, kMaxState, ts,
                       a.L);
      __syncthreads();
    }
    const uint8_t* tB = st + NROWT * ROWTILE;
    const uint8_t* tC = tB + BCTILE;
    const long t0 = (long)c * TL + sl * M;

    float dlu[M], dl[M], y[M];
    lds_items<T, M>(st, rowoff, segoff, dlu);
    lds_items<T, M>(st + ROWTILE, rowoff, segoff, dl);
#pragma unroll
    for (int i = 0; i < M; ++i) {
      float x = dl[i] + bias;
      if (a.softplus) x = softplus_f(x);
      if (t0 + i >= a.L) x = 0.f;  // beyond the sequence: a = 1, b = 0 -> state passes through
      dl[i] = x;
      y[i] = Dv * dlu[i];
      dlu[i] = x * dlu[i];
    }

#pragma unroll 1
    for (int n = 0; n < N; ++n) {
      const float A2 = sm_A2[rloc * kMaxState + n];
      const float hc = sm_h[rloc * kMaxState + n];
      float av[M], bv[M];
      lds_items<T, M>(tB, n * ROWB, segoff, bv);
      float P = 1.f, H = 0.f;
#pragma unroll
      for (int i = 0; i < M; ++i) {
        av[i] = ex2_approx(dl[i] * A2);
        bv[i] = dlu[i] * bv[i];
        H = fmaf(av[i], H, bv[i]);
        P *= av[i];
      }
#pragma unroll
      for (int off = 1; off < LPR; off <<= 1) {
        const float Pp = __shfl_up_sync(0xffffffffu, P, off, LPR);
        const float Hp = __shfl_up_sync(0xffffffffu, H, off, LPR);
        if (sl >= off) {
          H = fmaf(P, Hp, H);
          P *= Pp;
        }
      }
      float Pex = __shfl_up_sync(0xffffffffu, P, 1, LPR);
      float Hex = __shfl_up_sync(0xffffffffu, H, 1, LPR);
      if (sl == 0) {
        Pex = 1.f;
        Hex = 0.f;
      }
      float h = fmaf(Pex, hc, Hex);
      const float Pt = __shfl_sync(0xffffffffu, P, LPR - 1, LPR);
      const float Ht = __shfl_sync(0xffffffffu, H, LPR - 1, LPR);
      const float hnew = fmaf(Pt, hc, Ht);
      float cv[M];
      lds_items<T, M>(tC, n * ROWB, segoff, cv);
#pragma unroll
      for (int i = 0; i < M; ++i) {
        h = fmaf(av[i], h, bv[i]);
        y[i] = fmaf(cv[i], h, y[i]);
      }
      __syncwarp();
      if (sl == LPR - 1) {
        sm_h[rloc * kMaxState + n] = hnew;
        xrow[(long)c * N + n] = hnew;
      }
    }

    if (kHasZ) {
      float zz[M];
      lds_items<T, M>(st + 2 * ROWTILE, rowoff, segoff, zz);
#pragma unroll
      for (int i = 0; i < M; ++i) y[i] *= zz[i] * sigmoid_f(zz[i]);
    }
    stg_items<T, M>(outrow, y, t0, a.L, a.vec_out != 0);

    if (kTMA) {
      __syncthreads();  // every warp is done reading stage s
      if (tid == 0 && c + 2 < a.nchunks) issue(c + 2, s);
    }
  }
}

// ================================================================================================
// Backward
// ================================================================================================
template <typename T, int M, int LPR, int WARPS, bool kTMA, bool kHasZ>
__global__ void __launch_bounds__(WARPS * 32) scan_bwd_kernel(const __grid_constant__ ScanKArgs a) {
  using Cfg = ScanCfg<T, M, LPR, WARPS, kHasZ, true>;
  constexpr int R = Cfg::R, TL = Cfg::TL, ROWB = Cfg::ROWB, SEGB = Cfg::SEGB;
  constexpr int ROWTILE = Cfg::ROWTILE, BCTILE = Cfg::BCTILE, NROWT = Cfg::NROWT, STAGE = Cfg::STAGE;
  constexpr int NSTAGE = kTMA ? 2 : 1;
  constexpr int NT = WARPS * 32;
  constexpr int SLROW = TL * 4;  // bytes of one fp32 slab row

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* slabB = smem + NSTAGE * STAGE;
  uint8_t* slabC = slabB + Cfg::SLAB;
  uint8_t* tail = slabC + Cfg::SLAB;
  uint64_t* bars = reinterpret_cast<uint64_t*>(tail);
  float* sm_A2 = reinterpret_cast<float*>(tail + 64);
  float* sm_dhc = sm_A2 + R * kMaxState;    // dh carried in from the (later) chunk processed before
  float* sm_anx = sm_dhc + R * kMaxState;   // a of the first step of that later chunk
  float* sm_dA = sm_anx + R * kMaxState;    // dA accumulator over chunks
  float* sm_hcg = sm_dA + R * kMaxState;    // generic path: checkpoint carry-in

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
    sm_dhc[i] = 0.f;
    sm_anx[i] = 0.f;
    sm_dA[i] = 0.f;
    sm_hcg[i] = 0.f;
  }

  auto issue = [&](int c, int s) {
    uint8_t* st = smem + s * STAGE;
    mbar_arrive_expect_tx(&bars[s], Cfg::STAGE_TX + (c > 0 ? R * kMaxState * 4 : 0));
    const int c1 = c * (TL / Cfg::INNER);
    tma_load_4d(st, &a.tm_u, &bars[s], 0, c1, d0, b);
    tma_load_4d(st + ROWTILE, &a.tm_delta, &bars[s], 0, c1, d0, b);
    tma_load_4d(st + 2 * ROWTILE, &a.tm_dout, &bars[s], 0, c1, d0, b);
    if (kHasZ) tma_load_4d(st + 3 * ROWTILE, &a.tm_z, &bars[s], 0, c1, d0, b);
    tma_load_5d(st + NROWT * ROWTILE, &a.tm_B, &bars[s], 0, c1, 0, g, b);
    tma_load_5d(st + NROWT * ROWTILE + BCTILE, &a.tm_C, &bars[s], 0, c1, 0, g, b);
    if (c > 0) {  // state at the end of chunk c-1 == carry into chunk c (TMA path has N == 16)
      float* ck = reinterpret_cast<float*>(st + Cfg::STAGE_TX);
      for (int r = 0; r < R; ++r)
        bulk_load_1d(ck + r * kMaxState,
                     a.x + (((long)b * a.dim + d0 + r) * a.nchunks + (c - 1)) * kMaxState, kMaxState * 4, &bars[s]);
    }
  };

  if (kTMA && tid == 0) {
    mbar_init(&bars[0], 1);
    mbar_init(&bars[1], 1);
    fence_mbar_init();
  }
  __syncthreads();
  if (kTMA && tid == 0) {
    issue(a.nchunks - 1, 0);
    if (a.nchunks > 1) issue(a.nchunks - 2, 1);
  }

  const float Dv = a.D ? a.D[d] : 0.f;
  const float bias = a.bias ? a.bias[d] : 0.f;
  const long rowlin = ((long)b * a.dim + d) * a.L;  // gradients are contiguous (batch, dim, L)
  T* durow = reinterpret_cast<T*>(a.du) + rowlin;
  T* ddrow = reinterpret_cast<T*>(a.ddelta) + rowlin;
  T* dzrow = kHasZ ? reinterpret_cast<T*>(a.dz) + rowlin : nullptr;
  const uint32_t rowoff = rloc * ROWB, segoff = sl * SEGB;
  const uint32_t slab_off = rloc * SLROW + sl * (M * 4);
  float dD_acc = 0.f, db_acc = 0.f;

  for (int it = 0; it < a.nchunks; ++it) {
    const int c = a.nchunks - 1 - it;
    const int s = kTMA ? (it & 1) : 0;
    uint8_t* st = smem + s * STAGE;
    const float* ck;
    if (kTMA) {
      mbar_wait(&bars[s], (it >> 1) & 1);
      ck = reinterpret_cast<const float*>(st + Cfg::STAGE_TX);
    } else {
      __syncthreads();
      const long ts = (long)c * TL;
      coop_fill<T, TL>(st, reinterpret_cast<const T*>(a.u) + (long)b * a.u_bs + (long)d0 * a.u_ds, a.u_ds, R, R, ts, a.L);
      coop_fill<T, TL>(st + ROWTILE, reinterpret_cast<const T*>(a.delta) + (long)b * a.dl_bs + (long)d0 * a.dl_ds,
                       a.dl_ds, R, R, ts, a.L);
      coop_fill<T, TL>(st + 2 * ROWTILE, reinterpret_cast<const T*>(a.dout) + (long)b * a.do_bs + (long)d0 * a.do_ds,
                       a.do_ds, R, R, ts, a.L);
      if (kHasZ)
        coop_fill<T, TL>(st + 3 * ROWTILE, reinterpret_cast<const T*>(a.z) + (long)b * a.z_bs + (long)d0 * a.z_ds,
                         a.z_ds, R, R, ts, a.L);
      coop_fill<T, TL>(st + NROWT * ROWTILE, reinterpret_cast<const T*>(a.B) + (long)b * a.B_bs + (long)g * a.B_gs,
                       a.B_ns, N, kMaxState, ts, a.L);
      coop_fill<T, TL>(st + NROWT * ROWTILE + BCTILE,
                       reinterpret_cast<const T*>(a.C) + (long)b * a.C_bs + (long)g * a.C_gs, a.C_ns, N, kMaxState, ts,
                       a.L);
      if (c > 0)
        for (int i = tid; i < R * kMaxState; i += NT) {
          const int r = i / kMaxState, n = i % kMaxState;
          sm_hcg[i] = n < N ? a.x[(((long)b * a.dim + d0 + r) * a.nchunks + (c - 1)) * N + n] : 0.f;
        }
      __syncthreads();
      ck = sm_hcg;
    }
    const uint8_t* tB = st + NROWT * ROWTILE;
    const uint8_t* tC = tB + BCTILE;
    const long t0 = (long)c * TL + sl * M;

    float uu[M], dl[M], dy[M], dlu[M], sB[M], ddl[M];
    float yv[kHasZ ? M : 1], dzf[kHasZ ? M : 1];
    lds_items<T, M>(st, rowoff, segoff, uu);
    lds_items<T, M>(st + ROWTILE, rowoff, segoff, dl);
    lds_items<T, M>(st + 2 * ROWTILE, rowoff, segoff, dy);
    if (kHasZ) {
      float zz[M];
      lds_items<T, M>(st + 3 * ROWTILE, rowoff, segoff, zz);
#pragma unroll
      for (int i = 0; i < M; ++i) {
        const float sg = sigmoid_f(zz[i]);
        dzf[i] = dy[i] * sg * (1.f + zz[i] * (1.f - sg));  // dout * d silu(z)/dz
        dy[i] = dy[i] * zz[i] * sg;                         // dout * silu(z)
        yv[i] = Dv * uu[i];
      }
    }
#pragma unroll
    for (int i = 0; i < M; ++i) {
      float x = dl[i] + bias;
      if (a.softplus) x = softplus_f(x);
      if (t0 + i >= a.L) {
        x = 0.f;
        dy[i] = 0.f;
      }
      dl[i] = x;
      dlu[i] = x * uu[i];
      sB[i] = 0.f;
      ddl[i] = 0.f;
    }

#pragma unroll 1
    for (int n = 0; n < N; ++n) {
      const float A2 = sm_A2[rloc * kMaxState + n];
      const float An = A2 * kLn2;
      const float hc = c > 0 ? ck[rloc * kMaxState + n] : 0.f;
      const float dhc = sm_dhc[rloc * kMaxState + n];
      const float anx = sm_anx[rloc * kMaxState + n];
      float av[M], ah[M], bv[M], cv[M], tmp[M];
      lds_items<T, M>(tB, n * ROWB, segoff, bv);
      lds_items<T, M>(tC, n * ROWB, segoff, cv);

      // ---- forward inside the chunk: fold, scan across lanes, then replay with the true carry ----
      float P = 1.f, H = 0.f;
#pragma unroll
      for (int i = 0; i < M; ++i) {
        av[i] = ex2_approx(dl[i] * A2);
        H = fmaf(av[i], H, dlu[i] * bv[i]);
        P *= av[i];
      }
#pragma unroll
      for (int off = 1; off < LPR; off <<= 1) {
        const float Pp = __shfl_up_sync(0xffffffffu, P, off, LPR);
        const float Hp = __shfl_up_sync(0xffffffffu, H, off, LPR);
        if (sl >= off) {
          H = fmaf(P, Hp, H);
          P *= Pp;
        }
      }
      float Pex = __shfl_up_sync(0xffffffffu, P, 1, LPR);
      float Hex = __shfl_up_sync(0xffffffffu, H, 1, LPR);
      if (sl == 0) {
        Pex = 1.f;
        Hex = 0.f;
      }
      float h = fmaf(Pex, hc, Hex);
#pragma unroll
      for (int i = 0; i < M; ++i) {
        ah[i] = av[i] * h;                    // a_t * h_{t-1}
        h = fmaf(dlu[i], bv[i], ah[i]);       // h_t
        tmp[i] = dy[i] * h;                   // dC_t[n] contribution of this row
        if (kHasZ) yv[i] = fmaf(cv[i], h, yv[i]);
        cv[i] *= dy[i];                       // from here on cv holds C_t[n] * dy_t
      }
#pragma unroll
      for (int j = 0; j < M / 4; ++j)
        *reinterpret_cast<float4*>(slabC + swz128(slab_off + 16 * j)) =
            make_float4(tmp[4 * j], tmp[4 * j + 1], tmp[4 * j + 2], tmp[4 * j + 3]);

      // ---- reverse inside the chunk: dh_t = a_{t+1} dh_{t+1} + C_t dy_t ----
      float anl = __shfl_down_sync(0xffffffffu, av[0], 1, LPR);
      if (sl == LPR - 1) anl = anx;
      float Q = 1.f, G = 0.f;
#pragma unroll
      for (int i = M - 1; i >= 0; --i) {
        const float aup = (i == M - 1) ? anl : av[i + 1];
        G = fmaf(aup, G, cv[i]);
        Q *= aup;
      }
#pragma unroll
      for (int off = 1; off < LPR; off <<= 1) {
        const float Qn = __shfl_down_sync(0xffffffffu, Q, off, LPR);
        const float Gn = __shfl_down_sync(0xffffffffu, G, off, LPR);
        if (sl + off < LPR) {
          G = fmaf(Q, Gn, G);
          Q *= Qn;
        }
      }
      float Qex = __shfl_down_sync(0xffffffffu, Q, 1, LPR);
      float Gex = __shfl_down_sync(0xffffffffu, G, 1, LPR);
      if (sl == LPR - 1) {
        Qex = 1.f;
        Gex = 0.f;
      }
      float dh = fmaf(Qex, dhc, Gex);
      const float Q0 = __shfl_sync(0xffffffffu, Q, 0, LPR);
      const float G0 = __shfl_sync(0xffffffffu, G, 0, LPR);
      const float a0 = __shfl_sync(0xffffffffu, av[0], 0, LPR);
      const float dhnew = fmaf(Q0, dhc, G0);
      float gs = 0.f;
#pragma unroll
      for (int i = M - 1; i >= 0; --i) {
        const float aup = (i == M - 1) ? anl : av[i + 1];
        dh = fmaf(aup, dh, cv[i]);
        sB[i] = fmaf(dh, bv[i], sB[i]);
        const float gq = dh * ah[i];
        ddl[i] = fmaf(An, gq, ddl[i]);
        gs = fmaf(dl[i], gq, gs);
        tmp[i] = dh * dlu[i];                 // dB_t[n] contribution of this row
      }
#pragma unroll
      for (int j = 0; j < M / 4; ++j)
        *reinterpret_cast<float4*>(slabB + swz128(slab_off + 16 * j)) =
            make_float4(tmp[4 * j], tmp[4 * j + 1], tmp[4 * j + 2], tmp[4 * j + 3]);
#pragma unroll
      for (int off = LPR / 2; off > 0; off >>= 1) gs += __shfl_xor_sync(0xffffffffu, gs, off, LPR);
      __syncwarp();
      if (sl == 0) {
        sm_dA[rloc * kMaxState + n] += gs;
        sm_dhc[rloc * kMaxState + n] = dhnew;
        sm_anx[rloc * kMaxState + n] = a0;
      }

      // ---- reduce dB / dC over the R rows of this CTA, then one global update per (n, t) ----
      __syncthreads();
      {
        constexpr int NV = TL / 4;  // float4 outputs per array
        for (int k = tid; k < 2 * NV; k += NT) {
          const int arr = k / NV, t4 = k - arr * NV;
          const uint8_t* slab = arr ? slabC : slabB;
          float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
          for (int r = 0; r < R; ++r) {
            const float4 v = *reinterpret_cast<const float4*>(slab + swz128(r * SLROW + t4 * 16));
            acc.x += v.x;
            acc.y += v.y;
            acc.z += v.z;
            acc.w += v.w;
          }
          const long tg = (long)c * TL + t4 * 4;
          float* dst = (arr ? a.dC : a.dB) + (((long)b * a.ngroups + g) * N + n) * a.L + tg;
          if (a.vec_grad && tg + 4 <= a.L) {
            if (bpg == 1)
              *reinterpret_cast<float4*>(dst) = acc;
            else
              atomicAdd(reinterpret_cast<float4*>(dst), acc);
          } else {
            const float vals[4] = {acc.x, acc.y, acc.z, acc.w};
            for (int q = 0; q < 4; ++q)
              if (tg + q < a.L) {
                if (bpg == 1)
                  dst[q] = vals[q];
                else
                  atomicAdd(dst + q, vals[q]);
              }
          }
        }
      }
      __syncthreads();
    }

    // ---- per-(row, t) epilogue ----
    float outv[M];
#pragma unroll
    for (int i = 0; i < M; ++i) outv[i] = fmaf(dl[i], sB[i], Dv * dy[i]);  // du
    stg_items<T, M>(durow, outv, t0, a.L, a.vec_grad != 0);
#pragma unroll
    for (int i = 0; i < M; ++i) {
      float gd = fmaf(uu[i], sB[i], ddl[i]);  // d loss / d dl
      // d softplus(x)/dx = sigmoid(x) = 1 - exp(-softplus(x)) = -expm1(-dl)
      if (a.softplus) gd *= -expm1f(-dl[i]);
      if (t0 + i >= a.L) gd = 0.f;
      outv[i] = gd;
      db_acc += gd;
      dD_acc = fmaf(dy[i], uu[i], dD_acc);
    }
    stg_items<T, M>(ddrow, outv, t0, a.L, a.vec_grad != 0);
    if (kHasZ) {
#pragma unroll
      for (int i = 0; i < M; ++i) outv[i] = dzf[i] * yv[i];
      stg_items<T, M>(dzrow, outv, t0, a.L, a.vec_grad != 0);
    }

    if (kTMA) {
      // the trailing __syncthreads of the n-loop already fenced all reads of stage s
      if (tid == 0 && it + 2 < a.nchunks) issue(c - 2, s);
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
