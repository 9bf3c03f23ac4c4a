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
// Parallel decomposition (DESIGN.md "Kernels"; the cost model behind it is tools/ubench/pipes.cu):
//   * a TILE is R rows (b, d) of ONE (batch, group) times TL consecutive time steps.  Tiles are
//     handed out by an atomic ticket counter to persistent CTAs in chunk-major order (all row blocks
//     of chunk 0, then chunk 1, ...; the backward walks the chunks last to first), so the machine is
//     full for every shape -- rows x L, not rows, is the unit of parallelism;
//   * the state a tile needs from its predecessor along L (h for the forward, dh for the backward)
//     travels through L2 in 64-bit {value, tag} slots written with one relaxed 8-byte store and
//     polled with relaxed 8-byte loads (no fences: value and tag are a single-copy-atomic word);
//     the hand-off is per state, so consecutive chunks of a row run concurrently, one state apart;
//   * inside a tile one warp owns RPW = 32/LPR rows; lane = segment * RPW + row, i.e. the lanes that
//     hold the same M-step time segment of different rows are ADJACENT: their B_t/C_t reads hit the
//     same 16-byte words, which the shared-memory pipe serves in one wavefront per 128 unique bytes
//     (measured: LDS.128 costs 2 instead of 4 cycles) -- the kernels are bound by that pipe, not by HBM;
//   * per state n a lane folds its M steps sequentially, the LPR lane aggregates (P, H) of
//     h <- P*h + H are combined with a Kogge-Stone scan over warp shuffles (stride RPW), the product
//     of a over a lane's segment is exp2(A2 * sum(dl)) -- one MUFU instead of M multiplies;
//   * tiles are moved with TMA (cp.async.bulk.tensor, 128B swizzle) into single-buffered row tiles
//     (refilled for the next ticket as soon as every lane holds its values in registers) and
//     double-buffered B/C tiles, each on its own mbarrier; a cooperative loader with the same smem
//     layout serves shapes TMA cannot express;
//   * the forward writes h every NZ_CHUNK steps (the reference ABI's `x`); the backward recomputes h
//     inside a tile from that checkpoint, runs the reverse scan for dh interleaved with the forward
//     one, accumulates dA / dD / d(bias) in fp32 and reduces dB / dC first across the rows of a warp
//     (one shuffle per value) and then across the warps of the CTA in shared memory.
#pragma once

#include "../../include/nnuzoo_b200.h"  // NZ_CHUNK
#include "nz_common.cuh"

#ifndef NZ_FWD_UNROLL
#define NZ_FWD_UNROLL 2  // measured: 2.73 vs 2.83 clk/elt/SM (profiles/r01_kernel_tuning.md)
#endif
#ifndef NZ_BWD_UNROLL
#define NZ_BWD_UNROLL 1
#endif
#define NZ_PRAGMA_(x) _Pragma(#x)
#define NZ_UNROLL(n) NZ_PRAGMA_(unroll n)
#ifndef NZ_FWD_SPLIT
#define NZ_FWD_SPLIT 0  // 1: fold / replay a lane's segment as two independent half-chains (measured slower: 2.74 vs 2.60)
#endif
#ifndef NZ_BWD_REDUCE_ALL
#define NZ_BWD_REDUCE_ALL 0  // 1: every thread reduces one dB/dC element per state (balanced) instead of a rotating warp pair
#endif
#ifndef NZ_EXP_NOSLAB
#define NZ_EXP_NOSLAB 0  // timing experiment (wrong results): the backward without the dB/dC slab hand-off and reduction
#endif
#ifndef NZ_EXP_NOHSCAN
#define NZ_EXP_NOHSCAN 0  // timing experiment (wrong results): what the backward would cost without the h fold + scan
#endif
#ifndef NZ_BWD_REDUCE_LATE
#define NZ_BWD_REDUCE_LATE 1  // reduce slab(n-1) 0: after the scans of state n, 1: before writing slab(n) (7.62 -> 7.27), 2: after it
#endif
#ifndef NZ_FWD_FINE_DIRECT
#define NZ_FWD_FINE_DIRECT 0  // fine checkpoints: 1 = one 4-byte store per (block, state), 0 = staged 16-byte vectors
#endif
#ifndef NZ_BWD_KEEPB
#define NZ_BWD_KEEPB 0  // 1: keep B_t[n] in registers instead of re-reading the tile for sum_n dh*B (costs 8 registers)
#endif

namespace nz {

constexpr int kCkpt = NZ_CHUNK;  // the C ABI's checkpoint interval: steps between two saved states h

struct alignas(64) ScanKArgs {
  CUtensorMap tm_u, tm_delta, tm_z, tm_dout, tm_B, tm_C;
  const void *u, *delta, *z, *dout, *B, *C;
  const float *A, *D, *bias;
  void *out, *du, *ddelta, *dz;
  float *x, *dA, *dB, *dC, *dD, *dbias;
  unsigned long long* carry;  // [batch*dim][2][16] {value, tag} slots of the chained hand-off
  unsigned* ticket;           // dynamic tile counter (zeroed by the host before the launch)
  unsigned long long* trace;  // optional (tools only): 8 words per tile {ticket, smid, t_start, t_loop, wait, t_end,
                              // t state-0 carry received, t last-state carry received} in ns
  long L;
  long u_bs, u_ds, dl_bs, dl_ds, z_bs, z_ds, o_bs, o_ds, do_bs, do_ds;
  long B_bs, B_gs, B_ns, C_bs, C_gs, C_ns, A_ds;
  int batch, dim, dstate, ngroups, dpg;
  int nrb;        // row blocks per group
  int nrb_total;  // batch * ngroups * nrb
  int nchunks;    // tiles along L
  int nck;        // checkpoints along L (ceil(L / kCkpt))
  int ntiles;     // nrb_total * nchunks
  int softplus;
  int claim_late;  // 1: claim the next ticket at the END of a tile instead of one tile ahead.  With fewer row
                   // blocks than CTA slots every tile waits on its predecessor; a ticket claimed early then
                   // sits idle in a busy CTA while later tiles of its chain already occupy other CTAs
                   // (priority inversion: the whole chain stalls until that CTA comes round)
  int skew;      // states the later chunk's tile must be ahead before a dependent tile starts polling
  int vec_out;   // out rows 16-byte aligned -> vector stores
  int out_f32;   // 16-bit instantiations only: `out` is fp32 (NzScanDesc::out_f32)
  // chunk-parallel forward for few-rows / long-L launches (capi.cu: run_scan_fwd_cp).  0: chained hand-off (default);
  // 1: aggregate pass -- every tile starts from h = 0, publishes its (prod a, h_end) per (row, state) and skips the
  // replay / read-out; 2: final pass -- the carries come from cp_hin, nothing is polled or published.
  // Arrays are [row][state][chunk] fp32.
  int cp_mode;
  float *cp_P, *cp_H;
  const float* cp_hin;
  int vec_grad;  // du/ddelta/dz and dB/dC rows 16-byte aligned
  float* xf;     // fine checkpoints (batch, dim, nbt, 16): h at the end of every NZ_FINE-step block (kFineCk forwards)
  long nbt;      // fine blocks per row (L / NZ_FINE)
};

template <typename T, int M, int LPR, int WARPS, bool kHasZ, bool kBwd>
struct ScanCfg {
  static constexpr int RPW = 32 / LPR;            // rows per warp
  static constexpr int R = RPW * WARPS;           // rows per CTA
  static constexpr int TL = LPR * M;              // tile length (time steps)
  static constexpr int ES = sizeof(T);
  static constexpr int ROWB = TL * ES;            // bytes of one tile row
  static constexpr int SEGB = M * ES;             // bytes of one lane's segment
  static constexpr int INNER = 128 / ES;          // elements in one 128-byte TMA box row
  static constexpr int ROWTILE = R * ROWB;
  static constexpr int BCTILE = kMaxState * ROWB;
  static constexpr int NROWT = (kBwd ? 3 : 2) + (kHasZ ? 1 : 0);  // u, delta, [dout], [z]
  static constexpr int ROWS_TX = NROWT * ROWTILE;                 // bytes TMA delivers per tile (rows)
  static constexpr int ROWS_REGION = ((ROWS_TX + 1023) / 1024) * 1024;
  static constexpr int BC_TX = 2 * BCTILE;                        // bytes per B/C stage
  static constexpr int SLROW = TL * 4;                            // one fp32 slab row
  static constexpr int SLAB = kBwd ? R * SLROW : 0;               // fp32 [R][TL]: one array of one slab buffer
  static constexpr int NRW = (2 * TL / 4) / 32;                   // warps reducing one state's dB + dC slabs
  static constexpr int GSB = kBwd ? kMaxState * R * (LPR / 2) * 4 : 0;  // dA partials [n][row][segment pair]
  static constexpr int SMALL = 256 + (kBwd ? 3 : 2) * R * kMaxState * 4;
  static_assert(!kBwd || (NRW >= 1 && WARPS % NRW == 0), "reducer warps must tile the CTA");
  static_assert(RPW == 1 || RPW == 2, "one or two rows per warp");
  static_assert(TL % kCkpt == 0, "a tile is a whole number of checkpoint intervals");
  static_assert(!kBwd || TL == kCkpt, "the backward restarts from a checkpoint at every tile");
  static_assert(ROWTILE % 1024 == 0, "row tiles must keep the 1024-byte swizzle alignment");
  static_assert(SEGB >= 16 && SEGB % 16 == 0, "a lane's segment must be whole 16-byte vectors");
  static_assert(M % 2 == 0, "packed fp32x2 math works on time pairs");
  static constexpr int FINE_STAGE = WARPS * (M / NZ_FINE) * 4 * 32 * 4;  // per warp [block][state & 3][lane] fp32
  static constexpr size_t smem_bytes(bool tma, bool fine = false) {
    return 1024 + (size_t)ROWS_REGION + (size_t)(tma ? 2 : 1) * BC_TX + 5 * (size_t)SLAB + (size_t)GSB + SMALL +
           (fine ? (size_t)FINE_STAGE : 0);
  }
};

// ---- packed fp32x2 helpers (FFMA2 / FMUL2 / FADD2 on sm_100) ----
__device__ __forceinline__ float2 f2(float a, float b) { return make_float2(a, b); }
__device__ __forceinline__ float2 mul2(float2 a, float2 b) { return __fmul2_rn(a, b); }
__device__ __forceinline__ float2 fma2(float2 a, float2 b, float2 c) { return __ffma2_rn(a, b, c); }
__device__ __forceinline__ float2 sub2(float2 a, float2 b) { return __fadd2_rn(a, f2(-b.x, -b.y)); }

__device__ __forceinline__ unsigned long long gtime_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
__device__ __forceinline__ unsigned smid() {
  unsigned r;
  asm volatile("mov.u32 %0, %%smid;" : "=r"(r));
  return r;
}

// ---- chained hand-off slots ----
// {value, tag} travel as ONE 64-bit scalar access: the PTX memory model guarantees single-copy atomicity for an aligned
// b64 access, not for a v2.b32 vector access (which it treats as two scalar accesses in unspecified order).
__device__ __forceinline__ void slot_store(unsigned long long* p, float v, unsigned tag) {
  const unsigned long long w = ((unsigned long long)tag << 32) | (unsigned long long)__float_as_uint(v);
  asm volatile("st.relaxed.gpu.global.b64 [%0], %1;" ::"l"(p), "l"(w) : "memory");
}
__device__ __forceinline__ void slot_load(const unsigned long long* p, float& v, unsigned& tag) {
  unsigned long long w;
  asm volatile("ld.relaxed.gpu.global.b64 %0, [%1];" : "=l"(w) : "l"(p) : "memory");
  v = __uint_as_float((unsigned)(w & 0xffffffffull));
  tag = (unsigned)(w >> 32);
}

// Poll a hand-off slot until it carries `expect`.  (Backing off with nanosleep when the slot's current tag shows
// that the producer is several tiles away was measured and rejected: no gain on chain-bound shapes -- polling
// traffic is not what limits them -- and 30 % slower on cfg 1, where the producer is only two tiles away.)
template <bool kDescending>
__device__ __forceinline__ void slot_wait(const unsigned long long* p, unsigned expect, float& v, unsigned& tag) {
  while (tag != expect) slot_load(p, v, tag);
}

struct TileId {
  int b, g, rb, c, d0, rows_valid, pad0, pad1;  // 32 bytes: four of them sit in shared memory next to the ticket ring
};
template <int R, bool kBwd>
__device__ __forceinline__ TileId decode_tile(const ScanKArgs& a, int t) {
  TileId x;
  const int ci = t / a.nrb_total;
  int w = t - ci * a.nrb_total;
  x.c = kBwd ? a.nchunks - 1 - ci : ci;
  x.rb = w % a.nrb;
  w /= a.nrb;
  x.g = w % a.ngroups;
  x.b = w / a.ngroups;
  x.d0 = x.g * a.dpg + x.rb * R;
  x.rows_valid = min(R, a.dpg - x.rb * R);
  return x;
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

// ---- explicit shared-space accesses (32-bit addresses: no generic-pointer arithmetic in the hot loops) ----
__device__ __forceinline__ uint4 lds128(uint32_t a) {
  uint4 v;
  asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(a));
  return v;
}
__device__ __forceinline__ float lds32(uint32_t a) {
  float v;
  asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(a));
  return v;
}
__device__ __forceinline__ void sts128(uint32_t a, float x, float y, float z, float w) {
  asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(a), "f"(x), "f"(y), "f"(z), "f"(w) : "memory");
}
__device__ __forceinline__ void sts32(uint32_t a, float x) {
  asm volatile("st.shared.f32 [%0], %1;" ::"r"(a), "f"(x) : "memory");
}
// Hide a loop-invariant value's recipe from the optimiser: at 128 registers ptxas otherwise
// rematerialises it (address swizzles, the sum of dl) inside every trip of the state loop.
__device__ __forceinline__ uint32_t keep(uint32_t x) {
  asm volatile("" : "+r"(x));
  return x;
}
__device__ __forceinline__ float keep(float x) {
  asm volatile("" : "+f"(x));
  return x;
}

// A lane's M items out of a swizzled dense-row tile.  `pre[j]` = swz128(segoff + 16 j) is the lane's
// swizzled in-row offset of its j-th 16-byte vector (loop invariant); a row adds row*ROWB to the
// address and (row * ROWB/128) & 7 to the swizzle key -- the two never overlap because ROWB is a
// power of two <= 1024, so the key folds into one XOR per row.
template <int M, typename T>
struct LanePre {
  static constexpr int kVec = M * (int)sizeof(T) / 16;
  uint32_t pre[kVec];
  __device__ __forceinline__ explicit LanePre(uint32_t segoff) {
#pragma unroll
    for (int j = 0; j < kVec; ++j) pre[j] = keep(swz128(segoff + 16u * j));
  }
};
template <typename T, int M, int ROWB>
__device__ __forceinline__ void lds_seg(uint32_t tile_s, int row, const LanePre<M, T>& lp, float (&v)[M]) {
  static_assert((ROWB & (ROWB - 1)) == 0 && ROWB >= 128 && ROWB <= 1024, "row pitch must be a power of two in [128, 1024]");
  constexpr int kPer = 16 / (int)sizeof(T);
  const uint32_t rowx = (((uint32_t)row * (ROWB / 128)) & 7u) << 4;
  const uint32_t base = tile_s + (uint32_t)row * ROWB;
#pragma unroll
  for (int j = 0; j < LanePre<M, T>::kVec; ++j) {
    unpack16<T>(lds128(base + (lp.pre[j] ^ rowx)), &v[j * kPer]);
  }
}

// Kogge-Stone steps over the segments of a row; the lanes of one row are RPW apart.
__device__ __forceinline__ void ks_up_w(float& P, float& H, int delta) { ks_up<32>(P, H, delta); }
__device__ __forceinline__ void ks_down_w(float& Q, float& G, int delta) { ks_down<32>(Q, G, delta); }
template <int RPW>
__device__ __forceinline__ float ks_enter_up_w(float P, float H, float carry) {
  float h;
  asm volatile(
      "{\n\t.reg .pred p;\n\t.reg .f32 pp, hp;\n\t"
      "shfl.sync.up.b32 pp|p, %1, %4, 0, 0xffffffff;\n\t"
      "shfl.sync.up.b32 hp, %2, %4, 0, 0xffffffff;\n\t"
      "mov.f32 %0, %3;\n\t"
      "@p fma.rn.f32 %0, pp, %3, hp;\n\t}"
      : "=f"(h)
      : "f"(P), "f"(H), "f"(carry), "n"(RPW));
  return h;
}
template <int RPW>
__device__ __forceinline__ float ks_enter_down_w(float Q, float G, float carry) {
  float h;
  asm volatile(
      "{\n\t.reg .pred p;\n\t.reg .f32 qq, gg;\n\t"
      "shfl.sync.down.b32 qq|p, %1, %4, 0x1f, 0xffffffff;\n\t"
      "shfl.sync.down.b32 gg, %2, %4, 0x1f, 0xffffffff;\n\t"
      "mov.f32 %0, %3;\n\t"
      "@p fma.rn.f32 %0, qq, %3, gg;\n\t}"
      : "=f"(h)
      : "f"(Q), "f"(G), "f"(carry), "n"(RPW));
  return h;
}

// ================================================================================================
// Forward
// ================================================================================================
#ifndef NZ_FWD_MINB
#define NZ_FWD_MINB 2  // resident CTAs per SM the forward is compiled for (register cap 65536 / (256 * MINB))
#endif
template <typename T, int M, int LPR, int WARPS, int NQ, bool kTMA, bool kHasZ, bool kFineCk = false>
__global__ void __launch_bounds__(WARPS * 32, (kHasZ && sizeof(T) == 4 && M * LPR > 128) ? 1 : NZ_FWD_MINB)
    scan_fwd_kernel(const __grid_constant__ ScanKArgs a) {
  static_assert(!kFineCk || (kTMA && M % NZ_FINE == 0), "fine checkpoints: TMA path, whole blocks per lane");
  using Cfg = ScanCfg<T, M, LPR, WARPS, kHasZ, false>;
  constexpr int R = Cfg::R, TL = Cfg::TL, ROWB = Cfg::ROWB, SEGB = Cfg::SEGB, RPW = Cfg::RPW;
  constexpr int ROWTILE = Cfg::ROWTILE, BCTILE = Cfg::BCTILE;
  constexpr int NT = WARPS * 32;
  constexpr int H2 = M / 2;
  constexpr int CKSEG = kCkpt / M;   // lane segments per checkpoint interval
  constexpr int CKPT = TL / kCkpt;   // checkpoints per tile
  constexpr int APT = (R * kMaxState + NT - 1) / NT;

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* rows = smem;
  uint8_t* bcs = smem + Cfg::ROWS_REGION;
  uint8_t* tail = bcs + (kTMA ? 2 : 1) * Cfg::BC_TX;
  uint64_t* bars = reinterpret_cast<uint64_t*>(tail);            // [0] rows, [1],[2] B/C stages
  volatile int* tk = reinterpret_cast<volatile int*>(tail + 32);  // ring of 4 tickets
  TileId* tq = reinterpret_cast<TileId*>(tail + 128);              // ... and their decoded tile coordinates
  float* sm_A2 = reinterpret_cast<float*>(tail + 256);
  float* sm_hin = sm_A2 + R * kMaxState;  // h carried into the tile when the previous chunk had already finished
  const uint32_t rows_s = keep(smem_u32(rows)), bcs_s = rows_s + Cfg::ROWS_REGION;
  const uint32_t a2_s0 = rows_s + (uint32_t)(reinterpret_cast<uint8_t*>(sm_A2) - rows);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  // staging of the fine checkpoints: [warp][block of the lane's segment][state & 3][lane]
  [[maybe_unused]] const uint32_t fine_w =
      a2_s0 + (uint32_t)(Cfg::SMALL - 256) + (uint32_t)(warp * (M / NZ_FINE) * 4 * 32 * 4 + lane * 4);
  const int sl = lane / RPW;                      // segment (time) index inside the row
  const int rloc = warp * RPW + lane % RPW;       // row inside the CTA
  const int N = kTMA ? kMaxState : a.dstate;  // the TMA path only runs d_state == 16
  const uint32_t segoff = sl * SEGB;
  const LanePre<M, T> lp(segoff);
  const int NP = (N + NQ - 1) / NQ * NQ;  // states are processed NQ at a time; rows >= N of the tiles are zero

  auto issue_rows = [&](const TileId& q) {
    mbar_arrive_expect_tx(&bars[0], Cfg::ROWS_TX);
    const int c1 = q.c * (TL / Cfg::INNER);
    tma_load_4d(rows, &a.tm_u, &bars[0], 0, c1, q.d0, q.b);
    tma_load_4d(rows + ROWTILE, &a.tm_delta, &bars[0], 0, c1, q.d0, q.b);
    if (kHasZ) tma_load_4d(rows + 2 * ROWTILE, &a.tm_z, &bars[0], 0, c1, q.d0, q.b);
  };
  auto issue_bc = [&](const TileId& q, int s) {
    uint8_t* st = bcs + s * Cfg::BC_TX;
    mbar_arrive_expect_tx(&bars[1 + s], Cfg::BC_TX);
    const int c1 = q.c * (TL / Cfg::INNER);
    tma_load_5d(st, &a.tm_B, &bars[1 + s], 0, c1, 0, q.g, q.b);
    tma_load_5d(st + BCTILE, &a.tm_C, &bars[1 + s], 0, c1, 0, q.g, q.b);
  };

  if (tid == 0) {
    mbar_init(&bars[0], 1);
    mbar_init(&bars[1], 1);
    mbar_init(&bars[2], 1);
    fence_mbar_init();
    const int t0 = (int)atomicAdd(a.ticket, 1u);
    tk[0] = t0;
    if (t0 < a.ntiles) {
      const TileId q = decode_tile<R, false>(a, t0);
      tq[0] = q;
      if (kTMA) {
        issue_rows(q);
        issue_bc(q, 0);
      }
    }
  }
  __syncthreads();

  for (int k = 0;; ++k) {
    const int t = tk[k & 3];
    if (t >= a.ntiles) break;
    const TileId q = tq[k & 3];  // decoded once by thread 0 when it claimed the ticket
    const int c = q.c, s = kTMA ? (k & 1) : 0;
    const bool row_ok = kTMA || rloc < q.rows_valid;  // the TMA path only runs whole row blocks
    const int d = q.d0 + (row_ok ? rloc : 0);
    const long rowg = (long)q.b * a.dim + d;
    // Claim the NEXT ticket now (one tile ahead, not two: tiles then start in ticket order up to the
    // spread of one tile duration, so a tile's predecessor along L is almost always already running);
    // its B/C stage is free, its row tiles are requested once this tile's rows sit in registers.
    unsigned nxt = 0;
    if (tid == 0 && !a.claim_late) nxt = atomicAdd(a.ticket, 1u);
    unsigned long long tr_start = 0, tr_loop = 0, tr_wait = 0, tr_recv0 = 0, tr_recvN = 0;
    if (a.trace && tid == 0) tr_start = gtime_ns();

    const float Dv = a.D ? __ldg(a.D + d) : 0.f;
    const float bias = a.bias ? __ldg(a.bias + d) : 0.f;
    // per (row, state) tile constants, one per thread: A and -- if the previous chunk's tile has already
    // published it -- the h carry (the common case when row blocks outnumber the CTAs; otherwise the
    // state loop polls per state, see below)
    float a2pre[APT], hpre[APT];
    bool all_in = true;
#pragma unroll
    for (int j = 0; j < APT; ++j) {
      const int i = tid + j * NT, r = i / kMaxState, n = i % kMaxState;
      const bool ok = i < R * kMaxState && r < q.rows_valid && n < N;
      a2pre[j] = ok ? __ldg(a.A + (long)(q.d0 + r) * a.A_ds + n) * kLog2e : 0.f;
      hpre[j] = 0.f;
      if (ok && c > 0) {
        if (a.cp_mode == 0) {
          unsigned tag;
          slot_load(a.carry + ((((long)q.b * a.dim + q.d0 + r) * 2 + ((c - 1) & 1)) * kMaxState + n), hpre[j], tag);
          all_in = all_in && tag == (unsigned)c;
        } else if (a.cp_mode == 2) {
          hpre[j] = __ldg(a.cp_hin + (((long)q.b * a.dim + q.d0 + r) * kMaxState + n) * a.nchunks + c);
        }
      }
    }
    const unsigned long long* cin = a.carry + (rowg * 2 + ((c - 1) & 1)) * kMaxState;
    const bool chained = c > 0 && row_ok && a.cp_mode == 0;
    bool fast;  // every h carry of the tile was already there: no polling in the state loop

    if constexpr (kTMA) {
      mbar_wait(&bars[0], k & 1);
    } else {
      const long ts = (long)c * TL;
      const int rv = q.rows_valid;
      coop_fill<T, TL>(rows, reinterpret_cast<const T*>(a.u) + (long)q.b * a.u_bs + (long)q.d0 * a.u_ds, a.u_ds, rv, R, ts, a.L);
      coop_fill<T, TL>(rows + ROWTILE, reinterpret_cast<const T*>(a.delta) + (long)q.b * a.dl_bs + (long)q.d0 * a.dl_ds,
                       a.dl_ds, rv, R, ts, a.L);
      if (kHasZ)
        coop_fill<T, TL>(rows + 2 * ROWTILE, reinterpret_cast<const T*>(a.z) + (long)q.b * a.z_bs + (long)q.d0 * a.z_ds,
                         a.z_ds, rv, R, ts, a.L);
      coop_fill<T, TL>(bcs, reinterpret_cast<const T*>(a.B) + (long)q.b * a.B_bs + (long)q.g * a.B_gs, a.B_ns, N,
                       kMaxState, ts, a.L);
      coop_fill<T, TL>(bcs + BCTILE, reinterpret_cast<const T*>(a.C) + (long)q.b * a.C_bs + (long)q.g * a.C_gs, a.C_ns, N,
                       kMaxState, ts, a.L);
#pragma unroll
      for (int j = 0; j < APT; ++j)
        if (tid + j * NT < R * kMaxState) {
          sm_A2[tid + j * NT] = a2pre[j];
          sm_hin[tid + j * NT] = hpre[j];
        }
      if (tid == 0 && !a.claim_late) {
        tk[(k + 1) & 3] = (int)nxt;
        if ((int)nxt < a.ntiles) tq[(k + 1) & 3] = decode_tile<R, false>(a, (int)nxt);
      }
      fast = __syncthreads_and(all_in) != 0;
    }
    const long t0 = (long)c * TL + sl * M;
    const int nvalid = (int)max(0L, min((long)M, a.L - t0));  // steps of this lane's segment inside the sequence

    float dlu[M], dl[M], y[M], zz[kHasZ ? M : 1];
    lds_seg<T, M, ROWB>(rows_s, rloc, lp, dlu);
    lds_seg<T, M, ROWB>(rows_s + ROWTILE, rloc, lp, dl);
    if constexpr (kHasZ) lds_seg<T, M, ROWB>(rows_s + 2 * ROWTILE, rloc, lp, zz);
    float dlsum = 0.f;
#pragma unroll
    for (int i = 0; i < M; ++i) {
      float x = dl[i] + bias;
      if (a.softplus) x = softplus_f(x);
      if (i >= nvalid) x = 0.f;  // beyond the sequence: a = 1, b = 0 -> state passes through
      dl[i] = x;
      dlsum += x;
      y[i] = Dv * dlu[i];
      dlu[i] = x * dlu[i];
    }
    if constexpr (kTMA) {
#pragma unroll
      for (int j = 0; j < APT; ++j)
        if (tid + j * NT < R * kMaxState) {
          sm_A2[tid + j * NT] = a2pre[j];
          sm_hin[tid + j * NT] = hpre[j];
        }
      // the row values now live in registers: request the next ticket's tiles (rows + the free B/C stage)
      fast = __syncthreads_and(all_in) != 0;
      if (tid == 0 && !a.claim_late) {
        tk[(k + 1) & 3] = (int)nxt;
        if ((int)nxt < a.ntiles) {
          const TileId qn = decode_tile<R, false>(a, (int)nxt);
          tq[(k + 1) & 3] = qn;
          issue_rows(qn);
          issue_bc(qn, (k + 1) & 1);
        }
      }
      mbar_wait(&bars[1 + s], (k >> 1) & 1);
    }
    const uint32_t tB = bcs_s + s * Cfg::BC_TX, tC = tB + BCTILE;
    dlsum = keep(dlsum);
    [[maybe_unused]] float dlsum_a = 0.f, dlsum_b = 0.f;
#if NZ_FWD_SPLIT
#pragma unroll
    for (int i = 0; i < M / 2; ++i) {
      dlsum_a += dl[i];
      dlsum_b += dl[i + M / 2];
    }
    dlsum_a = keep(dlsum_a);
    dlsum_b = keep(dlsum_b);
#endif
    T* outrow = reinterpret_cast<T*>(a.out) + (long)q.b * a.o_bs + (long)d * a.o_ds;
    float* xrow = a.x + rowg * (long)a.nck * N;

    // when the previous chunk's tile is still running: start once it is a few states ahead, so that the
    // one-state-ahead prefetch of the per-state polling finds its data
    float cv_next = 0.f;
    unsigned ct_next = 0;
    if (!fast && chained) {
      if (a.skew >= 0) {
        const int nsk = min(a.skew, N - 1);
        unsigned tg;
        float dummy;
        do {
          slot_load(cin + nsk, dummy, tg);
        } while (tg != (unsigned)c);
      }
      slot_load(cin, cv_next, ct_next);
    }

    if (a.trace && tid == 0) tr_loop = gtime_ns();
    // per-state pointers advance by NQ per trip (keeps 64-bit address arithmetic out of the loop body)
    const unsigned long long* cin_p = cin;
    const int coff = (c & 1) ? kMaxState : -kMaxState;  // the outgoing ring slot sits 16 slots after / before the incoming one
    uint32_t a2_p = a2_s0 + rloc * (kMaxState * 4);    // sm_hin follows at a fixed distance
    const int my_ck = c * CKPT + (sl + 1) / CKSEG - 1;  // checkpoint this lane's segment end belongs to (if any)
    const bool ck_lane = row_ok && (sl + 1) % CKSEG == 0 && my_ck < a.nck;
    float* x_p = xrow + (long)my_ck * N;
    NZ_UNROLL(NZ_FWD_UNROLL)
    for (int n = 0; n < NP; n += NQ, cin_p += NQ, a2_p += NQ * 4, x_p += NQ) {  // NQ states per trip
      float hc[NQ], P[NQ], H[NQ];
      [[maybe_unused]] float Pa[NQ], Pb[NQ];
      unsigned ctag[NQ];
      float av[NQ][M], bv[NQ][M];
      if (fast) {
#pragma unroll
        for (int qi = 0; qi < NQ; ++qi) {
          hc[qi] = lds32(a2_p + (R * kMaxState + qi) * 4);
          ctag[qi] = (unsigned)c;
        }
      } else {
        // state n was prefetched during the previous trip; fetch the others / the next one now
        hc[0] = cv_next;
        ctag[0] = ct_next;
#pragma unroll
        for (int qi = 1; qi < NQ; ++qi) {
          hc[qi] = 0.f;
          ctag[qi] = 0;
          if (chained) slot_load(cin_p + qi, hc[qi], ctag[qi]);
        }
        if (chained && n + NQ < NP) slot_load(cin_p + NQ, cv_next, ct_next);
      }
#pragma unroll
      for (int qi = 0; qi < NQ; ++qi) {
        const float A2 = lds32(a2_p + qi * 4);
#if NZ_FWD_SPLIT
        // the segment is folded / replayed as two independent half-chains (halves the serial FFMA latency)
        Pa[qi] = ex2_approx(A2 * dlsum_a);  // product of a over the first / second half of the segment
        P[qi] = Pa[qi] * ex2_approx(A2 * dlsum_b);
        Pb[qi] = ex2_approx(A2 * dlsum_b);
#else
        P[qi] = ex2_approx(A2 * dlsum);  // product of a over this lane's segment
#endif
        lds_seg<T, M, ROWB>(tB, n + qi, lp, bv[qi]);
#pragma unroll
        for (int kk = 0; kk < H2; ++kk) {
          const float2 x2 = mul2(f2(dl[2 * kk], dl[2 * kk + 1]), f2(A2, A2));
          av[qi][2 * kk] = ex2_approx(x2.x);
          av[qi][2 * kk + 1] = ex2_approx(x2.y);
          const float2 b2 = mul2(f2(dlu[2 * kk], dlu[2 * kk + 1]), f2(bv[qi][2 * kk], bv[qi][2 * kk + 1]));
          bv[qi][2 * kk] = b2.x;
          bv[qi][2 * kk + 1] = b2.y;
        }
        H[qi] = 0.f;
      }
#if NZ_FWD_SPLIT
      float Ha[NQ];
#pragma unroll
      for (int qi = 0; qi < NQ; ++qi) Ha[qi] = 0.f;
#pragma unroll
      for (int i = 0; i < M / 2; ++i) {
#pragma unroll
        for (int qi = 0; qi < NQ; ++qi) {
          Ha[qi] = fmaf(av[qi][i], Ha[qi], bv[qi][i]);
          H[qi] = fmaf(av[qi][i + M / 2], H[qi], bv[qi][i + M / 2]);
        }
      }
#pragma unroll
      for (int qi = 0; qi < NQ; ++qi) H[qi] = fmaf(Pb[qi], Ha[qi], H[qi]);  // aggregate of the whole segment
#else
#pragma unroll
      for (int i = 0; i < M; ++i) {
#pragma unroll
        for (int qi = 0; qi < NQ; ++qi) H[qi] = fmaf(av[qi][i], H[qi], bv[qi][i]);
      }
#endif
#pragma unroll
      for (int off = 1; off < LPR; off <<= 1) {
#pragma unroll
        for (int qi = 0; qi < NQ; ++qi) ks_up_w(P[qi], H[qi], off * RPW);
      }
      // the chained carry must have arrived by now
      if (!fast && chained) {
        const unsigned long long w0 = (a.trace && tid == 0) ? gtime_ns() : 0;
#pragma unroll
        for (int qi = 0; qi < NQ; ++qi) {
          slot_wait<false>(cin_p + qi, (unsigned)c, hc[qi], ctag[qi]);
        }
        if (a.trace && tid == 0) {
          const unsigned long long w1 = gtime_ns();
          tr_wait += w1 - w0;
          if (n == 0) tr_recv0 = w1;
          tr_recvN = w1;
        }
      }
      float h[NQ];
#pragma unroll
      for (int qi = 0; qi < NQ; ++qi) {
        h[qi] = ks_enter_up_w<RPW>(P[qi], H[qi], hc[qi]);
        const float hend = fmaf(P[qi], hc[qi], H[qi]);  // state at the end of this lane's segment
        if (sl == LPR - 1 && row_ok) {
          if (a.cp_mode == 0) {
            slot_store(const_cast<unsigned long long*>(cin_p) + coff + qi, hend, (unsigned)c + 1u);
          } else if (a.cp_mode == 1 && n + qi < N) {
            const long ix = (rowg * kMaxState + n + qi) * a.nchunks + c;
            a.cp_P[ix] = P[qi];   // inclusive aggregate of the last segment = the whole tile
            a.cp_H[ix] = H[qi];
          }
        }
        if (ck_lane && n + qi < N && a.cp_mode != 1) x_p[qi] = hend;
      }
      if (a.cp_mode == 1) continue;  // aggregate pass: no replay, no read-out
#if NZ_FWD_SPLIT
#pragma unroll
      for (int qi = 0; qi < NQ; ++qi) Ha[qi] = fmaf(Pa[qi], h[qi], Ha[qi]);  // state in the middle of the segment
#pragma unroll
      for (int i = 0; i < M / 2; ++i) {  // replay both halves with their true carry-ins; bv is overwritten by h_t
#pragma unroll
        for (int qi = 0; qi < NQ; ++qi) {
          h[qi] = fmaf(av[qi][i], h[qi], bv[qi][i]);
          bv[qi][i] = h[qi];
          Ha[qi] = fmaf(av[qi][i + M / 2], Ha[qi], bv[qi][i + M / 2]);
          bv[qi][i + M / 2] = Ha[qi];
        }
#else
#pragma unroll
      for (int i = 0; i < M; ++i) {  // replay with the true carry-in; bv is overwritten by h_t
#pragma unroll
        for (int qi = 0; qi < NQ; ++qi) {
          h[qi] = fmaf(av[qi][i], h[qi], bv[qi][i]);
          bv[qi][i] = h[qi];
        }
#endif
      }
      if constexpr (kFineCk) {
        // h at the end of each 8-step block of the lane's segment -> xf[row][block][state]
#if NZ_FWD_FINE_DIRECT
        // one 4-byte store per (block, state): the 16 states of a block fill two sectors over 16 trips (merged in L2)
#pragma unroll
        for (int qi = 0; qi < NQ; ++qi) {
#pragma unroll
          for (int bk = 0; bk < M / NZ_FINE; ++bk) {
            const long blkg = (long)c * (TL / NZ_FINE) + sl * (M / NZ_FINE) + bk;
            if (blkg < a.nbt) a.xf[(rowg * a.nbt + blkg) * kMaxState + n + qi] = bv[qi][bk * NZ_FINE + NZ_FINE - 1];
          }
        }
#else
        // four states are gathered through a lane-private staging column so that the store is one 16-byte vector per block
#pragma unroll
        for (int qi = 0; qi < NQ; ++qi) {
          const int nn = n + qi;
          const uint32_t fs = fine_w + (uint32_t)((nn & 3) << 7);
#pragma unroll
          for (int bk = 0; bk < M / NZ_FINE; ++bk) sts32(fs + bk * 512, bv[qi][bk * NZ_FINE + NZ_FINE - 1]);
          if ((nn & 3) == 3) {
#pragma unroll
            for (int bk = 0; bk < M / NZ_FINE; ++bk) {
              const long blkg = (long)c * (TL / NZ_FINE) + sl * (M / NZ_FINE) + bk;
              const float4 v = make_float4(lds32(fine_w + bk * 512), lds32(fine_w + bk * 512 + 128),
                                           lds32(fine_w + bk * 512 + 256), lds32(fine_w + bk * 512 + 384));
              if (blkg < a.nbt) *reinterpret_cast<float4*>(a.xf + (rowg * a.nbt + blkg) * kMaxState + (nn - 3)) = v;
            }
          }
        }
#endif
      }
#pragma unroll
      for (int qi = 0; qi < NQ; ++qi) {
        float cv[M];
        lds_seg<T, M, ROWB>(tC, n + qi, lp, cv);
#pragma unroll
        for (int kk = 0; kk < H2; ++kk) {
          float2 y2 = f2(y[2 * kk], y[2 * kk + 1]);
          y2 = fma2(f2(cv[2 * kk], cv[2 * kk + 1]), f2(bv[qi][2 * kk], bv[qi][2 * kk + 1]), y2);
          y[2 * kk] = y2.x;
          y[2 * kk + 1] = y2.y;
        }
      }
    }

    if constexpr (kHasZ) {
#pragma unroll
      for (int i = 0; i < M; ++i) y[i] *= zz[i] * sigmoid_f(zz[i]);
    }
    if (row_ok && a.cp_mode != 1) {
      if (sizeof(T) == 2 && a.out_f32)
        stg_items<float, M>(reinterpret_cast<float*>(a.out) + (long)q.b * a.o_bs + (long)d * a.o_ds, y, t0, a.L,
                            a.vec_out != 0);
      else
        stg_items<T, M>(outrow, y, t0, a.L, a.vec_out != 0);
    }
    if (a.trace && tid == 0) {
      unsigned long long* tr = a.trace + (long)t * 8;
      tr[6] = tr_recv0;
      tr[7] = tr_recvN;
      tr[0] = (unsigned long long)t | ((unsigned long long)fast << 32);
      tr[1] = smid();
      tr[2] = tr_start;
      tr[3] = tr_loop;
      tr[4] = tr_wait;
      tr[5] = gtime_ns();
    }

    // every warp is done with this tile's B/C stage, sm_A2 and sm_hin
    __syncthreads();
    if (a.claim_late) {  // few row blocks, long L: claim just in time (see ScanKArgs::claim_late)
      if (tid == 0) {
        nxt = atomicAdd(a.ticket, 1u);
        tk[(k + 1) & 3] = (int)nxt;
        if ((int)nxt < a.ntiles) {
          const TileId qn = decode_tile<R, false>(a, (int)nxt);
          tq[(k + 1) & 3] = qn;
          if (kTMA) {
            issue_rows(qn);
            issue_bc(qn, (k + 1) & 1);
          }
        }
      }
      __syncthreads();
    }

  }
}

// ================================================================================================
// Backward
// ================================================================================================
// 16-byte vector add to global memory (one RED per four floats)
__device__ __forceinline__ void red_add_v4(float* dst, float4 v) {
  asm volatile("red.relaxed.gpu.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w)
               : "memory");
}

#ifndef NZ_BWD_MINB
#define NZ_BWD_MINB 2  // resident CTAs per SM the backward is compiled for
#endif
template <typename T, int M, int LPR, int WARPS, bool kTMA, bool kHasZ>
__global__ void __launch_bounds__(WARPS * 32, (kHasZ && sizeof(T) == 4) ? 1 : NZ_BWD_MINB)
    scan_bwd_kernel(const __grid_constant__ ScanKArgs a) {
  using Cfg = ScanCfg<T, M, LPR, WARPS, kHasZ, true>;
  constexpr int R = Cfg::R, TL = Cfg::TL, ROWB = Cfg::ROWB, SEGB = Cfg::SEGB, RPW = Cfg::RPW;
  constexpr int ROWTILE = Cfg::ROWTILE, BCTILE = Cfg::BCTILE, SLROW = Cfg::SLROW, SLAB = Cfg::SLAB;
  constexpr int NT = WARPS * 32;
  constexpr int H2 = M / 2;
  constexpr int APT = (R * kMaxState + NT - 1) / NT;
  constexpr int NRW = Cfg::NRW;          // warps that reduce one state's slabs (one float4 of dB or dC per lane)
  constexpr int NGRP = WARPS / NRW;      // the reducing warps rotate over the states
  constexpr int QPA = TL / 4;            // float4 items per array

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* rows = smem;
  uint8_t* bcs = smem + Cfg::ROWS_REGION;
  uint8_t* slabs = bcs + (kTMA ? 2 : 1) * Cfg::BC_TX;           // [2 buffers][dB, dC][R][TL] fp32
  uint8_t* ucopy = slabs + 4 * SLAB;                            // [R][TL] fp32: the tile's u, re-read by the epilogue
  float* sm_gs = reinterpret_cast<float*>(ucopy + SLAB);        // [n][row][segment pair] dA partials of this tile
  uint8_t* tail = reinterpret_cast<uint8_t*>(sm_gs) + Cfg::GSB;
  uint64_t* bars = reinterpret_cast<uint64_t*>(tail);  // [0] rows, [1],[2] B/C stages, [3],[4] slab full, [5],[6] slab free
  volatile int* tk = reinterpret_cast<volatile int*>(tail + 64);
  TileId* tq = reinterpret_cast<TileId*>(tail + 128);
  float* sm_A2 = reinterpret_cast<float*>(tail + 256);
  float* sm_hc = sm_A2 + R * kMaxState;   // h carried into the tile (forward checkpoints)
  float* sm_dhc = sm_hc + R * kMaxState;  // dh carried into the tile when the later chunk had already finished
  const uint32_t rows_s = keep(smem_u32(rows)), bcs_s = rows_s + Cfg::ROWS_REGION;
  const uint32_t slabs_s = rows_s + (uint32_t)(slabs - rows), gs_s0 = rows_s + (uint32_t)(reinterpret_cast<uint8_t*>(sm_gs) - rows);
  const uint32_t a2_s0 = rows_s + (uint32_t)(reinterpret_cast<uint8_t*>(sm_A2) - rows);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int sl = lane / RPW;
  const int rw = lane % RPW;
  const int rloc = warp * RPW + rw;
  const int N = a.dstate;  // (a compile-time 16 on the TMA path measured 6% slower: register allocation)
  const uint32_t segoff = sl * SEGB;
  const LanePre<M, T> lp(segoff);
  // slab addressing (loop invariant): row = rloc, the lane's M floats.  Odd rows flip 16-byte-chunk
  // bit 0 so that the 8 lanes of a quarter warp (four segments x two rows) hit 8 different bank groups.
  uint32_t slab_w[M / 4];
#pragma unroll
  for (int j = 0; j < M / 4; ++j) slab_w[j] = keep(swz128(rloc * SLROW + sl * (M * 4) + 16 * j) ^ ((rloc & 1) ? 16u : 0u));
  // reducer role: lane -> one float4 (four time steps) of dB (red_arr 0) or dC (1)
  const int red_item = (warp % NRW) * 32 + lane;
  const int red_arr = red_item / QPA, red_q = red_item % QPA;
  const uint32_t red_in = keep(swz128((uint32_t)red_q * 16u));
  [[maybe_unused]] const uint32_t red_in1 = keep(swz128((uint32_t)(tid % TL) * 4u));

  auto issue_rows = [&](const TileId& q) {
    mbar_arrive_expect_tx(&bars[0], Cfg::ROWS_TX);
    const int c1 = q.c * (TL / Cfg::INNER);
    tma_load_4d(rows, &a.tm_u, &bars[0], 0, c1, q.d0, q.b);
    tma_load_4d(rows + ROWTILE, &a.tm_delta, &bars[0], 0, c1, q.d0, q.b);
    tma_load_4d(rows + 2 * ROWTILE, &a.tm_dout, &bars[0], 0, c1, q.d0, q.b);
    if (kHasZ) tma_load_4d(rows + 3 * ROWTILE, &a.tm_z, &bars[0], 0, c1, q.d0, q.b);
  };
  auto issue_bc = [&](const TileId& q, int s) {
    uint8_t* st = bcs + s * Cfg::BC_TX;
    mbar_arrive_expect_tx(&bars[1 + s], Cfg::BC_TX);
    const int c1 = q.c * (TL / Cfg::INNER);
    tma_load_5d(st, &a.tm_B, &bars[1 + s], 0, c1, 0, q.g, q.b);
    tma_load_5d(st + BCTILE, &a.tm_C, &bars[1 + s], 0, c1, 0, q.g, q.b);
  };

  if (tid == 0) {
    mbar_init(&bars[0], 1);
    mbar_init(&bars[1], 1);
    mbar_init(&bars[2], 1);
    mbar_init(&bars[3], NT);
    mbar_init(&bars[4], NT);
    mbar_init(&bars[5], NZ_BWD_REDUCE_ALL ? NT : NRW * 32);
    mbar_init(&bars[6], NZ_BWD_REDUCE_ALL ? NT : NRW * 32);
    fence_mbar_init();
    const int t0 = (int)atomicAdd(a.ticket, 1u);
    tk[0] = t0;
    if (t0 < a.ntiles) {
      const TileId q = decode_tile<R, true>(a, t0);
      tq[0] = q;
      if (kTMA) {
        issue_rows(q);
        issue_bc(q, 0);
      }
    }
  }
  __syncthreads();

  unsigned g = 0;  // states processed so far by this CTA: slab buffer g & 1, barrier parity (g >> 1) & 1

  for (int k = 0;; ++k) {
    const int t = tk[k & 3];
    if (t >= a.ntiles) break;
    const TileId q = tq[k & 3];  // decoded once by thread 0 when it claimed the ticket
    const int c = q.c, s = kTMA ? (k & 1) : 0;
    const bool row_ok = kTMA || rloc < q.rows_valid;  // the TMA path only runs whole row blocks
    const int d = q.d0 + (row_ok ? rloc : 0);
    const long rowg = (long)q.b * a.dim + d;
    const int bpg = a.nrb;
    unsigned nxt = 0;  // the next ticket, claimed one tile ahead (see the forward kernel)
    if (tid == 0 && !a.claim_late) nxt = atomicAdd(a.ticket, 1u);

    const float Dv = a.D ? __ldg(a.D + d) : 0.f;
    const float bias = a.bias ? __ldg(a.bias + d) : 0.f;
    // per (row, state) tile constants, one per thread: A, the h checkpoint, and -- if the tile of the
    // later chunk has already published it -- the dh carry (the common case when row blocks outnumber
    // the CTAs; otherwise the state loop polls per state, see below)
    float a2pre[APT], hcpre[APT], dhpre[APT];
    bool all_in = true;
#pragma unroll
    for (int j = 0; j < APT; ++j) {
      const int i = tid + j * NT, r = i / kMaxState, n = i % kMaxState;
      const bool ok = i < R * kMaxState && r < q.rows_valid && n < N;
      a2pre[j] = ok ? __ldg(a.A + (long)(q.d0 + r) * a.A_ds + n) * kLog2e : 0.f;
      hcpre[j] = (ok && c > 0) ? __ldg(a.x + (((long)q.b * a.dim + q.d0 + r) * a.nck + (c - 1)) * N + n) : 0.f;
      dhpre[j] = 0.f;
      if (ok && c + 1 < a.nchunks) {
        unsigned tag;
        slot_load(a.carry + ((((long)q.b * a.dim + q.d0 + r) * 2 + ((c + 1) & 1)) * kMaxState + n), dhpre[j], tag);
        all_in = all_in && tag == (unsigned)c + 2u;
      }
    }
    // dl of the first step of the later chunk (a_{t+1} of the reverse recurrence at the tile end)
    float dlfirst_next = 0.f;
    if (sl == LPR - 1 && row_ok && (long)(c + 1) * TL < a.L) {
      float x = Elem<T>::to_f(reinterpret_cast<const T*>(a.delta)[(long)q.b * a.dl_bs + (long)d * a.dl_ds + (long)(c + 1) * TL]) + bias;
      if (a.softplus) x = softplus_f(x);
      dlfirst_next = x;
    }
    const unsigned long long* cin = a.carry + (rowg * 2 + ((c + 1) & 1)) * kMaxState;
    const bool chained = c + 1 < a.nchunks && row_ok;

    bool fast;  // every dh carry of the tile was already there: no polling in the state loop
    if constexpr (kTMA) {
      mbar_wait(&bars[0], k & 1);
    } else {
      const long ts = (long)c * TL;
      const int rv = q.rows_valid;
      coop_fill<T, TL>(rows, reinterpret_cast<const T*>(a.u) + (long)q.b * a.u_bs + (long)q.d0 * a.u_ds, a.u_ds, rv, R, ts, a.L);
      coop_fill<T, TL>(rows + ROWTILE, reinterpret_cast<const T*>(a.delta) + (long)q.b * a.dl_bs + (long)q.d0 * a.dl_ds,
                       a.dl_ds, rv, R, ts, a.L);
      coop_fill<T, TL>(rows + 2 * ROWTILE, reinterpret_cast<const T*>(a.dout) + (long)q.b * a.do_bs + (long)q.d0 * a.do_ds,
                       a.do_ds, rv, R, ts, a.L);
      if (kHasZ)
        coop_fill<T, TL>(rows + 3 * ROWTILE, reinterpret_cast<const T*>(a.z) + (long)q.b * a.z_bs + (long)q.d0 * a.z_ds,
                         a.z_ds, rv, R, ts, a.L);
      coop_fill<T, TL>(bcs, reinterpret_cast<const T*>(a.B) + (long)q.b * a.B_bs + (long)q.g * a.B_gs, a.B_ns, N,
                       kMaxState, ts, a.L);
      coop_fill<T, TL>(bcs + BCTILE, reinterpret_cast<const T*>(a.C) + (long)q.b * a.C_bs + (long)q.g * a.C_gs, a.C_ns, N,
                       kMaxState, ts, a.L);
#pragma unroll
      for (int j = 0; j < APT; ++j)
        if (tid + j * NT < R * kMaxState) {
          sm_A2[tid + j * NT] = a2pre[j];
          sm_hc[tid + j * NT] = hcpre[j];
          sm_dhc[tid + j * NT] = dhpre[j];
        }
      if (tid == 0 && !a.claim_late) {
        tk[(k + 1) & 3] = (int)nxt;
        if ((int)nxt < a.ntiles) tq[(k + 1) & 3] = decode_tile<R, true>(a, (int)nxt);
      }
      fast = __syncthreads_and(all_in) != 0;
    }
    const long t0 = (long)c * TL + sl * M;
    const int nvalid = (int)max(0L, min((long)M, a.L - t0));  // steps of this lane's segment inside the sequence

    float dl[M], dy[M], dlu[M], sB[M], ddl[M];
    float yv[kHasZ ? M : 1], dzf[kHasZ ? M : 1];
    lds_seg<T, M, ROWB>(rows_s, rloc, lp, dlu);
    lds_seg<T, M, ROWB>(rows_s + ROWTILE, rloc, lp, dl);
    // u is needed again only by the epilogue: park it in shared memory (the row tiles are refilled early)
#pragma unroll
    for (int j = 0; j < M / 4; ++j)
      *reinterpret_cast<float4*>(ucopy + slab_w[j]) = make_float4(dlu[4 * j], dlu[4 * j + 1], dlu[4 * j + 2], dlu[4 * j + 3]);
    lds_seg<T, M, ROWB>(rows_s + 2 * ROWTILE, rloc, lp, dy);
    if constexpr (kHasZ) {
      float zz[M];
      lds_seg<T, M, ROWB>(rows_s + 3 * ROWTILE, rloc, lp, zz);
#pragma unroll
      for (int i = 0; i < M; ++i) {
        const float sg = sigmoid_f(zz[i]);
        dzf[i] = dy[i] * sg * (1.f + zz[i] * (1.f - sg));  // dout * d silu(z)/dz
        dy[i] = dy[i] * zz[i] * sg;                         // dout * silu(z)
        yv[i] = Dv * dlu[i];
      }
    }
    float dlsum = 0.f;
#pragma unroll
    for (int i = 0; i < M; ++i) {
      float x = dl[i] + bias;
      if (a.softplus) x = softplus_f(x);
      if (i >= nvalid) {
        x = 0.f;
        dy[i] = 0.f;
      }
      dl[i] = x;
      dlsum += x;
      dlu[i] = x * dlu[i];
      sB[i] = 0.f;
      ddl[i] = 0.f;
    }
    // dl of the step right after this lane's segment (next segment's first step, or the first step
    // of the later chunk for the last segment): a_{t+1} of the reverse recurrence at the segment end
    float dlnext = __shfl_down_sync(0xffffffffu, dl[0], RPW);
    if (sl == LPR - 1) dlnext = dlfirst_next;
    const float qsum = keep(dlsum - dl[0] + dlnext);  // sum of dl over (segment shifted by one step)
    dlsum = keep(dlsum);

    if constexpr (kTMA) {
#pragma unroll
      for (int j = 0; j < APT; ++j)
        if (tid + j * NT < R * kMaxState) {
          sm_A2[tid + j * NT] = a2pre[j];
          sm_hc[tid + j * NT] = hcpre[j];
          sm_dhc[tid + j * NT] = dhpre[j];
        }
      fast = __syncthreads_and(all_in) != 0;  // row values are in registers: request the next ticket's tiles
      if (tid == 0 && !a.claim_late) {
        tk[(k + 1) & 3] = (int)nxt;
        if ((int)nxt < a.ntiles) {
          const TileId qn = decode_tile<R, true>(a, (int)nxt);
          tq[(k + 1) & 3] = qn;
          issue_rows(qn);
          issue_bc(qn, (k + 1) & 1);
        }
      }
      mbar_wait(&bars[1 + s], (k >> 1) & 1);
    }
    const uint32_t tB = bcs_s + s * Cfg::BC_TX, tC = tB + BCTILE;
    float* dG = (red_arr ? a.dC : a.dB) + (((long)q.b * a.ngroups + q.g) * N) * a.L + (long)c * TL + red_q * 4;
    const bool red_vec = a.vec_grad && (long)c * TL + red_q * 4 + 4 <= a.L;

    // dB/dC of state n (CTA-local state counter gp): the NRW reducing warps add the R slab rows
    auto reduce_slab = [&](int n, unsigned gp) {
      const int bq = gp & 1;
#if NZ_EXP_NOSLAB
      return;  // timing experiment (wrong results): no slab hand-off, no reduction
#endif
      mbar_wait(&bars[3 + bq], (gp >> 1) & 1);
#if NZ_BWD_REDUCE_ALL
      {  // every thread adds the R rows of ONE element (balanced: no warp carries a whole reduction)
        static_assert(!NZ_BWD_REDUCE_ALL || 2 * TL == NT, "one dB/dC element per thread");
        const int arr = tid / TL, tt = tid - arr * TL;
        const uint32_t sb1 = slabs_s + (bq * 2 + arr) * SLAB;
        float acc0 = 0.f, acc1 = 0.f;
#pragma unroll
        for (int r = 0; r < R; r += 2) {
          const uint32_t k0 = (uint32_t)((((r * (SLROW / 128)) & 7) << 4));
          const uint32_t k1 = (uint32_t)(((((r + 1) * (SLROW / 128)) & 7) << 4) ^ 16);
          acc0 += lds32(sb1 + r * SLROW + (red_in1 ^ k0));
          acc1 += lds32(sb1 + (r + 1) * SLROW + (red_in1 ^ k1));
        }
        mbar_arrive(&bars[5 + bq]);
        if ((long)c * TL + tt < a.L) {
          float* dst1 = (arr ? a.dC : a.dB) + (((long)q.b * a.ngroups + q.g) * N + n) * a.L + (long)c * TL + tt;
          if (bpg == 1) *dst1 = acc0 + acc1; else atomicAdd(dst1, acc0 + acc1);
        }
        return;
      }
#endif
      const uint32_t sb = slabs_s + (bq * 2 + red_arr) * SLAB;
      float2 lo = f2(0.f, 0.f), hi = f2(0.f, 0.f);
#pragma unroll
      for (int r = 0; r < R; ++r) {  // row r adds r*SLROW to the address and its swizzle key / parity flip to the offset
        const uint32_t key = (uint32_t)((((r * (SLROW / 128)) & 7) << 4) ^ ((r & 1) ? 16 : 0));
        const uint4 v = lds128(sb + r * SLROW + (red_in ^ key));
        lo = __fadd2_rn(lo, f2(__uint_as_float(v.x), __uint_as_float(v.y)));
        hi = __fadd2_rn(hi, f2(__uint_as_float(v.z), __uint_as_float(v.w)));
      }
      mbar_arrive(&bars[5 + bq]);
      float* dst = dG + (long)n * a.L;
      if (red_vec) {
        if (bpg == 1)
          *reinterpret_cast<float4*>(dst) = make_float4(lo.x, lo.y, hi.x, hi.y);
        else
          red_add_v4(dst, make_float4(lo.x, lo.y, hi.x, hi.y));
      } else {
        const float vv[4] = {lo.x, lo.y, hi.x, hi.y};
        const long tg = (long)c * TL + red_q * 4;
#pragma unroll
        for (int e = 0; e < 4; ++e)
          if (tg + e < a.L) {
            if (bpg == 1) dst[e] = vv[e]; else atomicAdd(dst + e, vv[e]);
          }
      }
    };

    // when the later chunk's tile is still running: start once it is a few states ahead, so that the
    // one-state-ahead prefetch of the per-state polling below finds its data
    float cv_next = 0.f;
    unsigned ct_next = 0;
    if (!fast && chained) {
      if (a.skew >= 0) {
        const int nsk = min(a.skew, N - 1);
        unsigned tg;
        float dummy;
        do {
          slot_load(cin + nsk, dummy, tg);
        } while (tg != (unsigned)c + 2u);
      }
      slot_load(cin, cv_next, ct_next);
    }
    uint32_t gsp = gs_s0 + (rloc * (LPR / 2) + (sl >> 1)) * 4;

    // per-state pointers advance by one per trip (keeps 64-bit address arithmetic out of the loop body)
    const unsigned long long* cin_p = cin;
    const int coff = (c & 1) ? kMaxState : -kMaxState;  // outgoing ring slot (c & 1) relative to the incoming one ((c + 1) & 1)
    uint32_t a2_p = a2_s0 + rloc * (kMaxState * 4);  // sm_hc / sm_dhc follow at fixed distances
    [[maybe_unused]] float exp_hin = 1.f;
    [[maybe_unused]] const float* exp_x = a.x + rowg * a.nck * N + sl;
    [[maybe_unused]] const int exp_max = a.nck * N - 16;
    NZ_UNROLL(NZ_BWD_UNROLL)
    for (int n = 0; n < N; ++n, ++g, ++cin_p, a2_p += 4, gsp += R * LPR / 2 * 4) {
      const float A2 = lds32(a2_p);
      const float An = A2 * kLn2;
      const float hc = lds32(a2_p + R * kMaxState * 4);
      float dhc = lds32(a2_p + 2 * R * kMaxState * 4);
      unsigned dtag = 0;
      if (!fast) {
        dhc = cv_next;
        dtag = ct_next;
        if (chained && n + 1 < N) slot_load(cin_p + 1, cv_next, ct_next);
      }
      float av[M], bv[M], cv[M], hh[M];
      [[maybe_unused]] float Bk[M];
      lds_seg<T, M, ROWB>(tB, n, lp, bv);
      lds_seg<T, M, ROWB>(tC, n, lp, cv);
#if !NZ_EXP_NOHSCAN
      float P = ex2_approx(A2 * dlsum);    // prod a over the segment
#endif
      float Q = ex2_approx(A2 * qsum);     // prod a over the segment shifted by one step
      const float anl = ex2_approx(A2 * dlnext);
      float cdy[M];
#pragma unroll
      for (int kk = 0; kk < H2; ++kk) {
        const float2 x2 = mul2(f2(dl[2 * kk], dl[2 * kk + 1]), f2(A2, A2));
        av[2 * kk] = ex2_approx(x2.x);
        av[2 * kk + 1] = ex2_approx(x2.y);
        const float2 b2 = mul2(f2(dlu[2 * kk], dlu[2 * kk + 1]), f2(bv[2 * kk], bv[2 * kk + 1]));
        hh[2 * kk] = b2.x;  // hh holds b_t until the replay overwrites it with h_t
        hh[2 * kk + 1] = b2.y;
#if NZ_BWD_KEEPB
        Bk[2 * kk] = bv[2 * kk];
        Bk[2 * kk + 1] = bv[2 * kk + 1];
#endif
        const float2 c2 = mul2(f2(cv[2 * kk], cv[2 * kk + 1]), f2(dy[2 * kk], dy[2 * kk + 1]));
        cdy[2 * kk] = c2.x;
        cdy[2 * kk + 1] = c2.y;
      }
      // ---- both folds (independent chains): forward for h, reverse for dh ----
      float G = 0.f;
#if NZ_EXP_NOHSCAN
#pragma unroll
      for (int i = 0; i < M; ++i) {
        const int j = M - 1 - i;
        G = fmaf(j == M - 1 ? anl : av[j + 1], G, cdy[j]);
      }
#pragma unroll
      for (int off = 1; off < LPR; off <<= 1) ks_down_w(Q, G, off * RPW);
#else
      float H = 0.f;
#pragma unroll
      for (int i = 0; i < M; ++i) {
        H = fmaf(av[i], H, hh[i]);
        const int j = M - 1 - i;
        G = fmaf(j == M - 1 ? anl : av[j + 1], G, cdy[j]);
      }
      // ---- both Kogge-Stone scans across the segments of the row, interleaved ----
#pragma unroll
      for (int off = 1; off < LPR; off <<= 1) {
        ks_up_w(P, H, off * RPW);
        ks_down_w(Q, G, off * RPW);
      }
#endif
      if (!fast && chained) {
        slot_wait<true>(cin_p, (unsigned)c + 2u, dhc, dtag);
      }
#if NZ_EXP_NOHSCAN
      float h = hc * exp_hin;  // timing experiment only: a prefetched per-lane global load stands in
      exp_hin = __ldg(exp_x + min((n + 1) * 16, exp_max));
#else
      float h = ks_enter_up_w<RPW>(P, H, hc);
#endif
      float dh = ks_enter_down_w<RPW>(Q, G, dhc);
      if (sl == 0 && row_ok)  // dh leaving the tile
        slot_store(const_cast<unsigned long long*>(cin_p) + coff, fmaf(Q, dhc, G), (unsigned)c + 1u);
#if !NZ_BWD_REDUCE_LATE
      // the reducing warps of the previous state add its slabs while this state's scans are in flight
      if (n > 0 && (NZ_BWD_REDUCE_ALL || (int)((g - 1) % NGRP) == warp / NRW)) reduce_slab(n - 1, g - 1);
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
      for (int kk = 0; kk < H2; ++kk) {
        const float2 h2 = f2(hh[2 * kk], hh[2 * kk + 1]);
        const float2 d2 = f2(dd[2 * kk], dd[2 * kk + 1]);
        const float2 dy2 = f2(dy[2 * kk], dy[2 * kk + 1]);
        const float2 dlu2 = f2(dlu[2 * kk], dlu[2 * kk + 1]);
        const float2 c2 = mul2(dy2, h2);  // dC_t[n] contribution of this row
        vC[2 * kk] = c2.x;
        vC[2 * kk + 1] = c2.y;
        const float2 ah2 = sub2(h2, f2(bv[2 * kk], bv[2 * kk + 1]));  // a_t h_{t-1}
        const float2 gq2 = mul2(d2, ah2);
        float2 ddl2 = f2(ddl[2 * kk], ddl[2 * kk + 1]);
        ddl2 = fma2(f2(An, An), gq2, ddl2);
        ddl[2 * kk] = ddl2.x;
        ddl[2 * kk + 1] = ddl2.y;
        gs2 = fma2(f2(dl[2 * kk], dl[2 * kk + 1]), gq2, gs2);
        const float2 b2 = mul2(d2, dlu2);  // dB_t[n] contribution of this row
        vB[2 * kk] = b2.x;
        vB[2 * kk + 1] = b2.y;
      }
      {
#if NZ_BWD_KEEPB
        const float (&Bv)[M] = Bk;
#else
        float Bv[M];
        lds_seg<T, M, ROWB>(tB, n, lp, Bv);
#endif
#pragma unroll
        for (int kk = 0; kk < H2; ++kk) {
          float2 s2 = f2(sB[2 * kk], sB[2 * kk + 1]);
          s2 = fma2(f2(dd[2 * kk], dd[2 * kk + 1]), f2(Bv[2 * kk], Bv[2 * kk + 1]), s2);
          sB[2 * kk] = s2.x;
          sB[2 * kk + 1] = s2.y;
        }
        if constexpr (kHasZ) {
          float Cv[M];
          lds_seg<T, M, ROWB>(tC, n, lp, Cv);
#pragma unroll
          for (int kk = 0; kk < H2; ++kk) {
            float2 y2 = f2(yv[2 * kk], yv[2 * kk + 1]);
            y2 = fma2(f2(Cv[2 * kk], Cv[2 * kk + 1]), f2(hh[2 * kk], hh[2 * kk + 1]), y2);
            yv[2 * kk] = y2.x;
            yv[2 * kk + 1] = y2.y;
          }
        }
      }
      // this lane's share of dA[n]: summed over the segments at the end of the tile
      {
        const float gsl = gs2.x + gs2.y;
        const float gsum = gsl + __shfl_down_sync(0xffffffffu, gsl, RPW);  // this segment + the next one
        if ((sl & 1) == 0) sts32(gsp, gsum);
      }

#if NZ_BWD_REDUCE_LATE == 1
      if (n > 0 && (NZ_BWD_REDUCE_ALL || (int)((g - 1) % NGRP) == warp / NRW)) reduce_slab(n - 1, g - 1);
#endif
      // ---- dB/dC: every row writes its products into slab buffer g & 1 (reduced two states later at the latest) ----
#if !NZ_EXP_NOSLAB
      if (g >= 2) mbar_wait(&bars[5 + (g & 1)], ((g >> 1) - 1) & 1);  // the buffer's previous content has been consumed
#endif
      {
        const uint32_t sb = slabs_s + (g & 1) * (2 * SLAB);
#pragma unroll
        for (int j = 0; j < M / 4; ++j) {
          sts128(sb + slab_w[j], vB[4 * j], vB[4 * j + 1], vB[4 * j + 2], vB[4 * j + 3]);
          sts128(sb + SLAB + slab_w[j], vC[4 * j], vC[4 * j + 1], vC[4 * j + 2], vC[4 * j + 3]);
        }
      }
#if !NZ_EXP_NOSLAB
      mbar_arrive(&bars[3 + (g & 1)]);  // my part of slab(n) is written
#endif
#if NZ_BWD_REDUCE_LATE == 2
      if (n > 0 && (NZ_BWD_REDUCE_ALL || (int)((g - 1) % NGRP) == warp / NRW)) reduce_slab(n - 1, g - 1);
#endif
    }
    // the last state's slabs of this tile
    if (NZ_BWD_REDUCE_ALL || (int)((g - 1) % NGRP) == warp / NRW) reduce_slab(N - 1, g - 1);

    // ---- per-(row, t) epilogue ----
    const long rowlin = rowg * a.L;  // gradients are contiguous (batch, dim, L)
    float outv[M];
#pragma unroll
    for (int i = 0; i < M; ++i) outv[i] = fmaf(dl[i], sB[i], Dv * dy[i]);  // du
    if (row_ok) stg_items<T, M>(reinterpret_cast<T*>(a.du) + rowlin, outv, t0, a.L, a.vec_grad != 0);
    float dD_acc = 0.f, db_acc = 0.f;
    float uu[M];
#pragma unroll
    for (int j = 0; j < M / 4; ++j) {
      const float4 v = *reinterpret_cast<const float4*>(ucopy + slab_w[j]);
      uu[4 * j] = v.x, uu[4 * j + 1] = v.y, uu[4 * j + 2] = v.z, uu[4 * j + 3] = v.w;
    }
#pragma unroll
    for (int i = 0; i < M; ++i) {
      float gd = fmaf(uu[i], sB[i], ddl[i]);  // d loss / d dl
      // d softplus(x)/dx = sigmoid(x) = 1 - exp(-softplus(x))
      if (a.softplus) gd *= sigmoid_from_softplus(dl[i]);
      if (i >= nvalid) gd = 0.f;
      outv[i] = gd;
      db_acc += gd;
      dD_acc = fmaf(dy[i], uu[i], dD_acc);
    }
    if (row_ok) stg_items<T, M>(reinterpret_cast<T*>(a.ddelta) + rowlin, outv, t0, a.L, a.vec_grad != 0);
    if constexpr (kHasZ) {
#pragma unroll
      for (int i = 0; i < M; ++i) outv[i] = dzf[i] * yv[i];
      if (row_ok) stg_items<T, M>(reinterpret_cast<T*>(a.dz) + rowlin, outv, t0, a.L, a.vec_grad != 0);
    }
    // (dim)-shaped reductions: over the segments here, over batch and chunks with fp32 atomics
#pragma unroll
    for (int off = 16; off >= RPW; off >>= 1) {
      dD_acc += __shfl_xor_sync(0xffffffffu, dD_acc, off);
      db_acc += __shfl_xor_sync(0xffffffffu, db_acc, off);
    }
    if (sl == 0 && row_ok) {
      if (a.dD) atomicAdd(a.dD + d, dD_acc);
      if (a.dbias) atomicAdd(a.dbias + d, db_acc);
    }

    // every warp is done with this tile's B/C stage, sm_A2 / sm_hc / sm_dhc and has written its dA partials
    __syncthreads();
    if (a.claim_late) {  // few row blocks, long L: claim just in time (see ScanKArgs::claim_late)
      if (tid == 0) {
        nxt = atomicAdd(a.ticket, 1u);
        tk[(k + 1) & 3] = (int)nxt;
        if ((int)nxt < a.ntiles) {
          const TileId qn = decode_tile<R, true>(a, (int)nxt);
          tq[(k + 1) & 3] = qn;
          if (kTMA) {
            issue_rows(qn);
            issue_bc(qn, (k + 1) & 1);
          }
        }
      }
      __syncthreads();
    }

#pragma unroll
    for (int j = 0; j < APT; ++j) {
      const int i = tid + j * NT, r = i % R, n = i / R;  // consecutive threads: consecutive rows (64-byte pitch)
      if (i < R * kMaxState && r < q.rows_valid && n < N) {
        const float4* p = reinterpret_cast<const float4*>(sm_gs + (n * R + r) * (LPR / 2));
        float acc = 0.f;
#pragma unroll
        for (int v = 0; v < LPR / 8; ++v) {
          const float4 w4 = p[(v + (r >> 2)) & (LPR / 8 - 1)];  // rotated: a quarter warp covers all 8 bank groups
          acc += (w4.x + w4.y) + (w4.z + w4.w);
        }
        atomicAdd(a.dA + (long)(q.d0 + r) * N + n, acc);
      }
    }
    __syncthreads();  // sm_gs may be rewritten by the next tile
  }
}

// Launch wrappers instantiated per element type in scan_inst_*.cu
template <typename T>
cudaError_t launch_scan_fwd(const ScanKArgs& a, bool tma, bool has_z, cudaStream_t stream);
template <typename T>
cudaError_t launch_scan_bwd(const ScanKArgs& a, bool tma, bool has_z, cudaStream_t stream);

}  // namespace nz
