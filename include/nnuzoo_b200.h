/*
 * nnuzoo_b200.h -- C ABI of the B200-native selective-scan / SS2D hot path.
 *
 * This is the drop-in boundary.  The reference reaches its native code
 * through a pybind module that is NOT in its tree (`selective_scan_cuda`, from mamba_ssm):
 *
 *   selective_scan_cuda.fwd(u, delta, A, B, C, D, z, delta_bias, delta_softplus) -> [out, x, (out_z)]
 *       called at nnunetv2/nets/seg_mamba/selective_scan_interface.py:37
 *   selective_scan_cuda.bwd(u, delta, A, B, C, D, z, delta_bias, dout, x, out, dz,
 *                           delta_softplus, recompute_out_z) -> [du, ddelta, dA, dB, dC, dD, ddelta_bias, (dz)]
 *       called at nnunetv2/nets/seg_mamba/selective_scan_interface.py:62
 *
 * nz_scan_fwd / nz_scan_bwd replace exactly those two entry points; nz_cross_scan /
 * nz_cross_merge replace the CrossScan / CrossMerge tensor shuffles that the reference writes
 * inline in PyTorch (m2net.py:175-177, :202-206 + :218; ssnd2net.py:249-255, :285-299).
 *
 * Conventions
 *   - plain C, no torch types; every pointer is a DEVICE pointer unless the name ends in _host
 *   - the caller owns and allocates every buffer (outputs, checkpoints, gradient accumulators);
 *     kernels never allocate and never synchronise the device
 *   - all launches go to the `stream` argument (a cudaStream_t passed as void*); the calls are
 *     re-entrant, hold no global state and are CUDA-graph capturable
 *   - strides are in ELEMENTS; the innermost (sequence) stride of u/delta/B/C/z/dout is 1, the same
 *     constraint the reference enforces at selective_scan_interface.py:19-30
 *   - return value: 0 on success, a negative NZ_E* code otherwise; nz_last_error() returns a
 *     thread-local human-readable message.  No exceptions cross this boundary.
 */
#ifndef NNUZOO_B200_H
#define NNUZOO_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define NZ_ABI_VERSION 4

/* element types of u / delta / B / C / z / out / dout / du / ddelta / dz */
#define NZ_F32 0
#define NZ_BF16 1
#define NZ_F16 2

/* error codes */
#define NZ_OK 0
#define NZ_EINVAL (-1)      /* bad shape / stride / dtype / null pointer */
#define NZ_EUNSUPPORTED (-2) /* valid request outside what the kernels implement (e.g. d_state > 16) */
#define NZ_ECUDA (-3)       /* a CUDA runtime / driver call failed */

/* The forward saves the state h every NZ_CHUNK steps ("x" of the reference ABI); the backward
 * recomputes inside a chunk from it. */
#ifndef NZ_CHUNK
#define NZ_CHUNK 128
#endif
#define NZ_MAX_DSTATE 16
/* Optional second, finer set of checkpoints: h at the end of every NZ_FINE steps (NzScanDesc::xf).  It feeds the
 * row-per-lane backward (csrc/scan_rl_kernels.cuh), which then needs no intra-warp scan at all. */
#define NZ_FINE 8

typedef struct NzScanDesc {
  /* ---- problem ---- */
  int32_t batch;          /* B                                                     */
  int32_t dim;            /* K*D rows per batch entry                              */
  int32_t dstate;         /* N (<= NZ_MAX_DSTATE)                                  */
  int32_t ngroups;        /* G: B/C are (batch, G, N, L); dim % G == 0             */
  int64_t seqlen;         /* L                                                     */
  int32_t dtype;          /* NZ_F32 / NZ_BF16 / NZ_F16                             */
  int32_t delta_softplus; /* 0 / 1                                                 */
  int32_t force_generic;  /* 1: never use the TMA path (testing)                   */
  int32_t out_f32;        /* 1 with a 16-bit dtype: `out` is fp32 (16-bit operands in, fp32 result out -- SS2D under
                             autocast keeps an fp32 out_y, m2net.py:185-200); ignored for NZ_F32 and by nz_scan_bwd */

  /* ---- forward inputs (selective_scan_cuda.fwd arguments) ---- */
  const void* u;            /* (batch, dim, L)                 */
  const void* delta;        /* (batch, dim, L)                 */
  const float* A;           /* (dim, N) fp32, real             */
  const void* B;            /* (batch, G, N, L)                */
  const void* C;            /* (batch, G, N, L)                */
  const float* D;           /* (dim) fp32 or NULL              */
  const void* z;            /* (batch, dim, L) or NULL         */
  const float* delta_bias;  /* (dim) fp32 or NULL              */
  int64_t u_stride[2];      /* batch, dim                      */
  int64_t delta_stride[2];
  int64_t z_stride[2];
  int64_t B_stride[3];      /* batch, group, state             */
  int64_t C_stride[3];
  int64_t A_stride;         /* row stride of A (state stride 1) */

  /* ---- forward outputs ---- */
  void* out;                /* (batch, dim, L): y + D*u, times SiLU(z) when z != NULL  */
  int64_t out_stride[2];
  float* x;                 /* (batch, dim, nz_scan_num_chunks(L), N) fp32 contiguous:
                               state at the end of each chunk; x[..., -1, :] is last_state */

  /* ---- backward inputs (in addition to the forward inputs and x) ---- */
  const void* dout;         /* (batch, dim, L)                 */
  int64_t dout_stride[2];

  /* ---- backward outputs ---- */
  void* du;                 /* (batch, dim, L), contiguous, dtype                         */
  void* ddelta;             /* (batch, dim, L), contiguous, dtype                         */
  void* dz;                 /* (batch, dim, L), contiguous, dtype; NULL iff z == NULL     */
  float* dA;                /* (dim, N) fp32, contiguous      -- ACCUMULATED INTO: caller zeroes */
  float* dB;                /* (batch, G, N, L) fp32 contiguous -- ACCUMULATED INTO: caller zeroes, unless        */
  float* dC;                /* (batch, G, N, L) fp32 contiguous    nz_scan_bwd_overwrites_dbc() says 1            */
  float* dD;                /* (dim) fp32 or NULL             -- ACCUMULATED INTO: caller zeroes */
  float* ddelta_bias;       /* (dim) fp32 or NULL             -- ACCUMULATED INTO: caller zeroes */

  /* ---- scratch (both directions) ---- */
  void* workspace;          /* >= nz_scan_workspace_bytes() bytes, 256-byte aligned; the call zeroes it on
                               `stream` itself.  Holds the tile ticket counter and the {value, tag} slots
                               through which consecutive chunks of a row hand their state over.  Two calls
                               that may run concurrently need separate workspaces. */
  int64_t workspace_bytes;

  /* ---- fine checkpoints (ABI v3; optional) ---- */
  float* xf;                /* NULL, or nz_scan_fine_bytes() bytes: (batch, dim, L / NZ_FINE, N) fp32, h at the end of
                               every NZ_FINE-step block.  Written by nz_scan_fwd when non-NULL and the problem
                               qualifies (nz_scan_fine_bytes() > 0; ignored otherwise); nz_scan_bwd given the same buffer
                               (and a workspace of nz_scan_workspace_bytes_bwd() bytes) runs the row-per-lane backward,
                               otherwise the warp-scan backward that only needs x. */

  /* ---- folded SS2D directions (ABI v4; both 0 for a plain selective_scan_fn call) ----
   * SS2D's CrossScan (m2net.py:175-177) feeds the scan x row-major, x column-major and the L-flips of both.  A flipped
   * direction is the same recurrence run from t = L-1 down to 0 over the UN-flipped arrays, so the flipped copies need
   * not exist: rev_mask bit g makes group g walk the sequence backwards (u, delta, B, C, z are read and out, du, ddelta,
   * dB, dC are written at their un-flipped positions; last_state / x then describe t = 0), and u_gdiv = 2 lets the
   * forward and the backward walker of one array share its rows.  Needs the row-per-lane kernels: d_state 16, groups of
   * a multiple of 32 rows, TMA-expressible operands, xf and the nz_scan_workspace_bytes_bwd() scratch in BOTH
   * directions; NZ_EUNSUPPORTED otherwise. */
  int32_t rev_mask;         /* bit g (g < 32): group g runs reversed in time */
  int32_t u_gdiv;           /* 0 / 1: u is (batch, dim, L).  n > 1: u is (batch, dim / n, L) and group g reads the rows of
                               group g / n (u_stride[1] is still the row stride); ngroups % n == 0 */
} NzScanDesc;

/* Bytes of scratch a call with this batch / dim needs (same for forward and backward). */
int64_t nz_scan_workspace_bytes(const NzScanDesc* desc);

/* Optional larger scratch size: for launches with far fewer row blocks than SMs and many chunks (the 1-D nets'
 * (2, 64, 2 M) scans, batch-1 inference) nz_scan_fwd runs chunk-parallel -- aggregate pass, combine, final pass -- instead
 * of handing the state from chunk to chunk, provided `workspace_bytes` >= this value (else it uses the chained scheme).
 * Equals nz_scan_workspace_bytes() for every other shape.  Results are bit-identical either way. */
int64_t nz_scan_workspace_bytes_cp(const NzScanDesc* desc);

/* Bytes of the fine-checkpoint buffer xf this problem can use, or 0 when the row-per-lane backward does not apply to it
 * (needs d_state 16, groups of a multiple of 32 rows, TMA-expressible operands: 16-byte aligned bases and strides, rows a
 * whole number of 128-byte lines).  Only the problem fields, u / delta / B / C / z pointers and strides are read. */
int64_t nz_scan_fine_bytes(const NzScanDesc* desc);

/* Scratch size for nz_scan_bwd: nz_scan_workspace_bytes() plus, when desc->xf is set and the row-per-lane backward
 * applies, the per-chunk aggregates of its chunk-parallel decomposition. */
int64_t nz_scan_workspace_bytes_bwd(const NzScanDesc* desc);

/* 1 when nz_scan_bwd on this problem OVERWRITES dB / dC (every element has a single owner tile, or the call zero-fills
 * them itself before it accumulates -- the row-per-lane backward's aggregate pass does that for free), so the caller
 * need not zero them; 0 when it accumulates with atomics into caller-zeroed buffers. */
int nz_scan_bwd_overwrites_dbc(const NzScanDesc* desc);

/* Number of NZ_CHUNK-long chunks (second-to-last extent of the checkpoint tensor x). */
int64_t nz_scan_num_chunks(int64_t seqlen);

/* replaces selective_scan_cuda.fwd (selective_scan_interface.py:37) */
int nz_scan_fwd(const NzScanDesc* desc, void* stream);

/* replaces selective_scan_cuda.bwd (selective_scan_interface.py:62) */
int nz_scan_bwd(const NzScanDesc* desc, void* stream);

/*
 * CrossScan: x (batch, dim, *spatial) -> xs (batch, K, dim, L), K = 2 * nspatial orders:
 *   nspatial == 2 (H, W):    k0 row-major, k1 column-major, k2/k3 their L-flips  (m2net.py:175-177)
 *   nspatial == 3 (Z, H, W): k0 "z h w", k1 "w z h", k2 "h w z", k3..5 flips     (ssnd2net.py:250-255)
 * Pure data movement, bit-exact.  x and xs are contiguous; dtype is the element type of both.
 */
int nz_cross_scan(const void* x, void* xs, int32_t dtype, int32_t batch, int32_t dim, int32_t nspatial,
                  const int64_t* spatial, void* stream);

/*
 * CrossMerge: out_y (batch, K, dim, L) fp32 -> y (batch, dim, L) fp32 in row-major spatial order,
 * summed in the reference's association order (m2net.py:202-206 + :218; ssnd2net.py:286-298).
 * mode 0 = reference (3-D: reproduces the reference's reuse of direction 1 / 4 and its ignoring of
 * directions 2 / 5 bit-exactly), mode 1 = "fixed" 3-D merge (never used for parity).
 */
int nz_cross_merge(const float* out_y, float* y, int32_t batch, int32_t dim, int32_t nspatial,
                   const int64_t* spatial, int32_t mode, void* stream);

/* Adjoint of nz_cross_merge: dy (batch, dim, L) -> d_out_y (batch, K, dim, L).  (The adjoint of
 * nz_cross_scan is nz_cross_merge with mode 1, so it needs no entry point of its own.) */
int nz_cross_merge_bwd(const float* dy, float* d_out_y, int32_t batch, int32_t dim, int32_t nspatial,
                       const int64_t* spatial, int32_t mode, void* stream);

/*
 * Folded 2-D CrossScan (the data-movement half of nz_ss2d below): x (batch, dim, H, W) -> xs2 (batch, 2, dim, L) =
 * {row-major walk (a copy), column-major walk}.  The two L-flipped directions of m2net.py:176 are not materialised: the
 * scan walks these two arrays backwards (NzScanDesc::rev_mask).  nz_cross_merge_pair is the adjoint:
 * dx[p] = dxs2[:, 0][p] + dxs2[:, 1][t(p)] (one fp32 add, rounded to dtype).
 */
int nz_cross_scan_pair(const void* x, void* xs2, int32_t dtype, int32_t batch, int32_t dim, int32_t H, int32_t W,
                       void* stream);
int nz_cross_merge_pair(const void* dxs2, void* dx, int32_t dtype, int32_t batch, int32_t dim, int32_t H, int32_t W,
                        void* stream);

/*
 * Depthwise causal conv1d (+ SiLU) of the 1-D Mamba block: out[b,d,l] = act(bias[d] + sum_k w[d,k] x[b,d,l-(W-1)+k])
 * with x = 0 left of the sequence -- `self.act(self.conv1d(x)[..., :seqlen])`, mamba_simple.py:316-317, which the
 * reference's fast path takes from causal_conv1d_cuda.causal_conv1d_fwd / _bwd (selective_scan_interface.py:177,
 * :247-252).  W <= 4.  Strides in elements, innermost stride 1.  dx is written contiguous (batch, dim, L);
 * dweight (dim, W) and dbias (dim) are ACCUMULATED INTO (caller zeroes), either may be NULL.
 */
typedef struct NzConv1dDesc {
  int32_t batch, dim, width, dtype; /* dtype of x / out / dout / dx: NZ_F32 / NZ_BF16 / NZ_F16 */
  int64_t seqlen;
  int32_t silu;                     /* 1: SiLU activation fused, 0: plain convolution */
  int32_t reverse;                  /* 1: the convolution of the L-flipped sequence, written un-flipped:
                                       out[l] = act(bias + sum_k w[k] x[l + (W-1) - k]), x = 0 right of the sequence
                                       (what `conv1d(x.flip(-1)).flip(-1)` computes: mamba_simple.py:250-262,
                                       mamba_nd2net.py:638-641) -- the flipped copies are never made */
  const void* x;                    /* (batch, dim, L) */
  const float* weight;              /* (dim, W) fp32 */
  const float* bias;                /* (dim) fp32 or NULL */
  void* out;                        /* forward output (batch, dim, L) */
  const void* dout;                 /* backward input */
  void* dx;                         /* backward output, contiguous */
  float* dweight;
  float* dbias;
  int64_t x_stride[2], out_stride[2], dout_stride[2]; /* batch, dim */
} NzConv1dDesc;
int nz_causal_conv1d_fwd(const NzConv1dDesc* desc, void* stream);
int nz_causal_conv1d_bwd(const NzConv1dDesc* desc, void* stream);
int64_t nz_sizeof_conv1d_desc(void);

/*
 * Weight gradient of SS2D's per-direction projections (the einsums at m2net.py:179 and :182, whose autograd
 * backward the reference leaves to a library GEMM):
 *     dW[k, m, n] += sum_{b, l} G[b, k, m, l] * X[b, k, n, l]
 * x_proj: G = d x_dbl (B, K, R+2N, L), X = xs (B, K, D, L), dW (K, R+2N, D);
 * dt_proj: G = d dts (B, K, D, L), X = the dt rows of x_dbl (B, K, R, L) as the strided split view, dW (K, D, R).
 * g_stride / x_stride: element strides of (batch, direction, row); the sequence stride is 1.  dW is fp32,
 * contiguous and ACCUMULATED INTO (caller zeroes).  M * N <= 10240.
 */
int nz_proj_wgrad(const void* G, const void* X, float* dW, int32_t g_dtype, int32_t x_dtype, int32_t batch, int32_t K,
                  int32_t M, int32_t N, int64_t L, const int64_t* g_stride, const int64_t* x_stride, void* stream);

/*
 * Channels-last LayerNorm over the last axis of a contiguous (rows, C) array -- ln_1 / out_norm of every VSS block
 * (m2net.py:524, :220) and the norms of patch merge / expand (:241, :286-290); the reference calls nn.LayerNorm.
 * x / dx have element type in_dtype, y / dy out_dtype (equal, or one of the two fp32: under autocast the reference
 * normalises in fp32 and the consumer casts); gamma / beta (may be NULL) and the saved statistics mean / rstd (rows
 * each) are fp32; biased variance, y = (x - mean) * rstd * gamma + beta.  dgamma / dbeta (C each, may be NULL) are
 * ACCUMULATED INTO.  Supported: C / E a power of two, C <= 32 * 32 with E = 4 (fp32 involved) or 8 elements;
 * nz_layernorm_supported tells (M2Net: C = 16 .. 1024).
 */
int nz_layernorm_supported(int32_t C, int32_t in_dtype, int32_t out_dtype);
int nz_layernorm_fwd(const void* x, const float* gamma, const float* beta, void* y, float* mean, float* rstd,
                     int64_t rows, int32_t C, int32_t in_dtype, int32_t out_dtype, float eps, void* stream);
int nz_layernorm_bwd(const void* dy, const void* x, const float* mean, const float* rstd, const float* gamma, void* dx,
                     float* dgamma, float* dbeta, int64_t rows, int32_t C, int32_t in_dtype, int32_t out_dtype,
                     void* stream);

/*
 * SS2D's depthwise 3x3 convolution (padding 1, stride 1) fused with its SiLU -- `self.act(self.conv2d(x))`,
 * m2net.py:69-77 and :214-215; the reference calls nn.Conv2d(groups = d_inner) + nn.SiLU.  x, y, dy, dx: contiguous
 * (batch, dim, H, W) of `dtype`; weight (dim, 1, 3, 3) and bias (dim, may be NULL) fp32; silu = 0 gives the plain
 * convolution.  The backward recomputes the pre-activation; dweight (dim, 3, 3) / dbias (dim), either may be NULL, are
 * ACCUMULATED INTO.
 */
int nz_dwconv3x3_fwd(const void* x, const float* weight, const float* bias, void* y, int32_t dtype, int32_t batch,
                     int32_t dim, int32_t H, int32_t W, int32_t silu, void* stream);
int nz_dwconv3x3_bwd(const void* x, const void* dy, const float* weight, const float* bias, void* dx, float* dweight,
                     float* dbias, int32_t dtype, int32_t batch, int32_t dim, int32_t H, int32_t W, int32_t silu,
                     void* stream);

/*
 * SS2D's post-scan chain fused (2-D, 4 directions): CrossMerge -> (B, L, D) -> LayerNorm over D -> * SiLU(z)
 * (m2net.py:202-206, :218-221).  out_y (batch, 4, D, L) fp32 as nz_scan_fwd writes it; z rows of D contiguous elements
 * (z_stride = {batch, position} element strides: the chunk view of in_proj's output goes in as it is); gamma / beta fp32
 * (may be NULL); out (batch, L, D) contiguous of out_dtype.  The forward also writes what the backward needs: the merged
 * y (batch, L, D) fp32 -- bit-identical to nz_cross_merge -- and the row statistics mean / rstd (batch * L).
 * Backward: dout (batch, L, D) of out_dtype -> d_out_y (batch, 4, D, L) of grad_dtype (fp32, or the scan's 16-bit operand
 * dtype so that nz_scan_bwd reads it directly), dz (batch, L, D) contiguous of z_dtype, dgamma / dbeta ACCUMULATED INTO.
 * D: 32, 64, 128 or 256 (nz_ss2d_epilogue_supported) -- d_model 16 .. 128, every SS2D of M2Net.
 */
int nz_ss2d_epilogue_supported(int32_t D);
int nz_ss2d_epilogue_fwd(const float* out_y, const void* z, const int64_t* z_stride, const float* gamma, const float* beta,
                         void* out, float* y_merged, float* mean, float* rstd, int32_t z_dtype, int32_t out_dtype,
                         int32_t batch, int32_t D, int32_t H, int32_t W, float eps, void* stream);
int nz_ss2d_epilogue_bwd(const void* dout, const float* y_merged, const float* mean, const float* rstd, const void* z,
                         const int64_t* z_stride, const float* gamma, const float* beta, void* d_out_y, void* dz,
                         float* dgamma, float* dbeta, int32_t z_dtype, int32_t out_dtype, int32_t grad_dtype,
                         int32_t batch, int32_t D, int32_t H, int32_t W, void* stream);

/* The same two kernels on the FOLDED direction layout: out_y / d_out_y (batch, 4, D, L) hold {row-major forward, row-major
 * backward, column-major forward, column-major backward}, none of them flipped -- what nz_scan_fwd / nz_scan_bwd write and
 * read with ngroups = 4, rev_mask = 0b1010, u_gdiv = 2 over xs2 of nz_cross_scan_pair.  Sum order unchanged:
 * ((y[0] + y[1]) + T y[2]) + T y[3] = ((y0 + flip y2) + T y1) + T flip y3 of m2net.py:202-206, :218, bit for bit. */
int nz_ss2d_epilogue_fwd_folded(const float* out_y, const void* z, const int64_t* z_stride, const float* gamma,
                                const float* beta, void* out, float* y_merged, float* mean, float* rstd, int32_t z_dtype,
                                int32_t out_dtype, int32_t batch, int32_t D, int32_t H, int32_t W, float eps, void* stream);
int nz_ss2d_epilogue_bwd_folded(const void* dout, const float* y_merged, const float* mean, const float* rstd, const void* z,
                                const int64_t* z_stride, const float* gamma, const float* beta, void* d_out_y, void* dz,
                                float* dgamma, float* dbeta, int32_t z_dtype, int32_t out_dtype, int32_t grad_dtype,
                                int32_t batch, int32_t D, int32_t H, int32_t W, void* stream);

/*
 * Sliding-window accumulate: the per-tile arithmetic of the predictor's inner loop as one launch --
 *   mirror average   prediction = net(x); prediction += flip(net(flip(x, c)), c) ...; prediction /= passes
 *                                                           (inference/predict_from_raw_data.py:549-565)
 *   accumulate       prediction = prediction.to(results dtype) * gaussian; predicted_logits[sl] += prediction;
 *                    n_predictions[sl[1:]] += gaussian                                        (:617-623)
 * with a rounding after every one of those steps, so the accumulators are bit-identical to the PyTorch expressions.
 *   pred     (nmirror * ntiles, heads, t0, t1, t2) contiguous in pred_dtype: pass m of tile t at index m * ntiles + t;
 *            pass 0 is un-mirrored, pass m was computed on the tile flipped along the axes in mirror_masks[m] (bit a =
 *            tile axis a) and is read back through the same flip.  nmirror is 1, 2, 4 or 8.
 *   gaussian (t0, t1, t2) in res_dtype (all ones when the reference's use_gaussian is off)
 *   logits   (heads, v0, v1, v2), n_pred (v0, v1, v2) in res_dtype, updated in place
 *   offsets  (ntiles, 3) HOST array: origin of each tile in the volume.  A 2-D tile of a 3-D volume has t0 = 1.
 * The tiles of ONE call must not overlap in the volume (each element has one owner thread); overlapping tiles go in
 * separate calls, which the stream orders.  ntiles <= NZ_SW_MAX_TILES.
 */
#define NZ_SW_MAX_TILES 32
#define NZ_SW_MAX_MIRRORS 8
int nz_sw_accumulate(const void* pred, int32_t pred_dtype, int32_t nmirror, const int32_t* mirror_masks, int32_t ntiles,
                     int32_t heads, const int64_t* tile, const void* gaussian, int32_t res_dtype, void* logits,
                     void* n_pred, const int64_t* vol, const int64_t* offsets, void* stream);

/*
 * Host-buffer entry points (what a non-PyTorch caller of the reference's operator would bind):
 * every pointer in `desc` is a HOST pointer, strides as above; the call stages host -> device,
 * runs nz_scan_fwd (and nz_scan_bwd when desc->dout != NULL) and copies the results back,
 * synchronising `stream` before returning.  Used by bench.py for the end-to-end number.
 */
int nz_scan_fwd_bwd_host(const NzScanDesc* desc_host, void* stream);

/* Bind the calling thread to `device` inside this library's CUDA runtime instance (the library
 * links cudart statically; one process per GPU calls this once with its LOCAL_RANK device). */
int nz_set_device(int device);

/* sizeof(NzScanDesc) as compiled into the library (lets FFI bindings verify their struct mirror) */
int64_t nz_sizeof_scan_desc(void);

/* Tools only: when non-NULL, forward launches write 6 u64 words per tile (ticket, SM, start / loop /
 * wait / end times in ns) into this device buffer (tools/trace_tiles.py).  NULL switches it off. */
void nz_debug_set_trace(void* device_buffer);

const char* nz_last_error(void);
int nz_abi_version(void);
/* number of kernel launches this library has issued in the calling process (bench.py gpu_launches) */
int64_t nz_launch_count(void);

#ifdef __cplusplus
}
#endif
#endif /* NNUZOO_B200_H */
