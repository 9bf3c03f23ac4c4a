// cross_kernels.cu -- CrossScan / CrossMerge sequence permutations of SS2D (2-D, 4 directions) and
// SSND (3-D, 6 directions) as stand-alone, bit-exact data-movement kernels.
//
// Reference statements (pure PyTorch tensor shuffles, materialised copies):
//   scan  2-D: nnunetv2/nets/m2net.py:175-177       3-D: nnunetv2/nets/ssnd2net.py:250-255
//   merge 2-D: nnunetv2/nets/m2net.py:202-206, :218 3-D: nnunetv2/nets/ssnd2net.py:286-298
//
// Index maps (l = position in the scanned sequence, src = row-major spatial offset):
//   2-D  k0: src = l                      k1: src = (l % H) * W + l / H           k2,k3: same at L-1-l
//   3-D  k0: src = l                      k1 "w z h": l = (w*Z + z)*H + h         k2 "h w z": l = (h*W + w)*Z + z
//        k3..5: same at L-1-l
// The merge sums in the reference's left-to-right order with separate fp32 adds, so results are
// bit-identical to the PyTorch expressions.  One thread per output element, writes coalesced.
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdlib.h>

#include "../../include/nnuzoo_b200.h"

namespace nz {
void count_launch(int n);

struct Dims {
  long Z, H, W, L;
  int nsp;
};

// spatial offset visited at step l of forward direction kf (0..nsp-1)
__device__ __forceinline__ long src_of(const Dims& s, int kf, long l) {
  if (kf == 0) return l;
  if (s.nsp == 2) {  // column-major walk
    const long h = l % s.H, w = l / s.H;
    return h * s.W + w;
  }
  if (kf == 1) {  // "w z h"
    const long h = l % s.H, z = (l / s.H) % s.Z, w = l / (s.H * s.Z);
    return (z * s.H + h) * s.W + w;
  }
  // kf == 2, "h w z"
  const long z = l % s.Z, w = (l / s.Z) % s.W, h = l / (s.Z * s.W);
  return (z * s.H + h) * s.W + w;
}

// step of forward direction kf that visits spatial offset src (inverse of src_of)
__device__ __forceinline__ long step_of(const Dims& s, int kf, long src) {
  if (kf == 0) return src;
  if (s.nsp == 2) {
    const long w = src % s.W, h = src / s.W;
    return w * s.H + h;
  }
  const long w = src % s.W, h = (src / s.W) % s.H, z = src / (s.W * s.H);
  if (kf == 1) return (w * s.Z + z) * s.H + h;
  return (h * s.W + w) * s.Z + z;
}

template <typename T>
__global__ void cross_scan_kernel(const T* __restrict__ x, T* __restrict__ xs, long rows /*batch*dim*/, int dim,
                                  Dims s) {
  const int K = 2 * s.nsp;
  const long total = rows * K * s.L;
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    const long l = i % s.L;
    const long q = i / s.L;  // (b, k, d) flattened
    const int dd = (int)(q % dim);
    const int k = (int)((q / dim) % K);
    const long b = q / ((long)dim * K);
    const int kf = k % s.nsp;
    const long ll = k >= s.nsp ? s.L - 1 - l : l;
    xs[i] = x[(b * dim + dd) * s.L + src_of(s, kf, ll)];
  }
}

// out_y (b, K, d, L) -> y (b, d, L)
__global__ void cross_merge_kernel(const float* __restrict__ oy, float* __restrict__ y, long rows, int dim, Dims s,
                                   int mode) {
  const int K = 2 * s.nsp;
  const long total = rows * s.L;
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    const long l = i % s.L;
    const long q = i / s.L;
    const int dd = (int)(q % dim);
    const long b = q / dim;
    const float* base = oy + (b * K * dim + dd) * s.L;  // direction k at base + k*dim*L
    const long ks = (long)dim * s.L;
    float acc = base[l];                                      // out_y[:, 0]
    acc = acc + base[s.nsp * ks + (s.L - 1 - l)];             // + inv_y[:, 0]
    if (s.nsp == 2) {
      const long j = step_of(s, 1, l);
      acc = acc + base[ks + j];                               // + wh_y
      acc = acc + base[3 * ks + (s.L - 1 - j)];               // + invwh_y
    } else {
      const long j1 = step_of(s, 1, l);
      acc = acc + base[ks + j1];                              // + y_wzh        (ssnd2net.py:291)
      acc = acc + base[4 * ks + (s.L - 1 - j1)];              // + inv_y_wzh    (:292)
      if (mode == 0) {
        // :295-296 -- direction 1 again, through a (W, Z, H)-shaped view whose axes are relabelled
        // "h w z": output flat index l decomposes over sizes (H, W, Z) as (i2, i0, i1).
        const long i1 = l % s.Z, i0 = (l / s.Z) % s.W, i2 = l / (s.Z * s.W);
        const long j2 = (i0 * s.Z + i1) * s.H + i2;
        acc = acc + base[ks + j2];
        acc = acc + base[4 * ks + (s.L - 1 - j2)];
      } else {
        const long j2 = step_of(s, 2, l);
        acc = acc + base[2 * ks + j2];
        acc = acc + base[5 * ks + (s.L - 1 - j2)];
      }
    }
    y[i] = acc;
  }
}

// adjoint of cross_merge: dy (b, d, L) -> d_out_y (b, K, d, L)
__global__ void cross_merge_bwd_kernel(const float* __restrict__ dy, float* __restrict__ doy, long rows, int dim,
                                       Dims s, int mode) {
  const int K = 2 * s.nsp;
  const long total = rows * K * s.L;
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    const long l = i % s.L;
    const long q = i / s.L;
    const int dd = (int)(q % dim);
    const int k = (int)((q / dim) % K);
    const long b = q / ((long)dim * K);
    const float* g = dy + (b * dim + dd) * s.L;
    const int kf = k % s.nsp;
    const long ll = k >= s.nsp ? s.L - 1 - l : l;  // step index in the un-flipped direction
    float v;
    if (s.nsp == 2 || mode != 0) {
      v = g[src_of(s, kf, ll)];
    } else if (kf == 0) {
      v = g[ll];
    } else if (kf == 1) {
      // contributions through :291/:292 and through the relabelled view of :295/:296
      const long i2 = ll % s.H, i1 = (ll / s.H) % s.Z, i0 = ll / (s.H * s.Z);
      v = g[src_of(s, 1, ll)] + g[(i2 * s.W + i0) * s.Z + i1];
    } else {
      v = 0.f;  // directions 2 and 5 never reach the reference's output
    }
    doy[i] = v;
  }
}


// ---------------------------------------------------------------------------------------------------------------
// 2-D fast path: the same maps, tiled.  A CTA owns a 32 x 32 (h, w) tile of one (batch, channel) plane and moves it
// through shared memory, so both the row-major walks (k0, k2: contiguous in w) and the column-major walks (k1, k3:
// contiguous in h) touch whole 128-byte lines; the one-thread-per-element kernels above read the transposed
// directions with a stride of H or W elements (one useful element per 32-byte sector).  Same values, same fp32
// association order, so still bit-identical to the reference expressions.
constexpr int kCT = 32;

// Direction slots of the (b, K, d, L) array a tiled pass touches; -1 = not written / read by this pass.
//   copy / copy_flip : the row-major walk of the (H, W) matrix and its L-flip
//   tr / tr_flip     : the column-major walk t(p) = w * H + h and its L-flip
//   zero_a / zero_b  : slots this pass fills with zeros (the 3-D reference merge never reads directions 2 / 5)
// The 3-D walks are the same two maps on re-factored matrices: "w z h" (ssnd2net.py:251) is the column-major walk of the
// (Z*H) x W matrix, "h w z" (:252) of the Z x (H*W) matrix, and the reference's second un-permute of direction 1
// (:295-296) of the H x (W*Z) matrix -- so every 3-D direction moves through the same 32 x 32 shared-memory tiles.
struct Slots {
  int K, copy, copy_flip, tr, tr_flip, zero_a, zero_b;
  int accumulate;  // scan: add into tr / tr_flip instead of overwriting (fp32 only); merge: start from y instead of copy
};

template <typename T>
__device__ __forceinline__ void put(T* p, T v, bool acc) { *p = v; (void)acc; }
template <>
__device__ __forceinline__ void put<uint32_t>(uint32_t* p, uint32_t v, bool acc) {
  if (acc) {
    float* f = reinterpret_cast<float*>(p);
    *f = *f + __uint_as_float(v);
  } else {
    *p = v;
  }
}

template <typename T>
__global__ void __launch_bounds__(256) cross_scan2d_tiled_kernel(const T* __restrict__ x, T* __restrict__ xs, long planes,
                                                                 int dim, long H, long W, int tiles_w, int tiles_h,
                                                                 const Slots sl) {
  __shared__ T tile[kCT][kCT + 1];
  const long L = H * W;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 32 x 8
  const bool acc = sl.accumulate != 0;
  for (long blk = blockIdx.x;; blk += gridDim.x) {
    const long plane = blk / ((long)tiles_w * tiles_h);
    const int tr = (int)(blk % ((long)tiles_w * tiles_h));
    if (plane >= planes) break;
    const long h0 = (long)(tr / tiles_w) * kCT, w0 = (long)(tr % tiles_w) * kCT;
    const long b = plane / dim;
    const int dd = (int)(plane % dim);
    const T* src = x + plane * L;
    auto slot = [&](int k) { return k < 0 ? (T*)nullptr : xs + ((b * sl.K + k) * dim + dd) * L; };
    T* o0 = slot(sl.copy);
    T* o1 = slot(sl.tr);
    T* o2 = slot(sl.copy_flip);
    T* o3 = slot(sl.tr_flip);
    T* z0 = slot(sl.zero_a);
    T* z1 = slot(sl.zero_b);
#pragma unroll
    for (int r = 0; r < kCT; r += 8) {
      const long h = h0 + ty + r, w = w0 + tx;
      if (h < H && w < W) {
        const T v = src[h * W + w];
        tile[ty + r][tx] = v;
        if (o0) o0[h * W + w] = v;
        if (o2) o2[L - 1 - (h * W + w)] = v;
        if (z0) z0[h * W + w] = T(0);
        if (z1) z1[h * W + w] = T(0);
      }
    }
    __syncthreads();
#pragma unroll
    for (int r = 0; r < kCT; r += 8) {
      const long w = w0 + ty + r, h = h0 + tx;
      if (h < H && w < W) {
        const T v = tile[tx][ty + r];
        if (o1) put<T>(o1 + (w * H + h), v, acc);
        if (o3) put<T>(o3 + (L - 1 - (w * H + h)), v, acc);
      }
    }
    __syncthreads();
  }
}

// out_y (b, K, d, L) -> y (b, d, L): ((start + tr[t(p)]) + tr_flip[L-1-t(p)]), t(p) = w * H + h, where
// start = copy[p] + copy_flip[L-1-p], or y[p] itself when sl.accumulate (the second pass of the 3-D merge).  Separate
// fp32 adds in the reference's left-to-right order (m2net.py:218, ssnd2net.py:298).
__global__ void __launch_bounds__(256) cross_merge2d_tiled_kernel(const float* __restrict__ oy, float* __restrict__ y,
                                                                  long planes, int dim, long H, long W, int tiles_w,
                                                                  int tiles_h, const Slots sl) {
  __shared__ float t1[kCT][kCT + 1], t3[kCT][kCT + 1];
  const long L = H * W;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  for (long blk = blockIdx.x;; blk += gridDim.x) {
    const long plane = blk / ((long)tiles_w * tiles_h);
    const int tr = (int)(blk % ((long)tiles_w * tiles_h));
    if (plane >= planes) break;
    const long h0 = (long)(tr / tiles_w) * kCT, w0 = (long)(tr % tiles_w) * kCT;
    const long b = plane / dim;
    const int dd = (int)(plane % dim);
    auto slot = [&](int k) { return oy + ((b * sl.K + (k < 0 ? 0 : k)) * dim + dd) * L; };
    const float* i0 = slot(sl.copy);
    const float* i1 = slot(sl.tr);
    const float* i2 = slot(sl.copy_flip);
    const float* i3 = slot(sl.tr_flip);
#pragma unroll
    for (int r = 0; r < kCT; r += 8) {  // column-major directions: contiguous in h
      const long w = w0 + ty + r, h = h0 + tx;
      if (h < H && w < W) {
        t1[ty + r][tx] = i1[w * H + h];
        t3[ty + r][tx] = i3[L - 1 - (w * H + h)];
      }
    }
    __syncthreads();
#pragma unroll
    for (int r = 0; r < kCT; r += 8) {
      const long h = h0 + ty + r, w = w0 + tx;
      if (h < H && w < W) {
        const long p = h * W + w;
        float acc;
        if (sl.accumulate) {
          acc = y[plane * L + p];
        } else {
          acc = i0[p];
          acc = acc + i2[L - 1 - p];
        }
        acc = acc + t1[tx][ty + r];
        acc = acc + t3[tx][ty + r];
        y[plane * L + p] = acc;
      }
    }
    __syncthreads();
  }
}

// Folded 2-D CrossScan: only the two ARRAYS the four directions walk -- xs2 (b, 2, d, L) = {x row-major (a copy), x
// column-major} -- the two flipped directions read them backwards inside the scan (NzScanDesc::rev_mask) instead of
// getting flipped copies.  Same tiling as above.
template <typename T>
__global__ void __launch_bounds__(256) cross_scan2d_pair_kernel(const T* __restrict__ x, T* __restrict__ xs, long planes,
                                                                int dim, long H, long W, int tiles_w, int tiles_h) {
  __shared__ T tile[kCT][kCT + 1];
  const long L = H * W;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 32 x 8
  for (long blk = blockIdx.x;; blk += gridDim.x) {
    const long plane = blk / ((long)tiles_w * tiles_h);
    const int tr = (int)(blk % ((long)tiles_w * tiles_h));
    if (plane >= planes) break;
    const long h0 = (long)(tr / tiles_w) * kCT, w0 = (long)(tr % tiles_w) * kCT;
    const long b = plane / dim;
    const int dd = (int)(plane % dim);
    const T* src = x + plane * L;
    T* o0 = xs + ((b * 2 + 0) * dim + dd) * L;
    T* o1 = xs + ((b * 2 + 1) * dim + dd) * L;
#pragma unroll
    for (int r = 0; r < kCT; r += 8) {
      const long h = h0 + ty + r, w = w0 + tx;
      if (h < H && w < W) {
        const T v = src[h * W + w];
        tile[ty + r][tx] = v;
        o0[h * W + w] = v;
      }
    }
    __syncthreads();
#pragma unroll
    for (int r = 0; r < kCT; r += 8) {
      const long w = w0 + ty + r, h = h0 + tx;
      if (h < H && w < W) o1[w * H + h] = tile[tx][ty + r];
    }
    __syncthreads();
  }
}

// its adjoint: dxs2 (b, 2, d, L) -> dx (b, d, L) = dxs2[:, 0][p] + dxs2[:, 1][t(p)], one fp32 add, rounded to T
template <typename T>
__device__ __forceinline__ float pair_ld(const T* p);
template <>
__device__ __forceinline__ float pair_ld<float>(const float* p) { return *p; }
template <>
__device__ __forceinline__ float pair_ld<__nv_bfloat16>(const __nv_bfloat16* p) { return __bfloat162float(*p); }
template <>
__device__ __forceinline__ float pair_ld<__half>(const __half* p) { return __half2float(*p); }
template <typename T>
__device__ __forceinline__ void pair_st(T* p, float v);
template <>
__device__ __forceinline__ void pair_st<float>(float* p, float v) { *p = v; }
template <>
__device__ __forceinline__ void pair_st<__nv_bfloat16>(__nv_bfloat16* p, float v) { *p = __float2bfloat16_rn(v); }
template <>
__device__ __forceinline__ void pair_st<__half>(__half* p, float v) { *p = __float2half_rn(v); }

template <typename T>
__global__ void __launch_bounds__(256) cross_merge2d_pair_kernel(const T* __restrict__ g, T* __restrict__ dx, long planes,
                                                                 int dim, long H, long W, int tiles_w, int tiles_h) {
  __shared__ float t1[kCT][kCT + 1];
  const long L = H * W;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  for (long blk = blockIdx.x;; blk += gridDim.x) {
    const long plane = blk / ((long)tiles_w * tiles_h);
    const int tr = (int)(blk % ((long)tiles_w * tiles_h));
    if (plane >= planes) break;
    const long h0 = (long)(tr / tiles_w) * kCT, w0 = (long)(tr % tiles_w) * kCT;
    const long b = plane / dim;
    const int dd = (int)(plane % dim);
    const T* i0 = g + ((b * 2 + 0) * dim + dd) * L;
    const T* i1 = g + ((b * 2 + 1) * dim + dd) * L;
#pragma unroll
    for (int r = 0; r < kCT; r += 8) {
      const long w = w0 + ty + r, h = h0 + tx;
      if (h < H && w < W) t1[ty + r][tx] = pair_ld<T>(i1 + w * H + h);
    }
    __syncthreads();
#pragma unroll
    for (int r = 0; r < kCT; r += 8) {
      const long h = h0 + ty + r, w = w0 + tx;
      if (h < H && w < W) {
        const long p = h * W + w;
        pair_st<T>(dx + plane * L + p, pair_ld<T>(i0 + p) + t1[tx][ty + r]);
      }
    }
    __syncthreads();
  }
}

static int tiled_grid(long planes, long H, long W, int* tw, int* th) {
  *tw = (int)((W + kCT - 1) / kCT), *th = (int)((H + kCT - 1) / kCT);
  const long blocks = planes * *tw * *th, cap = 148L * 32;
  return (int)(blocks < cap ? blocks : cap);
}

static int launch_cfg(long total, int* grid) {
  const int threads = 256;
  long blocks = (total + threads - 1) / threads;
  const long cap = 148L * 16;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  *grid = (int)blocks;
  return threads;
}

static bool make_dims(int nsp, const int64_t* sp, Dims* d) {
  if (!sp || (nsp != 2 && nsp != 3)) return false;
  d->nsp = nsp;
  d->Z = nsp == 3 ? sp[0] : 1;
  d->H = sp[nsp - 2];
  d->W = sp[nsp - 1];
  if (d->Z < 1 || d->H < 1 || d->W < 1) return false;
  d->L = d->Z * d->H * d->W;
  return true;
}

}  // namespace nz

extern "C" {

int nz_cross_scan(const void* x, void* xs, int32_t dtype, int32_t batch, int32_t dim, int32_t nspatial,
                  const int64_t* spatial, void* stream) {
  nz::Dims s;
  if (!x || !xs || batch < 1 || dim < 1 || !nz::make_dims(nspatial, spatial, &s)) return NZ_EINVAL;
  const long rows = (long)batch * dim;
  int grid;
  const int threads = nz::launch_cfg(rows * 2 * nspatial * s.L, &grid);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if ((dtype == NZ_F32 || dtype == NZ_BF16 || dtype == NZ_F16) && !getenv("NZ_CROSS_GENERIC")) {
    // 2-D: one tiled pass.  3-D: two -- the (Z*H) x W matrix gives k0, k1 "w z h" and their flips k3, k4, the
    // Z x (H*W) matrix gives k2 "h w z" and k5 (see Slots).
    auto pass = [&](long Hm, long Wm, const nz::Slots& sl) {
      int tw, th;
      const int g = nz::tiled_grid(rows, Hm, Wm, &tw, &th);
      if (dtype == NZ_F32)
        nz::cross_scan2d_tiled_kernel<uint32_t><<<g, 256, 0, st>>>(static_cast<const uint32_t*>(x),
                                                                   static_cast<uint32_t*>(xs), rows, dim, Hm, Wm, tw, th, sl);
      else
        nz::cross_scan2d_tiled_kernel<uint16_t><<<g, 256, 0, st>>>(static_cast<const uint16_t*>(x),
                                                                   static_cast<uint16_t*>(xs), rows, dim, Hm, Wm, tw, th, sl);
      nz::count_launch(1);
    };
    if (nspatial == 2) {
      pass(s.H, s.W, nz::Slots{4, 0, 2, 1, 3, -1, -1, 0});
    } else {
      pass(s.Z * s.H, s.W, nz::Slots{6, 0, 3, 1, 4, -1, -1, 0});
      pass(s.Z, s.H * s.W, nz::Slots{6, -1, -1, 2, 5, -1, -1, 0});
    }
    return cudaGetLastError() == cudaSuccess ? NZ_OK : NZ_ECUDA;
  }
  if (dtype == NZ_F32)
    nz::cross_scan_kernel<uint32_t><<<grid, threads, 0, st>>>(static_cast<const uint32_t*>(x),
                                                              static_cast<uint32_t*>(xs), rows, dim, s);
  else if (dtype == NZ_BF16 || dtype == NZ_F16)
    nz::cross_scan_kernel<uint16_t><<<grid, threads, 0, st>>>(static_cast<const uint16_t*>(x),
                                                              static_cast<uint16_t*>(xs), rows, dim, s);
  else
    return NZ_EINVAL;
  nz::count_launch(1);
  return cudaGetLastError() == cudaSuccess ? NZ_OK : NZ_ECUDA;
}

int nz_cross_merge(const float* out_y, float* y, int32_t batch, int32_t dim, int32_t nspatial,
                   const int64_t* spatial, int32_t mode, void* stream) {
  nz::Dims s;
  if (!out_y || !y || batch < 1 || dim < 1 || !nz::make_dims(nspatial, spatial, &s)) return NZ_EINVAL;
  const long rows = (long)batch * dim;
  if (!getenv("NZ_CROSS_GENERIC")) {
    auto pass = [&](long Hm, long Wm, const nz::Slots& sl) {
      int tw, th;
      const int g = nz::tiled_grid(rows, Hm, Wm, &tw, &th);
      nz::cross_merge2d_tiled_kernel<<<g, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(out_y, y, rows, dim, Hm, Wm,
                                                                                           tw, th, sl);
      nz::count_launch(1);
    };
    if (nspatial == 2) {
      pass(s.H, s.W, nz::Slots{4, 0, 2, 1, 3, -1, -1, 0});
    } else {
      // ((((y0 + inv0) + wzh) + inv_wzh) + third) + inv_third  (ssnd2net.py:298): the first four terms on the
      // (Z*H) x W matrix, the last two added onto y by a second pass -- reference mode reads direction 1 / 4 again
      // through the H x (W*Z) matrix (:295-296), "fixed" mode direction 2 / 5 through the Z x (H*W) one
      pass(s.Z * s.H, s.W, nz::Slots{6, 0, 3, 1, 4, -1, -1, 0});
      if (mode == 0) pass(s.H, s.W * s.Z, nz::Slots{6, -1, -1, 1, 4, -1, -1, 1});
      else pass(s.Z, s.H * s.W, nz::Slots{6, -1, -1, 2, 5, -1, -1, 1});
    }
    return cudaGetLastError() == cudaSuccess ? NZ_OK : NZ_ECUDA;
  }
  int grid;
  const int threads = nz::launch_cfg(rows * s.L, &grid);
  nz::cross_merge_kernel<<<grid, threads, 0, reinterpret_cast<cudaStream_t>(stream)>>>(out_y, y, rows, dim, s, mode);
  nz::count_launch(1);
  return cudaGetLastError() == cudaSuccess ? NZ_OK : NZ_ECUDA;
}

int nz_cross_merge_bwd(const float* dy, float* d_out_y, int32_t batch, int32_t dim, int32_t nspatial,
                       const int64_t* spatial, int32_t mode, void* stream) {
  nz::Dims s;
  if (!dy || !d_out_y || batch < 1 || dim < 1 || !nz::make_dims(nspatial, spatial, &s)) return NZ_EINVAL;
  const long rows = (long)batch * dim;
  if (!getenv("NZ_CROSS_GENERIC")) {  // the adjoint of the merge is the scan permutation applied to dy (fp32)
    auto pass = [&](long Hm, long Wm, const nz::Slots& sl) {
      int tw, th;
      const int g = nz::tiled_grid(rows, Hm, Wm, &tw, &th);
      nz::cross_scan2d_tiled_kernel<uint32_t><<<g, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
          reinterpret_cast<const uint32_t*>(dy), reinterpret_cast<uint32_t*>(d_out_y), rows, dim, Hm, Wm, tw, th, sl);
      nz::count_launch(1);
    };
    if (nspatial == 2) {
      pass(s.H, s.W, nz::Slots{4, 0, 2, 1, 3, -1, -1, 0});
    } else if (mode == 0) {
      // reference merge: directions 1 / 4 receive dy twice (two different un-permutes), 2 / 5 nothing
      pass(s.Z * s.H, s.W, nz::Slots{6, 0, 3, 1, 4, 2, 5, 0});
      pass(s.H, s.W * s.Z, nz::Slots{6, -1, -1, 1, 4, -1, -1, 1});
    } else {
      pass(s.Z * s.H, s.W, nz::Slots{6, 0, 3, 1, 4, -1, -1, 0});
      pass(s.Z, s.H * s.W, nz::Slots{6, -1, -1, 2, 5, -1, -1, 0});
    }
    return cudaGetLastError() == cudaSuccess ? NZ_OK : NZ_ECUDA;
  }
  int grid;
  const int threads = nz::launch_cfg(rows * 2 * nspatial * s.L, &grid);
  nz::cross_merge_bwd_kernel<<<grid, threads, 0, reinterpret_cast<cudaStream_t>(stream)>>>(dy, d_out_y, rows, dim, s,
                                                                                          mode);
  nz::count_launch(1);
  return cudaGetLastError() == cudaSuccess ? NZ_OK : NZ_ECUDA;
}


int nz_cross_scan_pair(const void* x, void* xs2, int32_t dtype, int32_t batch, int32_t dim, int32_t H, int32_t W,
                       void* stream) {
  if (!x || !xs2 || batch < 1 || dim < 1 || H < 1 || W < 1) return NZ_EINVAL;
  const long rows = (long)batch * dim;
  int tw, th;
  const int g = nz::tiled_grid(rows, H, W, &tw, &th);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (dtype == NZ_F32)
    nz::cross_scan2d_pair_kernel<uint32_t><<<g, 256, 0, st>>>(static_cast<const uint32_t*>(x), static_cast<uint32_t*>(xs2),
                                                              rows, dim, H, W, tw, th);
  else if (dtype == NZ_BF16 || dtype == NZ_F16)
    nz::cross_scan2d_pair_kernel<uint16_t><<<g, 256, 0, st>>>(static_cast<const uint16_t*>(x), static_cast<uint16_t*>(xs2),
                                                              rows, dim, H, W, tw, th);
  else
    return NZ_EINVAL;
  nz::count_launch(1);
  return cudaGetLastError() == cudaSuccess ? NZ_OK : NZ_ECUDA;
}

int nz_cross_merge_pair(const void* dxs2, void* dx, int32_t dtype, int32_t batch, int32_t dim, int32_t H, int32_t W,
                        void* stream) {
  if (!dxs2 || !dx || batch < 1 || dim < 1 || H < 1 || W < 1) return NZ_EINVAL;
  const long rows = (long)batch * dim;
  int tw, th;
  const int g = nz::tiled_grid(rows, H, W, &tw, &th);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (dtype == NZ_F32)
    nz::cross_merge2d_pair_kernel<float><<<g, 256, 0, st>>>(static_cast<const float*>(dxs2), static_cast<float*>(dx), rows,
                                                            dim, H, W, tw, th);
  else if (dtype == NZ_BF16)
    nz::cross_merge2d_pair_kernel<__nv_bfloat16><<<g, 256, 0, st>>>(static_cast<const __nv_bfloat16*>(dxs2),
                                                                    static_cast<__nv_bfloat16*>(dx), rows, dim, H, W, tw, th);
  else if (dtype == NZ_F16)
    nz::cross_merge2d_pair_kernel<__half><<<g, 256, 0, st>>>(static_cast<const __half*>(dxs2), static_cast<__half*>(dx),
                                                             rows, dim, H, W, tw, th);
  else
    return NZ_EINVAL;
  nz::count_launch(1);
  return cudaGetLastError() == cudaSuccess ? NZ_OK : NZ_ECUDA;
}

}  // extern "C"
