// pipes.cu -- per-SM pipe throughput micro-benchmarks for the scan kernels' cost model (sm_100a).
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/ubench/pipes.bin tools/ubench/pipes.cu
//   ./tools/ubench/pipes.bin
//
// Every test runs one CTA of WARPS warps on every SM, each warp executing ITER trips of an unrolled
// body of independent instructions; the SM-cycle cost per warp-instruction is
//   (clock64 span of the CTA) / (ITER * body * WARPS).
// Results feed DESIGN.md "cost model" (shared-memory wavefronts, SHFL, MUFU, FFMA2).
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <vector>

#define CK(x)                                                                      \
  do {                                                                             \
    cudaError_t e_ = (x);                                                          \
    if (e_ != cudaSuccess) {                                                       \
      printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); \
      return 1;                                                                    \
    }                                                                              \
  } while (0)

constexpr int ITER = 2000;

__device__ __forceinline__ float ex2(float x) {
  float y;
  asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

struct Span {
  long long t0, t1;
};

__device__ __forceinline__ void span_begin(long long& t0) {
  __syncthreads();
  t0 = clock64();
}
__device__ __forceinline__ void span_end(Span* out, long long t0) {
  long long t1 = clock64();
  __shared__ long long smin, smax;
  if (threadIdx.x == 0) {
    smin = 0x7fffffffffffffffLL;
    smax = 0;
  }
  __syncthreads();
  atomicMin(&smin, t0);
  atomicMax(&smax, t1);
  __syncthreads();
  if (threadIdx.x == 0) {
    out[blockIdx.x].t0 = smin;
    out[blockIdx.x].t1 = smax;
  }
}

// mode 0: MUFU.EX2 x8 independent
__global__ void k_mufu(Span* out, float* sink) {
  float v[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) v[i] = -0.001f * (threadIdx.x + i);
  long long t0;
  span_begin(t0);
  for (int it = 0; it < ITER; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = ex2(v[i]);
  }
  span_end(out, t0);
  float s = 0;
  for (int i = 0; i < 8; ++i) s += v[i];
  if (s == 123.f) sink[0] = s;
}

// SHFL x8 independent
__global__ void k_shfl(Span* out, float* sink) {
  float v[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) v[i] = threadIdx.x + i;
  long long t0;
  span_begin(t0);
  for (int it = 0; it < ITER; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = __shfl_up_sync(0xffffffffu, v[i], 1);
  }
  span_end(out, t0);
  float s = 0;
  for (int i = 0; i < 8; ++i) s += v[i];
  if (s == 123.f) sink[0] = s;
}

// MUFU + SHFL interleaved (8 + 8)
__global__ void k_mufu_shfl(Span* out, float* sink) {
  float v[8], w[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    v[i] = -0.001f * (threadIdx.x + i);
    w[i] = threadIdx.x + i;
  }
  long long t0;
  span_begin(t0);
  for (int it = 0; it < ITER; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      v[i] = ex2(v[i]);
      w[i] = __shfl_up_sync(0xffffffffu, w[i], 1);
    }
  }
  span_end(out, t0);
  float s = 0;
  for (int i = 0; i < 8; ++i) s += v[i] + w[i];
  if (s == 123.f) sink[0] = s;
}

// FFMA x16 independent chains
__global__ void k_ffma(Span* out, float* sink) {
  float v[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = threadIdx.x + i;
  const float a = 0.999f, b = 0.001f;
  long long t0;
  span_begin(t0);
  for (int it = 0; it < ITER; ++it) {
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = fmaf(v[i], a, b);
  }
  span_end(out, t0);
  float s = 0;
  for (int i = 0; i < 16; ++i) s += v[i];
  if (s == 123.f) sink[0] = s;
}

// FFMA2 x16 independent chains (32 flops-pairs)
__global__ void k_ffma2(Span* out, float* sink) {
  float2 v[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = make_float2(threadIdx.x + i, threadIdx.x - i);
  const float2 a = make_float2(0.999f, 0.998f), b = make_float2(0.001f, 0.002f);
  long long t0;
  span_begin(t0);
  for (int it = 0; it < ITER; ++it) {
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = __ffma2_rn(v[i], a, b);
  }
  span_end(out, t0);
  float s = 0;
  for (int i = 0; i < 16; ++i) s += v[i].x + v[i].y;
  if (s == 123.f) sink[0] = s;
}

// LDS.128 with a lane -> 16-byte-chunk map given by pattern:
//   0: every lane its own chunk (32 unique, conflict free)
//   1: all lanes the same chunk (broadcast)
//   2: lane l and l+16 share (16 unique; duplicates in different quarter-warps)
//   3: lanes (2j, 2j+1) share (16 unique; duplicates adjacent)
//   4: 8 unique: chunk = lane % 8
//   5: 8 unique: chunk = lane / 4
//   6: 4 unique: chunk = lane / 8  (one address per quarter-warp)
//   7: 2 unique: chunk = lane / 16
//   8: 32 unique but 2-way bank conflict (stride 32 bytes... chunk = 2*lane)
template <int VEC>  // 4: LDS.128, 2: LDS.64, 1: LDS.32
__global__ void k_lds(Span* out, float* sink, int pattern) {
  extern __shared__ float4 sm[];
  for (int i = threadIdx.x; i < 2048; i += blockDim.x) sm[i] = make_float4(i, 1, 2, 3);
  const int lane = threadIdx.x & 31;
  int chunk;
  switch (pattern) {
    case 0: chunk = lane; break;
    case 1: chunk = 0; break;
    case 2: chunk = lane & 15; break;
    case 3: chunk = lane >> 1; break;
    case 4: chunk = lane & 7; break;
    case 5: chunk = lane >> 2; break;
    case 6: chunk = lane >> 3; break;
    case 7: chunk = lane >> 4; break;
    default: chunk = 2 * lane; break;
  }
  // byte offset of this lane's element; for VEC<4 the unit shrinks accordingly
  const char* base = reinterpret_cast<const char*>(sm) + chunk * (4 * VEC);
  float acc[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) acc[i] = 0.f;
  long long t0;
  span_begin(t0);
  for (int it = 0; it < ITER; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const char* p = base + i * 1024 + (it & 3) * 8192;
      if (VEC == 4) {
        float4 q;
        asm volatile("ld.volatile.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(q.x), "=f"(q.y), "=f"(q.z), "=f"(q.w)
                     : "r"((uint32_t)__cvta_generic_to_shared(p)));
        acc[i] += q.x + q.w;
      } else if (VEC == 2) {
        float2 q;
        asm volatile("ld.volatile.shared.v2.f32 {%0,%1}, [%2];" : "=f"(q.x), "=f"(q.y)
                     : "r"((uint32_t)__cvta_generic_to_shared(p)));
        acc[i] += q.x + q.y;
      } else {
        float q;
        asm volatile("ld.volatile.shared.f32 %0, [%1];" : "=f"(q) : "r"((uint32_t)__cvta_generic_to_shared(p)));
        acc[i] += q;
      }
    }
  }
  span_end(out, t0);
  float s = 0;
  for (int i = 0; i < 8; ++i) s += acc[i];
  if (s == 123.f) sink[0] = s;
}

// STS.128, every lane its own chunk
__global__ void k_sts(Span* out, float* sink) {
  extern __shared__ float4 sm[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float4* base = sm + (warp & 3) * 256 + lane;
  long long t0;
  span_begin(t0);
  for (int it = 0; it < ITER; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      asm volatile("st.volatile.shared.v4.f32 [%0], {%1,%2,%3,%4};" ::"r"((uint32_t)__cvta_generic_to_shared(base + i * 32)),
                   "f"((float)it), "f"((float)i), "f"((float)lane), "f"((float)warp)
                   : "memory");
    }
  }
  span_end(out, t0);
  if (sm[threadIdx.x].x == 123.25f) sink[0] = 1.f;
}

// L2 round trip: one thread chases its own tagged 8-byte slot (ld.relaxed.gpu after st.relaxed.gpu)
__global__ void k_l2_latency(long long* out, unsigned long long* slots) {
  if (threadIdx.x != 0) return;
  unsigned long long* s = slots + blockIdx.x * 16;
  long long t0 = clock64();
  unsigned long long v = 0;
  for (int it = 1; it <= 200; ++it) {
    asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(s), "l"((unsigned long long)it) : "memory");
    do {
      asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(s) : "memory");
    } while (v != (unsigned long long)it);
  }
  out[blockIdx.x] = (clock64() - t0) / 200;
}

// producer/consumer hand-off latency between two CTAs through L2 (ping-pong)
__global__ void k_pingpong(long long* out, unsigned long long* slots) {
  if (threadIdx.x != 0) return;
  const int pair = blockIdx.x >> 1, side = blockIdx.x & 1;
  unsigned long long* mine = slots + (pair * 2 + side) * 16;
  unsigned long long* other = slots + (pair * 2 + (side ^ 1)) * 16;
  long long t0 = clock64();
  unsigned long long v;
  for (int it = 1; it <= 200; ++it) {
    if (side == 0) {
      asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(other), "l"((unsigned long long)it) : "memory");
      do {
        asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(mine) : "memory");
      } while (v != (unsigned long long)it);
    } else {
      do {
        asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(mine) : "memory");
      } while (v != (unsigned long long)it);
      asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(other), "l"((unsigned long long)it) : "memory");
    }
  }
  out[blockIdx.x] = (clock64() - t0) / 200;  // one round trip = two hops
}

// LDS.128 (32 unique chunks) interleaved 1:1 with SHFL: do they share the shared-memory data pipe?
__global__ void k_lds_shfl(Span* out, float* sink) {
  extern __shared__ float4 sm[];
  for (int i = threadIdx.x; i < 2048; i += blockDim.x) sm[i] = make_float4(i, 1, 2, 3);
  const int lane = threadIdx.x & 31;
  const char* base = reinterpret_cast<const char*>(sm) + lane * 16;
  float acc[8], w[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    acc[i] = 0.f;
    w[i] = lane + i;
  }
  long long t0;
  span_begin(t0);
  for (int it = 0; it < ITER; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const char* p = base + i * 1024 + (it & 3) * 8192;
      float4 q;
      asm volatile("ld.volatile.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(q.x), "=f"(q.y), "=f"(q.z), "=f"(q.w)
                   : "r"((uint32_t)__cvta_generic_to_shared(p)));
      acc[i] += q.x + q.w;
      w[i] = __shfl_up_sync(0xffffffffu, w[i], 1);
    }
  }
  span_end(out, t0);
  float s = 0;
  for (int i = 0; i < 8; ++i) s += acc[i] + w[i];
  if (s == 123.f) sink[0] = s;
}

static double report(const char* name, Span* d_span, int nblk, int warps, int body) {
  std::vector<Span> h(nblk);
  cudaMemcpy(h.data(), d_span, nblk * sizeof(Span), cudaMemcpyDeviceToHost);
  double tot = 0;
  for (int i = 0; i < nblk; ++i) tot += double(h[i].t1 - h[i].t0);
  const double cyc = tot / nblk;
  const double per = cyc / (double(ITER) * body * warps);
  printf("%-44s warps/SM %2d  %8.3f SM-clk per warp-instr   (%6.2f lanes/clk/SM)\n", name, warps, per, 32.0 / per);
  return per;
}

int main() {
  int nsm = 0;
  CK(cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, 0));
  Span* d_span;
  float* d_sink;
  CK(cudaMalloc(&d_span, nsm * sizeof(Span)));
  CK(cudaMalloc(&d_sink, 1024));
  printf("SMs: %d\n", nsm);
  for (int warps : {16}) {
    const int nt = warps * 32;
    k_mufu<<<nsm, nt>>>(d_span, d_sink);
    CK(cudaDeviceSynchronize());
    report("MUFU.EX2", d_span, nsm, warps, 8);
    k_shfl<<<nsm, nt>>>(d_span, d_sink);
    CK(cudaDeviceSynchronize());
    report("SHFL.UP", d_span, nsm, warps, 8);
    k_mufu_shfl<<<nsm, nt>>>(d_span, d_sink);
    CK(cudaDeviceSynchronize());
    report("MUFU+SHFL pair (per pair)", d_span, nsm, warps, 8);
    k_ffma<<<nsm, nt>>>(d_span, d_sink);
    CK(cudaDeviceSynchronize());
    report("FFMA", d_span, nsm, warps, 16);
    k_ffma2<<<nsm, nt>>>(d_span, d_sink);
    CK(cudaDeviceSynchronize());
    report("FFMA2", d_span, nsm, warps, 16);
  }
  CK(cudaFuncSetAttribute(k_lds<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536));
  CK(cudaFuncSetAttribute(k_lds<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536));
  CK(cudaFuncSetAttribute(k_lds<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536));
  CK(cudaFuncSetAttribute(k_lds_shfl, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536));
  CK(cudaFuncSetAttribute(k_sts, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536));
  const char* pn[] = {"32 unique", "broadcast (1 unique)", "16 unique: l, l+16 share", "16 unique: 2j,2j+1 share",
                      "8 unique: lane%8", "8 unique: lane/4", "4 unique: lane/8", "2 unique: lane/16",
                      "32 unique, stride 2 chunks"};
  for (int warps : {8, 16}) {
    for (int p = 0; p < 9; ++p) {
      char nm[96];
      k_lds<4><<<nsm, warps * 32, 65536>>>(d_span, d_sink, p);
      CK(cudaDeviceSynchronize());
      snprintf(nm, sizeof nm, "LDS.128 %s", pn[p]);
      report(nm, d_span, nsm, warps, 8);
    }
    for (int p = 0; p < 9; ++p) {
      char nm[96];
      k_lds<2><<<nsm, warps * 32, 65536>>>(d_span, d_sink, p);
      CK(cudaDeviceSynchronize());
      snprintf(nm, sizeof nm, "LDS.64  %s", pn[p]);
      report(nm, d_span, nsm, warps, 8);
    }
    for (int p = 0; p < 2; ++p) {
      char nm[96];
      k_lds<1><<<nsm, warps * 32, 65536>>>(d_span, d_sink, p);
      CK(cudaDeviceSynchronize());
      snprintf(nm, sizeof nm, "LDS.32  %s", pn[p]);
      report(nm, d_span, nsm, warps, 8);
    }
    k_lds_shfl<<<nsm, warps * 32, 65536>>>(d_span, d_sink);
    CK(cudaDeviceSynchronize());
    report("LDS.128(32 unique)+SHFL pair (per pair)", d_span, nsm, warps, 8);
    k_sts<<<nsm, warps * 32, 65536>>>(d_span, d_sink);
    CK(cudaDeviceSynchronize());
    report("STS.128 32 unique", d_span, nsm, warps, 8);
  }
  // L2 latencies
  long long* d_lat;
  unsigned long long* d_slots;
  CK(cudaMalloc(&d_lat, 1024 * sizeof(long long)));
  CK(cudaMalloc(&d_slots, 1024 * 16 * sizeof(unsigned long long)));
  CK(cudaMemset(d_slots, 0, 1024 * 16 * sizeof(unsigned long long)));
  k_l2_latency<<<nsm, 32>>>(d_lat, d_slots);
  CK(cudaDeviceSynchronize());
  {
    std::vector<long long> h(nsm);
    cudaMemcpy(h.data(), d_lat, nsm * sizeof(long long), cudaMemcpyDeviceToHost);
    long long mn = 1 << 30, mx = 0, s = 0;
    for (auto v : h) {
      mn = v < mn ? v : mn;
      mx = v > mx ? v : mx;
      s += v;
    }
    printf("st.relaxed.gpu + ld.relaxed.gpu own slot: min %lld avg %lld max %lld clk\n", mn, s / nsm, mx);
  }
  CK(cudaMemset(d_slots, 0, 1024 * 16 * sizeof(unsigned long long)));
  k_pingpong<<<(nsm / 2) * 2, 32>>>(d_lat, d_slots);
  CK(cudaDeviceSynchronize());
  {
    const int n = (nsm / 2) * 2;
    std::vector<long long> h(n);
    cudaMemcpy(h.data(), d_lat, n * sizeof(long long), cudaMemcpyDeviceToHost);
    long long mn = 1 << 30, mx = 0, s = 0;
    for (auto v : h) {
      mn = v < mn ? v : mn;
      mx = v > mx ? v : mx;
      s += v;
    }
    printf("CTA<->CTA ping-pong round trip (2 hops) via L2: min %lld avg %lld max %lld clk\n", mn, s / n, mx);
  }
  return 0;
}
