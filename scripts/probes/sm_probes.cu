// microbench.cu -- small probes of the sm_100a SM used to size the sampler kernels:
//   * shared-memory LDS.128 throughput under different lane->address sharing patterns
//     (how many wavefronts a multicast 128-bit load costs);
//   * FP32 FFMA rate of the register-blocked Toeplitz update the sampler forward issues
//     (acc[t][k] += a[t] * v[t + 2k], 8 x 21 accumulators per thread).
// Diagnostics only, built into their own library (scripts/probe_sm.py: libb200probes.so = this file +
// csrc/common.cu); the product library libb200corr.so does not contain them.
#include "common.cuh"

namespace {

__device__ __forceinline__ int lds_pattern_chunk(int pattern, int lane) {
  const int li = lane >> 3, lj = lane & 7;
  switch (pattern) {
    case 0: return lane;                    // 32 distinct 16-B chunks (512 B)
    case 1: return lane >> 2;               // 8 distinct, 4 consecutive lanes share
    case 2: return lane & 7;                // 8 distinct, every quarter-warp reads all 8 (bwd v-load)
    case 3: return (lane >> 3) * 9;         // 4 distinct, one per quarter-warp (bwd G-load)
    case 4: return 0;                       // 1 distinct (full broadcast)
    case 5: return (lane & 3) + 4 * (lane >> 4);  // 4 distinct per half-warp, halves differ
    case 6: return (li + lj) * 21;          // fwd in2 load as first written: 11 distinct rows
    case 7: {                               // fwd in2 load remapped: 7 distinct rows per half-warp
      const int i = (lane >> 2) & 3, j = (lane & 3) + 4 * (lane >> 4);
      return (i + j) * 21;
    }
    case 8: return lane & 15;               // 16 distinct, halves identical
    case 9: return (lane & 3) * 21;         // 4 distinct rows, every quarter identical
    case 10: return (lane >> 1);            // 16 distinct, pairs share
    default: return lane;
  }
}

__global__ void __launch_bounds__(512) lds_probe_kernel(float *sink, int iters, int pattern,
                                                        unsigned long long *cycles) {
  extern __shared__ float4 sm4[];
  for (int i = threadIdx.x; i < 4096; i += blockDim.x)
    sm4[i] = make_float4((float)i, 1.f, 2.f, 3.f);
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  const uint32_t base =
      b200dev::smem_u32(sm4) + 16u * (uint32_t)(lds_pattern_chunk(pattern, lane) + (warp & 7) * 256);
  uint32_t a0 = 0;
  unsigned long long t0 = clock64();
#pragma unroll 1
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int u = 0; u < 16; ++u) {
      uint32_t x, y, z, w;
      // distinct immediate offset per unrolled load + memory clobber: nothing can be folded
      asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];"
                   : "=r"(x), "=r"(y), "=r"(z), "=r"(w)
                   : "r"(base + (uint32_t)u * 2048u)
                   : "memory");
      a0 ^= x ^ y ^ z ^ w;
    }
  }
  unsigned long long t1 = clock64();
  if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
  if (a0 == 0x12345u) sink[0] = (float)a0;
}

// packed FFMA2 peak: 16 independent float2 chains per thread, operands in registers
__global__ void __launch_bounds__(256) ffma2_peak_kernel(float *sink, int iters, float seed) {
  float2 acc[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) acc[i] = make_float2(seed + threadIdx.x + i, seed - i);
  const float2 x = make_float2(1.0000001f + seed, 0.9999999f - seed);
  const float2 y = make_float2(0.5f * seed, 1.f - seed);
#pragma unroll 1
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int u = 0; u < 8; ++u)
#pragma unroll
      for (int i = 0; i < 16; ++i) acc[i] = __ffma2_rn(acc[i], x, y);
  }
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 16; ++i) s += acc[i].x + acc[i].y;
  if (s == 12345.678f) sink[0] = s;
}

// register-blocked update of the sampler forward with FFMA2, operands refreshed from shared memory
// with the real access pattern replaced by a conflict-free broadcast (isolates the FMA pipe)
__global__ void __launch_bounds__(256, 1) ffma_toeplitz_kernel(float *sink, int iters, float seed) {
  __shared__ float4 sm[64];
  if (threadIdx.x < 64) sm[threadIdx.x] = make_float4(seed + threadIdx.x, 1.f, seed, 2.f);
  __syncthreads();
  float2 acc[4][21];
#pragma unroll
  for (int t = 0; t < 4; ++t)
#pragma unroll
    for (int k = 0; k < 21; ++k) acc[t][k] = make_float2(0.f, 0.f);
#pragma unroll 1
  for (int it = 0; it < iters; ++it) {
    const float4 a0 = sm[(it & 3)], a1 = sm[(it & 3) + 4];
    const float2 ap[4] = {make_float2(a0.x, a0.y), make_float2(a0.z, a0.w), make_float2(a1.x, a1.y),
                          make_float2(a1.z, a1.w)};
#pragma unroll
    for (int sg = 0; sg < 12; ++sg) {
      const float4 v4 = sm[8 + sg + (it & 3) * 12];
      const float2 vp[2] = {make_float2(v4.x, v4.y), make_float2(v4.z, v4.w)};
#pragma unroll
      for (int uu = 0; uu < 2; ++uu)
#pragma unroll
        for (int tp = 0; tp < 4; ++tp) {
          const int d = 2 * sg + uu - tp;
          if (d >= 0 && d < 21) acc[tp][d] = __ffma2_rn(ap[tp], vp[uu], acc[tp][d]);
        }
    }
  }
  float s = 0.f;
#pragma unroll
  for (int t = 0; t < 4; ++t)
#pragma unroll
    for (int k = 0; k < 21; ++k) s += acc[t][k].x + acc[t][k].y;
  if (s == 12345.678f) sink[0] = s;
}

}  // namespace

// ---- window-gather ceiling: every thread owns a contiguous slice of `slice_sectors` 32-byte sectors
// (one RAFT query's H_l x W_l volume slice) and reads a window of `rows` rows x 2 sectors at a random
// position inside it, row pitch `pitch_sectors` -- the access pattern of the correlation lookup with
// nothing else attached (all loads in flight at once, 2048 threads per SM).
template <int ROWS>
__global__ void __launch_bounds__(256)
gather_probe_kernel(const float *__restrict__ buf, size_t nslices, int slice_sectors, int pitch_sectors,
                    float *sink) {
  const size_t tid = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  uint32_t h = (uint32_t)tid;
  h ^= h >> 16; h *= 0x7feb352du; h ^= h >> 15; h *= 0x846ca68bu; h ^= h >> 16;
  const int rows_avail = slice_sectors / pitch_sectors - ROWS;
  const size_t base = (tid % nslices) * slice_sectors + (h % (uint32_t)rows_avail) * pitch_sectors +
                      ((h >> 8) % (uint32_t)(pitch_sectors - 1));
  float v[ROWS][2][8];
#pragma unroll
  for (int r = 0; r < ROWS; ++r)
#pragma unroll
    for (int k = 0; k < 2; ++k) {
      const float *p = buf + (base + (size_t)r * pitch_sectors + k) * 8;
      asm volatile("ld.global.nc.L1::no_allocate.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                   : "=f"(v[r][k][0]), "=f"(v[r][k][1]), "=f"(v[r][k][2]), "=f"(v[r][k][3]), "=f"(v[r][k][4]),
                     "=f"(v[r][k][5]), "=f"(v[r][k][6]), "=f"(v[r][k][7])
                   : "l"(p));
    }
  float acc = 0.f;
#pragma unroll
  for (int r = 0; r < ROWS; ++r)
#pragma unroll
    for (int k = 0; k < 2; ++k)
#pragma unroll
      for (int j = 0; j < 8; ++j) acc += v[r][k][j];
  if (acc == 123.456f) sink[0] = acc;
}

extern "C" {

// cycles per LDS.128 warp-instruction per SM at `warps` resident warps (one CTA per SM).
int b200corr_probe_lds(int pattern, int warps, int iters, float *cycles_per_lds, void *stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  B200_CHECK(warps >= 1 && warps <= 16 && iters > 0 && cycles_per_lds, "probe_lds: bad arguments");
  const int blocks = b200::num_sms();
  float *sink = nullptr;
  unsigned long long *cyc = nullptr;
  B200_CUDA(cudaMalloc(&sink, sizeof(float)));
  B200_CUDA(cudaMalloc(&cyc, sizeof(unsigned long long) * blocks));
  const size_t smem = 4096 * sizeof(float4);
  B200_CUDA(cudaFuncSetAttribute(lds_probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  lds_probe_kernel<<<blocks, warps * 32, smem, stream>>>(sink, 4, pattern, cyc);
  lds_probe_kernel<<<blocks, warps * 32, smem, stream>>>(sink, iters, pattern, cyc);
  B200_LAUNCH_OK("lds_probe_kernel");
  B200_CUDA(cudaStreamSynchronize(stream));
  unsigned long long h[256];
  B200_CUDA(cudaMemcpy(h, cyc, sizeof(unsigned long long) * (blocks < 256 ? blocks : 256),
                       cudaMemcpyDeviceToHost));
  double sum = 0;
  const int n = blocks < 256 ? blocks : 256;
  for (int i = 0; i < n; ++i) sum += (double)h[i];
  *cycles_per_lds = (float)(sum / n / ((double)iters * 16 * warps));
  cudaFree(sink);
  cudaFree(cyc);
  return 0;
}

// achieved TFLOP/s of back-to-back packed FFMA2 (2 FMAs per lane per instruction)
int b200corr_probe_ffma2_peak(int iters, float *tflops, void *stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  B200_CHECK(iters > 0 && tflops, "probe_ffma2_peak: bad arguments");
  float *sink = nullptr;
  B200_CUDA(cudaMalloc(&sink, sizeof(float)));
  cudaEvent_t e0, e1;
  B200_CUDA(cudaEventCreate(&e0));
  B200_CUDA(cudaEventCreate(&e1));
  const int blocks = b200::num_sms() * 8;
  ffma2_peak_kernel<<<blocks, 256, 0, stream>>>(sink, 8, 0.f);
  float best = 1e30f;
  for (int rep = 0; rep < 5; ++rep) {
    B200_CUDA(cudaEventRecord(e0, stream));
    ffma2_peak_kernel<<<blocks, 256, 0, stream>>>(sink, iters, 0.f);
    B200_CUDA(cudaEventRecord(e1, stream));
    B200_CUDA(cudaEventSynchronize(e1));
    float ms = 0.f;
    B200_CUDA(cudaEventElapsedTime(&ms, e0, e1));
    if (ms < best) best = ms;
  }
  B200_LAUNCH_OK("ffma2_peak_kernel");
  *tflops = (float)(2.0 * 2 * 16 * 8 * (double)iters * blocks * 256 / (best * 1e-3) / 1e12);
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  cudaFree(sink);
  return 0;
}

// achieved TFLOP/s of the 8x21 Toeplitz FFMA block (256 threads/SM, 168 accumulators each)
int b200corr_probe_ffma_toeplitz(int iters, float *tflops, void *stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  B200_CHECK(iters > 0 && tflops, "probe_ffma_toeplitz: bad arguments");
  float *sink = nullptr;
  B200_CUDA(cudaMalloc(&sink, sizeof(float)));
  cudaEvent_t e0, e1;
  B200_CUDA(cudaEventCreate(&e0));
  B200_CUDA(cudaEventCreate(&e1));
  const int blocks = b200::num_sms() * 4;
  ffma_toeplitz_kernel<<<blocks, 256, 0, stream>>>(sink, 8, 0.f);
  float best = 1e30f;
  for (int rep = 0; rep < 5; ++rep) {
    B200_CUDA(cudaEventRecord(e0, stream));
    ffma_toeplitz_kernel<<<blocks, 256, 0, stream>>>(sink, iters, 0.f);
    B200_CUDA(cudaEventRecord(e1, stream));
    B200_CUDA(cudaEventSynchronize(e1));
    float ms = 0.f;
    B200_CUDA(cudaEventElapsedTime(&ms, e0, e1));
    if (ms < best) best = ms;
  }
  B200_LAUNCH_OK("ffma_toeplitz_kernel");
  *tflops = (float)(2.0 * 168 * (double)iters * blocks * 256 / (best * 1e-3) / 1e12);
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  cudaFree(sink);
  return 0;
}

// window rows gathered per second (in units of 1e9) when nothing but the gather runs: `nslices`
// slices of slice_bytes each (buffer provided by the caller, >= nslices * slice_bytes, 32-B aligned),
// one 10-row x 64-byte window per slice and launch, row pitch pitch_bytes
int b200corr_measure_gather_peak(const float *buf, long long nslices, int slice_bytes, int pitch_bytes,
                                 float *grows_per_s, void *stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  B200_CHECK(buf && grows_per_s && nslices >= 256 && slice_bytes % 32 == 0 && pitch_bytes % 32 == 0 &&
                 pitch_bytes >= 64 && slice_bytes / pitch_bytes > 10 && ((uintptr_t)buf & 31) == 0,
             "measure_gather_peak: bad arguments");
  float *sink = nullptr;
  B200_CUDA(cudaMalloc(&sink, sizeof(float)));
  cudaEvent_t e0, e1;
  B200_CUDA(cudaEventCreate(&e0));
  B200_CUDA(cudaEventCreate(&e1));
  // 8 windows per slice and launch so that a launch lasts long enough to time
  const long long threads = nslices * 8;
  const int blocks = (int)((threads + 255) / 256);
  gather_probe_kernel<10><<<blocks, 256, 0, stream>>>(buf, (size_t)nslices, slice_bytes / 32, pitch_bytes / 32, sink);
  float best = 1e30f;
  for (int rep = 0; rep < 3; ++rep) {
    B200_CUDA(cudaEventRecord(e0, stream));
    gather_probe_kernel<10><<<blocks, 256, 0, stream>>>(buf, (size_t)nslices, slice_bytes / 32, pitch_bytes / 32, sink);
    B200_CUDA(cudaEventRecord(e1, stream));
    B200_CUDA(cudaEventSynchronize(e1));
    float ms = 0.f;
    B200_CUDA(cudaEventElapsedTime(&ms, e0, e1));
    if (ms < best) best = ms;
  }
  B200_LAUNCH_OK("gather_probe_kernel");
  *grows_per_s = (float)((double)blocks * 256 * 10 / (best * 1e-3) / 1e9);
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  cudaFree(sink);
  return 0;
}

}  // extern "C"
