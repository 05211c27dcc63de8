// lds_width_probe.cu -- shared-memory cost (cycles per warp-wide load per SM) of LDS.32 / LDS.64 / LDS.128 under the
// lane -> address sharing patterns of the sampler backward's grad_output loads.  Round 1 measured for LDS.128:
// max(2, slots / 8) wavefronts, only an aligned lane pair reading the same 16 bytes shares a slot.  Question of
// round 2: does a narrower load make the "two addresses per warp" pattern (lane parity = pixel row) cheaper per byte?
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o lds_width_probe.bin lds_width_probe.cu
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

__device__ __forceinline__ int pattern_slot(int pattern, int lane) {
  switch (pattern) {
    case 0: return 0;              // full broadcast
    case 1: return lane & 1;       // two addresses, alternating lanes (pixel row = lane parity)
    case 2: return lane >> 4;      // two addresses, one per half warp
    case 3: return lane >> 1;      // 16 addresses, aligned pairs share
    case 4: return lane & 15;      // 16 addresses, lanes l and l+16 share
    default: return lane;          // 32 distinct
  }
}

template <int BYTES>
__global__ void __launch_bounds__(512) probe(float *sink, int iters, int pattern, unsigned long long *cycles) {
  extern __shared__ float4 sm4[];
  for (int i = threadIdx.x; i < 4096; i += blockDim.x) sm4[i] = make_float4((float)i, 1.f, 2.f, 3.f);
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  // slots are BYTES apart within a 1 KB region per warp: conflict-free when distinct
  const uint32_t base = (uint32_t)__cvta_generic_to_shared(sm4) + (uint32_t)(pattern_slot(pattern, lane) * BYTES) + (warp & 7) * 4096u;
  uint32_t a0 = 0;
  unsigned long long t0 = clock64();
#pragma unroll 1
  for (int it = 0; it < iters; ++it) {
    const uint32_t b2 = base + (uint32_t)(it & 1) * 512u;   // iteration-dependent: the loads stay inside the loop
#pragma unroll
    for (int u = 0; u < 16; ++u) {
      uint32_t x = 0, y = 0, z = 0, w = 0;
      if (BYTES == 16)
        asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(x), "=r"(y), "=r"(z), "=r"(w) : "r"(b2 + u * 1024u) : "memory");
      else if (BYTES == 8)
        asm volatile("ld.shared.v2.u32 {%0,%1}, [%2];" : "=r"(x), "=r"(y) : "r"(b2 + u * 1024u) : "memory");
      else
        asm volatile("ld.shared.u32 %0, [%1];" : "=r"(x) : "r"(b2 + u * 1024u) : "memory");
      a0 ^= x ^ y ^ z ^ w;
    }
  }
  unsigned long long t1 = clock64();
  if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
  if (a0 == 0x12345u) sink[0] = (float)a0;
}

template <int BYTES>
double run(int pattern, float *sink, unsigned long long *cyc) {
  const int iters = 2000, warps = 16, blocks = 148;
  cudaFuncSetAttribute(probe<BYTES>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536);
  probe<BYTES><<<blocks, warps * 32, 65536>>>(sink, iters, pattern, cyc);
  probe<BYTES><<<blocks, warps * 32, 65536>>>(sink, iters, pattern, cyc);
  cudaDeviceSynchronize();
  unsigned long long h[148];
  cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
  double m = 0;
  for (int i = 0; i < blocks; ++i) m += (double)h[i];
  m /= blocks;
  return m / ((double)iters * 16 * warps);     // cycles per warp-wide load, SM-wide
}

int main() {
  float *sink;
  unsigned long long *cyc;
  cudaMalloc(&sink, 4);
  cudaMalloc(&cyc, 148 * 8);
  const char *names[6] = {"full broadcast", "2 addr, lane parity", "2 addr, half warps", "16 addr, aligned pairs share",
                          "16 addr, l and l+16 share", "32 distinct"};
  printf("%-32s %10s %10s %10s   (cycles per warp-wide load per SM; bytes delivered per lane 4 / 8 / 16)\n", "pattern", "LDS.32", "LDS.64", "LDS.128");
  for (int p = 0; p < 6; ++p)
    printf("%-32s %10.2f %10.2f %10.2f\n", names[p], run<4>(p, sink, cyc), run<8>(p, sink, cyc), run<16>(p, sink, cyc));
  printf("status: %s\n", cudaGetErrorString(cudaDeviceSynchronize()));
  return 0;
}
