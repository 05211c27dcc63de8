// store_probe.cu -- how fast can a B200 absorb 1.0 GB of fp32 stores in the all-pairs volume pattern?
//   pattern 0: contiguous (every warp instruction writes 512 consecutive bytes)
//   pattern 1: volume tiles: a CTA writes 128 query rows x 8 patch rows x 128 B, query stride 30720 B,
//              patch-row stride 640 B; a warp instruction = 4 query rows x 128 B (the kernel's epilogue)
//   pattern 2: as 1 but a warp instruction = 32 query rows x 16 B (direct from registers)
//   pattern 3: as 1 but the CTA's tile is 128 query rows x 8 patch rows x 640 B (whole rows: 5 x-tiles)
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o store_probe.bin store_probe.cu
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

constexpr int HW = 7680, H = 48, W = 160, B = 4;

template <int PATTERN>
__global__ void __launch_bounds__(256) probe(float *vol, int ntiles) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const float4 v = make_float4(1.f, 2.f, 3.f, 4.f);
  if (PATTERN == 0) {
    const size_t total4 = (size_t)B * HW * HW / 4;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total4; i += (size_t)gridDim.x * blockDim.x)
      __stcs(reinterpret_cast<float4 *>(vol) + i, v);
    return;
  }
  // tiles: (b, mt, ny, nx) with nx fastest, like the kernel
  for (int t = blockIdx.x; t < ntiles; t += gridDim.x) {
    const int nt = t % 30, mt = (t / 30) % 60, b = t / 1800;
    const int y0 = (nt / 5) * 8, x0 = (nt % 5) * 32;
    const int m0 = mt * 128 + warp * 16;   // 8 warps x 16 query rows
    if (PATTERN == 1) {
      for (int r = 0; r < 8; ++r)
#pragma unroll
        for (int it = 0; it < 4; ++it) {
          const int row = m0 + it * 4 + (lane >> 3);
          __stcs(reinterpret_cast<float4 *>(vol + (((size_t)b * HW + row) * H + y0 + r) * W + x0 + 4 * (lane & 7)), v);
        }
    } else if (PATTERN == 2) {
      // lane = query row (16 rows per warp -> two half-warps write two patch rows)
      for (int r = 0; r < 8; r += 2)
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const int row = m0 + (lane & 15);
          __stcs(reinterpret_cast<float4 *>(vol + (((size_t)b * HW + row) * H + y0 + r + (lane >> 4)) * W + x0 + 4 * j), v);
        }
    } else if (PATTERN == 3) {
      // whole rows: tile index reinterpretation -- 5 x-tiles handled by the same CTA back to back
      const int band = t % 6, mt3 = (t / 6) % 60, b3 = t / 360;
      if (t >= ntiles / 5) return;
      const int mm0 = mt3 * 128 + warp * 16;
      for (int r = 0; r < 8; ++r)
        for (int it = 0; it < 20; ++it) {
          const int idx = it * 32 + lane;            // 16 rows x 40 float4
          const int row = mm0 + idx / 40, c4 = idx % 40;
          __stcs(reinterpret_cast<float4 *>(vol + (((size_t)b3 * HW + row) * H + band * 8 + r) * W + 4 * c4), v);
        }
    }
  }
}

template <int PATTERN>
void run(const char *name, float *vol, int ctas_per_sm = 4) {
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  const int ntiles = B * 60 * 30;
  for (int i = 0; i < 2; ++i) probe<PATTERN><<<148 * ctas_per_sm, 256>>>(vol, ntiles);
  cudaEventRecord(e0);
  const int reps = 10;
  for (int i = 0; i < reps; ++i) probe<PATTERN><<<148 * ctas_per_sm, 256>>>(vol, ntiles);
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms;
  cudaEventElapsedTime(&ms, e0, e1);
  const double bytes = (double)B * HW * HW * 4;
  printf("%-60s %7.1f us  %7.1f GB/s\n", name, ms / reps * 1e3, bytes / (ms / reps * 1e-3) / 1e9);
}

int main() {
  float *vol;
  cudaMalloc(&vol, (size_t)B * HW * HW * 4);
  run<0>("contiguous", vol);
  run<1>("volume tiles, 4 rows x 128 B per instruction", vol);
  run<2>("volume tiles, 16 B per lane, lane = query row", vol);
  run<3>("row bands (5 x-tiles per CTA), coalesced", vol);
  run<1>("volume tiles, 8 warps per SM", vol, 1);
  run<1>("volume tiles, 16 warps per SM", vol, 2);
  run<0>("contiguous, 8 warps per SM", vol, 1);
  cudaError_t e = cudaDeviceSynchronize();
  printf("status: %s\n", cudaGetErrorString(e));
  return 0;
}
