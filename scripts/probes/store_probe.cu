// store_probe.cu -- how fast can a B200 absorb 1.0 GB of fp32 stores in the all-pairs volume pattern?
//   pattern 0: contiguous (every warp instruction writes 512 consecutive bytes)
//   pattern 1: volume tiles: a CTA writes 128 query rows x 8 patch rows x 128 B, query stride 30720 B,
//              patch-row stride 640 B; a warp instruction = 4 query rows x 128 B (the kernel's epilogue)
//   pattern 2: as 1 but a warp instruction = 32 query rows x 16 B (direct from registers)
//   pattern 3: as 1 but the CTA's tile is 128 query rows x 8 patch rows x 640 B (whole rows: 5 x-tiles)
//   pattern 4: transposed epilogue (lane = key): blocked layout, a CTA writes 256 query slices x two 8x8 tiles;
//              a warp instruction = ONE query x 128 B (32 lanes x 4 B, half a tile), 32 queries back to back
//   pattern 5: as 4 plus the pooled levels as that epilogue would write them (level 1: 4 queries x 2 x 16 B per
//              instruction; level 2: 16-byte pieces; level 3: 8-byte pieces, lane = query)
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

// transposed epilogue (lane = key): tile = (b, 256-query block, 8x32 key patch); CTA pair member r takes the 16
// columns x0 + 16 r of the 8 patch rows.  8 warps: quarter q = warp & 3, ch = warp >> 2 -> 128 queries.
//   LAYOUT 0: a quarter = half an 8x8 tile (4 rows x 8 columns): level 0 = one 128-byte run per query
//   LAYOUT 1: a quarter = 2 rows x 16 columns: level 0 = two 64-byte runs (two tiles) per query
// LEVELS bit 0: level 1 (LAYOUT 0: 4 queries x 2 x 16 B per instruction; LAYOUT 1: 4 queries x one 32-byte sector),
//        bit 1: level 2 as 16-byte pieces (lane = query), bit 2: level 3 as 8-byte pieces
template <int LAYOUT, int LEVELS>
__global__ void __launch_bounds__(256) probe_t(float *vol, float *l1, float *l2, float *l3, int ntiles) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int q = warp & 3, ch = warp >> 2;
  const int rank = blockIdx.x & 1, unit = blockIdx.x >> 1, nunits = gridDim.x >> 1;
  const size_t S0 = 48 * 160, S1 = 24 * 80, S2 = 12 * 40, S3 = 6 * 20;
  for (int t = unit; t < ntiles; t += nunits) {
    const int nt = t % 30, mt = (t / 30) % 30, b = t / 900;
    const int ty = nt / 5, tc0 = (nt % 5) * 4 + 2 * rank;
    const size_t qrow0 = (size_t)b * HW + mt * 256 + ch * 128;
    float *p0, *p1;
    int k;
    if (LAYOUT == 0) {
      const int tc = tc0 + (q >> 1), h = q & 1;
      p0 = vol + qrow0 * S0 + ((size_t)ty * 20 + tc) * 64 + h * 32 + lane;
      const int yy = lane >> 3, xx = lane & 7;
      k = (yy & 1) * 2 + (xx & 1);
      const int y1 = ty * 4 + 2 * h + (yy >> 1), x1 = 4 * tc + (xx >> 1);
      p1 = l1 + qrow0 * S1 + (size_t)((y1 >> 3) * 10 + (x1 >> 3)) * 64 + (y1 & 7) * 8 + (x1 & 7);
    } else {
      const int yy = lane >> 4, xx = lane & 15;
      p0 = vol + qrow0 * S0 + ((size_t)ty * 20 + tc0 + (xx >> 3)) * 64 + (2 * q + yy) * 8 + (xx & 7);
      k = yy * 2 + (xx & 1);
      const int y1 = ty * 4 + q, x1 = 4 * tc0 + (xx >> 1);
      p1 = l1 + qrow0 * S1 + (size_t)((y1 >> 3) * 10 + (x1 >> 3)) * 64 + (y1 & 7) * 8 + (x1 & 7);
    }
#pragma unroll 1
    for (int c4 = 0; c4 < 4; ++c4) {
#pragma unroll
      for (int i = 0; i < 32; ++i) __stcs(p0 + (size_t)(c4 * 32 + i) * S0, 1.f + i);
      if (LEVELS & 1) {
#pragma unroll
        for (int g = 0; g < 8; ++g) __stcs(p1 + (size_t)(c4 * 32 + 4 * g + k) * S1, 2.f + g);
      }
    }
    // levels 2 / 3 after the shared-memory regroup: thread = query, 128 queries per 4-warp group
    const size_t n = qrow0 + q * 32 + lane;
    if (LEVELS & 2) {
      __stcs(reinterpret_cast<float4 *>(l2 + n * S2 + (ty * 2) * 40 + (nt % 5) * 8 + 4 * rank), make_float4(1, 2, 3, 4));
      __stcs(reinterpret_cast<float4 *>(l2 + n * S2 + (ty * 2 + 1) * 40 + (nt % 5) * 8 + 4 * rank), make_float4(1, 2, 3, 4));
    }
    if (LEVELS & 4)
      __stcs(reinterpret_cast<float2 *>(l3 + n * S3 + ty * 20 + (nt % 5) * 4 + 2 * rank), make_float2(1, 2));
  }
}

template <int LAYOUT, int LEVELS>
void run_t(const char *name, float *vol, float *l1, float *l2, float *l3) {
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  const int ntiles = B * 30 * 30;
  for (int i = 0; i < 2; ++i) probe_t<LAYOUT, LEVELS><<<148, 256>>>(vol, l1, l2, l3, ntiles);
  cudaEventRecord(e0);
  const int reps = 10;
  for (int i = 0; i < reps; ++i) probe_t<LAYOUT, LEVELS><<<148, 256>>>(vol, l1, l2, l3, ntiles);
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms;
  cudaEventElapsedTime(&ms, e0, e1);
  const double bytes = (double)B * HW * HW * 4 * (1.0 + (LEVELS & 1 ? 0.25 : 0) + (LEVELS & 2 ? 0.0625 : 0) + (LEVELS & 4 ? 0.015625 : 0));
  printf("%-72s %7.1f us  %7.1f GB/s\n", name, ms / reps * 1e3, bytes / (ms / reps * 1e-3) / 1e9);
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
  float *l1, *l2, *l3;
  cudaMalloc(&l1, (size_t)B * HW * HW);
  cudaMalloc(&l2, (size_t)B * HW * HW / 4);
  cudaMalloc(&l3, (size_t)B * HW * HW / 16);
  run_t<0, 0>("transposed, half-tile quarters, level 0 (128 B per inst)", vol, l1, l2, l3);
  run_t<0, 1>("transposed, half-tile quarters, levels 0-1 (2 x 16 B pieces)", vol, l1, l2, l3);
  run_t<1, 0>("transposed, 2x16 quarters, level 0 (2 x 64 B per inst)", vol, l1, l2, l3);
  run_t<1, 1>("transposed, 2x16 quarters, levels 0-1 (32 B sectors)", vol, l1, l2, l3);
  run_t<1, 3>("transposed, 2x16 quarters, levels 0-2 (16 B pieces)", vol, l1, l2, l3);
  run_t<1, 7>("transposed, 2x16 quarters, levels 0-3 (8 B pieces)", vol, l1, l2, l3);
  run_t<1, 6>("transposed, 2x16 quarters, levels 0,2,3", vol, l1, l2, l3);
  cudaError_t e = cudaDeviceSynchronize();
  printf("status: %s\n", cudaGetErrorString(e));
  return 0;
}
