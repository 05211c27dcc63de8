// gather_probe2.cu -- does the window-gather ceiling of the RAFT lookup move with the REQUEST SHAPE?
// Round 1 found ~38e9 window rows/s for "one thread = one query, one 32-byte load per window row" (gather_probe.cu)
// and the lookup kernel sits on it whatever its occupancy, ILP or L2 prefetching (round 2).  This probe reads the
// same bytes -- for each of Q query slices (blocked layout: 120 tiles of 256 B), a window of 2 x 2 tiles, 5 rows of
// each tile (rows 3..7 of the upper, 0..4 of the lower tiles: 20 sectors of 32 B) -- in three ways:
//   A  lane = query, 20 x LDG.256 per thread (the lookup kernel's pattern)
//   B  8 lanes = one tile, lane r loads row r if it is needed: the 5 rows of a tile leave the SM as ONE 256-byte-
//      line request pair instead of 5 sector requests
//   C  lane = query, cp.async.bulk of the 160 contiguous bytes of a tile's rows into shared memory (4 per query)
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o gather_probe2.bin gather_probe2.cu
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

constexpr int SLICE_FLOATS = 7680;   // 48 x 160 level-0 slice = 6 x 20 tiles of 64 floats
constexpr int TILES_W = 20;

__device__ __forceinline__ void ldg256(const float *p, float (&v)[8]) {
  asm volatile("ld.global.nc.L1::no_allocate.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=f"(v[0]), "=f"(v[1]), "=f"(v[2]), "=f"(v[3]), "=f"(v[4]), "=f"(v[5]), "=f"(v[6]), "=f"(v[7])
               : "l"(p));
}
__device__ __forceinline__ uint32_t hash32(uint32_t x) {
  x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16;
  return x;
}
// upper-left tile of query q's window
__device__ __forceinline__ int window_tile(uint32_t q) {
  const uint32_t h = hash32(q);
  return (h % 5) * TILES_W + ((h >> 8) % 19);
}

__global__ void __launch_bounds__(128) probe_a(const float *buf, int Q, float *sink) {
  const int q = blockIdx.x * 128 + threadIdx.x;
  if (q >= Q) return;
  const float *s = buf + (size_t)q * SLICE_FLOATS + window_tile(q) * 64;
  float v[20][8];
#pragma unroll
  for (int t = 0; t < 4; ++t) {
    const float *tp = s + ((t >> 1) * TILES_W + (t & 1)) * 64 + ((t >> 1) ? 0 : 3 * 8);
#pragma unroll
    for (int r = 0; r < 5; ++r) ldg256(tp + r * 8, v[t * 5 + r]);
  }
  float acc = 0.f;
#pragma unroll
  for (int k = 0; k < 20; ++k)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc += v[k][j];
  if (acc == 123.456f) sink[0] = acc;
}

// 8 lanes per tile: a warp instruction covers 4 tiles = one query's window; a warp walks 32 queries
__global__ void __launch_bounds__(128) probe_b(const float *buf, int Q, float *sink) {
  const int warp = (blockIdx.x * 128 + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  const int t = lane >> 3, r = lane & 7;
  const bool need = (t >> 1) ? r <= 4 : r >= 3;
  float acc = 0.f;
  float v[8][8];
#pragma unroll 1
  for (int g = 0; g < 32; g += 8) {
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const int q = warp * 32 + g + k;
#pragma unroll
      for (int j = 0; j < 8; ++j) v[k][j] = 0.f;
      if (q < Q && need) {
        const float *tp = buf + (size_t)q * SLICE_FLOATS + (window_tile(q) + (t >> 1) * TILES_W + (t & 1)) * 64 + r * 8;
        ldg256(tp, v[k]);
      }
    }
#pragma unroll
    for (int k = 0; k < 8; ++k)
#pragma unroll
      for (int j = 0; j < 8; ++j) acc += v[k][j];
  }
  if (acc == 123.456f) sink[0] = acc;
}

// bulk copies: lane = query, 4 x 160 B into this thread's shared-memory slot, one mbarrier per CTA
__global__ void __launch_bounds__(64) probe_c(const float *buf, int Q, float *sink) {
  __shared__ __align__(128) float slot[64][4][40];
  __shared__ uint64_t bar;
  const int q = blockIdx.x * 64 + threadIdx.x;
  const uint32_t bar_a = (uint32_t)__cvta_generic_to_shared(&bar);
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar_a), "r"(64));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (q < Q) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar_a), "r"(640) : "memory");
    const float *s = buf + (size_t)q * SLICE_FLOATS + window_tile(q) * 64;
#pragma unroll
    for (int t = 0; t < 4; ++t) {
      const float *tp = s + ((t >> 1) * TILES_W + (t & 1)) * 64 + ((t >> 1) ? 0 : 3 * 8);
      asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                       (uint32_t)__cvta_generic_to_shared(&slot[threadIdx.x][t][0])),
                   "l"(tp), "r"(160), "r"(bar_a)
                   : "memory");
    }
  } else {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar_a) : "memory");
  }
  uint32_t ok = 0;
  while (!ok)
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok) : "r"(bar_a), "r"(0) : "memory");
  float acc = 0.f;
#pragma unroll
  for (int t = 0; t < 4; ++t)
    for (int j = 0; j < 40; j += 8) acc += slot[threadIdx.x][t][j];
  if (acc == 123.456f) sink[0] = acc;
}

template <class F>
void run(const char *name, F launch, int Q) {
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  for (int i = 0; i < 3; ++i) launch();
  cudaEventRecord(e0);
  const int reps = 20;
  for (int i = 0; i < reps; ++i) launch();
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms;
  cudaEventElapsedTime(&ms, e0, e1);
  const double us = ms / reps * 1e3;
  printf("%-58s %8.1f us  %6.1f G window rows/s  %7.1f GB/s of the 640 requested bytes per window\n", name, us,
         Q * 10.0 / us / 1e3, Q * 640.0 / us / 1e3);
}

int main() {
  const int Q = 30720 * 4;            // one lookup's windows (4 levels' worth of level-0-like slices): 3.8 GB of slices
  float *buf, *sink;
  const size_t bytes = (size_t)Q * SLICE_FLOATS * 4;
  if (cudaMalloc(&buf, bytes) != cudaSuccess) { printf("alloc failed\n"); return 1; }
  cudaMalloc(&sink, 4);
  cudaMemset(buf, 0, bytes);
  run("A lane = query, 20 x LDG.256 (lookup kernel's pattern)", [&] { probe_a<<<(Q + 127) / 128, 128>>>(buf, Q, sink); }, Q);
  run("B 8 lanes = one tile, rows coalesced into line requests", [&] { probe_b<<<(Q / 32 * 32 + 127) / 128, 128>>>(buf, Q, sink); }, Q);
  run("C lane = query, 4 x cp.async.bulk of 160 B into smem", [&] { probe_c<<<(Q + 63) / 64, 64>>>(buf, Q, sink); }, Q);
  cudaError_t e = cudaDeviceSynchronize();
  printf("status: %s\n", cudaGetErrorString(e));
  return 0;
}
