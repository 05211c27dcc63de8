// gather_probe.cu -- how fast can a B200 read scattered 32/64-byte pieces out of a 1 GB buffer?
// Patterns: 0 = linear stream (reference), 1 = random 32-B sectors, 2 = random 64-B pairs,
// 3 = lookup-like: random window of 10 rows x 64 B at a 640-B row pitch (one window per thread).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o gather_probe.bin gather_probe.cu
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

__device__ __forceinline__ void ldg256(const float *p, float (&v)[8]) {
  asm volatile("ld.global.nc.L1::no_allocate.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=f"(v[0]), "=f"(v[1]), "=f"(v[2]), "=f"(v[3]), "=f"(v[4]), "=f"(v[5]), "=f"(v[6]), "=f"(v[7])
               : "l"(p));
}
__device__ __forceinline__ uint32_t hash32(uint32_t x) {
  x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16;
  return x;
}

template <int PATTERN, int PER_THREAD>
__global__ void __launch_bounds__(256) probe(const float *buf, size_t nsect, float *sink) {
  const size_t tid = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  const size_t nthreads = gridDim.x * (size_t)blockDim.x;
  float acc = 0.f;
  float v[PER_THREAD][8];
#pragma unroll
  for (int k = 0; k < PER_THREAD; ++k) {
    size_t s;
    if (PATTERN == 0) s = (tid + k * nthreads) % nsect;
    else if (PATTERN == 1) s = hash32((uint32_t)(tid * PER_THREAD + k)) % nsect;
    else if (PATTERN == 2) s = ((hash32((uint32_t)((tid * PER_THREAD + k) >> 1)) % (nsect / 2)) * 2) + (k & 1);
    else if (PATTERN == 4) {  // lookup-like: thread = query with its own contiguous 30720-B slice (960 sectors),
      // window of PER_THREAD/2 rows x 64 B at a random position inside the slice, 640-B row pitch
      const size_t slice = (tid % (nsect / 960)) * 960;
      const uint32_t h = hash32((uint32_t)tid);
      s = slice + (h % 38) * 20 + ((h >> 8) % 18) + (k >> 1) * 20 + (k & 1);
    } else if (PATTERN == 5) {  // as 4 but one sector per row
      const size_t slice = (tid % (nsect / 960)) * 960;
      const uint32_t h = hash32((uint32_t)tid);
      s = slice + (h % 38) * 20 + ((h >> 8) % 19) + k * 20;
    } else {  // window: base sector random, row k/2 at 640 B = 20 sectors pitch, 2 sectors per row
      const size_t base = hash32((uint32_t)tid) % (nsect - 20 * PER_THREAD);
      s = base + (k >> 1) * 20 + (k & 1);
    }
    ldg256(buf + s * 8, v[k]);
  }
#pragma unroll
  for (int k = 0; k < PER_THREAD; ++k)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc += v[k][j];
  if (acc == 123.456f) sink[0] = acc;
}

template <int PATTERN, int PER_THREAD>
void run(const char *name, const float *buf, size_t nsect, float *sink, int blocks) {
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  for (int i = 0; i < 3; ++i) probe<PATTERN, PER_THREAD><<<blocks, 256>>>(buf, nsect, sink);
  cudaEventRecord(e0);
  const int reps = 10;
  for (int i = 0; i < reps; ++i) probe<PATTERN, PER_THREAD><<<blocks, 256>>>(buf, nsect, sink);
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms;
  cudaEventElapsedTime(&ms, e0, e1);
  const double bytes = (double)blocks * 256 * PER_THREAD * 32;
  printf("%-44s %2d x 32 B/thread, %6d blocks: %7.1f us  %7.1f GB/s (requested bytes)\n", name, PER_THREAD, blocks,
         ms / reps * 1e3, bytes / (ms / reps * 1e-3) / 1e9);
}

int main() {
  const size_t bytes = (size_t)1 << 30;
  float *buf, *sink;
  cudaMalloc(&buf, bytes);
  cudaMalloc(&sink, 4);
  cudaMemset(buf, 0, bytes);
  const size_t nsect = bytes / 32;
  // ~160 MB requested per launch (like one lookup)
  run<0, 8>("linear stream", buf, nsect, sink, 2560);
  run<1, 8>("random 32-B sectors", buf, nsect, sink, 2560);
  run<2, 8>("random 64-B pairs", buf, nsect, sink, 2560);
  run<3, 20>("window 10 rows x 64 B @640 B pitch", buf, nsect, sink, 1024);
  run<1, 20>("random 32-B sectors", buf, nsect, sink, 1024);
  run<1, 8>("random 32-B sectors (4x more data)", buf, nsect, sink, 10240);
  run<2, 8>("random 64-B pairs (4x more data)", buf, nsect, sink, 10240);
  run<0, 8>("linear stream (4x more data)", buf, nsect, sink, 10240);
  // one thread per query slice: 30720 x 32 = 983040 threads = 3840 blocks
  run<4, 20>("per-slice window 10 rows x 64 B", buf, nsect, sink, 3840);
  run<4, 24>("per-slice window 12 rows x 64 B", buf, nsect, sink, 3840);
  run<5, 10>("per-slice window 10 rows x 32 B", buf, nsect, sink, 3840);
  run<5, 12>("per-slice window 12 rows x 32 B", buf, nsect, sink, 3840);
  cudaError_t e = cudaDeviceSynchronize();
  printf("status: %s\n", cudaGetErrorString(e));
  return 0;
}
