// l2_read_probe.cu -- what can the SMs pull out of the L2 with plain loads?  The ceiling of alt_cuda_corr's
// forward, which re-reads every 1 KB feature vector of a 10x10 window per query (3.07 GB requested, 1.9 GB after L1
// hits, per level-0 call at B=4).  A buffer that fits the 126 MB L2 is read `passes` times:
//   pattern 0: every warp instruction reads 512 contiguous bytes (LDG.128, what altcorr does), L1 bypassed (.cg)
//   pattern 1: the same through the L1 (ld.global.ca / default)
//   pattern 2: LDG.128 with L1::no_allocate
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o _build/l2_read_probe.bin l2_read_probe.cu
#include <cuda_runtime.h>
#include <stdio.h>

template <int PATTERN>
__global__ void __launch_bounds__(256) probe(const float4 *buf, size_t n4, int passes, float *sink) {
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (int p = 0; p < passes; ++p) {
    // rotate the start so that an SM does not re-read what its own L1 holds
    size_t i = ((size_t)blockIdx.x * blockDim.x + threadIdx.x + (size_t)p * 1237 * blockDim.x) % n4;
    for (size_t k = 0; k < n4 / stride; ++k) {
      float4 v;
      if (PATTERN == 0) v = __ldcg(buf + i);
      else if (PATTERN == 1) v = buf[i];
      else asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(buf + i));
      acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
      i += stride;
      if (i >= n4) i -= n4;
    }
  }
  if (acc.x + acc.y + acc.z + acc.w == 12345.678f) sink[0] = acc.x;
}

template <int PATTERN>
void run(const char *name, const float4 *buf, size_t n4, int ctas_per_sm, float *sink) {
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  const int passes = 20;
  probe<PATTERN><<<148 * ctas_per_sm, 256>>>(buf, n4, 2, sink);
  cudaEventRecord(e0);
  probe<PATTERN><<<148 * ctas_per_sm, 256>>>(buf, n4, passes, sink);
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms;
  cudaEventElapsedTime(&ms, e0, e1);
  const size_t stride = (size_t)148 * ctas_per_sm * 256;
  const double bytes = (double)(n4 / stride) * stride * 16.0 * passes;
  printf("%-44s %d CTAs/SM  %8.1f GB/s\n", name, ctas_per_sm, bytes / (ms * 1e-3) / 1e9);
}

int main() {
  const size_t bytes = 48u << 20;
  float4 *buf;
  float *sink;
  cudaMalloc(&buf, bytes);
  cudaMalloc(&sink, 4);
  cudaMemset(buf, 0, bytes);
  for (int c : {2, 4, 8}) {
    run<0>("L2 read, LDG.128 .cg (L1 bypass)", buf, bytes / 16, c, sink);
    run<1>("L2 read, LDG.128 default (through L1)", buf, bytes / 16, c, sink);
    run<2>("L2 read, LDG.128 nc L1::no_allocate", buf, bytes / 16, c, sink);
  }
  printf("status: %s\n", cudaGetErrorString(cudaDeviceSynchronize()));
  return 0;
}
