// toeplitz_probe.cu -- does the sampler forward's inner loop gain from more warps per scheduler?
// The loop: per channel, 8 in1 values (2 LDS.128) + a window of in2 (LDS.128s) -> acc[t][k] += a[t] * v[t + 2k] as
// packed FFMA2 (pixel pairs).  Variant <NK, THREADS>: NK displacements per thread (21 = the kernel: 168 accumulators,
// 245 registers, 2 warps per scheduler; 11 = the displacements split over two threads: 88 accumulators, window 28
// floats, 3-4 warps per scheduler).  Operands come from conflict-free broadcast loads (isolates pipe + latency).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o _build/toeplitz_probe.bin toeplitz_probe.cu
#include <cuda_runtime.h>
#include <stdio.h>

#if __CUDA_ARCH__ >= 1000 || !defined(__CUDA_ARCH__)
#define FFMA2(a, b, c) __ffma2_rn(a, b, c)
#endif

template <int NK, int THREADS, int MINB>
__global__ void __launch_bounds__(THREADS, MINB) probe(float *sink, int iters, float seed) {
  constexpr int WINQ = (8 + 2 * (NK - 1) + 3) / 4;   // window quads: 12 for NK = 21, 7 for NK = 11
  __shared__ float4 sm[256];
  if (threadIdx.x < 256) sm[threadIdx.x] = make_float4(seed + threadIdx.x, 1.f, seed, 2.f);
  __syncthreads();
  float2 acc[4][NK];
#pragma unroll
  for (int t = 0; t < 4; ++t)
#pragma unroll
    for (int k = 0; k < NK; ++k) acc[t][k] = make_float2(0.f, 0.f);
#pragma unroll 2
  for (int it = 0; it < iters; ++it) {
    const float4 a0 = sm[(it & 3)], a1 = sm[(it & 3) + 4];
    const float2 ap[4] = {make_float2(a0.x, a0.y), make_float2(a0.z, a0.w), make_float2(a1.x, a1.y), make_float2(a1.z, a1.w)};
#pragma unroll
    for (int sg = 0; sg < WINQ; ++sg) {
      const float4 v4 = sm[8 + sg + (it & 3) * 16];
      const float2 vp[2] = {make_float2(v4.x, v4.y), make_float2(v4.z, v4.w)};
#pragma unroll
      for (int uu = 0; uu < 2; ++uu)
#pragma unroll
        for (int tp = 0; tp < 4; ++tp) {
          const int d = 2 * sg + uu - tp;
          if (d >= 0 && d < NK) acc[tp][d] = FFMA2(ap[tp], vp[uu], acc[tp][d]);
        }
    }
  }
  float s = 0.f;
#pragma unroll
  for (int t = 0; t < 4; ++t)
#pragma unroll
    for (int k = 0; k < NK; ++k) s += acc[t][k].x + acc[t][k].y;
  if (s == 12345.678f) sink[0] = s;
}

template <int NK, int THREADS, int MINB>
void run(const char *name, float *sink) {
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  const int iters = 4000, blocks = 148 * MINB;
  probe<NK, THREADS, MINB><<<blocks, THREADS>>>(sink, 8, 0.f);
  float best = 1e30f;
  for (int r = 0; r < 5; ++r) {
    cudaEventRecord(e0);
    probe<NK, THREADS, MINB><<<blocks, THREADS>>>(sink, iters, 0.f);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    best = ms < best ? ms : best;
  }
  cudaFuncAttributes fa;
  cudaFuncGetAttributes(&fa, probe<NK, THREADS, MINB>);
  printf("%-52s regs %3d  %6.2f TFLOP/s\n", name, fa.numRegs, 2.0 * 8 * NK * (double)iters * blocks * THREADS / (best * 1e-3) / 1e12);
}

int main() {
  float *sink;
  cudaMalloc(&sink, 4);
  run<21, 256, 1>("21 k per thread, 2 warps/scheduler (the kernel)", sink);
  run<11, 256, 1>("11 k per thread, 2 warps/scheduler", sink);
  run<11, 384, 1>("11 k per thread, 3 warps/scheduler", sink);
  run<11, 512, 1>("11 k per thread, 4 warps/scheduler", sink);
  run<11, 256, 2>("11 k per thread, 2 CTAs x 2 warps/scheduler", sink);
  run<7, 512, 1>("7 k per thread, 4 warps/scheduler", sink);
  run<7, 768, 1>("7 k per thread, 6 warps/scheduler", sink);
  printf("status: %s\n", cudaGetErrorString(cudaDeviceSynchronize()));
  return 0;
}
