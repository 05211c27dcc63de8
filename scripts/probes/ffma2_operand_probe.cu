// ffma2_operand_probe.cu -- what does a packed FFMA2 stream cost when its operands come from the register file
// instead of the operand-reuse latches?  16 warps per SM (4 per scheduler), NACC independent accumulator pairs.
//   P1: acc[i] = fma2(x, y, acc[i])                 x, y fixed            (the peak probe)
//   P2: acc[i] = fma2(a[i % 4], y, acc[i])          A cycles over 4 pairs, B fixed
//   P3: acc[i] = fma2(a[i % 4], b[i % 3], acc[i])   A and B both change every instruction
//   P4: acc[i] = fma2(a[i / 8 % 4], b[i % 8], ...)  B changes every instruction, A every 8th
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o _build/ffma2_operand_probe.bin ffma2_operand_probe.cu
#include <cuda_runtime.h>
#include <stdio.h>

template <int P, int NACC>
__global__ void __launch_bounds__(512, 1) probe(float *sink, int iters, float seed) {
  float2 acc[NACC], a[4], b[8];
#pragma unroll
  for (int i = 0; i < NACC; ++i) acc[i] = make_float2(seed + threadIdx.x + i, seed - i);
#pragma unroll
  for (int i = 0; i < 4; ++i) a[i] = make_float2(1.0f + seed * (i + 1), 1.0f - seed * (i + 2));
#pragma unroll
  for (int i = 0; i < 8; ++i) b[i] = make_float2(0.5f * seed * (i + 1), seed * (i + 3));
#pragma unroll 1
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int u = 0; u < 4; ++u)
#pragma unroll
      for (int i = 0; i < NACC; ++i) {
        float2 x = a[0], y = b[0];
        if (P == 2) x = a[i % 4];
        if (P == 3) { x = a[i % 4]; y = b[i % 3]; }
        if (P == 4) { x = a[(i / 8) % 4]; y = b[i % 8]; }
        acc[i] = __ffma2_rn(x, y, acc[i]);
      }
    // keep a[] / b[] from being folded: rotate them slowly
    const float2 t = a[0]; a[0] = a[1]; a[1] = a[2]; a[2] = a[3]; a[3] = t;
  }
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < NACC; ++i) s += acc[i].x + acc[i].y;
  if (s == 12345.678f) sink[0] = s;
}

template <int P, int NACC>
void run(const char *name, float *sink) {
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  const int iters = 2000, blocks = 148;
  probe<P, NACC><<<blocks, 512>>>(sink, 8, 0.f);
  float best = 1e30f;
  for (int r = 0; r < 5; ++r) {
    cudaEventRecord(e0);
    probe<P, NACC><<<blocks, 512>>>(sink, iters, 0.f);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    best = ms < best ? ms : best;
  }
  cudaFuncAttributes fa;
  cudaFuncGetAttributes(&fa, probe<P, NACC>);
  printf("%-44s NACC %2d regs %3d  %6.2f TFLOP/s\n", name, NACC, fa.numRegs, 4.0 * 4 * NACC * (double)iters * blocks * 512 / (best * 1e-3) / 1e12);
}

int main() {
  float *sink;
  cudaMalloc(&sink, 4);
  run<1, 16>("P1 x, y fixed", sink);
  run<2, 16>("P2 A cycles (4), B fixed", sink);
  run<3, 16>("P3 A and B change every instruction", sink);
  run<4, 16>("P4 B changes every instruction", sink);
  run<1, 40>("P1 x, y fixed", sink);
  run<2, 40>("P2 A cycles (4), B fixed", sink);
  run<3, 40>("P3 A and B change every instruction", sink);
  run<4, 40>("P4 B changes every instruction", sink);
  printf("status: %s\n", cudaGetErrorString(cudaDeviceSynchronize()));
  return 0;
}
