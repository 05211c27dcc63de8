// issue_port_probe.cu -- what does a non-FMA instruction cost an FFMA2 stream that runs at the FP32 peak?
// 84 independent FFMA2 per iteration (x, y fixed: the FMAs read nothing from the loads) + NLD shared-memory loads
// (broadcast address, off one base register; one LOP3 per load keeps it alive), 8 warps per SM.
// Measured on a B200 (TFLOP/s): no loads 67.2; 7 / 14 / 28 x LDS.128: 61.6 / 58.5 / 51.4; 14 / 28 / 56 x LDS.32:
// 58.6 / 51.4 / 43.8.  The cost is per INSTRUCTION, not per byte (14 x LDS.32 = 14 x LDS.128), and it follows
//     FMA utilisation = 2 N_ffma2 / (2 N_ffma2 + N_other)        (28 others: 85.7 % -> 87 measured; 56: 75 -> 76.5)
// i.e. an FFMA2 holds its scheduler's issue port for two cycles and the FP32 peak IS the issue rate: at the peak
// there are no spare issue slots, every load, address computation, branch or barrier poll displaces an FMA cycle.
// (The name is historical: the first hypothesis was a register-file write-port conflict; the byte-independence of
// the cost rules that out.)
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o _build/issue_port_probe.bin issue_port_probe.cu
#include <cuda_runtime.h>
#include <stdio.h>

template <int NLD, int WIDTH>
__global__ void __launch_bounds__(256, 1) probe(float *sink, int iters, float seed) {
  __shared__ float4 sm[64];
  if (threadIdx.x < 64) sm[threadIdx.x] = make_float4(seed + threadIdx.x, 1.f, seed, 2.f);
  __syncthreads();
  float2 acc[84];
#pragma unroll
  for (int i = 0; i < 84; ++i) acc[i] = make_float2(seed + threadIdx.x + i, seed - i);
  const float2 x = make_float2(1.0000001f + seed, 0.9999999f - seed), y = make_float2(0.5f * seed, 1.f - seed);
  unsigned mix = 0;
#pragma unroll 1
  for (int it = 0; it < iters; ++it) {
    // loads off one base address with immediate offsets (no per-load integer work); the results are dead: the
    // volatile asm keeps the loads, nothing waits for them except the register write-back itself
    const unsigned base = (unsigned)__cvta_generic_to_shared(&sm[(it & 1) * 32]);
#pragma unroll
    for (int l = 0; l < NLD; ++l) {
      unsigned u0, u1, u2, u3;
      if (WIDTH == 16)
        asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(u0), "=r"(u1), "=r"(u2), "=r"(u3) : "r"(base + (l % 32) * 16));
      else
        asm volatile("ld.shared.u32 %0, [%1];" : "=r"(u0) : "r"(base + (l % 32) * 16));
      mix ^= u0;   // one LOP3 per load keeps it alive (a v4 load still writes its four registers)
    }
#pragma unroll
    for (int i = 0; i < 84; ++i) acc[i] = __ffma2_rn(x, y, acc[i]);
  }
  float s = __uint_as_float(mix & 1);
#pragma unroll
  for (int i = 0; i < 84; ++i) s += acc[i].x + acc[i].y;
  if (s == 12345.678f) sink[0] = s;
}

template <int NLD, int WIDTH>
void run(float *sink) {
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  const int iters = 4000, blocks = 148;
  probe<NLD, WIDTH><<<blocks, 256>>>(sink, 8, 0.f);
  float best = 1e30f;
  for (int r = 0; r < 5; ++r) {
    cudaEventRecord(e0);
    probe<NLD, WIDTH><<<blocks, 256>>>(sink, iters, 0.f);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    best = ms < best ? ms : best;
  }
  const double tf = 4.0 * 84 * (double)iters * blocks * 256 / (best * 1e-3) / 1e12;
  printf("%2d x LDS.%-3d (+ 1 LOP3 each) per 84 FFMA2: %6.2f TFLOP/s   (issue model 168/(168+2*NLD): %5.1f %% of the NLD=0 line)\n", NLD,
         WIDTH * 8, tf, 100.0 * 168 / (168 + 2 * NLD));
}

int main() {
  float *sink;
  cudaMalloc(&sink, 4);
  run<0, 16>(sink);
  run<7, 16>(sink);
  run<14, 16>(sink);
  run<28, 16>(sink);
  run<14, 4>(sink);
  run<28, 4>(sink);
  run<56, 4>(sink);
  printf("status: %s\n", cudaGetErrorString(cudaDeviceSynchronize()));
  return 0;
}
