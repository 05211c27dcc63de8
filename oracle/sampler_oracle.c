/*
 * oracle/sampler_oracle.c -- CPU restatement of the reference spatial correlation sampler.
 *
 * TEST INFRASTRUCTURE ONLY.  Only tests/, __graft_entry__.smoke() and the cpu_baseline /
 * --impl reference legs of bench.py may load this.  The product (libb200corr.so) never does.
 *
 * Follows, loop for loop, the reference CPU path
 *   /root/reference/models/Pytorch-Correlation-extension/Correlation_Module/correlation.cpp
 *     correlate_patch        :9-37    -> patch_dot_*()
 *     correlate_patch_grad   :40-73   -> patch_grad_*()
 *     correlation_cpp_forward :75-124 -> sampler_oracle_forward_*()
 *     correlation_cpp_backward:126-178-> sampler_oracle_backward_*()
 * Same conventions: patch radius = (patch-1)/2 (integer division, :88-89,:140-141), output size
 * (in + 2*pad - ((k-1)*dil+1))/stride + 1 (:90-94), a product is DROPPED unless both the input1 tap
 * and the shifted input2 tap are inside the image (WITHIN_BOUNDS on both, :24,:28), and per output
 * element the channel / kernel-tap accumulation order is c -> i -> j (:20-35).
 *
 * Parity pinned: tests/test_oracle_cpu.py checks this file against the compiled reference
 * (oracle/_ref, built by oracle/build_ref.py from the reference sources in place) and against the
 * committed golden vectors in tests/golden/ that were generated from that compiled reference.
 *
 * Threading: the reference parallelises forward over (n, ph) and backward over n only
 * (:98-100, :148-149).  Here forward does the same; backward parallelises over (n, c), which keeps
 * the per-element accumulation order (ph, pw, h, w, i, j) of the reference and is therefore
 * bit-identical to it, just not single-threaded at batch 1.
 */
#include <stddef.h>
#include <string.h>

#define INSIDE(a, b, A, B) ((a) >= 0 && (a) < (A) && (b) >= 0 && (b) < (B))

#define DEFINE_ORACLE(T, SUF)                                                                      \
  /* correlation.cpp:9-37 */                                                                       \
  static void patch_dot_##SUF(const T *in1, const T *in2, T *dst, int C, int H, int W, int kH,     \
                              int kW, int dilH, int dilW, int u, int v, int shiftU, int shiftV) {  \
    for (int c = 0; c < C; ++c) {                                                                  \
      const T *p1 = in1 + (size_t)c * H * W;                                                       \
      const T *p2 = in2 + (size_t)c * H * W;                                                       \
      for (int i = 0; i < kH; ++i) {                                                               \
        int i1 = u + i * dilH, i2 = i1 + shiftU;                                                   \
        if (!INSIDE(i1, i2, H, H)) continue;                                                       \
        for (int j = 0; j < kW; ++j) {                                                             \
          int j1 = v + j * dilW, j2 = j1 + shiftV;                                                 \
          if (!INSIDE(j1, j2, W, W)) continue;                                                     \
          *dst += p1[(size_t)i1 * W + j1] * p2[(size_t)i2 * W + j2];                               \
        }                                                                                          \
      }                                                                                            \
    }                                                                                              \
  }                                                                                                \
  /* correlation.cpp:75-124 */                                                                     \
  void sampler_oracle_forward_##SUF(const T *in1, const T *in2, T *out, int B, int C, int H,       \
                                    int W, int kH, int kW, int patchH, int patchW, int padH,       \
                                    int padW, int dilH, int dilW, int dpH, int dpW, int dH,        \
                                    int dW) {                                                      \
    const int radH = (patchH - 1) / 2, radW = (patchW - 1) / 2;                                    \
    const int oH = (H + 2 * padH - ((kH - 1) * dilH + 1)) / dH + 1;                                \
    const int oW = (W + 2 * padW - ((kW - 1) * dilW + 1)) / dW + 1;                                \
    memset(out, 0, sizeof(T) * (size_t)B * patchH * patchW * oH * oW);                             \
    _Pragma("omp parallel for collapse(2) schedule(dynamic)")                                      \
    for (int n = 0; n < B; ++n)                                                                    \
      for (int ph = 0; ph < patchH; ++ph)                                                          \
        for (int pw = 0; pw < patchW; ++pw)                                                        \
          for (int h = 0; h < oH; ++h)                                                             \
            for (int w = 0; w < oW; ++w)                                                           \
              patch_dot_##SUF(in1 + (size_t)n * C * H * W, in2 + (size_t)n * C * H * W,            \
                              out + ((((size_t)n * patchH + ph) * patchW + pw) * oH + h) * oW + w, \
                              C, H, W, kH, kW, dilH, dilW, -padH + h * dH, -padW + w * dW,         \
                              (ph - radH) * dpH, (pw - radW) * dpW);                               \
  }                                                                                                \
  /* correlation.cpp:40-73 + 126-178, restricted to one channel plane (order preserved) */         \
  void sampler_oracle_backward_##SUF(const T *in1, const T *in2, const T *gout, T *gin1, T *gin2,  \
                                     int B, int C, int H, int W, int oH, int oW, int kH, int kW,   \
                                     int patchH, int patchW, int padH, int padW, int dilH,         \
                                     int dilW, int dpH, int dpW, int dH, int dW) {                 \
    const int radH = (patchH - 1) / 2, radW = (patchW - 1) / 2;                                    \
    memset(gin1, 0, sizeof(T) * (size_t)B * C * H * W);                                            \
    memset(gin2, 0, sizeof(T) * (size_t)B * C * H * W);                                            \
    _Pragma("omp parallel for collapse(2) schedule(dynamic)")                                      \
    for (int n = 0; n < B; ++n)                                                                    \
      for (int c = 0; c < C; ++c) {                                                                \
        const T *p1 = in1 + ((size_t)n * C + c) * H * W;                                           \
        const T *p2 = in2 + ((size_t)n * C + c) * H * W;                                           \
        T *g1 = gin1 + ((size_t)n * C + c) * H * W;                                                \
        T *g2 = gin2 + ((size_t)n * C + c) * H * W;                                                \
        for (int ph = 0; ph < patchH; ++ph)                                                        \
          for (int pw = 0; pw < patchW; ++pw) {                                                    \
            const int shiftU = (ph - radH) * dpH, shiftV = (pw - radW) * dpW;                      \
            const T *go = gout + (((size_t)n * patchH + ph) * patchW + pw) * oH * oW;              \
            for (int h = 0; h < oH; ++h)                                                           \
              for (int w = 0; w < oW; ++w) {                                                       \
                const T g = go[(size_t)h * oW + w];                                                \
                const int u = -padH + h * dH, v = -padW + w * dW;                                  \
                for (int i = 0; i < kH; ++i) {                                                     \
                  int i1 = u + i * dilH, i2 = i1 + shiftU;                                         \
                  if (!INSIDE(i1, i2, H, H)) continue;                                             \
                  for (int j = 0; j < kW; ++j) {                                                   \
                    int j1 = v + j * dilW, j2 = j1 + shiftV;                                       \
                    if (!INSIDE(j1, j2, W, W)) continue;                                           \
                    g2[(size_t)i2 * W + j2] += g * p1[(size_t)i1 * W + j1];                        \
                    g1[(size_t)i1 * W + j1] += g * p2[(size_t)i2 * W + j2];                        \
                  }                                                                                \
                }                                                                                  \
              }                                                                                    \
          }                                                                                        \
      }                                                                                            \
  }

DEFINE_ORACLE(float, f32)
DEFINE_ORACLE(double, f64)

/* Output spatial size helper, correlation.cpp:90-94. */
int sampler_oracle_out_size(int in, int pad, int k, int dil, int stride) {
  return (in + 2 * pad - ((k - 1) * dil + 1)) / stride + 1;
}
