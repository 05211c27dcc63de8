// raft_volume_bwd.cu -- backward of the all-pairs volume + pyramid w.r.t. both feature maps.
//
// Reference: autograd through CorrBlock.__init__ (models/raft/corr.py:55-64,98-106):
//     vol0 = f1^T f2 / sqrt(C),  vol_{l+1} = avg_pool2d(vol_l, 2, 2)
// gives, for the per-level gradients G_l that the lookups leave behind (b200corr_lookup_backward),
//     dF1 = scale * fold(G)   . F2^T          fold = sum_l unpool_l(G_l) / 4^l  (a pass over the whole pyramid)
//     dF2 = scale * fold(G)^T . F1
// i.e. three avg_pool2d backward passes over (B*HW)^2-sized tensors and two cuBLAS GEMMs in the reference.
//
// Average pooling commutes with the contraction (vol_l = f1 . pool_l(f2), the identity AlternateCorrBlock is
// built on, corr.py:114-137), so the fold moves from the gradient VOLUME (1.25 GB at B=4, 48x160) to the
// FEATURE maps (31 MB):
//     dF1        = scale * sum_l G_l . pool_l(F2)^T                     one GEMM, K runs over the keys of all levels
//     dpool_l(F2) = scale * G_l^T . F1                                  one GEMM per level
//     dF2        = sum_l unpool_l(dpool_l(F2)) / 4^l                    elementwise on feature-sized tensors
// The gradient pyramid is read exactly twice and never folded.
//
// Kernels:
//   * volgrad_tc_kernel<MODE>: tcgen05 (kind::tf32) GEMM, M = 128 queries (MODE 0) or 128 keys of one level
//     (MODE 1), N = channels, K in blocks of 32 through a 4-stage TMA ring; the operands are the fp32 tensors
//     themselves (the tensor core reads the upper 19 bits: TF32 by truncation), G_l straight from the gradient
//     pyramid -- K-major for dF1, MN-major for the transposed product (no transposed copy of a 1 GB tensor).
//     Work items (256 rows) are split along K so that 148 SMs get ~5 equal waves; partial sums land with red.global.add.
//   * gemm_nt_simt_kernel: exact fp32 (precision "fp32", and every shape the tensor maps cannot express).
#include <cstdlib>
#include <cstring>

#include "tcgen05.cuh"

namespace {
using namespace b200dev;

constexpr int kMaxLv = 8;

// ------------------------------------------------------------------------------ feature pooling
// out[b,c,y,x] = mean of in[b,c,2y..2y+1,2x..2x+1]   (avg_pool2d(2, 2), floor mode)
__global__ void __launch_bounds__(256)
pool_feat_kernel(const float *__restrict__ in, float *__restrict__ out, long long planes, int Hi, int Wi) {
  const int Ho = Hi / 2, Wo = Wi / 2;
  const long long total = planes * Ho * Wo;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int x = (int)(i % Wo);
    const long long t = i / Wo;
    const int y = (int)(t % Ho);
    const long long pl = t / Ho;
    const float *p = in + (pl * Hi + 2 * y) * Wi + 2 * x;
    out[i] = 0.25f * ((p[0] + p[1]) + (p[Wi] + p[Wi + 1]));
  }
}

// g2[b,c,y,x] += sum_{l>=1} dpool_l[b,c,y>>l,x>>l] / 4^l   (cells past the floor-mode extent contribute nothing)
struct UnpoolParams {
  const float *lv[kMaxLv];
  int LH[kMaxLv], LW[kMaxLv];
  int num_levels;
};
__global__ void __launch_bounds__(256)
unpool_acc_kernel(float *__restrict__ g2, const UnpoolParams p, long long planes, int H, int W) {
  const long long total = planes * H * W;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int x = (int)(i % W);
    const long long t = i / W;
    const int y = (int)(t % H);
    const long long pl = t / H;
    float s = g2[i], w = 0.25f;
    for (int l = 1; l < p.num_levels; ++l, w *= 0.25f) {
      const int yl = y >> l, xl = x >> l;
      if (yl < p.LH[l] && xl < p.LW[l]) s += w * p.lv[l][(pl * p.LH[l] + yl) * p.LW[l] + xl];
    }
    g2[i] = s;
  }
}

// ------------------------------------------------------------------------------ exact SIMT GEMM
// C[m][n] (+)= alpha * sum_k A(m,k) * B(n,k); element strides for every operand; batch = blockIdx.z.
struct SimtGemm {
  const float *A, *B;
  float *C;
  long long sAm, sAk, sBn, sBk, sCm, sCn, bA, bB, bC;
  int M, N, K;
  float alpha;
  int accumulate;
};
__global__ void __launch_bounds__(256)
gemm_nt_simt_kernel(const SimtGemm g) {
  constexpr int TM = 64, TN = 64, TK = 16;
  __shared__ float As[TK][TM + 4], Bs[TK][TN + 4];
  const int tid = threadIdx.x, tm = tid & 15, tn = tid >> 4;
  const int m0 = blockIdx.x * TM, n0 = blockIdx.y * TN;
  const float *A = g.A + blockIdx.z * g.bA, *B = g.B + blockIdx.z * g.bB;
  float acc[4][4] = {};
  for (int k0 = 0; k0 < g.K; k0 += TK) {
    // loaders walk the unit-stride axis fastest
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int e = tid + 256 * i;
      int mm, kk;
      if (g.sAk == 1) { kk = e & 15; mm = e >> 4; } else { mm = e & 63; kk = e >> 6; }
      const int m = m0 + mm, k = k0 + kk;
      As[kk][mm] = (m < g.M && k < g.K) ? A[m * g.sAm + k * g.sAk] : 0.f;
      int nn;
      if (g.sBk == 1) { kk = e & 15; nn = e >> 4; } else { nn = e & 63; kk = e >> 6; }
      const int n = n0 + nn;
      const int k2 = k0 + kk;
      Bs[kk][nn] = (n < g.N && k2 < g.K) ? B[n * g.sBn + k2 * g.sBk] : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < TK; ++kk) {
      float a[4], b[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) { a[i] = As[kk][tm * 4 + i]; b[i] = Bs[kk][tn * 4 + i]; }
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }
  float *C = g.C + blockIdx.z * g.bC;
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int m = m0 + tm * 4 + i, n = n0 + tn * 4 + j;
      if (m < g.M && n < g.N) {
        float *c = C + m * g.sCm + n * g.sCn;
        *c = g.accumulate ? *c + g.alpha * acc[i][j] : g.alpha * acc[i][j];
      }
    }
}

// ------------------------------------------------------------------------------ tcgen05 GEMM
namespace vb {
// One work item covers TWO 128-row MMA tiles (256 rows, two TMEM accumulators) so that every 32 KB block of the
// N operand (the feature map) fetched through the L2 serves 256 rows: with single tiles the feature-map re-reads
// were two thirds of the L2->SM traffic (3.7 GB per product at B=4) and the kernels ran at 0.30 / 0.33 ms.
constexpr int BM = 128, TM = 2 * BM, BK = 32, NST = 3, THREADS = 320, TMEM_COLS = 512;
constexpr int A_BYTES = TM * BK * 4, B_BYTES = 256 * BK * 4, STAGE_BYTES = A_BYTES + B_BYTES;
constexpr int SMEM_BYTES = 1024 + NST * STAGE_BYTES + 256;
constexpr int KSPLIT = 6;

struct Maps {
  CUtensorMap a[kMaxLv];   // gradient level l: (HW_l, HW, B)
  CUtensorMap b[kMaxLv];   // MODE 0: pool_l(F2) as (HW_l, C, B);  MODE 1: b[0] = F1 as (HW, C, B)
};
struct Seg {
  short level;
  int kb0, nkb;            // K blocks [kb0, kb0 + nkb) of that level
};
struct Params {
  int B, C, N, HW, num_levels;        // N = channels rounded up to 16 (MMA N, rows of the B box)
  int HWl[kMaxLv];
  float *out[kMaxLv];                 // MODE 0: out[0] = dF1;  MODE 1: out[l] = dpool_l(F2)
  float scale;
  // MODE 0: item = (b, mtile, split); the K range of split s is segs[s][0..nseg[s])
  int mtiles0, nseg[KSPLIT];
  Seg segs[KSPLIT][kMaxLv];
  // MODE 1: item = (level, b, mtile, split): prefix[l] = first item of level l; K = queries in blocks of 32
  int prefix[kMaxLv + 1], mtiles[kMaxLv], kb_per_split, kb_total;
  int total_items;
};

// MN-major TF32 operand tile.  For 32-bit MN-major operands the tensor core knows ONE swizzled layout
// (cutlass sm100_common.inl: "for mn-major tf32 operands, SW128_32B is the only available smem layout"): rows of
// 128 bytes = 32 consecutive M elements at one k, 128-byte swizzle with 32-byte atomicity (byte-address bits
// [5,7) XOR bits [7,9): the pattern repeats every 4 rows).  TMA writes it with
// CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B; descriptor layout type 1 (SWIZZLE_128B_BASE32B), canonical form
// ((8,n),(4,k)) in 16-byte units: LBO = distance between 32-element groups along M, SBO = distance between
// 4-row groups along K.
__device__ __forceinline__ uint64_t umma_desc_mnmajor_sw128_32b(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);
  d |= (uint64_t)(lbo_bytes >> 4) << 16;
  d |= (uint64_t)(sbo_bytes >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)1 << 61;                     // SWIZZLE_128B_BASE32B
  return d;
}
}  // namespace vb

template <int MODE>
__global__ void __launch_bounds__(vb::THREADS, 1)
volgrad_tc_kernel(const __grid_constant__ vb::Maps maps, const vb::Params p) {
  using namespace vb;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t *sm = smem_raw + (base - raw);
  uint64_t *bars = reinterpret_cast<uint64_t *>(sm + NST * STAGE_BYTES);
  uint64_t *full_bar = bars, *empty_bar = bars + NST, *acc_full = bars + 2 * NST, *acc_empty = acc_full + 1;
  uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(acc_full + 2);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    for (int l = 0; l < p.num_levels; ++l) {
      tma_prefetch_desc(&maps.a[l]);
      if (MODE == 0 || l == 0) tma_prefetch_desc(&maps.b[l]);
    }
    for (int s = 0; s < NST; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    mbar_init(acc_full, 1);
    mbar_init(acc_empty, 8);
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, TMEM_COLS);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  // item -> (level, batch, first row, K segments)
  auto decode = [&](int item, int &lvl, int &b, int &m0, int &split) {
    if (MODE == 0) {
      split = item % KSPLIT;
      const int t = item / KSPLIT;
      m0 = (t % p.mtiles0) * TM;
      b = t / p.mtiles0;
      lvl = 0;
    } else {
      lvl = 0;
      while (lvl + 1 < p.num_levels && p.prefix[lvl + 1] <= item) ++lvl;
      int r = item - p.prefix[lvl];
      split = r % KSPLIT;
      r /= KSPLIT;
      m0 = (r % p.mtiles[lvl]) * TM;
      b = r / p.mtiles[lvl];
    }
  };
  // number of K blocks of an item (both the producer and the MMA thread walk the same sequence)
  auto item_kblocks = [&](int split) {
    if (MODE == 0) {
      int n = 0;
      for (int s = 0; s < p.nseg[split]; ++s) n += p.segs[split][s].nkb;
      return n;
    }
    const int k0 = split * p.kb_per_split;
    const int k1 = k0 + p.kb_per_split < p.kb_total ? k0 + p.kb_per_split : p.kb_total;
    return k1 > k0 ? k1 - k0 : 0;
  };

  if (warp == 0) {
    if (lane == 0) {
      uint32_t it = 0;
      for (int item = blockIdx.x; item < p.total_items; item += gridDim.x) {
        int lvl, b, m0, split;
        decode(item, lvl, b, m0, split);
        auto load = [&](int l, int kb) {
          const int st = it % NST;
          mbar_wait(&empty_bar[st], ((it / NST) & 1) ^ 1);
          uint8_t *a = sm + st * STAGE_BYTES;
          mbar_arrive_expect_tx(&full_bar[st], A_BYTES + (uint32_t)p.N * BK * 4);
          if (MODE == 0) {
            tma_load_3d(a, &maps.a[l], &full_bar[st], kb * BK, m0, b);              // G_l rows = queries, K = keys
            tma_load_3d(a + 16384, &maps.a[l], &full_bar[st], kb * BK, m0 + BM, b);
            tma_load_3d(a + A_BYTES, &maps.b[l], &full_bar[st], kb * BK, 0, b);     // pool_l(F2): rows = channels
          } else {
#pragma unroll
            for (int j = 0; j < 8; ++j)                                             // G_l: K = queries (rows), M = keys
              tma_load_3d(a + j * 4096, &maps.a[l], &full_bar[st], m0 + 32 * j, kb * BK, b);
            tma_load_3d(a + A_BYTES, &maps.b[0], &full_bar[st], kb * BK, 0, b);     // F1: rows = channels, K = queries
          }
          ++it;
        };
        if (MODE == 0) {
          for (int s = 0; s < p.nseg[split]; ++s)
            for (int kb = 0; kb < p.segs[split][s].nkb; ++kb) load(p.segs[split][s].level, p.segs[split][s].kb0 + kb);
        } else {
          const int n = item_kblocks(split);
          for (int kb = 0; kb < n; ++kb) load(lvl, split * p.kb_per_split + kb);
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      const uint32_t idesc = umma_idesc_tf32(BM, p.N) | (MODE == 1 ? (1u << 15) : 0u);   // MODE 1: A is MN-major
      uint32_t it = 0, ic = 0;
      for (int item = blockIdx.x; item < p.total_items; item += gridDim.x, ++ic) {
        int lvl, b, m0, split;
        decode(item, lvl, b, m0, split);
        const int n = item_kblocks(split);
        mbar_wait(acc_empty, (ic & 1) ^ 1);
        tc_fence_after();
        for (int kb = 0; kb < n; ++kb, ++it) {
          const int st = it % NST;
          mbar_wait(&full_bar[st], (it / NST) & 1);
          tc_fence_after();
          const uint32_t a_addr = base + st * STAGE_BYTES;
          const uint64_t bdesc = umma_desc_kmajor_sw128(a_addr + A_BYTES);
#pragma unroll
          for (int h = 0; h < 2; ++h)
#pragma unroll
            for (int k4 = 0; k4 < 4; ++k4) {
              const uint64_t adesc = MODE == 0 ? umma_desc_kmajor_sw128(a_addr + h * 16384) + 2 * k4
                                               : umma_desc_mnmajor_sw128_32b(a_addr + h * 16384 + k4 * 1024, 4096, 512);
              umma_tf32(tmem_base + h * 256, adesc, bdesc + 2 * k4, idesc, (kb | k4) != 0);
            }
          umma_commit(&empty_bar[st]);
        }
        umma_commit(acc_full);
      }
    }
  } else {
    // ================= epilogue: warps 2..9 = (accumulator h, TMEM lane quarter)
    const int wq = warp & 3, h = (warp - 2) >> 2;
    uint32_t ic = 0;
    for (int item = blockIdx.x; item < p.total_items; item += gridDim.x, ++ic) {
      int lvl, b, m0, split;
      decode(item, lvl, b, m0, split);
      const int n = item_kblocks(split);
      const int m = m0 + h * BM + wq * 32 + lane;
      const int mlim = MODE == 0 ? p.HW : p.HWl[lvl];
      float *out = (MODE == 0 ? p.out[0] : p.out[lvl]) + (size_t)b * p.C * mlim + m;
      mbar_wait(acc_full, ic & 1);
      tc_fence_after();
      if (n > 0) {
        for (int c0 = 0; c0 < p.C; c0 += 32) {
          float v[32];
          tmem_ld_32x32(tmem_base + h * 256 + c0 + ((uint32_t)(wq * 32) << 16), v);
          tmem_ld_wait();
          if (m < mlim) {
#pragma unroll
            for (int i = 0; i < 32; ++i)
              if (c0 + i < p.C) atomicAdd(out + (size_t)(c0 + i) * mlim, p.scale * v[i]);   // RED.ADD.F32: lanes = consecutive m
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(acc_empty);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, TMEM_COLS);
}

int level_dims(int H, int W, int num_levels, int *LH, int *LW, int *HWl) {
  for (int l = 0; l < num_levels; ++l) {
    LH[l] = H >> l;
    LW[l] = W >> l;
    // avg_pool2d floor mode applied l times equals H >> l
    HWl[l] = LH[l] * LW[l];
    if (HWl[l] < 1) return -1;
  }
  return 0;
}

size_t pooled_floats(int B, int C, int H, int W, int num_levels) {
  size_t n = 0;
  for (int l = 1; l < num_levels; ++l) n += (size_t)B * C * (H >> l) * (W >> l);
  return n;
}

}  // namespace

extern "C" {

size_t b200corr_volume_backward_workspace_bytes(int num_levels, int B, int C, int H, int W) {
  if (num_levels < 1 || B < 0 || C < 1 || H < 1 || W < 1) return 0;
  return 2 * pooled_floats(B, C, H, W, num_levels) * sizeof(float) + 256;   // pool_l(F2) and d pool_l(F2), l >= 1
}

int b200corr_volume_backward(const float *const *h_grad_levels, int num_levels, const float *fmap1,
                             const float *fmap2, float *grad_fmap1, float *grad_fmap2, int B, int C, int H, int W,
                             float scale, int precision, void *workspace, size_t workspace_bytes, void *stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  B200_CHECK(num_levels >= 1 && num_levels <= kMaxLv, "volume_backward: num_levels must be in [1, %d]", kMaxLv);
  B200_CHECK(B >= 0 && C >= 1 && H >= 1 && W >= 1, "volume_backward: bad sizes");
  if (B == 0) return 0;
  B200_CHECK(h_grad_levels && fmap1 && fmap2 && grad_fmap1 && grad_fmap2, "volume_backward: null pointer");
  int LH[kMaxLv], LW[kMaxLv], HWl[kMaxLv];
  B200_CHECK(level_dims(H, W, num_levels, LH, LW, HWl) == 0, "volume_backward: a pyramid level is empty");
  const size_t need = b200corr_volume_backward_workspace_bytes(num_levels, B, C, H, W);
  B200_CHECK(num_levels == 1 || (workspace && workspace_bytes >= need), "volume_backward: workspace too small (%zu < %zu)",
             workspace_bytes, need);
  const int HW = H * W;
  // workspace: pooled F2 levels, then the gradients w.r.t. them
  const float *f2p[kMaxLv];
  float *g2p[kMaxLv];
  f2p[0] = fmap2;
  g2p[0] = grad_fmap2;
  {
    float *w = reinterpret_cast<float *>((((uintptr_t)workspace) + 127) & ~(uintptr_t)127);
    const size_t half = pooled_floats(B, C, H, W, num_levels);
    float *pp = w, *gp = w + half;
    for (int l = 1; l < num_levels; ++l) {
      f2p[l] = pp;
      g2p[l] = gp;
      pp += (size_t)B * C * HWl[l];
      gp += (size_t)B * C * HWl[l];
      const long long total = (long long)B * C * HWl[l];
      const int nb = (int)((total + 255) / 256 < 65535 ? (total + 255) / 256 : 65535);
      pool_feat_kernel<<<nb, 256, 0, stream>>>(f2p[l - 1], const_cast<float *>(f2p[l]), (long long)B * C, LH[l - 1], LW[l - 1]);
      B200_LAUNCH_OK("pool_feat_kernel");
    }
  }
  for (int l = 0; l < num_levels; ++l) B200_CHECK(h_grad_levels[l], "volume_backward: null gradient level %d", l);

  // tensor-core path: TMA needs 16-byte aligned bases and row pitches, the MMA N <= 256
  bool tc = precision != B200CORR_PREC_FP32 && C <= 256 && (long long)B * HW < (1ll << 31);
  for (int l = 0; l < num_levels && tc; ++l)
    tc = HWl[l] % 4 == 0 && ((uintptr_t)h_grad_levels[l] & 15) == 0 && ((uintptr_t)f2p[l] & 15) == 0;
  tc = tc && ((uintptr_t)fmap1 & 15) == 0 && HW % 4 == 0 && !getenv("B200CORR_VOLBWD_SIMT");

  if (!tc) {
    for (int l = 0; l < num_levels; ++l) {
      SimtGemm g;
      // dF1[b][c][q] (+)= scale * sum_k G_l[b,q,k] * pool_l(F2)[b,c,k]
      g.A = h_grad_levels[l]; g.sAm = HWl[l]; g.sAk = 1; g.bA = (long long)HW * HWl[l];
      g.B = f2p[l]; g.sBn = HWl[l]; g.sBk = 1; g.bB = (long long)C * HWl[l];
      g.C = grad_fmap1; g.sCm = 1; g.sCn = HW; g.bC = (long long)C * HW;
      g.M = HW; g.N = C; g.K = HWl[l]; g.alpha = scale; g.accumulate = l > 0;
      gemm_nt_simt_kernel<<<dim3((g.M + 63) / 64, (g.N + 63) / 64, B), 256, 0, stream>>>(g);
      B200_LAUNCH_OK("gemm_nt_simt_kernel");
      // d pool_l(F2)[b][c][k] = scale * sum_q F1[b,c,q] * G_l[b,q,k]
      g.A = h_grad_levels[l]; g.sAm = 1; g.sAk = HWl[l]; g.bA = (long long)HW * HWl[l];
      g.B = fmap1; g.sBn = HW; g.sBk = 1; g.bB = (long long)C * HW;
      g.C = g2p[l]; g.sCm = 1; g.sCn = HWl[l]; g.bC = (long long)C * HWl[l];
      g.M = HWl[l]; g.N = C; g.K = HW; g.alpha = scale; g.accumulate = 0;
      gemm_nt_simt_kernel<<<dim3((g.M + 63) / 64, (g.N + 63) / 64, B), 256, 0, stream>>>(g);
      B200_LAUNCH_OK("gemm_nt_simt_kernel");
    }
  } else {
    using namespace vb;
    static bool done0[64], done1[64];
    if (int e = b200::set_max_smem_once((const void *)volgrad_tc_kernel<0>, SMEM_BYTES, done0)) return e;
    if (int e = b200::set_max_smem_once((const void *)volgrad_tc_kernel<1>, SMEM_BYTES, done1)) return e;
    // the partial sums of the K splits land with atomics
    B200_CUDA(cudaMemsetAsync(grad_fmap1, 0, sizeof(float) * (size_t)B * C * HW, stream));
    for (int l = 0; l < num_levels; ++l)
      B200_CUDA(cudaMemsetAsync(g2p[l], 0, sizeof(float) * (size_t)B * C * HWl[l], stream));
    Params p;
    memset(&p, 0, sizeof(p));
    p.B = B; p.C = C; p.N = (C + 15) / 16 * 16; p.HW = HW; p.num_levels = num_levels; p.scale = scale;
    Maps m0, m1;
    memset(&m0, 0, sizeof(m0));
    memset(&m1, 0, sizeof(m1));
    for (int l = 0; l < num_levels; ++l) {
      p.HWl[l] = HWl[l];
      const uint64_t dg[3] = {(uint64_t)HWl[l], (uint64_t)HW, (uint64_t)B};
      const uint64_t sg[3] = {4, (uint64_t)HWl[l] * 4, (uint64_t)HWl[l] * 4 * HW};
      const uint32_t box_k[3] = {32, 128, 1};     // MODE 0: 128 query rows x 32 keys (K-major), two boxes per stage
      const uint32_t box_mn[3] = {32, 32, 1};     // MODE 1: 32 query rows (K) x 32 keys (M, contiguous)
      if (int e = b200::make_tensor_map(&m0.a[l], CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, h_grad_levels[l], dg, sg, box_k,
                                        CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B)) return e;
      if (int e = b200::make_tensor_map(&m1.a[l], CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, h_grad_levels[l], dg, sg, box_mn,
                                        CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B)) return e;
      const uint64_t df[3] = {(uint64_t)HWl[l], (uint64_t)C, (uint64_t)B};
      const uint64_t sf[3] = {4, (uint64_t)HWl[l] * 4, (uint64_t)HWl[l] * 4 * C};
      const uint32_t box_f[3] = {32, (uint32_t)p.N, 1};
      if (int e = b200::make_tensor_map(&m0.b[l], CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, f2p[l], df, sf, box_f,
                                        CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B)) return e;
    }
    {
      const uint64_t df[3] = {(uint64_t)HW, (uint64_t)C, (uint64_t)B};
      const uint64_t sf[3] = {4, (uint64_t)HW * 4, (uint64_t)HW * 4 * C};
      const uint32_t box_f[3] = {32, (uint32_t)p.N, 1};
      if (int e = b200::make_tensor_map(&m1.b[0], CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, fmap1, df, sf, box_f,
                                        CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B)) return e;
    }
    // ---- MODE 0: K = the keys of all levels, cut into KSPLIT runs of equal length
    Params p0 = p;
    p0.out[0] = grad_fmap1;
    p0.mtiles0 = (HW + TM - 1) / TM;
    {
      int kbl[kMaxLv], total = 0;
      for (int l = 0; l < num_levels; ++l) { kbl[l] = (HWl[l] + BK - 1) / BK; total += kbl[l]; }
      const int per = (total + KSPLIT - 1) / KSPLIT;
      int l = 0, kb = 0;
      for (int s = 0; s < KSPLIT; ++s) {
        int left = per;
        p0.nseg[s] = 0;
        while (left > 0 && l < num_levels) {
          const int take = kbl[l] - kb < left ? kbl[l] - kb : left;
          if (take > 0) {
            Seg &sg = p0.segs[s][p0.nseg[s]++];
            sg.level = (short)l; sg.kb0 = kb; sg.nkb = take;
            kb += take; left -= take;
          }
          if (kb >= kbl[l]) { ++l; kb = 0; }
        }
      }
    }
    p0.total_items = B * p0.mtiles0 * KSPLIT;
    int grid = p0.total_items < b200::num_sms() ? p0.total_items : b200::num_sms();
    volgrad_tc_kernel<0><<<grid, THREADS, SMEM_BYTES, stream>>>(m0, p0);
    B200_LAUNCH_OK("volgrad_tc_kernel<0>");
    // ---- MODE 1: per level, M = keys, K = the HW queries of one sample
    Params p1 = p;
    p1.kb_total = (HW + BK - 1) / BK;
    p1.kb_per_split = (p1.kb_total + KSPLIT - 1) / KSPLIT;
    int items = 0;
    for (int l = 0; l < num_levels; ++l) {
      p1.out[l] = g2p[l];
      p1.mtiles[l] = (HWl[l] + TM - 1) / TM;
      p1.prefix[l] = items;
      items += B * p1.mtiles[l] * KSPLIT;
    }
    for (int l = num_levels; l <= kMaxLv; ++l) p1.prefix[l] = items;
    p1.total_items = items;
    grid = items < b200::num_sms() ? items : b200::num_sms();
    volgrad_tc_kernel<1><<<grid, THREADS, SMEM_BYTES, stream>>>(m1, p1);
    B200_LAUNCH_OK("volgrad_tc_kernel<1>");
  }
  if (num_levels > 1) {
    UnpoolParams u;
    u.num_levels = num_levels;
    for (int l = 0; l < kMaxLv; ++l) {
      u.lv[l] = l < num_levels ? g2p[l] : nullptr;
      u.LH[l] = l < num_levels ? LH[l] : 0;
      u.LW[l] = l < num_levels ? LW[l] : 0;
    }
    const long long total = (long long)B * C * HW;
    const int nb = (int)((total + 255) / 256 < 65535 ? (total + 255) / 256 : 65535);
    unpool_acc_kernel<<<nb, 256, 0, stream>>>(grad_fmap2, u, (long long)B * C, H, W);
    B200_LAUNCH_OK("unpool_acc_kernel");
  }
  return 0;
}

}  // extern "C"
