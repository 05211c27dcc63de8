// lookup_convc1.cu -- RAFT correlation lookup fused with the motion encoder's first convolution
// (SURVEY.md 8(f) row 3).
//
// Reference: every RAFT iteration runs  corr = corr_fn(coords1)  (models/raft/raft.py:189; corr.py:72-96) and
// BasicMotionEncoder.forward starts with  cor = F.relu(self.convc1(corr))  (models/raft/update.py:104,111), a
// 1x1 convolution L*(2r+1)^2 -> 256 (324 -> 256 for the basic model, 196 -> 96 for the small one).  The
// (B, 324, H, W) lookup result -- 39.8 MB per iteration at B=4, 48x160 -- is written, read back once by the
// convolution and never used again (only with return_feat_maps, raft.py:191-192).
//
// Here the lookup result never leaves the SM: a CTA owns 128 query pixels; four 4-warp groups gather and
// sample 32 queries each exactly as lookup_fwd_kernel does (same staging, same coordinate arithmetic, same
// bilinear sums), but the (2r+1)^2 samples of a level are rounded to TF32 and written as rows of a K-major
// operand tile in shared memory.  Level by level, one thread issues tcgen05.mma (kind::tf32, M = 128 queries,
// N = output channels in halves of <= 128, K = the level's taps padded to a multiple of 32) against that
// level's slice of the weights, streamed by TMA through a small ring; the fp32 accumulator (128 x n_out) lives
// in TMEM across the levels.  Epilogue: tcgen05.ld -> + bias -> ReLU -> out[b, c, q] (lane = query: every
// store instruction writes one full 128-byte line of one output channel).
//
// Numerics: the samples and the weights are rounded to TF32 (cvt.rna), products accumulate in fp32 -- what
// cuDNN does for this convolution under torch's default `torch.backends.cudnn.allow_tf32 = True`; against an
// fp32 convolution |err| <= 2^-10 * sum_k |w_k * corr_k| (+ fp32 accumulation), asserted in the tests.
#include "raft_lookup.cuh"
#include "tcgen05.cuh"

namespace {
using namespace b200lookup;
using namespace b200dev;

namespace fz {
constexpr int TILE_Q = 128;          // queries per CTA tile = MMA M
constexpr int GROUPS = 4;            // gather groups of 4 warps x 32 queries
constexpr int GATHER_THREADS = 512;
constexpr int THREADS = GATHER_THREADS + 64;   // + warp 16: TMA producer, warp 17: MMA issuer / TMEM allocator
constexpr int NWST = 3;              // weight ring stages
constexpr int W_STAGE_BYTES = 128 * 128;       // <= 128 output channels x 32 taps x 4 bytes
constexpr int TMEM_COLS = 256;

template <int R>
struct G {
  static constexpr int N = 2 * R + 1, NT = N * N, WS = 2 * R + 4;
  static constexpr int KB = (NT + 31) / 32;    // 32-tap blocks per level
  static constexpr int A_BYTES = KB * TILE_Q * 128;
  static constexpr int WIN_FLOATS = WS * kCols * 32;
  static constexpr int TAB = 2 * N * 32;
  static constexpr int SMEM_BYTES = 1024 /*align*/ + A_BYTES + NWST * W_STAGE_BYTES +
                                    GROUPS * (WIN_FLOATS * 4 + TAB * 8) + 256 /*barriers*/;
};

struct Params {
  LookupParams lp;
  const float *bias;     // [n_out] or nullptr
  float *out;            // [B, n_out, HW]
  int n_out, nh, rows_h; // output channels, halves, channels per half (<= 128, multiple of 16)
  int relu;
  int tiles_per_b, total_tiles;
};
}  // namespace fz

__device__ __forceinline__ void bar_sync_named(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

// prepared weights [L][n_out][KP]: tap k of level l is input channel l*NT + k (corr.py:95-96 concatenates the
// levels along the channel axis); TF32-rounded, zero beyond NT
__global__ void __launch_bounds__(256)
convc1_prep_kernel(const float *__restrict__ w, float *__restrict__ wp, int L, int n_out, int NT, int KP) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= L * n_out * KP) return;
  const int k = i % KP, n = (i / KP) % n_out, l = i / (KP * n_out);
  wp[i] = k < NT ? to_tf32_rna(w[(size_t)n * (L * NT) + l * NT + k]) : 0.f;
}

template <int R>
__global__ void __launch_bounds__(fz::THREADS, 1)
lookup_convc1_kernel(const __grid_constant__ CUtensorMap mapW, const fz::Params p, const float *__restrict__ coords) {
  using namespace fz;
  using GG = G<R>;
  constexpr int N = GG::N, WS = GG::WS, RPW = (WS + 3) / 4, KB = GG::KB;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;        // 128-byte swizzle atoms want 1024-byte alignment
  uint8_t *sm = smem_raw + (base - raw);
  uint8_t *a_tile = sm;                                 // [KB][128 rows][128 B], K-major, SW128
  uint8_t *w_ring = sm + GG::A_BYTES;                   // [NWST][rows_h][128 B]
  float *win_all = reinterpret_cast<float *>(w_ring + NWST * W_STAGE_BYTES);
  int *tabr_all = reinterpret_cast<int *>(win_all + GROUPS * GG::WIN_FLOATS);
  float *taba_all = reinterpret_cast<float *>(tabr_all + GROUPS * GG::TAB);
  uint64_t *bars = reinterpret_cast<uint64_t *>(taba_all + GROUPS * GG::TAB);
  uint64_t *w_full = bars, *w_empty = bars + NWST;
  uint64_t *a_full = bars + 2 * NWST, *a_empty = a_full + 1, *acc_full = a_full + 2, *acc_empty = a_full + 3;
  uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(a_full + 4);
  __shared__ float bias_s[256];      // a global load per output channel in the epilogue costs a DRAM/L2 round trip each

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (tid < 256) bias_s[tid] = (p.bias && tid < p.n_out) ? p.bias[tid] : 0.f;
  const LookupParams &lp = p.lp;
  const int L = lp.num_levels;

  if (tid == 0) {
    tma_prefetch_desc(&mapW);
    for (int s = 0; s < NWST; ++s) {
      mbar_init(&w_full[s], 1);
      mbar_init(&w_empty[s], 1);
    }
    mbar_init(a_full, GATHER_THREADS / 32);   // one arrival per gather warp
    mbar_init(a_empty, 1);
    mbar_init(acc_full, 1);
    mbar_init(acc_empty, GATHER_THREADS / 32);
    fence_barrier_init();
  }
  // zero the operand tile once: the padding taps (k >= NT) must read as 0 and are never written again
  for (int i = tid; i < GG::A_BYTES / 16; i += THREADS)
    reinterpret_cast<float4 *>(a_tile)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  if (warp == 17) {
    tmem_alloc(tmem_slot, TMEM_COLS);
    tmem_relinquish();
  }
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 16) {
    // ================= TMA producer: the weight slices, (level, tap block, channel half) in MMA order
    if (lane == 0) {
      uint32_t it = 0;
      for (int t = blockIdx.x; t < p.total_tiles; t += gridDim.x)
        for (int l = 0; l < L; ++l)
          for (int kb = 0; kb < KB; ++kb)
            for (int h = 0; h < p.nh; ++h, ++it) {
              const int st = it % NWST;
              mbar_wait(&w_empty[st], ((it / NWST) & 1) ^ 1);
              mbar_arrive_expect_tx(&w_full[st], (uint32_t)p.rows_h * 128u);
              tma_load_3d(w_ring + st * W_STAGE_BYTES, &mapW, &w_full[st], kb * 32, h * p.rows_h, l);
            }
    }
  } else if (warp == 17) {
    // ================= MMA issuer
    if (lane == 0) {
      const uint32_t idesc = umma_idesc_tf32(TILE_Q, p.rows_h);
      uint32_t it = 0, lc = 0, tc = 0;
      for (int t = blockIdx.x; t < p.total_tiles; t += gridDim.x, ++tc) {
        mbar_wait(acc_empty, (tc & 1) ^ 1);            // the epilogue of the previous tile has drained TMEM
        tc_fence_after();
        for (int l = 0; l < L; ++l, ++lc) {
          mbar_wait(a_full, lc & 1);                   // every gather warp has written its rows of level l
          tc_fence_after();
          for (int kb = 0; kb < KB; ++kb)
            for (int h = 0; h < p.nh; ++h, ++it) {
              const int st = it % NWST;
              mbar_wait(&w_full[st], (it / NWST) & 1);
              tc_fence_after();
              const uint64_t adesc = umma_desc_kmajor_sw128(base + kb * (TILE_Q * 128));
              const uint64_t bdesc = umma_desc_kmajor_sw128(base + GG::A_BYTES + st * W_STAGE_BYTES);
#pragma unroll
              for (int k4 = 0; k4 < 4; ++k4)
                umma_tf32(tmem_base + h * 128, adesc + 2 * k4, bdesc + 2 * k4, idesc, (l | kb | k4) != 0);
              umma_commit(&w_empty[st]);
            }
          umma_commit(a_empty);                        // the operand tile may be overwritten with the next level
        }
        umma_commit(acc_full);
      }
    }
  } else {
    // ================= gather groups + epilogue
    const int grp = warp >> 2, wg = warp & 3;           // group, warp inside the group
    float *win = win_all + grp * GG::WIN_FLOATS;
    int(*tab_r)[32] = reinterpret_cast<int(*)[32]>(tabr_all + grp * GG::TAB);
    float(*tab_a)[32] = reinterpret_cast<float(*)[32]>(taba_all + grp * GG::TAB);
    const int row = grp * 32 + lane;                    // row of the operand tile = query inside the tile
    uint8_t *a_row = a_tile + row * 128;
    const int rsw = row & 7;
    uint32_t lc = 0, tc = 0;
    for (int t = blockIdx.x; t < p.total_tiles; t += gridDim.x, ++tc) {
      const int b = t / p.tiles_per_b;
      const int q = (t - b * p.tiles_per_b) * TILE_Q + row;
      const bool q_ok = q < lp.HW;
      float cx = 0.f, cy = 0.f;
      if (q_ok) {
        cx = coords[((size_t)b * 2 + 0) * lp.HW + q];
        cy = coords[((size_t)b * 2 + 1) * lp.HW + q];
      }
      for (int lvl = 0; lvl < L; ++lvl, ++lc) {
        const int mode = lp.mode, path = lp.path[lvl];
        StageArgs a;
        a.q_ok = q_ok;
        a.LH = lp.LH[lvl];
        a.LW = lp.LW[lvl];
        int xlo, xhi;
        const int slvl = lvl + lp.first_level;
        a.ox = window_origin<R>(cx, slvl, xlo, xhi);
        a.oy = window_origin<R>(cy, slvl, a.ylo, a.yhi);
        const int shift = path == PATH_SCALAR ? 0 : (a.ox & 3);
        a.clo = shift + xlo; a.chi = shift + xhi;
        a.slice = lp.lvl[lvl] + ((size_t)b * lp.HW + (q_ok ? q : 0)) * (size_t)lp.slice[lvl];
        a.blocked = lp.blocked[lvl] != 0;
        a.tiles_w = lp.tiles_w[lvl];
        float sv[RPW][24];
        if (path == PATH_SECTOR) stage_load<PATH_SECTOR, RPW, WS>(a, wg * RPW, sv);
        else if (path == PATH_VEC4) stage_load<PATH_VEC4, RPW, WS>(a, wg * RPW, sv);
        else stage_load<PATH_SCALAR, RPW, WS>(a, wg * RPW, sv);
        // the group's previous level is fully sampled before its tables / window tile are overwritten
        bar_sync_named(1 + grp, 128);
        const float smx = (float)(a.LW - 1), smy = (float)(a.LH - 1);
        const float ismx = __frcp_rn(smx), ismy = __frcp_rn(smy);
#pragma unroll 1
        for (int e = wg; e < 2 * N; e += 4) {
          const bool isy = e >= N;
          int rel;
          float frac;
          one_tap<R>(isy ? cy : cx, slvl, isy ? e - N : e, isy ? smy : smx, isy ? ismy : ismx, mode, isy ? a.oy : a.ox, rel, frac);
          const int lo = isy ? a.ylo : xlo, hi = isy ? a.yhi : xhi;
          if (rel >= 0 && (rel < lo || rel + 1 > hi)) { rel = -1; frac = 0.f; }
          tab_r[e][lane] = rel;
          tab_a[e][lane] = frac;
        }
        if (path == PATH_SECTOR) stage_store<PATH_SECTOR, RPW, WS>(win, lane, a, wg * RPW, sv);
        else if (path == PATH_VEC4) stage_store<PATH_VEC4, RPW, WS>(win, lane, a, wg * RPW, sv);
        else stage_store<PATH_SCALAR, RPW, WS>(win, lane, a, wg * RPW, sv);
        bar_sync_named(1 + grp, 128);
        // the MMAs of the previous level have read the operand tile
        mbar_wait(a_empty, (lc & 1) ^ 1);

        // ---- sample: the same sums as lookup_fwd_kernel, written as TF32 taps k = i*N + j of this query's row
        int rxs[N];
        float axs[N], bxs[N];
        bool fast = true;
#pragma unroll
        for (int i = 0; i < N; ++i) {
          rxs[i] = tab_r[i][lane];
          axs[i] = tab_a[i][lane];
          bxs[i] = 1.f - axs[i];
          fast = fast && rxs[i] == rxs[0] + i && rxs[0] >= 0;
        }
        const float *wl = win + lane;
#pragma unroll 1
        for (int j = wg; j < N; j += 4) {
          const int ry = tab_r[N + j][lane];
          const float ay = tab_a[N + j][lane], by = 1.f - ay;
          float vals[N];
          if (fast && ry >= 0) {
            const float *r0 = wl + (ry * kCols + shift + rxs[0]) * 32;
            float c0[N + 1], c1[N + 1];
#pragma unroll
            for (int k = 0; k <= N; ++k) {
              c0[k] = r0[k * 32];
              c1[k] = r0[(kCols + k) * 32];
            }
#pragma unroll
            for (int i = 0; i < N; ++i) vals[i] = bilerp(c0[i], c0[i + 1], c1[i], c1[i + 1], axs[i], bxs[i], ay, by);
          } else {
#pragma unroll
            for (int i = 0; i < N; ++i) {
              const int rx = rxs[i];
              float v = 0.f;
              if (rx >= 0 && ry >= 0) {
                const float *r0 = wl + (ry * kCols + shift + rx) * 32;
                v = bilerp(r0[0], r0[32], r0[kCols * 32], r0[(kCols + 1) * 32], axs[i], bxs[i], ay, by);
              }
              vals[i] = v;
            }
          }
#pragma unroll
          for (int i = 0; i < N; ++i) {
            const int k = i * N + j, kb = k >> 5, kk = k & 31;
            // K-major, 128-byte swizzle: 16-byte chunk index XOR (row % 8) inside the row's 128 bytes
            float *dst = reinterpret_cast<float *>(a_row + kb * (TILE_Q * 128) + ((((kk >> 2) ^ rsw) << 4) | ((kk & 3) << 2)));
            *dst = q_ok ? to_tf32_rna(vals[i]) : 0.f;
          }
        }
        fence_proxy_async();       // generic-proxy stores -> visible to the tensor core's async-proxy reads
        __syncwarp();
        if (lane == 0) mbar_arrive(a_full);
      }
      // ---- epilogue: TMEM lanes 32*(warp%4).. are query rows 32*(warp%4)..; the four warps sharing a lane
      // quarter split the output channels in blocks of 32
      mbar_wait(acc_full, tc & 1);
      tc_fence_after();
      {
        const int rq = (warp & 3) * 32 + lane;
        const int qe = (t - b * p.tiles_per_b) * TILE_Q + rq;
        const bool ok = qe < lp.HW;
        float *ob = p.out + ((size_t)b * p.n_out) * lp.HW + qe;
        for (int cb = warp >> 2; cb * 32 < p.n_out; cb += 4) {
          const int c0 = cb * 32;
          // channel c lives in accumulator half c / rows_h at column c % rows_h
          const int h = c0 / p.rows_h, col = c0 - h * p.rows_h;
          float v[32];
          tmem_ld_32x32(tmem_base + h * 128 + col + ((uint32_t)((warp & 3) * 32) << 16), v);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            float x = v[i] + bias_s[c0 + i];
            if (p.relu) x = fmaxf(x, 0.f);
            if (ok) ob[(size_t)(c0 + i) * lp.HW] = x;
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
  if (warp == 17) tmem_dealloc(tmem_base, TMEM_COLS);
}

// ---------------------------------------------------------------------------------------------
// The same convolution as a stand-alone tcgen05 GEMM behind the plain lookup kernel ("pipelined" variant):
//     out[b, n, q] = act(sum_k W[n, k] * X[b, k, q] + bias[n])          X = the (B, K, H*W) lookup result
// X is consumed where the lookup kernel left it: (B, K, HW) with q contiguous is an MN-major A operand
// (M = 128 queries, TMA boxes of 32 queries x 32 channels, 128-byte swizzle with 32-byte atoms -- the one layout
// the tensor core accepts for 32-bit MN-major operands), W (n_out, K) is a K-major B operand as it lies.  At B <= 4
// the 39.8 MB of X are still in the 126 MB L2 when this kernel reads them.  TMEM holds two accumulators so that
// the epilogue (bias, ReLU, coalesced stores: lane = query) of a tile overlaps the MMAs of the next.
namespace cv {
constexpr int BM = 128, BK = 32, NST = 3, THREADS = 320;      // warp 0: TMA, warp 1: MMA, warps 2..9: epilogue
constexpr int A_BYTES = 2 * BM * BK * 4, B_BYTES = 256 * BK * 4, STAGE_BYTES = A_BYTES + B_BYTES;
constexpr int SMEM_BYTES = 1024 + NST * STAGE_BYTES + 256;
struct Params {
  int B, K, N, NP, HW, pairs_per_b, total_pairs, kblocks, relu;   // NP = N rounded up to 16
  const float *bias;
  float *out;
};
__device__ __forceinline__ uint64_t desc_mnmajor_sw128_32b(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  return (uint64_t)((smem_addr & 0x3FFFF) >> 4) | ((uint64_t)(lbo_bytes >> 4) << 16) | ((uint64_t)(sbo_bytes >> 4) << 32) |
         ((uint64_t)1 << 46) | ((uint64_t)1 << 61);
}
}  // namespace cv

// One CTA = a PAIR of 128-query tiles (256 queries, two TMEM accumulators): every 32-channel block of the weights
// (32 KB through the L2 per load) then serves 256 queries -- with single tiles the weight re-reads were 60 % of
// this kernel's L2->SM traffic and the MMA thread sat on the TMA barrier (ncu: tensor pipe 19 % active).
__global__ void __launch_bounds__(cv::THREADS, 1)
conv1x1_mn_kernel(const __grid_constant__ CUtensorMap mapX, const __grid_constant__ CUtensorMap mapW, const cv::Params p) {
  using namespace cv;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t *sm = smem_raw + (base - raw);
  uint64_t *bars = reinterpret_cast<uint64_t *>(sm + NST * STAGE_BYTES);
  uint64_t *full_bar = bars, *empty_bar = bars + NST, *tfull = bars + 2 * NST;
  uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(tfull + 1);
  __shared__ float bias_s[256];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < 256; i += THREADS) bias_s[i] = (p.bias && i < p.N) ? p.bias[i] : 0.f;
  if (threadIdx.x == 0) {
    tma_prefetch_desc(&mapX);
    tma_prefetch_desc(&mapW);
    for (int s = 0; s < NST; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    mbar_init(tfull, 1);
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const int pr = blockIdx.x;                        // one pair per CTA
  const int b = pr / p.pairs_per_b, m0 = (pr - b * p.pairs_per_b) * (2 * BM);

  if (warp == 0) {
    if (lane == 0) {
      for (int kb = 0; kb < p.kblocks; ++kb) {
        const int st = kb % NST;
        mbar_wait(&empty_bar[st], ((kb / NST) & 1) ^ 1);
        uint8_t *a = sm + st * STAGE_BYTES;
        mbar_arrive_expect_tx(&full_bar[st], A_BYTES + (uint32_t)p.NP * BK * 4);
#pragma unroll
        for (int j = 0; j < 8; ++j) tma_load_3d(a + j * 4096, &mapX, &full_bar[st], m0 + 32 * j, kb * BK, b);
        tma_load_3d(a + A_BYTES, &mapW, &full_bar[st], kb * BK, 0, 0);
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      const uint32_t idesc = umma_idesc_tf32(BM, p.NP) | (1u << 15);      // A MN-major
      for (int kb = 0; kb < p.kblocks; ++kb) {
        const int st = kb % NST;
        mbar_wait(&full_bar[st], (kb / NST) & 1);
        tc_fence_after();
        const uint32_t a_addr = base + st * STAGE_BYTES;
        const uint64_t bdesc = umma_desc_kmajor_sw128(a_addr + A_BYTES);
#pragma unroll
        for (int h = 0; h < 2; ++h)
#pragma unroll
          for (int k4 = 0; k4 < 4; ++k4)
            umma_tf32(tmem_base + h * 256, desc_mnmajor_sw128_32b(a_addr + h * 16384 + k4 * 1024, 4096, 512),
                      bdesc + 2 * k4, idesc, (kb | k4) != 0);
        umma_commit(&empty_bar[st]);
      }
      umma_commit(tfull);
    }
  } else {
    // epilogue: warps 2..9; warp w reads TMEM lane quarter w % 4 of accumulator (w - 2) / 4
    const int wq = warp & 3, h = (warp - 2) >> 2;
    const int q = m0 + h * BM + wq * 32 + lane;
    const bool ok = q < p.HW;
    float *ob = p.out + ((size_t)b * p.N) * p.HW + q;
    mbar_wait(tfull, 0);
    tc_fence_after();
    for (int c0 = 0; c0 < p.N; c0 += 32) {
      float v[32];
      tmem_ld_32x32(tmem_base + h * 256 + c0 + ((uint32_t)(wq * 32) << 16), v);
      tmem_ld_wait();
#pragma unroll
      for (int i = 0; i < 32; ++i) {
        if (c0 + i < p.N) {
          float x = v[i] + bias_s[c0 + i];
          if (p.relu) x = fmaxf(x, 0.f);
          if (ok) ob[(size_t)(c0 + i) * p.HW] = x;
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, 512);
}

int taps_padded(int radius) { return (((2 * radius + 1) * (2 * radius + 1)) + 31) / 32 * 32; }

}  // namespace

extern "C" {

size_t b200corr_lookup_convc1_weight_bytes(int num_levels, int radius, int n_out) {
  if (num_levels < 1 || radius < 1 || radius > 4 || n_out < 1) return 0;
  return (size_t)num_levels * n_out * taps_padded(radius) * sizeof(float);
}

int b200corr_lookup_convc1_prepare(const float *weight, float *wprep, int num_levels, int radius, int n_out,
                                   void *stream) {
  B200_CHECK(weight && wprep, "lookup_convc1_prepare: null pointer");
  B200_CHECK(num_levels >= 1 && num_levels <= kMaxLevels && radius >= 1 && radius <= 4 && n_out >= 1,
             "lookup_convc1_prepare: bad sizes");
  const int NT = (2 * radius + 1) * (2 * radius + 1), KP = taps_padded(radius);
  const int total = num_levels * n_out * KP;
  convc1_prep_kernel<<<(total + 255) / 256, 256, 0, (cudaStream_t)stream>>>(weight, wprep, num_levels, n_out, NT, KP);
  B200_LAUNCH_OK("convc1_prep_kernel");
  return 0;
}

int b200corr_lookup_convc1_forward(const float *const *h_levels, int num_levels, int blocked_levels,
                                   const float *coords, const float *wprep, const float *bias, float *out, int B,
                                   int H, int W, int radius, int mode, int n_out, int relu, void *stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  if (B == 0) return 0;
  fz::Params p;
  if (int e = fill_params(p.lp, h_levels, nullptr, num_levels, B, H, W, radius, mode, "lookup_convc1_forward")) return e;
  B200_CHECK(coords && wprep && out, "lookup_convc1_forward: null pointer");
  B200_CHECK(n_out % 32 == 0 && n_out >= 32 && n_out <= 256 && (n_out <= 128 || n_out % 2 == 0),
             "lookup_convc1_forward: n_out = %d not supported (multiples of 32 up to 256)", n_out);
  B200_CHECK(((uintptr_t)wprep & 15) == 0, "lookup_convc1_forward: prepared weights must be 16-byte aligned");
  for (int l = 0; l < num_levels; ++l) {
    const uintptr_t a = (uintptr_t)p.lp.lvl[l];
    p.lp.path[l] = (p.lp.LW[l] % 8 == 0 && a % 32 == 0) ? PATH_SECTOR : (p.lp.LW[l] % 4 == 0 && a % 16 == 0) ? PATH_VEC4 : PATH_SCALAR;
    if ((blocked_levels >> l) & 1) {
      B200_CHECK(l <= 1 && W % 8 == 0 && a % 32 == 0, "lookup_convc1_forward: level %d cannot be in the blocked layout", l);
      int hp, wp;
      b200corr_blocked_level_dims(l, H, W, &hp, &wp);
      p.lp.blocked[l] = 1; p.lp.path[l] = PATH_SECTOR; p.lp.tiles_w[l] = wp / 8; p.lp.slice[l] = (long long)hp * wp;
    }
  }
  B200_CHECK((blocked_levels >> num_levels) == 0, "lookup_convc1_forward: blocked_levels names a level that is not there");
  p.bias = bias; p.out = out; p.n_out = n_out; p.relu = relu;
  p.nh = n_out > 128 ? 2 : 1;
  p.rows_h = n_out / p.nh;
  B200_CHECK(p.rows_h % 16 == 0, "lookup_convc1_forward: n_out / %d must be a multiple of 16", p.nh);
  p.tiles_per_b = (H * W + fz::TILE_Q - 1) / fz::TILE_Q;
  p.total_tiles = p.tiles_per_b * B;
  const int KP = taps_padded(radius);
  CUtensorMap mapW;
  const uint64_t dims[3] = {(uint64_t)KP, (uint64_t)n_out, (uint64_t)num_levels};
  const uint64_t strides[3] = {4, (uint64_t)KP * 4, (uint64_t)KP * 4 * n_out};
  const uint32_t box[3] = {32, (uint32_t)p.rows_h, 1};
  if (int e = b200::make_tensor_map(&mapW, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, wprep, dims, strides, box,
                                    CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B))
    return e;
  const int grid = p.total_tiles < b200::num_sms() ? p.total_tiles : b200::num_sms();
#define LC1_LAUNCH(RR)                                                                                         \
  {                                                                                                            \
    static bool done[64];                                                                                      \
    const int smem = fz::G<RR>::SMEM_BYTES;                                                                    \
    if (int e = b200::set_max_smem_once((const void *)lookup_convc1_kernel<RR>, smem, done)) return e;         \
    lookup_convc1_kernel<RR><<<grid, fz::THREADS, smem, stream>>>(mapW, p, coords);                            \
  }
  switch (radius) {
    case 1: LC1_LAUNCH(1) break;
    case 2: LC1_LAUNCH(2) break;
    case 3: LC1_LAUNCH(3) break;
    default: LC1_LAUNCH(4) break;
  }
#undef LC1_LAUNCH
  B200_LAUNCH_OK("lookup_convc1_kernel");
  return 0;
}

int b200corr_conv1x1_forward(const float *x, const float *weight, const float *bias, float *out, int B, int K,
                             int N, int HW, int relu, void *stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  if (B == 0) return 0;
  B200_CHECK(x && weight && out, "conv1x1_forward: null pointer");
  B200_CHECK(B > 0 && K >= 1 && N >= 1 && N <= 256 && HW >= 1, "conv1x1_forward: bad sizes (n_out <= 256)");
  B200_CHECK(K % 4 == 0 && HW % 4 == 0, "conv1x1_forward: K and H*W must be multiples of 4 (16-byte TMA pitches)");
  B200_CHECK((((uintptr_t)x | (uintptr_t)weight) & 15) == 0, "conv1x1_forward: operands must be 16-byte aligned");
  cv::Params p;
  p.B = B; p.K = K; p.N = N; p.NP = (N + 15) / 16 * 16; p.HW = HW; p.relu = relu; p.bias = bias; p.out = out;
  p.pairs_per_b = (HW + 2 * cv::BM - 1) / (2 * cv::BM);
  p.total_pairs = p.pairs_per_b * B;
  p.kblocks = (K + cv::BK - 1) / cv::BK;
  CUtensorMap mapX, mapW;
  {
    const uint64_t d[3] = {(uint64_t)HW, (uint64_t)K, (uint64_t)B};
    const uint64_t s[3] = {4, (uint64_t)HW * 4, (uint64_t)HW * 4 * K};
    const uint32_t box[3] = {32, 32, 1};
    if (int e = b200::make_tensor_map(&mapX, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, x, d, s, box,
                                      CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B)) return e;
  }
  {
    const uint64_t d[3] = {(uint64_t)K, (uint64_t)N, 1};
    const uint64_t s[3] = {4, (uint64_t)K * 4, (uint64_t)K * 4 * N};
    const uint32_t box[3] = {32, (uint32_t)p.NP, 1};
    if (int e = b200::make_tensor_map(&mapW, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, weight, d, s, box,
                                      CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B)) return e;
  }
  static bool done[64];
  if (int e = b200::set_max_smem_once((const void *)conv1x1_mn_kernel, cv::SMEM_BYTES, done)) return e;
  conv1x1_mn_kernel<<<p.total_pairs, cv::THREADS, cv::SMEM_BYTES, stream>>>(mapX, mapW, p);
  B200_LAUNCH_OK("conv1x1_mn_kernel");
  return 0;
}

}  // extern "C"
