// raft_allpairs.cu -- RAFT all-pairs correlation volume + average-pool pyramid.
//
// Replaces CorrBlock.__init__ / CorrBlock.corr of the reference (models/raft/corr.py:55-64,98-106):
//     vol0[b, p1, p2] = (1/sqrt(C)) * sum_c f1[b,c,p1] * f2[b,c,p2]        (torch.matmul + divide)
//     vol_{l+1}       = avg_pool2d(vol_l, 2, stride 2) over p2 = (y, x)     (3 more passes)
// which costs one cuBLAS SGEMM plus four elementwise passes over a 236 MB/sample volume.
//
// Here (precision TF32, the default):
//   1. prep kernel: NCHW fp32 features -> [B][HW][Cp] (K-major) rounded to TF32 (cvt.rna), Cp = C
//      rounded up to 32, zero padded.  7.9 MB/sample per map -- noise next to the volume.
//   2. one persistent warp-specialised tcgen05 kernel run by CTA PAIRS (cluster of 2, cta_group::2).
//      Pair tile = 256 query pixels (M, 128 per CTA) x one 8x32 spatial patch of key pixels (N = 256),
//      K = Cp in blocks of 32 TF32 (128-byte swizzled rows).  Per CTA:
//        warp 0    : TMA producer -- its own 128 query rows (A: [Cp,HW,B] box 32x128) and HALF of the
//                    key patch (B: [Cp,W,H,B] box 32x32x4), 5-stage ring; the bytes of both CTAs are
//                    counted on the leader's full barrier
//        warp 1    : (leader CTA only) tcgen05.mma.cta_group::2.kind::tf32 issuer, M256 N256 K8, 4 per
//                    k-block; accumulators in each CTA's TMEM, double buffered (2 x 256 columns);
//                    commits are multicast to both CTAs' barriers
//        warp 2    : TMEM allocator
//        warps 4-11: epilogue.  All 8 warps drain ONE accumulator at a time: warps 4-7 take patch rows
//                    0-3, warps 8-11 patch rows 4-7 of the same 128 query rows; a thread owns one query
//                    row, one patch row (32 columns) at a time with tcgen05.ld, so the 2x2 and 4x4
//                    average pools are register-local and the 8x8 pool needs one 8-float hand-over
//                    between the halves.  Levels 0 and 1 go through a swizzled per-warp staging tile
//                    and leave as full 128/64-byte lines (st.global.cs); levels 2-3 are written
//                    straight from registers.
//      The volume is written exactly once and never re-read: 313 MB/sample instead of ~1.1 GB.
//      Measured (profiles/, DESIGN.md 2.3, scripts/ap_matrix.sh): MMAs alone 0.19 ms, + pooled
//      levels 0.25 ms, + level 0 0.33 ms at B=4 -- the store-only pattern of this kernel runs at
//      5.4 TB/s (scripts/probes/store_probe.cu), so the remaining gap is imperfect overlap: the
//      operand reads of the MMAs, the TMA fills and the staging round trip of the epilogue share the
//      shared-memory banks.  Tried and dropped: TMA stores for level 0 (32 scattered 128-byte rows per
//      box: slower), 256-bit stores straight from registers (32 sectors per instruction: 1.8 TB/s),
//      a cluster-scope release on the TMEM hand-over (puts a MEMBAR.GPU in front of every arrival).
//      B200CORR_ALLPAIRS_CTAS=1 selects the single-CTA variant of the same kernel.
//   Precision TF32X3: the same kernel run as a K-extended GEMM over the split operands x = hi + lo
//      (both TF32): passes lo(f1)*hi(f2), hi*lo, hi*hi accumulate into the same TMEM tile -- three times
//      the MMAs, fp32-level accuracy (the dropped lo*lo term is 2^-22 relative per product).
//   Precision FP32 (exact, also the path for W % 4 != 0): a plain SIMT tile GEMM + pooling kernels.
//
// Error bound of the TF32 path (documented in DESIGN.md, asserted in tests): inputs are rounded to
// 11 significant bits (rel. error <= 2^-11 each), products accumulate in fp32, so
//     |vol - exact| <= (2^-10 + 2^-22) * scale * sum_c |f1_c * f2_c|  (+ fp32 accumulation error).
#include <cstdlib>

#include <cuda_fp16.h>

#include "tcgen05.cuh"

namespace {
using namespace b200dev;

// ------------------------------------------------------------------------------ prep (transpose)
// in [B][C][HW] -> out [B][HW][Cp], tf32-rounded, channels >= C zero.  With out_lo (TF32x3): the
// residual x - tf32(x), itself rounded to TF32, so that x = hi + lo to ~2^-22 relative.
// One launch converts BOTH feature maps (blockIdx.z = map * B + b; the maps may differ in size): a CTA
// transposes 32 pixels x 128 channels, reads 128-byte runs of a channel plane and writes 512-byte runs of a
// pixel row (two launches of 32 x 32 tiles took 2 x 14.7 us at B=4, 4.3 TB/s; this one ~21 us).
struct PrepMap {
  const float *in;
  float *out, *out_lo;
  int HW;
};
__global__ void __launch_bounds__(256)
prep_kmajor_tf32_kernel(PrepMap m0, PrepMap m1, int B, int C, int Cp) {
  // [pixel][channel ^ pixel]: the transposing writes (lane = pixel) and the 16-byte reads (lane = channel quad)
  // are both bank-conflict free; the XOR permutes a quad's elements by pixel & 3, which is warp-uniform below
  __shared__ __align__(16) float tile[32][128];
  const int which = blockIdx.z / B, b = blockIdx.z - which * B;
  const PrepMap m = which ? m1 : m0;
  const int p0 = blockIdx.x * 32, c0 = blockIdx.y * 128;
  if (p0 >= m.HW) return;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 32 x 8
#pragma unroll
  for (int i = 0; i < 16; ++i) {
    const int cl = ty + 8 * i, c = c0 + cl, p = p0 + tx;
    float v = 0.f;
    if (c < C && p < m.HW) v = __ldg(m.in + ((size_t)b * C + c) * m.HW + p);
    tile[tx][cl ^ tx] = v;
  }
  __syncthreads();
  // thread -> (pixel ty + 8 i, channels c0 + 4 tx .. + 3): a warp writes 512 contiguous bytes of one pixel row
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int r = ty + 8 * i, p = p0 + r, c = c0 + 4 * tx;
    if (p < m.HW && c < Cp) {
      const float4 e = *reinterpret_cast<const float4 *>(&tile[r][(4 * tx) ^ (r & 28)]);
      // channel j of the quad sits at element j ^ (r & 3); r & 3 == ty & 3 for every i
      const bool s0 = ty & 1, s1 = ty & 2;
      const float a0 = s0 ? e.y : e.x, a1 = s0 ? e.x : e.y, a2 = s0 ? e.w : e.z, a3 = s0 ? e.z : e.w;
      float4 v, hi, lo;
      v.x = s1 ? a2 : a0; v.y = s1 ? a3 : a1; v.z = s1 ? a0 : a2; v.w = s1 ? a1 : a3;
      hi.x = to_tf32_rna(v.x); hi.y = to_tf32_rna(v.y); hi.z = to_tf32_rna(v.z); hi.w = to_tf32_rna(v.w);
      *reinterpret_cast<float4 *>(m.out + ((size_t)b * m.HW + p) * Cp + c) = hi;
      if (m.out_lo) {
        lo.x = to_tf32_rna(v.x - hi.x); lo.y = to_tf32_rna(v.y - hi.y);
        lo.z = to_tf32_rna(v.z - hi.z); lo.w = to_tf32_rna(v.w - hi.w);
        *reinterpret_cast<float4 *>(m.out_lo + ((size_t)b * m.HW + p) * Cp + c) = lo;
      }
    }
  }
}

// ------------------------------------------------------------------------------ tcgen05 kernel
namespace tc {
constexpr int BM = 128, BN = 256, BK = 32;
constexpr int PH = 8, PW = 32;  // key patch
constexpr int EPI_ROW_BYTES = 32 * 128;                 // one staging tile: 32 query rows x 32 floats
constexpr int EPI_WARPS = 8;
constexpr int THREADS = 384;
constexpr int TMEM_COLS = 512;
// CTAS = 1: one CTA per 128 x 256 tile.  CTAS = 2: a CTA pair (cluster of 2, tcgen05 cta_group::2)
// per 256 x 256 tile -- each CTA stages its 128 query rows and HALF of the key patch (4 of the 8
// patch rows), so a CTA pulls 256 KB instead of 384 KB of operands through the L2 per 128 x 256
// outputs, and the smaller stages allow a deeper ring.
template <int CTAS>
struct Cfg {
  static constexpr int NST = CTAS == 2 ? 5 : 3;
  static constexpr int EPI_WARP_BYTES = EPI_ROW_BYTES;   // per-warp staging tile: one patch row of 32 queries
  static constexpr int A_BYTES = BM * BK * 4, B_BYTES = (BN / CTAS) * BK * 4, STAGE_BYTES = A_BYTES + B_BYTES;
  // + level-2 rows handed from the upper-half to the lower-half epilogue warps (4 pairs x 2 slots x 1 KB)
  static constexpr int XCH_BYTES = 4 * 2 * 1024;
  static constexpr int SMEM_BYTES = NST * STAGE_BYTES + EPI_WARPS * EPI_WARP_BYTES + XCH_BYTES + 1024 /*align*/ + 512 /*barriers*/;
};

struct Params {
  int B, HW, H, W, KB;       // HW = query pixels (rows of the volume), H x W = key map; KB = Cp / 32
  int passes;                // 1: TF32.  3: TF32x3 -- the k loop runs over lo*hi, hi*lo, hi*hi (K-extended GEMM)
  int MT, NTY, NTX;          // tile counts (MT counts 128*CTAS-row blocks)
  float scale;
  int debug;                 // B200CORR_DEBUG bits (diagnostics): 1 skip level-0 stores, 2 skip pooled stores
  int TW0, TW1;              // BLK kernels: 8x8 tiles per row of a (padded) level-0 / level-1 slice
  long long S0, S1;          // BLK kernels: floats per (padded) level-0 / level-1 slice
  float *lvl0;               // level 0
  float *lvl[3];             // levels 1..3 (nullptr if not requested)
  int LH[3], LW[3];
};
}  // namespace tc

// two floats -> packed fp16 pair (lo in the low half), round to nearest, finite values saturate at +-65504
__device__ __forceinline__ uint32_t pack_half2_sat(float lo, float hi) {
  uint32_t d;
  asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(hi), "f"(lo));
  return d;
}

// N consecutive floats starting at column x of a row of width WL
template <int N>
__device__ __forceinline__ void store_row_vec(float *dst, const float (&v)[N], int x, int WL,
                                              bool vec_ok) {
  if (vec_ok) {
#pragma unroll
    for (int j = 0; j < N; j += 4)
      if (x + j < WL) *reinterpret_cast<float4 *>(dst + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
  } else {
#pragma unroll
    for (int j = 0; j < N; ++j)
      if (x + j < WL) dst[j] = v[j];
  }
}

// Pooled level whose width is not a multiple of 4 (no 16-byte stores): every lane holds NV consecutive
// values of ITS query row; scalar stores straight from registers would be 32 scattered 4-byte writes per
// instruction.  Transpose through the warp's staging tile instead, so that NV consecutive lanes write NV
// consecutive floats of one row (whole sectors).  `stage` holds 32 x (NV + 1) floats.
template <int NV>
__device__ __forceinline__ void store_rows_coalesced(float *stage, const float (&v)[NV], float *level, size_t qbase,
                                                     int mrow0, int HW, int row_elems, int y, int x, int WL, int lane) {
  __syncwarp();
#pragma unroll
  for (int j = 0; j < NV; ++j) stage[lane * (NV + 1) + j] = v[j];
  __syncwarp();
#pragma unroll
  for (int it = 0; it < NV; ++it) {
    const int idx = it * 32 + lane, qr = idx / NV, e = idx % NV;
    if (mrow0 + qr < HW && x + e < WL)
      level[((qbase + qr) * row_elems) + (size_t)y * WL + x + e] = stage[qr * (NV + 1) + e];
  }
}

// BLK (CTA pairs only): levels 0 and 1 are written in the blocked layout -- a slice is a row-major grid of 8x8
// tiles, each tile 64 consecutive floats (256 B).  The key patch is staged as four 8-column sub-boxes so that
// the accumulator columns of one epilogue step (32 of them) are one half tile (4 rows x 8 columns): still one
// 128-byte line per query row and store, but the lines of a query's patch are 1 KB contiguous, and the lookup's
// 10-row windows touch 4-6 tiles of 256 B instead of 10 rows 640 B apart.
// HALF (BLK only): the two blocked levels are stored as fp16 (8x8 tiles of 64 halves = 128 B): half the volume bytes
// for the build to write and the lookups to read.  The values are rounded once, from the fp32 accumulator / the fp32
// 2x2 mean (cvt.rn.satfinite: relative 2^-11, the size of the TF32 input rounding; |v| > 65504 saturates); levels 2-3
// stay fp32.  Staging rows shrink to 64 B (level 0) and 32 B (level 1): half the shared-memory round trip as well.
template <int CTAS, bool BLK = false, bool HALF = false>
__global__ void __launch_bounds__(tc::THREADS, 1)
allpairs_tc_kernel(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapB,
                   const __grid_constant__ CUtensorMap mapAlo, const __grid_constant__ CUtensorMap mapBlo,
                   const tc::Params p) {
  using namespace tc;
  using C = Cfg<CTAS>;
  constexpr int NST = C::NST, A_BYTES = C::A_BYTES, STAGE_BYTES = C::STAGE_BYTES, EPI_WARP_BYTES = C::EPI_WARP_BYTES;
  extern __shared__ uint8_t smem_raw[];
  // 1024-byte alignment for the 128-byte swizzle atoms
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t *sm = smem_raw + (base - raw);
  uint8_t *epi = sm + NST * STAGE_BYTES;
  float *xch = reinterpret_cast<float *>(epi + EPI_WARPS * EPI_WARP_BYTES);
  uint64_t *bars = reinterpret_cast<uint64_t *>(epi + EPI_WARPS * EPI_WARP_BYTES + C::XCH_BYTES);
  uint64_t *full_bar = bars, *empty_bar = bars + NST;
  uint64_t *tfull_bar = bars + 2 * NST, *tempty_bar = bars + 2 * NST + 2;
  uint64_t *xfull_bar = bars + 2 * NST + 4, *xempty_bar = bars + 2 * NST + 12;   // [pair][slot]
  uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(bars + 2 * NST + 20);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // CTA pair: rank 0 (the leader) owns the full / TMEM-empty barriers and issues the MMAs
  const uint32_t rank = CTAS == 2 ? cluster_ctarank() : 0;
  const int unit = blockIdx.x / CTAS, nunits = gridDim.x / CTAS;   // tile scheduler granularity

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&mapA);
    tma_prefetch_desc(&mapB);
    if (p.passes == 3) {
      tma_prefetch_desc(&mapAlo);
      tma_prefetch_desc(&mapBlo);
    }
    for (int s = 0; s < NST; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&tfull_bar[s], 1);
      mbar_init(&tempty_bar[s], EPI_WARPS * CTAS);   // every epilogue warp of every CTA reads every buffer
    }
    for (int s = 0; s < 8; ++s) {
      mbar_init(&xfull_bar[s], 1);
      mbar_init(&xempty_bar[s], 1);
    }
    fence_barrier_init();
  }
  if (warp == 2) {
    if (CTAS == 2) { tmem_alloc_2cta(tmem_slot, TMEM_COLS); tmem_relinquish_2cta(); }
    else { tmem_alloc(tmem_slot, TMEM_COLS); tmem_relinquish(); }
  }
  tc_fence_before();
  __syncthreads();
  if (CTAS == 2) cluster_sync_all();   // the peer's barriers exist before anything signals them
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int NT = p.NTY * p.NTX;
  const int total = p.B * p.MT * NT;

  if (warp == 0) {
    // ================= TMA producer (every CTA: its own A rows and its share of the key patch)
    if (lane == 0) {
      uint32_t it = 0;
      uint32_t full_cluster[NST];   // CTAS == 2: the leader's full barriers, as cluster addresses
      if (CTAS == 2)
        for (int s = 0; s < NST; ++s) full_cluster[s] = map_to_cta(&full_bar[s], 0);
      for (int t = unit; t < total; t += nunits) {
        const int nt = t % NT, mt = (t / NT) % p.MT, b = t / (NT * p.MT);
        const int y0 = (nt / p.NTX) * PH, x0 = (nt % p.NTX) * PW, m0 = (mt * CTAS + (int)rank) * BM;
        for (int kk = 0; kk < p.KB * p.passes; ++kk, ++it) {
          // TF32x3: pass 0 = lo(f1) * hi(f2), pass 1 = hi * lo, pass 2 = hi * hi (small terms first)
          const int pass = p.passes == 1 ? 2 : kk / p.KB, kb = kk - (p.passes == 1 ? 0 : pass * p.KB);
          const CUtensorMap *ma = pass == 0 ? &mapAlo : &mapA, *mb = pass == 1 ? &mapBlo : &mapB;
          const int st = it % NST;
          mbar_wait(&empty_bar[st], ((it / NST) & 1) ^ 1);
          uint8_t *a = sm + st * STAGE_BYTES;
          if (CTAS == 2) {
            // one arrival (the leader's) expects the bytes of both CTAs
            if (rank == 0) mbar_arrive_expect_tx(&full_bar[st], 2 * STAGE_BYTES);
            tma_load_3d_2cta(a, ma, full_cluster[st], kb * BK, m0, b);
            if constexpr (BLK) {
              // the key map seen as (K, x % 8, y, x / 8, b): one box lands in the order [tile column][row][column]
              tma_load_5d_2cta(a + A_BYTES, mb, full_cluster[st], kb * BK, 0, y0 + (int)rank * (PH / 2), x0 >> 3, b);
            } else {
              tma_load_4d_2cta(a + A_BYTES, mb, full_cluster[st], kb * BK, x0, y0 + (int)rank * (PH / 2), b);
            }
          } else {
            mbar_arrive_expect_tx(&full_bar[st], STAGE_BYTES);
            tma_load_3d(a, ma, &full_bar[st], kb * BK, m0, b);
            tma_load_4d(a + A_BYTES, mb, &full_bar[st], kb * BK, x0, y0, b);
          }
        }
      }
    }
  } else if (warp == 1) {
    // ================= MMA issuer (single thread; of the leader CTA for a pair)
    if (lane == 0 && rank == 0) {
      constexpr uint32_t idesc = umma_idesc_tf32(BM * CTAS, BN);
      uint32_t it = 0, lt = 0;
      for (int t = unit; t < total; t += nunits, ++lt) {
        const int buf = lt & 1;
        mbar_wait(&tempty_bar[buf], ((lt >> 1) & 1) ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + buf * BN;
        for (int kb = 0; kb < p.KB * p.passes; ++kb, ++it) {
          const int st = it % NST;
          mbar_wait(&full_bar[st], (it / NST) & 1);
          tc_fence_after();
          const uint32_t a_addr = base + st * STAGE_BYTES;
          const uint64_t adesc = umma_desc_kmajor_sw128(a_addr);
          const uint64_t bdesc = umma_desc_kmajor_sw128(a_addr + A_BYTES);
#pragma unroll
          for (int k4 = 0; k4 < 4; ++k4) {  // 8 tf32 = 32 bytes per MMA: +2 in the 16-byte address field
            if (CTAS == 2) umma_tf32_2cta(d_tmem, adesc + 2 * k4, bdesc + 2 * k4, idesc, (kb | k4) != 0);
            else umma_tf32(d_tmem, adesc + 2 * k4, bdesc + 2 * k4, idesc, (kb | k4) != 0);
          }
          if (CTAS == 2) umma_commit_2cta(&empty_bar[st]); else umma_commit(&empty_bar[st]);
        }
        if (CTAS == 2) umma_commit_2cta(&tfull_bar[buf]); else umma_commit(&tfull_bar[buf]);
      }
    }
  } else if (warp >= 4) {
    // ================= epilogue: 8 warps drain ONE accumulator at a time.  Warp w may only touch TMEM
    // lanes 32*(w%4) .. +31, so warps 4-7 (half 0) take patch rows 0-3 and warps 8-11 (half 1) patch
    // rows 4-7 of the same 128 query rows.  A tile's stores then run at the SM's full store rate and
    // the buffer goes back to the MMA warp after half the rows; the MMAs of the next tile run in the
    // other buffer meanwhile.  (Two groups on alternate tiles each got half the store bandwidth and
    // the MMA warp waited for TMEM: per-tile time (drain + mma) / 2 instead of max(drain, mma).)
    // A thread owns one query row and 4 x 32 keys, one patch row (32 columns) at a time with
    // tcgen05.ld, so the 2x2 and 4x4 average pools are register-local; the 8x8 pool needs the level-2
    // row of the other half, handed from half 0 to half 1 through shared memory.
    const int wq = warp & 3, half = (warp - 4) >> 2;
    uint8_t *ebuf = epi + (warp - 4) * EPI_WARP_BYTES;
    const bool m_dbg = !(p.debug & 2);
    const bool v1 = p.lvl[0] && (p.LW[0] % 4 == 0), v2 = p.lvl[1] && (p.LW[1] % 4 == 0),
               v3 = p.lvl[2] && (p.LW[2] % 4 == 0);
    const uint32_t tempty_leader[2] = {CTAS == 2 ? map_to_cta(&tempty_bar[0], 0) : 0u,
                                       CTAS == 2 ? map_to_cta(&tempty_bar[1], 0) : 0u};
    for (uint32_t lt = 0; (int)(unit + lt * nunits) < total; ++lt) {
      const int t = unit + lt * nunits;
      const int buf = lt & 1;
      const int nt = t % NT, mt = (t / NT) % p.MT, b = t / (NT * p.MT);
      const int y0 = (nt / p.NTX) * PH, x0 = (nt % p.NTX) * PW;
      int mrow0 = (mt * CTAS + (int)rank) * BM + wq * 32;
      int m = mrow0 + lane;
      const bool m_ok = m < p.HW && m_dbg;
      size_t q = (size_t)b * p.HW + (m_ok ? m : 0);
      if (p.debug & 4) {   // diagnostics: every tile stores into the first 128 query rows (L2-resident)
        mrow0 = wq * 32;
        m = mrow0 + lane;
        q = m;
      }
      const int bst = (p.debug & 4) ? 0 : b;
      mbar_wait(&tfull_bar[buf], (lt >> 1) & 1);
      tc_fence_after();
      const uint32_t taddr = tmem_base + buf * BN + half * (PH / 2) * 32 + ((uint32_t)(wq * 32) << 16);

      // global pointers of this tile (coalesced write-out: lane -> (row, 16-byte chunk))
      float *vol0 = p.lvl0;
      const int wrow = lane >> 3, wchunk = lane & 7;      // level 0: 4 rows x 8 chunks per instruction
      const int xrow = lane >> 2, xchunk = lane & 3;      // level 1: 8 rows x 4 chunks per instruction
      float prev[32], p1prev[16], p2[8];
#pragma unroll
      for (int r = 0; r < PH / 2; ++r) {
        const int pr = half * (PH / 2) + r;   // patch row
        float cur[32];
        tmem_ld_32x32(taddr + r * 32, cur);
        tmem_ld_wait();
        if (r == PH / 2 - 1) {
          // last TMEM read of this accumulator by this warp: hand it back before the math / stores
          tc_fence_before();
          __syncwarp();
          if (lane == 0) {
            if (CTAS == 2) mbar_arrive_cluster(tempty_leader[buf]); else mbar_arrive(&tempty_bar[buf]);
          }
        }
#pragma unroll
        for (int i = 0; i < 32; ++i) cur[i] *= p.scale;
        // ---- level 0: one patch row of 32 query rows -> swizzled staging -> full-line streaming stores
        __syncwarp();   // previous readers of the staging tile are done
        if constexpr (HALF) {
          // half a tile = 32 halves = 64 B per query row: staged like level 1 of the fp32 kernel
          uint32_t h[16];
#pragma unroll
          for (int i = 0; i < 16; ++i) h[i] = pack_half2_sat(cur[2 * i], cur[2 * i + 1]);
#pragma unroll
          for (int j = 0; j < 4; ++j)
            *reinterpret_cast<uint4 *>(ebuf + lane * 64 + ((j ^ ((lane >> 1) & 3)) << 4)) =
                make_uint4(h[4 * j], h[4 * j + 1], h[4 * j + 2], h[4 * j + 3]);
          __syncwarp();
          if (!(p.debug & 1)) {
            const size_t toff = ((size_t)(y0 >> 3) * p.TW0 + (x0 >> 3) + r) * 64 + half * 32 + 8 * xchunk;   // halves
            __half *vol0h = reinterpret_cast<__half *>(vol0);
#pragma unroll
            for (int it = 0; it < 4; ++it) {
              const int row = it * 8 + xrow;
              const uint4 v = *reinterpret_cast<const uint4 *>(ebuf + row * 64 + ((xchunk ^ ((row >> 1) & 3)) << 4));
              const int mm = mrow0 + row;
              if (mm < p.HW && x0 + 8 * r < p.W)
                __stcs(reinterpret_cast<uint4 *>(vol0h + ((size_t)bst * p.HW + mm) * (size_t)p.S0 + toff), v);
            }
          }
        } else {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const int off = lane * 128 + ((j ^ (lane & 7)) << 4);
          *reinterpret_cast<float4 *>(ebuf + off) = make_float4(cur[4 * j], cur[4 * j + 1], cur[4 * j + 2], cur[4 * j + 3]);
        }
        __syncwarp();
        }
        if (!HALF && !(p.debug & 1)) {
          if constexpr (BLK) {
            // cur = tile column r of this half: rows 4*half..4*half+3 x columns x0+8r..x0+8r+7 = half a tile
            const size_t toff = ((size_t)(y0 >> 3) * p.TW0 + (x0 >> 3) + r) * 64 + half * 32 + 4 * wchunk;
            const size_t slice = (size_t)p.S0;
#pragma unroll
            for (int it = 0; it < 8; ++it) {
              const int row = it * 4 + wrow;
              const float4 v = *reinterpret_cast<const float4 *>(ebuf + row * 128 + ((wchunk ^ (row & 7)) << 4));
              const int mm = mrow0 + row;
              if (mm < p.HW && x0 + 8 * r < p.W)
                __stcs(reinterpret_cast<float4 *>(vol0 + ((size_t)bst * p.HW + mm) * slice + toff), v);
            }
          } else {
          const int y = y0 + pr, x = x0 + 4 * wchunk;
#pragma unroll
          for (int it = 0; it < 8; ++it) {
            const int row = it * 4 + wrow;
            const float4 v = *reinterpret_cast<const float4 *>(ebuf + row * 128 + ((wchunk ^ (row & 7)) << 4));
            const int mm = mrow0 + row;
            if (mm < p.HW && y < p.H && x < p.W)
              __stcs(reinterpret_cast<float4 *>(vol0 + (((size_t)bst * p.HW + mm) * p.H + y) * p.W + x), v);
          }
          }
        }
        if constexpr (BLK) {
          // 2x2 means inside the half tile: [2 rows][4 columns]
          float p1h[8];
#pragma unroll
          for (int yy = 0; yy < 2; ++yy)
#pragma unroll
            for (int xx = 0; xx < 4; ++xx)
              p1h[yy * 4 + xx] = (((cur[(2 * yy) * 8 + 2 * xx] + cur[(2 * yy) * 8 + 2 * xx + 1]) + cur[(2 * yy + 1) * 8 + 2 * xx]) +
                                  cur[(2 * yy + 1) * 8 + 2 * xx + 1]) * 0.25f;
          if (r & 1) {
            // tile columns (r-1, r) -> level 1: [2 rows][8 columns] = 64 contiguous bytes of a level-1 tile
            float p1[16];
#pragma unroll
            for (int yy = 0; yy < 2; ++yy)
#pragma unroll
              for (int xx = 0; xx < 4; ++xx) {
                p1[yy * 8 + xx] = prev[yy * 4 + xx];
                p1[yy * 8 + 4 + xx] = p1h[yy * 4 + xx];
              }
            if (HALF && p.lvl[0] && m_dbg) {
              // [2 rows][8 columns] = 16 halves = one 32-byte sector of a level-1 tile per query row
              const int y1 = y0 / 2 + 2 * half;
              const size_t toff = ((size_t)(y1 >> 3) * p.TW1 + (x0 >> 4) + (r >> 1)) * 64 + (y1 & 7) * 8;   // halves
              uint32_t h1[8];
#pragma unroll
              for (int i = 0; i < 8; ++i) h1[i] = pack_half2_sat(p1[2 * i], p1[2 * i + 1]);
              __syncwarp();
#pragma unroll
              for (int j = 0; j < 2; ++j)
                *reinterpret_cast<uint4 *>(ebuf + lane * 32 + ((j ^ ((lane >> 2) & 1)) << 4)) =
                    make_uint4(h1[4 * j], h1[4 * j + 1], h1[4 * j + 2], h1[4 * j + 3]);
              __syncwarp();
              __half *l1h = reinterpret_cast<__half *>(p.lvl[0]);
#pragma unroll
              for (int it = 0; it < 2; ++it) {
                const int row = it * 16 + (lane >> 1), ch = lane & 1;
                const uint4 v = *reinterpret_cast<const uint4 *>(ebuf + row * 32 + ((ch ^ ((row >> 2) & 1)) << 4));
                const int mm = mrow0 + row;
                if (mm < p.HW && x0 / 2 + 8 * (r >> 1) < p.LW[0])
                  __stcs(reinterpret_cast<uint4 *>(l1h + ((size_t)bst * p.HW + mm) * (size_t)p.S1 + toff + 8 * ch), v);
              }
            }
            if (!HALF && p.lvl[0] && m_dbg) {
              const int y1 = y0 / 2 + 2 * half;                    // first of the two level-1 rows
              const size_t toff = ((size_t)(y1 >> 3) * p.TW1 + (x0 >> 4) + (r >> 1)) * 64 + (y1 & 7) * 8 + 4 * xchunk;
              const size_t slice = (size_t)p.S1;
              __syncwarp();
#pragma unroll
              for (int j = 0; j < 4; ++j)
                *reinterpret_cast<float4 *>(ebuf + lane * 64 + ((j ^ ((lane >> 1) & 3)) << 4)) =
                    make_float4(p1[4 * j], p1[4 * j + 1], p1[4 * j + 2], p1[4 * j + 3]);
              __syncwarp();
#pragma unroll
              for (int it = 0; it < 4; ++it) {
                const int row = it * 8 + xrow;
                const float4 v = *reinterpret_cast<const float4 *>(ebuf + row * 64 + ((xchunk ^ ((row >> 1) & 3)) << 4));
                const int mm = mrow0 + row;
                if (mm < p.HW && x0 / 2 + 8 * (r >> 1) < p.LW[0])
                  __stcs(reinterpret_cast<float4 *>(p.lvl[0] + ((size_t)bst * p.HW + mm) * slice + toff), v);
              }
            }
            // level 2 (means of the two level-1 rows): 4 of this half's 8 values per tile-column pair
#pragma unroll
            for (int j = 0; j < 4; ++j)
              p2[(r >> 1) * 4 + j] = (((p1[2 * j] + p1[2 * j + 1]) + p1[8 + 2 * j]) + p1[8 + 2 * j + 1]) * 0.25f;
          } else {
#pragma unroll
            for (int i = 0; i < 8; ++i) prev[i] = p1h[i];
          }
        }
        if (BLK ? false : (r & 1)) {
          // ---- level 1: 2x2 means of rows (r-1, r) -> 16 values per query row
          float p1[16];
#pragma unroll
          for (int j = 0; j < 16; ++j)
            p1[j] = (((prev[2 * j] + prev[2 * j + 1]) + cur[2 * j]) + cur[2 * j + 1]) * 0.25f;
          if (p.lvl[0] && m_dbg) {
            const int y1 = y0 / 2 + (pr >> 1), x1 = x0 / 2;
            if (v1) {
              // staged like level 0: [32 rows][64 B], chunk swizzled by (row >> 1) & 3
              __syncwarp();
#pragma unroll
              for (int j = 0; j < 4; ++j)
                *reinterpret_cast<float4 *>(ebuf + lane * 64 + ((j ^ ((lane >> 1) & 3)) << 4)) =
                    make_float4(p1[4 * j], p1[4 * j + 1], p1[4 * j + 2], p1[4 * j + 3]);
              __syncwarp();
#pragma unroll
              for (int it = 0; it < 4; ++it) {
                const int row = it * 8 + xrow;
                const float4 v = *reinterpret_cast<const float4 *>(ebuf + row * 64 + ((xchunk ^ ((row >> 1) & 3)) << 4));
                const int mm = mrow0 + row, xx = x1 + 4 * xchunk;
                if (mm < p.HW && y1 < p.LH[0] && xx < p.LW[0])
                  __stcs(reinterpret_cast<float4 *>(p.lvl[0] + (((size_t)bst * p.HW + mm) * p.LH[0] + y1) * p.LW[0] + xx), v);
              }
            } else if (y1 < p.LH[0]) {
              store_rows_coalesced<16>(reinterpret_cast<float *>(ebuf), p1, p.lvl[0], (size_t)bst * p.HW + mrow0, mrow0,
                                       p.HW, p.LH[0] * p.LW[0], y1, x1, p.LW[0], lane);
            }
          }
          if (r == PH / 2 - 1) {
            // ---- level 2: means of two level-1 rows -> 8 values
#pragma unroll
            for (int j = 0; j < 8; ++j)
              p2[j] = (((p1prev[2 * j] + p1prev[2 * j + 1]) + p1[2 * j]) + p1[2 * j + 1]) * 0.25f;
          } else {
#pragma unroll
            for (int j = 0; j < 16; ++j) p1prev[j] = p1[j];
          }
        } else if (!BLK) {
#pragma unroll
          for (int i = 0; i < 32; ++i) prev[i] = cur[i];
        }
        {
          if (r == PH / 2 - 1) {
            if (p.lvl[1] && m_dbg) {
              const int y2 = y0 / 4 + half, x2 = x0 / 4;
              if (y2 < p.LH[1]) {
                if (v2) {
                  if (m_ok) store_row_vec(p.lvl[1] + (q * p.LH[1] + y2) * p.LW[1] + x2, p2, x2, p.LW[1], true);
                } else {
                  store_rows_coalesced<8>(reinterpret_cast<float *>(ebuf), p2, p.lvl[1], (size_t)bst * p.HW + mrow0, mrow0,
                                          p.HW, p.LH[1] * p.LW[1], y2, x2, p.LW[1], lane);
                }
              }
            }
            if (p.lvl[2]) {
              // ---- level 3 = mean of the two level-2 rows of this patch: half 0 hands its row over
              const int slot = wq * 2 + buf;
              float *xs = xch + slot * 256 + lane;   // [value][lane]
              if (half == 0) {
                mbar_wait(&xempty_bar[slot], ((lt >> 1) & 1) ^ 1);
#pragma unroll
                for (int j = 0; j < 8; ++j) xs[j * 32] = p2[j];
                __syncwarp();
                if (lane == 0) mbar_arrive(&xfull_bar[slot]);
              } else {
                mbar_wait(&xfull_bar[slot], (lt >> 1) & 1);
                float p2prev[8];
#pragma unroll
                for (int j = 0; j < 8; ++j) p2prev[j] = xs[j * 32];
                __syncwarp();
                if (lane == 0) mbar_arrive(&xempty_bar[slot]);
                float p3[4];
#pragma unroll
                for (int j = 0; j < 4; ++j)
                  p3[j] = (((p2prev[2 * j] + p2prev[2 * j + 1]) + p2[2 * j]) + p2[2 * j + 1]) * 0.25f;
                const int y3 = y0 / 8, x3 = x0 / 8;
                if (y3 < p.LH[2] && m_dbg) {
                  if (v3) {
                    if (m_ok) store_row_vec(p.lvl[2] + (q * p.LH[2] + y3) * p.LW[2] + x3, p3, x3, p.LW[2], true);
                  } else {
                    store_rows_coalesced<4>(reinterpret_cast<float *>(ebuf), p3, p.lvl[2], (size_t)bst * p.HW + mrow0, mrow0,
                                            p.HW, p.LH[2] * p.LW[2], y3, x3, p.LW[2], lane);
                  }
                }
              }
            }
          }
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (CTAS == 2) cluster_sync_all();   // nobody leaves while the peer can still signal its barriers
  if (warp == 2) {
    tc_fence_after();
    if (CTAS == 2) tmem_dealloc_2cta(tmem_base, TMEM_COLS); else tmem_dealloc(tmem_base, TMEM_COLS);
  }
}

// ------------------------------------------------------------------------------ exact fp32 path
// vol0[b, m, n] = scale * sum_c f1[b,c,m] * f2[b,c,n]; 64x64 tile, 256 threads, 4x4 per thread
__global__ void __launch_bounds__(256)
allpairs_simt_kernel(const float *__restrict__ f1, const float *__restrict__ f2, float *__restrict__ vol,
                     int C, int HW, float scale) {
  __shared__ float As[16][64], Bs[16][64];
  const int b = blockIdx.z, m0 = blockIdx.y * 64, n0 = blockIdx.x * 64;
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  float acc[4][4] = {};
  const float *a = f1 + (size_t)b * C * HW, *bb = f2 + (size_t)b * C * HW;
  for (int c0 = 0; c0 < C; c0 += 16) {
    for (int i = threadIdx.x; i < 16 * 64; i += 256) {
      const int k = i >> 6, j = i & 63;
      As[k][j] = (c0 + k < C && m0 + j < HW) ? a[(size_t)(c0 + k) * HW + m0 + j] : 0.f;
      Bs[k][j] = (c0 + k < C && n0 + j < HW) ? bb[(size_t)(c0 + k) * HW + n0 + j] : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < 16; ++k) {
      const float4 av = *reinterpret_cast<const float4 *>(&As[k][ty * 4]);
      const float4 bv = *reinterpret_cast<const float4 *>(&Bs[k][tx * 4]);
      const float ar[4] = {av.x, av.y, av.z, av.w}, br[4] = {bv.x, bv.y, bv.z, bv.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(ar[i], br[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int m = m0 + ty * 4 + i;
    if (m >= HW) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int n = n0 + tx * 4 + j;
      if (n < HW) vol[((size_t)b * HW + m) * HW + n] = acc[i][j] * scale;
    }
  }
}

// out[q, y, x] = mean of the 2x2 block of in[q, 2y.., 2x..] (floor mode), q = flattened leading dims
__global__ void __launch_bounds__(256)
avgpool2_kernel(const float *__restrict__ in, float *__restrict__ out, long long Q, int Hi, int Wi) {
  const int Ho = Hi / 2, Wo = Wi / 2;
  const long long total = Q * Ho * Wo;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int x = (int)(i % Wo);
    const long long t = i / Wo;
    const int y = (int)(t % Ho);
    const long long q = t / Ho;
    const float *s = in + (q * Hi + 2 * y) * Wi + 2 * x;
    out[i] = (((s[0] + s[1]) + s[Wi]) + s[Wi + 1]) * 0.25f;
  }
}

int launch_pool(const float *in, float *out, long long Q, int Hi, int Wi, cudaStream_t stream) {
  const long long total = Q * (Hi / 2) * (Wi / 2);
  if (total == 0) return 0;
  long long blocks = (total + 255) / 256;
  const long long cap = (long long)b200::num_sms() * 16;
  avgpool2_kernel<<<(int)(blocks < cap ? blocks : cap), 256, 0, stream>>>(in, out, Q, Hi, Wi);
  B200_LAUNCH_OK("avgpool2_kernel");
  return 0;
}

}  // namespace

extern "C" {

size_t b200corr_allpairs_rect_workspace_bytes(int B, int C, int H1, int W1, int H2, int W2, int precision) {
  if (precision == B200CORR_PREC_FP32) return 0;
  const size_t Cp = (size_t)(C + 31) / 32 * 32;
  return (precision == B200CORR_PREC_TF32X3 ? 2 : 1) * (size_t)B * ((size_t)H1 * W1 + (size_t)H2 * W2) * Cp * sizeof(float);
}

size_t b200corr_allpairs_workspace_bytes(int B, int C, int H, int W, int precision) {
  return b200corr_allpairs_rect_workspace_bytes(B, C, H, W, H, W, precision);
}

int b200corr_allpairs_pyramid(const float *f1, const float *f2, float *const *h_levels,
                              int num_levels, int B, int C, int H, int W, float scale,
                              int precision, void *workspace, size_t workspace_bytes, void *stream_) {
  return b200corr_allpairs_pyramid_rect(f1, f2, h_levels, num_levels, B, C, H, W, H, W, scale, precision, workspace,
                                        workspace_bytes, stream_);
}

int b200corr_allpairs_pyramid_rect(const float *f1, const float *f2, float *const *h_levels,
                                   int num_levels, int B, int C, int H1, int W1, int H, int W, float scale,
                                   int precision, void *workspace, size_t workspace_bytes, void *stream_) {
  return b200corr_allpairs_pyramid_layout(f1, f2, h_levels, num_levels, B, C, H1, W1, H, W, scale, precision, 0,
                                          workspace, workspace_bytes, stream_);
}

void b200corr_blocked_level_dims(int level, int H, int W, int *Hp, int *Wp) {
  const int H0 = (H + 7) / 8 * 8;                 // level 0: whole tile rows; W % 8 == 0 is a precondition
  if (level == 0) { *Hp = H0; *Wp = W; return; }
  *Hp = (H0 / 2 + 7) / 8 * 8;                     // level 1: the H0 / 2 rows the patches produce, in whole tiles
  *Wp = (W + 15) / 16 * 8;                        //          columns of the last tile past W / 2 hold zeros
}

int b200corr_allpairs_blocked_levels(int num_levels, int H, int W, int precision) {
  if (precision == B200CORR_PREC_FP32 || b200::num_sms() % 2 != 0) return 0;
  if (W % 8 != 0 || H < 1) return 0;
  return num_levels == 1 ? 1 : 3;
}

int b200corr_allpairs_pyramid_layout(const float *f1, const float *f2, float *const *h_levels,
                                     int num_levels, int B, int C, int H1, int W1, int H, int W, float scale,
                                     int precision, int blocked_levels, void *workspace, size_t workspace_bytes,
                                     void *stream_) {
  return b200corr_allpairs_pyramid_storage(f1, f2, h_levels, num_levels, B, C, H1, W1, H, W, scale, precision,
                                           blocked_levels, 0, workspace, workspace_bytes, stream_);
}

int b200corr_allpairs_pyramid_storage(const float *f1, const float *f2, float *const *h_levels,
                                      int num_levels, int B, int C, int H1, int W1, int H, int W, float scale,
                                      int precision, int blocked_levels, int half_levels, void *workspace,
                                      size_t workspace_bytes, void *stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  B200_CHECK(half_levels == 0 || (blocked_levels != 0 && half_levels == blocked_levels),
             "allpairs_pyramid: fp16 storage covers exactly the blocked levels (half_levels %d, blocked_levels %d)",
             half_levels, blocked_levels);
  B200_CHECK(blocked_levels == 0 || blocked_levels == b200corr_allpairs_blocked_levels(num_levels, H, W, precision),
             "allpairs_pyramid: blocked_levels %d not available for this problem (ask b200corr_allpairs_blocked_levels)",
             blocked_levels);
  B200_CHECK(num_levels >= 1 && num_levels <= 8, "allpairs_pyramid: num_levels must be in [1, 8]");
  B200_CHECK(B >= 0 && C >= 1 && H >= 1 && W >= 1 && H1 >= 1 && W1 >= 1, "allpairs_pyramid: bad sizes");
  B200_CHECK(precision == B200CORR_PREC_TF32 || precision == B200CORR_PREC_TF32X3 || precision == B200CORR_PREC_FP32,
             "allpairs_pyramid: precision %d not available (TF32 = 0, TF32X3 = 1, FP32 = 2)", precision);
  if (B == 0) return 0;
  B200_CHECK(f1 && f2 && h_levels && h_levels[0], "allpairs_pyramid: null pointer");
  const int HW = H * W;        // key pixels (the maps of f2 and of every level)
  const int HWq = H1 * W1;     // query pixels (f1)
  int LH[8], LW[8];
  LH[0] = H; LW[0] = W;
  for (int l = 1; l < num_levels; ++l) {
    LH[l] = LH[l - 1] / 2;
    LW[l] = LW[l - 1] / 2;
    B200_CHECK(h_levels[l] || LH[l] * LW[l] == 0, "allpairs_pyramid: null level %d", l);
  }
  const long long Q = (long long)B * HWq;
  int first_unpooled = 1;  // first level that still has to be produced by the pooling kernel

  if (precision == B200CORR_PREC_FP32 || W % 4 != 0) {
    B200_CHECK(HWq == HW, "allpairs_pyramid_rect: the exact fp32 / W %% 4 != 0 path needs equal map sizes");
    dim3 grid((HW + 63) / 64, (HW + 63) / 64, B);
    allpairs_simt_kernel<<<grid, 256, 0, stream>>>(f1, f2, h_levels[0], C, HW, scale);
    B200_LAUNCH_OK("allpairs_simt_kernel");
  } else {
    const int Cp = (C + 31) / 32 * 32;
    // CTA pairs (cta_group::2) by default; B200CORR_ALLPAIRS_CTAS=1 selects the single-CTA kernel
    int ctas = 2;
    {
      const char *e = blocked_levels ? nullptr : getenv("B200CORR_ALLPAIRS_CTAS");   // the blocked layout needs CTA pairs
      if (e && atoi(e) == 1) ctas = 1;
      if (b200::num_sms() % 2) ctas = 1;
    }
    const bool x3 = precision == B200CORR_PREC_TF32X3;
    const size_t nfeat1 = (size_t)B * HWq * Cp, nfeat2 = (size_t)B * HW * Cp;
    const size_t need = (x3 ? 2 : 1) * (nfeat1 + nfeat2) * sizeof(float);
    B200_CHECK(workspace && workspace_bytes >= need,
               "allpairs_pyramid: workspace too small (%zu < %zu bytes)", workspace_bytes, need);
    B200_CHECK(((uintptr_t)workspace & 127) == 0 && ((uintptr_t)h_levels[0] & 15) == 0,
               "allpairs_pyramid: workspace must be 128-byte and level 0 16-byte aligned");
    float *f1t = (float *)workspace, *f2t = f1t + nfeat1;
    float *f1lo = x3 ? f2t + nfeat2 : nullptr, *f2lo = x3 ? f1lo + nfeat1 : nullptr;
    {
      const int hw_max = HWq > HW ? HWq : HW;
      dim3 pgrid((hw_max + 31) / 32, (Cp + 127) / 128, 2 * B);
      const PrepMap m0{f1, f1t, f1lo, HWq}, m1{f2, f2t, f2lo, HW};
      prep_kmajor_tf32_kernel<<<pgrid, 256, 0, stream>>>(m0, m1, B, C, Cp);
      B200_LAUNCH_OK("prep_kmajor_tf32_kernel");
    }

    CUtensorMap mapA, mapB, mapAlo, mapBlo;
    for (int which = 0; which < (x3 ? 2 : 1); ++which) {
      CUtensorMap &mapA_ = which ? mapAlo : mapA, &mapB_ = which ? mapBlo : mapB;
      const float *f1t_ = which ? f1lo : f1t, *f2t_ = which ? f2lo : f2t;
    {
      const uint64_t dims[3] = {(uint64_t)Cp, (uint64_t)HWq, (uint64_t)B};
      const uint64_t str[3] = {4, (uint64_t)Cp * 4, (uint64_t)HWq * Cp * 4};
      const uint32_t box[3] = {tc::BK, tc::BM, 1};
      if (int e = b200::make_tensor_map(&mapA_, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, f1t_, dims, str, box,
                                        CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B))
        return e;
    }
    if (blocked_levels) {
      // (K, x % 8, y, x / 8, b): a [32][8][4][4] box is one half patch in tile order (see the BLK kernel)
      const uint64_t dims[5] = {(uint64_t)Cp, 8, (uint64_t)H, (uint64_t)(W / 8), (uint64_t)B};
      const uint64_t str[5] = {4, (uint64_t)Cp * 4, (uint64_t)W * Cp * 4, (uint64_t)8 * Cp * 4, (uint64_t)HW * Cp * 4};
      const uint32_t box[5] = {tc::BK, 8, (uint32_t)(tc::PH / ctas), 4, 1};
      if (int e = b200::make_tensor_map(&mapB_, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 5, f2t_, dims, str, box,
                                        CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B))
        return e;
    } else {
      const uint64_t dims[4] = {(uint64_t)Cp, (uint64_t)W, (uint64_t)H, (uint64_t)B};
      const uint64_t str[4] = {4, (uint64_t)Cp * 4, (uint64_t)W * Cp * 4, (uint64_t)HW * Cp * 4};
      const uint32_t box[4] = {tc::BK, (uint32_t)tc::PW, (uint32_t)(tc::PH / ctas), 1};
      if (int e = b200::make_tensor_map(&mapB_, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, f2t_, dims, str, box,
                                        CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B))
        return e;
    }
    }
    if (!x3) { mapAlo = mapA; mapBlo = mapB; }
    tc::Params p;
    p.B = B; p.HW = HWq; p.H = H; p.W = W; p.KB = Cp / 32;
    p.passes = x3 ? 3 : 1;
    p.MT = (HWq + tc::BM * ctas - 1) / (tc::BM * ctas);
    p.NTY = (H + tc::PH - 1) / tc::PH;
    p.NTX = (W + tc::PW - 1) / tc::PW;
    p.scale = scale;
    {
      int hp, wp;
      b200corr_blocked_level_dims(0, H, W, &hp, &wp);
      p.TW0 = wp / 8; p.S0 = (long long)hp * wp;
      b200corr_blocked_level_dims(1, H, W, &hp, &wp);
      p.TW1 = wp / 8; p.S1 = (long long)hp * wp;
    }
    p.lvl0 = h_levels[0];
    {
      const char *dbg = getenv("B200CORR_DEBUG");
      p.debug = dbg ? atoi(dbg) : 0;
    }
    for (int l = 0; l < 3; ++l) {
      const bool want = l + 1 < num_levels && LH[l + 1] * LW[l + 1] > 0;
      p.lvl[l] = want ? h_levels[l + 1] : nullptr;
      p.LH[l] = want ? LH[l + 1] : 0;
      p.LW[l] = want ? LW[l + 1] : 0;
      if (want) B200_CHECK(((uintptr_t)h_levels[l + 1] & 15) == 0, "allpairs_pyramid: level %d misaligned", l + 1);
    }
    first_unpooled = 4;
    const int total = B * p.MT * p.NTY * p.NTX;
    if (ctas == 2) {
      static bool attr_done[64] = {};  // per device
      if (int e = b200::set_max_smem_once((const void *)allpairs_tc_kernel<2, false>, tc::Cfg<2>::SMEM_BYTES, attr_done)) return e;
      const int pairs = b200::num_sms() / 2;
      cudaLaunchConfig_t cfg = {};
      cfg.gridDim = dim3(2 * (total < pairs ? total : pairs));
      cfg.blockDim = dim3(tc::THREADS);
      cfg.dynamicSmemBytes = tc::Cfg<2>::SMEM_BYTES;
      cfg.stream = stream;
      cudaLaunchAttribute attr[1];
      attr[0].id = cudaLaunchAttributeClusterDimension;
      attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
      cfg.attrs = attr;
      cfg.numAttrs = 1;
      if (blocked_levels && half_levels) {
        static bool attr_blkh[64] = {};
        if (int e = b200::set_max_smem_once((const void *)allpairs_tc_kernel<2, true, true>, tc::Cfg<2>::SMEM_BYTES, attr_blkh)) return e;
        B200_CUDA(cudaLaunchKernelEx(&cfg, allpairs_tc_kernel<2, true, true>, mapA, mapB, mapAlo, mapBlo, p));
      } else if (blocked_levels) {
        static bool attr_blk[64] = {};
        if (int e = b200::set_max_smem_once((const void *)allpairs_tc_kernel<2, true>, tc::Cfg<2>::SMEM_BYTES, attr_blk)) return e;
        B200_CUDA(cudaLaunchKernelEx(&cfg, allpairs_tc_kernel<2, true>, mapA, mapB, mapAlo, mapBlo, p));
      } else {
        B200_CUDA(cudaLaunchKernelEx(&cfg, allpairs_tc_kernel<2, false>, mapA, mapB, mapAlo, mapBlo, p));
      }
    } else {
      static bool attr_done[64] = {};  // per device
      if (int e = b200::set_max_smem_once((const void *)allpairs_tc_kernel<1, false>, tc::Cfg<1>::SMEM_BYTES, attr_done)) return e;
      const int grid = total < b200::num_sms() ? total : b200::num_sms();
      allpairs_tc_kernel<1, false><<<grid, tc::THREADS, tc::Cfg<1>::SMEM_BYTES, stream>>>(mapA, mapB, mapAlo, mapBlo, p);
    }
    B200_LAUNCH_OK("allpairs_tc_kernel");
  }
  for (int l = first_unpooled; l < num_levels; ++l)
    if (LH[l] * LW[l] > 0)
      if (int e = launch_pool(h_levels[l - 1], h_levels[l], Q, LH[l - 1], LW[l - 1], stream)) return e;
  return 0;
}

}  // extern "C"
