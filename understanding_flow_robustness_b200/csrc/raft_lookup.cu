// raft_lookup.cu -- RAFT correlation lookup (forward + backward) and pyramid backward.
//
// Replaces CorrBlock.__call__ + bilinear_sampler of the reference (models/raft/corr.py:72-96,
// models/raft/utils/utils.py:62-76): per pyramid level a CPU-built (2r+1)^2 offset grid copied to
// the device, one F.grid_sample launch, then cat + permute + contiguous -- 4 grid_sample launches,
// 4 H2D copies and 2 extra passes over the (B, 324, H, W) result per RAFT iteration.
//
// Here: ONE launch per lookup for all levels.  A 6-warp CTA owns 32 consecutive query pixels of one
// level; lane = query in every phase:
//   staging: warp w fetches rows 2w, 2w+1 of each query's (2r+4)^2 neighbourhood with sector-exact
//            256-bit loads (level width % 8 == 0), 128-bit loads (width % 4 == 0) or scalar loads --
//            only the rows and sectors the taps can touch, all of a warp's loads in flight together --
//            and parks them UNSHIFTED in a query-minor shared tile win[row][column][query]: any
//            per-query offset is bank-conflict free.  Everything outside the slice is staged as zero
//            (= grid_sample's zero padding).  (The kernel is NOT bound by this gather: with the loads
//            switched off it takes the same time -- DESIGN.md 2.4; its time is the chain of phases
//            of a CTA.)
//   taps:    the 2 x (2r+1) tap positions / fractions of every query are shared out over the warps;
//   sample:  warp w takes the y offsets j = w, w+6, ...; every store of out[b, l*81 + k, q] is a full
//            128-byte line; the result is written once, in its final (B, L*(2r+1)^2, H, W) layout.
// Channel order k = i*(2r+1) + j with i the x offset and j the y offset (corr.py:80-86).
//
// Coordinate arithmetic, mode B200CORR_LOOKUP_GRIDSAMPLE: the reference normalises the pixel
// coordinate to [-1,1] (utils.py:65-67) and grid_sample un-normalises it again
// (ATen grid_sampler_unnormalize, align_corners=True: ((g + 1) / 2) * (size - 1)); both steps are
// reproduced operation by operation in fp32 (no FMA contraction) so that integer coordinates come
// back as e.g. 79.99999 exactly as in the reference.  Mode B200CORR_LOOKUP_DIRECT samples at the
// pixel coordinate itself (what alt_cuda_corr does).
//
// Backward (what autograd derives for the reference): the bilinear weights,
// times the output gradient, are added into the query's own slice of a dense per-level gradient
// volume -- same CTA / lane mapping as the forward, each warp owning three window rows, so neither
// shared nor global atomics are needed (see lookup_bwd_kernel).  Coordinates get no gradient
// (raft.py:188 detaches them).
#include "raft_lookup.cuh"

#ifndef B200_LOOKUP_BWD_RED
#define B200_LOOKUP_BWD_RED 1     // 73.9 -> 71.2 us per lookup backward at B=4 (scripts/time_lookup_bwd_variants.py); 0 = read-modify-write
#endif

namespace {
using namespace b200lookup;

// CTA = NW (6) warps x the same 32 queries of one level; lane = query everywhere.
//   warp w stages window rows w*RPW .. and computes every NW-th tap pair          -> one barrier
//   warp w samples the y offsets j = w, w+NW, ...                                 -> 128-byte stores
// Warps per CTA of the forward kernel.  6 x 2 window rows instead of 4 x 3: 78 instead of 105 registers, so the
// same 4 CTAs per SM hold 24 warps instead of 16, and every warp's chain (loads, parking, taps, samples) is a third
// shorter: 30.9 -> 29.8 us (KITTI), 28.3 -> 27.2 (Sintel), 35.6 -> 33.9 (FlyingThings) at B=4.  8 warps: 34.5 us (the
// per-warp prologue is redundant work); 6 warps squeezed to 64 registers for 5 CTAs per SM: 33.1 us (spills).
#ifndef B200_LOOKUP_NW
#define B200_LOOKUP_NW 6
#endif
#ifndef B200_LOOKUP_MINB
#define B200_LOOKUP_MINB 4        // CTAs per SM the register allocation aims at
#endif
constexpr int kFwdWarps = B200_LOOKUP_NW;

template <int R>
__global__ void __launch_bounds__(32 * kFwdWarps, B200_LOOKUP_MINB)
lookup_fwd_kernel(const LookupParams p, const float *__restrict__ coords, float *__restrict__ out) {
  constexpr int NW = kFwdWarps;
  constexpr int N = Geo<R>::N, WS = Geo<R>::WS, RPW0 = (WS + NW - 1) / NW, RPW = RPW0 < 2 ? 2 : RPW0;
  __shared__ float win[WS * kCols * 32];   // [row][column][lane]: conflict-free for any per-lane offset
  __shared__ int tab_r[2 * N][32];         // tap tables [axis * N + tap][lane]
  __shared__ float tab_a[2 * N][32];
  // levels interleaved across consecutive CTAs: DRAM-heavy level 0 and the cache-resident coarse
  // levels share every SM
  const int lvl = blockIdx.x % p.num_levels, q0 = (blockIdx.x / p.num_levels) * QT, b = blockIdx.z;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int q = q0 + lane;
  const int mode = p.mode, path = p.path[lvl];
  StageArgs a;
  a.q_ok = q < p.HW;
  a.LH = p.LH[lvl];
  a.LW = p.LW[lvl];
  float cx = 0.f, cy = 0.f;
  if (a.q_ok) {
    cx = coords[((size_t)b * 2 + 0) * p.HW + q];
    cy = coords[((size_t)b * 2 + 1) * p.HW + q];
  }
  int xlo, xhi;
  const int slvl = lvl + p.first_level;   // coordinate scale of this level
  a.ox = window_origin<R>(cx, slvl, xlo, xhi);
  a.oy = window_origin<R>(cy, slvl, a.ylo, a.yhi);
  // staged column of window column 0, and the staged columns the taps touch
  const int shift = path == PATH_SCALAR ? 0 : (a.ox & 3);
  a.clo = shift + xlo; a.chi = shift + xhi;
  a.blocked = p.blocked[lvl] != 0;
  {
    const size_t e0 = ((size_t)b * p.HW + (a.q_ok ? q : 0)) * (size_t)p.slice[lvl];   // in elements
    a.slice = path == PATH_HALF ? reinterpret_cast<const float *>(reinterpret_cast<const __half *>(p.lvl[lvl]) + e0) : p.lvl[lvl] + e0;
  }
  a.tiles_w = p.tiles_w[lvl];

  // ---- this warp's window rows: all loads in flight together, parked in the tile after the tap tables
  float sv[RPW][24];
  if (path == PATH_SECTOR) stage_load<PATH_SECTOR, RPW, WS>(a, warp * RPW0, sv);
  else if (path == PATH_HALF) stage_load_half<WS, RPW, NW>(a, warp, sv);
  else if (path == PATH_VEC4) stage_load<PATH_VEC4, RPW, WS>(a, warp * RPW0, sv);
  else stage_load<PATH_SCALAR, RPW, WS>(a, warp * RPW0, sv);
  // ---- this warp's share of the 2N tap table entries
  const float smx = (float)(a.LW - 1), smy = (float)(a.LH - 1);
  const float ismx = p.inv_w[lvl], ismy = p.inv_h[lvl];   // host-computed RN reciprocals
#pragma unroll 1
  for (int t = warp; t < N; t += NW) {
    int relx, rely;
    float fracx, fracy;
    two_taps<R>(cx, cy, slvl, t, smx, smy, ismx, ismy, mode, a.ox, a.oy, relx, fracx, rely, fracy);
    // a tap outside the staged rows / columns cannot happen (window_origin); drop it if it does
    if (relx >= 0 && (relx < xlo || relx + 1 > xhi)) { relx = -1; fracx = 0.f; }
    if (rely >= 0 && (rely < a.ylo || rely + 1 > a.yhi)) { rely = -1; fracy = 0.f; }
    tab_r[t][lane] = relx;
    tab_a[t][lane] = fracx;
    tab_r[N + t][lane] = rely;
    tab_a[N + t][lane] = fracy;
  }
  if (path == PATH_SECTOR) stage_store<PATH_SECTOR, RPW, WS>(win, lane, a, warp * RPW0, sv);
  else if (path == PATH_HALF) stage_store_half<WS, RPW, NW>(win, lane, a, warp, sv);
  else if (path == PATH_VEC4) stage_store<PATH_VEC4, RPW, WS>(win, lane, a, warp * RPW0, sv);
  else stage_store<PATH_SCALAR, RPW, WS>(win, lane, a, warp * RPW0, sv);
  __syncthreads();
  if (!a.q_ok) return;

  // ---- sample: every store is one full 128-byte line
  int rxs[N];
  float axs[N], bxs[N];
  bool fast = true;   // x taps consecutive and inside the window (the common case)
#pragma unroll
  for (int i = 0; i < N; ++i) {
    rxs[i] = tab_r[i][lane];
    axs[i] = tab_a[i][lane];
    bxs[i] = 1.f - axs[i];
    fast = fast && rxs[i] == rxs[0] + i && rxs[0] >= 0;
  }
  const int nchan = p.num_levels * N * N;
  float *outq = out + ((size_t)b * nchan + (size_t)lvl * N * N) * p.HW + q;
  const size_t istride = (size_t)N * p.HW;
  const float *wl = win + lane;
#pragma unroll 1
  for (int j = warp; j < N; j += NW) {
    const int ry = tab_r[N + j][lane];
    const float ay = tab_a[N + j][lane], by = 1.f - ay;
    float *o = outq + (size_t)j * p.HW;
    if (fast && ry >= 0) {
      // N+1 consecutive columns of both rows, then every tap is 4 products of registers
      const float *r0 = wl + (ry * kCols + shift + rxs[0]) * 32;
      float c0[N + 1], c1[N + 1];
#pragma unroll
      for (int k = 0; k <= N; ++k) {
        c0[k] = r0[k * 32];
        c1[k] = r0[(kCols + k) * 32];
      }
#pragma unroll
      for (int i = 0; i < N; ++i) {
        *o = bilerp(c0[i], c0[i + 1], c1[i], c1[i + 1], axs[i], bxs[i], ay, by);
        o += istride;
      }
    } else {
#pragma unroll
      for (int i = 0; i < N; ++i) {
        const int rx = rxs[i];
        float v = 0.f;
        if (rx >= 0 && ry >= 0) {
          const float *r0 = wl + (ry * kCols + shift + rx) * 32;
          v = bilerp(r0[0], r0[32], r0[kCols * 32], r0[(kCols + 1) * 32], axs[i], bxs[i], ay, by);
        }
        *o = v;
        o += istride;
      }
    }
  }
}

// plain (coherent) 256-bit / 128-bit accesses for the read-modify-write of the gradient slices
__device__ __forceinline__ void ld256(const float *p, float (&v)[8]) {
  asm volatile("ld.global.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=f"(v[0]), "=f"(v[1]), "=f"(v[2]), "=f"(v[3]), "=f"(v[4]), "=f"(v[5]), "=f"(v[6]), "=f"(v[7])
               : "l"(p)
               : "memory");
}
__device__ __forceinline__ void st256(float *p, const float (&v)[8]) {
  asm volatile("st.global.v8.f32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(p), "f"(v[0]), "f"(v[1]), "f"(v[2]), "f"(v[3]),
               "f"(v[4]), "f"(v[5]), "f"(v[6]), "f"(v[7])
               : "memory");
}

// add one accumulated window row (staged columns 0..15 of `wrow`, lane-minor) into the query's slice
template <int PATH>
__device__ __forceinline__ void flush_row(const float *wrow, float *grow, int ox, int LW, int clo, int chi) {
  if (PATH == PATH_SECTOR) {
    const int c0 = ox & ~7, off = ox & 4;   // staged column s sits at loaded column s + off
#pragma unroll
    for (int g = 0; g < 3; ++g) {
      const int x = c0 + 8 * g;
      if (x < 0 || x >= LW || chi + off < 8 * g || clo + off >= 8 * g + 8) continue;
      float v[8];
      ld256(grow + x, v);
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const int sidx = 8 * g + k - off;
        if (sidx >= 0 && sidx < kCols) v[k] += wrow[sidx * 32];
      }
      st256(grow + x, v);
    }
  } else if (PATH == PATH_VEC4) {
    const int c0 = ox & ~3;
#pragma unroll
    for (int g = 0; g < 4; ++g) {
      const int x = c0 + 4 * g;
      if (x < 0 || x >= LW || chi < 4 * g || clo >= 4 * g + 4) continue;
      float4 v = *reinterpret_cast<float4 *>(grow + x);
      v.x += wrow[(4 * g) * 32]; v.y += wrow[(4 * g + 1) * 32];
      v.z += wrow[(4 * g + 2) * 32]; v.w += wrow[(4 * g + 3) * 32];
      *reinterpret_cast<float4 *>(grow + x) = v;
    }
  } else {
#pragma unroll
    for (int k = 0; k < 12; ++k) {
      const int x = ox + k;
      const float a = wrow[k * 32];
      if (x >= 0 && x < LW && a != 0.f) grow[x] += a;
    }
  }
}

// schedulable 256-bit load (no volatile / memory clobber): the flush below puts all of a warp's sector loads in
// flight before the first add -- the serialised load -> add -> store chain was 7.8 long_scoreboard stalls per issue
__device__ __forceinline__ void ld256_sched(const float *p, float (&v)[8]) {
  asm("ld.global.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
      : "=f"(v[0]), "=f"(v[1]), "=f"(v[2]), "=f"(v[3]), "=f"(v[4]), "=f"(v[5]), "=f"(v[6]), "=f"(v[7])
      : "l"(p));
}
__device__ __forceinline__ void prefetch_l2(const float *p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }

// which (window row r of this warp, sector g) pairs the flush touches, and their addresses
template <int RPW, int WS>
struct SectorPlan {
  bool ok[RPW][3];
  float *ptr[RPW][3];
  __device__ __forceinline__ void make(float *slice, int warp, int oy, int ox, int LH, int LW, int ylo, int yhi, int clo,
                                       int chi) {
    const int c0 = ox & ~7, off = ox & 4;
#pragma unroll
    for (int r = 0; r < RPW; ++r) {
      const int wr = warp * RPW + r, y = oy + wr;
      const bool row_ok = wr < WS && wr >= ylo && wr <= yhi && y >= 0 && y < LH;
#pragma unroll
      for (int g = 0; g < 3; ++g) {
        const int x = c0 + 8 * g;
        ok[r][g] = row_ok && x >= 0 && x < LW && !(chi + off < 8 * g || clo + off >= 8 * g + 8);
        ptr[r][g] = slice + (size_t)(row_ok ? y : 0) * LW + (ok[r][g] ? x : 0);
      }
    }
  }
};

// Backward of the lookup (what autograd derives for grid_sample): the bilinear weights of every tap,
// times its output gradient, are added into the query's own slice of the dense gradient pyramid.
// Same decomposition as the forward kernel: CTA = 6 warps x 32 consecutive queries of one level,
// lane = query.  Warp w OWNS window rows 2w, 2w+1 (4 warps x 3 rows: 71.0 us, 6 x 2: 67.8, 8: 71.7, 12 x 1: 80.0): it gathers every contribution to those rows (in
// registers when the taps sit on consecutive window positions -- the common case -- else by
// read-modify-write of its rows of the shared tile), then adds the rows into the slice with
// sector-sized read-modify-writes.  A slice is touched by exactly one CTA per launch and a row by
// exactly one warp: no atomics anywhere, deterministic.
#ifndef B200_LOOKUP_BWD_NW
#define B200_LOOKUP_BWD_NW 6      // warps per CTA of the backward kernel (each owns ceil(WS / NW) window rows): 4 -> 6 warps 71.0 -> 67.8 us
#endif
constexpr int kBwdWarps = B200_LOOKUP_BWD_NW;

template <int R>
__global__ void __launch_bounds__(32 * kBwdWarps, 4)
lookup_bwd_kernel(const LookupParams p, const float *__restrict__ coords,
                  const float *__restrict__ gout) {
  constexpr int N = Geo<R>::N, WS = Geo<R>::WS, RPW = (WS + kBwdWarps - 1) / kBwdWarps;
  __shared__ float win[WS * kCols * 32];   // [row][column][lane]
  __shared__ int tab_r[2 * N][32];
  __shared__ float tab_a[2 * N][32];
  const int lvl = blockIdx.x % p.num_levels, q0 = (blockIdx.x / p.num_levels) * QT, b = blockIdx.z;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int q = q0 + lane;
  const bool q_ok = q < p.HW;
  const int mode = p.mode, path = p.path[lvl];
  const int LH = p.LH[lvl], LW = p.LW[lvl];
  float cx = 0.f, cy = 0.f;
  if (q_ok) {
    cx = coords[((size_t)b * 2 + 0) * p.HW + q];
    cy = coords[((size_t)b * 2 + 1) * p.HW + q];
  }
  int xlo, xhi, ylo, yhi;
  const int slvl = lvl + p.first_level;   // coordinate scale of this level
  const int ox = window_origin<R>(cx, slvl, xlo, xhi), oy = window_origin<R>(cy, slvl, ylo, yhi);
  const int shift = path == PATH_SCALAR ? 0 : (ox & 3);
  // the sectors of the gradient slice this warp will read-modify-write at the end: start them towards the L2 now,
  // the tap tables and the gathering below hide the DRAM latency
  float *const gslice = p.glvl[lvl] + ((size_t)b * p.HW + (q_ok ? q : 0)) * LH * LW;
  if (path == PATH_SECTOR && q_ok) {
    SectorPlan<RPW, WS> sp;
    sp.make(gslice, warp, oy, ox, LH, LW, ylo, yhi, shift + xlo, shift + xhi);
#pragma unroll
    for (int r = 0; r < RPW; ++r)
#pragma unroll
      for (int g = 0; g < 3; ++g)
        if (sp.ok[r][g]) prefetch_l2(sp.ptr[r][g]);
  }
  const float smx = (float)(LW - 1), smy = (float)(LH - 1);
  const float ismx = p.inv_w[lvl], ismy = p.inv_h[lvl];   // host-computed RN reciprocals
#pragma unroll 1
  for (int t = warp; t < N; t += kBwdWarps) {
    int relx, rely;
    float fracx, fracy;
    two_taps<R>(cx, cy, slvl, t, smx, smy, ismx, ismy, mode, ox, oy, relx, fracx, rely, fracy);
    if (relx >= 0 && (relx < xlo || relx + 1 > xhi)) { relx = -1; fracx = 0.f; }   // same rule as the forward
    if (rely >= 0 && (rely < ylo || rely + 1 > yhi)) { rely = -1; fracy = 0.f; }
    tab_r[t][lane] = relx;
    tab_a[t][lane] = fracx;
    tab_r[N + t][lane] = rely;
    tab_a[N + t][lane] = fracy;
  }
  __syncthreads();

  int rxs[N];
  float axs[N], bxs[N];
  bool fast = true;   // x and y taps on consecutive window positions
#pragma unroll
  for (int i = 0; i < N; ++i) {
    rxs[i] = tab_r[i][lane];
    axs[i] = tab_a[i][lane];
    bxs[i] = 1.f - axs[i];
    fast = fast && rxs[i] == rxs[0] + i && rxs[0] >= 0;
  }
  const int ry0 = tab_r[N][lane];
#pragma unroll
  for (int j = 1; j < N; ++j) fast = fast && tab_r[N + j][lane] == ry0 + j;
  fast = __all_sync(0xffffffffu, fast && ry0 >= 0);
  const int nchan = p.num_levels * N * N;
  const float *gq = gout + ((size_t)b * nchan + (size_t)lvl * N * N) * p.HW + q;
  float *wl = win + lane;

  if (fast) {
#pragma unroll 1
    for (int r = 0; r < RPW; ++r) {
      const int wr = warp * RPW + r;
      if (wr >= WS) break;
      float acc[N + 1];
#pragma unroll
      for (int k = 0; k <= N; ++k) acc[k] = 0.f;
      // row wr collects the upper taps of y offset j = wr - ry0 and the lower taps of j = wr - 1 - ry0
#pragma unroll
      for (int half = 0; half < 2; ++half) {
        const int j = wr - half - ry0;
        const bool ok = q_ok && j >= 0 && j < N;
        const int jj = ok ? j : 0;
        const float ay = tab_a[N + jj][lane];
        const float wy = half ? ay : 1.f - ay;
#pragma unroll
        for (int i = 0; i < N; ++i) {
          const float g = ok ? __ldg(gq + (size_t)(i * N + jj) * p.HW) : 0.f;
          acc[i] = __fmaf_rn(g, __fmul_rn(bxs[i], wy), acc[i]);
          acc[i + 1] = __fmaf_rn(g, __fmul_rn(axs[i], wy), acc[i + 1]);
        }
      }
      float *dst = wl + (wr * kCols) * 32;
#pragma unroll
      for (int c = 0; c < kCols; ++c) dst[c * 32] = 0.f;
#pragma unroll
      for (int k = 0; k <= N; ++k) dst[(shift + rxs[0] + k) * 32] = acc[k];
    }
  } else {
    // general case: this warp's rows are a private accumulator for whatever lands on them
#pragma unroll 1
    for (int r = 0; r < RPW; ++r) {
      const int wr = warp * RPW + r;
      if (wr >= WS) break;
#pragma unroll
      for (int c = 0; c < kCols; ++c) wl[(wr * kCols + c) * 32] = 0.f;
    }
#pragma unroll 1
    for (int j = 0; j < N; ++j) {
      const int ry = tab_r[N + j][lane];
      const float ay = tab_a[N + j][lane];
#pragma unroll 1
      for (int half = 0; half < 2; ++half) {
        const int row = ry + half;
        const bool mine = q_ok && ry >= 0 && row >= warp * RPW && row < warp * RPW + RPW && row < WS;
        const float wy = half ? ay : 1.f - ay;
#pragma unroll
        for (int i = 0; i < N; ++i) {
          if (mine && rxs[i] >= 0) {
            const float g = __ldg(gq + (size_t)(i * N + j) * p.HW);
            float *c = wl + (row * kCols + shift + rxs[i]) * 32;
            c[0] = __fmaf_rn(g, __fmul_rn(bxs[i], wy), c[0]);
            c[32] = __fmaf_rn(g, __fmul_rn(axs[i], wy), c[32]);
          }
        }
      }
    }
  }
  __syncwarp();
  if (!q_ok) return;

  // ---- add this warp's rows into the slice (only rows / sectors the taps can have touched)
  float *slice = gslice;
  const int clo = shift + xlo, chi = shift + xhi;
  if (path == PATH_SECTOR) {
    // all sector loads of this warp's rows first, then the adds and the stores
    SectorPlan<RPW, WS> sp;
    sp.make(slice, warp, oy, ox, LH, LW, ylo, yhi, clo, chi);
    const int off = ox & 4;
#if B200_LOOKUP_BWD_RED
    // fire-and-forget vector reductions at the L2 instead of load -> add -> store through the SM: every address is
    // still added to by exactly one lane per launch (launches are stream-ordered), so the sums stay deterministic
#pragma unroll
    for (int r = 0; r < RPW; ++r) {
      const float *wrow = wl + ((warp * RPW + r) * kCols) * 32;
#pragma unroll
      for (int g = 0; g < 3; ++g) {
        if (!sp.ok[r][g]) continue;
        float a[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          const int sidx = 8 * g + k - off;
          a[k] = (sidx >= 0 && sidx < kCols) ? wrow[sidx * 32] : 0.f;
        }
        asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(sp.ptr[r][g]), "f"(a[0]), "f"(a[1]), "f"(a[2]), "f"(a[3]) : "memory");
        asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(sp.ptr[r][g] + 4), "f"(a[4]), "f"(a[5]), "f"(a[6]), "f"(a[7]) : "memory");
      }
    }
    return;
#endif
    float v[RPW][3][8];
#pragma unroll
    for (int r = 0; r < RPW; ++r)
#pragma unroll
      for (int g = 0; g < 3; ++g)
        if (sp.ok[r][g]) ld256_sched(sp.ptr[r][g], v[r][g]);
#pragma unroll
    for (int r = 0; r < RPW; ++r) {
      const float *wrow = wl + ((warp * RPW + r) * kCols) * 32;
#pragma unroll
      for (int g = 0; g < 3; ++g) {
        if (!sp.ok[r][g]) continue;
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          const int sidx = 8 * g + k - off;
          if (sidx >= 0 && sidx < kCols) v[r][g][k] += wrow[sidx * 32];
        }
        st256(sp.ptr[r][g], v[r][g]);
      }
    }
    return;
  }
#pragma unroll 1
  for (int r = 0; r < RPW; ++r) {
    const int wr = warp * RPW + r, y = oy + wr;
    if (wr >= WS || wr < ylo || wr > yhi || y < 0 || y >= LH) continue;
    const float *wrow = wl + (wr * kCols) * 32;
    float *grow = slice + (size_t)y * LW;
    if (path == PATH_SECTOR) flush_row<PATH_SECTOR>(wrow, grow, ox, LW, clo, chi);
    else if (path == PATH_VEC4) flush_row<PATH_VEC4>(wrow, grow, ox, LW, clo, chi);
    else flush_row<PATH_SCALAR>(wrow, grow, ox, LW, clo, chi);
  }
}

// fine[q, y, x] += coarse[q, y/2, x/2] / 4   for y < 2*Hc, x < 2*Wc   (backward of avg_pool2d(2,2))
__global__ void __launch_bounds__(256)
pool_bwd_kernel(float *__restrict__ fine, const float *__restrict__ coarse, long long Q, int Hf, int Wf) {
  const int Hc = Hf / 2, Wc = Wf / 2;
  const long long total = Q * Hf * Wf;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int x = (int)(i % Wf);
    const long long t = i / Wf;
    const int y = (int)(t % Hf);
    const long long q = t / Hf;
    if (y < 2 * Hc && x < 2 * Wc) fine[i] += 0.25f * coarse[(q * Hc + (y >> 1)) * Wc + (x >> 1)];
  }
}

// the same for Wf % 4 == 0 and 16-byte aligned levels: one float4 of a fine row (and the two coarse values under it)
// per thread and step, row index arithmetic in 32 bits -- the fold is a pure streaming pass (read-modify-write of the
// fine level), 2.5x faster than the scalar kernel (1.10 -> 0.45 ms for the four levels at B=4, 48x160)
template <typename IDX>   // unsigned (rows * Wf / 4 < 2^31: 32-bit divisions) or long long
__global__ void __launch_bounds__(256)
pool_bwd_vec4_kernel(float *__restrict__ fine, const float *__restrict__ coarse, long long rows, int Hf, int Wf) {
  const int Hc = Hf / 2, Wc = Wf / 2;
  const IDX W4 = (IDX)(Wf / 4), HF = (IDX)Hf;
  const IDX total = (IDX)rows * W4;   // rows = Q * Hf
  const IDX stride = (IDX)gridDim.x * blockDim.x;
  for (IDX i = (IDX)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
    const IDX row = i / W4;
    const int x4 = (int)(i - row * W4);
    const IDX q = row / HF;
    const int y = (int)(row - q * HF);
    if (y >= 2 * Hc) continue;
    const float *c = coarse + ((size_t)q * Hc + (y >> 1)) * Wc + 2 * x4;
    const float c0 = 0.25f * c[0], c1 = 0.25f * c[1];   // Wf % 4 == 0: both coarse columns exist
    float4 *f = reinterpret_cast<float4 *>(fine + (size_t)row * Wf + 4 * x4);
    float4 v = *f;
    v.x += c0; v.y += c0; v.z += c1; v.w += c1;
    *f = v;
  }
}

}  // namespace

extern "C" {

int b200corr_lookup_forward(const float *const *h_levels, int num_levels, const float *coords,
                            float *out, int B, int H, int W, int radius, int mode, void *stream_) {
  return b200corr_lookup_forward_from(h_levels, num_levels, 0, coords, out, B, H, W, radius, mode, stream_);
}

int b200corr_lookup_forward_from(const float *const *h_levels, int num_levels, int first_level,
                                 const float *coords, float *out, int B, int H, int W, int radius, int mode,
                                 void *stream_) {
  return b200corr_lookup_forward_layout(h_levels, num_levels, first_level, 0, coords, out, B, H, W, radius, mode, stream_);
}

int b200corr_lookup_forward_layout(const float *const *h_levels, int num_levels, int first_level, int blocked_levels,
                                   const float *coords, float *out, int B, int H, int W, int radius, int mode,
                                   void *stream_) {
  return b200corr_lookup_forward_storage(h_levels, num_levels, first_level, blocked_levels, 0, coords, out, B, H, W,
                                         radius, mode, stream_);
}

int b200corr_lookup_forward_storage(const float *const *h_levels, int num_levels, int first_level, int blocked_levels,
                                    int half_levels, const float *coords, float *out, int B, int H, int W, int radius,
                                    int mode, void *stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  B200_CHECK((half_levels & ~blocked_levels) == 0, "lookup_forward: only blocked levels can be stored in fp16 (half_levels %d, blocked_levels %d)",
             half_levels, blocked_levels);
  LookupParams p;
  if (B == 0) return 0;   // empty tensors have no storage: nothing to validate, nothing to do
  if (int e = fill_params(p, h_levels, nullptr, num_levels, B, H, W, radius, mode, "lookup_forward", first_level)) return e;
  B200_CHECK(coords && out, "lookup_forward: null pointer");
  for (int l = 0; l < num_levels; ++l) {
    const uintptr_t a = (uintptr_t)p.lvl[l];
    p.path[l] = (p.LW[l] % 8 == 0 && a % 32 == 0) ? PATH_SECTOR : (p.LW[l] % 4 == 0 && a % 16 == 0) ? PATH_VEC4 : PATH_SCALAR;
    if ((blocked_levels >> l) & 1) {
      // blocked levels are the two fine levels of a pyramid built by b200corr_allpairs_pyramid_layout: padded tiles,
      // zeros in the padding columns, sector loads whatever the true width is
      B200_CHECK(first_level == 0 && l <= 1 && W % 8 == 0 && a % 32 == 0,
                 "lookup_forward: level %d (%dx%d) cannot be in the blocked layout", l, p.LH[l], p.LW[l]);
      int hp, wp;
      b200corr_blocked_level_dims(l, H, W, &hp, &wp);
      const bool hl = (half_levels >> l) & 1;
      p.blocked[l] = hl ? 2 : 1; p.path[l] = hl ? PATH_HALF : PATH_SECTOR; p.tiles_w[l] = wp / 8; p.slice[l] = (long long)hp * wp;
    }
  }
  B200_CHECK((blocked_levels >> num_levels) == 0, "lookup_forward: blocked_levels names a level that is not there");
  dim3 grid(((p.HW + QT - 1) / QT) * num_levels, 1, B);
  switch (radius) {
    case 1: lookup_fwd_kernel<1><<<grid, 32 * kFwdWarps, 0, stream>>>(p, coords, out); break;
    case 2: lookup_fwd_kernel<2><<<grid, 32 * kFwdWarps, 0, stream>>>(p, coords, out); break;
    case 3: lookup_fwd_kernel<3><<<grid, 32 * kFwdWarps, 0, stream>>>(p, coords, out); break;
    default: lookup_fwd_kernel<4><<<grid, 32 * kFwdWarps, 0, stream>>>(p, coords, out); break;
  }
  B200_LAUNCH_OK("lookup_fwd_kernel");
  return 0;
}

int b200corr_lookup_backward(float *const *h_grad_levels, int num_levels, const float *coords,
                             const float *grad_out, int B, int H, int W, int radius, int mode,
                             void *stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  LookupParams p;
  if (B == 0) return 0;
  if (int e = fill_params(p, nullptr, h_grad_levels, num_levels, B, H, W, radius, mode, "lookup_backward")) return e;
  B200_CHECK(coords && grad_out, "lookup_backward: null pointer");
  for (int l = 0; l < num_levels; ++l) {
    const uintptr_t a = (uintptr_t)p.glvl[l];
    p.path[l] = (p.LW[l] % 8 == 0 && a % 32 == 0) ? PATH_SECTOR : (p.LW[l] % 4 == 0 && a % 16 == 0) ? PATH_VEC4 : PATH_SCALAR;
  }
  dim3 grid(((p.HW + QT - 1) / QT) * num_levels, 1, B);
  switch (radius) {
    case 1: lookup_bwd_kernel<1><<<grid, 32 * kBwdWarps, 0, stream>>>(p, coords, grad_out); break;
    case 2: lookup_bwd_kernel<2><<<grid, 32 * kBwdWarps, 0, stream>>>(p, coords, grad_out); break;
    case 3: lookup_bwd_kernel<3><<<grid, 32 * kBwdWarps, 0, stream>>>(p, coords, grad_out); break;
    default: lookup_bwd_kernel<4><<<grid, 32 * kBwdWarps, 0, stream>>>(p, coords, grad_out); break;
  }
  B200_LAUNCH_OK("lookup_bwd_kernel");
  return 0;
}

int b200corr_pyramid_backward(float *const *h_grad_levels, int num_levels, int B, int H, int W,
                              void *stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  LookupParams p;
  if (B == 0) return 0;
  if (int e = fill_params(p, nullptr, h_grad_levels, num_levels, B, H, W, 1, 0, "pyramid_backward")) return e;
  const long long Q = (long long)B * H * W;
  for (int l = num_levels - 1; l >= 1; --l) {
    const long long total = Q * p.LH[l - 1] * p.LW[l - 1];
    long long blocks = (total + 255) / 256;
    const long long cap = (long long)b200::num_sms() * 16;
    if (p.LW[l - 1] % 4 == 0 && ((uintptr_t)p.glvl[l - 1] & 15) == 0) {
      const long long tot4 = total / 4;
      blocks = (tot4 + 255) / 256;
      const int nb = (int)(blocks < cap ? blocks : cap);
      if (tot4 < (1ll << 31))
        pool_bwd_vec4_kernel<unsigned><<<nb, 256, 0, stream>>>(p.glvl[l - 1], p.glvl[l], Q * p.LH[l - 1], p.LH[l - 1], p.LW[l - 1]);
      else
        pool_bwd_vec4_kernel<long long><<<nb, 256, 0, stream>>>(p.glvl[l - 1], p.glvl[l], Q * p.LH[l - 1], p.LH[l - 1], p.LW[l - 1]);
    } else {
      pool_bwd_kernel<<<(int)(blocks < cap ? blocks : cap), 256, 0, stream>>>(p.glvl[l - 1], p.glvl[l], Q,
                                                                            p.LH[l - 1], p.LW[l - 1]);
    }
    B200_LAUNCH_OK("pool_bwd_kernel");
  }
  return 0;
}

}  // extern "C"
