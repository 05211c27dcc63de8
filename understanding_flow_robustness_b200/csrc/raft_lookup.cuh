// raft_lookup.cuh -- pieces of the RAFT lookup shared by raft_lookup.cu (lookup forward / backward) and
// lookup_convc1.cu (the lookup fused with the motion encoder's 1x1 convolution): level parameters, the
// coordinate arithmetic of grid_sample reproduced operation by operation, and the window staging.
#pragma once
#include <cuda_fp16.h>
#include "common.cuh"

namespace b200lookup {

constexpr int kMaxLevels = 8;
constexpr int QT = 32;  // queries per CTA

struct LookupParams {
  const float *lvl[kMaxLevels];
  float *glvl[kMaxLevels];
  int LH[kMaxLevels], LW[kMaxLevels];
  int path[kMaxLevels];  // access flavour per level (sector / 16-byte / scalar), from width and alignment
  int num_levels, B, HW, radius, mode;
  int blocked[kMaxLevels];   // forward only: 1 = the level is stored as 8x8 tiles of 64 floats (b200corr.h), 2 = as 8x8 tiles of 64 fp16 values
  int tiles_w[kMaxLevels];   // blocked levels: tiles per row of the padded slice
  long long slice[kMaxLevels];   // floats per query slice (LH * LW, or the padded size of a blocked level)
  float inv_w[kMaxLevels], inv_h[kMaxLevels];   // RN(1 / (LW - 1)), RN(1 / (LH - 1)): the hoisted reciprocals of the coordinate round trip
  int first_level;   // pyramid level of list entry 0: entry i has extent (H, W) >> (first_level + i), coordinate scale 2^-(first_level + i)
};

template <int R>
struct Geo {
  static constexpr int N = 2 * R + 1;      // taps per axis
  static constexpr int WS = 2 * R + 4;     // window per axis: the taps + one guard row/column on either side
};

// 256-bit / 128-bit read-only loads that do not pollute L1 (every sector is used exactly once)
__device__ __forceinline__ void ldg256(const float *p, float (&v)[8]) {
  asm volatile("ld.global.nc.L1::no_allocate.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=f"(v[0]), "=f"(v[1]), "=f"(v[2]), "=f"(v[3]), "=f"(v[4]), "=f"(v[5]), "=f"(v[6]), "=f"(v[7])
               : "l"(p));
}
__device__ __forceinline__ void ldg128(const float *p, float (&v)[4]) {
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(v[0]), "=f"(v[1]), "=f"(v[2]), "=f"(v[3])
               : "l"(p));
}

// grid_sample's accumulation nw*w + ne*w + sw*w + se*w with the contraction pinned, so that the
// fast and the general sampling loop (and ATen's own FMA-contracted kernel) agree bit for bit
__device__ __forceinline__ float bilerp(float c00, float c01, float c10, float c11, float ax, float bx,
                                        float ay, float by) {
  float v = __fmul_rn(c00, __fmul_rn(bx, by));
  v = __fmaf_rn(c01, __fmul_rn(ax, by), v);
  v = __fmaf_rn(c10, __fmul_rn(bx, ay), v);
  return __fmaf_rn(c11, __fmul_rn(ax, ay), v);
}

// a / b correctly rounded from y = RN(1/b) (Markstein: q = RN(a*y), r = a - b*q exactly by FMA,
// RN(q + r*y) is the correctly rounded quotient for every b whose significand is not all ones --
// b is a small integer here).  Zero, non-finite and huge operands take the IEEE routine.
__device__ __forceinline__ float div_by(float a, float b, float y) {
  if (!(fabsf(a) < 1e30f) || !(b >= 1.f)) return __fdiv_rn(a, b);
  const float q = __fmul_rn(a, y);
  const float r = __fmaf_rn(-q, b, a);
  return __fmaf_rn(r, y, q);
}

// Pixel coordinate the reference ends up sampling at, along one axis of size sm1 + 1:
//   corr.py:84-86       centroid / 2**i + delta          (division by a power of two is exact)
//   utils.py:66         g = 2 * x / (W - 1) - 1
//   grid_sampler_unnormalize, align_corners=True:  ((g + 1) / 2) * (W - 1)
// every operation rounded separately in fp32 (no FMA contraction); the one real division uses the
// hoisted reciprocal of (size - 1): same value, a third of the instructions.
__device__ __forceinline__ float sample_coord_fast(float c, int lvl, int off, float sm1, float inv_sm1, int mode) {
  const float x = __fadd_rn(__fmul_rn(c, 1.0f / (float)(1 << lvl)), (float)off);
  if (mode == B200CORR_LOOKUP_DIRECT) return x;
  const float g = __fsub_rn(div_by(__fmul_rn(2.0f, x), sm1, inv_sm1), 1.0f);
  return __fmul_rn(__fmul_rn(__fadd_rn(g, 1.0f), 0.5f), sm1);   // x / 2 == x * 0.5 exactly
}

// one tap of one axis: position relative to the window origin `org` (or -1) and fraction
template <int R>
__device__ __forceinline__ void one_tap(float c, int lvl, int t, float sm1, float inv_sm1, int mode, int org,
                                        int &rel, float &frac) {
  constexpr int WS = 2 * R + 4;
  const float x = sample_coord_fast(c, lvl, t - R, sm1, inv_sm1, mode);
  const float fx = floorf(x);
  if (fabsf(fx) < 1e8f) {
    rel = (int)fx - org;
    frac = x - fx;
    // the round trip moves x by a few ulp at most: rel is within [0, WS-2]; clamp defensively
    if (rel < 0 || rel > WS - 2) { rel = -1; frac = 0.f; }
  } else {
    rel = 0; frac = x - x;  // NaN propagates like in the reference
  }
}

// The x tap and the y tap of index t in one pass: the coordinate arithmetic above on packed fp32 pairs
// (add / mul / fma .f32x2, sm_100: each half rounded exactly like the scalar instruction, so the values are the
// scalar routine's bit for bit).  The lookup kernels are bound by instruction issue, not by their gathers (with the
// window loads switched off a lookup takes the same 31 us), and the tap tables were a quarter of all instructions.
template <int R>
__device__ __forceinline__ void two_taps(float cx, float cy, int lvl, int t, float smx, float smy, float ismx, float ismy,
                                         int mode, int ox, int oy, int &relx, float &fracx, int &rely, float &fracy) {
  constexpr int WS = 2 * R + 4;
  const float sc = 1.0f / (float)(1 << lvl), off = (float)(t - R);
  float2 x = __fadd2_rn(__fmul2_rn(make_float2(cx, cy), make_float2(sc, sc)), make_float2(off, off));
  if (mode != B200CORR_LOOKUP_DIRECT) {
    const float2 a = __fmul2_rn(make_float2(2.0f, 2.0f), x);
    const float2 y = make_float2(ismx, ismy);
    const float2 q = __fmul2_rn(a, y);
    const float2 r = __ffma2_rn(q, make_float2(-smx, -smy), a);     // a - b q, exact
    float2 d = __ffma2_rn(r, y, q);
    // zero, non-finite and huge operands take the IEEE routine (div_by)
    if (!(fabsf(a.x) < 1e30f) || !(smx >= 1.f)) d.x = __fdiv_rn(a.x, smx);
    if (!(fabsf(a.y) < 1e30f) || !(smy >= 1.f)) d.y = __fdiv_rn(a.y, smy);
    const float2 g = __fadd2_rn(d, make_float2(-1.0f, -1.0f));
    x = __fmul2_rn(__fmul2_rn(__fadd2_rn(g, make_float2(1.0f, 1.0f)), make_float2(0.5f, 0.5f)), make_float2(smx, smy));
  }
  const float fx = floorf(x.x), fy = floorf(x.y);
  if (fabsf(fx) < 1e8f) {
    relx = (int)fx - ox;
    fracx = x.x - fx;
    if (relx < 0 || relx > WS - 2) { relx = -1; fracx = 0.f; }
  } else {
    relx = 0; fracx = x.x - x.x;
  }
  if (fabsf(fy) < 1e8f) {
    rely = (int)fy - oy;
    fracy = x.y - fy;
    if (rely < 0 || rely > WS - 2) { rely = -1; fracy = 0.f; }
  } else {
    rely = 0; fracy = x.y - x.y;
  }
}

// Window origin floor(c / 2^l) - R - 1 from the un-rounded centre, and the window rows/columns
// [lo, hi] the taps can touch: the coordinate arithmetic (fp32 add, and in grid_sample mode the
// normalise / un-normalise round trip) moves a sample position by < 1e-3 pixel for |c| < 1024, so
// unless the centre sits within 1/64 pixel of an integer, tap t lands exactly on window position
// t + 1 and the outermost row/column on either side is never read.
template <int R>
__device__ __forceinline__ int window_origin(float c, int lvl, int &lo, int &hi) {
  const float cl = c * (1.0f / (float)(1 << lvl));
  float fo = floorf(cl);
  const float f = cl - fo;
  const bool interior = fabsf(cl) < 1024.f && f > 0.015625f && f < 0.984375f;
  lo = interior ? 1 : 0;
  hi = interior ? 2 * R + 2 : 2 * R + 3;
  if (!(fabsf(fo) < 1e8f)) fo = -1e8f;  // non-finite / absurd coordinates: everything out of range
  return (int)fo - R - 1;
}

constexpr int kCols = 16;  // staged columns per window row: (ox & 3) + 2r + 4 <= 15 for r <= 4
enum { PATH_SCALAR = 0, PATH_VEC4 = 1, PATH_SECTOR = 2, PATH_HALF = 3 };   // PATH_HALF: blocked fp16 tiles, column geometry of PATH_SECTOR

struct StageArgs {
  const float *slice;  // this lane's H_l x W_l slice
  int oy, ox, LH, LW;  // window origin, level extent
  int ylo, yhi;        // window rows the taps touch
  int clo, chi;        // staged columns the taps touch
  bool q_ok;
  bool blocked;        // PATH_SECTOR / PATH_HALF: the slice is a grid of 8x8 tiles (64 consecutive values each)
  int tiles_w;         // tiles per row of a blocked slice
};

// Stage window rows r0 .. r0+NR-1 (those below WS) of this lane's query: win[row][column][lane].
// Staged column 0 is level column (ox & ~3) for the vector flavours, ox itself for the scalar one.
// Only the rows and the 32-byte sectors the taps touch are fetched; everything outside the slice is
// staged as zero (= grid_sample's zero padding).
// Two phases so that the tap tables (ALU work that needs only the coordinates) run while the loads are in flight:
// stage_load issues every global load of this warp's rows into registers, stage_store parks them in the tile.
template <int PATH, int NR, int WS>
__device__ __forceinline__ void stage_load(const StageArgs &a, int r0, float (&v)[NR][24]) {
  constexpr int NL = PATH == PATH_SECTOR ? 24 : PATH == PATH_VEC4 ? 16 : 12;   // loaded
  // first loaded column: sector aligned / 16-byte aligned / exact
  const int c0 = PATH == PATH_SECTOR ? (a.ox & ~7) : PATH == PATH_VEC4 ? (a.ox & ~3) : a.ox;
  const bool hi4 = (a.ox & 4) != 0;  // PATH_SECTOR: staged column 0 is loaded column 4
  // columns the taps touch, in loaded-column units
  const int llo = PATH == PATH_SECTOR && hi4 ? a.clo + 4 : a.clo, lhi = PATH == PATH_SECTOR && hi4 ? a.chi + 4 : a.chi;
#pragma unroll
  for (int r = 0; r < NR; ++r) {
    const int wr = r0 + r, y = a.oy + wr;
    const bool row_ok = a.q_ok && wr >= a.ylo && wr <= a.yhi && y >= 0 && y < a.LH;
    const float *src = a.slice + (size_t)(row_ok ? y : 0) * a.LW + c0;
#pragma unroll
    for (int k = 0; k < NL; ++k) v[r][k] = 0.f;
    if (PATH == PATH_SECTOR) {
#pragma unroll
      for (int g = 0; g < 3; ++g) {   // one 32-byte sector per load
        const int x = c0 + 8 * g;
        if (row_ok && x >= 0 && x < a.LW && lhi >= 8 * g && llo < 8 * g + 8) {
          const float *sp = a.blocked ? a.slice + ((size_t)((y >> 3) * a.tiles_w + (x >> 3)) * 64 + (y & 7) * 8)
                                      : src + 8 * g;
          ldg256(sp, *reinterpret_cast<float(*)[8]>(&v[r][8 * g]));
        }
      }
    } else if (PATH == PATH_VEC4) {
#pragma unroll
      for (int g = 0; g < 4; ++g) {
        const int x = c0 + 4 * g;
        if (row_ok && x >= 0 && x < a.LW && lhi >= 4 * g && llo < 4 * g + 4)
          ldg128(src + 4 * g, *reinterpret_cast<float(*)[4]>(&v[r][4 * g]));
      }
    } else {
#pragma unroll
      for (int k = 0; k < NL; ++k)
        if (row_ok && c0 + k >= 0 && c0 + k < a.LW) v[r][k] = __ldg(src + k);
    }
  }
}

template <int PATH, int NR, int WS>
__device__ __forceinline__ void stage_store(float *win, int lane, const StageArgs &a, int r0, const float (&v)[NR][24]) {
  constexpr int NC = PATH == PATH_SCALAR ? 12 : 16;                            // staged
  const bool hi4 = (a.ox & 4) != 0;
#pragma unroll
  for (int r = 0; r < NR; ++r) {
    float *dst = win + ((r0 + r) * kCols) * 32 + lane;
    if (r0 + r < WS) {
#pragma unroll
      for (int k = 0; k < NC; ++k) dst[k * 32] = (PATH == PATH_SECTOR) ? (hi4 ? v[r][k + 4] : v[r][k]) : v[r][k];
    }
  }
}


// ---- PATH_HALF: blocked fp16 tiles (a.slice addresses halves; the slice sizes count elements).  A 32-byte sector of
// such a tile holds TWO rows (y even, y + 1) of 8 columns, so the window is fetched by aligned row PAIRS: 5-6 pairs
// x 2-3 tile columns = ~12 sector loads per window instead of ~21.  Pair j of the window starts at row
// (oy & ~1) + 2 j, j = 0..6; warp w fetches pairs w and w + 4.  hv[jj][8 g + 4 rr + e]: packed halves 2e, 2e+1 of
// row rr of tile column g; everything stays packed until stage_store_half, so all loads of a warp are in flight
// together (a conversion right behind each load serialised them: 48.6 instead of 29.4 us per lookup).
template <int WS, int NR, int NW = 4>
__device__ __forceinline__ void stage_load_half(const StageArgs &a, int warp, float (&hv)[NR][24]) {
  static_assert(NR >= 2 && WS / 2 + 1 <= 2 * NW, "two row pairs per warp cover the window");
  const int c0 = a.ox & ~7, ybase = a.oy & ~1;
  const bool hi4 = (a.ox & 4) != 0;
  const int llo = hi4 ? a.clo + 4 : a.clo, lhi = hi4 ? a.chi + 4 : a.chi;
#pragma unroll
  for (int jj = 0; jj < 2; ++jj) {
    const int y0 = ybase + 2 * (warp + NW * jj);
#pragma unroll
    for (int k = 0; k < 24; ++k) hv[jj][k] = 0.f;
    bool need = false;
#pragma unroll
    for (int rr = 0; rr < 2; ++rr) {
      const int y = y0 + rr, wr = y - a.oy;
      need = need || (wr >= a.ylo && wr <= a.yhi && wr < WS && y >= 0 && y < a.LH);
    }
    need = need && a.q_ok;
#pragma unroll
    for (int g = 0; g < 3; ++g) {
      const int x = c0 + 8 * g;
      if (need && x >= 0 && x < a.LW && lhi >= 8 * g && llo < 8 * g + 8) {
        const __half *hp = reinterpret_cast<const __half *>(a.slice) +
                           ((size_t)((y0 >> 3) * a.tiles_w + (x >> 3)) * 64 + (y0 & 7) * 8);
        ldg256(reinterpret_cast<const float *>(hp), *reinterpret_cast<float(*)[8]>(&hv[jj][8 * g]));
      }
    }
  }
}

template <int WS, int NR, int NW = 4>
__device__ __forceinline__ void stage_store_half(float *win, int lane, const StageArgs &a, int warp, const float (&hv)[NR][24]) {
  const bool hi4 = (a.ox & 4) != 0;
  const int ybase = a.oy & ~1;
#pragma unroll
  for (int jj = 0; jj < 2; ++jj) {
#pragma unroll
    for (int rr = 0; rr < 2; ++rr) {
      const int y = ybase + 2 * (warp + NW * jj) + rr, wr = y - a.oy;
      if (wr < 0 || wr >= WS) continue;
      // rows the taps do not touch / outside the level are staged as zero (the tile rows past the extent of a
      // padded level are not zero in memory)
      const bool row_ok = a.q_ok && wr >= a.ylo && wr <= a.yhi && y >= 0 && y < a.LH;
      float *dst = win + (wr * kCols) * 32 + lane;
#pragma unroll
      for (int k = 0; k < 16; ++k) {
        // loaded column idx = k (+ 4 when the window starts in the upper half of a tile row): half idx & 1 of
        // register 8 (idx >> 3) + 4 rr + ((idx & 7) >> 1); idx and idx + 4 share the half position
        const float ra = hv[jj][8 * (k >> 3) + 4 * rr + ((k & 7) >> 1)];
        const float rb = hv[jj][8 * ((k + 4) >> 3) + 4 * rr + (((k + 4) & 7) >> 1)];
        const __half2 h2 = *reinterpret_cast<const __half2 *>(hi4 ? &rb : &ra);
        const float f = (k & 1) ? __high2float(h2) : __low2float(h2);
        dst[k * 32] = row_ok ? f : 0.f;
      }
    }
  }
}

inline int fill_params(LookupParams &p, const float *const *lv, float *const *glv, int num_levels, int B,
                int H, int W, int radius, int mode, const char *who, int first_level = 0) {
  B200_CHECK(num_levels >= 1 && num_levels <= kMaxLevels, "%s: num_levels must be in [1, %d]", who,
             kMaxLevels);
  B200_CHECK(radius >= 1 && radius <= 4, "%s: radius %d not instantiated (1..4)", who, radius);
  B200_CHECK(mode == B200CORR_LOOKUP_GRIDSAMPLE || mode == B200CORR_LOOKUP_DIRECT, "%s: bad mode", who);
  B200_CHECK(B >= 0 && H >= 1 && W >= 1, "%s: bad sizes", who);
  B200_CHECK(first_level >= 0 && first_level + num_levels <= 16, "%s: bad first_level %d", who, first_level);
  p.num_levels = num_levels; p.B = B; p.HW = H * W; p.radius = radius; p.mode = mode; p.first_level = first_level;
  int h = H >> first_level, w = W >> first_level;
  for (int l = 0; l < kMaxLevels; ++l) {
    p.inv_w[l] = 0.f; p.inv_h[l] = 0.f;
    p.lvl[l] = nullptr; p.glvl[l] = nullptr; p.LH[l] = 0; p.LW[l] = 0; p.path[l] = 0; p.blocked[l] = 0; p.tiles_w[l] = 0; p.slice[l] = 0;
    if (l < num_levels) {
      p.LH[l] = h; p.LW[l] = w; p.slice[l] = (long long)h * w;
      p.inv_w[l] = 1.0f / (float)(w - 1); p.inv_h[l] = 1.0f / (float)(h - 1);   // IEEE division = __frcp_rn (inf for a 1-pixel axis)
      B200_CHECK(h >= 1 && w >= 1, "%s: pyramid level %d is empty (%dx%d input)", who, l, H, W);
      if (lv) { B200_CHECK(lv[l], "%s: null level %d", who, l); p.lvl[l] = lv[l]; }
      if (glv) { B200_CHECK(glv[l], "%s: null gradient level %d", who, l); p.glvl[l] = glv[l]; }
      h /= 2; w /= 2;
    }
  }
  return 0;
}


}  // namespace b200lookup
