// sampler_fast_fwd.cu -- register-blocked, TMA-staged forward of the spatial correlation sampler
// for the FlowNetC / PWC-Net call-site structure (kernel_size 1, stride 1, padding 0):
//
//     out[n, ph, pw, h, w] = sum_c in1[n, c, h, w] * in2[n, c, h + (ph-RH)*dpH, w + (pw-RW)*dpW]
//
// Replaces correlation_cuda_forward + correlation_cuda_forward_kernel of the reference
// (correlation_cuda_kernel.cu:22-83, 236-273: one 32-thread block per output pixel, 441 serial
// displacements, thread-0 serial reduction, NHWC permute copies of both inputs).
//
// Design (DESIGN.md section "sampler forward"):
//   * Row displacements are multiples of dpH, so rows split into dpH independent "parity classes";
//     inside a class every displacement is a unit step on the sub-lattice of rows.  TMA loads only
//     the rows of one class (tensor-map element stride dpH on the row axis) straight from NCHW --
//     no layout copy -- and zero-fills everything outside the image, which IS the reference's
//     WITHIN_BOUNDS rule (correlation.cpp:24,28) for this structure.
//   * Work item = (pixel row s, column group g of 8 pixels, row displacement e).  One thread owns
//     one item and keeps its 8 x PW accumulators (168 for FlowNetC) in registers for the whole
//     channel loop: per channel it needs 8 in1 values and a window of 8+(PW-1)*dpW in2 values,
//     i.e. 14 LDS.128 for 168 FFMA -- the Toeplitz reuse that lifts the loop off the
//     shared-memory roof onto the FP32 pipe.
//   * A warp is 4 pixel rows x 8 items: the 4 lanes of an item column share in2 addresses only
//     through the hardware multicast of identical/neighbouring words and the 8 lanes of a pixel row
//     share the in1 address, so one warp-wide LDS.128 touches few distinct 16-byte chunks.
//   * Row displacements that fall outside the image for every row of a group are never computed
//     (their output planes are zero-filled by the same threads), so the work tracks the in-bounds
//     MAC count rather than the dense one.
//   * Persistent CTAs (one per SM, 256 threads, <=255 registers), static round-robin over units; a
//     3-stage mbarrier ring over 8-channel chunks that keeps running across unit boundaries.
#include "sampler_fast.cuh"

// tuning knobs (scripts/tune_sampler.py builds variants with -D...)
#ifndef B200_FWD_CC
#define B200_FWD_CC 8
#endif
#ifndef B200_FWD_NST
#define B200_FWD_NST 3
#endif
#ifndef B200_FWD_UNROLL
#define B200_FWD_UNROLL 2
#endif

namespace {
using namespace b200dev;

constexpr int kFwdUnroll = B200_FWD_UNROLL;

constexpr int odd4(int x) {  // round up to a multiple of 4 floats whose 16-byte chunk count is odd
  int c = (x + 3) / 4;
  if (c % 2 == 0) ++c;
  return 4 * c;
}

template <int PH_, int PW_, int DPW_>
struct FwdCfg {
  static constexpr int PH = PH_, PW = PW_, DPW = DPW_;
  static constexpr int T = 8;                          // pixels per thread
  static constexpr int RH = (PH - 1) / 2, RW = (PW - 1) / 2;
  static constexpr int HALO = RW * DPW;
  static constexpr int WIN = T + (PW - 1) * DPW;       // in2 window of one thread, floats
  static_assert(WIN % 4 == 0, "window must be a whole number of 16-byte chunks");
  static constexpr int GMAX = 5;                       // column groups one unit may span
  static constexpr int NRB = b200::kRowsPerGroup + PH - 1;  // in2 rows staged per unit
  static constexpr int NC1 = odd4(GMAX * T);
  static constexpr int NC2 = odd4(GMAX * T + (PW - 1) * DPW);
  static constexpr int CC = B200_FWD_CC;               // channels per pipeline stage
  static constexpr int NST = B200_FWD_NST;             // pipeline stages
  static constexpr int IN2_FLOATS = CC * NRB * NC2;
  static constexpr int IN1_FLOATS = CC * b200::kRowsPerGroup * NC1;
  static constexpr int STAGE_FLOATS = IN2_FLOATS + IN1_FLOATS;
  static constexpr int STAGE_BYTES = STAGE_FLOATS * 4;
  static_assert((IN2_FLOATS * 4) % 128 == 0 && STAGE_BYTES % 128 == 0, "TMA destination alignment");
  static constexpr int SMEM_BYTES = NST * STAGE_BYTES + 256;
  static constexpr int SLOTS = 64;                     // items per unit (8 warps x 8)
};

struct FwdParams {
  int B, C, H, W, dpH, NG, total_units;
  int wide_store;   // W % 8 == 0 and a 32-byte aligned output: one STG.256 per 8-pixel run
  // Output epilogue: out[n] starts out_bstride floats after out[n-1]; the sums are stored as
  // leaky_relu(sum * post_scale, post_slope).  Plain sampler: P*P*H*W, 1, 1 (x*1 is exact, so the values are
  // the sums); fused FlowNetC merge block: batch stride of the concat tensor, 1/C, the LeakyReLU slope.
  // (The always-on transform also happens to steer ptxas to a spill-free 245-register allocation of the
  // channel loop: 0.285 -> 0.273 ms at (8,256,48,160).)
  long long out_bstride;
  float post_scale, post_slope;
  b200::SamplerGroups g;
  int ctail;        // C % CC != 0: the maps are 4-D (W, H, C, B) so that TMA zero-fills the channels past C
};

// in2 sub-row range [Rlo, Rlo + nR) a row group needs, and its split into units of <= SLOTS items
// (item = (column group, in2 sub-row)); same arithmetic on host and device.
__host__ __device__ inline void fwd_group_geom(int NS, int s0, int RH, int NG, int GMAX, int SLOTS,
                                               int &Rlo, int &nR, int &nch, int &cs) {
  const int s_last = (s0 + b200::kRowsPerGroup - 1 < NS - 1) ? s0 + b200::kRowsPerGroup - 1 : NS - 1;
  Rlo = s0 - RH > 0 ? s0 - RH : 0;
  const int Rhi = s_last + RH < NS - 1 ? s_last + RH : NS - 1;
  nR = Rhi - Rlo + 1;
  const int nitems = NG * nR;
  int cap = (GMAX - 1) * nR;
  if (cap > SLOTS) cap = SLOTS;
  nch = (nitems + cap - 1) / cap;
  cs = (nitems + nch - 1) / nch;
}

struct Unit {
  int n, rp, s0, NS, Rlo, nR, item0, cnt, g0;
};

template <class Cfg>
__device__ __forceinline__ void decode_unit(const FwdParams &p, int u, Unit &x) {
  x.n = u / p.g.units_per_sample;
  const int r = u - x.n * p.g.units_per_sample;
  int gi = 0;
  while (gi + 1 < p.g.ngroups && p.g.prefix[gi + 1] <= r) ++gi;
  x.rp = p.g.rp[gi];
  x.s0 = p.g.s0[gi];
  x.NS = (p.H - x.rp + p.dpH - 1) / p.dpH;
  int nch, cs;
  fwd_group_geom(x.NS, x.s0, Cfg::RH, p.NG, Cfg::GMAX, Cfg::SLOTS, x.Rlo, x.nR, nch, cs);
  const int k = r - p.g.prefix[gi];
  x.item0 = k * cs;
  const int nitems = p.NG * x.nR;
  x.cnt = nitems - x.item0 < cs ? nitems - x.item0 : cs;
  x.g0 = x.item0 / x.nR;
}

__device__ __forceinline__ void stg128(float *p, float a, float b, float c, float d) {
  *reinterpret_cast<float4 *>(p) = make_float4(a, b, c, d);
}
// 256-bit global store (sm_100: STG.E.256), p must be 32-byte aligned
__device__ __forceinline__ void stg256(float *p, float2 a, float2 b, float2 c, float2 d) {
  asm volatile("st.global.v8.f32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(p), "f"(a.x), "f"(a.y), "f"(b.x),
               "f"(b.y), "f"(c.x), "f"(c.y), "f"(d.x), "f"(d.y)
               : "memory");
}

// Fused merge block: out is the channel slice [c_off, c_off + PH*PW) of a (B, c_total, H, W) tensor (batch
// stride p.out_bstride) and receives leaky_relu(sum * (1/C), slope) -- submodules.py:124-138 + FlowNetC.py:138,147.
// CTAIL: C % CC != 0, 4-D tensor maps (a compile-time switch: the extra branch in the producer costs the
// non-tail kernel its spill-free register allocation, 0.273 -> 0.277 ms).
template <class Cfg, bool CTAIL>
__global__ void __launch_bounds__(256, 1)
sampler_fwd_kernel(const __grid_constant__ CUtensorMap map1, const __grid_constant__ CUtensorMap map2,
                   float *__restrict__ out, const FwdParams p) {
  constexpr int PH = Cfg::PH, PW = Cfg::PW, DPW = Cfg::DPW, CC = Cfg::CC, NST = Cfg::NST;
  constexpr int NC1 = Cfg::NC1, NC2 = Cfg::NC2, NRB = Cfg::NRB;
  extern __shared__ __align__(128) float smem[];  // NST stages, then the mbarriers
  uint64_t *full_bar = reinterpret_cast<uint64_t *>(smem + NST * Cfg::STAGE_FLOATS);
  uint64_t *empty_bar = full_bar + NST;
  Unit &px = *reinterpret_cast<Unit *>(empty_bar + NST);  // producer-side unit (thread 0 only)

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  // Lane layout.  Measured on B200 (scripts/probe_sm.py + ncu wavefront counters): an LDS.128
  // returns 8 sixteen-byte slots per wavefront and only an aligned lane PAIR (2i, 2i+1) that reads
  // the same 16 bytes shares a slot; sharing between other lanes is not merged.  So the 4 pixel-row
  // lanes of an item are adjacent (li fastest) and all read the SAME in2 row: an item is
  // (column group, in2 sub-row R) and lane li handles the displacement e = R - s(li).  One in2
  // LDS.128 then costs 2 wavefronts instead of 4.
  const int li = lane & 3;   // pixel row inside the group (0..3)
  const int lj = lane >> 2;  // item inside the warp (0..7)

  if (tid == 0) {
    tma_prefetch_desc(&map1);
    tma_prefetch_desc(&map2);
    for (int s = 0; s < NST; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 8);
    }
    fence_barrier_init();
  }
  __syncthreads();

  const int nunits = p.total_units;
  const int cpu = CTAIL ? (p.C + CC - 1) / CC : p.C / CC;  // channel chunks per unit (CTAIL: the last one zero-filled past C)

  // ---- producer state (thread 0 only): the load stream runs NST-1 chunks ahead of the math
  int pu = blockIdx.x, pc = 0;
  uint32_t pq = 0;
  if (tid == 0 && pu < nunits) decode_unit<Cfg>(p, pu, px);
  auto issue = [&]() {
    if (pu >= nunits) return;
    const int st = pq % NST;
    float *dst = smem + st * Cfg::STAGE_FLOATS;
    mbar_arrive_expect_tx(&full_bar[st], Cfg::STAGE_BYTES);
    if constexpr (CTAIL) {
      tma_load_4d(dst, &map2, &full_bar[st], px.g0 * Cfg::T - Cfg::HALO, px.Rlo * p.dpH + px.rp, pc * CC, px.n);
      tma_load_4d(dst + Cfg::IN2_FLOATS, &map1, &full_bar[st], px.g0 * Cfg::T, px.s0 * p.dpH + px.rp, pc * CC,
                  px.n);
    } else {
      const int ch = px.n * p.C + pc * CC;
      tma_load_3d(dst, &map2, &full_bar[st], px.g0 * Cfg::T - Cfg::HALO,
                  px.Rlo * p.dpH + px.rp, ch);
      tma_load_3d(dst + Cfg::IN2_FLOATS, &map1, &full_bar[st], px.g0 * Cfg::T,
                  px.s0 * p.dpH + px.rp, ch);
    }
    ++pq;
    if (++pc == cpu) {
      pc = 0;
      pu += gridDim.x;
      if (pu < nunits) decode_unit<Cfg>(p, pu, px);
    }
  };
  if (tid == 0)
    for (int s = 0; s < NST - 1; ++s) issue();

  uint32_t q = 0;  // chunks consumed so far by this CTA
  for (int u = blockIdx.x; u < nunits; u += gridDim.x) {
    Unit x;
    decode_unit<Cfg>(p, u, x);
    const int slot = warp * 8 + lj;
    const bool active = slot < x.cnt;
    const int item = x.item0 + (active ? slot : 0);
    const int g = item / x.nR;
    const int rbox = item - g * x.nR;          // in2 row inside the staged box
    const int s = x.s0 + li;
    const int e = x.Rlo + rbox - s;            // row displacement this lane computes
    const bool valid = active && s < x.NS && e >= -Cfg::RH && e <= Cfg::RH;
    const int cb = (g - x.g0) * Cfg::T;        // column offset inside both boxes
    const int off2 = rbox * NC2 + cb;
    const int off1 = Cfg::IN2_FLOATS + li * NC1 + cb;

    // accumulators as pixel pairs: acc2[tp][k] = (acc[2tp][k], acc[2tp+1][k]) -> packed FFMA2
    float2 acc2[Cfg::T / 2][PW];
#pragma unroll
    for (int t = 0; t < Cfg::T / 2; ++t)
#pragma unroll
      for (int k = 0; k < PW; ++k) acc2[t][k] = make_float2(0.f, 0.f);

    for (int c = 0; c < cpu; ++c, ++q) {
      const int st = q % NST;
      if (tid == 0) {
        if (q > 0) mbar_wait(&empty_bar[(q - 1) % NST], ((q - 1) / NST) & 1);
        issue();
      }
      mbar_wait(&full_bar[st], (q / NST) & 1);
      const float *sb = smem + st * Cfg::STAGE_FLOATS + off2;
      const float *sa = smem + st * Cfg::STAGE_FLOATS + off1;
#pragma unroll kFwdUnroll
      for (int cc = 0; cc < CC; ++cc) {
        const float4 a0 = lds128(sa + cc * b200::kRowsPerGroup * NC1);
        const float4 a1 = lds128(sa + cc * b200::kRowsPerGroup * NC1 + 4);
        if constexpr (DPW % 2 == 0) {
          // pixel pair tp = (2tp, 2tp+1) meets window pair tp + k*DPW/2: one FFMA2 per (tp, k)
          const float2 ap[4] = {make_float2(a0.x, a0.y), make_float2(a0.z, a0.w),
                                make_float2(a1.x, a1.y), make_float2(a1.z, a1.w)};
#pragma unroll
          for (int sgm = 0; sgm < Cfg::WIN / 4; ++sgm) {
            const float4 v4 = lds128(sb + cc * NRB * NC2 + 4 * sgm);
            const float2 vp[2] = {make_float2(v4.x, v4.y), make_float2(v4.z, v4.w)};
#pragma unroll
            for (int uu = 0; uu < 2; ++uu) {
#pragma unroll
              for (int tp = 0; tp < 4; ++tp) {
                const int d = 2 * sgm + uu - tp;  // = k * DPW / 2
                if (d >= 0 && d % (DPW / 2) == 0 && d / (DPW / 2) < PW)
                  acc2[tp][d / (DPW / 2)] = __ffma2_rn(ap[tp], vp[uu], acc2[tp][d / (DPW / 2)]);
              }
            }
          }
        } else {
          const float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
#pragma unroll
          for (int sgm = 0; sgm < Cfg::WIN / 4; ++sgm) {
            const float4 v4 = lds128(sb + cc * NRB * NC2 + 4 * sgm);
            const float v[4] = {v4.x, v4.y, v4.z, v4.w};
#pragma unroll
            for (int uu = 0; uu < 4; ++uu) {
#pragma unroll
              for (int t = 0; t < Cfg::T; ++t) {
                const int d = 4 * sgm + uu - t;  // = k * DPW
                if (d >= 0 && d % DPW == 0 && d / DPW < PW) {
                  float2 &r = acc2[t / 2][d / DPW];
                  if (t % 2 == 0) r.x = fmaf(a[t], v[uu], r.x);
                  else r.y = fmaf(a[t], v[uu], r.y);
                }
              }
            }
          }
        }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&empty_bar[st]);
    }

    // ---- epilogue: every (n, ph, pw, h, w) of this unit's rows/columns is written exactly once
    if (valid) {
      const size_t HW = (size_t)p.H * p.W;
      const int h = s * p.dpH + x.rp;
      const int w0 = g * Cfg::T;
      const bool lo_ok = w0 < p.W, hi_ok = w0 + 4 < p.W;
      float *obase = out + (size_t)x.n * (size_t)p.out_bstride;   // plane (n, ph = 0, pw = 0) of this sample
      {
        const float ic = p.post_scale, sl = p.post_slope;
#pragma unroll
        for (int t = 0; t < Cfg::T / 2; ++t)
#pragma unroll
          for (int k = 0; k < PW; ++k) {
            float vx = acc2[t][k].x * ic, vy = acc2[t][k].y * ic;
            acc2[t][k].x = vx > 0.f ? vx : vx * sl;
            acc2[t][k].y = vy > 0.f ? vy : vy * sl;
          }
      }
      float *o = obase + ((size_t)(e + Cfg::RH) * PW) * HW + (size_t)h * p.W + w0;
      // one 32-byte sector per (displacement, pixel row): a single 256-bit store when rows are 32-B aligned
      const bool wide = p.wide_store && hi_ok;
      if (wide) {
#pragma unroll
        for (int k = 0; k < PW; ++k) stg256(o + k * HW, acc2[0][k], acc2[1][k], acc2[2][k], acc2[3][k]);
      } else {
#pragma unroll
        for (int k = 0; k < PW; ++k) {
          if (lo_ok) stg128(o + k * HW, acc2[0][k].x, acc2[0][k].y, acc2[1][k].x, acc2[1][k].y);
          if (hi_ok) stg128(o + k * HW + 4, acc2[2][k].x, acc2[2][k].y, acc2[3][k].x, acc2[3][k].y);
        }
      }
      // displacement planes whose in2 row lies outside the image are zero; the valid lanes of this
      // pixel row share them out by their rank among the valid rows
      const int below = Cfg::RH - s > 0 ? Cfg::RH - s : 0;                      // planes [0, below)
      const int above = Cfg::RH - (x.NS - 1 - s) > 0 ? Cfg::RH - (x.NS - 1 - s) : 0;  // [PH-above, PH)
      const int nvalid = PH - below - above;
      for (int qq = e + Cfg::RH - below; qq < below + above; qq += nvalid) {
        const int phc = qq < below ? qq : PH - above + (qq - below);
        float *z = obase + ((size_t)phc * PW) * HW + (size_t)h * p.W + w0;
        const float2 zz = make_float2(0.f, 0.f);
        for (int k = 0; k < PW; ++k) {
          if (wide) {
            stg256(z + k * HW, zz, zz, zz, zz);
          } else {
            if (lo_ok) stg128(z + k * HW, 0.f, 0.f, 0.f, 0.f);
            if (hi_ok) stg128(z + k * HW + 4, 0.f, 0.f, 0.f, 0.f);
          }
        }
      }
    }
  }
}

template <class Cfg>
int launch_fwd(const float *in1, const float *in2, float *out, int B, int C, int H, int W, int dpH,
               long long out_bstride, float post_scale, float post_slope, cudaStream_t stream) {
  FwdParams p;
  p.B = B; p.C = C; p.H = H; p.W = W; p.dpH = dpH;
  p.NG = (W + Cfg::T - 1) / Cfg::T;
  p.out_bstride = out_bstride; p.post_scale = post_scale; p.post_slope = post_slope;
  p.wide_store = (W % 8 == 0 && ((uintptr_t)out & 31) == 0 && out_bstride % 8 == 0) ? 1 : 0;
  int ng = 0, units = 0;
  for (int rp = 0; rp < dpH; ++rp) {
    const int NS = b200::sublattice_rows(H, dpH, rp);
    for (int s0 = 0; s0 < NS; s0 += b200::kRowsPerGroup) {
      B200_CHECK(ng < b200::kSamplerMaxGroups, "sampler_fast_forward: too many row groups");
      int Rlo, nR, nch, cs;
      fwd_group_geom(NS, s0, Cfg::RH, p.NG, Cfg::GMAX, Cfg::SLOTS, Rlo, nR, nch, cs);
      p.g.prefix[ng] = units;
      p.g.rp[ng] = (short)rp;
      p.g.s0[ng] = (short)s0;
      units += nch;
      ++ng;
    }
  }
  p.g.prefix[ng] = units;
  p.g.ngroups = ng;
  p.g.units_per_sample = units;
  p.total_units = units * B;
  if (p.total_units == 0) return 0;

  CUtensorMap map1, map2;
  // (W, H, B*C) when the channel chunks tile C exactly; (W, H, C, B) otherwise (PWC-Net's 196-channel level):
  // the last chunk of a sample then reads zeros past C instead of the next sample's channels
  p.ctail = (C % Cfg::CC != 0) ? 1 : 0;
  const int rank = p.ctail ? 4 : 3;
  const uint64_t dims[4] = {(uint64_t)W, (uint64_t)H, p.ctail ? (uint64_t)C : (uint64_t)B * C, (uint64_t)B};
  const uint64_t strides[4] = {4, (uint64_t)W * 4, (uint64_t)H * W * 4, (uint64_t)H * W * 4 * C};
  const uint32_t estr[4] = {1, (uint32_t)dpH, 1, 1};
  const uint32_t box2[4] = {(uint32_t)Cfg::NC2, (uint32_t)((Cfg::NRB - 1) * dpH + 1), (uint32_t)Cfg::CC, 1};
  const uint32_t box1[4] = {(uint32_t)Cfg::NC1, (uint32_t)((b200::kRowsPerGroup - 1) * dpH + 1),
                            (uint32_t)Cfg::CC, 1};
  if (int e = b200::make_tensor_map(&map2, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, rank, in2, dims, strides,
                                    box2, CU_TENSOR_MAP_SWIZZLE_NONE,
                                    CU_TENSOR_MAP_L2_PROMOTION_L2_128B, estr))
    return e;
  if (int e = b200::make_tensor_map(&map1, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, rank, in1, dims, strides,
                                    box1, CU_TENSOR_MAP_SWIZZLE_NONE,
                                    CU_TENSOR_MAP_L2_PROMOTION_L2_128B, estr))
    return e;

  const int grid = p.total_units < b200::num_sms() ? p.total_units : b200::num_sms();
  if (p.ctail) {
    auto kern = sampler_fwd_kernel<Cfg, true>;
    static bool attr_done[64] = {};  // per (kernel instantiation, device)
    if (int e = b200::set_max_smem_once((const void *)kern, Cfg::SMEM_BYTES, attr_done)) return e;
    kern<<<grid, 256, Cfg::SMEM_BYTES, stream>>>(map1, map2, out, p);
  } else {
    auto kern = sampler_fwd_kernel<Cfg, false>;
    static bool attr_done[64] = {};
    if (int e = b200::set_max_smem_once((const void *)kern, Cfg::SMEM_BYTES, attr_done)) return e;
    kern<<<grid, 256, Cfg::SMEM_BYTES, stream>>>(map1, map2, out, p);
  }
  B200_LAUNCH_OK("sampler_fwd_kernel");
  return 0;
}

}  // namespace

namespace b200 {

bool sampler_fast_fwd_shape_ok(int patchH, int patchW, int dpW) {
  return (patchH == 21 && patchW == 21 && dpW == 2) || (patchH == 9 && patchW == 9 && dpW == 1);
}

int sampler_fast_forward(const float *in1, const float *in2, float *out, int B, int C, int H, int W,
                         const int *q, cudaStream_t stream) {
  const int patchH = q[2], patchW = q[3], dpH = q[8], dpW = q[9];
  if (patchH == 21 && patchW == 21 && dpW == 2)
    return launch_fwd<FwdCfg<21, 21, 2>>(in1, in2, out, B, C, H, W, dpH, 441ll * H * W, 1.f, 1.f, stream);
  if (patchH == 9 && patchW == 9 && dpW == 1)
    return launch_fwd<FwdCfg<9, 9, 1>>(in1, in2, out, B, C, H, W, dpH, 81ll * H * W, 1.f, 1.f, stream);
  set_error("sampler_fast_forward: no instantiation for patch %dx%d dilation_patch_w %d", patchH,
            patchW, dpW);
  return -1;
}

// Fused merge block: `out` points at channel c_off of the (B, c_total, H, W) concat tensor.
int sampler_fast_forward_merge(const float *in1, const float *in2, float *out, int B, int C, int H, int W,
                               const int *q, long long out_bstride, float slope, cudaStream_t stream) {
  const int patchH = q[2], patchW = q[3], dpH = q[8], dpW = q[9];
  if (patchH == 21 && patchW == 21 && dpW == 2)
    return launch_fwd<FwdCfg<21, 21, 2>>(in1, in2, out, B, C, H, W, dpH, out_bstride, 1.0f / (float)C, slope, stream);
  if (patchH == 9 && patchW == 9 && dpW == 1)
    return launch_fwd<FwdCfg<9, 9, 1>>(in1, in2, out, B, C, H, W, dpH, out_bstride, 1.0f / (float)C, slope, stream);
  set_error("sampler_fast_forward_merge: no instantiation for patch %dx%d dilation_patch_w %d", patchH,
            patchW, dpW);
  return -1;
}

}  // namespace b200
