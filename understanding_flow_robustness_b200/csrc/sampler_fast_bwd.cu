// sampler_fast_bwd.cu -- register-blocked, TMA-staged backward of the spatial correlation sampler
// for the FlowNetC / PWC-Net call-site structure (kernel_size 1, stride 1, padding 0):
//
//   WHICH == 1:  gIn1[n,c,h,w] = sum_{e,k} G[n,e,k,h,w]           * in2[n,c,h+e*dpH, w+(k-RW)*dpW]
//   WHICH == 2:  gIn2[n,c,y,x] = sum_{e,k} G[n,e,k,y-e*dpH,x-dxk] * in1[n,c,y-e*dpH, x-dxk]
//
// (e = row displacement index - RH, dxk = (k-RW)*dpW.)  Gather form, no atomics, deterministic;
// replaces correlation_cuda_backward_kernel_input1/2 (correlation_cuda_kernel.cu:87-233, one launch
// per batch sample, 25-thread blocks, thread-(0,0) serial reduction).
//
// Both gradients are the same contraction: for a pixel row s and a source row R of the OTHER
// feature map (e = R - s for gIn1, e = s - R for gIn2) and 8 consecutive pixels t,
//     acc[t][c] += Gk[k'][t] * other[c][R][x0 + t + k'*dpW - HALO],   k' = 0..PW-1
// with Gk the grad_output row re-indexed (gIn2 walks k backwards and reads G at the source pixel).
// One thread owns 8 pixels x 8 channels (64 accumulators) and, per source row, holds half of the
// PW x 8 grad_output coefficients in registers (two passes over k), so each 16-byte window load of
// `other` feeds ~12 FFMA.  A warp is 4 pixel rows x 8 interleaved channels: the 4 row-lanes read the
// same `other` address (multicast), the 8 channel-lanes read the same grad_output address.
// Source rows outside the image are never visited, so work tracks the in-bounds MAC count.
// Persistent CTAs, one per SM, a TMA/mbarrier ring over source rows that runs across units.
#include <algorithm>
#include <cstdlib>
#include <queue>
#include <vector>

#include "sampler_fast.cuh"

namespace {
using namespace b200dev;

constexpr int odd4b(int x) {
  int c = (x + 3) / 4;
  if (c % 2 == 0) ++c;
  return 4 * c;
}
constexpr int round_up_c(int a, int b) { return (a + b - 1) / b * b; }

template <int PH_, int PW_, int DPW_, int WHICH_, int NCH_ = 8>
struct BwdCfg {
  static constexpr int PH = PH_, PW = PW_, DPW = DPW_, WHICH = WHICH_;
  static constexpr int T = 8, NCH = NCH_, CGROUPS = 4;
  static constexpr int CH_UNIT = 16 * NCH;            // 128 channels per unit (16 channel lanes x NCH)
  static constexpr int COLS = CGROUPS * T;            // 32 pixels per unit row
  static constexpr int RH = (PH - 1) / 2, RW = (PW - 1) / 2;
  static constexpr int HALO = RW * DPW;
  static constexpr int WIN = T + (PW - 1) * DPW;
  static_assert(WIN % 4 == 0, "window must be whole 16-byte chunks");
  static constexpr int NC = odd4b(COLS + (PW - 1) * DPW);   // staged columns of `other`
  static constexpr int OTHER_FLOATS = CH_UNIT * NC;
  // grad_output staging: WHICH 1 -> four separate [PW][GC] boxes, the box of row-lane li shifted
  // left by 4*li columns so the four lanes hit different banks; WHICH 2 -> one [4][PW][NC] box.
  static constexpr int GC = COLS + 12;
  static constexpr int GLI = round_up_c(PW * GC, 32);
  static constexpr int G_FLOATS = WHICH == 1 ? 4 * GLI : round_up_c(4 * PW * NC, 32);
  static constexpr int STAGE_FLOATS = OTHER_FLOATS + G_FLOATS;
  static constexpr int STAGE_BYTES = STAGE_FLOATS * 4;
  static_assert((OTHER_FLOATS * 4) % 128 == 0 && STAGE_BYTES % 128 == 0 && (GLI * 4) % 128 == 0,
                "TMA destination alignment");
  static constexpr int NST = (4 * STAGE_BYTES + 256 <= 227 * 1024) ? 4 : 3;
  static constexpr int SMEM_BYTES = NST * STAGE_BYTES + 256;
  static constexpr int KA = (PW + 1) / 2;             // first k' pass: [0, KA), second: [KA, PW)
};

struct BwdParams {
  const int *plan;   // optional LPT schedule: [grid+1] offsets then unit ids (device memory), or nullptr
  int B, C, H, W, dpH, NCT, NCB, total_units;
  int wide_store;   // W % 8 == 0 and 32-byte aligned gradient: 256-bit stores
  int ctail;        // C % CH_UNIT != 0: map_other is 4-D (W, H, C, B), channels past C read as zero and are not stored
  b200::SamplerGroups g;   // prefix[] unused here (every group has NCT*NCB units)
};

struct BUnit {
  int n, rp, s0, NS, c0, cb, Rlo, nsteps;
};

template <class Cfg>
__device__ __forceinline__ void decode_bunit(const BwdParams &p, int u, BUnit &x) {
  const int per_group = p.NCT * p.NCB;
  const int ups = p.g.ngroups * per_group;
  x.n = u / ups;
  int r = u - x.n * ups;
  const int gi = r / per_group;
  r -= gi * per_group;
  x.c0 = (r / p.NCB) * Cfg::COLS;
  x.cb = r % p.NCB;
  x.rp = p.g.rp[gi];
  x.s0 = p.g.s0[gi];
  x.NS = (p.H - x.rp + p.dpH - 1) / p.dpH;
  const int s_last = x.s0 + 3 < x.NS - 1 ? x.s0 + 3 : x.NS - 1;
  x.Rlo = x.s0 - Cfg::RH > 0 ? x.s0 - Cfg::RH : 0;
  const int Rhi = s_last + Cfg::RH < x.NS - 1 ? s_last + Cfg::RH : x.NS - 1;
  x.nsteps = Rhi - x.Rlo + 1;
}

// One k' pass [K0, K1): coefficients to registers, then stream the `other` window per channel.
template <class Cfg, int K0, int K1>
__device__ __forceinline__ void bwd_pass(float2 (&acc2)[Cfg::T / 2][Cfg::NCH], const float *gs,
                                         const float *vb) {
  constexpr int DPW = Cfg::DPW, PW = Cfg::PW, NC = Cfg::NC, T = Cfg::T;
  constexpr int NK = K1 - K0;
  float2 Gr[NK][T / 2];  // coefficient pairs (pixel 2tp, 2tp+1)
#pragma unroll
  for (int kk = 0; kk < NK; ++kk) {
    const int k = K0 + kk;
    if (Cfg::WHICH == 1) {
      const float4 g0 = lds128(gs + k * Cfg::GC);
      const float4 g1 = lds128(gs + k * Cfg::GC + 4);
      Gr[kk][0] = make_float2(g0.x, g0.y); Gr[kk][1] = make_float2(g0.z, g0.w);
      Gr[kk][2] = make_float2(g1.x, g1.y); Gr[kk][3] = make_float2(g1.z, g1.w);
    } else {
      const float *gp = gs + (PW - 1 - k) * NC + k * DPW;
      if ((k * DPW) % 4 == 0) {
        const float4 g0 = lds128(gp), g1 = lds128(gp + 4);
        Gr[kk][0] = make_float2(g0.x, g0.y); Gr[kk][1] = make_float2(g0.z, g0.w);
        Gr[kk][2] = make_float2(g1.x, g1.y); Gr[kk][3] = make_float2(g1.z, g1.w);
      } else if ((k * DPW) % 2 == 0) {
#pragma unroll
        for (int h = 0; h < 4; ++h) Gr[kk][h] = *reinterpret_cast<const float2 *>(gp + 2 * h);
      } else {
#pragma unroll
        for (int h = 0; h < 4; ++h) Gr[kk][h] = make_float2(gp[2 * h], gp[2 * h + 1]);
      }
    }
  }
  constexpr int MB = (K0 * DPW) / 4 * 4;            // first window float this pass touches (16-B aligned)
  constexpr int ME = T + (K1 - 1) * DPW;            // one past the last
  constexpr int NL = (ME - MB + 3) / 4;
#pragma unroll
  for (int ci = 0; ci < Cfg::NCH; ++ci) {
#pragma unroll
    for (int sg = 0; sg < NL; ++sg) {
      const float4 v4 = lds128(vb + ci * 16 * NC + MB + 4 * sg);
      if constexpr (DPW % 2 == 0) {
        // pixel pair tp meets window pair tp + k'*DPW/2: one FFMA2 per (tp, k')
        const float2 vp[2] = {make_float2(v4.x, v4.y), make_float2(v4.z, v4.w)};
#pragma unroll
        for (int uu = 0; uu < 2; ++uu) {
#pragma unroll
          for (int tp = 0; tp < T / 2; ++tp) {
            const int d = MB / 2 + 2 * sg + uu - tp;  // = k' * DPW / 2
            if (d >= 0 && d % (DPW / 2) == 0 && d / (DPW / 2) >= K0 && d / (DPW / 2) < K1)
              acc2[tp][ci] = __ffma2_rn(Gr[d / (DPW / 2) - K0][tp], vp[uu], acc2[tp][ci]);
          }
        }
      } else {
        const float v[4] = {v4.x, v4.y, v4.z, v4.w};
#pragma unroll
        for (int uu = 0; uu < 4; ++uu) {
#pragma unroll
          for (int t = 0; t < T; ++t) {
            const int d = MB + 4 * sg + uu - t;  // = k' * DPW
            if (d >= 0 && d % DPW == 0 && d / DPW >= K0 && d / DPW < K1) {
              const float2 g = Gr[d / DPW - K0][t / 2];
              float2 &r = acc2[t / 2][ci];
              if (t % 2 == 0) r.x = fmaf(g.x, v[uu], r.x);
              else r.y = fmaf(g.y, v[uu], r.y);
            }
          }
        }
      }
    }
  }
}

template <class Cfg>
__global__ void __launch_bounds__(288, 1)
sampler_bwd_kernel(const __grid_constant__ CUtensorMap map_other, const __grid_constant__ CUtensorMap map_g,
                   float *__restrict__ gin, const BwdParams p) {
  constexpr int NST = Cfg::NST, RH = Cfg::RH, PW = Cfg::PW, NC = Cfg::NC, T = Cfg::T;
  extern __shared__ __align__(128) float smem[];
  uint64_t *full_bar = reinterpret_cast<uint64_t *>(smem + NST * Cfg::STAGE_FLOATS);
  uint64_t *empty_bar = full_bar + NST;
  BUnit &px = *reinterpret_cast<BUnit *>(empty_bar + NST);  // producer-side unit (thread 0 only)

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  // Warp = 2 pixel rows (adjacent lanes: an aligned lane pair reads the same `other` address, the
  // only sharing an LDS.128 merges) x 16 interleaved channel lanes x 8 channels each = 128 channels.
  // The CTA's 8 warps are 4 column groups x 2 row pairs; a warp whose two rows are both outside the
  // +-RH band of the current source row skips the step as a whole (22 of 24 steps useful instead of
  // 84 of 96 lanes).
  const int li = (lane & 1) + 2 * (warp >> 2);   // pixel row inside the group (0..3)
  const int lj = lane >> 1;                      // channel lane (0..15)
  const int cgw = warp & 3;

  if (tid == 0) {
    tma_prefetch_desc(&map_other);
    tma_prefetch_desc(&map_g);
    for (int s = 0; s < NST; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 8);
    }
    fence_barrier_init();
  }
  __syncthreads();

  // unit sequence of this CTA: the host's LPT plan if given, else static round-robin
  const int *plan = p.plan;
  const int it_end = plan ? plan[blockIdx.x + 1] : p.total_units;
  const int it_step = plan ? 1 : (int)gridDim.x;
  const int it_begin = plan ? plan[blockIdx.x] : (int)blockIdx.x;
  const int *ulist = plan ? plan + gridDim.x + 1 : nullptr;
  const int nunits = it_end;

  // ---- producer (thread 0): walks (unit, source row) NST-1 steps ahead of the math
  int pu = it_begin, ps = 0;
  uint32_t pq = 0;
  if (tid == 256 && pu < nunits) decode_bunit<Cfg>(p, ulist ? ulist[pu] : pu, px);
  auto issue = [&]() {
    if (pu >= nunits) return;
    const int st = pq % NST;
    float *dst = smem + st * Cfg::STAGE_FLOATS;
    const int R = px.Rlo + ps;
    uint32_t bytes = Cfg::OTHER_FLOATS * 4;
    if (Cfg::WHICH == 1) {
      for (int l = 0; l < 4; ++l) {
        const int s = px.s0 + l, e = R - s;
        if (s < px.NS && e >= -RH && e <= RH) bytes += PW * Cfg::GC * 4;
      }
    } else {
      bytes += 4 * PW * NC * 4;
    }
    mbar_arrive_expect_tx(&full_bar[st], bytes);
    if (p.ctail)
      tma_load_4d(dst, &map_other, &full_bar[st], px.c0 - Cfg::HALO, R * p.dpH + px.rp, px.cb * Cfg::CH_UNIT, px.n);
    else
      tma_load_3d(dst, &map_other, &full_bar[st], px.c0 - Cfg::HALO, R * p.dpH + px.rp,
                  px.n * p.C + px.cb * Cfg::CH_UNIT);
    float *gd = dst + Cfg::OTHER_FLOATS;
    if (Cfg::WHICH == 1) {
      for (int l = 0; l < 4; ++l) {
        const int s = px.s0 + l, e = R - s;
        if (s < px.NS && e >= -RH && e <= RH)
          tma_load_5d(gd + l * Cfg::GLI, &map_g, &full_bar[st], px.c0 - 4 * l, s * p.dpH + px.rp, 0,
                      e + RH, px.n);
      }
    } else {
      tma_load_5d(gd, &map_g, &full_bar[st], px.c0 - Cfg::HALO, R * p.dpH + px.rp, 0,
                  px.s0 - R + RH, px.n);
    }
    ++pq;
    if (++ps == px.nsteps) {
      ps = 0;
      pu += it_step;
      if (pu < nunits) decode_bunit<Cfg>(p, ulist ? ulist[pu] : pu, px);
    }
  };
  if (warp == 8) {
    // ---- dedicated producer warp: one lane walks every (unit, source row) of this CTA, NST deep
    if (lane == 0) {
      for (uint32_t n = 0; pu < nunits; ++n) {
        if (n >= (uint32_t)NST) mbar_wait(&empty_bar[n % NST], ((n / NST) - 1) & 1);
        issue();
      }
    }
    return;
  }

  uint32_t q = 0;
  for (int ui = it_begin; ui < nunits; ui += it_step) {
    BUnit x;
    decode_bunit<Cfg>(p, ulist ? ulist[ui] : ui, x);
    const int s = x.s0 + li;
    // per-thread offsets inside a stage
    const int offv = lj * NC + cgw * T;
    const int offg = Cfg::OTHER_FLOATS +
                     (Cfg::WHICH == 1 ? li * Cfg::GLI + cgw * T + 4 * li : li * PW * NC + cgw * T);

    float2 acc[T / 2][Cfg::NCH];  // pixel pairs x channels
#pragma unroll
    for (int t = 0; t < T / 2; ++t)
#pragma unroll
      for (int c = 0; c < Cfg::NCH; ++c) acc[t][c] = make_float2(0.f, 0.f);

    for (int step = 0; step < x.nsteps; ++step, ++q) {
      const int st = q % NST;
      mbar_wait(&full_bar[st], (q / NST) & 1);
      const int R = x.Rlo + step;
      const int e = Cfg::WHICH == 1 ? R - s : s - R;
      const bool live = s < x.NS && e >= -RH && e <= RH;
      if (__any_sync(0xffffffffu, live) && live) {
        const float *vb = smem + st * Cfg::STAGE_FLOATS + offv;
        const float *gs = smem + st * Cfg::STAGE_FLOATS + offg;
        bwd_pass<Cfg, 0, Cfg::KA>(acc, gs, vb);
        bwd_pass<Cfg, Cfg::KA, PW>(acc, gs, vb);
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&empty_bar[st]);
    }

    if (s < x.NS) {
      const int h = s * p.dpH + x.rp;
      const int x0 = x.c0 + cgw * T;
      const size_t HW = (size_t)p.H * p.W;
      float *o = gin + ((size_t)x.n * p.C + x.cb * Cfg::CH_UNIT + lj) * HW +
                 (size_t)h * p.W + x0;
      const int ch0 = x.cb * Cfg::CH_UNIT + lj;   // channel of ci = 0 (ci-th channel: ch0 + 16 ci)
      const int nci = !p.ctail ? Cfg::NCH : (p.C - ch0 + 15) / 16;   // channels of this lane that exist
      if (p.wide_store && x0 + 4 < p.W) {
#pragma unroll
        for (int ci = 0; ci < Cfg::NCH; ++ci)   // one 32-byte sector per (channel, row): STG.256
          if (ci < nci)
          asm volatile("st.global.v8.f32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(o + (size_t)ci * 16 * HW),
                       "f"(acc[0][ci].x), "f"(acc[0][ci].y), "f"(acc[1][ci].x), "f"(acc[1][ci].y),
                       "f"(acc[2][ci].x), "f"(acc[2][ci].y), "f"(acc[3][ci].x), "f"(acc[3][ci].y)
                       : "memory");
      } else
#pragma unroll
      for (int ci = 0; ci < Cfg::NCH; ++ci) {
        if (x0 < p.W && ci < nci)
          *reinterpret_cast<float4 *>(o + (size_t)ci * 16 * HW) =
              make_float4(acc[0][ci].x, acc[0][ci].y, acc[1][ci].x, acc[1][ci].y);
        if (x0 + 4 < p.W && ci < nci)
          *reinterpret_cast<float4 *>(o + (size_t)ci * 16 * HW + 4) =
              make_float4(acc[2][ci].x, acc[2][ci].y, acc[3][ci].x, acc[3][ci].y);
      }
    }
  }
}

// fills the group table / unit counts shared by the launcher and the planner
// fills the group table / unit counts shared by the launcher and the planner
template <class Cfg>
int bwd_geometry(BwdParams &p, int B, int C, int H, int W, int dpH) {
  p.plan = nullptr;
  p.wide_store = 0;
  p.B = B; p.C = C; p.H = H; p.W = W; p.dpH = dpH;
  p.NCT = (W + Cfg::COLS - 1) / Cfg::COLS;
  p.NCB = (C + Cfg::CH_UNIT - 1) / Cfg::CH_UNIT;
  p.ctail = (C % Cfg::CH_UNIT != 0) ? 1 : 0;
  int ng = 0;
  for (int rp = 0; rp < dpH; ++rp) {
    const int NS = b200::sublattice_rows(H, dpH, rp);
    for (int s0 = 0; s0 < NS; s0 += b200::kRowsPerGroup) {
      B200_CHECK(ng < b200::kSamplerMaxGroups, "sampler_fast_backward: too many row groups");
      p.g.prefix[ng] = ng * p.NCT * p.NCB;
      p.g.rp[ng] = (short)rp;
      p.g.s0[ng] = (short)s0;
      ++ng;
    }
  }
  p.g.ngroups = ng;
  p.g.units_per_sample = ng * p.NCT * p.NCB;
  p.total_units = p.g.units_per_sample * B;
  return 0;
}

// source rows one unit of group gi walks (its cost)
template <class Cfg>
int bwd_group_steps(const BwdParams &p, int gi) {
  const int NS = b200::sublattice_rows(p.H, p.dpH, p.g.rp[gi]);
  const int s0 = p.g.s0[gi];
  const int s_last = s0 + 3 < NS - 1 ? s0 + 3 : NS - 1;
  const int Rlo = s0 - Cfg::RH > 0 ? s0 - Cfg::RH : 0;
  const int Rhi = s_last + Cfg::RH < NS - 1 ? s_last + Cfg::RH : NS - 1;
  return Rhi - Rlo + 1;
}

constexpr int kBwdUnitOverheadRows = 0;

static inline int bwd_grid(int total_units) {
  return total_units < b200::num_sms() ? total_units : b200::num_sms();
}

// Longest-processing-time-first schedule of the units over `grid` persistent CTAs.
// h_plan: [grid + 1] offsets, then total_units unit ids.  Host only.
template <class Cfg>
int bwd_plan(int B, int C, int H, int W, int dpH, int grid, int *h_plan, size_t bytes) {
  BwdParams p;
  if (int e = bwd_geometry<Cfg>(p, B, C, H, W, dpH)) return e;
  B200_CHECK(bytes >= sizeof(int) * ((size_t)grid + 1 + p.total_units), "sampler_backward_plan: buffer too small");
  const int per_group = p.NCT * p.NCB;
  std::vector<int> cost(p.g.ngroups);
  for (int gi = 0; gi < p.g.ngroups; ++gi) cost[gi] = bwd_group_steps<Cfg>(p, gi);
  std::vector<int> units(p.total_units);
  for (int u = 0; u < p.total_units; ++u) units[u] = u;
  // cost of a unit = the source rows it walks + a fixed part (drain of the accumulators, stores, ring turn-around)
  // worth B200CORR_BWD_UNIT_OVERHEAD rows (default below: fitted on a B200, scripts/time_bwd_plan.py)
  int overhead = kBwdUnitOverheadRows;
  if (const char *e = getenv("B200CORR_BWD_UNIT_OVERHEAD")) overhead = atoi(e);
  for (int &c : cost) c += overhead;
  auto cost_of = [&](int u) { return cost[(u % p.g.units_per_sample) / per_group]; };
  std::stable_sort(units.begin(), units.end(), [&](int a, int b) { return cost_of(a) > cost_of(b); });
  using Bin = std::pair<long long, int>;  // (load, cta)
  std::priority_queue<Bin, std::vector<Bin>, std::greater<Bin>> heap;
  for (int c = 0; c < grid; ++c) heap.push({0, c});
  std::vector<std::vector<int>> lists(grid);
  for (int u : units) {
    Bin b = heap.top();
    heap.pop();
    lists[b.second].push_back(u);
    heap.push({b.first + cost_of(u), b.second});
  }
  // LPT leaves the makespan up to 4-17 % above the mean when a CTA gets only 3-7 units (batch 4-8: 122 vs a mean
  // of 116.8 source rows at batch 8; the optimum, found by an integer program over the three unit sizes, is 118).
  // Local search on the most loaded CTA -- move one of its units, or swap one against a cheaper unit of another
  // CTA, whichever lowers max(load_a, load_b) most -- reaches that optimum in a few hundred steps.
  const char *ls_env = getenv("B200CORR_BWD_PLAN_LS");   // diagnostics: 0 = plain LPT
  if (!ls_env || atoi(ls_env) != 0) {
    std::vector<long long> load(grid, 0);
    for (int c = 0; c < grid; ++c)
      for (int u : lists[c]) load[c] += cost_of(u);
    for (int iter = 0; iter < 4 * grid + 64; ++iter) {
      int a = 0;
      for (int c = 1; c < grid; ++c)
        if (load[c] > load[a]) a = c;
      long long best = load[a];
      int best_b = -1, best_i = -1, best_j = -1;   // j < 0: move lists[a][i] to b; else swap with lists[b][j]
      for (int b = 0; b < grid; ++b) {
        if (b == a) continue;
        int seen_x = -1;
        for (size_t i = 0; i < lists[a].size(); ++i) {
          const int x = cost_of(lists[a][i]);
          if (x == seen_x) continue;              // lists are sorted by cost at first; a repeat changes nothing
          seen_x = x;
          const long long mv = std::max(load[a] - x, load[b] + x);
          if (mv < best) { best = mv; best_b = b; best_i = (int)i; best_j = -1; }
          int seen_y = -1;
          for (size_t j = 0; j < lists[b].size(); ++j) {
            const int y = cost_of(lists[b][j]);
            if (y >= x || y == seen_y) continue;
            seen_y = y;
            const long long sw = std::max(load[a] - x + y, load[b] + x - y);
            if (sw < best) { best = sw; best_b = b; best_i = (int)i; best_j = (int)j; }
          }
        }
      }
      if (best_b < 0) break;
      const int u = lists[a][best_i], x = cost_of(u);
      if (best_j < 0) {
        lists[a].erase(lists[a].begin() + best_i);
        lists[best_b].push_back(u);
        load[a] -= x; load[best_b] += x;
      } else {
        const int v = lists[best_b][best_j], y = cost_of(v);
        lists[a][best_i] = v; lists[best_b][best_j] = u;
        load[a] += y - x; load[best_b] += x - y;
      }
    }
    // long units first inside a CTA (what LPT produced, restored after the exchanges): the ring keeps running
    // across units, so the order only decides which unit is in flight when the kernel drains
    for (int c = 0; c < grid; ++c)
      std::stable_sort(lists[c].begin(), lists[c].end(), [&](int a2, int b2) { return cost_of(a2) > cost_of(b2); });
  }
  int off = 0;
  int *ids = h_plan + grid + 1;
  for (int c = 0; c < grid; ++c) {
    h_plan[c] = off;
    for (int u : lists[c]) ids[off++] = u;
  }
  h_plan[grid] = off;
  return 0;
}

template <class Cfg>
int launch_bwd(const float *other, const float *gout, float *gin, int B, int C, int H, int W,
               int dpH, const int *plan, cudaStream_t stream) {
  BwdParams p;
  if (int e = bwd_geometry<Cfg>(p, B, C, H, W, dpH)) return e;
  p.plan = plan;
  p.wide_store = (W % 8 == 0 && ((uintptr_t)gin & 31) == 0) ? 1 : 0;
  if (p.total_units == 0) return 0;

  CUtensorMap map_o, map_g;
  {
    const uint64_t dims[4] = {(uint64_t)W, (uint64_t)H, p.ctail ? (uint64_t)C : (uint64_t)B * C, (uint64_t)B};
    const uint64_t strides[4] = {4, (uint64_t)W * 4, (uint64_t)H * W * 4, (uint64_t)H * W * 4 * C};
    const uint32_t box[4] = {(uint32_t)Cfg::NC, 1, (uint32_t)Cfg::CH_UNIT, 1};
    if (int e = b200::make_tensor_map(&map_o, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, p.ctail ? 4 : 3, other, dims,
                                      strides, box, CU_TENSOR_MAP_SWIZZLE_NONE,
                                      CU_TENSOR_MAP_L2_PROMOTION_L2_128B))
      return e;
  }
  {
    const uint64_t HW = (uint64_t)H * W;
    const uint64_t dims[5] = {(uint64_t)W, (uint64_t)H, (uint64_t)Cfg::PW, (uint64_t)Cfg::PH, (uint64_t)B};
    const uint64_t strides[5] = {4, (uint64_t)W * 4, HW * 4, HW * 4 * Cfg::PW, HW * 4 * Cfg::PW * Cfg::PH};
    const uint32_t box1[5] = {(uint32_t)Cfg::GC, 1, (uint32_t)Cfg::PW, 1, 1};
    const uint32_t box2[5] = {(uint32_t)Cfg::NC, 1, (uint32_t)Cfg::PW, 4, 1};
    if (int e = b200::make_tensor_map(&map_g, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 5, gout, dims,
                                      strides, Cfg::WHICH == 1 ? box1 : box2,
                                      CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B))
      return e;
  }
  auto kern = sampler_bwd_kernel<Cfg>;
  static bool attr_done[64] = {};  // per (kernel instantiation, device)
  if (int e = b200::set_max_smem_once((const void *)kern, Cfg::SMEM_BYTES, attr_done)) return e;
  const int grid = bwd_grid(p.total_units);
  kern<<<grid, 288, Cfg::SMEM_BYTES, stream>>>(map_o, map_g, gin, p);
  B200_LAUNCH_OK(Cfg::WHICH == 1 ? "sampler_bwd_kernel<gIn1>" : "sampler_bwd_kernel<gIn2>");
  return 0;
}

}  // namespace

namespace b200 {

// channels per unit: 128 (8 per thread) when C allows, else 32 (2 per thread: PWC-Net's 96/64/32)
#define B200_BWD_DISPATCH(FN, ...)                                                         \
  do {                                                                                     \
    if (patchH == 21 && patchW == 21 && dpW == 2) {                                         \
      if (C % 128 == 0) { FN(21, 21, 2, 8, __VA_ARGS__); } else { FN(21, 21, 2, 2, __VA_ARGS__); } \
    } else if (patchH == 9 && patchW == 9 && dpW == 1) {                                    \
      if (C % 128 == 0) { FN(9, 9, 1, 8, __VA_ARGS__); } else { FN(9, 9, 1, 2, __VA_ARGS__); }     \
    }                                                                                      \
  } while (0)

int sampler_fast_backward(const float *in1, const float *in2, const float *gout, float *gin1,
                          float *gin2, int B, int C, int H, int W, const int *q, const int *plan,
                          cudaStream_t stream) {
  const int patchH = q[2], patchW = q[3], dpH = q[8], dpW = q[9];
#define RUN(PH, PW, DP, N, dummy)                                                                     \
  {                                                                                                   \
    if (int e = launch_bwd<BwdCfg<PH, PW, DP, 1, N>>(in2, gout, gin1, B, C, H, W, dpH, plan, stream)) \
      return e;                                                                                       \
    return launch_bwd<BwdCfg<PH, PW, DP, 2, N>>(in1, gout, gin2, B, C, H, W, dpH, plan, stream);      \
  }
  B200_BWD_DISPATCH(RUN, 0);
#undef RUN
  set_error("sampler_fast_backward: no instantiation for patch %dx%d dilation_patch_w %d", patchH,
            patchW, dpW);
  return -1;
}

// number of ints of the backward plan (0 if the problem has no units)
size_t sampler_fast_backward_plan_ints(int B, int C, int H, int W, const int *q) {
  BwdParams p;
  const int patchH = q[2], patchW = q[3], dpW = q[9];
  int e = -1;
#define GEO(PH, PW, DP, N, dummy) e = bwd_geometry<BwdCfg<PH, PW, DP, 1, N>>(p, B, C, H, W, q[8]);
  B200_BWD_DISPATCH(GEO, 0);
#undef GEO
  if (e || p.total_units == 0) return 0;
  return (size_t)bwd_grid(p.total_units) + 1 + p.total_units;
}

int sampler_fast_backward_plan(int B, int C, int H, int W, const int *q, int *h_plan, size_t bytes) {
  const size_t n = sampler_fast_backward_plan_ints(B, C, H, W, q);
  if (n == 0) return 0;
  const int patchH = q[2], patchW = q[3], dpW = q[9];
  BwdParams p;
#define PLAN(PH, PW, DP, N, dummy)                                                                  \
  {                                                                                                 \
    bwd_geometry<BwdCfg<PH, PW, DP, 1, N>>(p, B, C, H, W, q[8]);                                    \
    return bwd_plan<BwdCfg<PH, PW, DP, 1, N>>(B, C, H, W, q[8], bwd_grid(p.total_units), h_plan, bytes); \
  }
  B200_BWD_DISPATCH(PLAN, 0);
#undef PLAN
  return -1;
}

// The structure the register-blocked kernels cover (everything else runs on sampler_generic.cu).
bool sampler_fast_applicable(int B, int C, int H, int W, const int *q, int dtype, int backward) {
  (void)B;
  if (dtype != B200CORR_F32) return false;
  if (q[0] != 1 || q[1] != 1 || q[4] != 0 || q[5] != 0 || q[10] != 1 || q[11] != 1) return false;
  const int patchH = q[2], patchW = q[3], dpH = q[8], dpW = q[9];
  const bool shape = (patchH == 21 && patchW == 21 && dpW == 2) || (patchH == 9 && patchW == 9 && dpW == 1);
  if (!shape) return false;
  if (W % 4 != 0 || W < 4 || H < 1) return false;
  if (dpH < 1 || dpH > 8) return false;
  if ((long long)B * C >= (1ll << 31)) return false;
  int ng = 0;
  for (int rp = 0; rp < dpH; ++rp) ng += (sublattice_rows(H, dpH, rp) + kRowsPerGroup - 1) / kRowsPerGroup;
  if (ng > kSamplerMaxGroups) return false;
  // any channel count: chunks that do not tile C exactly read zeros past C through 4-D tensor maps
  // (backward: only the 32-channel units; the 128-channel units are picked for C % 128 == 0)
  (void)backward;
  return C >= 1;
}

}  // namespace b200
