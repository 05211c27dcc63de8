// sampler_fast.cuh -- shared host-side geometry of the register-blocked sampler kernels.
#pragma once
#include "common.cuh"

namespace b200 {

// Rows of one "row parity class" rp (rows h with h % dpH == rp) form a sub-lattice on which every
// row displacement of the patch is a unit step.  A row group is 4 consecutive sub-lattice rows.
constexpr int kSamplerMaxGroups = 160;
constexpr int kRowsPerGroup = 4;

struct SamplerGroups {
  int ngroups;
  int units_per_sample;
  int prefix[kSamplerMaxGroups + 1];   // first unit (within a sample) of each row group
  short rp[kSamplerMaxGroups];         // row parity class of the group
  short s0[kSamplerMaxGroups];         // first sub-lattice row of the group
};

static inline int sublattice_rows(int H, int dpH, int rp) { return H > rp ? (H - rp + dpH - 1) / dpH : 0; }

}  // namespace b200
