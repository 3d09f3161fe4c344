// tma_maps.hpp -- tensor maps (TMA descriptors) of the arrays k_sweep_fast streams, and their host-side encoder.
//
// A sub-tile of 32 spots is fetched by four cp.async.bulk.tensor instructions (SASS UTMALDG) issued by one lane:
//   Y     double [d][ld]       two boxes of [d rows][16 spots], 128-byte rows, SWIZZLE_128B (the 16-byte chunks of a
//                              row are XOR-ed with the row index mod 8, so both the per-spot column reads and the
//                              DMMA fragment reads of the landing buffer are bank-conflict free without padding)
//   ell   int32  [maxdeg][ld]  one box of [maxdeg][32]
//   orig  int32  [1][ld]       one box of [1][32]
// Box coordinates are element-granular (no alignment requirement on a segment's first spot) and columns beyond ld
// are zero-filled by the hardware.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>

#include <cstdint>

namespace bsb {

struct alignas(64) SweepMaps {
  CUtensorMap y, ell, orig;
};

// Returns 0 on success; on failure `msg` names the step.  ell may be null (statistics-only passes).
int encode_sweep_maps(const double *Y, int64_t ld, int d, const int32_t *ell, int maxdeg, const int32_t *orig,
                      SweepMaps *out, const char **msg);

}   // namespace bsb
