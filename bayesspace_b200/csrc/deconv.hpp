// deconv.hpp -- kernel arguments of the iterate_deconv sweep (spatialEnhance).
//
// HBM layout (parents in internal order ip = 0..n0-1, sorted by (group, parent colour, id);
// subspot (r, ip) lives at r*ld0 + ip, so a warp that handles consecutive parents streams
// S contiguous runs per PC column):
//   Y     double [d][S][ld0]   current subspot PCs (centred)
//   Y0    double [d][ld0]      parent PCs (centred), the anchor of the jitter prior
//   z     uint8  [S][ld0]      w double [S][ld0]
//   ell   int32  [maxdeg][S][ld0]  subspot neighbours as internal subspot ids, -1 padded
//   orig0 int32  [ld0]         original parent id (keys the Philox counters: s = r*n0 + j0)
//   A     double [np][S][ld0]  adaptive (RAM) proposal factors, packed lower triangle (optional)
//   nacc  int32  [ld0]         accepted Y moves per parent (adaptive acceptance rate)
#pragma once
#include "layout.hpp"

namespace bsb {

struct DeconvArgs {
  double *Y;
  const double *Y0;
  uint8_t *z;
  double *w;
  const int32_t *ell;
  int maxdeg;
  const int32_t *orig0;
  double *A;
  int32_t *nacc;
  int64_t ld0, n0;
  int S;
  const int32_t *seg_begin, *seg_count, *seg_slot;   // segments over PARENTS
  int nseg;
  const double *params;
  ParamLayout PL;
  double *partials;
  StatLayout SL;
  const uint32_t *iter;
  uint32_t key0, key1;
  double gamma, c, err_sd;
  int tdist, adaptive, adapt_before, stats_only;
  int *err;
  // All colour classes in ONE launch (when every CTA of the launch is resident at once): only the label scan depends on
  // the neighbouring parents' labels, so a CTA runs its Y / w updates at once and waits with its label scan until
  // every CTA of the previous colour class has published its labels (done[c]: segments of colour c finished, counted
  // over all iterations; target = iteration * segments of that colour).  fuse == 0: one launch per colour class.
  int fuse, ncol;
  const int32_t *colour_seg_start;   // ncol + 1 offsets into the launch list (colour-major)
  uint32_t *done;                    // ncol counters
};

}   // namespace bsb
