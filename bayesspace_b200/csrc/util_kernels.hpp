// util_kernels.hpp -- launchers of the dimension-agnostic staging kernels.
#pragma once
#include <atomic>
#include <cstdint>
#include <cuda_runtime.h>

#include "layout.hpp"

namespace bsb {
extern std::atomic<long long> g_launches;
void launch_fill_f64(double *p, int64_t n, double v, cudaStream_t s);
void launch_colmeans(const double *Y, int64_t n, int d, double *ybar, cudaStream_t s);
void launch_permute_center(const double *Y, int64_t src_ld, int64_t count, int d, const int32_t *orig,
                           const double *ybar, double *Yp, int64_t dst_ld, cudaStream_t s, int64_t orig_off = 0);
void launch_colsums_groups(const double *Y, int64_t ldy, const int64_t *group_lo, int64_t row_off, int d, int g_first,
                           int g_count, double *gsums, cudaStream_t s);
void launch_halo_pack(const uint8_t *z, const int32_t *pos, int count, uint8_t *buf, cudaStream_t s);
void launch_halo_unpack(uint8_t *z, const int32_t *pos, int count, const uint8_t *buf, cudaStream_t s);
void launch_snapshot(const uint8_t *z, const double *w, const int32_t *orig, int64_t n, int32_t *z_out,
                     double *w_out, cudaStream_t s, int64_t orig_off = 0);
void launch_snapshot_Y(const double *Yp, int64_t src_ld, int64_t count, int d, const int32_t *orig,
                       const double *ybar, double *out, int64_t dst_ld, cudaStream_t s);
void launch_mode_accumulate(const int32_t *z1, int64_t n, int q, uint32_t row, uint32_t *cnt, uint32_t *first,
                            cudaStream_t s);
void launch_mode_final(const uint32_t *cnt, const uint32_t *first, int64_t n, int q, int32_t *out, cudaStream_t s);
void launch_running_mean(const double *snap, double *mean, int64_t len, double inv_count, cudaStream_t s);
void launch_reduce_groups(const double *partials, const int32_t *group_start, int ngroups, int E,
                          double *gsum, cudaStream_t s, const uint32_t *gepoch = nullptr, size_t par_stride = 0);
void launch_factor(const double *mat, int nl, int d, int form, double *rooti, double *logdet, int *err,
                   cudaStream_t s);
void launch_scatter_probe(const double *stats, const StatLayout &SL, const double *mu, int nl, double *S,
                          cudaStream_t s);
}   // namespace bsb
