// enhance_utils.cu -- the two helpers of /root/reference/src/compare_enhanced.cpp behind the C ABI:
//   map_subspot2ref  :47-92    nearest reference spot(s) of every subspot (brute force, all exact ties reported)
//   compute_corr     :94-132   row-wise Pearson correlation of two equally shaped matrices
// The reference runs both as OpenMP loops over rows with a custom matrix-append reduction (RedMat, :12-42); here a
// row is a thread (nearest search, reference coordinates tiled through shared memory) or a warp (correlation).
// Distances are sums of squares accumulated in column order with separately rounded multiplies and adds -- the same
// arithmetic as arma::sum(pow(v, 2), 1) -- so that "dists == min(dists)" (:72) finds the same ties.
#include <algorithm>
#include <cstring>
#include <vector>

#include "chain.hpp"

namespace bsb {
namespace {

constexpr int kRefTile = 256;

// pass 1 (idx == nullptr): best[i], ties[i]; pass 2: the tied reference indices, ascending, at idx[start[i]..]
__global__ void __launch_bounds__(256) k_nearest_ref(const double *sub, int64_t ns, const double *ref, int64_t nr, int k,
                                                     double *best, int32_t *ties, const int64_t *start, int32_t *idx) {
  extern __shared__ double s_ref[];   // kRefTile x k, row-major
  const int64_t i = (int64_t) blockIdx.x * blockDim.x + threadIdx.x;
  double x[8];
  for (int c = 0; c < k; c++) x[c] = i < ns ? sub[(size_t) c * ns + i] : 0.0;
  double bmin = idx ? (i < ns ? best[i] : 0.0) : __longlong_as_double(0x7ff0000000000000ll);
  int32_t cnt = 0;
  int64_t at = idx && i < ns ? start[i] : 0;
  for (int64_t j0 = 0; j0 < nr; j0 += kRefTile) {
    const int m = (int) min((int64_t) kRefTile, nr - j0);
    __syncthreads();
    for (int e = threadIdx.x; e < m * k; e += blockDim.x) {
      const int jj = e / k, c = e - jj * k;
      s_ref[e] = ref[(size_t) c * nr + j0 + jj];
    }
    __syncthreads();
    if (i < ns)
      for (int jj = 0; jj < m; jj++) {
        double dist = 0.0;
        for (int c = 0; c < k; c++) {
          const double v = __dsub_rn(x[c], s_ref[jj * k + c]);
          dist = __dadd_rn(dist, __dmul_rn(v, v));   // no FMA contraction: bit-equal to the CPU sum of squares
        }
        if (idx) {
          if (dist == bmin) idx[at++] = (int32_t) (j0 + jj);
        } else if (dist < bmin) {
          bmin = dist;
          cnt = 1;
        } else if (dist == bmin) {
          cnt++;
        }
      }
  }
  if (!idx && i < ns) { best[i] = bmin; ties[i] = cnt; }
}

// one warp per row: cor(m1[i, ], m2[i, ]) = sum (a - abar)(b - bbar) / sqrt(sum (a - abar)^2 sum (b - bbar)^2)
__global__ void __launch_bounds__(256) k_row_corr(const double *m1, const double *m2, int64_t nrow, int64_t ncol, double *out) {
  const int64_t i = ((int64_t) blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (i >= nrow) return;
  double sa = 0.0, sb = 0.0;
  for (int64_t c = lane; c < ncol; c += 32) { sa += m1[(size_t) c * nrow + i]; sb += m2[(size_t) c * nrow + i]; }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) { sa += __shfl_xor_sync(0xffffffffu, sa, o); sb += __shfl_xor_sync(0xffffffffu, sb, o); }
  const double ma = sa / (double) ncol, mb = sb / (double) ncol;
  double sab = 0.0, saa = 0.0, sbb = 0.0;
  for (int64_t c = lane; c < ncol; c += 32) {
    const double a = m1[(size_t) c * nrow + i] - ma, b = m2[(size_t) c * nrow + i] - mb;
    sab += a * b; saa += a * a; sbb += b * b;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    sab += __shfl_xor_sync(0xffffffffu, sab, o);
    saa += __shfl_xor_sync(0xffffffffu, saa, o);
    sbb += __shfl_xor_sync(0xffffffffu, sbb, o);
  }
  if (lane == 0) out[i] = sab / sqrt(saa * sbb);
}

}   // namespace
}   // namespace bsb

using namespace bsb;

#define CKE(call)                                                                                           \
  do {                                                                                                      \
    cudaError_t e_ = (call);                                                                                \
    if (e_ != cudaSuccess)                                                                                  \
      return fail(BSB200_ERR_CUDA, "CUDA error %s at %s:%d", cudaGetErrorString(e_), __FILE__, __LINE__);   \
  } while (0)

extern "C" {

int bsb200_map_subspot2ref(int32_t device, const double *subspot_coords, int64_t ns, const double *ref_coords, int64_t nr,
                           int32_t k, int64_t *row_ptr, int32_t *ref_idx, int64_t cap, double *min_dist2, int64_t *nrows) {
  if (!subspot_coords || !ref_coords || ns <= 0 || nr <= 0 || k < 1 || !nrows) return fail(BSB200_ERR_INVALID, "Invalid arguments!");
  if (k > 8) return fail(BSB200_ERR_UNSUPPORTED, "coordinates with %d columns (max 8)", k);
  int rc = select_device(device);
  if (rc) return rc;
  DevBuf<double> dsub, dref, dbest;
  DevBuf<int32_t> dties, didx;
  DevBuf<int64_t> dstart;
  if ((rc = dsub.alloc((size_t) ns * k)) || (rc = dref.alloc((size_t) nr * k)) || (rc = dbest.alloc((size_t) ns)) ||
      (rc = dties.alloc((size_t) ns)))
    return rc;
  CKE(cudaMemcpy(dsub.p, subspot_coords, sizeof(double) * ns * k, cudaMemcpyHostToDevice));
  CKE(cudaMemcpy(dref.p, ref_coords, sizeof(double) * nr * k, cudaMemcpyHostToDevice));
  const unsigned grid = (unsigned) ((ns + 255) / 256);
  const size_t smem = sizeof(double) * kRefTile * k;
  k_nearest_ref<<<grid, 256, smem>>>(dsub.p, ns, dref.p, nr, k, dbest.p, dties.p, nullptr, nullptr);
  g_launches++;
  std::vector<int32_t> ties((size_t) ns);
  CKE(cudaMemcpy(ties.data(), dties.p, sizeof(int32_t) * ns, cudaMemcpyDeviceToHost));
  std::vector<int64_t> start((size_t) ns + 1, 0);
  for (int64_t i = 0; i < ns; i++) start[(size_t) i + 1] = start[(size_t) i] + ties[(size_t) i];
  *nrows = start[(size_t) ns];
  if (row_ptr) std::memcpy(row_ptr, start.data(), sizeof(int64_t) * (ns + 1));
  if (min_dist2) CKE(cudaMemcpy(min_dist2, dbest.p, sizeof(double) * ns, cudaMemcpyDeviceToHost));
  if (!ref_idx) return BSB200_OK;   // sizing pass
  if (cap < *nrows) return fail(BSB200_ERR_BUFFER, "ref_idx holds %lld entries, %lld needed", (long long) cap, (long long) *nrows);
  if ((rc = didx.alloc((size_t) *nrows)) || (rc = dstart.alloc((size_t) ns + 1))) return rc;
  CKE(cudaMemcpy(dstart.p, start.data(), sizeof(int64_t) * (ns + 1), cudaMemcpyHostToDevice));
  k_nearest_ref<<<grid, 256, smem>>>(dsub.p, ns, dref.p, nr, k, dbest.p, dties.p, dstart.p, didx.p);
  g_launches++;
  CKE(cudaGetLastError());
  CKE(cudaMemcpy(ref_idx, didx.p, sizeof(int32_t) * *nrows, cudaMemcpyDeviceToHost));
  return BSB200_OK;
}

int bsb200_compute_corr(int32_t device, const double *m1, const double *m2, int64_t nrow, int64_t ncol, double *cor) {
  if (!m1 || !m2 || !cor || nrow <= 0 || ncol <= 0) return fail(BSB200_ERR_INVALID, "Input matrices should be of the same shape.");
  int rc = select_device(device);
  if (rc) return rc;
  DevBuf<double> d1, d2, dout;
  if ((rc = d1.alloc((size_t) nrow * ncol)) || (rc = d2.alloc((size_t) nrow * ncol)) || (rc = dout.alloc((size_t) nrow))) return rc;
  CKE(cudaMemcpy(d1.p, m1, sizeof(double) * nrow * ncol, cudaMemcpyHostToDevice));
  CKE(cudaMemcpy(d2.p, m2, sizeof(double) * nrow * ncol, cudaMemcpyHostToDevice));
  k_row_corr<<<(unsigned) ((nrow * 32 + 255) / 256), 256>>>(d1.p, d2.p, nrow, ncol, dout.p);
  g_launches++;
  CKE(cudaGetLastError());
  CKE(cudaMemcpy(cor, dout.p, sizeof(double) * nrow, cudaMemcpyDeviceToHost));
  return BSB200_OK;
}

}   // extern "C"
