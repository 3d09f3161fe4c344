// kernels.hpp -- host-visible launchers of the CUDA kernels.
#pragma once
#include <cuda_runtime.h>

#include "layout.hpp"

namespace bsb {

struct LoglikArgs {
  const double *Y;      // n x d col-major (device)
  int64_t n;
  int q, nl;
  const double *mu;     // q x d row-major
  const double *rooti;  // nl x d x d col-major upper-triangular factor applied to (y - mu)
  const double *logdet; // nl
  const double *w;      // n or nullptr
  int transposed;       // 0: z = R (y-mu)  (Q1 / precision form); 1: z = (y-mu) R (gallery original)
  double *out;          // n x q col-major
};

struct DeconvArgs;   // deconv_kernels.cuh

// Per-dimension kernel table: every kernel that keeps a spot's d PCs in registers is instantiated
// once per d (1..kMaxD) in its own translation unit (sweep_kernels.cu with -DBSB_D=d).
struct KernelsForD {
  void (*sweep)(int tdist, int vvv, const SweepArgs &a, int grid, cudaStream_t s);
  size_t (*sweep_smem)(int tdist, int vvv, const SweepArgs &a);
  int (*sweep_max_blocks_per_sm)(int tdist, int vvv, size_t smem);
  void (*loglik)(const LoglikArgs &a, cudaStream_t s);
  void (*deconv)(const DeconvArgs &a, int grid, cudaStream_t s);
  size_t (*deconv_smem)(const DeconvArgs &a);
  // k_sweep_fast (sweep_fast.cu): shared precision, q <= kFastMaxQ
  void (*sweep_fast)(int tdist, const SweepArgs &a, int grid, cudaStream_t s);
  size_t (*sweep_fast_smem)(int tdist, const SweepArgs &a);
};
constexpr int kFastMaxQ = 16;
constexpr int kSweepPad = 64;   // elements every array a sweep streams is over-allocated by (bulk copies read whole runs)
KernelsForD *kernel_table();   // kMaxD + 1 entries, index = d; null pointers where not built
struct KernelRegistrar {
  KernelRegistrar(int d, const KernelsForD &k);
};
struct FastRegistrar {
  FastRegistrar(int d, void (*launch)(int, const SweepArgs &, int, cudaStream_t), size_t (*smem)(int, const SweepArgs &));
};
// Picks the sweep kernel for (model, precision, q) and launches it.
void launch_sweep_any(const KernelsForD &K, int tdist, int vvv, const SweepArgs &a, int grid, cudaStream_t s);
size_t sweep_smem_any(const KernelsForD &K, int tdist, int vvv, const SweepArgs &a);

void launch_param(const ParamArgs &a, cudaStream_t stream);

// Opts `kernel` into `bytes` of dynamic shared memory on the CURRENT device.  The attribute is per device (per
// context), so the size already granted is remembered per device ordinal (`granted`: one static array per kernel
// instantiation); a failure (more than the device offers) is returned to the caller instead of surfacing as an
// 'invalid argument' of the launch.
template <typename Kernel>
inline cudaError_t ensure_dynamic_smem(Kernel kernel, size_t bytes, size_t (&granted)[64]) {
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return e;
  if (dev >= 0 && dev < 64 && bytes <= granted[dev]) return cudaSuccess;
  e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) bytes);
  if (e == cudaSuccess && dev >= 0 && dev < 64) granted[dev] = bytes;
  return e;
}
// Shared memory one CTA may use on the current device (opt-in maximum), 0 when it cannot be queried.
size_t device_smem_limit();

}   // namespace bsb
