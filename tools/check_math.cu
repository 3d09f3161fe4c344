// Accuracy of the hand-written elementary functions of the sweep (rng.cuh: cos2pi_unit, sqrt_small, rcp_moderate)
// against CUDA's cospi(), sqrt() and division on Philox-uniform inputs.
//   nvcc -gencode arch=compute_100a,code=sm_100a -I bayesspace_b200/csrc -o /tmp/check_math tools/check_math.cu && /tmp/check_math
#include <cstdio>
#include "rng.cuh"
using namespace bsb;
__global__ void k(double *err, int n) {
  double ec = 0, es = 0, er = 0;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const uint4 o = philox4x32_10((uint32_t) i, 0u, 1u, 77u, Key{1u, 2u});
    const double u0 = u53(o.x, o.y), u1 = u53(o.z, o.w);
    ec = fmax(ec, fabs(cos2pi_unit(u1) - cospi(2.0 * u1)));
    double sn, cs;
    sincos2pi_unit(u1, sn, cs);
    ec = fmax(ec, fmax(fabs(sn - sinpi(2.0 * u1)), fabs(cs - cospi(2.0 * u1))));
    const double a = -2.0 * log(u0) * (i % 3 == 0 ? 1e-12 : (i % 3 == 1 ? 1.0 : 30.0));
    es = fmax(es, fabs(sqrt_small(a) - sqrt(a)) / sqrt(a));
    const double x = 2.0 + a * a * a;
    er = fmax(er, fabs(rcp_moderate(x) - 1.0 / x) * x);
  }
  atomicMax((unsigned long long *) &err[0], (unsigned long long) __double_as_longlong(ec));
  atomicMax((unsigned long long *) &err[1], (unsigned long long) __double_as_longlong(es));
  atomicMax((unsigned long long *) &err[2], (unsigned long long) __double_as_longlong(er));
}
int main() {
  double *err, h[3];
  cudaMalloc(&err, 24);
  cudaMemset(err, 0, 24);
  k<<<592, 256>>>(err, 1 << 26);
  cudaMemcpy(h, err, 24, cudaMemcpyDeviceToHost);
  printf("max |cos2pi_unit / sincos2pi_unit - cospi / sinpi| = %.3g   max rel |sqrt_small - sqrt| = %.3g   max rel |rcp_moderate - 1/x| = %.3g  (%s)\n", h[0], h[1], h[2],
         cudaGetErrorString(cudaGetLastError()));
  return !(h[0] < 1e-15 && h[1] < 4e-16 && h[2] < 4e-16);
}
