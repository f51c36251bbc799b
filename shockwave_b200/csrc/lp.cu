// swb_lp_solve: a batch of small sparse linear programs, one CTA each (lp_core.cuh).  The packing policies
// (scheduler/policies/policy.py:68-193 and the *WithPacking classes) hand cvxpy/ECOS one LP over
// (job combination x worker type) columns per get_allocation(); here the host builds the CSC arrays and the whole
// simplex runs on the device.  Programs of one batch share the sparsity pattern (colp, rowi) and differ in the
// values / right-hand sides / costs: that is the shape of a multi-section over a scalar (finish-time fairness:
// S candidate ratios per launch, one CTA per candidate, 148 SMs busy).
#include "swb_internal.h"
#include "lp_core.cuh"

namespace swb {

__global__ void __launch_bounds__(1024, 1) lp_simplex_kernel(LpLaunch L) {
  __shared__ double sv[64];
  __shared__ int si[64];
  const int s = blockIdx.x;
  const size_t m = L.m, n = L.n;
  lp::Problem P;
  P.m = L.m; P.n = L.n; P.colp = L.colp; P.rowi = L.rowi;
  P.val = L.val + (size_t)s * L.nnz;
  P.c = L.c + s * n;
  P.b = L.b + s * m;
  lp::Work W;
  W.Binv = L.Binv + (size_t)s * m * m;
  W.Bm = L.Bm + (size_t)s * m * m;
  double *v = L.vec + (size_t)s * 5 * m;
  W.xB = v; W.y = v + m; W.alpha = v + 2 * m; W.cB = v + 3 * m; W.prow = v + 4 * m;
  W.basis = L.basis + s * m;
  W.where = L.where + s * (n + m + 1);
  W.x = L.x + s * n;
  W.out = L.out + s * 8;
  W.sv = sv; W.si = si;
  lp::simplex(P, W, L.max_iter);
}

cudaError_t launch_lp(const LpLaunch &L, cudaStream_t st) {
  // small programs: fewer threads, cheaper barriers
  int threads = 1024;
  if (L.m <= 96 && L.n <= 2048) threads = 256;
  else if (L.m <= 256 && L.n <= 8192) threads = 512;
  lp_simplex_kernel<<<L.S, threads, 0, st>>>(L);
  return cudaGetLastError();
}

}  // namespace swb
