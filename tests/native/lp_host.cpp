// Host build of shockwave_b200/csrc/lp_core.cuh (one "thread", barriers compiled out): lets the CPU tests run the
// device simplex's pivoting logic against HiGHS without a GPU.  TEST TOOLING ONLY: nothing under shockwave_b200/
// loads this library, and libswb200.so has no CPU path.
#include <stdlib.h>
#include <vector>
#include "../../shockwave_b200/csrc/lp_core.cuh"

extern "C" int lp_host_solve(int m, int n, const int *colp, const int *rowi, const double *val, const double *c,
                             const double *b, int max_iter, double *x, double *out) {
  using namespace swb::lp;
  Problem P{m, n, colp, rowi, val, c, b};
  std::vector<double> Binv((size_t)m * m), Bm((size_t)m * m), vec(5 * (size_t)m), sv(64);
  std::vector<int> basis(m), where(n + m + 1), si(64);
  Work W;
  W.Binv = Binv.data(); W.Bm = Bm.data();
  W.xB = vec.data(); W.y = W.xB + m; W.alpha = W.y + m; W.cB = W.alpha + m; W.prow = W.cB + m;
  W.basis = basis.data(); W.where = where.data(); W.x = x; W.out = out; W.sv = sv.data(); W.si = si.data();
  simplex(P, W, max_iter);
  return (int)out[1];
}
