// policy.cu — Gavel policies' get_allocation() on the GPU for a pooled (homogeneous) worker pool.
//
// Replaces the cvxpy -> ECOS / Gurobi LP solves of
//   MaxMinFairnessPolicy[WithPerf]        scheduler/policies/max_min_fairness.py:53-113
//   FinishTimeFairnessPolicy[WithPerf]    scheduler/policies/finish_time_fairness.py:66-157
//   MinTotalDurationPolicy[WithPerf]      scheduler/policies/min_total_duration.py:55-135
//   ThroughputNormalizedByCostSumWithPerfSLOs (MST)  scheduler/policies/max_sum_throughput.py:49-108
// for the case every worker type with capacity has the same throughput per job (always true for the
// non-Perf wrappers, which overwrite the matrix with the v100 column / with 1.0, and for Shockwave's
// homogeneous clusters, scheduler/scheduler.py:1128).  With one pooled type of N workers the base
// constraints (policy.py:58-65) read  0 <= x_j <= 1,  sum_j sf_j x_j <= N  and every one of these LPs
// has a closed form or a monotone 1-D search, evaluated here with block reductions in float64.
// The host splits x_j over the real worker types in proportion to their capacities.
#include <math.h>

#include "swb_common.cuh"
#include "swb_internal.h"

namespace swb {

__device__ __forceinline__ unsigned long long pol_order_key(double v) {
  unsigned long long b = (unsigned long long)__double_as_longlong(v);
  return (b & 0x8000000000000000ull) ? ~b : (b | 0x8000000000000000ull);
}

__global__ void __launch_bounds__(1024, 1) policy_kernel(PolicyLaunch L) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  double *red = reinterpret_cast<double *>(smem_raw);
  BlockRed br(red);
  const int J = L.J;
  const double N = L.N;
  double obj = 0.0;
  int status = 0;

  if (L.mode == SWB_POL_MAXMIN) {
    // max z : coef_j x_j >= z   ->  z = min(min_j coef_j, N / sum_j sf_j/coef_j),  x_j = z/coef_j
    double mn = 1e300, s = 0.0;
    for (int j = threadIdx.x; j < J; j += blockDim.x) {
      const double c = L.coef[j];
      mn = fmin(mn, c);
      s += L.sf[j] / c;
    }
    mn = br.min(mn);
    s = br.sum(s);
    const double z = fmin(mn, N / s);
    for (int j = threadIdx.x; j < J; j += blockDim.x) L.x[j] = fmin(1.0, z / L.coef[j]);
    obj = z;
  } else if (L.mode == SWB_POL_FTF) {
    // min rho : (t_j + n_j/(thr_j x_j)) / den_j <= rho   ->  x_j(rho) = n_j / (thr_j (rho den_j - t_j))
    // feasible(rho) <=> every rho den_j > t_j, x_j <= 1, sum sf_j x_j <= N ; monotone in rho
    double lo = 0.0, hi = 1.0;
    auto feasible = [&](double rho, bool write) -> bool {
      double used = 0.0, bad = 0.0;
      for (int j = threadIdx.x; j < J; j += blockDim.x) {
        const double room = rho * L.den[j] - L.t[j];
        double xj = 2.0;
        if (room > 0.0) xj = L.n[j] / (L.coef[j] * room);
        if (xj > 1.0 + 1e-12) bad += 1.0;
        used += L.sf[j] * fmin(xj, 1.0);
        if (write) L.x[j] = fmin(xj, 1.0);
      }
      br.sum2(used, bad);
      return bad == 0.0 && used <= N * (1.0 + 1e-12);
    };
    int it = 0;
    while (!feasible(hi, false) && it < 200) { lo = hi; hi *= 2.0; ++it; }
    if (it >= 200) status = 1;
    for (it = 0; it < 200 && hi - lo > 1e-13 * hi; ++it) {
      const double mid = 0.5 * (lo + hi);
      if (feasible(mid, false)) hi = mid; else lo = mid;
    }
    feasible(hi, true);
    obj = hi;
  } else if (L.mode == SWB_POL_MTD) {
    // the reference's own bisection on T (min_total_duration.py:105-131), LP feasibility in closed form
    auto feasible = [&](double T, bool write) -> bool {
      double used = 0.0, bad = 0.0;
      for (int j = threadIdx.x; j < J; j += blockDim.x) {
        const double xj = L.n[j] / (T * L.coef[j]);
        if (xj > 1.0 + 1e-12) bad += 1.0;
        used += L.sf[j] * fmin(xj, 1.0);
        if (write) L.x[j] = fmin(xj, 1.0);
      }
      br.sum2(used, bad);
      return bad == 0.0 && used <= N * (1.0 + 1e-12);
    };
    double max_T = 1000000.0, min_T = 100.0, last_max_T = max_T, best = -1.0;
    for (int outer = 0; outer < 40 && best < 0.0; ++outer) {
      while (1.05 * min_T < max_T) {
        const double T = (min_T + max_T) / 2.0;
        if (feasible(T, false)) { best = T; max_T = T; } else min_T = T;
      }
      max_T = last_max_T * 10.0;
      min_T = last_max_T;
      last_max_T *= 10.0;
    }
    if (best < 0.0) { status = 1; best = max_T; }
    feasible(best, true);
    obj = best;
  } else if (L.mode == SWB_POL_MAXSUM) {
    // max sum_j v_j x_j : fractional knapsack by v_j/sf_j (coef = v_j), stable by job index
    int npad = 64;
    while (npad < J) npad <<= 1;
    unsigned long long *key = reinterpret_cast<unsigned long long *>(smem_raw + 2 * 64 * sizeof(double));
    unsigned short *idx = reinterpret_cast<unsigned short *>(key + npad);
    double *pref = reinterpret_cast<double *>(idx + npad);
    for (int i = threadIdx.x; i < npad; i += blockDim.x) {
      key[i] = i < J ? pol_order_key(L.coef[i] / L.sf[i]) : 0ull;
      idx[i] = (unsigned short)(i < J ? i : 0xffff);
    }
    __syncthreads();
    for (int k = 2; k <= npad; k <<= 1)
      for (int j = k >> 1; j > 0; j >>= 1) {
        for (int i = threadIdx.x; i < npad; i += blockDim.x) {
          const int p = i ^ j;
          if (p > i) {
            const unsigned long long a = key[i], b = key[p];
            const unsigned short ia = idx[i], ib = idx[p];
            const bool a_first = (a > b) || (a == b && ia < ib);
            if (((i & k) == 0) != a_first) { key[i] = b; key[p] = a; idx[i] = ib; idx[p] = ia; }
          }
        }
        __syncthreads();
      }
    // sequential prefix in sorted order (J <= 8192; one thread, the sums must be order-exact)
    if (threadIdx.x == 0) {
      double used = 0.0, tot = 0.0;
      for (int i = 0; i < J; ++i) {
        const int j = idx[i];
        const double room = N - used;
        double xj = 0.0;
        if (room > 0.0 && L.coef[j] > 0.0) xj = fmin(1.0, room / L.sf[j]);
        used += xj * L.sf[j];
        tot += xj * L.coef[j];
        L.x[j] = xj;
      }
      pref[0] = tot;
    }
    __syncthreads();
    obj = pref[0];
  } else if (L.mode == SWB_POL_ISOLATED) {
    // isolated.py:35-55 / proportional.py:26-43 / gandiva_fair_proportional.py:26-41 (pooled):
    // coef_j carries the per-job divisor (sf_j for Isolated, 1 otherwise); rows are normalised to <= 1
    for (int j = threadIdx.x; j < J; j += blockDim.x) {
      const double v = (N / (double)J) / L.coef[j];
      L.x[j] = v > 1.0 ? 1.0 : v;
    }
    obj = 0.0;
  }
  if (threadIdx.x == 0) { L.out[0] = obj; L.out[1] = (double)status; }
}

cudaError_t launch_policy(const PolicyLaunch &L, cudaStream_t st) {
  int npad = 64;
  while (npad < L.J) npad <<= 1;
  size_t smem = 2 * 64 * sizeof(double) + (size_t)npad * (8 + 2) + 64;
  // function attributes are per device: one flag per device ordinal (one process may drive several GPUs)
  static bool attr_done[64] = {false};
  int dev_ = 0;
  cudaGetDevice(&dev_);
  bool &attr_set = attr_done[dev_ & 63];
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(policy_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    if (e != cudaSuccess) return e;
    attr_set = true;
  }
  policy_kernel<<<1, 1024, smem, st>>>(L);
  return cudaGetLastError();
}

}  // namespace swb
