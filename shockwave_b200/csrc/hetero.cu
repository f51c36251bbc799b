// hetero.cu — Gavel policies' get_allocation() on the GPU for HETEROGENEOUS worker types (W <= 4; max-sum W <= 3).
//
// Replaces the cvxpy -> ECOS / Gurobi LP solves of
//   MaxMinFairnessPolicyWithPerf        scheduler/policies/max_min_fairness.py:53-113
//   FinishTimeFairnessPolicyWithPerf    scheduler/policies/finish_time_fairness.py:66-157
//   MinTotalDurationPolicyWithPerf      scheduler/policies/min_total_duration.py:55-135
//   ThroughputNormalizedByCostSumWithPerf[SLOs]                  scheduler/policies/max_sum_throughput.py:49-108
// when the per-type throughputs of a job differ (k80 / p100 / v100 columns of the throughput table).
//
// All four programs share the base constraints (policy.py:58-65)  x >= 0, sum_w x_jw <= 1, sum_j sf_j x_jw <= N_w
// and differ in one scalar theta that is searched by bisection:
//   MAXMIN  max theta : sum_w a_jw x_jw >= theta                        (a = coef)
//   FTF     min theta : sum_w a_jw x_jw >= n_j / (theta den_j - t_j)     (a = throughput)
//   MTD     the reference's own bisection on T : sum_w a_jw x_jw >= n_j / T
//   MAXSUM  max theta : sum_jw a_jw x_jw >= theta                        (a = throughput / cost)
// For a fixed theta the question "is there an x" couples the jobs only through R <= 4 rows (the W capacity
// rows, plus the value row for MAXSUM).  That is solved by Dantzig-Wolfe decomposition on those rows:
//   * pricing: for a price vector pi on the rows every job picks, independently, its cheapest vertex of
//     { x >= 0, sum_w x_w <= 1, a_j.x >= r_j } (W single-type vertices and W(W-1) two-type mixes that use the
//     whole time) — one pass over the jobs, a block reduction of R sums;  the result is a column e in R^R
//     (relative excess of every row) and the lower bound pi.e on the value of the matrix game below;
//   * master: max_{pi in simplex} min_i pi.e_i over the columns seen so far — a matrix game with R rows, solved
//     EXACTLY by enumerating its bases (k rows x k columns, k <= R) in parallel over the CTA;
//     its column mix mu is a convex combination of job vertices, hence satisfies every job constraint, and is
//     feasible for the coupling rows as soon as  max_w sum_i mu_i e_iw <= 0  (checked directly: certificate);
//   * pi.e > 0 at any pi is a Farkas certificate of infeasibility.
// Every claim the bisection uses is therefore certified; the tolerance only decides when to stop refining a
// theta that is within 1e-12 (relative excess) of critical.  One CTA, float64, no CPU fallback.
#include <math.h>

#include "swb_common.cuh"
#include "swb_internal.h"

namespace swb {

#define HT_MAXR 4
#define HT_MAXC 40
#define HT_TOL 1e-12
#define HT_WF_LOWER (1.0 - 1e-9)

struct HtShared {
  double E[HT_MAXC][HT_MAXR];    // columns: relative excess per row
  double PI[HT_MAXC][HT_MAXR];   // price vector that generated the column
  double pi[HT_MAXR];            // current master prices
  double mu[HT_MAXC];            // current master column mix
  double v;                      // master value
  double e[HT_MAXR];             // last priced column
  double lb;                     // pi.e of the last priced column
  int bad;                       // some job cannot meet its requirement at all
  int ok;                        // master found a base
  // support of the best feasible point so far
  double bPI[HT_MAXR][HT_MAXR], bMU[HT_MAXR], bTheta;
  int bK;
  double tPI[HT_MAXR][HT_MAXR], tMU[HT_MAXR];   // support of the latest feasible check
  int tK;
  double red[2][32][HT_MAXR + 2];
  double candv[32];
  double dbg[12];                // reason counters of the infeasibility claims (diagnostics)
  int candt[32];
};

struct HtCtx {
  const HeteroLaunch &L;
  HtShared &S;
  int R, W, phase;
  double vscale;     // MAXSUM: scale of the value row
  __device__ HtCtx(const HeteroLaunch &l, HtShared &s) : L(l), S(s), phase(0) {}
};

// requirement of job j at parameter theta; returns false when the job can never be satisfied
__device__ __forceinline__ bool ht_requirement(const HeteroLaunch &L, int j, double theta, double &r) {
  if (L.mode == SWB_POL_MAXMIN) { r = theta; return true; }
  if (L.mode == SWB_POL_FTF) {
    const double room = theta * L.den[j] - L.t[j];
    if (!(room > 0.0)) return false;
    r = L.n[j] / room;
    return true;
  }
  if (L.mode == SWB_POL_MTD) { r = L.n[j] / theta; return true; }
  if (L.mode == SWB_POL_WFILL) {   // net_j = a_j.x_j / prop_j >= lower_j + theta / mult_j  (water_filling.py:127-171)
    // the lower bounds carry a relative slack of 1e-9 (the reference's ECOS works to ~1e-8): they are the previous
    // iteration's optimum, i.e. exactly tight against capacity, and must stay feasible under one ulp of re-association
    const double m = L.n[j];
    r = L.den[j] * (L.t[j] * HT_WF_LOWER + (m > 0.0 ? theta / m : 0.0));
    return true;
  }
  r = 0.0;
  return true;
}

__device__ __forceinline__ bool ht_response(const double *a, int W, double r, const double *q, double *x);
__device__ __forceinline__ bool ht_sum_mode(int mode) { return mode == SWB_POL_MAXSUM || mode == SWB_POL_WFZ; }

// Bottleneck detection of the water-filling iteration (max_min_fairness_water_filling.py:191-305): every job keeps
// net_j >= so_far_j; an active job may additionally rise to so_far_j * slack and then counts 1 (the z_j of the
// reference's MILP, relaxed to [0, 1] by mixing the two vertices).  Returns the job's level at the prices (p0, q).
__device__ __forceinline__ bool ht_response_wfz(const HeteroLaunch &L, int j, const double *a, int W, double sf, double p0,
                                                const double *q, double *x, double &level) {
  const double m = L.n[j];
  const double cprev = L.wf_c ? *L.wf_c : 0.0;
  const double r0 = L.den[j] * (L.t[j] * HT_WF_LOWER + (m > 0.0 ? cprev / m : 0.0));
  level = 0.0;
  if (!ht_response(a, W, r0, q, x)) return false;
  if (m > 0.0) {
    double x1[HT_MAXR];
    if (ht_response(a, W, r0 * L.wf_slack, q, x1)) {
      double c0 = 0.0, c1 = 0.0;
      for (int w = 0; w < W; ++w) { c0 += q[w] * x[w]; c1 += q[w] * x1[w]; }
      if (p0 - sf * c1 > -sf * c0) { level = 1.0; for (int w = 0; w < W; ++w) x[w] = x1[w]; }
    }
  }
  return true;
}

// cheapest vertex of job j for per-type prices q (already divided by N_w; the common factor sf_j is dropped).
// Writes the vertex into x[W]; returns false if no vertex meets the requirement.
__device__ __forceinline__ bool ht_response(const double *a, int W, double r, const double *q, double *x) {
  double best = 1e300;
  int bu = -1, bv = -1;
  double bxu = 0.0;
  if (r <= 0.0) { for (int w = 0; w < W; ++w) x[w] = 0.0; return true; }
  for (int u = 0; u < W; ++u) {
    const double au = a[u];
    if (au >= r && au > 0.0) {
      const double xu = r / au, c = q[u] * xu;
      if (c < best) { best = c; bu = u; bv = -1; bxu = xu; }
    }
  }
  for (int u = 0; u < W; ++u)
    for (int v = 0; v < W; ++v) {
      const double au = a[u], av = a[v];
      if (u != v && au > r && av < r && av > 0.0) {
        const double xu = (r - av) / (au - av), c = q[u] * xu + q[v] * (1.0 - xu);
        if (c < best) { best = c; bu = u; bv = v; bxu = xu; }
      }
    }
  for (int w = 0; w < W; ++w) x[w] = 0.0;
  if (bu < 0) return false;
  x[bu] = bxu;
  if (bv >= 0) x[bv] = 1.0 - bxu;
  return true;
}

// MAXSUM vertex: the type with the best (value price * a_w - capacity price * sf) if positive, else idle
__device__ __forceinline__ void ht_response_sum(const double *a, int W, double sf, double p0, const double *q, double *x) {
  double best = 0.0;
  int bu = -1;
  for (int u = 0; u < W; ++u) {
    const double gain = p0 * a[u] - q[u] * sf;
    if (gain > best) { best = gain; bu = u; }
  }
  for (int w = 0; w < W; ++w) x[w] = 0.0;
  if (bu >= 0) x[bu] = 1.0;
}

// MAXSUM with an SLO floor (max_sum_throughput.py:87-93): the job must also reach thr.x >= need.  The best vertex of
// { x >= 0, sum x <= 1, thr.x >= need } for the gains c_w = value price * a_w - capacity price * sf: idle (need = 0),
// full time on one type, the floor met exactly on one type, or both rows tight on two types.  false: no type can
// meet the floor even at full time.
__device__ __forceinline__ bool ht_response_sum_slo(const double *a, const double *thr, int W, double sf, double need,
                                                    double p0, const double *q, double *x) {
  double c[HT_MAXR];
  for (int w = 0; w < W; ++w) c[w] = p0 * a[w] - q[w] * sf;
  double best = -1e300;
  int bu = -1, bv = -1;
  double bxu = 0.0;
  if (need <= 0.0) { best = 0.0; }
  for (int u = 0; u < W; ++u) {
    if (thr[u] >= need && thr[u] > 0.0) {
      if (c[u] > best) { best = c[u]; bu = u; bv = -1; bxu = 1.0; }                  // full time on u
      if (need > 0.0) {
        const double xu = need / thr[u], gn = c[u] * xu;                             // floor met exactly on u
        if (gn > best) { best = gn; bu = u; bv = -1; bxu = xu; }
      }
    }
  }
  if (need > 0.0)
    for (int u = 0; u < W; ++u)
      for (int v = 0; v < W; ++v)
        if (u != v && thr[u] > need && thr[v] < need) {                              // both rows tight
          const double xu = (need - thr[v]) / (thr[u] - thr[v]), gn = c[u] * xu + c[v] * (1.0 - xu);
          if (gn > best) { best = gn; bu = u; bv = v; bxu = xu; }
        }
  for (int w = 0; w < W; ++w) x[w] = 0.0;
  if (best <= -1e299) return false;
  if (bu >= 0) { x[bu] = bxu; if (bv >= 0) x[bv] = 1.0 - bxu; }
  return true;
}

// block-wide sum of n <= HT_MAXR + 2 doubles per thread (one barrier, ping-pong scratch)
__device__ __forceinline__ void ht_block_sum(HtCtx &C, double *v, int n) {
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31, nw = (blockDim.x + 31) >> 5;
  for (int i = 0; i < n; ++i) v[i] = warp_sum(v[i]);
  if (l == 0) for (int i = 0; i < n; ++i) C.S.red[C.phase][w][i] = v[i];
  __syncthreads();
  for (int i = 0; i < n; ++i) v[i] = warp_sum(l < nw ? C.S.red[C.phase][l][i] : 0.0);
  C.phase ^= 1;
}

// pricing pass at S.pi: fills S.e, S.lb, S.bad (block-uniform after return)
template <int W, int R>
__device__ __noinline__ void ht_price(HtCtx &C, double theta) {
  const HeteroLaunch &L = C.L;
  double q[HT_MAXR], acc[HT_MAXR + 2];
  const bool sum_mode = ht_sum_mode(L.mode);
  for (int w = 0; w < W; ++w) q[w] = C.S.pi[sum_mode ? w + 1 : w] / L.N[w];
  const double p0 = sum_mode ? C.S.pi[0] / C.vscale : 0.0;
  for (int i = 0; i < HT_MAXR + 2; ++i) acc[i] = 0.0;
  for (int j = threadIdx.x; j < L.J; j += blockDim.x) {
    double a[HT_MAXR], x[HT_MAXR];
    for (int w = 0; w < W; ++w) a[w] = L.a[(size_t)j * W + w];
    const double sf = L.sf[j];
    if (L.mode == SWB_POL_WFZ) {
      double level;
      if (!ht_response_wfz(L, j, a, W, sf, p0, q, x, level)) acc[HT_MAXR] += 1.0;
      else { acc[0] += level; for (int w = 0; w < W; ++w) acc[w + 1] += sf * x[w]; }
    } else if (sum_mode) {
      bool okj = true;
      if (L.t) {      // SLO floors: t = needed throughput, den = instance cost per type (thr = a * cost)
        double thr[HT_MAXR];
        for (int w = 0; w < W; ++w) thr[w] = a[w] * L.den[w];
        okj = ht_response_sum_slo(a, thr, W, sf, L.t[j], p0, q, x);
      } else {
        ht_response_sum(a, W, sf, p0, q, x);
      }
      if (!okj) acc[HT_MAXR] += 1.0;
      double val = 0.0;
      for (int w = 0; w < W; ++w) { val += a[w] * x[w]; acc[w + 1] += sf * x[w]; }
      acc[0] += val;
    } else {
      double r;
      bool okj = ht_requirement(L, j, theta, r);
      if (okj) okj = ht_response(a, W, r, q, x);
      if (!okj) acc[HT_MAXR] += 1.0;
      else for (int w = 0; w < W; ++w) acc[w] += sf * x[w];
    }
  }
  ht_block_sum(C, acc, HT_MAXR + 1);
  if (threadIdx.x == 0) {
    double lb = 0.0;
    if (sum_mode) {
      C.S.e[0] = (theta - acc[0]) / C.vscale;
      for (int w = 0; w < W; ++w) C.S.e[w + 1] = acc[w + 1] / L.N[w] - 1.0;
    } else {
      for (int w = 0; w < W; ++w) C.S.e[w] = acc[w] / L.N[w] - 1.0;
    }
    for (int r = 0; r < R; ++r) lb += C.S.pi[r] * C.S.e[r];
    C.S.lb = lb;
    C.S.bad = acc[HT_MAXR] > 0.0 ? 1 : 0;
  }
  __syncthreads();
}

// Gaussian elimination with partial pivoting of the n x n system in A (augmented, n+1 columns); false if singular
__device__ __forceinline__ bool ht_solve(double A[HT_MAXR + 1][HT_MAXR + 2], int n, double *sol) {
  for (int c = 0; c < n; ++c) {
    int pr = c;
    double pv = fabs(A[c][c]);
    for (int r = c + 1; r < n; ++r) if (fabs(A[r][c]) > pv) { pv = fabs(A[r][c]); pr = r; }
    if (pv < 1e-300) return false;
    if (pr != c) for (int k = c; k <= n; ++k) { const double t = A[c][k]; A[c][k] = A[pr][k]; A[pr][k] = t; }
    const double inv = 1.0 / A[c][c];
    for (int r = c + 1; r < n; ++r) {
      const double f = A[r][c] * inv;
      if (f != 0.0) for (int k = c; k <= n; ++k) A[r][k] -= f * A[c][k];
    }
  }
  for (int r = n - 1; r >= 0; --r) {
    double s = A[r][n];
    for (int k = r + 1; k < n; ++k) s -= A[r][k] * sol[k];
    sol[r] = s / A[r][r];
  }
  return true;
}

__device__ __forceinline__ int ht_choose(int n, int k) {
  if (k < 0 || k > n) return 0;
  if (k == 0) return 1;
  if (k == 1) return n;
  if (k == 2) return n * (n - 1) / 2;
  if (k == 3) return n * (n - 1) * (n - 2) / 6;
  return n * (n - 1) * (n - 2) / 6 * (n - 3) / 4;
}

// the q-th k-subset (k <= 4) of {0..n-1} in colex order
__device__ __forceinline__ void ht_unrank(int q, int k, int *c) {
  for (int i = k; i >= 1; --i) {
    int x = i - 1;
    while (ht_choose(x + 1, i) <= q) ++x;
    c[i - 1] = x;
    q -= ht_choose(x, i);
  }
}

// master: exact solution of the matrix game over the M columns in S.E.  `forced` >= 0: only bases that contain
// that column are enumerated (the optimum must contain the column that was just added).  Fills S.pi, S.mu, S.v, S.ok.
template <int R>
__device__ __noinline__ void ht_master(HtCtx &C, int M, int forced) {
  HtShared &S = C.S;
  double bestv = -1e300;
  int bk = 0, bcols[HT_MAXR];
  double bpi[HT_MAXR], bmu[HT_MAXR];
  const int Mo = forced >= 0 ? M - 1 : M;                   // columns to choose the rest of the base from
  for (int k = 1; k <= R && k <= M; ++k) {
    const int kc = forced >= 0 ? k - 1 : k;                 // columns still to choose
    const int ncs = ht_choose(Mo, kc);
    int nrs = 0;
    for (int m = 1; m < (1 << R); ++m) nrs += (__popc(m) == k) ? 1 : 0;
    const int total = ncs * nrs;
    for (int q = threadIdx.x; q < total; q += blockDim.x) {
      const int cs = q / nrs, rsi = q - cs * nrs;
      int rowmask = 0;
      for (int m = 1, seen = 0; m < (1 << R); ++m)
        if (__popc(m) == k) { if (seen == rsi) { rowmask = m; break; } ++seen; }
      int cols[HT_MAXR];
      ht_unrank(cs, kc, cols);
      if (forced >= 0) {
        for (int i = 0; i < kc; ++i) if (cols[i] >= forced) cols[i] += 1;   // skip the forced index
        cols[kc] = forced;
      }
      int rows[HT_MAXR], nr = 0;
      for (int r = 0; r < R; ++r) if (rowmask >> r & 1) rows[nr++] = r;
      // primal: pi on `rows` with pi.e_c = v for the base columns, sum pi = 1
      double A[HT_MAXR + 1][HT_MAXR + 2], sol[HT_MAXR + 1];
      for (int i = 0; i < k; ++i) {
        for (int r = 0; r < k; ++r) A[i][r] = S.E[cols[i]][rows[r]];
        A[i][k] = -1.0; A[i][k + 1] = 0.0;
      }
      for (int r = 0; r < k; ++r) A[k][r] = 1.0;
      A[k][k] = 0.0; A[k][k + 1] = 1.0;
      if (!ht_solve(A, k + 1, sol)) continue;
      double pi[HT_MAXR];
      bool good = true;
      for (int r = 0; r < R; ++r) pi[r] = 0.0;
      for (int r = 0; r < k; ++r) { pi[rows[r]] = sol[r]; if (sol[r] < -1e-12) good = false; }
      if (!good) continue;
      const double v = sol[k];
      if (!(v > bestv)) continue;
      const double slack = 1e-11 * fmax(1.0, fabs(v));
      for (int i = 0; i < M && good; ++i) {
        double d = 0.0;
        for (int r = 0; r < R; ++r) d += pi[r] * S.E[i][r];
        if (d < v - slack) good = false;
      }
      if (!good) continue;
      // dual: mu on the base columns with sum_i mu_i e_i[row] = v' on `rows`, sum mu = 1
      for (int r = 0; r < k; ++r) {
        for (int i = 0; i < k; ++i) A[r][i] = S.E[cols[i]][rows[r]];
        A[r][k] = -1.0; A[r][k + 1] = 0.0;
      }
      for (int i = 0; i < k; ++i) A[k][i] = 1.0;
      A[k][k] = 0.0; A[k][k + 1] = 1.0;
      double sm[HT_MAXR + 1];
      if (!ht_solve(A, k + 1, sm)) continue;
      for (int i = 0; i < k; ++i) if (sm[i] < -1e-9) good = false;
      // ... and the rows outside the base (pi = 0 there) must not exceed the value under this column mix; a base that
      // passes both tests is primal AND dual feasible, hence optimal, and its mix certifies  max_w sum mu e = v
      for (int r = 0; r < R && good; ++r) {
        if (rowmask >> r & 1) continue;
        double d = 0.0;
        for (int i = 0; i < k; ++i) d += sm[i] * S.E[cols[i]][r];
        if (d > v + slack) good = false;
      }
      if (!good) continue;
      bestv = v; bk = k;
      double tot = 0.0;
      for (int i = 0; i < k; ++i) { bcols[i] = cols[i]; bmu[i] = fmax(sm[i], 0.0); tot += bmu[i]; }
      for (int i = 0; i < k; ++i) bmu[i] /= tot;
      for (int r = 0; r < R; ++r) bpi[r] = fmax(pi[r], 0.0);
    }
  }
  // block arg-max (lowest thread wins ties)
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31, nw = (blockDim.x + 31) >> 5;
  double wv = warp_max(bestv);
  unsigned int bal = __ballot_sync(SWB_FULL, bestv == wv);
  const int wt = (w << 5) + (__ffs(bal) - 1);
  __syncthreads();
  if (l == 0) { S.candv[w] = wv; S.candt[w] = wt; }
  __syncthreads();
  double gv = -1e300;
  int gt = 0;
  for (int i = 0; i < nw; ++i) if (S.candv[i] > gv) { gv = S.candv[i]; gt = S.candt[i]; }
  if (threadIdx.x == 0) S.ok = gv > -1e299 ? 1 : 0;
  if ((int)threadIdx.x == gt && gv > -1e299) {
    double tot = 0.0;
    for (int r = 0; r < R; ++r) tot += bpi[r];
    for (int r = 0; r < R; ++r) S.pi[r] = bpi[r] / tot;
    for (int i = 0; i < M; ++i) S.mu[i] = 0.0;
    for (int i = 0; i < bk; ++i) S.mu[bcols[i]] = bmu[i];
    S.v = bestv;
  }
  __syncthreads();
}

// is there an x at parameter theta?  On success the support (<= R columns) is left in S.tPI / S.tMU / S.tK.
template <int W, int R>
__device__ __noinline__ bool ht_feasible(HtCtx &C, double theta, int &rounds) {
  HtShared &S = C.S;
  int M = 0;      // columns; every thread keeps the same count
  __syncthreads();
  // seeds: the prices that refuted / supported the previous theta, the prices of the support of the best feasible
  // point so far (near the optimum the same job vertices stay optimal, so these columns usually certify the new
  // theta at once), the unit vectors, the uniform vector
  double keep[HT_MAXR];
  for (int r = 0; r < R; ++r) keep[r] = S.pi[r];
  const int nb = S.bK;
  __syncthreads();
  for (int sd = 0; sd < nb + R + 2; ++sd) {
    if (threadIdx.x == 0) {
      for (int r = 0; r < R; ++r) {
        double p;
        if (sd == 0) p = keep[r];
        else if (sd <= nb) p = S.bPI[sd - 1][r];
        else if (sd <= nb + R) p = (r == sd - nb - 1) ? 1.0 : 0.0;
        else p = 1.0 / R;
        S.pi[r] = p;
      }
    }
    __syncthreads();
    ht_price<W, R>(C, theta);
    ++rounds;
    if (S.bad) { if (threadIdx.x == 0) S.dbg[0] += 1; return false; }
    if (S.lb > HT_TOL) { if (threadIdx.x == 0) S.dbg[1] += 1; return false; }   // Farkas certificate: no x at this theta
    bool done = true;
    for (int r = 0; r < R; ++r) if (S.e[r] > 0.0) done = false;
    if (threadIdx.x == 0) {
      for (int r = 0; r < R; ++r) { S.E[M][r] = S.e[r]; S.PI[M][r] = S.pi[r]; }
      if (done) { S.tK = 1; S.tMU[0] = 1.0; for (int r = 0; r < R; ++r) S.tPI[0][r] = S.pi[r]; }
    }
    ++M;
    __syncthreads();
    if (done) return true;                    // this single vertex already fits every row
    if (nb > 0 && sd == nb) {
      // early master over {last prices, previous support}
      ht_master<R>(C, M, -1);
      if (S.ok) {
        double U = -1e300;
        for (int r = 0; r < R; ++r) {
          double t = 0.0;
          for (int i = 0; i < M; ++i) t += S.mu[i] * S.E[i][r];
          U = fmax(U, t);
        }
        if (U <= 0.0) {
          __syncthreads();
          if (threadIdx.x == 0) {
            int k = 0;
            for (int i = 0; i < M; ++i)
              if (S.mu[i] > 0.0 && k < HT_MAXR) { S.tMU[k] = S.mu[i]; for (int r = 0; r < R; ++r) S.tPI[k][r] = S.PI[i][r]; ++k; }
            S.tK = k;
          }
          __syncthreads();
          return true;
        }
      }
      __syncthreads();
    }
  }
  int forced = -1;
  for (int it = 0; it < 400; ++it) {
    ht_master<R>(C, M, forced);
    if (!S.ok) {
      if (forced >= 0) { forced = -1; continue; }    // numerical trouble with the restricted search: full search
      if (threadIdx.x == 0) S.dbg[2] += 1;
      return false;
    }
    // certificate: the column mix fits every row
    double U = -1e300;
    for (int r = 0; r < R; ++r) {
      double s = 0.0;
      for (int i = 0; i < M; ++i) s += S.mu[i] * S.E[i][r];
      U = fmax(U, s);
    }
    if (U <= 0.0) {
      if (threadIdx.x == 0) {
        int k = 0;
        for (int i = 0; i < M; ++i)
          if (S.mu[i] > 0.0 && k < HT_MAXR) { S.tMU[k] = S.mu[i]; for (int r = 0; r < R; ++r) S.tPI[k][r] = S.PI[i][r]; ++k; }
        S.tK = k;
      }
      __syncthreads();
      return true;
    }
    const double v = S.v;
    ht_price<W, R>(C, theta);
    ++rounds;
    if (S.lb > HT_TOL) { if (threadIdx.x == 0) S.dbg[3] += 1; return false; }   // Farkas certificate
    if (v - S.lb <= HT_TOL) {
      // the master value and the pricing bound meet: no x.  The restricted base search is only guaranteed to find
      // the optimum when the master optimum was unique, so the claim is confirmed with the full search first
      // (the column just priced is not added: at this point it cuts nothing).
      if (forced >= 0) { forced = -1; continue; }
      if (threadIdx.x == 0) { S.dbg[4] += 1; S.dbg[6] = v; S.dbg[7] = S.lb; S.dbg[8] = theta; S.dbg[9] = M; }
      return false;
    }
    if (threadIdx.x == 0) {
      int m = M;
      if (m == HT_MAXC) {
        // evict the column with the largest slack at the current prices (never a base column: their slack is 0)
        int worst = 0; double ws = -1.0;
        for (int i = 0; i < m; ++i) {
          if (S.mu[i] > 0.0) continue;
          double d = 0.0;
          for (int r = 0; r < R; ++r) d += S.pi[r] * S.E[i][r];
          if (d - v > ws) { ws = d - v; worst = i; }
        }
        for (int r = 0; r < R; ++r) { S.E[worst][r] = S.E[m - 1][r]; S.PI[worst][r] = S.PI[m - 1][r]; }
        S.mu[worst] = S.mu[m - 1];
        m = m - 1;
      }
      for (int r = 0; r < R; ++r) { S.E[m][r] = S.e[r]; S.PI[m][r] = S.pi[r]; }
    }
    if (M < HT_MAXC) ++M;                     // at the cap one column was evicted and the new one took the last slot
    __syncthreads();
    forced = M - 1;
  }
  if (threadIdx.x == 0) S.dbg[5] += 1;
  return false;
}

__device__ void ht_accept(HtCtx &C, double theta) {
  HtShared &S = C.S;
  __syncthreads();
  if (threadIdx.x == 0) {
    S.bK = S.tK; S.bTheta = theta;
    for (int i = 0; i < S.tK; ++i) { S.bMU[i] = S.tMU[i]; for (int r = 0; r < C.R; ++r) S.bPI[i][r] = S.tPI[i][r]; }
  }
  __syncthreads();
}

template <int W, int R>
__global__ void __launch_bounds__(1024, 1) hetero_kernel(HeteroLaunch L) {
  __shared__ HtShared S;
  HtCtx C(L, S);
  const int J = L.J;
  C.W = W;
  C.R = R;
  C.vscale = 1.0;
  if (threadIdx.x == 0) {
    S.bK = 0; S.bTheta = 0.0;
    for (int i = 0; i < 12; ++i) S.dbg[i] = 0.0;
    for (int r = 0; r < C.R; ++r) S.pi[r] = 1.0 / C.R;
  }
  __syncthreads();
  int rounds = 0, checks = 0, status = 0;
  double obj = 0.0;

  if (L.mode == SWB_POL_MAXMIN) {
    // theta <= min_j max_w a_jw (a job alone cannot exceed its best type at full time)
    double m = 1e300;
    for (int j = threadIdx.x; j < J; j += blockDim.x) {
      double b = 0.0;
      for (int w = 0; w < W; ++w) b = fmax(b, L.a[(size_t)j * W + w]);
      m = fmin(m, b);
    }
    { BlockRed br(&S.red[0][0][0]); m = br.min(m); __syncthreads(); }
    double lo = 0.0, hi = m;
    ++checks;
    if (hi > 0.0 && ht_feasible<W, R>(C, hi, rounds)) { ht_accept(C, hi); lo = hi; }
    for (int it = 0; it < 200 && hi - lo > 1e-12 * hi; ++it) {
      const double mid = 0.5 * (lo + hi);
      ++checks;
      if (ht_feasible<W, R>(C, mid, rounds)) { ht_accept(C, mid); lo = mid; } else hi = mid;
    }
    obj = lo;
  } else if (L.mode == SWB_POL_FTF) {
    double lo = 0.0, hi = 1.0;
    int it = 0;
    bool found = false;
    for (; it < 200; ++it) {
      ++checks;
      if (ht_feasible<W, R>(C, hi, rounds)) { ht_accept(C, hi); found = true; break; }
      lo = hi; hi *= 2.0;
    }
    if (!found) status = 1;
    for (it = 0; found && it < 200 && hi - lo > 1e-12 * hi; ++it) {
      const double mid = 0.5 * (lo + hi);
      ++checks;
      if (ht_feasible<W, R>(C, mid, rounds)) { ht_accept(C, mid); hi = mid; } else lo = mid;
    }
    obj = hi;
  } else if (L.mode == SWB_POL_MTD) {
    // the reference's own bisection on T (min_total_duration.py:105-131)
    double max_T = 1000000.0, min_T = 100.0, last_max_T = max_T, best = -1.0;
    for (int outer = 0; outer < 40 && best < 0.0; ++outer) {
      while (1.05 * min_T < max_T) {
        const double T = (min_T + max_T) / 2.0;
        ++checks;
        if (ht_feasible<W, R>(C, T, rounds)) { ht_accept(C, T); best = T; max_T = T; } else min_T = T;
      }
      max_T = last_max_T * 10.0;
      min_T = last_max_T;
      last_max_T *= 10.0;
    }
    if (best < 0.0) { status = 1; best = max_T; }
    obj = best;
  } else if (L.mode == SWB_POL_WFILL) {
    // max theta <= M with net_j >= lower_j + theta / mult_j for the active jobs: theta cannot pass the point where an
    // active job would need more than its best type at full time
    double m = L.wf_M;
    for (int j = threadIdx.x; j < J; j += blockDim.x) {
      const double mj = L.n[j];
      if (mj > 0.0) {
        double b = 0.0;
        for (int w = 0; w < W; ++w) b = fmax(b, L.a[(size_t)j * W + w]);
        m = fmin(m, (b / L.den[j] - L.t[j]) * mj);
      }
    }
    { BlockRed br(&S.red[0][0][0]); m = br.min(m); __syncthreads(); }
    double lo = 0.0, hi = fmax(m, 0.0);
    ++checks;
    if (!ht_feasible<W, R>(C, 0.0, rounds)) {
      status = 1;
    } else {
      ht_accept(C, 0.0);
      ++checks;
      if (hi > 0.0 && ht_feasible<W, R>(C, hi, rounds)) { ht_accept(C, hi); lo = hi; }
      for (int it = 0; it < 200 && hi - lo > 1e-12 * hi; ++it) {
        const double mid = 0.5 * (lo + hi);
        ++checks;
        if (ht_feasible<W, R>(C, mid, rounds)) { ht_accept(C, mid); lo = mid; } else hi = mid;
      }
    }
    obj = lo;
  } else if (L.mode == SWB_POL_WFZ) {
    double na = 0.0;
    for (int j = threadIdx.x; j < J; j += blockDim.x) na += (L.n[j] > 0.0) ? 1.0 : 0.0;
    { BlockRed br(&S.red[0][0][0]); na = br.sum(na); __syncthreads(); }
    C.vscale = na > 0.0 ? na : 1.0;
    double lo = 0.0, hi = na;
    ++checks;
    if (!ht_feasible<W, R>(C, 0.0, rounds)) {
      status = 1;
    } else {
      ht_accept(C, 0.0);
      ++checks;
      if (hi > 0.0 && ht_feasible<W, R>(C, hi, rounds)) { ht_accept(C, hi); lo = hi; }
      for (int it = 0; it < 200 && hi - lo > 1e-7 * fmax(hi, 1.0); ++it) {
        const double mid = 0.5 * (lo + hi);
        ++checks;
        if (ht_feasible<W, R>(C, mid, rounds)) { ht_accept(C, mid); lo = mid; } else hi = mid;
      }
    }
    obj = lo;
  } else {  // MAXSUM
    double ub = 0.0;
    for (int j = threadIdx.x; j < J; j += blockDim.x) {
      double b = 0.0;
      for (int w = 0; w < W; ++w) b = fmax(b, L.a[(size_t)j * W + w]);
      ub += b;
    }
    { BlockRed br(&S.red[0][0][0]); ub = br.sum(ub); __syncthreads(); }
    C.vscale = ub > 0.0 ? ub : 1.0;
    double lo = 0.0, hi = ub;
    bool base_ok = true;
    if (L.t) {          // with SLO floors even theta = 0 has to be shown feasible
      ++checks;
      base_ok = ht_feasible<W, R>(C, 0.0, rounds);
      if (base_ok) ht_accept(C, 0.0); else status = 1;
    }
    if (base_ok) {
      ++checks;
      if (ht_feasible<W, R>(C, hi, rounds)) { ht_accept(C, hi); lo = hi; }
      for (int it = 0; it < 200 && hi - lo > 1e-12 * hi; ++it) {
        const double mid = 0.5 * (lo + hi);
        ++checks;
        if (ht_feasible<W, R>(C, mid, rounds)) { ht_accept(C, mid); lo = mid; } else hi = mid;
      }
    }
    obj = lo;
  }

  // ---- x = sum_i mu_i * (vertex chosen by every job at the prices of support column i) -----------------
  __syncthreads();
  const int K = S.bK;
  const double theta = S.bTheta;
  const bool sum_mode = ht_sum_mode(L.mode);
  for (int j = threadIdx.x; j < J; j += blockDim.x) {
    double a[HT_MAXR], out[HT_MAXR];
    for (int w = 0; w < W; ++w) { a[w] = L.a[(size_t)j * W + w]; out[w] = 0.0; }
    const double sf = L.sf[j];
    double zj = 0.0;
    for (int i = 0; i < K; ++i) {
      double q[HT_MAXR], x[HT_MAXR];
      for (int w = 0; w < W; ++w) q[w] = S.bPI[i][sum_mode ? w + 1 : w] / L.N[w];
      if (L.mode == SWB_POL_WFZ) {
        double level = 0.0;
        ht_response_wfz(L, j, a, W, sf, S.bPI[i][0] / C.vscale, q, x, level);
        zj += S.bMU[i] * level;
      } else if (sum_mode) {
        if (L.t) {
          double thr[HT_MAXR];
          for (int w = 0; w < W; ++w) thr[w] = a[w] * L.den[w];
          ht_response_sum_slo(a, thr, W, sf, L.t[j], S.bPI[i][0] / C.vscale, q, x);
        } else {
          ht_response_sum(a, W, sf, S.bPI[i][0] / C.vscale, q, x);
        }
      } else {
        double r = 0.0;
        ht_requirement(L, j, theta, r);
        ht_response(a, W, r, q, x);
      }
      for (int w = 0; w < W; ++w) out[w] += S.bMU[i] * x[w];
    }
    for (int w = 0; w < W; ++w) L.x[(size_t)j * W + w] = out[w];
    if (L.mode == SWB_POL_WFZ && L.zout) L.zout[j] = zj;
  }
  if (threadIdx.x == 0) { L.out[0] = obj; L.out[1] = (double)status; L.out[2] = (double)rounds; L.out[3] = (double)checks;
    for (int i = 0; i < 12; ++i) L.out[4 + i] = S.dbg[i]; }
}

cudaError_t launch_hetero(const HeteroLaunch &L, cudaStream_t st) {
  const bool sum = (L.mode == SWB_POL_MAXSUM || L.mode == SWB_POL_WFZ);
  switch (L.W) {
    case 1: if (sum) hetero_kernel<1, 2><<<1, 1024, 0, st>>>(L); else hetero_kernel<1, 1><<<1, 1024, 0, st>>>(L); break;
    case 2: if (sum) hetero_kernel<2, 3><<<1, 1024, 0, st>>>(L); else hetero_kernel<2, 2><<<1, 1024, 0, st>>>(L); break;
    case 3: if (sum) hetero_kernel<3, 4><<<1, 1024, 0, st>>>(L); else hetero_kernel<3, 3><<<1, 1024, 0, st>>>(L); break;
    case 4: if (sum) return cudaErrorInvalidValue; hetero_kernel<4, 4><<<1, 1024, 0, st>>>(L); break;
    default: return cudaErrorInvalidValue;
  }
  return cudaGetLastError();
}

}  // namespace swb
