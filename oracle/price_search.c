/* CPU ORACLE / BASELINE (test infrastructure, not product code): plain-C restatement of the ALGORITHM of
 * shockwave_b200/csrc/solve.cu — the collapse of the reference MILP (scheduler/shockwave.py:504-711) to a concave
 * allocation over per-job round counts, solved by a clearing price nested in a search over the makespan threshold —
 * on host cores, float64 throughout, one scenario per OpenMP thread.
 *
 * Purpose: (1) `bench.py`'s `same_algorithm_cpu` figure, which separates what the B200 contributes from what the
 * algorithm contributes (the reference-vs-GPU ratio mixes both); (2) a second, independent implementation of the
 * count solve for tests/test_price_search_port.py (objective of the counts vs the HiGHS oracle and vs the kernel).
 * Follows solve.cu phase by phase (constants + FTF rows :phase 1, priorities shockwave.py:830-911, price bisection on
 * the bit pattern, makespan search, tie fill + completion); the fp32 response table of the kernel is replaced by
 * direct float64 marginal utilities, so prices inside a tie interval may resolve differently (same objective within
 * the reference's gap).  Only tests/ and bench.py's CPU legs may load it.
 *
 * Build: gcc -O3 -fopenmp -shared -fPIC oracle/price_search.c -o oracle/libprice_search.so -lm   (done by build()).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef struct {
  int J, T, G, B, round_ptr;
  double D, k, lam, rhomax;
  const double *bases, *logv;      /* B */
  const int32_t *g, *E, *c;        /* J */
  const double *dbar, *rem, *ftobj;
  /* work */
  double *a, *u0, *cap, *ws, *R;
  int *nmax, *nF;
  double slope[16];
} Scn;

static double plog(const Scn *s, double u) {
  if (u < 0) u = 0;
  if (u >= s->bases[s->B - 1]) return s->logv[s->B - 1];
  int b = 0;
  for (int i = 1; i < s->B - 1; ++i) if (u >= s->bases[i]) b = i;
  return s->slope[b] * (u - s->bases[b]) + s->logv[b];
}
static double util_of(const Scn *s, int j, int n) {
  double u = (s->D * n >= s->cap[j]) ? 1.0 : s->a[j] * n + s->u0[j];
  return s->ws[j] * plog(s, u);
}
static double rem_of(const Scn *s, int j, int n) {
  double done = fmin(s->D * n, s->cap[j]);
  return fmax(0.0, s->R[j] - done);
}
/* largest n with marginal utility per GPU above mu (utilities are concave in n) */
static int pref_n(const Scn *s, int j, double mu) {
  if (mu <= 0.0) return s->nmax[j];
  int lo = 0, hi = s->nmax[j];         /* invariant: marginal(lo) > mu*g or lo == 0 ; marginal(hi+1) <= mu*g */
  const double thr = mu * s->g[j];
  while (lo < hi) {
    int mid = (lo + hi + 1) >> 1;
    if (util_of(s, j, mid) - util_of(s, j, mid - 1) > thr) lo = mid; else hi = mid - 1;
  }
  return lo;
}
static int lower_n(const Scn *s, int j, double M) {
  double need = (s->R[j] - M) / s->D;
  int L = 0;
  if (need > 0.0) { L = (int)ceil(need - 1e-9); if (L > s->nmax[j]) L = s->nmax[j]; }
  return L > s->nF[j] ? L : s->nF[j];
}
static int job_n(const Scn *s, int j, double M, double mu) {
  int a = lower_n(s, j, M), b = pref_n(s, j, mu);
  return a > b ? a : b;
}
static long long cost_at(const Scn *s, double M, double mu) {
  long long t = 0;
  for (int j = 0; j < s->J; ++j) t += (long long)s->g[j] * job_n(s, j, M, mu);
  return t;
}
typedef struct { double hi, lo; long long cost_hi; } Price;
static Price solve_price(const Scn *s, double M, double mu_max, int prec, int *evals) {
  Price p; long long GT = (long long)s->G * s->T;
  long long c0 = cost_at(s, M, 0.0); (*evals)++;
  if (c0 <= GT) { p.hi = p.lo = 0.0; p.cost_hi = c0; return p; }
  uint64_t lob = 0, hib; memcpy(&hib, &mu_max, 8);
  double lo = 0.0, hi = mu_max;
  long long chi = cost_at(s, M, hi); (*evals)++;
  const uint64_t width = 1ull << (52 - prec);
  while (hib - lob > width) {
    uint64_t midb = lob + ((hib - lob) >> 1); double mid; memcpy(&mid, &midb, 8);
    long long cm = cost_at(s, M, mid); (*evals)++;
    if (cm <= GT) { hib = midb; hi = mid; chi = cm; if (cm == GT) break; } else { lob = midb; lo = mid; }
  }
  p.hi = hi; p.lo = lo; p.cost_hi = chi; return p;
}
typedef struct { double V, Meff, mu; } Phi;
static Phi phi_at(const Scn *s, double M, double mu_max, int *evals) {
  Price p = solve_price(s, M, mu_max, 12, evals);
  double w = 0.0, me = 0.0; long long GT = (long long)s->G * s->T;
  for (int j = 0; j < s->J; ++j) { int n = job_n(s, j, M, p.hi); w += util_of(s, j, n); me = fmax(me, rem_of(s, j, n)); }
  Phi r; r.Meff = me; r.mu = p.hi; r.V = w + p.hi * (double)(GT - p.cost_hi) - s->k * me; return r;
}

/* one scenario: writes n[J]; returns status (0 ok, 1 fallback); *objective = welfare - k makespan of the counts */
static int solve_one(Scn *s, int32_t *n_out, double *objective, int *evals_out) {
  const int J = s->J, T = s->T, G = s->G;
  const double D = s->D, share = fmin(1.0, (double)G / J), invJT = 1.0 / ((double)J * T);
  const double next_t = D * (double)(s->round_ptr + T);
  for (int b = 0; b + 1 < s->B; ++b) s->slope[b] = (s->logv[b + 1] - s->logv[b]) / (s->bases[b + 1] - s->bases[b]);
  long long infeasible = 0, forced = 0;
  for (int j = 0; j < J; ++j) {
    double Ef = s->E[j], cf = s->c[j], dbar = s->dbar[j], R = s->rem[j];
    double cap = dbar * (Ef - cf);
    int nfin = cap <= 0.0 ? 0 : (int)fmin(ceil(cap / D - 1e-9), 255.0);
    int nmax = nfin < T ? nfin : T;
    if (s->g[j] > G) nmax = 0;
    s->a[j] = D / (dbar * Ef); s->u0[j] = cf / Ef; s->cap[j] = cap; s->R[j] = R; s->nmax[j] = nmax;
    double capF = share * (s->rhomax * s->ftobj[j] - next_t);
    int nF = 0, bad = 0;
    if (capF < 0.0) bad = 1;
    else {
      double need = R - capF;
      if (need > 0.0) {
        if (need > cap * (1.0 + 1e-12) + 1e-9) bad = 1;
        else { double q = ceil(need / D - 1e-9); if (q > nmax) { if (q > T || s->g[j] > G) bad = 1; else nF = nmax; } else nF = (int)q; }
      }
    }
    s->nF[j] = nF; infeasible += bad; forced += (long long)s->g[j] * nF;
  }
  const long long GT = (long long)G * T;
  const int ftf_ok = (infeasible == 0) && (forced <= GT);
  double mu_max = 0.0, mfloor = 0.0, mtop = 0.0;
  for (int j = 0; j < J; ++j) {
    double w = 1.0;
    if (!ftf_ok) {
      s->nF[j] = 0;
      double ratio = (D * s->round_ptr + s->rem[j] / share) / s->ftobj[j];
      if (ratio > s->rhomax) { w = pow(ratio, (s->rem[j] < D) ? 1e2 : s->lam); if (w > 1e300) w = 1e300; }
    }
    s->ws[j] = w * invJT;
    if (s->nmax[j] > 0) mu_max = fmax(mu_max, (util_of(s, j, 1) - util_of(s, j, 0)) / s->g[j]);
    mfloor = fmax(mfloor, s->R[j] - fmin(D * s->nmax[j], s->cap[j]));
    mtop = fmax(mtop, s->R[j]);
  }
  mu_max = mu_max * (1.0 + 1e-6) + 1e-300; mfloor = fmax(0.0, mfloor);
  int evals = 0;
  const double INF_M = 1e300, eps = 1e-6 * D;
  double mmin = mfloor;
  if (cost_at(s, mfloor, mu_max) > GT) {
    double lo = mfloor, hi = mtop;
    for (int it = 0; it < 60 && hi - lo > 1e-9 * (1.0 + hi); ++it) {
      double mid = 0.5 * (lo + hi);
      if (cost_at(s, mid, mu_max) <= GT) hi = mid; else lo = mid;
      evals++;
    }
    mmin = hi;
  }
  Phi nat = phi_at(s, INF_M, mu_max, &evals);
  double best_V = nat.V, best_thr = INF_M;
  if (nat.Meff - eps >= mmin) {
    Phi t2 = phi_at(s, nat.Meff - eps, mu_max, &evals);
    if (t2.V > nat.V) {
      best_V = t2.V; best_thr = nat.Meff - eps;
      double lo = mmin, hi = nat.Meff - eps;
      Phi pl = phi_at(s, lo, mu_max, &evals);
      if (pl.V > best_V) { best_V = pl.V; best_thr = lo; }
      for (int it = 0; it < 48 && hi - lo > eps; ++it) {
        double mid = 0.5 * (lo + hi);
        Phi p1 = phi_at(s, mid, mu_max, &evals);
        if (p1.V > best_V) { best_V = p1.V; best_thr = mid; }
        double below = p1.Meff - eps;
        if (below < mmin) break;
        Phi p2 = phi_at(s, below, mu_max, &evals);
        if (p2.V > best_V) { best_V = p2.V; best_thr = below; }
        if (p2.V > p1.V) hi = below; else lo = mid;
      }
    }
  }
  Price pr = solve_price(s, best_thr, mu_max, 22, &evals);
  long long used = 0;
  for (int j = 0; j < J; ++j) { n_out[j] = job_n(s, j, best_thr, pr.hi); used += (long long)s->g[j] * n_out[j]; }
  long long left = GT - used;
  if (left > 0 && pr.lo < pr.hi)                  /* ties at the clearing price, in job order */
    for (int j = 0; j < J && left > 0; ++j) {
      int extra = job_n(s, j, best_thr, pr.lo) - n_out[j];
      if (extra > 0) { long long take = left / s->g[j]; if (take > extra) take = extra; n_out[j] += (int)take; left -= take * s->g[j]; }
    }
  for (int rep = 0; rep < 32 && left > 0; ++rep) {   /* completion: best remaining item that still fits */
    double bd = 0.0; int bj = -1;
    for (int j = 0; j < J; ++j)
      if (n_out[j] < s->nmax[j] && s->g[j] <= left) {
        double d = (util_of(s, j, n_out[j] + 1) - util_of(s, j, n_out[j])) / s->g[j];
        if (d > bd) { bd = d; bj = j; }
      }
    if (bj < 0) break;
    n_out[bj]++; left -= s->g[bj];
  }
  double w = 0.0, me = 0.0;
  for (int j = 0; j < J; ++j) { w += util_of(s, j, n_out[j]); me = fmax(me, rem_of(s, j, n_out[j])); }
  *objective = w - s->k * me;
  if (evals_out) *evals_out = evals;
  return ftf_ok ? 0 : 1;
}

/* S scenarios, per-scenario arrays [S][J]; k [S], round_ptr [S]; one scenario per OpenMP thread.
 * Outputs n [S][J], objective [S], status [S], evals [S].  Returns 0. */
int sw_price_search(int S, int J, int T, int G, int B, double D, const double *k, double lam, double rhomax,
                    const int32_t *round_ptr, const double *bases, const double *logv, const int32_t *g,
                    const int32_t *E, const int32_t *c, const double *dbar, const double *rem, const double *ftobj,
                    int32_t *n, double *objective, int32_t *status, int32_t *evals, int nthreads) {
#pragma omp parallel for schedule(dynamic, 1) num_threads(nthreads)
  for (int si = 0; si < S; ++si) {
    Scn s; memset(&s, 0, sizeof(s));
    const size_t o = (size_t)si * J;
    s.J = J; s.T = T; s.G = G; s.B = B; s.round_ptr = round_ptr[si]; s.D = D; s.k = k[si]; s.lam = lam; s.rhomax = rhomax;
    s.bases = bases; s.logv = logv; s.g = g + o; s.E = E + o; s.c = c + o; s.dbar = dbar + o; s.rem = rem + o; s.ftobj = ftobj + o;
    double *buf = (double *)malloc(sizeof(double) * 5 * (size_t)J);
    int *ib = (int *)malloc(sizeof(int) * 2 * (size_t)J);
    s.a = buf; s.u0 = buf + J; s.cap = buf + 2 * (size_t)J; s.ws = buf + 3 * (size_t)J; s.R = buf + 4 * (size_t)J;
    s.nmax = ib; s.nF = ib + J;
    int ev = 0;
    status[si] = solve_one(&s, n + o, objective + si, &ev);
    evals[si] = ev;
    free(buf); free(ib);
  }
  return 0;
}
