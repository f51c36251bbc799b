// Dense-inverse revised simplex for the SMALL, SPARSE-COLUMN linear programs of the packing policies
// (scheduler/policies/policy.py:68-193 `PolicyWithPacking`: one column per (job combination, worker type),
// each touching one capacity row, <= 2 per-job share rows and <= 2 per-job throughput rows).
//
//     maximise c'x   subject to   A x <= b,  x >= 0          (b may be negative: phase I with ONE artificial)
//
// One CTA per program; the m x m basis inverse lives in global memory (L2 resident: m <= 2048 -> 32 MiB), every
// O(m^2) step (dual vector, rank-one update, Gauss-Jordan refactorisation) is spread over the CTA's threads
// with coalesced row-major accesses, every O(n) step (pricing) is one thread per column.  fp64 throughout: the
// cvxpy/ECOS programs this replaces are float64 and the parity bar is 1e-6 on the objective.
//
// The SAME source compiles for the host (one "thread", barriers are no-ops): tests/lp_host.cpp builds it with g++
// so the pivoting logic is exercised against HiGHS without a GPU.  That host build is test tooling only — the
// library never calls it (libswb200 has no CPU path).
#pragma once
#include <math.h>
#include <stdint.h>

#ifdef __CUDACC__
#define LP_HD __host__ __device__ __forceinline__
#else
#define LP_HD inline
#endif

#ifdef __CUDA_ARCH__
#define LP_TID ((int)threadIdx.x)
#define LP_NT ((int)blockDim.x)
#define LP_SYNC() __syncthreads()
#else
#define LP_TID 0
#define LP_NT 1
#define LP_SYNC() ((void)0)
#endif

namespace swb {
namespace lp {

enum { ST_OPTIMAL = 0, ST_INFEASIBLE = 1, ST_UNBOUNDED = 2, ST_ITER_LIMIT = 3, ST_SINGULAR = 4 };

struct Problem {
  int m, n;            // rows, structural columns
  const int *colp;     // [n + 1]  CSC column pointers
  const int *rowi;     // [nnz]
  const double *val;   // [nnz]
  const double *c;     // [n]
  const double *b;     // [m]
};

struct Work {
  double *Binv;        // [m * m] row-major: row k belongs to basis position k
  double *Bm;          // [m * m] scratch of the refactorisation
  double *xB, *y, *alpha, *cB, *prow;   // [m] each
  int *basis;          // [m]      column at basis position k (structural j < n, slack n + i, artificial n + m)
  int *where;          // [n + m + 1] basis position of a column, -1 = nonbasic
  double *x;           // [n] out
  double *out;         // [8] out: objective, status, iterations, phase-I iterations, refactorisations, bland iterations
  double *sv;          // [64] reduction scratch (device: shared memory)
  int *si;             // [64]
};

// ---- CTA-wide reductions; every thread receives the result ----
// "best" = largest v, ties to the smaller index; i < 0 means "no candidate"
LP_HD bool better(double ov, int oi, double v, int i) { return oi >= 0 && (i < 0 || ov > v || (ov == v && oi < i)); }

LP_HD void team_best(double &v, int &i, double *sv, int *si) {
#ifdef __CUDA_ARCH__
  for (int o = 16; o > 0; o >>= 1) {
    double ov = __shfl_xor_sync(0xffffffffu, v, o);
    int oi = __shfl_xor_sync(0xffffffffu, i, o);
    if (better(ov, oi, v, i)) { v = ov; i = oi; }
  }
  const int w = LP_TID >> 5, l = LP_TID & 31, nw = (LP_NT + 31) >> 5;
  __syncthreads();                      // the scratch row may still be read by the previous reduction
  if (l == 0) { sv[w] = v; si[w] = i; }
  __syncthreads();
  v = (l < nw) ? sv[l] : 0.0;
  i = (l < nw) ? si[l] : -1;
  for (int o = 16; o > 0; o >>= 1) {
    double ov = __shfl_xor_sync(0xffffffffu, v, o);
    int oi = __shfl_xor_sync(0xffffffffu, i, o);
    if (better(ov, oi, v, i)) { v = ov; i = oi; }
  }
#else
  (void)v; (void)i; (void)sv; (void)si;
#endif
}

LP_HD double team_sum(double v, double *sv) {
#ifdef __CUDA_ARCH__
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  const int w = LP_TID >> 5, l = LP_TID & 31, nw = (LP_NT + 31) >> 5;
  __syncthreads();
  if (l == 0) sv[32 + w] = v;
  __syncthreads();
  v = (l < nw) ? sv[32 + l] : 0.0;
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
#else
  (void)sv;
#endif
  return v;
}

// alpha = Binv * (column q)   — q structural, slack or the artificial (a0_i = -1 where b_i < 0)
LP_HD void ftran(const Problem &P, const Work &W, int q) {
  const int m = P.m;
  for (int i = LP_TID; i < m; i += LP_NT) {
    const double *row = W.Binv + (size_t)i * m;
    double a = 0.0;
    if (q < P.n) {
      for (int k = P.colp[q]; k < P.colp[q + 1]; ++k) a = fma(row[P.rowi[k]], P.val[k], a);
    } else if (q < P.n + m) {
      a = row[q - P.n];
    } else {
      for (int r = 0; r < m; ++r)
        if (P.b[r] < 0.0) a -= row[r];
    }
    W.alpha[i] = a;
  }
  LP_SYNC();
}

// pivot on (basis position r, entering column q): rank-one update of Binv and xB
LP_HD void pivot(const Problem &P, const Work &W, int r, int q) {
  const int m = P.m;
  const double ar = W.alpha[r];
  const double inv = 1.0 / ar;
  for (int k = LP_TID; k < m; k += LP_NT) W.prow[k] = W.Binv[(size_t)r * m + k] * inv;
  LP_SYNC();
  const double theta = W.xB[r] * inv;
  const size_t mm = (size_t)m * m;
  for (size_t e = LP_TID; e < mm; e += LP_NT) {
    const int i = (int)(e / m), k = (int)(e - (size_t)i * m);
    if (i == r) {
      W.Binv[e] = W.prow[k];
    } else {
      const double ai = W.alpha[i];
      if (ai != 0.0) W.Binv[e] = fma(-ai, W.prow[k], W.Binv[e]);
    }
  }
  for (int i = LP_TID; i < m; i += LP_NT) {
    if (i == r) continue;
    double v = fma(-theta, W.alpha[i], W.xB[i]);
    W.xB[i] = v < 0.0 ? 0.0 : v;          // the ratio test keeps it >= -roundoff
  }
  LP_SYNC();
  if (LP_TID == 0) {
    W.xB[r] = theta < 0.0 ? 0.0 : theta;
    W.where[W.basis[r]] = -1;
    W.basis[r] = q;
    W.where[q] = r;
  }
  LP_SYNC();
}

// Binv <- inverse of the current basis matrix (Gauss-Jordan with partial pivoting on [B | I]), xB <- Binv b
// (artificial column included when basic).  Returns false when the basis is numerically singular.
LP_HD bool refactor(const Problem &P, const Work &W) {
  const int m = P.m;
  const size_t mm = (size_t)m * m;
  for (size_t e = LP_TID; e < mm; e += LP_NT) {
    W.Bm[e] = 0.0;
    const int i = (int)(e / m), k = (int)(e - (size_t)i * m);
    W.Binv[e] = (i == k) ? 1.0 : 0.0;
  }
  LP_SYNC();
  for (int k = LP_TID; k < m; k += LP_NT) {          // column k of B = column basis[k] of [A | I | a0]
    const int q = W.basis[k];
    if (q < P.n) {
      for (int t = P.colp[q]; t < P.colp[q + 1]; ++t) W.Bm[(size_t)P.rowi[t] * m + k] = P.val[t];
    } else if (q < P.n + m) {
      W.Bm[(size_t)(q - P.n) * m + k] = 1.0;
    } else {
      for (int r = 0; r < m; ++r)
        if (P.b[r] < 0.0) W.Bm[(size_t)r * m + k] = -1.0;
    }
  }
  LP_SYNC();
  for (int p = 0; p < m; ++p) {
    double bv = -1.0;
    int bi = -1;
    for (int i = p + LP_TID; i < m; i += LP_NT) {
      const double a = fabs(W.Bm[(size_t)i * m + p]);
      if (better(a, i, bv, bi)) { bv = a; bi = i; }
    }
    team_best(bv, bi, W.sv, W.si);
    if (bi < 0 || !(bv > 1e-13)) return false;
    if (bi != p) {
      for (int k = LP_TID; k < 2 * m; k += LP_NT) {
        double *M = (k < m) ? W.Bm : W.Binv;
        const int kk = (k < m) ? k : k - m;
        const double t = M[(size_t)p * m + kk];
        M[(size_t)p * m + kk] = M[(size_t)bi * m + kk];
        M[(size_t)bi * m + kk] = t;
      }
      LP_SYNC();
    }
    const double inv = 1.0 / W.Bm[(size_t)p * m + p];
    LP_SYNC();                                   // everyone has read the pivot before it is scaled
    for (int k = LP_TID; k < 2 * m; k += LP_NT) {
      double *M = (k < m) ? W.Bm : W.Binv;
      const int kk = (k < m) ? k : k - m;
      M[(size_t)p * m + kk] *= inv;
    }
    for (int i = LP_TID; i < m; i += LP_NT) W.alpha[i] = (i == p) ? 0.0 : W.Bm[(size_t)i * m + p];   // multipliers
    LP_SYNC();
    // columns < p of Bm are already unit columns and column p becomes one: only columns > p of Bm change,
    // but every column of the inverse does
    for (size_t e = LP_TID; e < 2 * mm; e += LP_NT) {
      const bool inB = e < mm;
      const size_t f = inB ? e : e - mm;
      const int i = (int)(f / m), k = (int)(f - (size_t)i * m);
      if (inB && k <= p) continue;
      const double ai = W.alpha[i];
      if (ai == 0.0) continue;
      double *M = inB ? W.Bm : W.Binv;
      M[f] = fma(-ai, M[(size_t)p * m + k], M[f]);
    }
    LP_SYNC();
    for (int i = LP_TID; i < m; i += LP_NT)
      if (i != p) W.Bm[(size_t)i * m + p] = 0.0;
    LP_SYNC();
  }
  for (int i = LP_TID; i < m; i += LP_NT) {
    const double *row = W.Binv + (size_t)i * m;
    double s = 0.0;
    for (int k = 0; k < m; ++k) s = fma(row[k], P.b[k], s);
    W.xB[i] = s < 0.0 ? 0.0 : s;
  }
  LP_SYNC();
  return true;
}

// One CTA solves one program.  max_iter bounds the pivots of both phases together.
LP_HD void simplex(const Problem &P, const Work &W, int max_iter) {
  const int m = P.m, n = P.n, ART = n + m;
  const double DTOL = 1e-9, PTOL = 1e-9, TIE = 1e-12;
  const int REFACTOR_EVERY = 96;

  // scale of the costs (reduced-cost tolerance is relative to it)
  double cmax = 0.0;
  {
    double v = -1.0; int i = -1;
    for (int j = LP_TID; j < n; j += LP_NT) { const double a = fabs(P.c[j]); if (better(a, j, v, i)) { v = a; i = j; } }
    team_best(v, i, W.sv, W.si);
    cmax = (i >= 0 && v > 0.0) ? v : 1.0;
  }
  // slack basis
  for (size_t e = LP_TID; e < (size_t)m * m; e += LP_NT) {
    const int i = (int)(e / m), k = (int)(e - (size_t)i * m);
    W.Binv[e] = (i == k) ? 1.0 : 0.0;
  }
  for (int j = LP_TID; j <= ART; j += LP_NT) W.where[j] = -1;
  LP_SYNC();
  for (int i = LP_TID; i < m; i += LP_NT) { W.basis[i] = n + i; W.where[n + i] = i; W.xB[i] = P.b[i]; }
  LP_SYNC();

  int phase = 2;
  {
    double v = 0.0; int r = -1;                   // most negative right-hand side
    for (int i = LP_TID; i < m; i += LP_NT) { const double a = -P.b[i]; if (a > 0.0 && better(a, i, v, r)) { v = a; r = i; } }
    team_best(v, r, W.sv, W.si);
    if (r >= 0) {
      phase = 1;
      ftran(P, W, ART);                           // Binv = I: alpha = a0
      // xB may be negative here: pivot() clamps at 0 only what the ratio test guarantees, so do this one by hand
      const double theta = -P.b[r];               // = xB_r / alpha_r with alpha_r = -1
      for (int i = LP_TID; i < m; i += LP_NT) W.xB[i] = (i == r) ? theta : P.b[i] - theta * W.alpha[i];
      LP_SYNC();
      for (int i = LP_TID; i < m; i += LP_NT)
        if (i != r && W.xB[i] < 0.0) W.xB[i] = 0.0;
      // Binv update for alpha = a0, pivot row r (alpha_r = -1)
      for (int k = LP_TID; k < m; k += LP_NT) W.prow[k] = -W.Binv[(size_t)r * m + k];
      LP_SYNC();
      for (size_t e = LP_TID; e < (size_t)m * m; e += LP_NT) {
        const int i = (int)(e / m), k = (int)(e - (size_t)i * m);
        if (i == r) W.Binv[e] = W.prow[k];
        else if (W.alpha[i] != 0.0) W.Binv[e] = fma(-W.alpha[i], W.prow[k], W.Binv[e]);
      }
      LP_SYNC();
      if (LP_TID == 0) { W.where[n + r] = -1; W.basis[r] = ART; W.where[ART] = r; }
      LP_SYNC();
    }
  }

  int it = 0, it1 = 0, nref = 0, nbland = 0, since_ref = 0, stall = 0, status = ST_ITER_LIMIT;
  bool verified = false;      // optimality re-checked on a fresh factorisation
  while (it < max_iter) {
    // ---- dual vector y = cB' Binv ----
    for (int k = LP_TID; k < m; k += LP_NT) {
      const int q = W.basis[k];
      W.cB[k] = (phase == 1) ? (q == ART ? -1.0 : 0.0) : (q < n ? P.c[q] : 0.0);
    }
    LP_SYNC();
    for (int i = LP_TID; i < m; i += LP_NT) {
      double s = 0.0;
      for (int k = 0; k < m; ++k) {
        const double ck = W.cB[k];
        if (ck != 0.0) s = fma(ck, W.Binv[(size_t)k * m + i], s);
      }
      W.y[i] = s;
    }
    LP_SYNC();
    // ---- pricing ----
    const bool bland = stall > 40;
    const double dt = DTOL * (phase == 1 ? 1.0 : cmax);
    double bv = 0.0; int q = -1;
    for (int j = LP_TID; j < n + m; j += LP_NT) {
      if (W.where[j] >= 0) continue;
      double d;
      if (j < n) {
        d = (phase == 1) ? 0.0 : P.c[j];
        for (int k = P.colp[j]; k < P.colp[j + 1]; ++k) d = fma(-W.y[P.rowi[k]], P.val[k], d);
      } else {
        d = -W.y[j - n];
      }
      if (d > dt) {
        const double score = bland ? -(double)j : d;
        if (better(score, j, bv, q)) { bv = score; q = j; }
      }
    }
    team_best(bv, q, W.sv, W.si);
    if (q < 0) {                                   // no improving column
      if (!verified && since_ref > 0) {            // confirm on a fresh inverse before declaring anything
        if (!refactor(P, W)) { status = ST_SINGULAR; break; }
        ++nref; since_ref = 0; verified = true;
        continue;
      }
      if (phase == 1) {
        const int pa = W.where[ART];
        const double x0 = pa >= 0 ? W.xB[pa] : 0.0;
        if (x0 > 1e-9) { status = ST_INFEASIBLE; break; }
        if (pa >= 0) {
          // artificial basic at level 0: pivot it out on any usable column of its row (degenerate pivot)
          double pv = 0.0; int pj = -1;
          for (int j = LP_TID; j < n + m; j += LP_NT) {
            if (W.where[j] >= 0) continue;
            double a = 0.0;
            const double *row = W.Binv + (size_t)pa * m;
            if (j < n) { for (int k = P.colp[j]; k < P.colp[j + 1]; ++k) a = fma(row[P.rowi[k]], P.val[k], a); }
            else a = row[j - n];
            a = fabs(a);
            if (a > 1e-7 && better(a, j, pv, pj)) { pv = a; pj = j; }
          }
          team_best(pv, pj, W.sv, W.si);
          if (pj >= 0) {
            ftran(P, W, pj);
            if (LP_TID == 0) W.xB[pa] = 0.0;
            LP_SYNC();
            pivot(P, W, pa, pj);
            ++it; ++since_ref;
          }
          // else: the row is redundant — the artificial stays basic at 0 and can never move
        }
        phase = 2; stall = 0; verified = false;
        continue;
      }
      status = ST_OPTIMAL;
      break;
    }
    verified = false;
    const double dq = bland ? 1.0 : bv;            // only its sign matters in Bland mode
    // ---- entering column in basis coordinates ----
    ftran(P, W, q);
    // ---- ratio test ----
    double amax = 0.0;
    {
      double v = -1.0; int i0 = -1;
      for (int i = LP_TID; i < m; i += LP_NT) { const double a = W.alpha[i]; if (a > 0.0 && better(a, i, v, i0)) { v = a; i0 = i; } }
      team_best(v, i0, W.sv, W.si);
      amax = i0 >= 0 ? v : 0.0;
    }
    const double pt = PTOL * (amax > 1.0 ? amax : 1.0);
    double tv = 0.0; int tr = -1;                  // maximise -ratio
    for (int i = LP_TID; i < m; i += LP_NT) {
      const double a = W.alpha[i];
      if (a > pt) { const double ratio = -(W.xB[i] / a); if (better(ratio, i, tv, tr)) { tv = ratio; tr = i; } }
    }
    team_best(tv, tr, W.sv, W.si);
    if (tr < 0) { status = (phase == 1) ? ST_SINGULAR : ST_UNBOUNDED; break; }
    const double theta = -tv;
    // among the rows tied at the minimum ratio: the artificial first, then (Bland) the smallest column index,
    // else the largest pivot element
    double sv2 = 0.0; int r = -1;
    for (int i = LP_TID; i < m; i += LP_NT) {
      const double a = W.alpha[i];
      if (a > pt && W.xB[i] / a <= theta + TIE * (1.0 + theta)) {
        double score = (W.basis[i] == ART) ? 1e300 : (bland ? -(double)W.basis[i] : a);
        if (better(score, i, sv2, r)) { sv2 = score; r = i; }
      }
    }
    team_best(sv2, r, W.sv, W.si);
    const bool art_leaves = (W.basis[r] == ART);
    LP_SYNC();
    pivot(P, W, r, q);
    ++it; ++since_ref;
    if (phase == 1) ++it1;
    if (bland) ++nbland;
    const double gain = dq * theta;
    stall = (gain > 1e-13) ? 0 : stall + 1;
    if (art_leaves) { phase = 2; stall = 0; }
    if (since_ref >= REFACTOR_EVERY) {
      if (!refactor(P, W)) { status = ST_SINGULAR; break; }
      ++nref; since_ref = 0;
    }
  }
  // ---- primal solution ----
  LP_SYNC();
  for (int j = LP_TID; j < n; j += LP_NT) W.x[j] = (W.where[j] >= 0) ? W.xB[W.where[j]] : 0.0;
  LP_SYNC();
  double o = 0.0;
  for (int j = LP_TID; j < n; j += LP_NT) o = fma(P.c[j], W.x[j], o);
  o = team_sum(o, W.sv);
  if (LP_TID == 0) {
    W.out[0] = o; W.out[1] = (double)status; W.out[2] = (double)it; W.out[3] = (double)it1;
    W.out[4] = (double)nref; W.out[5] = (double)nbland; W.out[6] = 0.0; W.out[7] = 0.0;
  }
  LP_SYNC();
}

}  // namespace lp
}  // namespace swb
