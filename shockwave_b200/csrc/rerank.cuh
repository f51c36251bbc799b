// rerank.cuh — local search of the fallback re-rank (rank_in_schedule_jobs, scheduler/shockwave.py:714-793):
// minimise sum_j prio_j * mean round index of job j with the per-job round counts fixed, per-round capacity G.
//
// With w_j = prio_j / n_j the objective is sum_j w_j * sum_{t in rounds(j)} t — a min-cost transportation problem
// (jobs x rounds, a job at most once per round) whose item sizes are the gang widths.  The priority round-sweep of
// place.cu builds a feasible schedule; this pass improves it by NEGATIVE-CYCLE CANCELLING on the round graph:
//   nodes     the T rounds + one slack node
//   edge a->c (a, c rounds) of width class g: the cheapest way to move g GPUs worth of jobs from round a to round c
//             = the job of width g present in a and absent from c with the smallest w_j (c - a), OR (nested widths, the
//             reference's {1,2,4,8} gangs) a composite of two disjoint items of width g/2 — so one 4-gang can trade places
//             with two 2-gangs or a 2-gang and two singles
//   edge slack->a  cost 0 (a round may simply lose a job);  edge c->slack cost 0 if round c has >= g idle GPUs
// A negative cycle of one width class leaves every round's load unchanged (or moves load into idle GPUs through the
// slack node) and lowers the objective by its (negative) cost; for unit widths cancelling until none is left is the
// exact min-cost-flow optimum, for mixed widths it is a local optimum of a neighbourhood that contains every
// same-class exchange chain and the nested cross-class ones.  Bellman-Ford on <= 129 nodes, one warp per width class.
// Measured on the 128 recorded fallback solves of the canonical trace (exact re-rank MILP of the same counts as the
// yardstick): excess median 0, p90 1e-3, max 7e-3 (sweep alone: median 2e-3, p90 1.4e-2, max 4.4e-2).
#pragma once
#include "swb_common.cuh"
#include "swb_internal.h"

namespace swb {

#define RR_MAXCLS 4
#define RR_ITEMJOBS 8

struct RrItem {               // 32 bytes
  double cost;
  unsigned short n;
  unsigned short job[RR_ITEMJOBS];
  unsigned short pad[3];
};

struct RrTop2 { RrItem it[2]; };   // best and second best item of one (pair, class), disjoint job sets

__device__ __forceinline__ void rr_insert(RrItem &b0, RrItem &b1, const RrItem &x) {
  if (x.cost < b0.cost) { b1 = b0; b0 = x; }
  else if (x.cost < b1.cost) b1 = x;
}

// Builds, for every ordered pair of rounds and every width class, the best two items.  One thread per pair.
__device__ void rr_build_items(RrTop2 *items, const unsigned long long *xm, const unsigned char *gs,
                               const unsigned char *remn, const unsigned char *nplan, const double *prio, int J, int T,
                               const int *clsw, int ncls) {
  const double INF = 1e300;
  for (int pr = threadIdx.x; pr < T * T; pr += blockDim.x) {
    const int a = pr / T, c = pr - a * T;
    RrItem best[RR_MAXCLS][2];
#pragma unroll
    for (int k = 0; k < RR_MAXCLS; ++k) { best[k][0].cost = INF; best[k][0].n = 0; best[k][1].cost = INF; best[k][1].n = 0; }
    if (a != c) {
      const int wa = a >> 6, wc = c >> 6;
      const unsigned long long ba = 1ull << (a & 63), bc = 1ull << (c & 63);
      const double dist = (double)(c - a);
      for (int j = 0; j < J; ++j) {
        if (!(xm[2 * j + wa] & ba) || (xm[2 * j + wc] & bc)) continue;
        const int n = nplan[j];
        if (n == 0 || remn[j] != 0) continue;
        const int g = gs[j];
        int k = -1;
#pragma unroll
        for (int q = 0; q < RR_MAXCLS; ++q) if (q < ncls && clsw[q] == g) k = q;
        if (k < 0) continue;
        RrItem x;
        x.cost = prio[j] / (double)n * dist; x.n = 1; x.job[0] = (unsigned short)j;
#pragma unroll
        for (int q = 0; q < RR_MAXCLS; ++q) if (q == k) rr_insert(best[q][0], best[q][1], x);
      }
      // composites: two disjoint items of half the width (nested gang widths)
#pragma unroll
      for (int k = 1; k < RR_MAXCLS; ++k) {
        if (k < ncls && clsw[k] == 2 * clsw[k - 1] && best[k - 1][1].cost < INF &&
            best[k - 1][0].n + best[k - 1][1].n <= RR_ITEMJOBS) {
          RrItem x;
          x.cost = best[k - 1][0].cost + best[k - 1][1].cost;
          x.n = (unsigned short)(best[k - 1][0].n + best[k - 1][1].n);
          for (int q = 0; q < best[k - 1][0].n; ++q) x.job[q] = best[k - 1][0].job[q];
          for (int q = 0; q < best[k - 1][1].n; ++q) x.job[best[k - 1][0].n + q] = best[k - 1][1].job[q];
          rr_insert(best[k][0], best[k][1], x);
        }
      }
    }
#pragma unroll
    for (int k = 0; k < RR_MAXCLS; ++k)
      if (k < ncls) { items[(size_t)k * T * T + pr].it[0] = best[k][0]; items[(size_t)k * T * T + pr].it[1] = best[k][1]; }
  }
}

// One warp: Bellman-Ford (Jacobi sweeps, double-buffered distances) from a virtual source on the class's round graph;
// returns the length of a negative cycle written to cyc[] (nodes in forward order), 0 if there is none.
// d: 2 x (T+1) doubles, pred: T+1 shorts of scratch.
__device__ int rr_find_cycle(const RrTop2 *items_k, const int *idle, int width, int T, double *d, short *pred,
                             short *cyc, double tol) {
  const int lane = threadIdx.x & 31, N = T + 1;
  const double INF = 1e300;
  double *dold = d, *dnew = d + (SWB_MAX_T + 1);
  for (int v = lane; v < N; v += 32) { dold[v] = 0.0; dnew[v] = 0.0; pred[v] = -1; }
  __syncwarp();
  int last = -1;
  for (int it = 0; it < N; ++it) {
    int changed = -1;
    for (int v = lane; v < N; v += 32) {
      double dv = dold[v];
      int pv = -2;
      for (int u = 0; u < N; ++u) {
        if (u == v) continue;
        double cuv;
        if (u == T) cuv = 0.0;                                       // slack -> round v
        else if (v == T) cuv = (idle[u] >= width) ? 0.0 : INF;       // round u -> slack (idle GPUs absorb the item)
        else cuv = items_k[u * T + v].it[0].cost;
        if (cuv >= INF * 0.5) continue;
        const double cand = dold[u] + cuv;
        if (cand < dv - tol) { dv = cand; pv = u; }
      }
      dnew[v] = dv;
      if (pv != -2) { pred[v] = (short)pv; changed = v; }
    }
    changed = __reduce_max_sync(SWB_FULL, changed);
    __syncwarp();
    double *t = dold; dold = dnew; dnew = t;
    last = changed;
    if (changed < 0) return 0;
  }
  // a relaxation in the N-th sweep: `last` hangs off a negative cycle of the predecessor graph; N steps back land on it
  int len = 0;
  if (lane == 0) {
    int x = last;
    for (int q = 0; q < N && x >= 0; ++q) x = pred[x];
    if (x >= 0) {
      int cur = x, guard = 0;
      do { cyc[len++] = (short)cur; cur = pred[cur]; } while (cur != x && cur >= 0 && ++guard <= N);
      if (cur != x) len = 0;
      for (int q = 0; q < len / 2; ++q) { const short t2 = cyc[q]; cyc[q] = cyc[len - 1 - q]; cyc[len - 1 - q] = t2; }
    }
  }
  len = __shfl_sync(SWB_FULL, len, 0);
  __syncwarp();
  return len;
}

// The local search.  All threads of the CTA call it.  xm: [J][2] round masks (in/out); idle: [T] (in/out).
// Returns the number of cancelled cycles.
__device__ __noinline__ int rr_local_search(RrTop2 *items, unsigned long long *xm, const unsigned char *gs,
                                            const unsigned char *remn, const unsigned char *nplan, const double *prio,
                                            int *idle, int J, int T, int max_iters) {
  __shared__ double ws_d[RR_MAXCLS * 2 * (SWB_MAX_T + 1)];
  __shared__ short ws_pred[RR_MAXCLS * (SWB_MAX_T + 1)];
  __shared__ short ws_cyc[RR_MAXCLS * (SWB_MAX_T + 2)];
  // width classes: the (up to 4) distinct widths among the placed jobs, ascending
  __shared__ int s_cls[RR_MAXCLS], s_ncls, s_found, s_len[RR_MAXCLS];
  __shared__ double s_scale;
  if (threadIdx.x == 0) {
    int n = 0, w[RR_MAXCLS];
    for (int j = 0; j < J && n <= RR_MAXCLS; ++j) {
      if (nplan[j] == 0) continue;
      const int g = gs[j];
      bool seen = false;
      for (int q = 0; q < n; ++q) seen |= (w[q] == g);
      if (!seen) { if (n == RR_MAXCLS) { n = RR_MAXCLS + 1; break; } w[n++] = g; }
    }
    if (n > RR_MAXCLS) n = 0;                     // more than 4 distinct widths: leave the sweep's schedule as it is
    for (int q = 1; q < n; ++q) for (int r = q; r > 0 && w[r - 1] > w[r]; --r) { int t = w[r]; w[r] = w[r - 1]; w[r - 1] = t; }
    for (int q = 0; q < n; ++q) s_cls[q] = w[q];
    s_ncls = n;
    double mx = 0.0;
    for (int j = 0; j < J; ++j) if (nplan[j] > 0 && isfinite(prio[j])) mx = fmax(mx, prio[j] / (double)nplan[j]);
    s_scale = mx * (double)T;
  }
  __syncthreads();
  const int ncls = s_ncls;
  if (ncls == 0) return 0;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  int done = 0;
  for (int iter = 0; iter < max_iters; ++iter) {
    rr_build_items(items, xm, gs, remn, nplan, prio, J, T, s_cls, ncls);
    if (threadIdx.x == 0) s_found = -1;
    __syncthreads();
    if (warp < ncls) {
      const int len = rr_find_cycle(items + (size_t)warp * T * T, idle, s_cls[warp], T,
                                    ws_d + warp * 2 * (SWB_MAX_T + 1), ws_pred + warp * (SWB_MAX_T + 1),
                                    ws_cyc + warp * (SWB_MAX_T + 2), 1e-13 * s_scale);
      if (lane == 0) s_len[warp] = len;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      // the lowest class that has a cycle; verify it is really negative (tolerances) and apply it
      for (int k = 0; k < ncls && s_found < 0; ++k) {
        const int len = s_len[k];
        if (len < 2) continue;
        const short *cyc = ws_cyc + k * (SWB_MAX_T + 2);
        const RrTop2 *itk = items + (size_t)k * T * T;
        double tot = 0.0;
        bool ok = true;
        for (int q = 0; q < len && ok; ++q) {
          const int u = cyc[q], v = cyc[(q + 1) % len];
          if (u == T) continue;
          if (v == T) { ok = idle[u] >= s_cls[k]; continue; }
          const RrItem &it = itk[u * T + v].it[0];
          if (it.cost >= 1e299) ok = false; else tot += it.cost;
        }
        if (!ok || !(tot < -1e-12 * s_scale)) continue;
        for (int q = 0; q < len; ++q) {
          const int u = cyc[q], v = cyc[(q + 1) % len];
          if (u == T) { idle[v] += s_cls[k]; continue; }           // v loses an item that no cycle edge brings back
          if (v == T) { idle[u] -= s_cls[k]; continue; }           // u keeps an item: its idle GPUs take it
          const RrItem &it = itk[u * T + v].it[0];
          for (int e = 0; e < it.n; ++e) {
            const int j = it.job[e];
            xm[2 * j + (u >> 6)] &= ~(1ull << (u & 63));
            xm[2 * j + (v >> 6)] |= 1ull << (v & 63);
          }
        }
        s_found = k;
      }
    }
    __syncthreads();
    if (s_found < 0) break;
    ++done;
  }
  return done;
}

}  // namespace swb
