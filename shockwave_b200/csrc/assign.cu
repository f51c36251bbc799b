// assign.cu — AlloX's min-cost assignment (scheduler/policies/allox.py:108-144) on the GPU.
//
// The reference builds q[i][k*n + j] = (k+1) * steps_i / throughput(i, type(worker j)) + times_since_start_i
// for m jobs and m*n (worker, queue position) columns and calls scipy.optimize.linear_sum_assignment (a
// Jonker-Volgenant shortest-augmenting-path solver).  Same algorithm here, one CTA: the outer loops (one
// augmentation per job, one Dijkstra step per scanned row) are sequential, every step relaxes ALL columns
// in parallel and finds the next column with a block arg-min.  The cost matrix is never materialised:
// a column decodes to (position k, worker j) and the cost is recomputed from p[i][type] and t[i].
// float64 throughout; exact (optimal) like the reference's solver, ties broken by lowest column index
// with free columns first.
#include <math.h>

#include "swb_common.cuh"
#include "swb_internal.h"

namespace swb {

__device__ __forceinline__ double ax_cost(const AssignLaunch &L, int i, int col) {
  const int k = col / L.n, j = col - k * L.n;
  return (double)(k + 1) * L.p[i * L.W + L.wtype[j]] + L.t[i];
}

__global__ void __launch_bounds__(1024, 1) assign_kernel(AssignLaunch L) {
  __shared__ double s_val[32];
  __shared__ int s_idx[32];
  __shared__ int s_i, s_sink;
  __shared__ double s_min;
  const int m = L.m, N = L.m * L.n;
  double *u = L.u, *v = L.v, *spc = L.spc;
  int *col4row = L.col4row, *row4col = L.row4col, *path = L.path;
  unsigned char *SC = L.inSC, *SR = L.inSR;
  for (int j = threadIdx.x; j < N; j += blockDim.x) { v[j] = 0.0; row4col[j] = -1; }
  for (int i = threadIdx.x; i < m; i += blockDim.x) { u[i] = 0.0; col4row[i] = -1; }
  __syncthreads();
  for (int cur = 0; cur < m; ++cur) {
    for (int j = threadIdx.x; j < N; j += blockDim.x) { spc[j] = INFINITY; SC[j] = 0; path[j] = -1; }
    for (int i = threadIdx.x; i < m; i += blockDim.x) SR[i] = 0;
    if (threadIdx.x == 0) { s_i = cur; s_sink = -1; s_min = 0.0; }
    __syncthreads();
    while (s_sink < 0) {
      const int i = s_i;
      const double minVal = s_min, ui = u[i];
      if (threadIdx.x == 0) SR[i] = 1;
      // relax every column not yet in the tree, track the best (value, free-first, lowest index)
      double bv = INFINITY;
      int bj = -1;
      for (int j = threadIdx.x; j < N; j += blockDim.x) {
        if (SC[j]) continue;
        const double r = minVal + ax_cost(L, i, j) - ui - v[j];
        double cur_spc = spc[j];
        if (r < cur_spc) { spc[j] = r; path[j] = i; cur_spc = r; }
        const bool better = cur_spc < bv || (cur_spc == bv && bj >= 0 && row4col[j] < 0 && row4col[bj] >= 0);
        if (better) { bv = cur_spc; bj = j; }
      }
      // block arg-min with the same tie rule
      for (int o = 16; o > 0; o >>= 1) {
        const double ov = __shfl_xor_sync(SWB_FULL, bv, o);
        const int oj = __shfl_xor_sync(SWB_FULL, bj, o);
        bool take = false;
        if (oj >= 0) {
          if (bj < 0 || ov < bv) take = true;
          else if (ov == bv) {
            const bool of = row4col[oj] < 0, bf = row4col[bj] < 0;
            take = (of && !bf) || (of == bf && oj < bj);
          }
        }
        if (take) { bv = ov; bj = oj; }
      }
      if ((threadIdx.x & 31) == 0) { s_val[threadIdx.x >> 5] = bv; s_idx[threadIdx.x >> 5] = bj; }
      __syncthreads();
      if (threadIdx.x < 32) {
        const int nw = blockDim.x >> 5;
        bv = threadIdx.x < nw ? s_val[threadIdx.x] : INFINITY;
        bj = threadIdx.x < nw ? s_idx[threadIdx.x] : -1;
        for (int o = 16; o > 0; o >>= 1) {
          const double ov = __shfl_xor_sync(SWB_FULL, bv, o);
          const int oj = __shfl_xor_sync(SWB_FULL, bj, o);
          bool take = false;
          if (oj >= 0) {
            if (bj < 0 || ov < bv) take = true;
            else if (ov == bv) {
              const bool of = row4col[oj] < 0, bf = row4col[bj] < 0;
              take = (of && !bf) || (of == bf && oj < bj);
            }
          }
          if (take) { bv = ov; bj = oj; }
        }
        if (threadIdx.x == 0) {
          s_min = bv; SC[bj] = 1;
          if (row4col[bj] < 0) s_sink = bj; else s_i = row4col[bj];
        }
      }
      __syncthreads();
    }
    // dual updates (scipy _lsap: u[cur] += minVal; u[i] += minVal - spc[col4row[i]]; v[j] -= minVal - spc[j])
    const double minVal = s_min;
    for (int i = threadIdx.x; i < m; i += blockDim.x) {
      if (i == cur) u[i] += minVal;
      else if (SR[i]) u[i] += minVal - spc[col4row[i]];
    }
    for (int j = threadIdx.x; j < N; j += blockDim.x)
      if (SC[j]) v[j] -= minVal - spc[j];
    __syncthreads();
    // augment along the alternating path (sequential, at most m links)
    if (threadIdx.x == 0) {
      int j = s_sink;
      while (true) {
        const int i = path[j];
        row4col[j] = i;
        const int prev = col4row[i];
        col4row[i] = j;
        j = prev;
        if (i == cur) break;
      }
    }
    __syncthreads();
  }
  double tot = 0.0;
  for (int i = threadIdx.x; i < m; i += blockDim.x) tot += ax_cost(L, i, col4row[i]);
  __shared__ double red[2 * 64];
  BlockRed br(red);
  tot = br.sum(tot);
  if (threadIdx.x == 0) L.out[0] = tot;
}

cudaError_t launch_assign(const AssignLaunch &L, cudaStream_t st) {
  assign_kernel<<<1, 1024, 0, st>>>(L);
  return cudaGetLastError();
}

}  // namespace swb
