// market.cu — dense projected-gradient / price-response iteration over the allocation tensor
// X[S][J][W][T] (scenario x job x worker type x planning round, fp32, t innermost).
//
// This is the general *volatile Fisher market* form of Shockwave's relaxation: a job may progress at a
// different rate r_jw (epochs per round) on every worker type and every (type, round) slot has its own
// capacity price — the formulation the reference leaves as future work ("we assume homogeneous
// hardware", scripts/drivers/simulate_scheduler_with_trace.py:76-78) and BASELINE.json's north_star
// names.  Objective (per scenario), same pieces as shockwave.py:565-568:
//     sum_j w_j plog((c_j + P_j)/E_j)/(J T) - k max_j max(0, R_j - dbar_j P_j),   P_j = sum_wt r_jw x_jwt
//     s.t. sum_j g_j x_jwt <= G_w  (per worker type and round),  sum_w x_jwt <= 1,  0 <= x <= 1.
// One iteration = market_dual_kernel (O(J + W T) per scenario: marginal utilities theta_j, capacity
// scale factors and price update from the previous pass' reductions) + market_step_kernel (the dense
// pass): x <- clip(x * colscale_wt + eta (theta_j r_jw - pi_wt g_j)/(theta_j r_jw + pi_wt g_j), 0, 1)
// (a proportional-response style, scale-free step), per-(job, round) budget
// normalisation over worker types, and — fused in the same pass — the row reduction P_j (warp
// shuffles) and the column reduction sum_j g_j x_jwt (registers -> shared memory -> one atomicAdd per
// column and CTA).  The dense pass reads X once and writes X once: 8 bytes per element, HBM bound.
// On homogeneous inputs (W = 1, r_jw = D/dbar_j) its fixed point is the relaxation solved exactly by
// solve.cu, which the tests use as the cross-check.
#include <math.h>

#include "swb_common.cuh"
#include "swb_internal.h"

namespace swb {

#define MK_THREADS 256
// row sums go through global atomics when the Q = T/4 threads of a job are not a power-of-two sub-warp
#define Q_ROW_ATOMICS(T) ((((T) >> 2) > 32) || ((((T) >> 2) & (((T) >> 2) - 1)) != 0))

__device__ __forceinline__ float4 ld_stream(const float4 *p) {
  float4 v;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
  return v;
}
__device__ __forceinline__ void st_stream(float4 *p, const float4 &v) {
  asm volatile("st.global.L1::no_allocate.v4.f32 [%0], {%1,%2,%3,%4};"
               :: "l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}

// Dense pass.  Grid (job tiles, S).  A thread owns one float4 of rounds (4 consecutive t) of a job and
// loops over that job's worker types, so the budget normalisation over w is thread-local.  Q = T/4
// threads cover a job; MK_THREADS/Q jobs per sweep; U sweeps are loaded before any is used so that
// U*W 16-byte loads per thread are in flight (the pass is HBM bound: memory-level parallelism first).
template <int W, int U>
__global__ void __launch_bounds__(MK_THREADS) market_step_kernel(MarketLaunch L) {
  extern __shared__ float sm[];           // [W*T] colscale | [W*T] price | [W*T] column accumulators
  const int s = blockIdx.y;
  const int T = L.T, J = L.J, WT = W * T;
  const int Q = T >> 2;                   // float4 groups per (job, type) row
  float *cs = sm, *pi = sm + WT, *acc = sm + 2 * WT;
  const float *cs_g = L.colscale + (size_t)s * WT, *pi_g = L.price + (size_t)s * WT;
  for (int i = threadIdx.x; i < WT; i += blockDim.x) { cs[i] = cs_g[i]; pi[i] = pi_g[i]; acc[i] = 0.f; }
  __syncthreads();
  const float eta = L.eta;
  const int q = threadIdx.x % Q;          // which 4 rounds
  const int jl = threadIdx.x / Q;         // job lane within the sweep
  const int jobs_per_sweep = MK_THREADS / Q;
  const int j0 = blockIdx.x * L.jobs_per_cta;
  const int j1 = min(J, j0 + L.jobs_per_cta);
  const bool active_lane = jl < jobs_per_sweep;
  float colacc[W][4];
  float csr[W][4], pir[W][4];             // this thread's columns never change: keep them in registers
#pragma unroll
  for (int w = 0; w < W; ++w)
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      colacc[w][e] = 0.f;
      csr[w][e] = active_lane ? cs[w * T + 4 * q + e] : 1.f;
      pir[w][e] = active_lane ? pi[w * T + 4 * q + e] : 0.f;
    }
  for (int jb = j0; jb < j1; jb += U * jobs_per_sweep) {
    float4 x[U][W];
    float theta[U], gj[U], r[U][W];
    bool live[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int j = jb + u * jobs_per_sweep + jl;
      live[u] = active_lane && j < j1;
      if (live[u]) {
        const size_t sj = (size_t)s * J + j;
        const float4 *xrow = reinterpret_cast<const float4 *>(L.X + sj * WT);
#pragma unroll
        for (int w = 0; w < W; ++w) x[u][w] = ld_stream(xrow + w * Q + q);
        theta[u] = L.theta[sj];
        gj[u] = (float)L.g[L.per_scn ? sj : j];
#pragma unroll
        for (int w = 0; w < W; ++w) r[u][w] = L.rate[(L.per_scn ? sj : (size_t)j) * W + w];
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int j = jb + u * jobs_per_sweep + jl;
      float rowp = 0.f;
      if (live[u]) {
        const size_t sj = (size_t)s * J + j;
        float4 *xrow = reinterpret_cast<float4 *>(L.X + sj * WT);
        float4 v[W];
        float4 tot = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int w = 0; w < W; ++w) {
          // scale-free response: relative surplus (gain - cost)/(gain + cost) in [-1, 1] — marginal
          // utilities span many orders of magnitude across jobs (PWL slope 61 vs 1.1, fallback priorities)
          const float gain = theta[u] * r[u][w];
          const float k0 = pir[w][0] * gj[u], k1 = pir[w][1] * gj[u], k2 = pir[w][2] * gj[u], k3 = pir[w][3] * gj[u];
          float4 y;
          y.x = fminf(fmaxf(fmaf(x[u][w].x, csr[w][0], eta * __fdividef(gain - k0, gain + k0 + 1e-30f)), 0.f), 1.f);
          y.y = fminf(fmaxf(fmaf(x[u][w].y, csr[w][1], eta * __fdividef(gain - k1, gain + k1 + 1e-30f)), 0.f), 1.f);
          y.z = fminf(fmaxf(fmaf(x[u][w].z, csr[w][2], eta * __fdividef(gain - k2, gain + k2 + 1e-30f)), 0.f), 1.f);
          y.w = fminf(fmaxf(fmaf(x[u][w].w, csr[w][3], eta * __fdividef(gain - k3, gain + k3 + 1e-30f)), 0.f), 1.f);
          v[w] = y;
          tot.x += y.x; tot.y += y.y; tot.z += y.z; tot.w += y.w;
        }
        // per-(job, round) budget normalisation: sum_w x_jwt <= 1   (policy.py:64 generalised per round)
        const float nx = tot.x > 1.f ? 1.f / tot.x : 1.f, ny = tot.y > 1.f ? 1.f / tot.y : 1.f;
        const float nz = tot.z > 1.f ? 1.f / tot.z : 1.f, nw = tot.w > 1.f ? 1.f / tot.w : 1.f;
#pragma unroll
        for (int w = 0; w < W; ++w) {
          float4 y = v[w];
          if (W > 1) { y.x *= nx; y.y *= ny; y.z *= nz; y.w *= nw; }
          st_stream(xrow + w * Q + q, y);
          rowp = fmaf(r[u][w], (y.x + y.y) + (y.z + y.w), rowp);
          colacc[w][0] = fmaf(gj[u], y.x, colacc[w][0]); colacc[w][1] = fmaf(gj[u], y.y, colacc[w][1]);
          colacc[w][2] = fmaf(gj[u], y.z, colacc[w][2]); colacc[w][3] = fmaf(gj[u], y.w, colacc[w][3]);
        }
      }
      // row reduction over the Q threads of the job: shuffles inside the warp when Q is a power of two
      // <= 32, else the partial sums go through global atomics (rowp is zeroed by the dual pass)
      if (!Q_ROW_ATOMICS(T)) {
        for (int o = Q >> 1; o > 0; o >>= 1) rowp += __shfl_xor_sync(SWB_FULL, rowp, o);
        if (live[u] && q == 0) L.rowp[(size_t)s * J + j] = rowp;
      } else if (live[u]) {
        atomicAdd(&L.rowp[(size_t)s * J + j], rowp);
      }
    }
  }
  // column reduction: registers -> shared -> global
  if (active_lane) {
#pragma unroll
    for (int w = 0; w < W; ++w) {
      const int c0 = w * T + 4 * q;
      atomicAdd(&acc[c0 + 0], colacc[w][0]); atomicAdd(&acc[c0 + 1], colacc[w][1]);
      atomicAdd(&acc[c0 + 2], colacc[w][2]); atomicAdd(&acc[c0 + 3], colacc[w][3]);
    }
  }
  __syncthreads();
  float *col_g = L.colload + (size_t)s * WT;
  for (int i = threadIdx.x; i < WT; i += blockDim.x) atomicAdd(&col_g[i], acc[i]);
}


// Fast path of the dense pass: rounds per job T = 4 Q with Q a power of two <= 32 fixed at compile time (T = 32 / 64 /
// 128), so the row reduction is a fixed shuffle tree, every index is a 32-bit offset from a per-scenario base pointer
// and nothing depends on T at run time.  Same arithmetic, same order of operations as market_step_kernel (the numpy
// restatement in tests/ref_market.py covers both); the generic kernel spent ~40 thread instructions per tensor element,
// 71 % issue-active, which held the pass at 0.81 of the HBM peak (ncu r02).
template <int W, int U, int Q>
__global__ void __launch_bounds__(MK_THREADS) market_step_fast(MarketLaunch L) {
  constexpr int T = 4 * Q, WT = W * T, JPS = MK_THREADS / Q;
  __shared__ float acc[WT];
  const int s = blockIdx.y, J = L.J;
  const int q = threadIdx.x & (Q - 1), jl = threadIdx.x / Q;
  const float *cs_g = L.colscale + (size_t)s * WT, *pi_g = L.price + (size_t)s * WT;
  for (int i = threadIdx.x; i < WT; i += MK_THREADS) acc[i] = 0.f;
  const float eta = L.eta;
  const int j0 = blockIdx.x * L.jobs_per_cta;
  const int j1 = min(J, j0 + L.jobs_per_cta);
  float colacc[W][4], csr[W][4], pir[W][4];
#pragma unroll
  for (int w = 0; w < W; ++w) {
    const float4 c4 = *reinterpret_cast<const float4 *>(cs_g + w * T + 4 * q);
    const float4 p4 = *reinterpret_cast<const float4 *>(pi_g + w * T + 4 * q);
    csr[w][0] = c4.x; csr[w][1] = c4.y; csr[w][2] = c4.z; csr[w][3] = c4.w;
    pir[w][0] = p4.x; pir[w][1] = p4.y; pir[w][2] = p4.z; pir[w][3] = p4.w;
#pragma unroll
    for (int e = 0; e < 4; ++e) colacc[w][e] = 0.f;
  }
  __syncthreads();
  const size_t sJ = (size_t)s * J;
  float4 *xs = reinterpret_cast<float4 *>(L.X + sJ * WT) + q;          // + (j W + w) Q : 32-bit offsets below
  const float *th = L.theta + sJ;
  float *rp = L.rowp + sJ;
  const int32_t *gp = L.g + (L.per_scn ? sJ : 0);
  const float *rt = L.rate + (L.per_scn ? sJ * W : 0);
  for (int jb = j0 + jl; jb < j1; jb += U * JPS) {
    float4 x[U][W];
    float theta[U], gj[U], r[U][W];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int j = jb + u * JPS;
      if (j < j1) {
#pragma unroll
        for (int w = 0; w < W; ++w) x[u][w] = ld_stream(xs + (j * W + w) * Q);
        theta[u] = th[j];
        gj[u] = (float)gp[j];
#pragma unroll
        for (int w = 0; w < W; ++w) r[u][w] = rt[j * W + w];
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int j = jb + u * JPS;
      const bool live = j < j1;
      float rowp = 0.f;
      if (live) {
        float4 v[W];
        float4 tot = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int w = 0; w < W; ++w) {
          const float gain = theta[u] * r[u][w];
          const float k0 = pir[w][0] * gj[u], k1 = pir[w][1] * gj[u], k2 = pir[w][2] * gj[u], k3 = pir[w][3] * gj[u];
          float4 y;
          y.x = fminf(fmaxf(fmaf(x[u][w].x, csr[w][0], eta * __fdividef(gain - k0, gain + k0 + 1e-30f)), 0.f), 1.f);
          y.y = fminf(fmaxf(fmaf(x[u][w].y, csr[w][1], eta * __fdividef(gain - k1, gain + k1 + 1e-30f)), 0.f), 1.f);
          y.z = fminf(fmaxf(fmaf(x[u][w].z, csr[w][2], eta * __fdividef(gain - k2, gain + k2 + 1e-30f)), 0.f), 1.f);
          y.w = fminf(fmaxf(fmaf(x[u][w].w, csr[w][3], eta * __fdividef(gain - k3, gain + k3 + 1e-30f)), 0.f), 1.f);
          v[w] = y;
          tot.x += y.x; tot.y += y.y; tot.z += y.z; tot.w += y.w;
        }
        const float nx = tot.x > 1.f ? 1.f / tot.x : 1.f, ny = tot.y > 1.f ? 1.f / tot.y : 1.f;
        const float nz = tot.z > 1.f ? 1.f / tot.z : 1.f, nw = tot.w > 1.f ? 1.f / tot.w : 1.f;
#pragma unroll
        for (int w = 0; w < W; ++w) {
          float4 y = v[w];
          if (W > 1) { y.x *= nx; y.y *= ny; y.z *= nz; y.w *= nw; }
          st_stream(xs + (j * W + w) * Q, y);
          rowp = fmaf(r[u][w], (y.x + y.y) + (y.z + y.w), rowp);
          colacc[w][0] = fmaf(gj[u], y.x, colacc[w][0]); colacc[w][1] = fmaf(gj[u], y.y, colacc[w][1]);
          colacc[w][2] = fmaf(gj[u], y.z, colacc[w][2]); colacc[w][3] = fmaf(gj[u], y.w, colacc[w][3]);
        }
      }
#pragma unroll
      for (int o = Q >> 1; o > 0; o >>= 1) rowp += __shfl_xor_sync(SWB_FULL, rowp, o);
      if (live && q == 0) rp[j] = rowp;
    }
  }
#pragma unroll
  for (int w = 0; w < W; ++w) {
    const int c0 = w * T + 4 * q;
    atomicAdd(&acc[c0 + 0], colacc[w][0]); atomicAdd(&acc[c0 + 1], colacc[w][1]);
    atomicAdd(&acc[c0 + 2], colacc[w][2]); atomicAdd(&acc[c0 + 3], colacc[w][3]);
  }
  __syncthreads();
  float *col_g = L.colload + (size_t)s * WT;
  for (int i = threadIdx.x; i < WT; i += MK_THREADS) atomicAdd(&col_g[i], acc[i]);
}

template <int W, int U>
static bool launch_fast(const MarketLaunch &L, dim3 grid, cudaStream_t st) {
  switch (L.T) {
    case 32: market_step_fast<W, U, 8><<<grid, MK_THREADS, 0, st>>>(L); return true;
    case 64: market_step_fast<W, U, 16><<<grid, MK_THREADS, 0, st>>>(L); return true;
    case 128: market_step_fast<W, U, 32><<<grid, MK_THREADS, 0, st>>>(L); return true;
    default: return false;
  }
}

// Per-scenario small pass: marginal utilities, makespan sub-gradient, capacity scale + price update.
__global__ void __launch_bounds__(1024) market_dual_kernel(MarketLaunch L) {
  __shared__ double red[2 * 64];
  __shared__ Pwl P;
  const int s = blockIdx.x;
  const int J = L.J, WT = L.W * L.T;
  const swb_params &prm = L.prm[s];
  BlockRed br(red);
  if (threadIdx.x == 0) {
    P.B = prm.nbases;
    for (int b = 0; b < prm.nbases; ++b) { P.base[b] = prm.bases[b]; P.logv[b] = prm.logv[b]; }
    for (int b = 0; b + 1 < prm.nbases; ++b)
      P.slope[b] = (prm.logv[b + 1] - prm.logv[b]) / (prm.bases[b + 1] - prm.bases[b]);
  }
  __syncthreads();
  const double invJT = 1.0 / ((double)J * (double)L.T);
  double welfare = 0.0, mx = 0.0;
  for (int j = threadIdx.x; j < J; j += blockDim.x) {
    const size_t sj = (size_t)s * J + j, ji = L.per_scn ? sj : j;
    const double Ef = L.E[ji], cf = L.c[ji];
    const double Pj = fmin((double)L.rowp[sj], Ef - cf);
    const double u = (cf + Pj) / Ef;
    welfare += plog(P, u);
    mx = fmax(mx, fmax(0.0, L.rem[ji] - L.dbar[ji] * Pj));
  }
  welfare = br.sum(welfare) * invJT;
  mx = br.max(mx);
  const double band = 1e-3 * prm.round_duration;
  double cnt = 0.0;
  for (int j = threadIdx.x; j < J; j += blockDim.x) {
    const size_t sj = (size_t)s * J + j, ji = L.per_scn ? sj : j;
    const double Ef = L.E[ji], cf = L.c[ji];
    const double Pj = fmin((double)L.rowp[sj], Ef - cf);
    const double remj = fmax(0.0, L.rem[ji] - L.dbar[ji] * Pj);
    cnt += (mx > 0.0 && remj >= mx - band && Pj < Ef - cf) ? 1.0 : 0.0;
  }
  cnt = br.sum(cnt);
  for (int j = threadIdx.x; j < J; j += blockDim.x) {
    const size_t sj = (size_t)s * J + j, ji = L.per_scn ? sj : j;
    const double Ef = L.E[ji], cf = L.c[ji];
    const double Praw = (double)L.rowp[sj];
    const double Pj = fmin(Praw, Ef - cf);
    const double u = (cf + Pj) / Ef;
    double th = 0.0;
    if (Praw < Ef - cf) {
      int b = 0;
      for (int i = 1; i < P.B - 1; ++i) b = (u >= P.base[i]) ? i : b;
      th = P.slope[b] / Ef * invJT;
      const double remj = fmax(0.0, L.rem[ji] - L.dbar[ji] * Pj);
      if (mx > 0.0 && remj >= mx - band && cnt > 0.0) th += prm.k * L.dbar[ji] / cnt;
    }
    L.theta[sj] = (float)(th * L.theta_scale);
    if (Q_ROW_ATOMICS(L.T)) L.rowp[sj] = 0.f;
  }
  // first call: start every price at the mean marginal density of that worker type
  __shared__ double s_init[SWB_MK_MAXW];
  if (L.init_price) {
    for (int w = 0; w < L.W; ++w) {
      double num = 0.0, den = 0.0;
      for (int j = threadIdx.x; j < J; j += blockDim.x) {
        const size_t sj = (size_t)s * J + j, ji = L.per_scn ? sj : j;
        num += (double)L.theta[sj] * (double)L.rate[ji * L.W + w];
        den += (double)L.g[ji];
      }
      br.sum2(num, den);
      if (threadIdx.x == 0) s_init[w] = den > 0.0 ? num / den : 0.0;
      __syncthreads();
    }
  }
  double viol = 0.0;
  for (int i = threadIdx.x; i < WT; i += blockDim.x) {
    const size_t si = (size_t)s * WT + i;
    const float cap = (float)L.Gw[i / L.T];
    const float load = L.colload[si];
    viol = fmax(viol, (double)(load / cap - 1.f));
    L.colscale[si] = (load > cap && load > 0.f) ? cap / load : 1.f;
    const float p0 = L.init_price ? (float)s_init[i / L.T] : L.price[si];
    L.price[si] = fmaxf(1e-30f, p0 * __expf(L.sigma * (load / cap - 1.f)));   // tatonnement, multiplicative
    L.colload[si] = 0.f;
  }
  viol = br.max(viol);
  // objective and worst relative capacity violation of the X the previous dense pass wrote
  if (threadIdx.x == 0) { L.obj[3 * s] = welfare - prm.k * mx; L.obj[3 * s + 1] = mx; L.obj[3 * s + 2] = viol; }
}

cudaError_t launch_market_iter(const MarketLaunch &L, cudaStream_t st, bool dense) {
  if (!dense) {
    market_dual_kernel<<<L.S, 1024, 0, st>>>(L);
    return cudaGetLastError();
  }
  const int Q = L.T / 4;
  dim3 grid((L.J + L.jobs_per_cta - 1) / L.jobs_per_cta, L.S);
  const size_t smem = 3 * (size_t)L.W * L.T * sizeof(float);
  (void)Q;
  // fast path: T in {32, 64, 128} and a scenario's slice of X addressable with 32-bit float4 offsets
  if ((size_t)L.J * L.W * (L.T / 4) < (1u << 30)) {
    bool done = false;
    switch (L.W) {
      case 1: done = launch_fast<1, 4>(L, grid, st); break;
      case 2: done = launch_fast<2, 4>(L, grid, st); break;
      case 3: done = launch_fast<3, 2>(L, grid, st); break;
      default: done = launch_fast<4, 2>(L, grid, st); break;
    }
    if (done) return cudaGetLastError();
  }
  switch (L.W) {
    case 1: market_step_kernel<1, 4><<<grid, MK_THREADS, smem, st>>>(L); break;
    case 2: market_step_kernel<2, 4><<<grid, MK_THREADS, smem, st>>>(L); break;
    case 3: market_step_kernel<3, 2><<<grid, MK_THREADS, smem, st>>>(L); break;
    default: market_step_kernel<4, 2><<<grid, MK_THREADS, smem, st>>>(L); break;
  }
  return cudaGetLastError();
}

}  // namespace swb
