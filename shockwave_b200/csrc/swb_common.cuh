// Shared device helpers for libswb200 (sm_100a).  No tensor cores anywhere: every kernel here is an
// elementwise / row-column reduction path, so the tools are warp shuffles,
// shared-memory staging and vectorised global accesses.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/swb200.h"

#define SWB_WARP 32
#define SWB_FULL 0xffffffffu

namespace swb {

// Piecewise-linear log through (base_b, logv_b): the reference's SOS2 model (shockwave.py:384-419)
// evaluated directly — for a concave maximise the LP picks adjacent breakpoints, so
// plog(u) == min_b(slope_b * (u - base_b) + logv_b) on [0,1].
struct Pwl {
  int B;
  double base[SWB_MAX_BASES];
  double logv[SWB_MAX_BASES];
  double slope[SWB_MAX_BASES];  // slope[b] of segment [base[b], base[b+1]], b < B-1
};

__device__ __forceinline__ double plog(const Pwl &P, double u) {
  u = fmax(u, 0.0);
  // exact at the last breakpoint (lambda_B = 1 in the reference's model): the fallback priorities
  // reach 1e21 and would turn the ~1e-17 residual of slope*(1-base)+logv into O(10) of objective
  if (u >= P.base[P.B - 1]) return P.logv[P.B - 1];
  int b = 0;
  for (int i = 1; i < P.B - 1; ++i) b = (u >= P.base[i]) ? i : b;
  return fma(P.slope[b], u - P.base[b], P.logv[b]);
}

__device__ __forceinline__ int warp_sum(int v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(SWB_FULL, v, o);
  return v;
}
__device__ __forceinline__ long long warp_sum(long long v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(SWB_FULL, v, o);
  return v;
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(SWB_FULL, v, o);
  return v;
}
__device__ __forceinline__ double warp_max(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(SWB_FULL, v, o));
  return v;
}
__device__ __forceinline__ double warp_min(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmin(v, __shfl_xor_sync(SWB_FULL, v, o));
  return v;
}

// Block reductions with ONE __syncthreads each: two scratch rows used alternately (ping-pong), so
// a warp that races ahead into reduction k+1 writes the other row and cannot reach reduction k+2
// before every warp has left reduction k.  `scratch` is 2 x 32 x 16 bytes.
struct BlockRed {
  double *s;       // [2][32][2]
  int phase;
  __device__ __forceinline__ BlockRed(double *scratch) : s(scratch), phase(0) {}

  // (sum of a, sum of b) over the block, both double
  __device__ __forceinline__ void sum2(double &a, double &b) {
    a = warp_sum(a);
    b = warp_sum(b);
    double *row = s + phase * 64;
    const int w = threadIdx.x >> 5, l = threadIdx.x & 31, nw = (blockDim.x + 31) >> 5;
    if (l == 0) { row[2 * w] = a; row[2 * w + 1] = b; }
    __syncthreads();
    double ra = (l < nw) ? row[2 * l] : 0.0, rb = (l < nw) ? row[2 * l + 1] : 0.0;
    a = warp_sum(ra);
    b = warp_sum(rb);
    phase ^= 1;
  }
  __device__ __forceinline__ double sum(double a) { double b = 0.0; sum2(a, b); return a; }
  __device__ __forceinline__ long long sumll(long long a) {
    a = warp_sum(a);
    long long *row = reinterpret_cast<long long *>(s + phase * 64);
    const int w = threadIdx.x >> 5, l = threadIdx.x & 31, nw = (blockDim.x + 31) >> 5;
    if (l == 0) row[w] = a;
    __syncthreads();
    long long r = (l < nw) ? row[l] : 0;
    a = warp_sum(r);
    phase ^= 1;
    return a;
  }
  // int32 sum on the hardware warp reduction (REDUX.SUM): two instructions and one barrier instead of twenty
  // 32-bit shuffles — the demand sums of the price search are evaluated hundreds of times per solve
  __device__ __forceinline__ int sumi(int a) {
    a = __reduce_add_sync(SWB_FULL, a);
    int *row = reinterpret_cast<int *>(s + phase * 64);
    const int w = threadIdx.x >> 5, l = threadIdx.x & 31, nw = (blockDim.x + 31) >> 5;
    if (l == 0) row[w] = a;
    __syncthreads();
    a = __reduce_add_sync(SWB_FULL, (l < nw) ? row[l] : 0);
    phase ^= 1;
    return a;
  }
  // three int32 sums with ONE barrier (the price search evaluates three candidate prices per pass)
  __device__ __forceinline__ void sumi3(int &a, int &b, int &c) {
    a = __reduce_add_sync(SWB_FULL, a); b = __reduce_add_sync(SWB_FULL, b); c = __reduce_add_sync(SWB_FULL, c);
    int *row = reinterpret_cast<int *>(s + phase * 64);       // 128 ints per row: [32 warps][3]
    const int w = threadIdx.x >> 5, l = threadIdx.x & 31, nw = (blockDim.x + 31) >> 5;
    if (l == 0) { row[3 * w] = a; row[3 * w + 1] = b; row[3 * w + 2] = c; }
    __syncthreads();
    a = __reduce_add_sync(SWB_FULL, (l < nw) ? row[3 * l] : 0);
    b = __reduce_add_sync(SWB_FULL, (l < nw) ? row[3 * l + 1] : 0);
    c = __reduce_add_sync(SWB_FULL, (l < nw) ? row[3 * l + 2] : 0);
    phase ^= 1;
  }
  __device__ __forceinline__ double max(double a) {
    a = warp_max(a);
    double *row = s + phase * 64;
    const int w = threadIdx.x >> 5, l = threadIdx.x & 31, nw = (blockDim.x + 31) >> 5;
    if (l == 0) row[w] = a;
    __syncthreads();
    double r = (l < nw) ? row[l] : -1.0e300;
    a = warp_max(r);
    phase ^= 1;
    return a;
  }
  __device__ __forceinline__ double min(double a) { return -max(-a); }
};

}  // namespace swb
