// gbm.cu — Monte-Carlo throughput forecast: geometric-Brownian-motion sample paths of the per-epoch
// duration of every active job.  NOT in the reference (the reference's only forecast is
// the deterministic Dirichlet-posterior remaining runtime, JobMetaData.py:315-370) — this kernel is the
// stochastic generalisation BASELINE.json's north_star asks for, and it reduces to that quantity
// exactly when sigma = mu = 0:
//     R_j^(p) = R0_j * (1/H_j) * sum_{h=1..H_j} exp((mu_j - sigma_j^2/2) h + sigma_j W_h^(p)),
//     W_h = sum of h iid N(0,1)   (a running sum = the prefix scan along the path),
// R0_j the deterministic forecast.  Per job the kernel returns sum_p R^(p) and sum_p (R^(p))^2 over the
// LOCAL paths; paths shard across GPUs and ONE allreduce of the [2][J] float64 vector follows
// (SURVEY.md §8e).  Random streams are keyed by (seed, job, GLOBAL pair id) — global paths 2q and 2q+1 are an
// antithetic pair driven by the same normals with opposite signs — so results do not depend on how the paths are
// sharded (up to float64 summation order).
// Bound: ALU/SFU (xorshift128+, Box-Muller, exp), not HBM: 40 bytes in and 16 bytes out per job.
#include <math.h>

#include "swb_common.cuh"
#include "swb_internal.h"

namespace swb {

__device__ __forceinline__ unsigned long long splitmix64(unsigned long long &x) {
  unsigned long long z = (x += 0x9E3779B97F4A7C15ull);
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}

struct XorShift128p {
  unsigned long long s0, s1;
  __device__ __forceinline__ unsigned long long next() {
    unsigned long long x = s0;
    const unsigned long long y = s1;
    s0 = y;
    x ^= x << 23;
    s1 = x ^ y ^ (x >> 17) ^ (y >> 26);
    return s1 + y;
  }
};

__device__ __forceinline__ float ex2_approx(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float lg2_approx(float x) { float y; asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float sqrt_approx(float x) { float y; asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }

__global__ void __launch_bounds__(256) gbm_kernel(GbmLaunch L) {
  __shared__ double red[2 * 64];
  const int j = blockIdx.x;
  int H;
  float mu, sg;
  if (L.slots) {
    const int slot = L.slots[j];
    mu = (float)L.tab_mu[slot]; sg = (float)L.tab_sigma[slot];
    H = L.Eo[j] - L.co[j];
    H = H < 0 ? 0 : (H > L.Hmax ? L.Hmax : H);
    if (mu == 0.0f && sg == 0.0f) H = 0;      // deterministic model: nothing to simulate (R = R0 on every path)
  } else {
    H = L.H[j];
    mu = (float)L.mu[j]; sg = (float)L.sigma[j];
  }
  const float drift = mu - 0.5f * sg * sg;
  const float drift2 = drift * 1.4426950408889634f, sg2 = sg * 1.4426950408889634f;   // exponents in base 2
  const float invH = H > 0 ? 1.0f / (float)H : 0.0f;
  const double R0 = L.R0[j];
  double s1 = 0.0, s2 = 0.0;
  // ANTITHETIC PAIRS: global paths 2q and 2q+1 share one stream of normals with opposite signs (W and -W) — half
  // the random numbers, logs, square roots and sin/cos per path, and a lower variance of the mean for free.  A
  // thread walks whole pairs; a shard boundary that cuts a pair (odd offset / count) simply skips the missing member.
  const long long first = L.path_offset, last = L.path_offset + L.P_local;     // [first, last)
  const long long q0 = first >> 1, q1 = (last + 1) >> 1;                         // pair ids touched
  for (long long q = q0 + threadIdx.x; q < q1; q += blockDim.x) {
    const unsigned long long gq = (unsigned long long)q;
    unsigned long long sm = L.seed ^ (0xD1B54A32D192ED03ull * (unsigned long long)(j + 1)) ^ (gq * 0x9E3779B97F4A7C15ull);
    XorShift128p rng;
    rng.s0 = splitmix64(sm);
    rng.s1 = splitmix64(sm) | 1ull;
    float W = 0.0f, accA = 0.0f, accB = 0.0f, hf = 1.0f;
    for (int h = 1; h <= H; h += 2, hf += 2.0f) {
      // Box-Muller: two normals per pair of 24-bit uniforms.  Every transcendental is ONE MUFU instruction (the
      // kernel is bound by instruction issue and the MUFU pipe, ncu r02): exponents are kept in base 2 (drift and
      // sigma pre-scaled by log2 e), the radius is sqrt(-2 ln2 * lg2 u1).
      const unsigned long long r = rng.next();
      const float u1 = ((float)((unsigned)(r >> 40)) + 1.0f) * (1.0f / 16777217.0f);  // (0,1)
      const float u2 = (float)((unsigned)(r >> 8) & 0xffffffu) * (1.0f / 16777216.0f);
      const float rad = sqrt_approx(-1.3862943611198906f * lg2_approx(u1));
      const float ang = 6.283185307179586f * u2;
      const float cs = __cosf(ang), sn = __sinf(ang);
      W = fmaf(rad, cs, W);
      float dh = drift2 * hf, sw = sg2 * W;
      accA += ex2_approx(dh + sw);
      accB += ex2_approx(dh - sw);
      if (h + 1 <= H) {
        W = fmaf(rad, sn, W);
        dh += drift2; sw = sg2 * W;
        accA += ex2_approx(dh + sw);
        accB += ex2_approx(dh - sw);
      }
    }
    const long long pa = 2 * q, pb = 2 * q + 1;
    if (pa >= first && pa < last) { const double R = H > 0 ? R0 * (double)(accA * invH) : R0; s1 += R; s2 += R * R; }
    if (pb >= first && pb < last) { const double R = H > 0 ? R0 * (double)(accB * invH) : R0; s1 += R; s2 += R * R; }
  }
  BlockRed br(red);
  br.sum2(s1, s2);
  if (threadIdx.x == 0) { L.out[j] = s1; L.out[L.J + j] = s2; }
}

cudaError_t launch_gbm(const GbmLaunch &L, cudaStream_t st) {
  gbm_kernel<<<L.J, 256, 0, st>>>(L);
  return cudaGetLastError();
}

__global__ void gbm_apply_kernel(GbmApplyLaunch L) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= L.J) return;
  const double mean = L.sums[j] / L.P_total;
  const double var = fmax(0.0, L.sums[L.J + j] / L.P_total - mean * mean);
  if (L.var_out) L.var_out[j] = var;
  const int slot = L.slots[j];
  if (L.tab_mu[slot] == 0.0 && L.tab_sigma[slot] == 0.0) return;   // deterministic job: forecast untouched
  const double R0 = L.rem[j];
  if (!(R0 > 0.0)) return;
  const double ratio = mean / R0;
  L.rem[j] = mean;
  if (L.rem_fb) L.rem_fb[j] *= ratio;
  if (L.bfkey) L.bfkey[j] *= ratio;
  if (L.bfkey_fb) L.bfkey_fb[j] *= ratio;
}

cudaError_t launch_gbm_apply(const GbmApplyLaunch &L, cudaStream_t st) {
  gbm_apply_kernel<<<(L.J + 255) / 256, 256, 0, st>>>(L);
  return cudaGetLastError();
}

__global__ void gbm_ensemble_kernel(int S, int J, double P, const double *sums, const double *z, double *rem_out) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= J) return;
  const double mean = sums[j] / P;
  const double sd = sqrt(fmax(0.0, sums[J + j] / P - mean * mean));
  for (int s = blockIdx.y; s < S; s += gridDim.y) rem_out[(size_t)s * J + j] = fmax(0.0, fma(z[s], sd, mean));
}

cudaError_t launch_gbm_ensemble(int S, int J, double P_total, const double *sums, const double *z, double *rem_out,
                                cudaStream_t st) {
  dim3 grid((J + 255) / 256, S < 64 ? S : 64);
  gbm_ensemble_kernel<<<grid, 256, 0, st>>>(S, J, P_total, sums, z, rem_out);
  return cudaGetLastError();
}

}  // namespace swb
