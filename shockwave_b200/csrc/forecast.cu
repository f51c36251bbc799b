// forecast.cu — batched dynamic-adaptation forecast: one warp per active job.
//
// Replaces, for every job of a re-solve, the Python walk the reference does through JobMetaData:
//   calibrate_profiled_epoch_duration   scheduler/JobMetaData.py:225-288
//   dirichlet_posterior_remaining_runtime            JobMetaData.py:315-370
//   interpolate_epoch_duration                       scheduler/shockwave.py:322-324
//   finish_time_uniform_share + finish_time_momentumed_average   shockwave.py:88-120, :480-501
//
// The static profile of a job (prefix sums of the pre-profiled epoch durations, the bs schedule, the
// sorted bs modes and the per-mode mean duration) is resident in HBM (uploaded once by swb_job_add);
// per call only epoch_progress and the two-number summary of the throughput timeline travel.
//
// Calibration is stateful in the reference (the rescale divides by the CURRENT epoch duration), so
// the kernel replays the reference's call sequence on one scalar `amp` per job:
//   [share]   cal ; elapsed ; cal(in dirichlet) -> finish-time estimate        shockwave.py:101-112
//   [solve]   cal ; dbar ; cal(in dirichlet) -> rem                            shockwave.py:370, :560
//   [fallbk]  cal ; cal(in dirichlet) -> rem used for the priorities           shockwave.py:863-866
//   [backfill]cal(in dirichlet) -> sort key of construct_schedules             shockwave.py:261-267
// and emits both continuations (with / without the fallback) — commit_amp_kernel keeps the one the
// solve actually took.
#include <math.h>

#include "swb_common.cuh"
#include "swb_internal.h"

namespace swb {

struct JobTab {
  const double *pp;     // prefix sums, E+1 entries
  const int32_t *bs;    // E entries
  int E, nm;
  double ns;
  const int32_t *modes;
  const double *modemean;
  double grd;           // gavel_round_duration
  bool has_tl;
  double meas_ns;
  int meas_end;
};

// JobMetaData.py:225-288 on the scalar state `amp`
__device__ __forceinline__ double calibrate(const JobTab &t, double amp) {
  if (!t.has_tl) return amp;
  const double range = t.grd * (double)t.meas_end;
  // first epoch i with pp[i+1] > range (the reference breaks there); E if none
  int lo = 0, hi = t.E;
  while (lo < hi) {
    const int mid = (lo + hi) >> 1;
    if (t.pp[mid + 1] > range) hi = mid; else lo = mid + 1;
  }
  const int full = lo;                       // epochs that fit entirely
  const int iep = full < t.E ? full : t.E - 1;  // loop variable after the for (JobMetaData.py:253-262)
  double pre_ns = (double)full * t.ns;
  const double deficit = range - t.pp[full];
  if (deficit > 0.0) {
    const double dur = (t.pp[iep + 1] - t.pp[iep]) * amp;
    pre_ns += t.ns * deficit / dur;
  }
  if (t.meas_ns <= 0.0 || pre_ns <= 0.0 || fabs(t.meas_ns - pre_ns) / pre_ns <= 0.4) return amp;
  return pre_ns / t.meas_ns;
}

// Python >= 3.12's builtin sum() over floats is Neumaier-compensated (CPython Python/bltinmodule.c);
// the reference calls it at JobMetaData.py:330 and :347 and truncates the second result with int(),
// so the forecast is sensitive to the last ulp of these two sums.  The golden fixtures were made
// under Python 3.12, hence the same algorithm here.
struct PySum {
  double s = 0.0, c = 0.0;
  __device__ __forceinline__ void add(double x) {
    const double t = s + x;
    if (fabs(s) >= fabs(x)) c += (s - t) + x; else c += (x - t) + s;
    s = t;
  }
  __device__ __forceinline__ double value() const { return (c != 0.0 && isfinite(c)) ? s + c : s; }
};

// JobMetaData.py:315-370; whole warp cooperates on the histogram, scalars are lane-uniform
__device__ double dirichlet(const JobTab &t, int c, double &amp) {
  const int lane = threadIdx.x & 31;
  const int nobs = (c + 1 < t.E) ? c + 1 : t.E;
  const double prior = (double)t.E / (double)t.nm;
  double post[SWB_MAX_MODES];
  PySum csum_acc;
  int cnts[SWB_MAX_MODES];
#pragma unroll 1
  for (int m = 0; m < t.nm; ++m) {
    const int mode = t.modes[m];
    int cnt = 0;
    for (int e = lane; e < nobs; e += 32) cnt += (t.bs[e] == mode) ? 1 : 0;
    cnt = warp_sum(cnt);
    cnts[m] = cnt;
    post[m] = prior + (double)cnt;
    csum_acc.add(post[m]);
  }
  const double csum = csum_acc.value();
  PySum rsum_acc;
#pragma unroll 1
  for (int m = 0; m < t.nm; ++m) {
    double v = (double)t.E * post[m] / csum;
    const double fl = floor(v);
    const double dec = fmin((double)cnts[m], fl > 0.0 ? fl : 0.0);
    v -= dec;
    post[m] = v;  // rebased, after the observed epochs were taken off
    rsum_acc.add(v);
  }
  long long inflated = (long long)(rsum_acc.value() + 1.0);
  const long long rem_epochs = (long long)t.E - c;
  if (inflated < rem_epochs) inflated = rem_epochs;
  if (inflated <= 0 || rem_epochs <= 0) return 1.0;
  amp = calibrate(t, amp);  // get_bs_epoch_duration_map, JobMetaData.py:302
  double out = 0.0;
#pragma unroll 1
  for (int m = 0; m < t.nm; ++m) out += post[m] * (t.modemean[m] * amp);
  out *= (double)rem_epochs / (double)inflated;
  return out;
}

__global__ void __launch_bounds__(256) forecast_kernel(ForecastLaunch L) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (warp >= L.J) return;
  const int slot = L.slots[warp];
  JobTab t;
  const int64_t off = L.tab_off[slot];
  t.E = L.tab_E[slot];
  t.pp = L.pool_prefix + off;
  t.bs = L.pool_bs + off;
  t.nm = L.tab_nmodes[slot];
  t.ns = L.tab_nsamples[slot];
  t.modes = L.tab_modes + (size_t)slot * SWB_MAX_MODES;
  t.modemean = L.tab_modemean + (size_t)slot * SWB_MAX_MODES;
  t.grd = L.gavel_round_duration;
  t.meas_end = L.meas_end[warp];
  t.has_tl = t.meas_end >= 0;
  t.meas_ns = L.meas_ns[warp];
  const int c = L.progress[warp];
  double amp = L.tab_amp[slot];
  const double share = fmin(1.0, (double)L.ngpus / (double)L.J);

  // [share]  shockwave.py:92-118
  double ftest = 0.0;
  if (L.reestimate_share) {
    amp = calibrate(t, amp);
    const double elapsed = t.pp[c < t.E ? c : t.E] * amp;
    const double ra = dirichlet(t, c, amp);
    ftest = L.tab_tsubmit[slot] + (elapsed + ra) / share;
    if (lane == 0) {
      const int cnt = L.ss_cnt[slot];
      if (cnt == 0) { L.ss_r0[slot] = L.round_ptr; L.ss_acc[slot] = 0.0; }
      else L.ss_acc[slot] += (double)(L.round_ptr - L.ss_rlast[slot]) * L.ss_vlast[slot];
      L.ss_rlast[slot] = L.round_ptr; L.ss_vlast[slot] = ftest; L.ss_cnt[slot] = cnt + 1;
    }
    __syncwarp();
  }
  // finish_time_momentumed_average(share_series[j], round_ptr)   shockwave.py:480-501
  double ftobj;
  {
    const int cnt = L.reestimate_share ? 1 : L.ss_cnt[slot];
    if (cnt == 0) ftobj = nan("");
    else {
      const int r0 = L.ss_r0[slot], rl = L.ss_rlast[slot];
      const double vl = L.ss_vlast[slot];
      double running;
      if (L.round_ptr - r0 <= 0) running = vl;
      else running = (L.ss_acc[slot] + (double)(L.round_ptr - rl) * vl) / (double)(L.round_ptr - r0);
      ftobj = 0.9 * running + (1.0 - 0.9) * vl;
    }
  }
  // [solve]
  amp = calibrate(t, amp);
  const int nmean = (c + 1 < t.E) ? c + 1 : t.E;
  const double dbar = t.pp[nmean] * amp / (double)nmean;
  const double rem = dirichlet(t, c, amp);
  // continuation without fallback: back-fill key = one more dirichlet() call (shockwave.py:261-267);
  // the state is NOT advanced here — commit_calibration_kernel replays the calls that really happen
  const double amp_ok = amp;
  double tmp = amp_ok;
  const double bf_ok = dirichlet(t, c, tmp);
  // continuation with fallback (shockwave.py:863-866)
  double amp_fb = calibrate(t, amp);
  const double rem_fb = dirichlet(t, c, amp_fb);
  tmp = amp_fb;
  const double bf_fb = dirichlet(t, c, tmp);
  if (lane == 0) {
    L.dbar[warp] = dbar; L.rem[warp] = rem; L.ftobj[warp] = ftobj; L.ftest[warp] = ftest;
    L.bfkey[warp] = bf_ok; L.bfkey_fb[warp] = bf_fb; L.rem_fb[warp] = rem_fb;
    L.amp_ok[warp] = amp_ok; L.amp_fb[warp] = amp_fb;
    L.g_out[warp] = L.tab_g[slot]; L.E_out[warp] = t.E; L.c_out[warp] = c;
  }
}

// After the solve: keep the continuation the solve took and replay the calibrate() calls that
// construct_schedules' sort key makes — one per (round with idle GPUs, job not scheduled in it)
// (shockwave.py:254-267 -> JobMetaData.py:355 -> :302).  dirichlet() returns before calibrating when
// the job has no epochs left (JobMetaData.py:349-353).
__global__ void commit_calibration_kernel(ForecastLaunch L, const swb_result *res, int fallback_host,
                                          const int32_t *ncal) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= L.J) return;
  const int slot = L.slots[j];
  const bool fb = res ? (res->status == SWB_ST_FALLBACK) : (fallback_host != 0);
  double amp = fb ? L.amp_fb[j] : L.amp_ok[j];
  JobTab t;
  const int64_t off = L.tab_off[slot];
  t.E = L.tab_E[slot];
  t.pp = L.pool_prefix + off;
  t.bs = L.pool_bs + off;
  t.nm = L.tab_nmodes[slot];
  t.ns = L.tab_nsamples[slot];
  t.modes = nullptr; t.modemean = nullptr;
  t.grd = L.gavel_round_duration;
  t.meas_end = L.meas_end[j];
  t.has_tl = t.meas_end >= 0;
  t.meas_ns = L.meas_ns[j];
  const int n = (L.progress[j] < t.E) ? ncal[j] : 0;
  for (int i = 0; i < n; ++i) amp = calibrate(t, amp);
  L.tab_amp[slot] = amp;
}

cudaError_t launch_commit_calibration(const ForecastLaunch &L, const swb_result *res, int fallback_host,
                                      const int32_t *ncal, cudaStream_t st) {
  commit_calibration_kernel<<<(L.J + 127) / 128, 128, 0, st>>>(L, res, fallback_host, ncal);
  return cudaGetLastError();
}

cudaError_t launch_forecast(const ForecastLaunch &L, cudaStream_t st) {
  const int warps_per_block = 8;
  const int blocks = (L.J + warps_per_block - 1) / warps_per_block;
  forecast_kernel<<<blocks, warps_per_block * 32, 0, st>>>(L);
  return cudaGetLastError();
}

}  // namespace swb
