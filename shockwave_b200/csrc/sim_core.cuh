// Round loop of the reference simulator for trace-driven runs of STATIC jobs on one worker type, one CTA per what-if
// scenario: `Scheduler.simulate()` scheduler/scheduler.py:1878-2250 with the parts of `_get_job_steps_and_finish_times`
// :1467-1512, `_get_num_steps` :1425-1465, `_done_callback` :4341-4700, `_update_throughput` :549-571, `_remove_job`
// :808-830 and `_update_shockwave_scheduler` :2270-2342 that act on such jobs.
//
// One step = "run round k with the jobs the policy chose, jump to the next event, book the progress, retire finished
// jobs, admit arrivals": phases (C) and (A) of the reference's loop fused, so the policy's selection is the only thing
// that happens between two steps.  All arithmetic is the reference's own, in its order, in IEEE double / int64:
//   steps of the round     n   = min(int(throughput * time_per_iteration), total_steps - steps_run)
//   finish time            fin = now + n / throughput
//   next timestamp         max over the running jobs (else the next arrival)
//   preemption overhead    a job that did NOT run in the previous round loses 20 s of a (nearly) full round:
//                          slow = (ex - 20) / ex, ex -= 20, n = int(n * slow)               (scheduler.py:1935-1965)
//   over-deadline rule     cumulative run time > int(1.5 * duration) retires the job       (scheduler.py:4378-4395)
//   measured throughput    n / ex, appended to the job's timeline under the round index     (scheduler.py:549-571)
//   timeline summary       ns += bs * (thr * round_duration * (round - previous round))      (JobMetaData.py:235-249,
//                          kept as a running sum: what ShockwaveScheduler's forecast reads)
//   epoch progress         floor(steps_run / ceil(dataset / batch size))                    (scheduler.py:2332-2335)
//
// Dynamic adaptation (accordion / gns batch-size rescaling) is table-driven: `_simulate_accordion` :1658-1727 and
// `_simulate_gns` :1604-1656 raise a request from the job's epoch (tables: critical-regime flag / gns batch size per
// epoch), `_scale_bs_and_iters` :4731-4935 applies it at the job's next completion callback (new batch size, throughput
// of that batch size from the throughput file, total steps and progress rescaled with the epoch count preserved).
// The micro-task failure branch (:4497-4570: a round with no steps and no run time; five in a row drop the job) is part
// of the loop: the canonical trace reaches it after a rescale rounds a job's progress up past its total.
//
// The SAME source compiles for the host (one "thread", barriers are no-ops): tests/native/sim_host.cpp builds it with
// g++ so the bookkeeping is checked bit for bit against the pinned restatement (oracle/sim_loop.py) without a GPU.
// That host build is test tooling only — libswb200 has no CPU path.
#pragma once
#include <math.h>
#include <stdint.h>

#ifdef __CUDACC__
#define SIM_HD __host__ __device__ __forceinline__
#else
#define SIM_HD inline
#endif

#ifdef __CUDA_ARCH__
#define SIM_TID ((int)threadIdx.x)
#define SIM_NT ((int)blockDim.x)
#define SIM_SYNC() __syncthreads()
// no FMA contraction where the reference rounds twice (a * b, then + c)
#define SIM_MUL(a, b) __dmul_rn((a), (b))
#define SIM_ADD(a, b) __dadd_rn((a), (b))
#else
#define SIM_MUL(a, b) ((a) * (b))
#define SIM_ADD(a, b) ((a) + (b))
#define SIM_TID 0
#define SIM_NT 1
#define SIM_SYNC() ((void)0)
#endif

namespace swb {
namespace sim {

enum { ERR_CAPACITY = 1, ERR_NO_EVENT = 2, ERR_PATTERN = 4, ERR_THROUGHPUT = 8 };
#define SIM_MAX_TYPES 8
#define SIM_MAX_FAILED_ATTEMPTS 5 /* scheduler.py:53 */
#define SIM_MAX_LEVELS 8
enum { QUEUED = 0, LIVE = 1, COMPLETED = 2 };
#define SIM_PREEMPTION_OVERHEAD 20.0

struct Trace {            // [J], shared by the scenarios
  int J;
  const double *arrival;
  const long long *total;
  const int *sf;
  const double *thr;
  const double *dur15;    // (double) int(1.5 * duration)
  const int *bs;          // ORIGINAL batch size
  const long long *spe;   // steps per epoch at the original batch size = ceil(dataset / batch size)
  const long long *ds;    // dataset length
  // dynamic adaptation (all null / 0 when every job is static)
  const int *mode;        // 0 static, 1 accordion, 2 gns
  const int *bs_max, *bs_min, *bs_big, *orig_locked, *acc_skip;
  const long long *pat_off;   // [J + 1]
  const int *pattern;
  int K;                  // levels per job
  const int *lvl_bs;      // [J][K]
  const double *lvl_thr;  // [J][K]
  // several worker types (static jobs only; null / 0 = one type): chosen[j] = 1 + index of the type the job runs on
  int W;
  const double *thr_w;    // [J][W] steps/s of the job on each type (<= 0: the job cannot run there, scheduler.py:1494-1504)
  const int *cap_w;       // [W] workers of each type
};

struct Scn {              // == swb_sim_scn
  double now, round_start, round_end;   // round_end: NaN = the reference's None
  int rounds, remaining, n_active, done, err, pad;
};

struct State {            // [J] of ONE scenario
  unsigned char *status, *ranprev;
  long long *steps_run, *nsteps;
  double *run_time, *latest, *jct, *fin, *tl_ns, *thr_meas;
  int *tl_prev, *tl_end, *epoch;
  // mutable copies of the trace constants a rescale changes + request / failure counters
  unsigned char *running, *flag, *fails;
  int *cbs;
  long long *ctotal, *cspe;
  double *cthr;
  double *last_ex;        // execution time booked for the job's latest round (the Gavel time accounting reads it)
};

struct Shared {           // SIM_NT entries each
  double *sd;
  long long *si;
};

SIM_HD double blk_max(double v, const Shared &sh) {
#ifdef __CUDA_ARCH__
  // warp shuffles, then one value per warp through shared memory (max is order-independent: same result as the loop)
  for (int o = 16; o; o >>= 1) {
    const double w = __shfl_xor_sync(0xffffffffu, v, o);
    v = w > v ? w : v;
  }
  if ((SIM_TID & 31) == 0) sh.sd[SIM_TID >> 5] = v;
  SIM_SYNC();
  double r = sh.sd[0];
  for (int i = 1; i < (SIM_NT >> 5); ++i) r = sh.sd[i] > r ? sh.sd[i] : r;
  SIM_SYNC();
  return r;
#else
  sh.sd[SIM_TID] = v;
  SIM_SYNC();
  if (SIM_TID == 0) {
    double r = sh.sd[0];
    for (int i = 1; i < SIM_NT; ++i) r = sh.sd[i] > r ? sh.sd[i] : r;
    sh.sd[0] = r;
  }
  SIM_SYNC();
  const double r = sh.sd[0];
  SIM_SYNC();
  return r;
#endif
}
SIM_HD double blk_min(double v, const Shared &sh) { return -blk_max(-v, sh); }
SIM_HD long long blk_sum(long long v, const Shared &sh) {
#ifdef __CUDA_ARCH__
  for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);      // integer sum: exact in any order
  if ((SIM_TID & 31) == 0) sh.si[SIM_TID >> 5] = v;
  SIM_SYNC();
  long long r = 0;
  for (int i = 0; i < (SIM_NT >> 5); ++i) r += sh.si[i];
  SIM_SYNC();
  return r;
#else
  sh.si[SIM_TID] = v;
  SIM_SYNC();
  if (SIM_TID == 0) {
    long long r = 0;
    for (int i = 0; i < SIM_NT; ++i) r += sh.si[i];
    sh.si[0] = r;
  }
  SIM_SYNC();
  const long long r = sh.si[0];
  SIM_SYNC();
  return r;
#endif
}

SIM_HD long long ceil_div(long long a, long long b) { return (a + b - 1) / b; }

// _simulate_accordion :1658-1727 / _simulate_gns :1604-1656: raise a rescale request from the job's progress.
// Returns 1 when the job's epoch is outside its table.
SIM_HD int request(const Trace &T, const State &X, int j) {
  const long long cur = ceil_div(X.steps_run[j], X.cspe[j]);          // _get_num_epochs :4722-4729
  const int *pat = T.pattern + T.pat_off[j];
  const long long len = T.pat_off[j + 1] - T.pat_off[j];
  const int bs = X.cbs[j], orig = T.bs[j];
  if (T.mode[j] == 1) {
    if (T.acc_skip[j]) return 0;
    if (cur >= len) return 1;
    const bool crit = pat[cur] != 0;
    if (bs == orig && !crit) {
      if (bs != T.bs_max[j]) X.flag[j] = 1;
    } else if (bs != orig && crit) {
      if (bs != T.bs_min[j]) X.flag[j] = 2;
    }
  } else {
    if (cur + 1 >= len) return 1;
    // get_gns_bs_pattern(.., max(760, cur + 2), ..) never scales its last entry (utils.py:801-1010)
    const int nxt = cur + 1 >= 759 ? orig : pat[cur + 1];
    if (nxt > bs || pat[cur] > bs) {
      if (bs != T.bs_max[j]) X.flag[j] = 1;
    }
  }
  return 0;
}

// _scale_bs_and_iters :4731-4935 for a job whose request flag is up
SIM_HD int rescale(const Trace &T, const State &X, int j) {
  if (T.orig_locked[j]) return 0;
  const int old = X.cbs[j];
  const int nw = T.mode[j] == 2 ? 2 * old : (X.flag[j] == 1 ? T.bs_big[j] : T.bs[j]);
  int k = -1;
  for (int i = 0; i < T.K; ++i)
    if (T.lvl_bs[(size_t)j * T.K + i] == nw && T.lvl_thr[(size_t)j * T.K + i] > 0.0) k = i;
  if (k < 0 || nw <= 0) return 0;                    // batch size not in the throughput file: request dropped (:4803-4818)
  const double factor = (double)nw / (double)old;
  const double it = 1.0 / factor;
  const long long spe_old = ceil_div(T.ds[j], old), spe_new = ceil_div(T.ds[j], nw);
  const long long old_epochs = ceil_div(X.ctotal[j], spe_old);
  long long new_total = (long long)ceil(SIM_MUL((double)X.ctotal[j], it));
  if (ceil_div(new_total, spe_new) != old_epochs) new_total = spe_new * old_epochs;
  const long long done_epochs = ceil_div(X.steps_run[j], spe_old);
  X.ctotal[j] = new_total;
  X.steps_run[j] = done_epochs * spe_new;
  X.cspe[j] = spe_new;
  X.cbs[j] = nw;
  X.cthr[j] = T.lvl_thr[(size_t)j * T.K + k];
  return 0;
}

// admission of the jobs that have arrived by `now` (scheduler.py:2040-2052) + count of the live jobs
SIM_HD int admit(const Trace &T, const State &X, double now, const Shared &sh) {
  long long live = 0;
  for (int j = SIM_TID; j < T.J; j += SIM_NT) {
    if (X.status[j] == QUEUED && T.arrival[j] <= now) X.status[j] = LIVE;
    live += X.status[j] == LIVE;
  }
  return (int)blk_sum(live, sh);
}

// iteration 0 of the loop: current timestamp = first arrival (scheduler.py:1847), nothing has run yet
SIM_HD void scenario_begin(const Trace &T, const State &X, Scn *scn, const Shared &sh) {
  for (int j = SIM_TID; j < T.J; j += SIM_NT) {
    X.status[j] = QUEUED; X.ranprev[j] = 0; X.steps_run[j] = 0; X.nsteps[j] = -1;
    X.run_time[j] = 0.0; X.latest[j] = NAN; X.jct[j] = NAN; X.fin[j] = 0.0; X.tl_ns[j] = 0.0; X.thr_meas[j] = 0.0;
    X.tl_prev[j] = 0; X.tl_end[j] = -1; X.epoch[j] = 0;
    X.running[j] = 0; X.flag[j] = 0; X.fails[j] = 0;
    X.cbs[j] = T.bs[j]; X.ctotal[j] = T.total[j]; X.cspe[j] = T.spe[j]; X.cthr[j] = T.thr[j]; X.last_ex[j] = 0.0;
  }
  SIM_SYNC();
  const double now = T.arrival[0];
  const int live = admit(T, X, now, sh);
  if (SIM_TID == 0) {
    scn->now = now; scn->round_start = 0.0; scn->round_end = NAN;      // scheduler.py:1819-1820
    scn->rounds = 0; scn->remaining = T.J; scn->n_active = live; scn->done = live == 0; scn->err = 0; scn->pad = 0;
  }
  SIM_SYNC();
}

// run round `scn->rounds` with the jobs flagged in chosen[J], then the head of the next iteration
SIM_HD void scenario_step(const Trace &T, const State &X, Scn *scn, const unsigned char *chosen, int ngpus, double tpi,
                          double grd, const Shared &sh) {
  Scn z = *scn;
  SIM_SYNC();
  if (z.done) return;
  // ---- (C) steps and finish times of the chosen jobs (scheduler.py:2224-2236)
  double mx = 0.0;
  long long used = 0, nrun = 0, badthr = 0;
  const bool het = T.thr_w != nullptr;                         // several worker types: throughput of the round's type
  long long used_w[SIM_MAX_TYPES];
  for (int w = 0; w < SIM_MAX_TYPES; ++w) used_w[w] = 0;
  for (int j = SIM_TID; j < T.J; j += SIM_NT) {
    if (chosen[j] && X.status[j] == LIVE) {
      double th = X.cthr[j];
      if (het) {
        int w = (int)chosen[j] - 1;
        if (w >= T.W) { w = T.W - 1; used_w[w] += (long long)ngpus + 1; }     // unknown type: reported as a capacity error
        th = T.thr_w[(size_t)j * T.W + w];
        used_w[w] += T.sf[j];
      }
      if (!(th > 0.0)) {                                        // the reference raises here (scheduler.py:1494-1504)
        ++badthr; X.running[j] = 0;
        continue;
      }
      const long long rem = X.ctotal[j] - X.steps_run[j];      // negative after a rescale that rounds the progress up
      long long n = (long long)SIM_MUL(th, tpi);
      if (n > rem) n = rem;
      double fin = z.now + (double)n / th;
      if (!(fin > z.now)) fin = z.now;                          // max_finish_time starts at the current timestamp (:1470)
      X.nsteps[j] = n; X.fin[j] = fin; X.running[j] = 1;
      mx = fin > mx ? fin : mx;
      used += T.sf[j]; ++nrun;
    } else {
      X.running[j] = 0;
    }
  }
  mx = blk_max(mx, sh);
  used = blk_sum(used, sh);
  nrun = blk_sum(nrun, sh);
  if (het) {
    for (int w = 0; w < T.W; ++w)
      if (blk_sum(used_w[w], sh) > (long long)T.cap_w[w]) z.err |= ERR_CAPACITY;
    if (blk_sum(badthr, sh)) z.err |= ERR_THROUGHPUT;
  } else if (used > ngpus) z.err |= ERR_CAPACITY;
  // ---- head of iteration c = k + 1: jump to the next event (scheduler.py:1901-1915)
  const int c = z.rounds + 1;
  double max_ts = 0.0;
  if (nrun > 0 && mx > 0.0) {
    max_ts = mx;
    if (!isnan(z.round_end)) z.round_start = z.round_end;
    z.round_end = mx;
  }
  double na = INFINITY;
  for (int j = SIM_TID; j < T.J; j += SIM_NT)
    if (X.status[j] == QUEUED && T.arrival[j] < na) na = T.arrival[j];
  na = blk_min(na, sh);
  if (max_ts > 0.0) z.now = max_ts;
  else if (na < INFINITY) z.now = na;
  else { z.err |= ERR_NO_EVENT; z.done = 1; }      // the reference would add None to a float here
  // ---- completions of the round (scheduler.py:1921-2014 + _done_callback)
  long long completed = 0, bad = 0;
  if (!z.done)
    for (int j = SIM_TID; j < T.J; j += SIM_NT) {
      const bool ran = X.running[j] != 0;
      if (ran) {
        double ex = X.fin[j] - z.round_start;
        long long n = X.nsteps[j];
        if (c != 1 && !X.ranprev[j]) {
          if (ex != 0.0 && tpi - 5.0 < ex) {
            const double slow = (ex - SIM_PREEMPTION_OVERHEAD) / ex;
            ex -= SIM_PREEMPTION_OVERHEAD;
            n = (long long)SIM_MUL((double)n, slow);
          }
        }
        X.latest[j] = X.fin[j];
        X.run_time[j] += ex;
        X.last_ex[j] = ex;
        const bool over = X.run_time[j] > T.dur15[j];
        bool done;
        if (n <= 0 && ex <= 0.0) {                  // micro-task failure (scheduler.py:4497-4570): nothing is booked
          X.fails[j] = (unsigned char)(X.fails[j] + 1);
          done = X.fails[j] >= SIM_MAX_FAILED_ATTEMPTS;
        } else {
          X.fails[j] = 0;
          X.steps_run[j] += n;
          done = X.ctotal[j] - X.steps_run[j] <= 0 || over;
        }
        const double tm = ex <= 0.0 ? 0.0 : (double)n / ex;
        X.thr_meas[j] = tm;
        X.tl_ns[j] = SIM_ADD(X.tl_ns[j], SIM_MUL((double)X.cbs[j], SIM_MUL(SIM_MUL(tm, grd), (double)(c - X.tl_prev[j]))));
        X.tl_prev[j] = c; X.tl_end[j] = c;
        if (X.flag[j] && T.mode) bad += rescale(T, X, j);
        X.flag[j] = 0;                              // reset in every completion callback (scheduler.py:4700-4712)
        if (done) {
          X.status[j] = COMPLETED;
          X.jct[j] = X.latest[j] - T.arrival[j];
          X.epoch[j] = -1;                          // the caller retires the job (its progress = all epochs)
          ++completed;
        } else {
          X.epoch[j] = (int)(X.steps_run[j] / X.cspe[j]);
        }
      }
      X.ranprev[j] = ran;
      if (T.mode && T.mode[j] && X.status[j] == LIVE) bad += request(T, X, j);     // scheduler.py:2019-2027
    }
  completed = blk_sum(completed, sh);
  bad = blk_sum(bad, sh);
  if (bad) z.err |= ERR_PATTERN;
  z.remaining -= (int)completed;
  // ---- arrivals (scheduler.py:2040-2052), end-of-run tests (:1895-1899, :2173-2178)
  int live = z.n_active;
  if (!z.done) {
    live = admit(T, X, z.now, sh);
    z.rounds = c;
    z.done = (z.remaining == 0 || live == 0) ? 1 : 0;
  }
  z.n_active = live;
  if (SIM_TID == 0) *scn = z;
  SIM_SYNC();
}

}  // namespace sim
}  // namespace swb
