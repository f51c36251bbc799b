// Host build of shockwave_b200/csrc/sim_core.cuh (one "thread", barriers compiled out): lets the CPU tests run the
// device round loop's bookkeeping against the pinned restatement (oracle/sim_loop.py) without a GPU.
// TEST TOOLING ONLY: nothing under shockwave_b200/ loads this library, and libswb200.so has no CPU path.
#include <stdlib.h>
#include <vector>
#include "../../shockwave_b200/csrc/sim_core.cuh"

using namespace swb::sim;

struct HostSim {
  int J;
  std::vector<double> arrival, thr, dur15;
  std::vector<long long> total, spe, ds;
  std::vector<int> sf, bs;
  std::vector<unsigned char> status, ranprev;
  std::vector<long long> steps_run, nsteps;
  std::vector<double> run_time, latest, jct, fin, tl_ns, thr_meas;
  std::vector<int> tl_prev, tl_end, epoch;
  std::vector<unsigned char> running, flag, fails;
  std::vector<int> cbs;
  std::vector<long long> ctotal, cspe;
  std::vector<double> cthr, last_ex;
  // dynamic tables
  std::vector<int> mode, bs_max, bs_min, bs_big, orig_locked, acc_skip, pattern, lvl_bs;
  std::vector<long long> pat_off;
  std::vector<double> lvl_thr;
  std::vector<double> thr_w;
  std::vector<int> cap_w;
  Scn scn;
  double sd[1];
  long long si[1];
  Trace T;
  State X;
  Shared sh;
};

extern "C" void *sim_host_create(int J, const double *arrival, const long long *total, const int *sf, const double *thr,
                                 const double *duration, const int *bs, const long long *dataset_len) {
  HostSim *h = new HostSim;
  h->J = J;
  h->arrival.assign(arrival, arrival + J); h->thr.assign(thr, thr + J); h->total.assign(total, total + J);
  h->sf.assign(sf, sf + J); h->bs.assign(bs, bs + J); h->ds.assign(dataset_len, dataset_len + J);
  h->dur15.resize(J); h->spe.resize(J);
  for (int j = 0; j < J; ++j) {
    h->dur15[j] = (double)(long long)(duration[j] * 1.5);
    h->spe[j] = (dataset_len[j] + bs[j] - 1) / bs[j];
  }
  h->status.resize(J); h->ranprev.resize(J); h->steps_run.resize(J); h->nsteps.resize(J);
  h->run_time.resize(J); h->latest.resize(J); h->jct.resize(J); h->fin.resize(J); h->tl_ns.resize(J); h->thr_meas.resize(J);
  h->tl_prev.resize(J); h->tl_end.resize(J); h->epoch.resize(J);
  h->running.resize(J); h->flag.resize(J); h->fails.resize(J); h->cbs.resize(J); h->ctotal.resize(J); h->cspe.resize(J);
  h->cthr.resize(J); h->last_ex.resize(J);
  h->T = Trace{};
  h->T.J = J; h->T.arrival = h->arrival.data(); h->T.total = h->total.data(); h->T.sf = h->sf.data();
  h->T.thr = h->thr.data(); h->T.dur15 = h->dur15.data(); h->T.bs = h->bs.data(); h->T.spe = h->spe.data(); h->T.ds = h->ds.data();
  h->X = State{h->status.data(), h->ranprev.data(), h->steps_run.data(), h->nsteps.data(), h->run_time.data(),
               h->latest.data(), h->jct.data(), h->fin.data(), h->tl_ns.data(), h->thr_meas.data(), h->tl_prev.data(),
               h->tl_end.data(), h->epoch.data(), h->running.data(), h->flag.data(), h->fails.data(), h->cbs.data(),
               h->ctotal.data(), h->cspe.data(), h->cthr.data(), h->last_ex.data()};
  h->sh = Shared{h->sd, h->si};
  return h;
}

extern "C" void sim_host_set_dynamic(void *p, const int *mode, const int *bs_max, const int *bs_min, const int *bs_big,
                                     const int *orig_locked, const int *acc_skip, const long long *pat_off,
                                     const int *pattern, int K, const int *lvl_bs, const double *lvl_thr) {
  HostSim *h = (HostSim *)p;
  const int J = h->J;
  h->mode.assign(mode, mode + J); h->bs_max.assign(bs_max, bs_max + J); h->bs_min.assign(bs_min, bs_min + J);
  h->bs_big.assign(bs_big, bs_big + J); h->orig_locked.assign(orig_locked, orig_locked + J);
  h->acc_skip.assign(acc_skip, acc_skip + J); h->pat_off.assign(pat_off, pat_off + J + 1);
  h->pattern.assign(pattern, pattern + pat_off[J]); h->pattern.push_back(0);
  h->lvl_bs.assign(lvl_bs, lvl_bs + (size_t)J * K); h->lvl_thr.assign(lvl_thr, lvl_thr + (size_t)J * K);
  h->T.mode = h->mode.data(); h->T.bs_max = h->bs_max.data(); h->T.bs_min = h->bs_min.data(); h->T.bs_big = h->bs_big.data();
  h->T.orig_locked = h->orig_locked.data(); h->T.acc_skip = h->acc_skip.data(); h->T.pat_off = h->pat_off.data();
  h->T.pattern = h->pattern.data(); h->T.K = K; h->T.lvl_bs = h->lvl_bs.data(); h->T.lvl_thr = h->lvl_thr.data();
}

extern "C" void sim_host_set_worker_types(void *p, int W, const double *thr, const int *cap) {
  HostSim *h = (HostSim *)p;
  h->thr_w.assign(thr, thr + (size_t)h->J * W); h->cap_w.assign(cap, cap + W);
  h->T.W = W; h->T.thr_w = h->thr_w.data(); h->T.cap_w = h->cap_w.data();
}

extern "C" void sim_host_begin(void *p, Scn *out) {
  HostSim *h = (HostSim *)p;
  scenario_begin(h->T, h->X, &h->scn, h->sh);
  *out = h->scn;
}

extern "C" void sim_host_step(void *p, const unsigned char *chosen, int ngpus, double tpi, double grd, Scn *out,
                              unsigned char *status, int *epoch, double *tl_ns, int *tl_end, double *thr_meas) {
  HostSim *h = (HostSim *)p;
  scenario_step(h->T, h->X, &h->scn, chosen, ngpus, tpi, grd, h->sh);
  *out = h->scn;
  for (int j = 0; j < h->J; ++j) {
    status[j] = h->status[j]; epoch[j] = h->epoch[j]; tl_ns[j] = h->tl_ns[j]; tl_end[j] = h->tl_end[j];
    thr_meas[j] = h->thr_meas[j];
  }
}

extern "C" void sim_host_results(void *p, double *jct, long long *steps_run, double *run_time) {
  HostSim *h = (HostSim *)p;
  for (int j = 0; j < h->J; ++j) { jct[j] = h->jct[j]; steps_run[j] = h->steps_run[j]; run_time[j] = h->run_time[j]; }
}

extern "C" void sim_host_job_state(void *p, long long *total, double *thr, int *bs, double *ex, double *fin,
                                   unsigned char *fails, unsigned char *ran) {
  HostSim *h = (HostSim *)p;
  for (int j = 0; j < h->J; ++j) {
    total[j] = h->ctotal[j]; thr[j] = h->cthr[j]; bs[j] = h->cbs[j]; ex[j] = h->last_ex[j]; fin[j] = h->fin[j];
    fails[j] = h->fails[j]; ran[j] = h->ranprev[j];
  }
}

extern "C" void sim_host_destroy(void *p) { delete (HostSim *)p; }
