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
  std::vector<long long> total, spe;
  std::vector<int> sf, bs;
  std::vector<unsigned char> status, ranprev;
  std::vector<long long> steps_run, nsteps;
  std::vector<double> run_time, latest, jct, fin, tl_ns, thr_meas;
  std::vector<int> tl_prev, tl_end, epoch;
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
  h->sf.assign(sf, sf + J); h->bs.assign(bs, bs + J);
  h->dur15.resize(J); h->spe.resize(J);
  for (int j = 0; j < J; ++j) {
    h->dur15[j] = (double)(long long)(duration[j] * 1.5);
    h->spe[j] = (dataset_len[j] + bs[j] - 1) / bs[j];
  }
  h->status.resize(J); h->ranprev.resize(J); h->steps_run.resize(J); h->nsteps.resize(J);
  h->run_time.resize(J); h->latest.resize(J); h->jct.resize(J); h->fin.resize(J); h->tl_ns.resize(J); h->thr_meas.resize(J);
  h->tl_prev.resize(J); h->tl_end.resize(J); h->epoch.resize(J);
  h->T = Trace{J, h->arrival.data(), h->total.data(), h->sf.data(), h->thr.data(), h->dur15.data(), h->bs.data(), h->spe.data()};
  h->X = State{h->status.data(), h->ranprev.data(), h->steps_run.data(), h->nsteps.data(), h->run_time.data(),
               h->latest.data(), h->jct.data(), h->fin.data(), h->tl_ns.data(), h->thr_meas.data(), h->tl_prev.data(),
               h->tl_end.data(), h->epoch.data()};
  h->sh = Shared{h->sd, h->si};
  return h;
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

extern "C" void sim_host_destroy(void *p) { delete (HostSim *)p; }
