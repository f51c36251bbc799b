"""CPU ORACLE backend (test infrastructure) for shockwave_b200/placement.py: the same call signature as the device
backend, answered by oracle/gavel_round.py — lets the UNMODIFIED reference simulator run with the restated
priority / selection / assignment step on a machine without a GPU, which is how that restatement is pinned on the
golden pickles (tests/test_oracle_gavel_round.py).  Only tests may import this."""
import numpy as np

from oracle import gavel_round as gr


class OracleBackend:
    def __init__(self, record=None):
        self.record = record

    def gavel_round(self, alloc, job_time, worker_time, thr, deficit, sf, capacity, type_order, worker_lists, prev,
                    isolated_plus=False, fifo=False):
        servers = {int(t): [list(worker_lists[ti])] for ti, t in enumerate(type_order)}
        prio, sel, asg = gr.gavel_round(alloc, job_time, worker_time, thr, deficit, sf, capacity,
                                        [int(t) for t in type_order], servers, prev, isolated_plus=isolated_plus,
                                        fifo=fifo)
        if self.record is not None:
            self.record.append(dict(alloc=np.array(alloc), job_time=np.array(job_time), worker_time=np.array(worker_time),
                                    thr=np.array(thr), deficit=np.array(deficit), sf=np.array(sf),
                                    capacity=np.array(capacity), type_order=[int(t) for t in type_order],
                                    worker_lists=[list(w) for w in worker_lists], prev=dict(prev),
                                    isolated_plus=isolated_plus, fifo=fifo,
                                    prio=prio, sel={k: list(v) for k, v in sel.items()},
                                    asg=[(j, tuple(w)) for j, w in asg.items()]))
        return prio, sel, list(asg.items())
