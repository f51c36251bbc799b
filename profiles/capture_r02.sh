#!/bin/bash
# ncu evidence of round 2 (run under gpurun on ONE GPU).  Launch list of the bench command + one `--set full` capture
# per kernel of the hot path; summaries: python profiles/summarize.py r02 (here, without a GPU).
set -x
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'swb::' -c 4000 --csv --log-file gpurun_out/launches_r02.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench_under_ncu.log 2>&1
cap() {  # name regex driver-section [launches to skip]
  ncu --set full --clock-control none --import-source on -k "regex:$2" -s ${4:-2} -c 1 -f -o gpurun_out/$1_r02 \
      python profiles/run_kernels.py $3 > gpurun_out/ncu_$1.log 2>&1
}
cap solve_cluster 'solve_kernel' single
cap solve_batched 'solve_kernel' batched
cap place_single 'place_kernel' single
cap place_single_fb 'place_kernel' single 6      # the fallback instance (priority sweep) of the same section
cap place_batched 'place_kernel' batched
cap gbm 'gbm_kernel' gbm
cap market 'market_step_fast' market
cap gavel 'gavel_round_kernel' gavel
cap place_rerank 'place_kernel' canonical
ls -la gpurun_out/*_r02.ncu-rep
