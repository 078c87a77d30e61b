#!/bin/bash
# ncu captures behind profiles/r2_*: run on the GPU box (one GPU); reports land in gpurun_out/
set -x
O=gpurun_out
rm -f $O/r2_*.ncu-rep
# launch list of the bench command (cold-cache, serialised: shares, not absolutes)
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r2_launches_bench.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-other-modes > $O/r2_bench_under_ncu.log 2>&1
# headline kernel, one launch of the bench workload
timeout 600 ncu --set full --import-source on --clock-control none -k regex:k_gap_global_blk -c 1 -o $O/r2_m2 -f \
    python tools/profile_run.py --reads 10000 --graph-bp 100000 > $O/r2_m2.log 2>&1
# pathwise: C3 (-m 5) and C4 (-m 9) at bench size
timeout 600 ncu --set full --import-source on --clock-control none -k regex:k_pathwise_tr -c 1 -o $O/r2_pw_m5 -f \
    python tools/pw_run.py c3 5 10000 > $O/r2_pw_m5.log 2>&1
timeout 600 ncu --set full --clock-control none -k regex:k_pathwise_tr -c 1 -o $O/r2_pw_m9 -f \
    python tools/pw_run.py c4 9 10000 > $O/r2_pw_m9.log 2>&1
# modes 0 / 3 (k_poa_lin) on the headline graph
timeout 300 ncu --set full --clock-control none -k regex:k_poa_lin -c 1 -o $O/r2_lin_m0 -f \
    python tools/profile_run.py --mode 0 --reads 2368 --graph-bp 100000 > $O/r2_lin_m0.log 2>&1
ls -la $O/*.ncu-rep
tail -2 $O/r2_m2.log $O/r2_pw_m5.log $O/r2_pw_m9.log $O/r2_lin_m0.log
