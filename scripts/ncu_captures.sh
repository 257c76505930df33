#!/bin/bash
# One `ncu --set full` capture per kernel this round changed or leans on, summarised as text under gpurun_out/ (copied to
# profiles/ by hand).  Run on the GPU box: gpurun --timeout 1500 -- bash scripts/ncu_captures.sh
set -u
OUT=gpurun_out
mkdir -p $OUT
METRICS='^gpu__time_duration.sum$|^dram__bytes_read.sum$|^dram__bytes_write.sum$|^gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed$|^sm__throughput.avg.pct_of_peak_sustained_elapsed$|^sm__inst_executed_pipe_fp64.sum$|^sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active$|^sm__inst_executed_pipe_tensor|^smsp__inst_executed.sum$|^launch__registers_per_thread$|^sm__warps_active.avg.pct_of_peak_sustained_active$|^smsp__issue_active.avg.pct_of_peak_sustained_active$|^lts__t_bytes.sum$|^launch__occupancy_limit|^smsp__pcsamp_warps_issue_stalled_(math_pipe_throttle|wait|long_scoreboard|short_scoreboard|barrier|lg_throttle|mio_throttle|selected|not_selected)$|^sm__cycles_elapsed.avg.per_second$'
summ() {  # $1 report, $2 text file
  ncu -i $1 --page raw --csv 2>/dev/null | python3 -c "
import csv, sys, re
rows = list(csv.reader(sys.stdin))
if len(rows) < 3: sys.exit(0)
hdr, units = rows[0], rows[1]
pat = re.compile(r'''$METRICS''')
for r in rows[2:]:
    d = dict(zip(hdr, r))
    print('== kernel:', d.get('Kernel Name','?')[:110], '| grid', d.get('Grid Size'), '| block', d.get('Block Size'))
    for h, u, v in zip(hdr, units, r):
        if pat.search(h): print(f'   {h:75s} {v:>18s} {u}')
" > $2
}
WHICH="${1:-1 2 3 4}"
if [[ " $WHICH " == *" 1 "* ]]; then
# 1. dominant kernel on the bench workload (512^3, one launch)
ncu --set full --clock-control none --import-source on -k regex:eval_zrun -s 3 -c 1 -o $OUT/r2_eval_zrun_512 -f python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu --no-extras > /dev/null 2>&1
summ $OUT/r2_eval_zrun_512.ncu-rep $OUT/r2_eval_zrun_kernel_ncu_full_512.txt
fi
if [[ " $WHICH " == *" 2 "* ]]; then
# 2. the symmetric solve at n = 6999: biggest trailing update, diagonal factor, panel solve, backward substitution
ncu --set full --clock-control none --import-source on -k regex:'syrk_kernel|potrf64|trsm_dmma|trsv_lt' -s 0 -c 12 -o $OUT/r2_chol_6999 -f python scripts/probes/prof_solve.py 1000 > /dev/null 2>&1
summ $OUT/r2_chol_6999.ncu-rep $OUT/r2_chol_kernels_ncu_full_n6999.txt
fi
if [[ " $WHICH " == *" 3 "* ]]; then
# 3. covariance assembly at n = 6999
ncu --set full --clock-control none --import-source on -k regex:'cov_ii|cov_ig|cov_gg|cov_du' -s 0 -c 4 -o $OUT/r2_cov_6999 -f python scripts/probes/prof_solve.py 1000 > /dev/null 2>&1
summ $OUT/r2_cov_6999.ncu-rep $OUT/r2_cov_kernels_ncu_full_n6999.txt
fi
if [[ " $WHICH " == *" 4 "* ]]; then
# 4. point-list evaluation with the fused activator on the multi-fault octree model: the largest launches of the last level
ncu --set full --clock-control none --import-source on -k regex:'eval_octet_kernel|eval_kernel' -s 330 -c 24 -o $OUT/r2_eval_points_cfg4 -f python scripts/profile_compute_model.py --levels 8 > /dev/null 2>&1
summ $OUT/r2_eval_points_cfg4.ncu-rep $OUT/r2_eval_points_kernel_ncu_full_cfg4.txt
fi
rm -f $OUT/*.ncu-rep
ls -la $OUT/*.txt
