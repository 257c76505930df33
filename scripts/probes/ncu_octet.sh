OUT=gpurun_out
ncu --set full --clock-control none --import-source on -k regex:'eval_octet_kernel' -s 50 -c 10 -o $OUT/oct -f python scripts/profile_compute_model.py --levels 8 > /dev/null 2>&1
ncu -i $OUT/oct.ncu-rep --page raw --csv 2>/dev/null | python3 -c "
import csv, sys, re
rows = list(csv.reader(sys.stdin))
hdr, units = rows[0], rows[1]
pat = re.compile(r'gpu__time_duration.sum|fp64_cycles_active.avg.pct_of_peak_sustained_active|issue_active.avg.pct|registers_per_thread\$|pcsamp_warps_issue_stalled_[a-z_]*\$|smsp__inst_executed.sum\$|inst_executed_pipe_fp64|inst_executed_pipe_lsu|inst_executed_pipe_xu|inst_executed_pipe_alu|inst_executed_pipe_fma|l1tex__data_bank_conflicts|smsp__warps_eligible|warps_active.avg.pct')
best = max(rows[2:], key=lambda r: float(dict(zip(hdr, r))['gpu__time_duration.sum'].replace(',','')))
d = dict(zip(hdr, best))
print('kernel', d['Kernel Name'][:80], d['Grid Size'])
for h, u, v in zip(hdr, units, best):
    if pat.search(h) and not h.endswith('_not_issued'): print(f'   {h:80s} {v:>16s} {u}')
" > $OUT/r2_eval_octet_kernel_ncu_full_cfg4_level7.txt
ncu -i $OUT/oct.ncu-rep --page source --csv --print-source sass 2>/dev/null | head -5 > $OUT/oct_source_head.txt
rm -f $OUT/oct.ncu-rep
cat $OUT/r2_eval_octet_kernel_ncu_full_cfg4_level7.txt
