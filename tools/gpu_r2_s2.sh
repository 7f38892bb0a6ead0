#!/bin/bash
OUT=gpurun_out/r02s2; mkdir -p $OUT
timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv3x3_slab -s 6 -c 1 -o $OUT/prof_slab python tools/run_once.py trunk 128 2 > $OUT/ncu.log 2>&1
ncu -i $OUT/prof_slab.ncu-rep --page raw --csv > $OUT/prof_slab_raw.csv 2>/dev/null
ncu -i $OUT/prof_slab.ncu-rep --page source --csv > $OUT/prof_slab_source.csv 2>/dev/null
rm -f $OUT/prof_slab.ncu-rep
python - <<'PY'
import csv
rows=list(csv.reader(l for l in open("gpurun_out/r02s2/prof_slab_raw.csv") if not l.startswith("==")))
hdr=rows[0]; r=rows[2]
for k in ["gpu__time_duration.sum","sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active","l1tex__data_pipe_tc_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed","l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed","smsp__issue_active.avg.pct_of_peak_sustained_active","l1tex__m_xbar2l1tex_read_bytes.sum","dram__bytes_read.sum","lts__throughput.avg.pct_of_peak_sustained_elapsed","sm__inst_executed_pipe_tc.sum","smsp__inst_executed.sum"]:
    if k in hdr: print(k, r[hdr.index(k)], rows[1][hdr.index(k)])
PY
python tools/stall_top.py $OUT/prof_slab_source.csv 14 | tail -16
