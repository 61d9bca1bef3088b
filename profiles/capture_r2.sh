#!/bin/bash
# Round-2 evidence capture (run under gpurun on ONE B200): plain bench first (must exit 0), then the ncu launch list of
# the same command, `ncu --set full` of the four GEMM launches of one layer and of one attention launch, and the DRAM
# counters of the HBM-bound kernels.  Outputs land in gpurun_out/; summaries are copied into profiles/ by hand.
# A number printed by a run under ncu is never a bench value.
set -x
TAG=${1:-r2}
CMD="python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --no-agreement"
$CMD > gpurun_out/plain_$TAG.log 2>&1 || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/launches_$TAG.csv $CMD > gpurun_out/ncu_launch_$TAG.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:gemm_tc_pair -s 50 -c 4 -f -o gpurun_out/gemm_$TAG $CMD > gpurun_out/ncu_gemm_$TAG.log 2>&1
ncu -i gpurun_out/gemm_$TAG.ncu-rep --page raw --csv > gpurun_out/gemm_raw_$TAG.csv 2>/dev/null
ncu --set full --clock-control none --import-source on -k regex:attention_kernel -s 12 -c 1 -f -o gpurun_out/att_$TAG $CMD > gpurun_out/ncu_att_$TAG.log 2>&1
ncu -i gpurun_out/att_$TAG.ncu-rep --page raw --csv > gpurun_out/att_raw_$TAG.csv 2>/dev/null
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,dram__throughput.avg.pct_of_peak_sustained_elapsed,lts__t_sector_hit_rate.pct --clock-control none -k regex:'ln_rows_vec|text_embed_vec|bias_build|exit_fused|visual_ln|im2col|ragged_gather|plan_rows|keymask' -s 45 -c 50 --csv --log-file gpurun_out/hbm_$TAG.csv $CMD > gpurun_out/ncu_hbm_$TAG.log 2>&1
echo done
