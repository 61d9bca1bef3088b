#!/bin/bash
# Round-1 evidence capture (run under gpurun): tests, bench, reference arm, ncu launch list, ncu --set full of the
# GEMM and attention kernels.  Outputs land in gpurun_out/; summaries are copied into profiles/ by hand.
set -x
TAG=${1:-r1}
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_$TAG.log 2>&1
python bench.py > gpurun_out/bench_$TAG.log 2>&1
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref_$TAG.log 2>&1
CMD="python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e"
$CMD > gpurun_out/plain_$TAG.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/launches_$TAG.csv $CMD > gpurun_out/ncu_launch_$TAG.log 2>&1
$CMD > gpurun_out/plain_$TAG.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:gemm_tc_pair -s 50 -c 4 -f -o gpurun_out/gemm_$TAG $CMD > gpurun_out/ncu_gemm_$TAG.log 2>&1
$CMD > gpurun_out/plain_$TAG.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:attention_kernel -s 12 -c 1 -f -o gpurun_out/att_$TAG $CMD > gpurun_out/ncu_att_$TAG.log 2>&1
$CMD > gpurun_out/plain_$TAG.log 2>&1 &&
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,dram__throughput.avg.pct_of_peak_sustained_elapsed,lts__t_sector_hit_rate.pct --clock-control none -k regex:'ln_rows_vec|text_embed_vec|bias_build|exit_fused|visual_ln|im2col' -s 42 -c 44 --csv --log-file gpurun_out/hbm_$TAG.csv $CMD > gpurun_out/ncu_hbm_$TAG.log 2>&1
echo done
