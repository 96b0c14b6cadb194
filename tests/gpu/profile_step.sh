#!/bin/bash
# ncu evidence for one training step (run under gpurun, one GPU): the launch list of a step, a full capture of the
# residual-block convolution and the counters of every row-stream launch.  Usage: tests/gpu/profile_step.sh <tag>
# The step is launched eagerly (SGGAN_GRAPH=0) so that every kernel is its own launch for the profiler.
tag=${1:-r02}
out=gpurun_out
mkdir -p $out
export SGGAN_GRAPH=0
cmd="python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --sustained-seconds 0"
$cmd > $out/${tag}_plain.log 2>&1 || { echo "plain run failed"; tail -5 $out/${tag}_plain.log; exit 1; }
# 1. every launch of the 4th step (3 warm-up steps before it) with its device time
ncu --metrics gpu__time_duration.sum --clock-control none -c 1600 --csv --log-file $out/${tag}_launches_all.csv $cmd \
  > $out/${tag}_ncu1.log 2>&1
# 2. full capture of two residual-block convolutions (launches 5, 6 of conv_gemm_pair_kernel: residual-block convolutions)
ncu --set full --clock-control none --import-source on -k regex:conv_gemm_pair_kernel -s 4 -c 2 -f -o $out/${tag}_conv_res $cmd \
  > $out/${tag}_ncu2.log 2>&1
# 3. counters of every row-stream launch of one step
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum,sm__throughput.avg.pct_of_peak_sustained_elapsed,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed \
  --clock-control none -k regex:"row_stream|act_bwd|fake_grad|prep_image|wgrad_gemm|wgrad_reduce|conv_gemm_swap|conv_gemm_tc|conv_gemm_pair" -s 560 -c 290 --csv \
  --log-file $out/${tag}_glue_counters.csv $cmd > $out/${tag}_ncu3.log 2>&1
ls -la $out | grep $tag
