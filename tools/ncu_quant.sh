#!/bin/bash
# ncu metrics of the streaming quantiser for a list of configurations (one GPU).  usage: tools/ncu_quant.sh <tag> [full]
# Each configuration: python tools/prof_quant.py <dtype> 7 64 sq 4096 11008 [stoc]; the 5th matching launch is captured.
tag=${1:-r02}; mode=${2:-quick}
M="gpu__time_duration.sum,smsp__inst_executed.sum,launch__registers_per_thread,launch__occupancy_limit_registers,sm__warps_active.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active,sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active,dram__throughput.avg.pct_of_peak_sustained_elapsed,dram__bytes_read.sum,dram__bytes_write.sum,smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio,smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio,smsp__average_warps_issue_stalled_wait_per_issue_active.ratio,smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio,smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio,smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio,smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio,smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio,smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio"
mkdir -p gpurun_out
for cfg in "f32 nearest" "f32 stoc" "bf16 nearest" "f16 nearest" "bf16 stoc"; do
  set -- $cfg; dt=$1; r=$2; extra=""; [ "$r" = "stoc" ] && extra="stoc"
  out=gpurun_out/${tag}_quant_${dt}_${r}
  if [ "$mode" = "full" ]; then
    ncu --set full --clock-control none --import-source on -k regex:quant_stream_kernel -s 4 -c 1 -o $out -f python tools/prof_quant.py $dt 7 64 sq 4096 11008 $extra > $out.log 2>&1
    ncu -i $out.ncu-rep --page raw --csv > $out.ncu_raw.csv 2>/dev/null
  else
    ncu --metrics $M --clock-control none -k regex:quant_stream_kernel -s 4 -c 1 --csv python tools/prof_quant.py $dt 7 64 sq 4096 11008 $extra > $out.quick.csv 2>&1
  fi
done
