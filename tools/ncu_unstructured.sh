#!/bin/bash
# ncu capture of the fused unstructured pipeline's four kernels (one GPU).  usage: tools/ncu_unstructured.sh <tag> <dtype> <order> <rows> <cols> [relu]
tag=$1; shift
out=gpurun_out/${tag}_unstructured_$1_$2_$3x$4$5
ncu --set full --clock-control none --import-source on -k regex:"sample_kernel|pass_a_kernel|refine_kernel|apply_kernel" -s 8 -c 4 -o $out -f python tools/prof_unstructured_fused.py "$@" > $out.log 2>&1
ncu -i $out.ncu-rep --page raw --csv > $out.ncu_raw.csv 2>/dev/null
ncu -i $out.ncu-rep --page source --csv --kernel-name regex:pass_a_kernel > $out.pass_a_source.csv 2>/dev/null
ncu -i $out.ncu-rep --page source --csv --kernel-name regex:sample_kernel > $out.sample_source.csv 2>/dev/null
ncu -i $out.ncu-rep --page source --csv --kernel-name regex:refine_kernel > $out.refine_source.csv 2>/dev/null
rm -f $out.ncu-rep
