#!/bin/bash
# A/B of compile-time variants of the streaming quantiser against the default build (profiles/r02_ab_quant_variants.log).
# Build the variant libraries first (only bfp_quant.cu differs; the other objects are the default build's):
#   cd quantization-sparsity-interplay_b200 && mkdir -p variants && for v in "R:-DBFP_REDUX_MAX" "RP:-DBFP_REDUX_MAX -DBFP_STOC_PINGPONG" "P:-DBFP_STOC_PINGPONG"; do
#     n=${v%%:*}; nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -fmad=false -Xcompiler -fPIC ${v#*:} -c csrc/bfp_quant.cu -o variants/bfp_quant_$n.o
#     nvcc -gencode arch=compute_100a,code=sm_100a -shared -o variants/libbfp_$n.so variants/bfp_quant_$n.o $(ls csrc/*.o | grep -v bfp_quant.o) -lcudart_static -lpthread -ldl -lrt; done
# (variants/ is git-ignored but travels to the GPU box).  Each variant runs the throughput table and its own parity suite.
for n in base R RP P; do
  if [ $n = base ]; then unset BFP_B200_LIB; else export BFP_B200_LIB=$PWD/quantization-sparsity-interplay_b200/variants/libbfp_$n.so; fi
  echo "=== $n"
  python tools/tune_quant.py --quick --dtypes f32,bf16,f16 --iters 20 2>&1 | grep -E "stoc|near" | grep -E "sq"
  python -m pytest tests/test_quant_gpu.py -x -q 2>&1 | tail -1
done
