for n in base R RP P; do
  if [ $n = base ]; then unset BFP_B200_LIB; else export BFP_B200_LIB=$PWD/quantization-sparsity-interplay_b200/variants/libbfp_$n.so; fi
  echo "=== $n"
  python tools/tune_quant.py --quick --dtypes f32,bf16,f16 --iters 20 2>&1 | grep -E "stoc|near" | grep -E "sq" 
  python -m pytest tests/test_quant_gpu.py -x -q 2>&1 | tail -1
done
