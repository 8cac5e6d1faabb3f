# A/B of the occupancy knobs of the stochastic / half-precision quantiser (profiles/r02_ab_quant_occupancy.log); variant libraries are built like
# tools/ab_quant_variants.sh describes, with -DBFP_STOC_MIN_CTAS=2|4 and -DBFP_HALF_MIN_CTAS=5.
for n in base C2 C4 H5; do
  if [ $n = base ]; then unset BFP_B200_LIB; else export BFP_B200_LIB=$PWD/quantization-sparsity-interplay_b200/variants/libbfp_$n.so; fi
  echo "=== $n"
  python tools/tune_quant.py --quick --dtypes f32,bf16,f16 --iters 20 2>&1 | grep -E "stoc|near" | grep -E "sq" | grep -E "stoc|bf16|f16"
done
