mkdir -p gpurun_out
for cfg in "quad 16000" "pandemic 100000" "pandemic128 10000" "opf 100000" "opf30 10000" "farmer 100000"; do
  set -- $cfg
  python tests/quick_bench.py $1 $2 > gpurun_out/s14_$1_$2.log 2>&1
done
