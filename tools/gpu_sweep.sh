mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "graph or same_x" 2>&1 | tail -3 > gpurun_out/s20_pytest.log
for cfg in "quad 16000" "pandemic 100000" "farmer 100000" "opf 100000"; do
  set -- $cfg
  IEXA_GRAPH=1 python tests/quick_bench.py $1 $2 2>&1 | tail -2 > gpurun_out/s20_$1.log
done
