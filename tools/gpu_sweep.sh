set -x
mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py tests/test_sympy_golden.py tests/test_edge_cases.py -m gpu -x -q 2>&1 | tail -3 > gpurun_out/s6_pytest.log
B="python bench.py --steps 100 --warmup 5 --no-e2e --no-cpu"
$B > gpurun_out/s6_addr.json 2>&1
