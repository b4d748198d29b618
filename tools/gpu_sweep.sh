set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -3 > gpurun_out/s9_pytest.log
B="python bench.py --steps 100 --warmup 5 --no-e2e --no-cpu"
$B > gpurun_out/s9_sincos.json 2>&1
IEXA_NO_SINCOS=1 $B > gpurun_out/s9_nosincos.json 2>&1
python bench.py --steps 300 --warmup 5 --no-e2e --no-cpu --supports 125000 > gpurun_out/s9_small.json 2>&1
python tests/quick_bench.py opf 100000 > gpurun_out/s9_opf.log 2>&1
