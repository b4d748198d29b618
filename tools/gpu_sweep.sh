set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -3 > gpurun_out/s8_pytest.log
B="python bench.py --steps 100 --warmup 5 --no-e2e --no-cpu"
$B > gpurun_out/s8_idx32.json 2>&1
IEXA_IDX64=1 $B > gpurun_out/s8_idx64.json 2>&1
IEXA_SCHED=t $B > gpurun_out/s8_idx32_table.json 2>&1
IEXA_HOIST=-1,-1,16,0,16 IEXA_MINBLOCKS=8,8,8,10,8 $B > gpurun_out/s8_idx32_h16c8.json 2>&1
python bench.py --steps 300 --warmup 5 --no-e2e --no-cpu --supports 125000 > gpurun_out/s8_small.json 2>&1
python tests/quick_bench.py opf30 10000 > gpurun_out/s8_opf30.log 2>&1
python tests/quick_bench.py opf 100000 > gpurun_out/s8_opf.log 2>&1
python tests/quick_bench.py pandemic 100000 > gpurun_out/s8_pand.log 2>&1
python tests/quick_bench.py farmer 100000 > gpurun_out/s8_farmer.log 2>&1
