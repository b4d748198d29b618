set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -3 > gpurun_out/s2_pytest.log
B="python bench.py --steps 100 --warmup 5 --no-e2e --no-cpu"
IEXA_PDL=0 $B > gpurun_out/s2_pdl0.json 2>gpurun_out/s2_pdl0.err
IEXA_PDL=1 $B > gpurun_out/s2_pdl1.json 2>gpurun_out/s2_pdl1.err
IEXA_BLOCK=256 $B > gpurun_out/s2_b256.json 2>gpurun_out/s2_b256.err
IEXA_MINBLOCKS=8,8,12,12,8 $B > gpurun_out/s2_mb12.json 2>gpurun_out/s2_mb12.err
IEXA_MINBLOCKS=8,8,16,12,8 $B > gpurun_out/s2_mb16.json 2>gpurun_out/s2_mb16.err
for p in 0 1; do IEXA_PDL=$p python tests/quick_bench.py pandemic 100000 > gpurun_out/s2_pand_pdl$p.log 2>&1; done
for p in 0 1; do IEXA_PDL=$p python bench.py --steps 200 --warmup 5 --no-e2e --no-cpu --supports 125000 > gpurun_out/s2_small_pdl$p.json 2>&1; done
