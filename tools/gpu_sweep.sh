mkdir -p gpurun_out
B="python bench.py --steps 100 --warmup 5 --no-e2e --no-cpu"
$B > gpurun_out/s18_base.json 2>&1
IEXA_PREFETCH=1 IEXA_HOIST=-1,-1,8,0,16 IEXA_MINBLOCKS=8,8,10,10,8 $B > gpurun_out/s18_pf_c8h.json 2>&1
IEXA_PREFETCH=1 IEXA_HOIST=-1,-1,4,0,8 IEXA_MINBLOCKS=8,8,10,10,8 $B > gpurun_out/s18_pf_c4_h8.json 2>&1
IEXA_PREFETCH=1 IEXA_HOIST=-1,-1,0,0,0 IEXA_MINBLOCKS=8,8,10,10,8 $B > gpurun_out/s18_pf_all0.json 2>&1
IEXA_PREFETCH=1 IEXA_HOIST=-1,-1,8,4,16 IEXA_MINBLOCKS=8,8,12,10,8 $B > gpurun_out/s18_pf_c12.json 2>&1
IEXA_PREFETCH=1 IEXA_HOIST=-1,-1,8,0,16 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "device_buffers" 2>&1 | tail -2 > gpurun_out/s18_pytest.log
