# parity of the non-default code-generation modes (each is a documented knob): GPU parity + fuzz tests per variant
mkdir -p gpurun_out
: > gpurun_out/knobs.log
run() { echo "== $*" >> gpurun_out/knobs.log; env "$@" python -m pytest tests/test_gpu_parity.py tests/test_fuzz.py tests/test_edge_cases.py -m gpu -x -q -p no:cacheprovider 2>&1 | tail -2 >> gpurun_out/knobs.log; }
run IEXA_SCHED=t
run IEXA_PDL=0
run IEXA_IDX32=1,1,1,1,1
run IEXA_IDX32=0,0,0,0,0 IEXA_NO_AFFINE=1 IEXA_NO_SINCOS=1
run IEXA_HOIST=0
run IEXA_HOIST=-1 IEXA_PREFETCH=1
run IEXA_BLOCK=64
run IEXA_BLOCK=256
run IEXA_STAGE=block
run IEXA_STAGE=tma
run IEXA_STAGE=vec2
run IEXA_CLASS_CHUNK=1 IEXA_CLASS_UNROLL=1
run IEXA_CLASS_CHUNK=3 IEXA_CLASS_UNROLL=4
# round 2 knobs
run IEXA_CLASS_SMEM=0
run IEXA_NO_CLASS_FUSION=1
run IEXA_PAD_WARP_TILES=1
run IEXA_NO_RIDERS=1 IEXA_NO_PAIR_REDUCE=1
run IEXA_NO_SCATTER_DIRECT=1
run IEXA_MINBLOCKS_PROD=8,8,8 IEXA_HOIST_PROD=0,0,0
run IEXA_ORDER=g
run IEXA_CACHE_DIR=off
run IEXA_NO_VMM=1
run IEXA_NO_COLUMN_SLICES=1 IEXA_NO_PFUNC_SHARDING=1
