# one shard of an 8-rank sharding of config 3, timed on ONE GPU (rank 4 of 8): which launch shape suits < 1-wave kernels
mkdir -p gpurun_out
: > gpurun_out/shard_knobs.log
export IEXA_QB_WORLD=8 IEXA_PROD=0 IEXA_QB_REPS=200
run() { echo "== $*" >> gpurun_out/shard_knobs.log; env "$@" python tests/quick_bench.py quad 1000000 2>&1 | grep -E "^(cons|jac|hess|step|eval3)" | cut -c1-150 >> gpurun_out/shard_knobs.log; }
run IEXA_X=0
run IEXA_BLOCK=64
run IEXA_BLOCK=256
run IEXA_PDL=0
run IEXA_BLOCK=64 IEXA_MINBLOCKS=16,16,16,16,16
run IEXA_BLOCK=96
cat gpurun_out/shard_knobs.log
