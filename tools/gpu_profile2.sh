# usage: tools/gpu_profile2.sh <tag> <quick_bench args...>     (run on the GPU box, from the repo root)
# 1. plain run (must exit 0 without ncu);  2. launch list (gpu__time_duration);  3. one --set full capture of the callback /
# product / csr kernels (first launches after the warm-up), raw page exported as CSV into gpurun_out/
TAG=$1; shift
mkdir -p gpurun_out
export IEXA_DUMP_DIR=$PWD/gpurun_out
python tests/quick_bench.py "$@" > gpurun_out/${TAG}_plain.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/${TAG}_plain.log; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file gpurun_out/${TAG}_launches.csv python tests/quick_bench.py "$@" > gpurun_out/${TAG}_ncu1.log 2>&1
IEXA_QB_REPS=1 ncu --set full --clock-control none --import-source on -k regex:'iexa_cb|csr_apply' -s ${SKIP:-0} -c ${COUNT:-40} -o gpurun_out/${TAG}_prof -f python tests/quick_bench.py "$@" > gpurun_out/${TAG}_ncu2.log 2>&1
ncu -i gpurun_out/${TAG}_prof.ncu-rep --page raw --csv > gpurun_out/${TAG}_full_raw.csv 2>/dev/null
rm -f gpurun_out/${TAG}_prof.ncu-rep.tmp
ls -la gpurun_out/${TAG}_*
