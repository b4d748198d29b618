set -x
mkdir -p gpurun_out
export IEXA_DUMP_DIR=$PWD/gpurun_out
TAG=${1:-v6}
python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu > gpurun_out/${TAG}_plain.log 2>&1 || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_launches.csv python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu > gpurun_out/${TAG}_ncu1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:iexa_cb -s 9 -c 3 -o gpurun_out/${TAG}_prof -f python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu > gpurun_out/${TAG}_ncu2.log 2>&1
