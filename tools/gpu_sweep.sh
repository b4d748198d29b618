set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -3 > gpurun_out/s10_pytest.log
python bench.py > gpurun_out/s10_bench.json 2> gpurun_out/s10_bench.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/s10_ref.json 2> gpurun_out/s10_ref.err
