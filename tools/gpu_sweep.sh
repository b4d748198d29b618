# full single-GPU check: GPU test suite, smoke, bench (both arms)
mkdir -p gpurun_out
( time python -m pytest tests -m gpu -x -q ) > gpurun_out/final_pytest.log 2>&1
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/final_smoke.log 2>&1
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/final_ref.json 2> gpurun_out/final_ref.err
python bench.py --steps 20 --warmup 3 > gpurun_out/final_bench_k20.json 2> gpurun_out/final_bench_k20.err
python bench.py > gpurun_out/final_bench.json 2> gpurun_out/final_bench.err
