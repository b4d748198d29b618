mkdir -p gpurun_out
( time python -m pytest tests -m gpu -x -q ) > gpurun_out/s22_pytest.log 2>&1
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/s22_smoke.log 2>&1
python bench.py > gpurun_out/s22_bench.json 2> gpurun_out/s22_bench.err
