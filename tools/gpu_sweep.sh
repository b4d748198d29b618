mkdir -p gpurun_out
( time python -m pytest tests/test_fuzz.py -m gpu -x -q ) > gpurun_out/s25_fuzz.log 2>&1
