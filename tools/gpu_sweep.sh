mkdir -p gpurun_out
( time python -m pytest tests/test_full_size.py -m gpu -x -q --durations=10 ) > gpurun_out/s15_fullsize.log 2>&1
free -g | head -2 >> gpurun_out/s15_fullsize.log; nproc >> gpurun_out/s15_fullsize.log
