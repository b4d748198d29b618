mkdir -p gpurun_out
( time python -m pytest tests/test_full_size.py tests/test_sympy_golden.py -m gpu -x -q --durations=5 ) > gpurun_out/s23_pytest.log 2>&1
