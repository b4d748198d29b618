mkdir -p gpurun_out
python -m pytest tests/test_csr.py tests/test_full_size.py -m gpu -x -q -k "csr" 2>&1 | tail -3 > gpurun_out/s21_pytest.log
python tests/quick_csr.py > gpurun_out/s21_csr.log 2>&1
