mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "device_objective" 2>&1 | tail -8 > gpurun_out/s34_pytest.log
