mkdir -p gpurun_out
for ch in 4 16; do
IEXA_CLASS_CHUNK=$ch python tests/quick_bench.py opf118 10000 2>&1 | tail -1 > gpurun_out/s31_opf118_ch$ch.log
IEXA_CLASS_CHUNK=$ch python tests/quick_bench.py opf30 10000 2>&1 | tail -1 > gpurun_out/s31_opf30_ch$ch.log
done
python -m pytest tests -m gpu -x -q 2>&1 | tail -2 > gpurun_out/s31_pytest.log
