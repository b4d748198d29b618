mkdir -p gpurun_out
for u in 4; do
( time IEXA_CLASS_UNROLL=$u python tests/quick_bench.py opf118 10000 ) 2>&1 | tail -5 > gpurun_out/s33_opf118_u$u.log
IEXA_CLASS_UNROLL=$u python tests/quick_bench.py opf30 10000 2>&1 | tail -1 > gpurun_out/s33_opf30_u$u.log
done
IEXA_CLASS_UNROLL=2 IEXA_CLASS_CHUNK=16 python tests/quick_bench.py opf118 10000 2>&1 | tail -1 > gpurun_out/s33_opf118_u2c16.log
