mkdir -p gpurun_out
./tools/stream_probe > gpurun_out/s28_probe.log 2>&1
