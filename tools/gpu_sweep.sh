mkdir -p gpurun_out
bash tools/gpu_profile.sh v8
