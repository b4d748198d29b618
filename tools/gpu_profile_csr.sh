# ncu launch list + full capture of the COO->CSR apply kernels at config 3 (tests/quick_csr.py); run on the GPU box from the repo root
TAG=${1:-r2_csr}
mkdir -p gpurun_out
python tests/quick_csr.py 1000000 > gpurun_out/${TAG}_plain.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/${TAG}_plain.log; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:csr_apply -c 200 --csv --log-file gpurun_out/${TAG}_launches.csv python tests/quick_csr.py 1000000 > gpurun_out/${TAG}_ncu1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:csr_apply -s 3 -c 2 -o gpurun_out/${TAG}_prof -f python tests/quick_csr.py 1000000 > gpurun_out/${TAG}_ncu2.log 2>&1
ncu -i gpurun_out/${TAG}_prof.ncu-rep --page raw --csv > gpurun_out/${TAG}_full_raw.csv 2>/dev/null
ls -la gpurun_out/${TAG}_*
