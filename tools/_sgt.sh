timeout 200 python tools/check_backends.py 40000 2048 4096 6 > gpurun_out/sgt3c_check_2048.txt 2>&1; echo "rc=$?"; tail -2 gpurun_out/sgt3c_check_2048.txt
timeout 200 python tools/check_backends.py 60000 512 1024 3 > gpurun_out/sgt3c_check_512.txt 2>&1; echo "rc=$?"; tail -1 gpurun_out/sgt3c_check_512.txt
timeout 200 python bench.py --workload c5 --rows 625000 --steps 4 --warmup 3 --no-e2e --no-cpu --no-parity --no-fit > gpurun_out/c5sgt3c_1gpu.json 2>> gpurun_out/c5sgt.err; echo "rc=$?"
