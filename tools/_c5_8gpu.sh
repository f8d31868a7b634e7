TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
timeout 400 $TR --master-port 29751 bench.py --gpus 8 --workload c5 --steps 4 --warmup 3 --no-e2e --no-cpu --no-fit > gpurun_out/r2_c5_8gpu.json 2> gpurun_out/r2_c5_8gpu.err
tail -c 600 gpurun_out/r2_c5_8gpu.json
DBGSOM_PROFILE=1 timeout 700 $TR --master-port 29752 tools/fit_config5.py --distributed --n 625000 --manifold --aligned --n-iter 400 --json gpurun_out/r2_fit_c5_8gpu.json > gpurun_out/r2_fit_c5_8gpu.log 2>&1
tail -c 1500 gpurun_out/r2_fit_c5_8gpu.log | tr '\r' '\n' | tail -4
nvidia-smi topo -m > gpurun_out/r2_topo8.txt 2>&1
