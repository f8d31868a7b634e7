TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
timeout 300 $TR --master-port 29751 bench.py --gpus 8 --workload c5 --steps 4 --warmup 3 --no-e2e --no-cpu --no-fit > gpurun_out/r2b_c5_8gpu.json 2> gpurun_out/r2b_c5_8gpu.err
tail -c 300 gpurun_out/r2b_c5_8gpu.json
DBGSOM_PROFILE=1 timeout 500 $TR --master-port 29752 tools/fit_config5.py --distributed --rows 625000 --manifold --aligned --n-iter 320 --coarse-frac 0.9 --json gpurun_out/r2b_fit_c5_8gpu.json > gpurun_out/r2b_fit_c5_8gpu.log 2>&1
tail -c 2500 gpurun_out/r2b_fit_c5_8gpu.log | tr '\r' '\n' | tail -4
