TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
DBGSOM_PROFILE=1 timeout 700 $TR --master-port 29752 tools/fit_config5.py --distributed --rows 625000 --manifold --aligned --n-iter 400 --json gpurun_out/r2_fit_c5_8gpu.json > gpurun_out/r2_fit_c5_8gpu.log 2>&1
tail -c 2500 gpurun_out/r2_fit_c5_8gpu.log | tr '\r' '\n' | tail -5
