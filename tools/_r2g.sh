python -m pytest tests -m gpu -x -q > gpurun_out/r2g_pytest.log 2>&1; tail -3 gpurun_out/r2g_pytest.log
DBGSOM_PROFILE=1 python tools/fit_config5.py --n 100000 --manifold --aligned --n-iter 200 --json gpurun_out/fit_c5_1gpu_100k.json > gpurun_out/fit_c5_1gpu_100k.log 2>&1; tail -3 gpurun_out/fit_c5_1gpu_100k.log
