for sg in 1 0; do
  DBGSOM_TC_SEGM=$sg python bench.py --workload c5 --rows 625000 --steps 4 --warmup 3 --no-e2e --no-cpu --no-parity --no-fit > gpurun_out/c5_1gpu_segm$sg.json 2>> gpurun_out/c5ab.err
done
python -m pytest tests -m gpu -x -q > gpurun_out/r2f_pytest.log 2>&1; tail -3 gpurun_out/r2f_pytest.log
