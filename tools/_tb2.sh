for sc in 2 4; do
  DBGSOM_TC_SEGM=0 DBGSOM_ACC_SCALE=$sc python bench.py --workload c5 --rows 625000 --steps 4 --warmup 3 --no-e2e --no-cpu --no-parity --no-fit > gpurun_out/c5tb_1gpu_chain_acc$sc.json 2>> gpurun_out/c5tb2.err
done
DBGSOM_TC_SEGM=0 DBGSOM_ACC_SCALE=4 python tools/check_backends.py 40000 2048 4096 6 > gpurun_out/tb_check_2048_chain_acc4.txt 2>&1; tail -3 gpurun_out/tb_check_2048_chain_acc4.txt
