python tools/check_backends.py 60000 512 1024 3 > gpurun_out/tb_check_512.txt 2>&1; tail -2 gpurun_out/tb_check_512.txt
python tools/check_backends.py 40000 2048 4096 6 > gpurun_out/tb_check_2048.txt 2>&1; tail -3 gpurun_out/tb_check_2048.txt
DBGSOM_TC_SEGM=0 python tools/check_backends.py 40000 2048 4096 6 > gpurun_out/tb_check_2048_chain.txt 2>&1; tail -3 gpurun_out/tb_check_2048_chain.txt
python tools/check_backends.py 150000 4096 1024 1 > gpurun_out/tb_check_4096.txt 2>&1; tail -2 gpurun_out/tb_check_4096.txt
for sg in 1 0; do
  DBGSOM_TC_SEGM=$sg python bench.py --workload c5 --rows 625000 --steps 4 --warmup 3 --no-e2e --no-cpu --no-parity --no-fit > gpurun_out/c5tb_1gpu_segm$sg.json 2>> gpurun_out/c5tb.err
done
