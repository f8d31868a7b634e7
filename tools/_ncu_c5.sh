CMD="python bench.py --workload c5 --rows 160000 --steps 2 --warmup 2 --no-e2e --no-cpu --no-parity --no-fit"
$CMD > gpurun_out/ncu_c5_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:bmu_cand_tensor -s 2 -c 1 -o gpurun_out/r2_c5_segm $CMD > gpurun_out/ncu_c5_a.log 2>&1
DBGSOM_TC_SEGM=0 $CMD > gpurun_out/ncu_c5_plain0.log 2>&1 && \
DBGSOM_TC_SEGM=0 ncu --set full --clock-control none --import-source on -k regex:bmu_cand_tensor -s 2 -c 1 -o gpurun_out/r2_c5_chain $CMD > gpurun_out/ncu_c5_b.log 2>&1
ls -la gpurun_out/*.ncu-rep
