set -x
for xu in 0 1; do
  DBGSOM_ACC_XU2=$xu python tools/bench_k2.py 10000000 256 4096 >> gpurun_out/k2ab.jsonl 2>&1
  DBGSOM_ACC_XU2=$xu python tools/bench_k2.py 12500000 128 4096 >> gpurun_out/k2ab.jsonl 2>&1
done
DBGSOM_ACC_XU2=0 python tools/bench_k2.py 10000000 256 4096 500 >> gpurun_out/k2ab.jsonl 2>&1
DBGSOM_ACC_XU2=0 python tools/bench_k2.py 2000000 784 400 >> gpurun_out/k2ab.jsonl 2>&1
DBGSOM_ACC_XU2=0 python tools/bench_k2.py 300000 4096 16384 >> gpurun_out/k2ab.jsonl 2>&1
for xu in 0 1; do
  DBGSOM_ACC_XU2=$xu python bench.py --no-e2e --no-cpu --no-strong --no-parity --no-fit > gpurun_out/k2ab_bench_xu$xu.json 2>> gpurun_out/k2ab.err
done
cat gpurun_out/k2ab.jsonl
