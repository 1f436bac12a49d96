#!/bin/bash
# Round-2 evidence in one gpurun call (1 GPU): bench lines, ncu launch list, ncu --set full of the three top kernels.
#   gpurun --timeout 2400 -- tools/final_profiles.sh         -> gpurun_out/r02_*  (copy / summarise into profiles/ afterwards)
cd "$(dirname "$0")/.."
python bench.py --steps 5 --warmup 3 > gpurun_out/r02_bench_final.json 2> gpurun_out/r02_bench_final.err || { tail -5 gpurun_out/r02_bench_final.err; exit 1; }
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02_bench_reference_arm.json 2> gpurun_out/r02_bench_reference_arm.err
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extras"
$CMD > gpurun_out/plain_final.log 2>&1 || { echo "plain run failed"; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_ncu_launches_final.csv $CMD > /dev/null 2>&1
for K in symbol_kernel_p vit_simd_forward vit_simd_traceback; do tools/ncu_big.sh $K r02_full_$K; done
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/r02_smi.csv
