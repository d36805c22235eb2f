#!/bin/bash
# Round-2 ncu captures (run on the GPU box through gpurun): launch lists and one --set full capture per dominant kernel.
set -x
O=gpurun_out
timeout -s KILL 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r02_launches_bench_eval.csv \
    python bench.py --steps 2 --warmup 3 --train-batch 0 --no-extras --cpu-sample 20000 > $O/r02_ncu_a.log 2>&1
timeout -s KILL 300 ncu --set full --clock-control none --import-source on -k regex:chain_umma_kernel -s 3 -c 1 -f -o $O/prof_r02_chain16 \
    python bench.py --steps 1 --warmup 3 --train-batch 0 --no-extras --cpu-sample 20000 > $O/r02_ncu_b.log 2>&1
timeout -s KILL 300 ncu --set full --clock-control none -k regex:chain_umma_pp_kernel -s 3 -c 1 -f -o $O/prof_r02_pp \
    python bench.py --workload two_moons_conditional --steps 1 --warmup 3 --train-batch 0 --no-extras --cpu-sample 20000 > $O/r02_ncu_c.log 2>&1
timeout -s KILL 300 ncu --set full --clock-control none -k 'regex:chain_umma_kernel|img_nt_kernel|img_tn_kernel' -s 8 -c 3 -f -o $O/prof_r02_train \
    python scripts/train_step_once.py 262144 0 262144 1 > $O/r02_ncu_d.log 2>&1
for n in chain16 pp train; do
    ncu -i $O/prof_r02_$n.ncu-rep --page raw --csv > $O/prof_r02_${n}_raw.csv 2>/dev/null
done
ls -la $O/prof_r02_*
