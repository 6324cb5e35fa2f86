#!/bin/bash
# Final artefacts of a round beside tools/gpu_profile_round.sh: sparse-path launch list + ncu --set full of its kernels,
# single-pair latencies, VGA / 4K bench lines, parameter sweep.  Output: gpurun_out/$1/
D=gpurun_out/${1:-r2}
mkdir -p $D
timeout 300 python tools/latency_bench.py > $D/latency.txt 2>&1; tail -2 $D/latency.txt
timeout 300 python tools/param_sweep.py > $D/param_sweep.txt 2>&1; tail -3 $D/param_sweep.txt
timeout 300 python bench.py --frame vga --batch 72 --no-cpu-baseline --no-extras > $D/bench_vga.json 2> $D/bench_vga.err
timeout 300 python bench.py --frame 4k --batch 18 --no-cpu-baseline --no-extras > $D/bench_4k.json 2> $D/bench_4k.err
timeout 300 python bench.py --mode lk --steps 30 --warmup 3 > $D/bench_lk.json 2> $D/bench_lk.err; tail -c 400 $D/bench_lk.json
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file $D/lk_launches.csv python tools/lk_profile.py 5 > $D/ncu_lk.log 2>&1
OFB_GRAPH=0 timeout 400 ncu --set full --clock-control none --import-source on -k regex:"k_lk_track|k_greedy_select|k_candidates|k_sobel_min_eig" -s 8 -c 4 -o $D/sparse_kernels -f python tools/lk_profile.py 2 > $D/ncu_sparse.log 2>&1
ls -la $D
