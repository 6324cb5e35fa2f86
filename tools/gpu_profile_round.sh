#!/bin/bash
# Round artefacts: bench line (N=1), reference arm, ncu launch list of the bench command, ncu --set full of the dominant
# kernel (level-0 plain launch of k_iter_v), PolyExp and the fused pyramid kernel.  Output: gpurun_out/$1/
D=gpurun_out/${1:-r2}
mkdir -p $D
timeout 900 python bench.py > $D/bench_n1.json 2> $D/bench_n1.err; tail -c 600 $D/bench_n1.json
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > $D/bench_ref.json 2> $D/bench_ref.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $D/bench_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extras > $D/ncu_bench.log 2>&1
OFB_STAGES=0 timeout 500 ncu --set full --clock-control none --import-source on -k regex:k_iter_v --launch-skip 10 --launch-count 1 -o $D/iter_v -f python tools/profile_run.py 1 36 > $D/ncu_iter.log 2>&1
OFB_OVERLAP=0 OFB_STAGES=0 timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_polyexp_march --launch-skip 7 --launch-count 1 -o $D/polyexp -f python tools/profile_run.py 2 36 > $D/ncu_px.log 2>&1
OFB_STAGES=0 timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_pyr_fast3 --launch-skip 1 --launch-count 1 -o $D/pyr3 -f python tools/profile_run.py 2 36 > $D/ncu_pyr.log 2>&1
ls -la $D
