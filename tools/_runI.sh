O=gpurun_out/r1h; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3 > $O/tests.log
timeout 600 python bench.py > $O/bench_n1.json 2> $O/bench_n1.err
timeout 300 python bench.py --mode lk --steps 5 --warmup 2 > $O/bench_lk.json 2> $O/bench_lk.err
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/lk_launches.csv python bench.py --mode lk --steps 3 --warmup 2 --no-cpu-baseline > $O/ncu_lk.log 2>&1
python __graft_entry__.py --smoke 2>&1 | tail -1 >> $O/tests.log
cat $O/tests.log; cut -c1-200 $O/bench_n1.json; cut -c1-200 $O/bench_lk.json
