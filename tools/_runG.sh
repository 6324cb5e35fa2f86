O=gpurun_out/r1g; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3 > $O/tests.log
timeout 600 python bench.py > $O/bench_n1.json 2> $O/bench_n1.err
timeout 300 python bench.py --mode sequence --no-cpu-baseline > $O/bench_seq.json 2> $O/bench_seq.err
timeout 300 python bench.py --frame 4k --batch 9 --steps 10 --no-cpu-baseline > $O/bench_4k.json 2> $O/bench_4k.err
timeout 300 python bench.py --frame vga --batch 72 --steps 10 --no-cpu-baseline > $O/bench_vga.json 2> $O/bench_vga.err
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/bench_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $O/ncu_bench.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_iter_v -s 21 -c 1 -o $O/iter_v -f python tools/profile_run.py 2 18 > $O/ncu_iter.log 2>&1
cat $O/tests.log; for f in n1 seq 4k vga; do echo "--- $f"; cut -c1-260 $O/bench_$f.json; tail -2 $O/bench_$f.err; done
