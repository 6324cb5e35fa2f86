mkdir -p gpurun_out/s3
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 > gpurun_out/s3/tests2.log
timeout 600 python bench.py > gpurun_out/s3/bench_b18.json 2> gpurun_out/s3/bench_b18.err
cat gpurun_out/s3/tests2.log; cat gpurun_out/s3/bench_b18.json; tail -3 gpurun_out/s3/bench_b18.err
