python -m pytest tests -m gpu -x -q 2>&1 | tail -n 3
python tools/pipe_trace.py 20 host > gpurun_out/r02_trace_lanes_host.log 2>&1
python tools/pipe_trace.py 20 dev > gpurun_out/r02_trace_lanes_dev.log 2>&1
GOMP_BENCH_LANES=1 python tools/pipe_trace.py 20 dev > gpurun_out/r02_trace_nolanes_dev.log 2>&1
GOMP_BENCH_EXTRA=0 python bench.py --steps 2 --warmup 1 > gpurun_out/pre_ncu.json 2> gpurun_out/pre_ncu.err && \
GOMP_BENCH_EXTRA=0 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_bench_launches_lanes.csv python bench.py --steps 2 --warmup 1 > gpurun_out/ncu_run.log 2>&1
tail -n 2 gpurun_out/r02_trace_*.log
wc -l gpurun_out/r02_bench_launches_lanes.csv
