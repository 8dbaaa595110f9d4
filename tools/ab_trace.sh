export CUDA_DEVICE_MAX_CONNECTIONS=32
GOMP_HEAD_CTAS=0 GOMP_MAX_CTAS=37 python tools/pipe_trace.py 20 host > gpurun_out/trace_lh.log 2>&1
GOMP_HEAD_CTAS=0 GOMP_MAX_CTAS=37 python tools/pipe_trace.py 20 dev > gpurun_out/trace_ld.log 2>&1
GOMP_HEAD_CTAS=0 python tools/pipe_trace.py 20 dev > gpurun_out/trace_bd.log 2>&1
for f in gpurun_out/trace_lh.log gpurun_out/trace_ld.log gpurun_out/trace_bd.log; do tail -n 1 $f; done
