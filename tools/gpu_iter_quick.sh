#!/bin/bash
# Quick development round on the GPU box: parity of the two-slot kernel, phase timers (profile build), one 96-step bench.
# usage: tools/gpu_iter_quick.sh <tag>   (build both libraries first, see tools/gpu_iter.sh)
tag=${1:-x}
python tools/duo_check.py 320 0 > gpurun_out/${tag}_check.log 2>&1
grep -v "^iters hist" gpurun_out/${tag}_check.log | tail -n 4
GOMP_CUDA_LIB=$PWD/osqp-solver_b200/libgomp_cuda_prof.so GOMP_PROFILE_PHASES=2 python tools/profile_case.py 592 1 > gpurun_out/${tag}_phases.log 2>&1
grep -h "per ADMM\|marks" gpurun_out/${tag}_phases.log
GOMP_BENCH_EXTRA=0 python bench.py --steps 96 --warmup 12 > gpurun_out/${tag}_b96.json 2> gpurun_out/${tag}_b96.err
python - <<PY
import json
d = json.loads(open("gpurun_out/${tag}_b96.json").read().strip().splitlines()[-1])
print("96 steps: value %.0f e2e %.0f ms/step %.2f" % (d["value"], d["e2e"]["value"], d["ms_per_step"]))
PY
