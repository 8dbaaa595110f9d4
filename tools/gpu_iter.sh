#!/bin/bash
# One development round on the GPU box: quick parity of the two-slot kernel, phase timers (profile build), bench (20 and 64 steps).
# usage: tools/gpu_iter.sh <tag>      (build both libraries first: make -C osqp-solver_b200 -j8 && make -C osqp-solver_b200 -j8 PROFILE=1 OUT=$PWD/osqp-solver_b200/libgomp_cuda_prof.so)
tag=${1:-x}
python tools/duo_check.py 320 0 > gpurun_out/${tag}_check.log 2>&1
cat gpurun_out/${tag}_check.log | grep -v "^iters hist"
GOMP_CUDA_LIB=$PWD/osqp-solver_b200/libgomp_cuda_prof.so GOMP_PROFILE_PHASES=2 python tools/profile_case.py 592 1 > gpurun_out/${tag}_phases.log 2>&1
grep -h "per ADMM\|marks" gpurun_out/${tag}_phases.log
python bench.py --steps 20 --warmup 5 > gpurun_out/${tag}_b20.json 2> gpurun_out/${tag}_b20.err
python bench.py --steps 96 --warmup 12 > gpurun_out/${tag}_b96.json 2> gpurun_out/${tag}_b96.err
python - <<PY
import json
for f in ("gpurun_out/${tag}_b20.json", "gpurun_out/${tag}_b96.json"):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, "value %.0f e2e %.0f ms/step %.2f" % (d["value"], d["e2e"]["value"], d["ms_per_step"]))
    except Exception as e:
        print(f, "failed", e)
PY
