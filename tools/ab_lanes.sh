# A/B of the SM lanes of bench.py (GOMP_BENCH_LANES, GOMP_BENCH_LANE_RAMP) at the driver's 20 steps and at 96
run() { # tag steps env...
  tag=$1; s=$2; shift 2
  env "$@" GOMP_BENCH_EXTRA=0 python bench.py --steps $s --warmup 5 > gpurun_out/ln_${tag}_${s}.json 2> gpurun_out/ln_${tag}_${s}.err
}
rm -f gpurun_out/ln_*
for r in ${RAMPS:-0 1 2 3}; do
  for s in ${STEPS:-20}; do
    run ramp$r $s GOMP_BENCH_LANE_RAMP=$r
  done
done
python - <<PY
import json,glob
for f in sorted(glob.glob("gpurun_out/ln_*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, "value %.0f e2e %.0f ms/step %.2f launches %d" % (d["value"], d["e2e"]["value"], d["ms_per_step"], d["gpu_launches"]))
    except Exception as e:
        print(f, "failed", e)
PY
