# A/B of the SM lanes of bench.py (GOMP_BENCH_LANES) and of the handles in flight (GOMP_BENCH_PIPE)
run() { # tag steps env...
  tag=$1; s=$2; shift 2
  env "$@" GOMP_BENCH_EXTRA=0 python bench.py --steps $s --warmup 5 > gpurun_out/ln_${tag}_${s}.json 2> gpurun_out/ln_${tag}_${s}.err
}
rm -f gpurun_out/ln_*
for cfg in ${CFGS:-12_4 16_4 16_6 20_4}; do
  p=${cfg%_*}; l=${cfg#*_}
  for s in ${STEPS:-20}; do
    run p${p}_l$l $s GOMP_BENCH_LANES=$l GOMP_BENCH_PIPE=$p
  done
done
python - <<PY
import json,glob
for f in sorted(glob.glob("gpurun_out/ln_*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, "value %.0f e2e %.0f (%d steps) ms/step %.2f launches %d" % (d["value"], d["e2e"]["value"], d["e2e"]["steps"], d["ms_per_step"], d["gpu_launches"]))
    except Exception as e:
        print(f, "failed", e)
PY
