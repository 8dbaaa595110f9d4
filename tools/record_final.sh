# Records of the round's final build (one GPU): smoke, GPU tests, both bench arms at the driver's settings, the default bench,
# then -- after those exited 0 without ncu -- the launch list of the bench command and one ncu --set full capture of the kernel.
set -x
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -n 1
python -m pytest tests -m gpu -x -q 2>&1 | tail -n 2
python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/final_ref20.json 2> gpurun_out/final_ref20.err
python bench.py --steps 20 --warmup 5 > gpurun_out/final_b20.json 2> gpurun_out/final_b20.err
python bench.py > gpurun_out/final_bdef.json 2> gpurun_out/final_bdef.err
python tools/profile_case.py 4096 1 > gpurun_out/final_case.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:gomp_duo -c 1 -o gpurun_out/final_duo -f python tools/profile_case.py 4096 1 > gpurun_out/final_ncu.log 2>&1
ncu -i gpurun_out/final_duo.ncu-rep --page raw --csv > gpurun_out/final_duo_raw.csv 2>/dev/null
GOMP_BENCH_EXTRA=0 python bench.py --steps 2 --warmup 1 > /dev/null 2>&1 && \
GOMP_BENCH_EXTRA=0 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/final_launches.csv python bench.py --steps 2 --warmup 1 > /dev/null 2>&1
ls -la gpurun_out/final_*
python - <<PY
import json
for f in ("gpurun_out/final_b20.json","gpurun_out/final_bdef.json","gpurun_out/final_ref20.json"):
    d=json.loads(open(f).read().strip().splitlines()[-1]); print(f, round(d["value"]), round(d["e2e"]["value"]), d.get("ms_per_step"), (d.get("roofline") or {}).get("frac"))
PY
