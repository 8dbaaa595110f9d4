"""profiles/r02_ncu_duo.json (the per-launch counts bench.py turns into pipe utilisation) from a raw ncu CSV page.
usage: python tools/ncu_json.py profiles/r02_ncu_duo_4096_raw.csv profiles/r02_ncu_duo.json [qps_per_launch]
The CSV is `ncu -i capture.ncu-rep --page raw --csv` of
  ncu --set full --clock-control none --import-source on -k regex:gomp_duo -c 1 python tools/profile_case.py 4096 1"""
import csv
import json
import sys

src, dst = sys.argv[1], sys.argv[2]
qps = int(sys.argv[3]) if len(sys.argv) > 3 else 4096
rows = list(csv.reader(open(src)))
head, units, vals = rows[0], rows[1], rows[2]


def get(name):
    i = head.index(name)
    v = float(vals[i].replace(",", ""))
    scale = {"Mbyte": 1e6, "Kbyte": 1e3, "Gbyte": 1e9, "byte": 1.0, "us": 1e-3, "ms": 1.0, "s": 1e3, "ns": 1e-6}.get(units[i], 1.0)
    return v * scale


rd, wr = get("dram__bytes_read.sum"), get("dram__bytes_write.sum")
inst, wav = get("smsp__inst_executed.sum"), get("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum")
out = {
    "kernel": "gomp_duo_kernel<9,4,14>",
    "capture": "ncu --set full --clock-control none --import-source on -k regex:gomp_duo -c 1, tools/profile_case.py 4096 1 (one launch, "
               "4096 QPs of BASELINE configs[1], two-slot TMEM back end; round 2, final build)",
    "qps_per_launch": qps,
    "dram_bytes_read": int(rd), "dram_bytes_write": int(wr), "dram_bytes_per_qp": (rd + wr) / qps,
    "gpu_time_duration_ms_solo": get("gpu__time_duration.sum"),
    "warp_instructions": int(inst), "warp_instructions_per_qp": inst / qps,
    "smem_wavefronts": int(wav), "smem_wavefronts_per_qp": wav / qps,
    "registers_per_thread": int(get("launch__registers_per_thread")),
    "issue_pct_of_active_solo": get("smsp__issue_active.avg.pct_of_peak_sustained_active"),
    "fp64_pipe_pct_of_active_solo": get("sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active"),
    "note_utilisation": "bench.py spreads these per-launch counts over its own measured step time (launches of consecutive steps overlap): "
                        "wavefronts / (148 SMs x cycles per step), warp instructions / (592 SM sub-partitions x cycles per step)",
    "report": src,
}
json.dump(out, open(dst, "w"), indent=1)
print(json.dumps(out, indent=1))
