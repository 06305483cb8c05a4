#!/bin/bash
# Round-2 profiling pass (run under gpurun): the launch list of the bench command and one `ncu --set full` capture of the
# dominant kernel of every config; summaries land in gpurun_out/ (copy what is to be judged into profiles/).
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches_bench_c2.csv \
  python bench.py --steps 2 --warmup 1 --only --no-cpu-baseline --no-simulation > gpurun_out/r02_launches_bench_c2.log 2>&1
tools/ncu_one.sh tdl_tma r02_c2_tma 2 --config C2 --no-simulation
tools/ncu_one.sh cdl_umma r02_c3_umma2 2 --config C3 --no-simulation
tools/ncu_one.sh tdl_window r02_c5_window 2 --config C5 --no-simulation
tools/ncu_one.sh tdl_poly64 r02_c2_f64 2 --config C2 --precision f64 --no-simulation
tools/ncu_one.sh spatial_gemm r02_c4_gemm 2 --config C4 --no-simulation
for t in r02_c2_tma r02_c3_umma2 r02_c5_window r02_c2_f64 r02_c4_gemm; do
  python - <<EOF
import csv, json
rows = list(csv.reader(open("gpurun_out/ncu_${t}_raw.csv")))
m = dict(zip(rows[0], zip(rows[1], rows[-1])))
scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
def b(k):
    u, v = m[k]
    return float(v.replace(",", "")) * scale[u]
print(json.dumps({"tag": "${t}", "kernel": m["Kernel Name"][1], "dram_bytes_read": b("dram__bytes_read.sum"),
                  "dram_bytes_write": b("dram__bytes_write.sum"), "duration": m["gpu__time_duration.sum"]}))
EOF
done | tee gpurun_out/r02_ncu_traffic_raw.jsonl
