for cfg in C5 C1; do for mode in auto poly_gather poly_window; do
timeout 300 python bench.py --config $cfg --only --steps 20 --no-simulation --no-cpu-baseline --no-e2e --sos-mode $mode > gpurun_out/r02_${cfg}_${mode}.json 2> gpurun_out/r02_${cfg}_${mode}.err
python - <<EOF
import json
try:
    d=json.loads(open("gpurun_out/r02_${cfg}_${mode}.json").read().strip().splitlines()[-1])
    print("$cfg $mode", "%.3g"%d["value"], d["ms_per_step"], d["roofline"]["kernel"], d["roofline"]["kernel_ms_per_step"], d["roofline"]["frac"], d["roofline"]["other_kernels_ms_per_step"], d["parity"]["rel_l2_vs_oracle"], d["run"]["plan"])
except Exception as e:
    print("$cfg $mode failed", e); print(open("gpurun_out/r02_${cfg}_${mode}.err").read()[-500:])
EOF
done; done
