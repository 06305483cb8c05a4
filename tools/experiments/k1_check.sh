# K1 change check: parity tests of the fading kernels + the bench lines whose steps contain sos_poly_coef
timeout 800 python -m pytest tests/test_fading_gpu.py tests/test_sinc.py tests/test_fading_plugin_gpu.py tests/test_oracle_golden.py -m gpu -q 2>&1 | grep -E "^E  |FAILED|passed|failed" | head -20
for cfg in C2 C5 C4 C1; do
timeout 300 python bench.py --config $cfg --only --steps 20 --no-simulation --no-cpu-baseline --no-e2e 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$cfg', '%.4g'%d['value'], d['ms_per_step'], d['roofline']['kernel'], d['roofline']['kernel_ms_per_step'], d['roofline']['other_kernels_ms_per_step'], d['parity']['rel_l2_vs_oracle'])"
done
timeout 300 python bench.py --config C5 --sinc --only --steps 20 --no-simulation --no-cpu-baseline --no-e2e 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('C5 sinc', '%.4g'%d['value'], d['ms_per_step'], d['roofline']['kernel'], d['roofline']['kernel_ms_per_step'], d['roofline']['other_kernels_ms_per_step'], d['parity']['rel_l2_vs_oracle'])"
