# drops/s of the C1 script through Simulation.run() for several lane / helper counts (reference on all cores once)
for lw in "64 15" "128 15" "256 15" "512 15" "256 16" "128 12"; do set -- $lw
timeout 200 python tools/simulation_campaign.py --config c1 --samples 400 --lanes $1 --workers $2 --skip-reference --skip-serial --out gpurun_out/c1_l$1_w$2.json > /dev/null 2>&1
python - <<EOF
import json
d=json.load(open("gpurun_out/c1_l$1_w$2.json")); g=d["gpu_batched"]; print("lanes $1 workers $2: %.0f drops/s"%g["drops_per_s"], g.get("owner_seconds"), "rounds", g["rounds"])
EOF
done
