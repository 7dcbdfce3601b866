#!/bin/bash
# multi-GPU check after a single-GPU change: bash scripts/gpu_multi_lite.sh N   (run under gpurun --gpus N)
# one in-run exchange check (owner_push / push_all as the world size selects) + the default bench line (weak AHDS
# series with the config-4 strong sub-run)
N=${1:-2}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29571"
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/r2_multi_env_n$N.txt
python -m gaussianip_b200.build > /dev/null 2>&1
timeout 300 $TR bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/r2_bench_n$N.json 2> gpurun_out/r2_bench_n$N.err
echo "bench rc $?"; tail -2 gpurun_out/r2_bench_n$N.err | cut -c1-300
python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r2_bench_n$N.json").read().strip().splitlines()[-1])
    print("n$N", round(d["value"],1), "ms", round(d["ms_per_step"],3), "e2e", round(d["e2e"]["value"],1), d["config"]["launch"][:20], d["config"].get("graph_error"), d.get("exchange_check"), (d.get("vcr") or {}).get("value"), (d.get("vcr") or {}).get("error"))
except Exception as e:
    print("ERR", e)
PY
