#!/bin/bash
# multi-GPU pass: bash scripts/gpu_multi.sh N [full]  (run under gpurun --gpus N)
N=${1:-2}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29571"
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/r2_multi_env_n$N.txt
if [ "$N" = "2" ] && [ "$2" = "full" ]; then
  timeout 600 python -m pytest tests/test_gpu_exchange.py -m gpu -q -rA -p no:cacheprovider > gpurun_out/r2_exchange_tests_n$N.txt 2>&1
  tail -3 gpurun_out/r2_exchange_tests_n$N.txt
fi
: > gpurun_out/r2_exchange_check_n$N.txt
for algo in push_all owner_push; do for sh in 0 3; do
  if [ "$N" != "2" ] && [ "$algo" = "push_all" ] && [ "$sh" = "3" ]; then continue; fi
  timeout 150 $TR scripts/exchange_check.py --algo $algo --sh $sh --steps 5 2>&1 | grep -E "^\{|Error|error" | tail -2 >> gpurun_out/r2_exchange_check_n$N.txt
done; done
cat gpurun_out/r2_exchange_check_n$N.txt
timeout 240 $TR bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/r2_bench_n$N.json 2> gpurun_out/r2_bench_n$N.err
echo "bench rc $?"; tail -2 gpurun_out/r2_bench_n$N.err | cut -c1-300
timeout 200 $TR bench.py --gpus $N --steps 20 --warmup 5 --exchange nccl --no-vcr > gpurun_out/r2_bench_n${N}_nccl.json 2> gpurun_out/r2_bench_n${N}_nccl.err
echo "nccl rc $?"
timeout 200 $TR bench.py --gpus $N --config vcr --steps 10 --warmup 3 > gpurun_out/r2_bench_n${N}_vcr.json 2> gpurun_out/r2_bench_n${N}_vcr.err
echo "vcr rc $?"
python - <<PY
import json
for v in ("", "_nccl", "_vcr"):
    try:
        d=json.loads(open(f"gpurun_out/r2_bench_n$N{v}.json").read().strip().splitlines()[-1])
        print("n$N"+v, round(d["value"],1), "ms", round(d["ms_per_step"],3), "e2e", round(d["e2e"]["value"],1), d["config"]["launch"][:20], d["config"].get("graph_error"), (d.get("exchange_check") or {}).get("ok"), (d.get("vcr") or {}).get("value"), (d.get("vcr") or {}).get("error"), (d.get("vcr") or {}).get("graph_error"))
    except Exception as e:
        print(v, "ERR", e)
PY
