#!/bin/bash
# round-2 GPU pass 1: full GPU test suite (no -x), bench default graph/eager, other configs
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total --format=csv > gpurun_out/r2a_env.txt 2>&1
nproc >> gpurun_out/r2a_env.txt
timeout 1800 python -m pytest tests -m gpu -q -rA --timeout 1500 -p no:cacheprovider > gpurun_out/r2a_tests.txt 2>&1
echo "pytest rc $?" >> gpurun_out/r2a_tests.txt
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r2a_bench.json 2> gpurun_out/r2a_bench.err
echo "bench rc $?" >> gpurun_out/r2a_bench.err
timeout 300 python bench.py --steps 20 --warmup 5 --graph off --no-cpu-baseline --no-vcr > gpurun_out/r2a_bench_eager.json 2> gpurun_out/r2a_bench_eager.err
timeout 300 python bench.py --config playback --steps 40 --warmup 5 --no-cpu-baseline > gpurun_out/r2a_bench_playback.json 2> gpurun_out/r2a_bench_playback.err
timeout 300 python bench.py --config c3 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2a_bench_c3.json 2> gpurun_out/r2a_bench_c3.err
tail -5 gpurun_out/r2a_tests.txt
tail -c 1500 gpurun_out/r2a_bench.json
tail -3 gpurun_out/r2a_bench.err gpurun_out/r2a_bench_eager.err gpurun_out/r2a_bench_playback.err gpurun_out/r2a_bench_c3.err
