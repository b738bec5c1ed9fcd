#!/bin/bash
# Mimics the round-end driver on a fresh box: build check, smoke, GPU tests, reference arm, own arm.
set -x
mkdir -p gpurun_out
T0=$(date +%s)
python -c 'import __graft_entry__ as g; g.build(); g.smoke()' > gpurun_out/flow_smoke.log 2>&1; echo "smoke rc=$? t=$(( $(date +%s)-T0 ))"
python -m pytest tests/ -x -q -m gpu > gpurun_out/flow_pytest.log 2>&1; echo "pytest rc=$? t=$(( $(date +%s)-T0 ))"
tail -3 gpurun_out/flow_pytest.log
python bench.py --impl reference --gpus 1 --steps 2 --warmup 1 > gpurun_out/flow_bench_ref.json 2> gpurun_out/flow_bench_ref.err; echo "ref rc=$? t=$(( $(date +%s)-T0 ))"
python bench.py > gpurun_out/flow_bench.json 2> gpurun_out/flow_bench.err; echo "bench rc=$? t=$(( $(date +%s)-T0 ))"
cat gpurun_out/flow_bench_ref.json gpurun_out/flow_bench.json
tail -3 gpurun_out/flow_smoke.log
