#!/bin/bash
# A/B runs of bench.py under environment switches.  Usage (under gpurun): bash tools/ab_bench.sh <tag> "VAR=a VAR2=b" "VAR=c" ...
TAG=$1; shift
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
: > gpurun_out/${TAG}_ab.txt
for cfg in "$@"; do
  out=$(env $cfg timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | tail -1)
  echo "$cfg :: $(echo "$out" | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('ms_per_step=%.3f value=%.0f e2e=%.0f e2e_resident=%.0f gemm_ms=%.3f launches=%d' % (d['ms_per_step'], d['value'], d['e2e']['value'], d['e2e']['maps_resident']['value'], d['roofline']['launch_ms'], d['gpu_launches']))" 2>&1)" | tee -a gpurun_out/${TAG}_ab.txt
done
