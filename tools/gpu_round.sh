#!/bin/bash
# One GPU-box visit: GPU tests (as the driver runs them), smoke, bench lines, ncu launch list and full captures.
# Usage (under gpurun): bash tools/gpu_round.sh <tag> [notests] [noncu]
TAG=${1:-rX}
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out
rm -f $O/parity_report.jsonl
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $O/${TAG}_smi.txt 2>&1
if [[ "$*" != *notests* ]]; then
  timeout 1500 python -m pytest tests -m gpu -q --maxfail=10 --durations=12 > $O/${TAG}_pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a $O/${TAG}_pytest_gpu.log
  tail -25 $O/${TAG}_pytest_gpu.log
  cp $O/parity_report.jsonl $O/${TAG}_parity_report.jsonl 2>/dev/null
  timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/${TAG}_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 $O/${TAG}_smoke.log
fi
timeout 600 python bench.py > $O/${TAG}_bench.json 2> $O/${TAG}_bench.err; echo "bench rc=$?"; tail -1 $O/${TAG}_bench.json
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > $O/${TAG}_bench_reference.json 2> $O/${TAG}_bench_reference.err; tail -1 $O/${TAG}_bench_reference.json
if [[ "$*" != *noncu* ]]; then
  CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline"
  $CMD > $O/${TAG}_plain.log 2>&1 &&
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file $O/${TAG}_launches.csv $CMD > $O/${TAG}_ncu_launch.log 2>&1
  echo "ncu launches rc=$?"
  $CMD > $O/${TAG}_plain2.log 2>&1 &&
  # gpurun_out may carry at most 64 MiB back: keep the full capture to ~20 launches (~2.5 MiB each with source)
  timeout 1200 ncu --set full --clock-control none --import-source on -k regex:"${NCU_KERNELS:-gemm_tc_pair_kernel|gnn_layer_kernel|spline_gather|sinkhorn_log|lap_topk|afau_attention|match_cls_stage2|ke_factored}" -s ${NCU_SKIP:-40} -c ${NCU_COUNT:-20} -o $O/${TAG}_prof $CMD > $O/${TAG}_ncu_full.log 2>&1
  echo "ncu full rc=$?"; ls -la $O/${TAG}_prof.ncu-rep
  # the association-graph layer kernels in their own small capture (the window above rarely reaches them)
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:"gnn_layer_kernel" -s 6 -c 3 -o $O/${TAG}_prof_gnn $CMD > $O/${TAG}_ncu_full_gnn.log 2>&1
  echo "ncu gnn rc=$?"
fi
ls -la $O | tail -12
