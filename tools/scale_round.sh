#!/bin/bash
# One visit to an N-GPU box (gpurun --gpus 8): every multi-GPU number of the round, largest N first.
#   bash tools/scale_round.sh <tag> [maxN]
# Writes gpurun_out/<tag>_bench_n{N}.json (bench.py head inference: device-timed, e2e, e2e with resident maps),
# gpurun_out/train_scale.jsonl (config 3: stage-1 DDP step, 8 pairs per GPU), gpurun_out/scale_configs.jsonl (config 4:
# 400 keypoints x 32 pairs per GPU), gpurun_out/ddp_train_check.json and the two-GPU device-guard test.
TAG=${1:-rS}
MAXN=${2:-8}
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out
rm -f $O/train_scale.jsonl $O/scale_configs.jsonl
nvidia-smi --query-gpu=index,name,clocks.sm,clocks.max.sm --format=csv > $O/${TAG}_smi.txt 2>&1
nvidia-smi topo -m > $O/${TAG}_topo.txt 2>&1
PORT=29511
run() {  # run <N> <script> <args...>
  local n=$1; shift
  PORT=$((PORT + 1))
  if [ "$n" = "1" ]; then timeout 600 python "$@"
  else timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $PORT "$@"; fi
}
for N in 8 4 2 1; do
  [ $N -gt $MAXN ] && continue
  echo "== bench.py N=$N"
  run $N bench.py --gpus $N --steps 10 --warmup 3 --no-cpu-baseline > $O/${TAG}_bench_n$N.json 2> $O/${TAG}_bench_n$N.err
  tail -1 $O/${TAG}_bench_n$N.json | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('N=%d value=%.0f ms=%.3f e2e=%.0f e2e_resident=%.0f h2d_gbs=%.1f' % (d['n_gpus'], d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e']['maps_resident']['value'], d['e2e']['h2d_link_gbs_measured']))"
done
for N in 8 4 2 1; do
  [ $N -gt $MAXN ] && continue
  echo "== config 3 (train) N=$N"
  if [ "$N" = "1" ]; then run 1 tools/bench_train.py --pairs-per-rank 8 --graph --tag $TAG 2> $O/${TAG}_train_n$N.err | tail -1 | cut -c1-900
  else run $N tools/bench_train.py --pairs-per-rank 8 --tag $TAG 2> $O/${TAG}_train_n$N.err | tail -1 | cut -c1-900; fi
done
for N in 8 4 2 1; do
  [ $N -gt $MAXN ] && continue
  echo "== config 4 (400 keypoints) N=$N"
  run $N tools/bench_scale.py --keypoints 400 --pairs-per-rank 32 2> $O/${TAG}_c4_n$N.err | tail -1 | cut -c1-500
done
echo "== config 4 ragged, N=1"
run 1 tools/bench_scale.py --keypoints 400 --pairs-per-rank 32 --ragged 2>> $O/${TAG}_c4_n1.err | tail -1 | cut -c1-500
echo "== DDP gradient all-reduce check N=$MAXN"
run $MAXN tools/ddp_train_check.py 2> $O/${TAG}_ddp_check.err | tail -1 | cut -c1-1500
echo "== two-GPU device guard test"
timeout 300 python -m pytest tests -m gpu -q -k "follow_their_tensors_device" 2>&1 | tail -2
cp $O/train_scale.jsonl $O/${TAG}_train_scale.jsonl 2>/dev/null
cp $O/scale_configs.jsonl $O/${TAG}_scale_configs.jsonl 2>/dev/null
ls -la $O | grep ${TAG} | head -40
