#!/usr/bin/env bash
# BASELINE configs 3 / 4 / 5 and the host-link micro-run on 1, 2, 4, 8 GPUs of one box (one process per GPU).
#   bash benchmarks/run_multi_gpu.sh OUTDIR "1 2 4 8"
# Every line of OUTDIR/*.jsonl is one JSON record printed by rank 0.
set -u
OUT=${1:-gpurun_out/multi}
NS=${2:-"1 2 4 8"}
mkdir -p "$OUT"
PORT=29511
run() {  # run N script args...
  local n=$1; shift
  PORT=$((PORT + 1))
  if [ "$n" = 1 ]; then python "$@"; else
    python -m torch.distributed.run --nnodes=1 --nproc-per-node "$n" --master-addr 127.0.0.1 --master-port $PORT "$@"; fi
}
for n in $NS; do
  echo "== N=$n train config 4 (reference dataflow)" >&2
  run $n benchmarks/train_step.py --steps 6 --warmup 2 --profile >> "$OUT/train_config4_reference_dataflow.jsonl" 2>> "$OUT/err.log"
  echo "== N=$n train config 4 (fused heads + channels_last_3d volume)" >&2
  run $n benchmarks/train_step.py --steps 6 --warmup 2 --profile --fuse-upsample --channels-last-3d >> "$OUT/train_config4_optins.jsonl" 2>> "$OUT/err.log"
  echo "== N=$n train config 3" >&2
  run $n benchmarks/train_step.py --steps 6 --warmup 2 --no-real >> "$OUT/train_config3.jsonl" 2>> "$OUT/err.log"
  echo "== N=$n sweep" >&2
  run $n benchmarks/sweep_sharded.py --out "$OUT/sweep_config5_n$n.json" > /dev/null 2>> "$OUT/err.log"
  echo "== N=$n h2d" >&2
  run $n benchmarks/h2d_micro.py >> "$OUT/h2d_micro.jsonl" 2>> "$OUT/err.log"
done
tail -n 3 "$OUT"/*.jsonl >&2
