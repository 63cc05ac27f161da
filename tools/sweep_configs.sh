#!/bin/bash
# BASELINE.json configs[3] / configs[4] shapes on one GPU: mixed corpus and the tick-record chunk-size sweep.
# Usage: bash tools/sweep_configs.sh [bytes] > profiles/rN_sweep.jsonl
B=${1:-1073741824}
for wl in decode64k compress128k; do
  python bench.py --workload $wl --corpus mixed --bytes $B --steps 3 --warmup 3 --no-cpu-baseline 2>/dev/null | tail -1
  for c in 4096 16384 65536 262144 1048576; do
    python bench.py --workload $wl --corpus tick --chunk $c --bytes $B --steps 3 --warmup 3 --no-cpu-baseline 2>/dev/null | tail -1
  done
done
