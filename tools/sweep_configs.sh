#!/bin/bash
# BASELINE.json configs[3] / configs[4] shapes: mixed corpus and the tick-record chunk-size sweep, both directions, on the
# GPUs of this launch (run under torchrun for N > 1: `torchrun ... tools/...` is not needed, bench.py reads RANK itself).
# Usage: bash tools/sweep_configs.sh [bytes] [launcher...] > profiles/rN_sweep.jsonl
#   e.g. bash tools/sweep_configs.sh 1073741824 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511
B=${1:-1073741824}; shift
L=${@:-python}
for wl in decode64k compress128k; do
  $L bench.py --workload $wl --corpus mixed --bytes $B --steps 3 --warmup 3 --no-cpu-baseline --no-extras 2>/dev/null | tail -1
  for c in 4096 16384 65536 262144 1048576; do
    $L bench.py --workload $wl --corpus tick --chunk $c --bytes $B --steps 3 --warmup 3 --no-cpu-baseline --no-extras 2>/dev/null | tail -1
  done
done
# dictionary decode on the tiny-message shape: 4 KiB log messages with a 32 KiB trained dictionary
$L bench.py --corpus log --chunk 4096 --bytes $((B / 4)) --dictionary 32768 --steps 3 --warmup 3 --no-extras 2>/dev/null | tail -1
$L bench.py --corpus log --chunk 4096 --bytes $((B / 4)) --steps 3 --warmup 3 --no-cpu-baseline --no-extras 2>/dev/null | tail -1
