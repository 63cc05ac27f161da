N=$1
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533"
timeout 600 $TR bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/r2_bench_decode_${N}gpu.json 2> gpurun_out/r2_multi_${N}.err; tail -1 gpurun_out/r2_bench_decode_${N}gpu.json | cut -c1-300
timeout 300 $TR bench.py --gpus $N --corpus mixed --steps 3 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/r2_bench_decode_mixed_${N}gpu.json 2>> gpurun_out/r2_multi_${N}.err; tail -1 gpurun_out/r2_bench_decode_mixed_${N}gpu.json | cut -c1-200
timeout 300 $TR bench.py --gpus $N --workload compress128k --corpus mixed --steps 3 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/r2_bench_compress_mixed_${N}gpu.json 2>> gpurun_out/r2_multi_${N}.err; tail -1 gpurun_out/r2_bench_compress_mixed_${N}gpu.json | cut -c1-200
if [ "$N" = "8" ]; then
  for wl in decode64k compress128k; do for c in 4096 65536 1048576; do
    timeout 300 $TR bench.py --gpus $N --workload $wl --corpus tick --chunk $c --steps 3 --warmup 3 --no-cpu-baseline --no-extras 2>> gpurun_out/r2_multi_${N}.err | tail -1 >> gpurun_out/r2_sweep_8gpu.jsonl
  done; done
  wc -l gpurun_out/r2_sweep_8gpu.jsonl
fi
