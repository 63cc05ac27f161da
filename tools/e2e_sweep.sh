python - <<'PY'
import torch, time
dev=torch.device("cuda:0")
h=torch.empty(1<<30,dtype=torch.uint8).pin_memory(); d=torch.empty(1<<30,dtype=torch.uint8,device=dev)
hs=torch.empty(220<<20,dtype=torch.uint8).pin_memory(); ds=torch.empty(220<<20,dtype=torch.uint8,device=dev)
for name,(a,b) in {"d2h 1GiB":(h,d),"h2d 220MB":(ds,hs)}.items():
    for _ in range(2): a.copy_(b,non_blocking=True)
    torch.cuda.synchronize(); t=time.perf_counter()
    for _ in range(5): a.copy_(b,non_blocking=True)
    torch.cuda.synchronize(); dt=(time.perf_counter()-t)/5
    print(name, "%.2f ms %.1f GB/s"%(dt*1e3,a.numel()/dt/1e9))
s1,s2=torch.cuda.Stream(),torch.cuda.Stream()
torch.cuda.synchronize(); t=time.perf_counter()
for _ in range(5):
    with torch.cuda.stream(s1): h.copy_(d,non_blocking=True)
    with torch.cuda.stream(s2): ds.copy_(hs,non_blocking=True)
torch.cuda.synchronize(); dt=(time.perf_counter()-t)/5
print("both directions %.2f ms"%(dt*1e3))
PY
for cfg in "8 32" "4 48" "8 16" "8 64" "6 32" "8 24" "2 128" "1 2048"; do set -- $cfg
  echo "streams $1 slice $2: $(ZSTDB200_STREAMS=$1 ZSTDB200_SLICE_MB=$2 python bench.py --steps 5 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c 'import sys,json; j=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(j["value"], j["e2e"]["value"])')"
done
