"""Multi-rank plumbing on CPU (gloo, world size 2): the path shards by frame with no data-path collective, so what
there is to test is the rendezvous, the per-rank shards, the max-over-ranks reduction and the reference arm's
"rank 0 alone prints" contract."""
import json
import os
import socket
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out_dir):
    import torch
    import torch.distributed as dist
    sys.path.insert(0, ROOT)
    from tools import corpus
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    shard = corpus.make("log", 1 << 20, shard=rank)
    t = torch.tensor([10.0 + rank], dtype=torch.float64)          # pretend per-rank time
    nbytes = torch.tensor([len(shard)], dtype=torch.int64)
    dist.barrier()
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dist.all_reduce(nbytes, op=dist.ReduceOp.SUM)
    np.save(os.path.join(out_dir, f"r{rank}.npy"), np.array([t.item(), nbytes.item(), int(shard[:4096].sum())]))
    dist.destroy_process_group()


def test_gloo_two_ranks_shards_and_reduction(tmp_path):
    import torch.multiprocessing as mp
    port = _free_port()
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    r0, r1 = np.load(tmp_path / "r0.npy"), np.load(tmp_path / "r1.npy")
    assert r0[0] == r1[0] == 11.0              # max over ranks, identical everywhere
    assert r0[1] == r1[1] == 2 * (1 << 20)     # whole-job bytes
    assert r0[2] != r1[2]                      # ranks work on different shards


def test_reference_arm_under_torchrun_prints_once():
    port = _free_port()
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", str(port), os.path.join(ROOT, "bench.py"), "--gpus", "2", "--impl", "reference", "--steps", "1",
           "--warmup", "0", "--bytes", str(4 << 20)]
    p = subprocess.run(cmd, capture_output=True, text=True, timeout=300, cwd=ROOT)
    assert p.returncode == 0, p.stderr[-2000:]
    lines = [l for l in p.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    line = json.loads(lines[0])
    assert line["impl"] == "reference" and line["n_gpus"] == 2 and line["cpu_baseline"]["kind"] == "port"
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["value"] > 0
