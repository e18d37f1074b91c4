"""N>1 host logic on CPU: two gloo ranks shard buckets disjointly and agree on the max-over-ranks time."""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = r'''
import os, sys, json
sys.path.insert(0, %r)
import torch, torch.distributed as dist
from ema_b200 import shard
dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
mine = shard.plan(6, rank, world, 25)
allp = [None] * world
dist.all_gather_object(allp, mine)
t = shard.reduce_max([1.0 + rank, 5.0 - rank], world)
if rank == 0:
    print(json.dumps({"plans": allp, "tmax": t}))
dist.destroy_process_group()
'''


def test_two_ranks_shard_disjointly(tmp_path):
    script = tmp_path / "w.py"
    script.write_text(WORKER % ROOT)
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
                        "--master-port", "29731", str(script)], capture_output=True, text=True, env=env, timeout=300)
    assert r.returncode == 0, r.stderr[-2000:]
    import json
    out = json.loads([l for l in r.stdout.splitlines() if l.startswith("{")][-1])
    a, b = out["plans"]
    assert len(a) == len(b) == 6 and not set(a) & set(b), "ranks must take disjoint buckets"
    assert sorted(a + b) == list(range(12))
    assert out["tmax"] == [2.0, 5.0]


def test_single_rank_plan():
    from ema_b200 import shard
    assert shard.plan(4, 0, 1, 3) == [0, 1, 2, 0]
    assert shard.reduce_max([1.5], 1) == [1.5]
