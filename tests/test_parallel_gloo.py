"""N>1 host logic on CPU: world_size-2 gloo run of the shard + statistics all-reduce."""
import os
import socket
import subprocess
import sys

import rgbd_b200
from rgbd_b200.parallel import shard_range

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = r"""
import os, sys, json
sys.path.insert(0, os.environ["REPO"])
import torch, torch.distributed as dist
import rgbd_b200
from rgbd_b200.parallel import shard_range, new_stats, add_pair_stats, allreduce_stats, summarize
dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
lo, hi = shard_range(7, rank, world)
st = new_stats()
for i in range(lo, hi):
    g = torch.Generator(); g.manual_seed(i)
    x = torch.rand(1, 3, 8, 8, generator=g); d = torch.rand(1, 1, 8, 8, generator=g)
    add_pair_stats(st, [[b"\0" * (i + 1)], [b"\0" * 4]], [[b"\0" * 8], [b"\0" * 4]], x, d, x * 0.5, d)
tot = allreduce_stats(st)
if rank == 0:
    print(json.dumps({"tot": tot, "sum": summarize(tot)}))
dist.destroy_process_group()
"""


def test_shard_range_covers_everything():
    for n in (0, 1, 7, 256):
        for world in (1, 2, 3, 8):
            spans = [shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1


def test_stats_allreduce_world2(tmp_path):
    import json
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    env = dict(os.environ, REPO=ROOT, OMP_NUM_THREADS="1")
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                          "--master-addr", "127.0.0.1", "--master-port", str(port), str(script)],
                         env=env, capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr[-2000:]
    line = [l for l in out.stdout.splitlines() if l.startswith("{")][-1]
    res = json.loads(line)
    assert res["tot"]["pairs"] == 7 and res["tot"]["pixels"] == 7 * 64
    assert res["tot"]["bits_r"] == 8.0 * (sum(range(1, 8)) + 4 * 7)
    assert res["tot"]["se_d"] == 0.0 and res["sum"]["psnr_d"] == 99.0
    assert 0 < res["sum"]["psnr_r"] < 30
