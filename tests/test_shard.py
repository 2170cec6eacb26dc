"""CPU-only: the N>1 host logic.  Shard plans partition the lit chunks, and the
oracle run shard by shard (each shard a separate world holding only its
resident chunks) reproduces the single-world result on every owned chunk —
the property the multi-GPU path relies on instead of a halo exchange.  Also a
world_size-2 gloo run of the bench's rendezvous / max-over-ranks plumbing."""
import os
import subprocess
import sys

import numpy as np
import pytest

import cases
import pyoracle
from vxgpu import shard, worlds

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("n", [1, 2, 3, 4, 8])
def test_plan_partitions_lit_chunks(n):
    shards = shard.plan(-1, -1, 18, 10, n)
    owned = [k for s in shards for k in s.owned]
    lit = {(cx, cz) for cx in range(0, 16) for cz in range(0, 8)}
    assert len(owned) == len(set(owned)) == len(lit) and set(owned) == lit
    for s in shards:
        ox, oz, w, d = s.window
        assert set(s.owned) <= set(s.lit) <= set(s.resident)
        assert len(s.resident) == w * d
        for (cx, cz) in s.owned:  # every owned chunk has its 8 neighbours resident
            assert all((cx + dx, cz + dz) in set(s.resident) for dx in (-1, 0, 1) for dz in (-1, 0, 1))


def _run(world, sh, table):
    ox, oz, w, d = sh.window
    be = pyoracle.OracleWorld(table, ox, oz, w, d, "port")
    for k in sh.resident:
        be.put_chunk(k[0], k[1], world.chunks[k])
    for k in sh.resident:
        be.prebuild_sky(*k)
    for k in sh.lit:
        be.light_chunk(k[0], k[1], True)
    out = {}
    for k in sh.owned:
        m = be.mesh_chunk(k[0], k[1])
        out[k] = (be.get_light(*k), m.vertices, m.indices, m.entry_floats,
                  np.frombuffer(m.entries.tobytes(), dtype=np.uint32))
    be.close()
    return out


def test_sharded_oracle_equals_single_world():
    table = cases.table()
    world = worlds.build_world("terrain", -1, -1, 6, 4, seed=77)
    rng = np.random.default_rng(5)
    for k in list(world.chunks):
        world.chunks[k] = worlds.sprinkle(world.chunks[k], rng, table, 10, 10)
    single = _run(world, shard.plan(-1, -1, 6, 4, 1)[0], table)
    for n in (2, 4):
        for sh in shard.plan(-1, -1, 6, 4, n):
            part = _run(world, sh, table)
            for k, arrs in part.items():
                for a, b in zip(arrs, single[k]):
                    assert a.shape == b.shape and np.array_equal(a, b), (n, sh.rank, k)


_GLOO = r"""
import os, sys, torch, torch.distributed as dist
sys.path.insert(0, sys.argv[1])
from vxgpu import shard
dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
shards = shard.plan(-1, -1, 10, 6, world)
mine = shards[rank]
t = torch.tensor([float(len(mine.owned)), 10.0 + rank], dtype=torch.float64)
total = t.clone(); dist.all_reduce(total, op=dist.ReduceOp.SUM)
mx = t.clone(); dist.all_reduce(mx, op=dist.ReduceOp.MAX)
dist.barrier()
if rank == 0:
    assert int(total[0].item()) == 8 * 4, total
    assert mx[1].item() == 10.0 + world - 1
    print("ok", world)
dist.destroy_process_group()
"""


def test_gloo_world_size_2(tmp_path):
    script = tmp_path / "g.py"
    script.write_text(_GLOO)
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", OMP_NUM_THREADS="1")
    out = subprocess.run(
        [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
         "--master-addr", "127.0.0.1", "--master-port", "29533", str(script),
         os.path.join(ROOT, "voxelengine-cpp_b200")],
        env=env, capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr[-2000:]
    assert "ok 2" in out.stdout
