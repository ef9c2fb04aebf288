"""world_size-2 (and 3) gloo runs on CPU of the multi-GPU host logic: row-block sharding, the rendezvous plumbing
bench.py uses (unique-id broadcast, max/sum over ranks). No GPU, no NCCL."""
import os
import subprocess
import sys
import textwrap

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = textwrap.dedent('''
    import os, sys
    sys.path.insert(0, os.path.join(%(root)r, "blackhole-simulation_b200"))
    sys.path.insert(0, %(root)r)
    import torch, torch.distributed as dist
    from gravitas_b200 import shard
    import bench
    dist.init_process_group("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    for H, W in ((2160, 3840), (4320, 7680), (83, 157), (5, 7), (1, 3)):
        for taa in (False, True):
            mine = (shard.shard_rows(H, rank, world), shard.traced_rows(H, rank, world, taa))
            allr = [None] * world
            dist.all_gather_object(allr, mine)
            own = [a[0] for a in allr]
            assert own[0][0] == 0 and own[-1][1] == H, own
            for a, b in zip(own, own[1:]):
                assert a[1] == b[0], own                      # blocks tile [0, H) with no gap or overlap
            rpr = shard.rows_per_rank(H, world)
            assert all(e - b <= rpr for b, e in own)
            il = [None] * world
            dist.all_gather_object(il, shard.interleaved_rows(H, rank, world, taa))
            assert sorted(sum(il, [])) == list(range(H))          # GVT_FLAG_ROW_INTERLEAVE: a partition of the rows too
            assert max(len(x) for x in il) - min(len(x) for x in il) <= (16 if taa else 1)
            cnt, padded = shard.gather_counts(W, H, world)
            assert cnt == rpr * W * 4 and padded >= H and padded - H < world
            for (b, e), (tb, te) in allr:
                if taa and e > b:
                    assert tb == max(0, b - 1) and te == min(H, e + 1)    # one halo row on interior edges
                else:
                    assert (tb, te) == (b, e)
    # the plumbing bench.py uses
    objs = [bytes(range(128)) if rank == 0 else None]
    dist.broadcast_object_list(objs, src=0)
    assert objs[0] == bytes(range(128))
    assert bench.allreduce_max(dist, float(rank + 1)) == float(world)
    assert bench.allreduce_sum(dist, float(rank + 1)) == world * (world + 1) / 2
    bench.barrier(dist)
    dist.destroy_process_group()
    print("rank", rank, "ok")
''')


@pytest.mark.parametrize("world", [2, 3])
def test_sharding_and_plumbing_under_gloo(tmp_path, world):
    script = tmp_path / "worker.py"
    script.write_text(WORKER % {"root": ROOT})
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
           "--master-addr", "127.0.0.1", "--master-port", str(29650 + world), str(script)]
    p = subprocess.run(cmd, capture_output=True, text=True, timeout=300, env=env)
    assert p.returncode == 0, p.stdout[-2000:] + p.stderr[-4000:]
    assert p.stdout.count("ok") == world
