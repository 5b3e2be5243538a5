"""world_size-2 gloo test of the N>1 host logic (independent replicas + counter reduction)."""
import os
import socket
import sys

import pytest
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, total_units, q):
    import torch.distributed as dist

    sys.path.insert(0, ROOT)
    import __graft_entry__ as graft

    rep = graft.load_package().replica
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    info = rep.RankInfo.from_env()
    mine = rep.shard_units(total_units, info.rank, info.world)
    counters = {"evals": 256 * len(mine), "batches": len(mine), "legal_moves": sum(mine)}
    rep.barrier()
    total, t_max = rep.aggregate(counters, elapsed_ms=10.0 * (rank + 1))
    q.put((rank, list(mine), total, t_max))
    dist.destroy_process_group()


def test_two_rank_shards_and_aggregation():
    world, total_units = 2, 11
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, total_units, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    units = sorted(res[0][1] + res[1][1])
    assert units == list(range(total_units))                  # disjoint cover, no exchange needed
    assert set(res[0][1]).isdisjoint(res[1][1])
    for _, _, total, t_max in res:
        assert total["batches"] == total_units and total["evals"] == 256 * total_units
        assert total["legal_moves"] == sum(range(total_units))
        assert t_max == 20.0                                    # max over ranks


def test_shard_counts_single_process():
    sys.path.insert(0, ROOT)
    import __graft_entry__ as graft

    rep = graft.load_package().replica
    assert rep.shard_counts(10, 4) == [3, 3, 2, 2]
    assert list(rep.shard_units(5, 0, 1)) == [0, 1, 2, 3, 4]
    total, t = rep.aggregate({"evals": 3}, 1.5)
    assert total["evals"] == 3 and t == 1.5
    with pytest.raises(ValueError):
        rep.shard_units(4, 2, 2)
    assert rep.whole_job_rate(1000, 500.0) == 2000.0
