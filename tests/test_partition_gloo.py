"""world_size-2 gloo test of the multi-GPU host logic (no GPU): every rank takes its round-robin
share of the frame batch, there is no data-path collective, and the only cross-rank operations are
the barrier and the MAX-over-ranks of the step time that bench.py reports."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from librempeg_b200 import partition as P


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, n_frames, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    mine = P.frames_for_rank(n_frames, rank, world)
    # each rank "processes" its frames independently: checksum of its frame indices
    local = torch.tensor([len(mine), sum(mine)], dtype=torch.int64)
    gathered = [torch.zeros(2, dtype=torch.int64) for _ in range(world)]
    dist.barrier()
    dist.all_gather(gathered, local)           # test-only: the product path has no collective
    t = torch.tensor([10.0 + rank], dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)   # bench.py: max over ranks of the device time
    q.put((rank, mine, [g.tolist() for g in gathered], float(t.item())))
    dist.barrier()
    dist.destroy_process_group()


def test_round_robin_partition_world2():
    world, n_frames = 2, 129
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_frames, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    frames = sorted(res[0][1] + res[1][1])
    assert frames == list(range(n_frames))                     # disjoint cover
    assert not set(res[0][1]) & set(res[1][1])
    assert res[0][2] == res[1][2] == [[65, sum(range(0, 129, 2))], [64, sum(range(1, 129, 2))]]
    assert res[0][3] == res[1][3] == 11.0                       # MAX over ranks
