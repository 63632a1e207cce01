"""world_size-2 gloo test of the N>1 host logic of bench.py (replica scaling: independent streams, job time =
max over ranks, counters add, value = all ranks' frames / slowest rank).  CPU only."""
import os
import socket
import sys

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    import bench

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    times = [100.0 + 50.0 * rank, 200.0 - 30.0 * rank]  # rank 1 is slower on the first, faster on the second
    (t0, t1), (launches,) = bench.reduce_over_ranks(times, [1000 + rank], world, torch.device("cpu"))
    dist.barrier()
    q.put((rank, t0, t1, launches, bench.aggregate_fps(world, 20, t0)))
    dist.destroy_process_group()


def test_replica_aggregation_world2():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, t0, t1, launches, fps in res:
        assert t0 == 150.0 and t1 == 200.0      # max over ranks
        assert launches == 2001                  # sum over ranks
        assert abs(fps - 2 * 20 / 0.150) < 1e-9  # whole-job frames / slowest rank


def test_single_rank_passthrough_and_frame_window():
    sys.path.insert(0, ROOT)
    import bench

    assert bench.reduce_over_ranks([1.5, 2.5], [7], 1, torch.device("cpu")) == ([1.5, 2.5], [7])
    assert bench.stream_frame_index(1) == (0, 1, 2)
    assert bench.stream_frame_index(bench.NFRAMES) == (bench.NFRAMES - 1, 0, 1)
    # algorithmic custom-op bytes per direction: light 10.6 (Correlation) + 6.7 (Warp) MB,
    # dense-4K 462.5 + 239.5 MB
    assert abs(bench.op_bytes(bench.WORKLOADS["1080p-light"]) / 2 / 1e6 - (10.57 + 6.73)) < 0.1
    assert abs(bench.op_bytes(bench.WORKLOADS["4k-dense"]) / 2 / 1e6 - (462.5 + 239.5)) < 1.0
