"""World-size-2 (gloo, CPU) test of the N>1 launch: sessions are partitioned per rank with NO data-path collective; the
only collectives are the timing barrier and the max-over-ranks reduce that bench.py performs."""
import os
import socket
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, n_streams, q):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import asr_streaming_b200 as A
    from asr_streaming_b200.scheduler import partition_streams
    from test_host_logic import FakeEngine
    cfg = A.ModelConfig(max_batch=64, max_sessions=64)
    mine = partition_streams(n_streams, world, rank)
    sch = A.SessionScheduler(FakeEngine(cfg))
    sess = {gid: sch.open() for gid in mine}
    for gid, s in sess.items():                                   # stream gid's audio is a function of gid only
        s.accept_waveform(np.full(10240, gid % 1000, np.int16))
    dist.barrier()
    out = sch.tick()
    elapsed = torch.tensor([0.001 * (rank + 1)], dtype=torch.float64)
    dist.all_reduce(elapsed, op=dist.ReduceOp.MAX)                # the bench's max-over-ranks timing reduce
    tokens = {gid: s.tokens for gid, s in sess.items()}
    q.put((rank, sorted(mine), tokens, float(elapsed.item()), len(out)))
    dist.barrier()
    dist.destroy_process_group()


def test_two_ranks_partition_sessions_without_exchange():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    n_streams = 37
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n_streams, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    res.sort()
    ids = res[0][1] + res[1][1]
    assert sorted(ids) == list(range(n_streams)) and abs(len(res[0][1]) - len(res[1][1])) <= 1
    assert res[0][3] == res[1][3] == 0.002                        # max over ranks
    assert res[0][4] + res[1][4] == n_streams
    # per-stream result depends only on the stream (no cross-rank / cross-stream term): compare with a 1-rank run
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import asr_streaming_b200 as A
    from test_host_logic import FakeEngine
    sch = A.SessionScheduler(FakeEngine(A.ModelConfig(max_batch=64, max_sessions=64)))
    single = {}
    for gid in range(n_streams):
        se = sch.open()
        se.accept_waveform(np.full(10240, gid % 1000, np.int16))
        single[gid] = se
    sch.tick()
    for r in res:
        for gid, toks in r[2].items():
            assert toks == single[gid].tokens
