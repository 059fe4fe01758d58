"""CPU tier: the N>1 host path (chunk sharding + ordered gather of .alc blobs) with world_size 2 over gloo.
The ranks produce their blobs with the CPU oracle (test stand-in for the per-GPU encoder; the CUDA encoder is
compared with the same oracle blobs in test_gpu_parity.py)."""
import os
import socket

import numpy as np
import pytest
import torch.multiprocessing as mp

import oracle as O
from __graft_entry__ import load_package

pkg = load_package()
from alice_codec_b200 import sharding  # noqa: E402

W, H, F, Q, WV, N_CHUNKS = 24, 12, 4, 80, 1, 5


def _blob(c):
    return O.encode(O.generate(O.G1, W, H, F, O.SEED + c), W, H, F, Q, WV)


def _worker(rank, world, port, q):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    mine = sharding.chunks_of_rank(N_CHUNKS, rank, world)
    out = sharding.gather_stream([_blob(c) for c in mine], N_CHUNKS, rank, world)
    dist.barrier()
    if rank == 0:
        q.put(sharding.concat_stream(out))
    else:
        assert out is None
    dist.destroy_process_group()


def test_partition_is_a_cover():
    for n in (0, 1, 5, 8, 17):
        for world in (1, 2, 4, 8):
            seen = sorted(c for r in range(world) for c in sharding.chunks_of_rank(n, r, world))
            assert seen == list(range(n))
    with pytest.raises(ValueError):
        sharding.chunks_of_rank(4, 2, 2)


def test_stream_split_roundtrip():
    blobs = [_blob(c) for c in range(3)] + [O.encode(np.zeros(0, np.uint8), 0, 0, 0, 90, 0)]
    assert sharding.split_stream(sharding.concat_stream(blobs)) == blobs
    with pytest.raises(ValueError):
        sharding.split_stream(b"nope" + blobs[0])


def test_two_rank_gather_matches_single_rank():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    stream = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    expected = sharding.concat_stream([_blob(c) for c in range(N_CHUNKS)])
    assert stream == expected
    assert len(sharding.split_stream(stream)) == N_CHUNKS
