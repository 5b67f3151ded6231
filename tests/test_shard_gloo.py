"""CPU, world_size 2 over gloo: the multi-GPU host logic (frame blocks, size all-gather, global offsets).  The encoder on
each rank is the oracle -- this test is about the sharding, not the kernels: the concatenation of the ranks' streams at
the gathered offsets must equal what one process produces for the whole batch."""
import os
import socket
import tempfile

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from imcvt_b200 import shard, synth
from tests._util import oracle_jls


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n_total, outdir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    first, count = shard.frame_block(rank, world, n_total)
    frames = synth.batch("texture", count, 40, 56, first=first) if count else np.zeros((0, 40, 56), np.uint8)
    streams = [oracle_jls(f, 0) for f in frames]
    sizes = shard.gather_sizes(torch.tensor([len(s) for s in streams], dtype=torch.int64), n_total, world)
    offs = shard.global_offsets(sizes)
    total = int(sizes.sum())
    path = os.path.join(outdir, "all.bin")
    if rank == 0:
        with open(path, "wb") as f:
            f.truncate(total)
    dist.barrier()
    with open(path, "r+b") as f:
        for i, s in enumerate(streams):
            f.seek(int(offs[first + i]))
            f.write(s)
    dist.barrier()
    np.save(os.path.join(outdir, f"sizes{rank}.npy"), sizes.numpy())
    dist.destroy_process_group()


def test_two_ranks_produce_the_single_process_output():
    n_total = 7                                     # odd: the last rank's block is short
    with tempfile.TemporaryDirectory() as d:
        mp.spawn(_worker, args=(2, _free_port(), n_total, d), nprocs=2, join=True)
        want = [oracle_jls(f, 0) for f in synth.batch("texture", n_total, 40, 56)]
        got = open(os.path.join(d, "all.bin"), "rb").read()
        assert got == b"".join(want)
        s0, s1 = np.load(os.path.join(d, "sizes0.npy")), np.load(os.path.join(d, "sizes1.npy"))
        assert list(s0) == [len(w) for w in want] and list(s1) == list(s0)


def test_frame_blocks_cover_the_batch_once():
    for world in (1, 2, 3, 4, 8):
        for n in (1, 5, 8, 64, 257):
            seen = []
            for r in range(world):
                first, count = shard.frame_block(r, world, n)
                seen += list(range(first, first + count))
            assert seen == list(range(n))
