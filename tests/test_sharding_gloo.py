"""N>1 host logic on CPU: two gloo ranks shard a list of streams, run Phase A (CPU oracle as the
stand-in for the device call: the sharding code is what is under test), and the union of the
shards must equal the single-process result; the reporting reduction must count every block once."""
import os
import socket
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import ROOT, load_setup
from vorbis_b200 import abi, shard


def test_stream_slice_partitions_exactly():
    for n in (0, 1, 7, 8, 10000):
        for world in (1, 2, 3, 8):
            got = []
            for r in range(world):
                lo, hi = shard.stream_slice(n, world, r)
                assert 0 <= lo <= hi <= n
                got.extend(range(lo, hi))
            assert got == list(range(n))
            sizes = [shard.stream_slice(n, world, r)[1] - shard.stream_slice(n, world, r)[0] for r in range(world)]
            assert max(sizes) - min(sizes) <= 1


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out_dir):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import pyoracle
    setup = load_setup("44k_stereo_q5")
    o = pyoracle.Oracle(setup)
    nstreams, bps, W = 5, 3, 0            # short blocks keep it quick; 5 streams do not divide by 2
    N, ch = setup.blocksize(W), setup.channels
    rng = np.random.default_rng(2024)
    pcm = rng.uniform(-0.5, 0.5, (nstreams, bps, ch, N)).astype(np.float32)
    lo, hi = shard.stream_slice(nstreams, world, rank)
    desc = np.zeros((hi - lo) * bps, abi.BLOCKDESC_DTYPE)
    res = o.phaseA(W, pcm[lo:hi].reshape(-1, ch, N), desc, streams=(hi - lo, bps)) if hi > lo else None
    total, ms = shard.reduce_report(dist, (hi - lo) * bps, 10.0 + rank)
    assert total == nstreams * bps and ms == 10.0 + world - 1
    if res is not None:
        np.savez(os.path.join(out_dir, "rank%d.npz" % rank), lo=lo, hi=hi, logmask=res["logmask"],
                 ampmax=res["ampmax_out"])
    dist.barrier()
    dist.destroy_process_group()


def test_two_ranks_cover_the_job(tmp_path, oracle_lib):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    setup = load_setup("44k_stereo_q5")
    o = oracle_lib.Oracle(setup)
    nstreams, bps, W = 5, 3, 0
    N, ch = setup.blocksize(W), setup.channels
    rng = np.random.default_rng(2024)
    pcm = rng.uniform(-0.5, 0.5, (nstreams, bps, ch, N)).astype(np.float32)
    whole = o.phaseA(W, pcm.reshape(-1, ch, N), np.zeros(nstreams * bps, abi.BLOCKDESC_DTYPE),
                     streams=(nstreams, bps))
    seen = 0
    for r in range(world):
        z = np.load(os.path.join(str(tmp_path), "rank%d.npz" % r))
        lo, hi = int(z["lo"]), int(z["hi"])
        assert np.array_equal(z["logmask"].view(np.uint32),
                              whole["logmask"][lo * bps:hi * bps].view(np.uint32))
        assert np.array_equal(z["ampmax"], whole["ampmax_out"][lo * bps:hi * bps])
        seen += hi - lo
    assert seen == nstreams
