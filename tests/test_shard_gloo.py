"""CPU test of the N>1 path (world_size 2, gloo): stream sharding is a partition, shards decode independently to the same
results as the unsharded batch (checked with the oracle standing in for the per-rank decoder), and timings reduce with MAX."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import load_golden


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from opus_codec_b200.shard import stream_range, max_over_ranks, aggregate_throughput
    from oracle import oraclepy
    g = load_golden("cfg2_mono_20ms_64k_cbr")
    S = g["packets"].shape[0]
    first, count = stream_range(rank, world, S)
    counts = [torch.zeros(1, dtype=torch.int64) for _ in range(world)]
    dist.all_gather(counts, torch.tensor([count]))
    assert sum(int(c) for c in counts) == S
    rng = np.stack([oraclepy.decode_stream(g["packets"][s], g["lens"][s], 960, 1)[1] for s in range(first, first + count)])
    np.save(os.path.join(out_dir, "rng_%d.npy" % rank), rng)
    dist.barrier()
    ms = max_over_ranks(10.0 + 5.0 * rank)            # rank 1 is "slower": MAX must win on every rank
    assert ms == 15.0
    assert aggregate_throughput(100, world, ms) == pytest.approx(100 * world / 0.015)
    dist.destroy_process_group()


def test_two_rank_sharding(tmp_path):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    g = load_golden("cfg2_mono_20ms_64k_cbr")
    got = np.concatenate([np.load(tmp_path / ("rng_%d.npy" % r)) for r in range(world)])
    assert (got == g["dec_rng"]).all()


def test_stream_range_is_a_partition():
    from opus_codec_b200.shard import stream_range
    for S in (1, 7, 4096, 131072, 131073):
        for W in (1, 2, 3, 4, 8):
            nxt = 0
            for r in range(W):
                f, c = stream_range(r, W, S)
                assert f == nxt and c >= 0
                nxt = f + c
            assert nxt == S
