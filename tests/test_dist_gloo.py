"""The N > 1 path on CPU: two gloo ranks shard a batch image-wise (no data-path collective), run
their shard through a stand-in for the per-image core, and the bookkeeping used by bench.py
(units of the whole job, max-over-ranks time, optional gather of rows) comes out right."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from detrpose_b200 import shard, synthetic
from oracle import msda_torch as otorch


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, n_images, out_path):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        w = synthetic.WORKLOADS["detrpose_n"]
        inp = synthetic.make_inputs(n_images, 6, w["H"], w["Dh"], ((8, 8), (4, 4)), w["P"], seed=5)
        shapes = inp["shapes"]
        mem, loc, att = shard.shard_batch([inp["memory"], inp["locations"], inp["attention"]], rank, world)
        start, stop = shard.image_range(n_images, rank, world)
        assert mem.shape[0] == stop - start == shard.local_images(n_images, rank, world)
        # every image is independent: a rank computes exactly its own rows (CPU stand-in for the kernels)
        local = otorch.core(otorch.make_value_list(mem, w["H"], shapes), shapes, loc, att).contiguous() \
            if mem.shape[0] else torch.zeros((0, 6, w["H"] * w["Dh"]))
        full = shard.gather_rows(local, n_images)
        total = shard.job_total(float(local.shape[0]))
        slowest = shard.max_over_ranks(1.0 + rank)
        if rank == 0:
            ref = otorch.core(otorch.make_value_list(inp["memory"], w["H"], shapes), shapes,
                              inp["locations"], inp["attention"])
            torch.save({"ok": torch.allclose(full, ref, atol=1e-6), "total": total, "slowest": slowest,
                        "rows": full.shape[0]}, out_path)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n_images", [5, 8])
def test_two_rank_image_sharding(tmp_path, n_images):
    out_path = str(tmp_path / "result.pt")
    port = _free_port()
    mp.spawn(_worker, args=(2, port, n_images, out_path), nprocs=2, join=True)
    res = torch.load(out_path)
    assert res["ok"]
    assert res["total"] == float(n_images)       # units all ranks processed
    assert res["slowest"] == 2.0                 # max over ranks, as bench.py times a step
    assert res["rows"] == n_images
