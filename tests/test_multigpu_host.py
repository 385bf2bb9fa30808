"""CPU, world_size 2 over gloo: the host-side logic of the slab decomposition (meshlesshydro_b200/multigpu.py):
every particle is owned by exactly one rank, owners follow the cell-layer ranges of mlh_slab_range, the
unique-id broadcast delivers rank 0's 128 bytes, and max-over-ranks timing reduction works."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from meshlesshydro_b200 import ic as IC, multigpu


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    results = {}
    for name, ic in (("kh", IC.kelvin_helmholtz(40, lattice=False)), ("sedov", IC.sedov(16)), ("fb", IC.fluid_block(30, jitter=0.05))):
        local, ids = multigpu.shard(ic, rank, world)
        layer, n_layers = multigpu.particle_layers(ic)
        lo, hi = multigpu.slab_range(n_layers, world, rank)
        assert np.all((layer[ids] >= lo) & (layer[ids] < hi))
        assert np.array_equal(local["x"], ic["x"][ids])
        gathered = [None] * world
        dist.all_gather_object(gathered, ids)
        allids = np.concatenate(gathered)
        assert len(allids) == len(ic["x"]) and np.array_equal(np.sort(allids), np.arange(len(ic["x"])))
        results[name] = len(ids)
    uid = multigpu.broadcast_unique_id(dist, rank, make_id=lambda: bytes(range(128)))
    assert uid == bytes(range(128))
    t = torch.tensor([float(rank + 1)], dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    assert t.item() == world
    np.save(os.path.join(out_dir, "rank%d.npy" % rank), np.array(list(results.values())))
    dist.destroy_process_group()


def test_sharding_world2(tmp_path):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    counts = sum(np.load(tmp_path / ("rank%d.npy" % r)) for r in range(world))
    assert list(counts) == [1600, 4096, 900]


def test_grid_matches_reference_formula():
    """search_grid == Domain::createGrid: cells = floor(L/h), size = L/cells (Domain.cpp:10-22)"""
    ic = IC.kelvin_helmholtz(100)
    bmin, bmax, cells, size = multigpu.search_grid(ic)
    assert list(cells) == [25, 25] and np.allclose(size, 0.04)
    ic = IC.sedov(21)
    bmin, bmax, cells, size = multigpu.search_grid(ic)
    assert np.all(size >= ic["h"]) and np.all(cells == np.floor((bmax - bmin) / ic["h"]))


def test_too_few_layers_rejected():
    from meshlesshydro_b200 import capi
    with pytest.raises(capi.MlhError):
        multigpu.shard(IC.kelvin_helmholtz(20, lattice=False, h_over_dx=4.0), 0, 8)  # 5 layers, 8 ranks
