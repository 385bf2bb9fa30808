"""Host-side plumbing of the slab decomposition (one process per GPU, torch.distributed for rendezvous only).

The data path never goes through torch: particles are sharded on the host, uploaded through the C ABI, and
all per-step exchanges (halo layers, migrants, dt, bbox) are NCCL calls inside libmlh_gpu.so (csrc/halo.cu).
torch.distributed is used to broadcast the 128-byte NCCL unique id and for barriers in bench/tests.
"""
import ctypes as C

import numpy as np

from . import capi

DBL_MIN = np.finfo(np.float64).tiny


def domain_limits(ic):
    """Particles::getDomainLimits (Particles.cpp:228-267) incl. quirk Q8 for the common case: minimum over all
    particles, maximum over original index >= 1 starting from DBL_MIN (quirk Q2)."""
    D = ic["dim"]
    lo, hi = [], []
    for k in ("x", "y", "z")[:D]:
        a = ic[k]
        lo.append(float(a.min()))
        hi.append(max(float(a[1:].max()), DBL_MIN))
    return np.array(lo + hi)


def search_grid(ic):
    """Domain::createGrid (Domain.cpp:9-54): cells per axis and cell sizes of the global search grid."""
    D = ic["dim"]
    box = np.asarray(ic["box"], dtype=np.float64) if ic.get("periodic") else domain_limits(ic)
    bmin, bmax = box[:D], box[D:2 * D]
    cells = np.floor((bmax - bmin) / ic["h"]).astype(np.int64)
    size = (bmax - bmin) / cells
    return bmin, bmax, cells, size


def slab_range(n_layers, nranks, rank):
    lib = capi.load_library()
    lo, hi = C.c_int(), C.c_int()
    rc = lib.mlh_slab_range(n_layers, nranks, rank, C.byref(lo), C.byref(hi))
    if rc != 0:
        raise capi.MlhError("mlh_slab_range(%d, %d, %d) failed" % (n_layers, nranks, rank))
    return lo.value, hi.value


def particle_layers(ic):
    """global cell layer of every particle along the slab axis (slowest-varying cell axis = last dimension),
    with the reference's cell formula (Particles.cpp:279-295)"""
    D = ic["dim"]
    bmin, bmax, cells, size = search_grid(ic)
    k = D - 1
    coord = ic[("x", "y", "z")[k]]
    layer = np.floor((coord - bmin[k]) / size[k]).astype(np.int64)
    layer[layer == cells[k]] -= 1
    return layer, int(cells[k])


def shard(ic, rank, nranks):
    """the particles rank `rank` owns: (local ic dict, global original ids)"""
    layer, n_layers = particle_layers(ic)
    if n_layers < 2 * nranks:
        raise capi.MlhError("slab decomposition needs >= 2 cell layers per rank (%d layers, %d ranks)" % (n_layers, nranks))
    lo, hi = slab_range(n_layers, nranks, rank)
    mine = np.nonzero((layer >= lo) & (layer < hi))[0]
    local = dict(ic)
    for key in ("x", "y", "z", "vx", "vy", "vz", "m", "u"):
        if ic.get(key) is not None:
            local[key] = np.ascontiguousarray(ic[key][mine])
    return local, mine.astype(np.int32)


def broadcast_unique_id(dist, rank, make_id=None):
    """rank 0 creates the NCCL unique id inside the library; everybody gets the 128 bytes"""
    import torch
    buf = C.create_string_buffer(128)
    if rank == 0:
        if make_id is not None:
            buf.raw = make_id()
        else:
            rc = capi.load_library().mlh_comm_unique_id(buf)
            if rc != 0:
                raise capi.MlhError("mlh_comm_unique_id failed (%d)" % rc)
    t = torch.frombuffer(bytearray(buf.raw), dtype=torch.uint8).clone()
    if dist.get_backend() == "nccl":
        t = t.cuda()
    dist.broadcast(t, src=0)
    return bytes(t.cpu().numpy().tobytes())


def create_sharded(cfg, ic, dist, capacity_factor=1.6):
    """Create this rank's context, join the library's communicator and upload the owned particles."""
    rank, world = dist.get_rank(), dist.get_world_size()
    cfg.rank, cfg.nranks = rank, world
    local, ids = shard(ic, rank, world)
    n_max = int(len(ic["x"]) / world * capacity_factor) + 8192
    cfg.capacity = max(n_max, int(len(ids) * capacity_factor) + 8192)
    gpu = capi.MfvGpu(cfg)
    uid = broadcast_unique_id(dist, rank)
    gpu._check(gpu.lib.mlh_comm_init(gpu.ctx, uid), "mlh_comm_init")
    gpu.upload(local, ids=ids)
    return gpu, local
