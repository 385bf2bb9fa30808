"""torchrun worker (2+ GPUs): run the sharded MFV step and compare with the single-GPU path on rank 0.
Usage: torchrun --nproc-per-node N tests/mgpu_worker.py <case> <steps>"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from meshlesshydro_b200 import capi, multigpu, ic as IC  # noqa: E402

CASES = {
    "kh": (lambda: IC.kelvin_helmholtz(64, lattice=False), "kh2d"),
    "kh_big": (lambda: IC.kelvin_helmholtz(300, lattice=True, jitter=0.2), "kh2d"),
    "sedov": (lambda: IC.sedov(32), "sedov3d"),
    "fb": (lambda: IC.fluid_block(80, jitter=0.05), "fb2d"),
}


def main():
    case, steps = sys.argv[1], int(sys.argv[2])
    rank, world, local_rank = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local_rank)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    factory, preset = CASES[case]
    ic = factory()
    D = ic["dim"]
    cfg = capi.make_config(preset, ic["h"], ic["gamma"], ic.get("box"), abs_mode=capi.ABS_INT_TRUNC, max_interactions=160)
    cfg.device = local_rank
    gpu, local = multigpu.create_sharded(cfg, ic, dist)
    s0 = gpu.sums()
    dts = [gpu.step() for _ in range(steps)]
    s1 = gpu.sums()
    flags = gpu.error_flags()
    st = gpu.download_state()
    names = ["x", "y", "vx", "vy", "m", "u"] + (["z", "vz"] if D == 3 else [])
    gathered = [None] * world
    dist.all_gather_object(gathered, {k: st[k] for k in names + ["ids"]})
    ok = True
    if rank == 0:
        N = len(ic["x"])
        full = {k: np.full(N, np.nan) for k in names}
        count = np.zeros(N, dtype=np.int64)
        for part in gathered:
            count[part["ids"]] += 1
            for k in names:
                full[k][part["ids"]] = part[k]
        assert np.all(count == 1), "ownership is not a partition after %d steps" % steps
        cfg1 = capi.make_config(preset, ic["h"], ic["gamma"], ic.get("box"), abs_mode=capi.ABS_INT_TRUNC, max_interactions=160)
        cfg1.device = local_rank
        one = capi.MfvGpu(cfg1)
        one.upload(ic)
        dts1 = [one.step() for _ in range(steps)]
        ref = one.download_state()
        worst = 0.0
        for k in names:
            d = np.max(np.abs(full[k] - ref[k]))
            worst = max(worst, d)
            if not np.array_equal(full[k], ref[k]):
                print("MISMATCH %s: max abs diff %.3e (%d values)" % (k, d, int((full[k] != ref[k]).sum())))
                ok = False
        if dts != dts1:
            print("MISMATCH dt:", dts, dts1)
            ok = False
        print("case=%s world=%d steps=%d flags=%d sums drift M %.2e E %.2e bitwise_equal=%s worst=%.2e"
              % (case, world, steps, flags, abs(s1[1] - s0[1]) / s0[1], abs(s1[2] - s0[2]) / s0[2], ok, worst))
    okt = torch.tensor([1 if ok else 0], device="cuda")
    dist.broadcast(okt, src=0)
    gpu.close()
    dist.destroy_process_group()
    sys.exit(0 if int(okt.item()) == 1 and flags == 0 else 1)


if __name__ == "__main__":
    main()
