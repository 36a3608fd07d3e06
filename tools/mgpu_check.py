"""Multi-GPU parity check, one process per GPU (launch with torchrun):

  python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
      tools/mgpu_check.py [nx ny nz] [steps]

Every rank runs its z-slab of a TGV on the cuda_c backend (NCCL halo / reduced-row exchange, all-to-all FFT,
all-reduced monitors); rank 0 runs the oracle's in-process P-rank emulation and compares slabs.
"""
import ctypes
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import torch
import torch.distributed as dist

import x3d2_b200 as X


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ.get("LOCAL_RANK", 0))
    args = [int(a) for a in sys.argv[1:] if not a.startswith("--")]
    dims = tuple(args[:3]) if len(args) >= 3 else (64, 64, 64 * world)
    steps = args[3] if len(args) >= 4 else 2
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    c, _ = X.load()
    ok = True
    modes = (False,) if "--fast-only" in sys.argv else (False, True)
    rng = np.random.default_rng(7)
    gf = [rng.standard_normal((dims[2], dims[1], dims[0])) for _ in range(3)]
    exp = rmon = None
    if rank == 0:  # the oracle's P-rank emulation, once for both modes
        import _oracle as O
        ref = O.World(dims, nproc_dir=(1, 1, world))
        ref.init_tgv()
        ref.step(steps)
        exp = list(ref.get_uvw()) + [ref.tds_solve(3, "der1st", gf[0])] + list(ref.transeq_dir(3, *gf)) + \
              [ref.divergence(*gf), ref.poisson(gf[0] - gf[0].mean())]
        rmon = ref.monitor()
        del ref
    for strict in modes:
        buf = [None]  # one ncclUniqueId per communicator
        if rank == 0:
            raw = ctypes.create_string_buffer(128)
            assert c.x3d2c_nccl_unique_id(raw) == 0, c.x3d2c_last_error()
            buf = [raw.raw]
        dist.broadcast_object_list(buf, src=0)
        sim = X.Sim(dims, nproc_dir=(1, 1, world), rank=rank, nproc=world, device=local, strict=strict,
                    nccl_unique_id=buf[0])
        sim.init_tgv()
        sim.step(steps)
        u, v, w = sim.get_uvw()
        mon = sim.monitor()
        # single operators on seeded data (each rank generates the global field and takes its slab)
        nzl = dims[2] // world
        sl = slice(rank * nzl, (rank + 1) * nzl)
        loc = [g[sl] for g in gf]
        tz = sim.tds_solve(3, "der1st", loc[0])
        tq = sim.transeq_dir(3, *loc)
        dv = sim.divergence(*loc)
        po = sim.poisson(loc[0] - gf[0].mean())
        parts = [u, v, w, tz, *tq, dv, po]
        gathered = [None] * world
        dist.all_gather_object(gathered, parts)
        if rank == 0:
            names = ["u", "v", "w", "tds_z", "transeq_z_du", "transeq_z_dv", "transeq_z_dw", "div", "poisson"]
            # velocities and the three transeq outputs are compared relative to the scale of the VECTOR (north_star:
            # "velocity fields within 1e-12 relative"): w of the Taylor-Green vortex starts at zero and stays orders of
            # magnitude below u, v, so its own maximum is not a meaningful scale; the per-field ratio is printed too
            scale = {0: max(np.abs(exp[i]).max() for i in (0, 1, 2)), 4: max(np.abs(exp[i]).max() for i in (4, 5, 6))}
            for i, name in enumerate(names):
                got = np.concatenate([g[i] for g in gathered], axis=0)
                own = np.abs(exp[i]).max()
                sc = scale[0] if i < 3 else (scale[4] if 4 <= i <= 6 else own)
                err = np.abs(got - exp[i]).max() / sc
                exact = np.array_equal(got, exp[i])
                print(f"[mgpu P={world} strict={strict}] {name:14s} rel err {err:.3e} (vs own max {np.abs(got - exp[i]).max() / own:.3e}) "
                      f"bit-exact={exact}", flush=True)
                ok &= err < 1e-12
            for k in ("enstrophy", "ke"):
                e = abs(mon[k] - rmon[k]) / rmon[k]
                print(f"[mgpu P={world} strict={strict}] {k:14s} rel err {e:.3e}")
                ok &= e < 1e-10
        sim.close()
    flag = torch.tensor([1 if ok else 0], device="cuda")
    dist.broadcast(flag, src=0)
    dist.destroy_process_group()
    if rank == 0:
        print("MGPU_CHECK", "PASS" if ok else "FAIL")
    sys.exit(0 if int(flag.item()) else 1)


if __name__ == "__main__":
    main()
