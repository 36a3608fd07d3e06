"""Per-operator device timing (CUDA events on the backend's stream) against the algorithmic bytes of SURVEY.md §8d.

usage: python tools/opbench.py [N | nx,ny,nz] [--strict] [--channel] [--ops=transeq_z,poisson,...]
       --channel: walls in y (Dirichlet), y stretched 'top-bottom' with beta = 0.259065151, L = (4, 2, 2), Re = 4200
                  (examples/channel/input.x3d; BASELINE.json configs[4] is 512,257,512)
       python -m torch.distributed.run --nnodes=1 --nproc-per-node P --master-addr 127.0.0.1 tools/opbench.py [N]
           (P ranks, z-slabs of N^3 points each: the z operators run the rank-split kernels + NCCL exchanges)
"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import x3d2_b200 as X

BYTES_PER_PT = {"transeq_x": 48, "transeq_y": 48, "transeq_z": 48, "tds_solve_x": 16, "tds_solve_y": 16, "tds_solve_z": 16,
                "reorder_x2y": 16, "reorder_x2z": 16, "reorder_y2x": 16, "reorder_y2z": 16, "reorder_z2y": 16,
                "reorder_z2c": 16, "reorder_c2z": 16, "reorder_z2x": 16, "sum_yintox": 24, "sum_zintox": 24, "vecadd": 24,
                "veccopy": 16, "poisson": 152, "pressure_correction": 256 + 152 + 208 + 72, "transeq": 384}


def time_op(sim, op, reps, stream):
    sim.bench_op(op, 2)
    sim.sync()
    with torch.cuda.stream(stream):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        sim.bench_op(op, reps)
        e1.record(stream)
    e1.synchronize()
    return e0.elapsed_time(e1) / reps


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 and sys.argv[1].isdigit() else 256
    grid = [int(a) for a in sys.argv[1].split(",")] if len(sys.argv) > 1 and "," in sys.argv[1] else [n, n, n]
    strict = "--strict" in sys.argv
    peak = 6541.8
    try:
        peak = json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))["hbm_gbs"]
    except Exception:
        pass
    world, rank = int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("RANK", 0))
    if world > 1:
        import ctypes
        import torch.distributed as dist
        local = int(os.environ.get("LOCAL_RANK", 0))
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        buf = [None]
        if rank == 0:
            raw = ctypes.create_string_buffer(128)
            assert X.load()[0].x3d2c_nccl_unique_id(raw) == 0
            buf = [raw.raw]
        dist.broadcast_object_list(buf, src=0)
        sim = X.Sim((grid[0], grid[1], grid[2] * world), nproc_dir=(1, 1, world), rank=rank, nproc=world, device=local, strict=strict,
                    nccl_unique_id=buf[0])
    elif "--channel" in sys.argv:
        sim = X.Sim(tuple(grid), strict=strict, L=(4.0, 2.0, 2.0), bcs=((0, 0), (2, 2), (0, 0)), Re=4200.0, dt=0.005,
                    stretching=("uniform", "top-bottom", "uniform"), beta=(1.0, 0.259065151, 1.0))
    else:
        sim = X.Sim(tuple(grid), strict=strict)
    if "--channel" not in sys.argv:
        sim.init_tgv()
    stream = torch.cuda.ExternalStream(sim.stream())
    npts = grid[0] * grid[1] * grid[2]
    rows = []
    only = [a.split("=", 1)[1].split(",") for a in sys.argv if a.startswith("--ops=")]
    for op, b in BYTES_PER_PT.items():
        if only and op not in only[0]:
            continue
        ms = time_op(sim, op, 5, stream)
        gbs = b * npts / ms / 1e6
        rows.append((op, ms, gbs, gbs / peak))
        if rank == 0:
            print(f"{op:22s} {ms:9.3f} ms  {gbs:9.1f} GB/s  {100 * gbs / peak:6.1f}% of measured {peak:.0f} GB/s", flush=True)
    if rank == 0:
        print(f"local grid {grid} per rank, ranks={world}, strict={strict}")
    sim.close()


if __name__ == "__main__":
    main()
