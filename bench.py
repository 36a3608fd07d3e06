#!/usr/bin/env python
"""bench.py — headline benchmark of the cuda_c backend: Taylor-Green vortex, FP64, RK3.

  python bench.py --gpus N --steps K --warmup W            (N > 1: launched under torchrun by the driver)
  python bench.py --impl reference --steps K --warmup W    (CPU arm: the oracle port of the reference OMP backend)

A "step" is one full RK3 time step of the reference's time loop (src/case/base_case.f90:246-289): three sub-stages
of transeq + time integration + pressure correction. Metric: Mpt-steps/s = global grid points * steps / time / 1e6.
Workload (weak scaling, 512^3 points per GPU): 512^3 (N=1, BASELINE.json configs[2]), 512x512x1024 (N=2),
512x1024x1024 (N=4), 1024^3 (N=8, configs[3]); slab decomposition nproc_dir = (1, 1, N).
One JSON line is printed by rank 0. It also carries (config.*): the drop-in number (`dropin_ms_per_step`: the same step
issued through the base_backend_t entry points only, exactly the reference solver's operator graph) and, at N = 1,
BASELINE.json configs[1] (`standalone_256`: tds_solve / transeq at 256^3 against the HBM roofline).

CPU legs (the oracle port; the reference binary cannot be built in this image): the thread count is set explicitly
through the oracle's ABI (torchrun exports OMP_NUM_THREADS=1) and the count reported is the one an OpenMP region
really got. `--impl reference` runs the per-GPU workload itself (512^3) when K + W steps fit its time budget at the
rate measured on this box, otherwise the largest z-shortened periodic box that does.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

ALGO_BYTES_PER_PT = {"transeq": 48, "tds_solve": 16}  # per transeq_{x,y,z} / tds_solve call
STEP_BYTES_PER_PT = 3888                                # whole RK3 step, the reference's operator graph
# What this backend's operator graph moves per RK3 step (DESIGN.md "bytes per step"): per stage transeq 192 (3 x 48
# kernels - the y sweep reads the x layout itself - and 3 x2z reorders of 16), divergence 112 (x solves store in the y
# layout: 3 x 16; y 24 + 16; z 24), Poisson 152, gradient + correction 152 (c2z pass 16, z 24, y 24 + 16, x solves load from
# the y layout: 3 x 24); the y / z sums of transeq fused with the RK3 update: 3 components x (48 + 56 + 56) per step
STEP_BYTES_MOVED_PER_PT = 3 * (192 + 112 + 152 + 152) + 3 * (48 + 56 + 56)


def grid_for(n_gpus, base):
    dims = [base, base, base]
    k, axis = n_gpus, 2
    while k > 1:  # double z, then y, then x: 1 -> (b,b,b), 2 -> (b,b,2b), 4 -> (b,2b,2b), 8 -> (2b,2b,2b)
        dims[axis] *= 2
        axis -= 1
        k //= 2
    return dims


def measured_peak():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return json.load(f)["hbm_gbs"], "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            p = [x.strip() for x in ln.split(",")]
            if len(p) < 7:
                continue
            try:
                sm.append(float(p[0]))
                mx = float(p[1])
            except ValueError:
                continue
            for name, val in zip(names, p[3:7]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(sm)}


def host_threads():
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


def cpu_port_run(dims, steps, warmup, threads=None):
    """Times the oracle port (C++/OpenMP restatement of the reference OMP backend) on the host cores.
    Returns (Mpt-steps/s, s per step, threads really used)."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import _oracle
    _oracle.set_num_threads(threads or host_threads())  # overrides OMP_NUM_THREADS (= 1 under torchrun)
    used = _oracle.num_threads()
    w = _oracle.World(tuple(dims))
    w.init_tgv()
    for _ in range(warmup):
        w.step(1)
    t0 = time.perf_counter()
    w.step(steps)
    dt = time.perf_counter() - t0
    del w
    pts = dims[0] * dims[1] * dims[2]
    return pts * steps / dt / 1e6, dt / steps, used


def cpu_sample_grid(size, n_steps_total, budget_s):
    """The CPU sample of the `size`^3-per-GPU workload: the block itself if n_steps_total steps fit the budget at the
    rate MEASURED here on a 256^3 step (fields of 128 MiB: not cache resident), else the same x-y lines with a shorter
    periodic z extent. Returns (dims, calibrated s per point-step)."""
    _, s_per_step, _ = cpu_port_run([256, 256, 256], 1, 1)
    per_pt = s_per_step / 256 ** 3
    cands = [[size, size, size], [size, size, size // 2], [size, size, size // 4], [size // 2] * 3, [size // 4] * 3]
    for dims in cands:
        if per_pt * dims[0] * dims[1] * dims[2] * n_steps_total * 1.15 <= budget_s:
            return dims, per_pt
    return cands[-1], per_pt


def workload_name(dims, size, world):
    return (f"Taylor-Green vortex {dims[0]}x{dims[1]}x{dims[2]} FP64 RK3 (Re=1600, dt=1e-3, compact6/classic), "
            f"{size}^3 points per GPU, nproc_dir=(1,1,{world})")


def run_reference(args):
    """--impl reference: the reference's own CPU path. The Fortran/MPI/2DECOMP build is impossible in this image
    (no gfortran, no MPI), so oracle/_ref does not exist and the oracle port is timed (kind = "port").
    Under torchrun only rank 0 works; the other ranks exit."""
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if rank != 0:
        return
    n_gpus = max(args.gpus, world)
    if args.cpu_size > 0:
        dims = [args.cpu_size] * 3
    else:
        # N = 1: the 512^3 workload itself unless the box is too slow for ~7 minutes; N > 1 (the N-GPU grid is N times
        # larger and does not fit a CPU run): one GPU's 512^3 share, bounded to ~4 minutes
        dims, _ = cpu_sample_grid(args.size, args.steps + args.warmup, 420.0 if n_gpus == 1 else 240.0)
    value, s_per_step, threads = cpu_port_run(dims, args.steps, args.warmup)
    gdims = grid_for(n_gpus, args.size)
    full = dims == gdims
    sample = (f"oracle port of the reference OMP backend (C++/OpenMP, strict IEEE), {threads} threads, TGV "
              f"{dims[0]}x{dims[1]}x{dims[2]} FP64 RK3, {args.steps} steps after {args.warmup} warm-up: " +
              ("the whole workload" if full else
               f"a bounded sample of the {gdims[0]}x{gdims[1]}x{gdims[2]} workload (same operators and per-point work on a "
               f"smaller periodic box; Mpt-steps/s is a per-point rate)"))
    out = {"impl": "reference", "metric": "Mpt-steps/s", "value": value, "unit": "Mpt-steps/s", "n_gpus": n_gpus,
           "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * s_per_step, "higher_is_better": True,
           "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
           "config": {"workload": workload_name(gdims, args.size, n_gpus)},
           "cpu_baseline": {"value": value, "unit": "Mpt-steps/s", "cores": threads, "kind": "port", "sample": sample},
           "e2e": {"value": value, "unit": "Mpt-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(out), flush=True)


def time_steps(sim, stream, torch, steps):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    sim.step(steps)
    e1.record(stream)
    e1.synchronize()
    return e0.elapsed_time(e1) / steps


def time_op(sim, stream, torch, op, reps=5):
    sim.bench_op(op, 2)
    sim.sync()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(stream)
    sim.bench_op(op, reps)
    b.record(stream)
    b.synchronize()
    return a.elapsed_time(b) / reps


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="cuda_c", choices=["cuda_c", "reference"])
    ap.add_argument("--size", type=int, default=512, help="grid points per direction per GPU (default 512)")
    ap.add_argument("--cpu-size", type=int, default=0, help="grid of the CPU baseline sample (0 = auto)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-standalone", action="store_true", help="skip the 256^3 standalone operator table")
    ap.add_argument("--strict", action="store_true", help="reference-order (bit-exact) kernels instead of the fast path")
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3
    if args.impl == "reference":
        return run_reference(args)

    import numpy as np
    import torch

    import x3d2_b200 as X

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus and world > 1:
        args.gpus = world
    torch.cuda.set_device(local_rank)
    nccl_id = nccl_id2 = None
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
        c, _ = X.load()
        buf = [None]
        if rank == 0:
            import ctypes
            raw = ctypes.create_string_buffer(128)
            assert c.x3d2c_nccl_unique_id(raw) == 0, c.x3d2c_last_error()
            buf = [raw.raw]
        dist.broadcast_object_list(buf, src=0)
        nccl_id = buf[0]
        buf2 = [None]  # a second communicator for the drop-in Sim
        if rank == 0:
            raw2 = ctypes.create_string_buffer(128)
            assert c.x3d2c_nccl_unique_id(raw2) == 0, c.x3d2c_last_error()
            buf2 = [raw2.raw]
        dist.broadcast_object_list(buf2, src=0)
        nccl_id2 = buf2[0]

    dims = grid_for(world, args.size)
    L = tuple(2 * np.pi * d / args.size for d in dims)
    sim = X.Sim(dims, nproc_dir=(1, 1, world), L=L, rank=rank, nproc=world, device=local_rank, strict=args.strict,
                nccl_unique_id=nccl_id)
    sim.init_tgv()
    stream = torch.cuda.ExternalStream(sim.stream())
    pts_global = dims[0] * dims[1] * dims[2]
    pts_local = pts_global // world

    def barrier_sim(s):
        s.sync()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()

    def barrier():
        barrier_sim(sim)

    # ---------------------------------------------------------------- device-resident timing (value)
    sim.step(args.warmup)
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    launches0 = sim.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    sim.step(args.steps)
    e1.record(stream)
    barrier()
    ms = e0.elapsed_time(e1)
    launches = sim.launch_count() - launches0
    clocks = sampler.stop() if rank == 0 else None
    if world > 1:
        t = torch.tensor([ms], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    ms_per_step = ms / args.steps
    value = pts_global / (ms_per_step * 1e-3) / 1e6
    mon = sim.monitor()

    # ---------------------------------------------------------------- roofline of the dominant kernel (transeq)
    peak, peak_src = measured_peak()
    roof = {}
    for op, key in (("transeq_x", "transeq"), ("tds_solve_x", "tds_solve")):
        op_ms = time_op(sim, stream, torch, op)
        roof[key] = {"ms": op_ms, "gbs": ALGO_BYTES_PER_PT[key] * pts_local / op_ms / 1e6}
    # DRAM bytes per transeq_x launch at 512^3 from the committed `ncu --set full` capture (dram__bytes_read.sum +
    # dram__bytes_write.sum; profiles/traffic.json names the capture and the commit it was taken at). A profiler
    # counter cannot be read inside an un-profiled timed run, so this is the per-launch constant of that capture.
    traffic, traffic_src = None, None
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            tj = json.load(f)
        if args.size == 512:
            traffic, traffic_src = tj.get("transeq_bytes_per_launch"), tj.get("source")
    except Exception:
        pass
    roofline = {"bound": "hbm", "kernel": "transeq_x (fused DistD2-TDS transeq, one call = 3 velocity components)",
                "achieved": roof["transeq"]["gbs"], "peak": peak, "unit": "GB/s", "frac": roof["transeq"]["gbs"] / peak,
                "traffic": traffic, "traffic_source": traffic_src, "peak_source": peak_src, "algorithmic_bytes_per_launch": 48 * pts_local,
                "ms_per_launch": roof["transeq"]["ms"],
                "tds_solve": {"achieved": roof["tds_solve"]["gbs"], "frac": roof["tds_solve"]["gbs"] / peak,
                              "ms_per_launch": roof["tds_solve"]["ms"], "algorithmic_bytes_per_launch": 16 * pts_local},
                # the timed step runs the FUSED operator graph of this repo's host layer: fraction = bytes it really moves
                "whole_step": {"graph": "fused host layer (extension entry points of include/x3d2c.h)",
                               "bytes_per_step": STEP_BYTES_MOVED_PER_PT * pts_local,
                               "achieved": STEP_BYTES_MOVED_PER_PT * pts_local / (ms_per_step * 1e-3) / 1e9,
                               "frac": STEP_BYTES_MOVED_PER_PT * pts_local / (ms_per_step * 1e-3) / 1e9 / peak,
                               "reference_graph_bytes_per_step": STEP_BYTES_PER_PT * pts_local,
                               "note": "the reference's operator graph would move 3888 B/pt (SURVEY.md 8d); that graph is "
                                       "timed separately as whole_step_dropin"}}
    # ---------------------------------------------------------------- end to end: host buffers in, host buffers out
    nz, ny, nx = sim.shape()
    host = [torch.empty((nz, ny, nx), dtype=torch.float64).pin_memory() for _ in range(3)]
    host_np = [h.numpy() for h in host]
    sim.get_uvw(out=host_np)
    e2e_steps = max(1, min(args.steps, 3))
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        sim.set_uvw(*host_np)          # H2D of the step's inputs (pinned)
        sim.step(1)
        sim.get_uvw(out=host_np)       # D2H of the step's result (synchronises)
    barrier()
    serial_s = (time.perf_counter() - t0) / e2e_steps
    # the same work as a stream of independent batches (Sim.step_batches): every batch is uploaded from the pinned host
    # arrays, advanced by one RK3 step and downloaded; the copies of neighbouring batches overlap the kernels
    host_out = [torch.empty((nz, ny, nx), dtype=torch.float64).pin_memory() for _ in range(3)]
    out_np = [h.numpy() for h in host_out]
    n_batches = max(4, min(3 * args.steps, 24))  # the first upload and the last download are not overlapped: amortised
    sim.step_batches(2, host_np, out_np)   # warm-up: staging blocks, copy streams
    barrier()
    t0 = time.perf_counter()
    sim.step_batches(n_batches, host_np, out_np)
    barrier()
    e2e_s = (time.perf_counter() - t0) / n_batches
    if world > 1:
        t = torch.tensor([e2e_s, serial_s], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_s, serial_s = float(t[0].item()), float(t[1].item())
    e2e = {"value": pts_global / e2e_s / 1e6, "unit": "Mpt-steps/s", "h2d_bytes_per_step": 3 * 8 * pts_local,
           "d2h_bytes_per_step": 3 * 8 * pts_local, "steps": n_batches, "ms_per_step": 1e3 * e2e_s,
           "how": "x3d2_b200.Sim.step_batches: independent batches, each = H2D of u, v, w from pinned host memory + one RK3 "
                  "step + D2H of u, v, w; uploads / downloads of neighbouring batches overlap the kernels on copy streams",
           "serial": {"value": pts_global / serial_s / 1e6, "ms_per_step": 1e3 * serial_s, "steps": e2e_steps,
                      "how": "set_uvw -> step -> get_uvw, each call synchronous (every step depends on the host copy of "
                             "the previous one)"}}

    sim.close()
    del sim

    # ---------------------------------------------------------------- drop-in path: base_backend_t entry points only
    # the unchanged reference solver's operator graph (solver.f90:291-389,693-739, vector_calculus.f90:142-332,
    # time_integrator.f90:166-231) issued call by call; 3888 B/pt per step
    dsim = X.Sim(dims, nproc_dir=(1, 1, world), L=L, rank=rank, nproc=world, device=local_rank, strict=args.strict,
                 nccl_unique_id=nccl_id2, base_ops=True)
    dsim.init_tgv()
    dstream = torch.cuda.ExternalStream(dsim.stream())
    dsim.step(3)
    barrier_sim(dsim)
    dl0 = dsim.launch_count()
    dsteps = max(1, min(args.steps, 5))
    dropin_ms = time_steps(dsim, dstream, torch, dsteps)
    dropin_launches = (dsim.launch_count() - dl0) // dsteps
    if world > 1:
        t = torch.tensor([dropin_ms], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dropin_ms = float(t.item())
    dmon = dsim.monitor()
    dsim.close()
    del dsim
    roofline["whole_step_dropin"] = {"graph": "reference operator graph through the base_backend_t entry points only",
                                     "bytes_per_step": STEP_BYTES_PER_PT * pts_local, "ms_per_step": dropin_ms,
                                     "achieved": STEP_BYTES_PER_PT * pts_local / (dropin_ms * 1e-3) / 1e9,
                                     "frac": STEP_BYTES_PER_PT * pts_local / (dropin_ms * 1e-3) / 1e9 / peak}

    # ---------------------------------------------------------------- BASELINE.json configs[1]: standalone 256^3
    standalone = None
    if world == 1 and not args.no_standalone:
        s256 = X.Sim((256, 256, 256), device=local_rank, strict=args.strict)
        s256.init_tgv()  # non-degenerate lines (a TGV field so that caches cannot exploit equal lines)
        st = torch.cuda.ExternalStream(s256.stream())
        standalone = {"grid": "256x256x256 FP64, TGV field at t = 0, periodic compact6", "reps": 20,
                      "note": "7 field blocks of 128 MiB cycle through a 126 MB L2: DRAM-resident"}
        for op, b in (("tds_solve_x", 16), ("tds_solve_y", 16), ("tds_solve_z", 16), ("transeq_x", 48),
                      ("transeq_y", 48), ("transeq_z", 48)):
            op_ms = time_op(s256, st, torch, op, reps=20)
            gbs = b * 256 ** 3 / op_ms / 1e6
            standalone[op] = {"ms": op_ms, "gbs": gbs, "frac": gbs / peak}
        s256.close()
        del s256

    # ---------------------------------------------------------------- CPU baseline (rank 0, N = 1 only)
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        # ~10-30 s of CPU work: the 512^3 workload itself (1 step after 1 warm-up step) if this box manages that,
        # else a z-shortened box; the grid and the thread count really used are stated
        if args.cpu_size > 0:
            cdims = [args.cpu_size] * 3
        else:
            cdims, _ = cpu_sample_grid(args.size, 2, 40.0)
        v, s_per_step, threads = cpu_port_run(cdims, 1, 1)
        cpu = {"value": v, "unit": "Mpt-steps/s", "cores": threads, "kind": "port",
               "sample": f"TGV {cdims[0]}x{cdims[1]}x{cdims[2]} FP64 RK3, 1 step after 1 warm-up step, oracle port of the "
                         f"reference OMP backend (C++/OpenMP, strict IEEE), {threads} threads, {1e3 * s_per_step:.0f} ms/step"}

    if rank == 0:
        out = {"metric": "Mpt-steps/s", "value": value, "unit": "Mpt-steps/s", "n_gpus": world, "steps": args.steps,
               "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
               "vs_baseline": None, "dtype": "f64", "data": "synthetic",
               "config": {"workload": workload_name(dims, args.size, world),
                          "dropin_ms_per_step": dropin_ms, "dropin_gpu_launches_per_step": dropin_launches,
                          "dropin_mpt_steps_per_s": pts_global / (dropin_ms * 1e-3) / 1e6,
                          "dropin_monitor_after_run": dmon, "standalone_256": standalone,
                          "mode": "strict (reference-order, bit-exact)" if args.strict else "fast (FMA)",
                          "l2": "every field block (8 B x points per GPU) exceeds the 126 MB L2; no flush needed",
                          "monitor_after_run": mon},
               "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": launches, "clocks": clocks}
        print(json.dumps(out), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
