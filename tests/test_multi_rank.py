"""N > 1 coverage.

CPU (gloo, world_size 2): a test of the host layer's DECOMPOSITION (x3d2h_decompose: offsets, extents, cyclic
neighbours) under real message passing. The halo exchange and the z-slab -> y-slab transpose are re-stated here with
torch / numpy index maps as a specification of what nccl.cu / poisson.cu have to deliver; the product's exchange
kernels themselves only run on GPUs (no CPU fallback exists), where
GPU (>= 2 devices): tools/mgpu_check.py under torchrun compares every rank's slab with the oracle's P-rank emulation,
at small shapes and at the per-rank shape of the benchmark.
"""
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _gloo_worker(rank, world, port, q):
    import torch
    import torch.distributed as dist
    sys.path.insert(0, ROOT)
    import x3d2_b200 as X
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    dims = (16, 8, 12 * world)
    d = X.decompose(dims, (1, 1, world), rank)
    nzl, off = d["vert_dims"][2], d["n_offset"][2]
    g = np.random.default_rng(3).standard_normal((dims[2], dims[1], dims[0]))
    loc = g[off:off + nzl]
    # sendrecv_fields (omp/sendrecv.f90:23-33): send_s -> prev, recv_e <- next, send_e -> next, recv_s <- prev
    prev, nxt = d["pprev"][2], d["pnext"][2]
    send_s, send_e = torch.from_numpy(loc[:4].copy()), torch.from_numpy(loc[-4:].copy())
    recv_s, recv_e = torch.empty_like(send_s), torch.empty_like(send_e)
    reqs = [dist.isend(send_s, prev), dist.irecv(recv_e, nxt), dist.isend(send_e, nxt), dist.irecv(recv_s, prev)]
    for r in reqs:
        r.wait()
    ok = np.array_equal(recv_s.numpy(), g[(off - 4) % dims[2]:(off - 4) % dims[2] + 4]) and \
        np.array_equal(recv_e.numpy(), g[(off + nzl) % dims[2]:(off + nzl) % dims[2] + 4])
    # FFT slab transpose (poisson.cu slab_pack_kernel): B(j, i, k_loc) -> blocks [r][k_loc][i][j_loc]; after the
    # all-to-all the received blocks are C(j_loc, i, k) with k = s * nz_loc + k_loc
    ny, nxh = dims[1], dims[0] // 2 + 1
    nyl = ny // world
    spec = np.random.default_rng(5).standard_normal((dims[2], nxh, ny))  # global (k, i, j), j fastest
    B = spec[off:off + nzl]                                                # this rank's z-slab
    send = [torch.from_numpy(np.ascontiguousarray(B[:, :, r * nyl:(r + 1) * nyl])) for r in range(world)]
    recv = [torch.empty_like(send[0]) for _ in range(world)]
    dist.all_to_all(recv, send) if dist.get_backend() != "gloo" else None
    if dist.get_backend() == "gloo":  # gloo has no all_to_all: same exchange with point-to-point messages
        reqs = []
        for r in range(world):
            if r == rank:
                recv[r].copy_(send[r])
            else:
                reqs += [dist.isend(send[r], r), dist.irecv(recv[r], r)]
        for r in reqs:
            r.wait()
    Cbuf = np.concatenate([t.numpy() for t in recv], axis=0)  # (k, i, j_loc)
    ok = ok and np.array_equal(Cbuf, spec[:, :, rank * nyl:(rank + 1) * nyl])
    # all-reduce of a scalar product
    s = torch.tensor([float((loc * loc).sum())], dtype=torch.float64)
    dist.all_reduce(s)
    ok = ok and abs(s.item() - float((g * g).sum())) < 1e-9
    q.put((rank, bool(ok)))
    dist.destroy_process_group()


def test_gloo_world2_exchange_patterns():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29600 + os.getpid() % 300
    procs = [ctx.Process(target=_gloo_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert sorted(res) == [(0, True), (1, True)]


@pytest.mark.gpu
@pytest.mark.parametrize("nz,z_path", [(128, "m1 (reference order)"), (256, "m4 (tma tiles)")])
def test_two_gpus_vs_oracle(nz, z_path):
    """z-slabs on 2 GPUs against the oracle's 2-rank emulation: 64 planes per rank use the reference-order
    DistD2 kernels with the 2x2 reduced systems, 128 planes per rank the distributed fast path (m3_edge.cu + TMA kernels)."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
           "127.0.0.1", "--master-port", "29533", os.path.join(ROOT, "tools", "mgpu_check.py"), "64", "64", str(nz), "2"]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900, env=dict(os.environ, X3D2C_TRACE="1"))
    print(r.stdout[-4000:], r.stderr[-2000:])
    assert r.returncode == 0 and "MGPU_CHECK PASS" in r.stdout
    assert f"transeq dir=3 ranks=2 -> {z_path}" in r.stderr and f"tds_solve dir=3 ranks=2 -> {z_path}" in r.stderr
