"""Generates tests/golden/*.npz from the oracle (CPU restatement of the reference OMP backend).

The reference itself cannot be built or run in this image (no Fortran compiler / MPI) and stores no
golden vectors, so these fixtures pin the ORACLE (regression) and are what the CUDA path is compared with on the GPU
box, where /root/reference does not exist. Run:  python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import _oracle as O  # noqa: E402


def main():
    # 1) TGV 32^3, RK3: velocity after 2 steps + monitor series
    W = O.World((32, 32, 32))
    W.init_tgv()
    series = [list(W.monitor().values())]
    for _ in range(2):
        W.step(1)
        series.append(list(W.monitor().values()))
    u, v, w = W.get_uvw()
    np.savez_compressed(os.path.join(HERE, "tgv32_rk3_2steps.npz"), u=u, v=v, w=w, series=np.array(series))
    # 2) TGV 64^3, RK3, 100 steps: enstrophy / kinetic-energy series (north_star: 1e-10 over 100 steps)
    W = O.World((64, 64, 64))
    W.init_tgv()
    s = []
    for i in range(101):
        if i:
            W.step(1)
        m = W.monitor()
        s.append([m["enstrophy"], m["ke"], m["div_u_max"]])
    np.savez_compressed(os.path.join(HERE, "tgv64_rk3_100steps_series.npz"), series=np.array(s))
    # 3) TGV 64^3, AB3 (the shipped examples/TGV default), 10 steps
    W = O.World((64, 64, 64), time_intg="AB3")
    W.init_tgv()
    s = []
    for i in range(11):
        if i:
            W.step(1)
        m = W.monitor()
        s.append([m["enstrophy"], m["ke"]])
    np.savez_compressed(os.path.join(HERE, "tgv64_ab3_10steps_series.npz"), series=np.array(s))
    # 4) operator fixtures on a seeded 64x32x48 field (inputs are regenerated from the seed by the tests).
    #    Each output is stored as a strided sample (tolerance checks) plus a SHA-256 of its bytes (bit-exact checks).
    import hashlib
    rng = np.random.default_rng(2024)
    W = O.World((64, 32, 48), Re=100.0)
    f, g, h = (rng.standard_normal(W.shape()) for _ in range(3))
    out = {}

    def put(name, arr):
        out[name] = arr.ravel()[::7].copy()
        out[name + "_sha"] = np.frombuffer(hashlib.sha256(np.ascontiguousarray(arr).tobytes()).digest(), dtype=np.uint8)

    put("f", f)
    for d in (1, 2, 3):
        for op in ("der1st", "der2nd", "stagder_v2p", "interpl_v2p"):
            put(f"tds_{d}_{op}", W.tds_solve(d, op, f))
        for arr, k in zip(W.transeq_dir(d, f, g, h), ("du", "dv", "dw")):
            put(f"transeq_{d}_{k}", arr)
    put("div", W.divergence(f, g, h))
    put("poisson", W.poisson(f - f.mean()))
    np.savez_compressed(os.path.join(HERE, "ops_64x32x48.npz"), **out)
    print("golden fixtures written")


if __name__ == "__main__":
    main()
