"""Run one backend op a few times (for ncu): python tools/run_op.py N op reps"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import x3d2_b200 as X

n, op, reps = int(sys.argv[1]), sys.argv[2], int(sys.argv[3])
sim = X.Sim((n, n, n))
sim.init_tgv()
sim.bench_op(op, reps)
sim.sync()
sim.close()
