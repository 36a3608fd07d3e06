"""Run a few TGV time steps (for ncu launch lists): python tools/run_steps.py N steps"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import x3d2_b200 as X

n, steps = int(sys.argv[1]), int(sys.argv[2])
sim = X.Sim((n, n, n))
sim.init_tgv()
sim.step(steps)
sim.sync()
print(sim.monitor())
sim.close()
