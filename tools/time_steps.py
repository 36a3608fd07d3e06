"""Time N TGV steps with CUDA-synchronised wall clock: python tools/time_steps.py n steps"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import x3d2_b200 as X

n, steps = int(sys.argv[1]), int(sys.argv[2])
sim = X.Sim((n, n, n))
sim.init_tgv()
sim.step(3)
sim.sync()
t0 = time.perf_counter()
sim.step(steps)
sim.sync()
print("%.3f ms/step" % (1e3 * (time.perf_counter() - t0) / steps), sim.monitor())
sim.close()
