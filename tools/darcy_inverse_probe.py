"""Where does the time of the small Darcy configuration go? (developer probe)"""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from nonlinpdes_gpsolver_b200 import InverseProblems
np.random.seed(9999)
d = InverseProblems.Darcy_flow2d(bdy=lambda x, y: 0, rhs=lambda x, y: 1)
d.sampled_pts(400, 100, 60)
d.get_observation(np.zeros(60), 1e-3)
init = np.random.normal(0, 1, 2400)
eng = d._engine()
for rep in range(3):
    T = {}
    def tick(name, fn):
        eng.sync(); t0 = time.perf_counter(); r = fn(); eng.sync(); T[name] = round((time.perf_counter() - t0) * 1e3, 2); return r
    tick("Gram_matrix", lambda: d.Gram_matrix("Gaussian", 0.2, 1e-8, "adaptive"))
    tick("Gram_Cholesky", d.Gram_Cholesky)
    tick("setup_gn", d._setup_gn)
    tick("inverse0", lambda: eng.inverse(0))
    tick("inverse1", lambda: eng.inverse(1))
    tick("set_z", lambda: eng.gn_set_z(init))
    tick("loss0", eng.gn_loss)
    for k in range(3):
        tick(f"step{k}", lambda: eng.gn_step(1.0))
    print(rep, T, flush=True)
