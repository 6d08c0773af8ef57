"""Generates tests/golden/sampling_golden.npz by importing the reference's own sampler
(/root/reference/src/sample_points.py is pure numpy, so it runs here without JAX).
Run in the build container only; the GPU box reads the committed .npz."""
import importlib.util
import os

import numpy as np

spec = importlib.util.spec_from_file_location("ref_sample_points", "/root/reference/src/sample_points.py")
ref = importlib.util.module_from_spec(spec)
spec.loader.exec_module(ref)

out = {}
cases = [("ell", np.array([[0.0, 1.0], [0.0, 1.0]]), False, 900, 124), ("bur", np.array([[0.0, 1.0], [-1.0, 1.0]]), True, 1000, 200),
         ("odd", np.array([[-0.5, 2.0], [1.0, 3.0]]), False, 37, 23), ("oddt", np.array([[0.0, 2.0], [-3.0, 1.0]]), True, 41, 20)]
for name, dom, td, N, Nb in cases:
    for seed in (0, 9999):
        np.random.seed(seed)
        Xd, Xb = ref.sampled_pts_rdm(N, Nb, dom, time_dependent=td)
        after = np.random.normal(0.0, 1.0, 3)          # RNG stream position after sampling
        out[f"rdm_{name}_{seed}_Xd"], out[f"rdm_{name}_{seed}_Xb"], out[f"rdm_{name}_{seed}_next"] = Xd, Xb, after
    Xd, Xb = ref.sampled_pts_grid(N, Nb, dom, time_dependent=td)
    out[f"grid_{name}_Xd"], out[f"grid_{name}_Xb"] = Xd, Xb
np.savez_compressed(os.path.join(os.path.dirname(os.path.abspath(__file__)), "sampling_golden.npz"), **out)
print("wrote", len(out), "arrays")
