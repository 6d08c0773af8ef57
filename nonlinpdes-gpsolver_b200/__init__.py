"""B200-native Gauss-Newton hot path of the GP-PDE solver (drop-in for the reference's src/ package).

Import as ``nonlinpdes_gpsolver_b200`` (the underscore alias at the repo root).  Same module and class
names as the reference: ``solver.solver_GP``, ``PDEs.{Nonlinear_elliptic2d, Burgers, Eikonal}``,
``InverseProblems.Darcy_flow2d``, ``Gram_matrice.{Gram_matrix_assembly, construct_Theta_test}``,
``kernels.{Gaussian_kernel, Anisotropic_Gaussian_kernel}``, ``sample_points``.
All numerics run in hand-written sm_100a CUDA behind the C ABI of ``include/gpp.h``; there is no CPU
fallback -- importing works anywhere, the first compute call raises if the library or a B200 is missing.
"""
__all__ = ["solver", "PDEs", "InverseProblems", "Gram_matrice", "kernels", "sample_points"]
__version__ = "0.1.0"
