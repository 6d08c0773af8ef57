"""Kernel classes with the reference's names and method signatures (src/kernels.py:8-179).

Every method evaluates the closed form of the corresponding jax.grad chain on the GPU through
gpp_kernel_eval (include/gpp.h) and accepts scalars or broadcastable arrays.  The Gram builders do not
go through these methods: they use the fused assembly kernel."""
from . import _lib

_ID, _D1, _D2, _D22, _LAP = 0, 1, 2, 3, 4
# method name -> (operator on x, operator on y)
_METHODS = {
    "kappa": (_ID, _ID),
    "D_x1_kappa": (_D1, _ID), "D_x2_kappa": (_D2, _ID), "DD_x2_kappa": (_D22, _ID),
    "D_y1_kappa": (_ID, _D1), "D_y2_kappa": (_ID, _D2), "DD_y2_kappa": (_ID, _D22),
    "D_x1_D_y1_kappa": (_D1, _D1), "D_x1_D_y2_kappa": (_D1, _D2), "D_x1_DD_y2_kappa": (_D1, _D22),
    "D_x2_D_y2_kappa": (_D2, _D2), "D_x2_D_y1_kappa": (_D2, _D1), "D_x2_DD_y2_kappa": (_D2, _D22),
    "DD_x2_DD_y2_kappa": (_D22, _D22),
    "Delta_x_kappa": (_LAP, _ID), "Delta_y_kappa": (_ID, _LAP), "Delta_x_Delta_y_kappa": (_LAP, _LAP),
    "Delta_x_D_y1_kappa": (_LAP, _D1), "Delta_x_D_y2_kappa": (_LAP, _D2),
}


class _KernelBase(object):
    _name = None

    def __init__(self):
        pass

    def _eval(self, ops, x1, x2, y1, y2, sigma):
        out = _lib.default_engine().kernel_eval(self._name, sigma, ops[0], ops[1], x1, x2, y1, y2)
        return out if out.shape else float(out)


def _make(name, ops):
    def method(self, x1, x2, y1, y2, sigma):
        return self._eval(ops, x1, x2, y1, y2, sigma)
    method.__name__ = name
    return method


class Gaussian_kernel(_KernelBase):
    """kappa = exp(-|x-y|^2 / (2 sigma^2))   (src/kernels.py:8-89)"""
    _name = "Gaussian"


class Anisotropic_Gaussian_kernel(_KernelBase):
    """kappa = exp(-((x1-y1)/s_t)^2 - ((x2-y2)/s_x)^2), sigma = [s_t, s_x]   (src/kernels.py:91-179)"""
    _name = "anisotropic_Gaussian"


for _n, _o in _METHODS.items():
    setattr(Gaussian_kernel, _n, _make(_n, _o))
    setattr(Anisotropic_Gaussian_kernel, _n, _make(_n, _o))
# src/kernels.py:163: unused duplicate of Delta_x_Delta_y_kappa on the anisotropic class
Anisotropic_Gaussian_kernel.Delta_x_y_kappa = _make("Delta_x_y_kappa", (_LAP, _LAP))
