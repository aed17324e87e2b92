"""Kernel structs with MLKernels' constructor argument order (test_kpca.jl:49).
The product side carries only (id, 3 params); the arithmetic lives in the CUDA
epilogues (csrc/fx_common.cuh ``kernel_map``)."""


class Kernel:
    kernel_id = -1

    def params(self):
        raise NotImplementedError


class GaussianKernel(Kernel):
    """k(x,y) = exp(-alpha ||x-y||^2)."""
    kernel_id = 0

    def __init__(self, alpha=1.0):
        if not alpha > 0:
            raise ValueError("alpha must be positive")
        self.alpha = float(alpha)

    def params(self):
        return (self.alpha, 0.0, 0.0)


class PolynomialKernel(Kernel):
    """k(x,y) = (a x'y + c)^d -- PolynomialKernel(a, c, d)."""
    kernel_id = 1

    def __init__(self, a=1.0, c=1.0, d=2.0):
        self.a, self.c, self.d = float(a), float(c), float(d)

    def params(self):
        return (self.a, self.c, self.d)


class SigmoidKernel(Kernel):
    """k(x,y) = tanh(a x'y + c)."""
    kernel_id = 2

    def __init__(self, a=1.0, c=0.0):
        self.a, self.c = float(a), float(c)

    def params(self):
        return (self.a, self.c, 0.0)


def from_spec(name, params):
    name = name.lower()
    if name in ("gaussian", "rbf"):
        return GaussianKernel(*params)
    if name in ("polynomial", "poly"):
        return PolynomialKernel(*params)
    if name == "sigmoid":
        return SigmoidKernel(*params)
    raise ValueError("unknown kernel %r" % name)
