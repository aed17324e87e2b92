"""featureextraction.jl_b200 -- B200-native Kernel PCA / Kernel LDA engine.

Drop-in for the kpca.jl / klda.jl solvers of trthatcher/FeatureExtraction.jl:
the public functions keep the reference's names and argument order and call a
C-ABI shared library (include/fxb200.h) of hand-written sm_100a CUDA kernels.
Import as ``import fxb200`` (root shim) -- the directory name is not a valid
Python identifier.  Importing never touches the oracle; calling anything
without the built CUDA library raises ``FxLibraryMissing``.
"""
from ._lib import (FxError, FxLibraryMissing, DimensionMismatch, SolverOpts, Stats, LIB_PATH,
                   PROTOTYPES, load as load_library)
from .kernels import Kernel, GaussianKernel, PolynomialKernel, SigmoidKernel, from_spec
from .api import (Context, Model, default_context, make_opts, kpca, transform_kpca, klda,
                  transform_klda, model_from_components, pca, transform_pca, lda, transform_lda, kernelmatrix, centerkernelmatrix, symm, center_symm,
                  syevd_small, bench_kernel)
from . import synth, sharding

__all__ = ["kpca", "transform_kpca", "klda", "transform_klda", "pca", "transform_pca", "lda", "transform_lda", "GaussianKernel",
           "PolynomialKernel", "SigmoidKernel", "Context", "Model"]
