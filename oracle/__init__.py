"""CPU oracle for the KPCA / KLDA hot path -- TEST INFRASTRUCTURE ONLY.

This package is a NumPy + SciPy(LAPACK) restatement of the reference
(trthatcher/FeatureExtraction.jl, Julia 0.4) for the path BASELINE.json names.
Every function cites the reference file:line it follows.

Who may import it: ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py``; there only as the
checker or the CPU arm, never on the product path.  The product
(``featureextraction.jl_b200``) never imports this package and fails loudly when
its CUDA library is missing.

Parity pinning status (see DESIGN.md section "Oracle"):
  * The reference cannot be executed here (no ``julia``; Julia-0.4 syntax; the
    MLKernels dependency is not vendored).  The oracle is pinned by restating
    every test under /root/reference/test (tests/test_oracle_reference_tests.py):
    the known-answer tests of ``test_common.jl`` and the cross-solver identity
    ``test_kpca.jl:44-56`` (KPCA with PolynomialKernel(1,0,1) == PCA), which is
    the only reference test crossing the MLKernels boundary.
  * PARITY UNPINNED by the reference's own tests: Gaussian / polynomial
    degree>1 / sigmoid kernel values, ``transform_kpca`` (broken upstream,
    kpca.jl:67), every KLDA function (``klda!`` is exported but undefined,
    klda.jl:42 solves the wrong pencil).  For those the oracle implements the
    semantics written down in SURVEY.md A.2 / A.3 and is cross-checked against
    scikit-learn's KernelPCA (tests/test_oracle_sklearn.py).

MLKernels (REQUIRE:2, ``MLKernels 0.1`` lower bound, not in /root/reference):
kernel definitions restated from its published API -- see oracle/kernels.py.
"""
from .common import (translate, perturb, regularize, syml, symu, syevd,
                     components_svd, components_eig, components_geig)
from .kernels import (Kernel, GaussianKernel, PolynomialKernel, SigmoidKernel,
                      kernelmatrix, centerkernelmatrix)
from .kpca import (pca, transform_pca, pca_components_eig, pca_components_svd,
                   kpca, kpca_topk, transform_kpca, transform_kpca_reference_intent)
from .klda import (class_counts, class_totals, center_rows, lda_matrices,
                   lda_components, lda, transform_lda, klda, transform_klda,
                   scatter_matrices)
