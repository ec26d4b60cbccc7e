"""CPU oracle for the GPAR-at-scale GP linear-algebra hot path.  TEST INFRASTRUCTURE ONLY.

This package is a float64 NumPy/SciPy (+ a small plain-C loop library, ``oracle/c``) restatement of
the algorithm the reference (TudorParas/GPAR-at-scale, Julia) runs on its hot path.  Only
``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference`` legs of
``bench.py`` may import it.  The product (``gpar-at-scale_b200``) never imports, links or calls
anything in here, and fails loudly when its CUDA library is missing.

PARITY UNPINNED.  The reference ships no tests, golden vectors or fixtures for this path, cannot
be executed here (no Julia), and all of its arithmetic lives in un-vendored dependencies without a
pinned version (``Project.toml:1-17`` has no ``[compat]``; ``Manifest.toml`` is git-ignored):
Stheno.jl (~0.6.x: kernels, ``pairwise``, ``cov``, ``logpdf``, posterior) and TemporalGPs.jl
(~0.1-0.2: ``to_sde``, LGSSM, ``decorrelate``, ``logpdf``, ``smooth``).  Their published algorithms
are restated here and anchored on the reference's own call sites (cited per function as
``file:line`` relative to the reference root).  What *is* pinned, in ``tests/``:

* the only self-check the reference contains (``examples/dtc_example.jl:8-64``): the LGSSM-whitened
  DTC objective equals the dense-Cholesky DTC;
* the mask identity of ``src/util.jl:57-96`` with its literal numbers;
* independent dense ground truths (direct Gaussian logpdf, dense GP posterior) and 50-digit mpmath
  spot values.

Details that come from memory of the dependencies and move results by <= 1e-12 relative are
flagged in the docstrings (``smooth``'s 1e-12 jitter, the 1e-18 default observation noise).
"""
from .params import unpack_gp, unpack_gpar, EQ, MATERN12, MATERN32, MATERN52, KERNEL_NAMES
from .kernels import (base_kernel, pairwise, scaled_kernel_matrix, get_time_mask, get_output_mask,
                      gpar_kernel_matrix, to_colvecs)
from .lgssm import (sde_matrices, transition, build_lgssm, kalman_decorrelate, kalman_logpdf,
                    kalman_smooth, dense_time_cov)
from .dtc import (compute_gpar_dtc_objective, gpar_dtc_collapsed, dtc_dense, dtc_diag, elbo_diag,
                  compute_q_u)
from .exact import exact_logpdf, exact_posterior
from .predict import sde_predictions, gpar_scaled_predict_given_eps
