"""CPU oracle for the SG-MCMC hot path -- TEST INFRASTRUCTURE, NOT PRODUCT.

This package is a plain NumPy restatement of the algorithms on the reference's
(MFreidank/pysgmcmc) sampler hot path.  It exists to *check* the CUDA path:
only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` /
``--impl reference`` legs of ``bench.py`` may import it.  Nothing under
``pysgmcmc_b200/`` imports it, and the product path raises when the CUDA
extension is missing instead of falling back to this code.

Parity status (see DESIGN.md "Oracle"):

* The reference itself cannot run here (TensorFlow 1.x, arspy and pymc3 are not
  installable in this image), so no ``oracle/_ref`` exists.
* PINNED by the reference's own fixtures: ``safe_divide`` and both BNN priors
  (bit-level float64 goldens of ``tests/bayesian_neural_network/test_priors.py``),
  the banana / gmm objective-function doctest optima, the ``-50.0`` known answer
  of ``docs/source/notebooks/api_quickstart.ipynb:1104`` and NumPy's own
  ``RandomState`` for the minibatch index stream (bit-exact).
* Sampler *trajectories*: **parity unpinned** -- the reference holds no golden
  trajectory and cannot be executed; the line-by-line restatement below is the pin.
* pymc3 ESS / Gelman-Rubin and arspy (third-party, absent from /root/reference):
  **parity unpinned** bit-wise; restated from their published algorithms.  The ESS
  estimator and the relativistic sampler are pinned STATISTICALLY by the reference's
  published ESS-vs-stepsize table (tests/golden/relativistic_ess_published.json,
  reproduced within ~1 % by tools/ess_vs_stepsize.py; the GPU estimator is tested to
  equal oracle/diagnostics.py exactly).

Every function cites the reference file:line it follows.
"""

from . import tensor_utils, samplers, targets, bnn, mt19937, philox, diagnostics, svgd  # noqa: F401
