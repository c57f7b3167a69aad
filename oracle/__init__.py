"""TEST INFRASTRUCTURE ONLY -- CPU restatement ("oracle") of the reference's spectral hot path.

Nothing under ``oracle/`` is part of the product.  Only ``tests/``,
``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference`` legs of
``bench.py`` may import it, and only as the checker or the timed CPU baseline.  The
product path (``pde_fluid_phi_b200``) never imports this package and fails loudly when
its CUDA library is missing.

Parity pinning: the reference ships no golden vectors for this path (SURVEY.md section
8c), so the oracle is pinned against outputs of the reference itself: the fixtures under
``tests/golden/*.npz`` were produced by importing the unmodified reference modules from
``/root/reference/src`` (script: ``tests/golden/make_golden.py``) and
``tests/test_oracle_golden.py`` checks both restatements here against them.

Two restatements live here:

* ``spectral_oracle``  -- torch-CPU, same operation order as the reference
  (``torch.fft.rfftn`` -> polynomial loops -> ``einsum`` -> projection -> ``irfftn``);
  gradients come from autograd exactly as they do in the reference.
* ``dense_fp64``       -- numpy float64, dense DFT matrices and hand-derived backward
  formulas.  This is the specification the CUDA kernels are
  written against; it localises a failure to one stage.
"""
