"""CPU oracle for the GPflowPILCO hot path — TEST INFRASTRUCTURE ONLY.

This package is a float64 CPU restatement (torch, so that ``torch.autograd`` doubles as the gradient
oracle) of the reference's moment-matching / pathwise rollout maths.  Every function cites the
reference ``file:line`` it follows (paths relative to the upstream repository root).

Rules:
  * only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference``
    legs may import this package, and only as the checker / the timed CPU baseline;
  * nothing under ``gpflowpilco_b200/`` imports it; the product path fails loudly without its CUDA library.

Parity pin status:
  * The upstream package cannot be imported with its real dependencies (tensorflow, gpflow,
    gpflow_sampling, tfp, gym are absent).  ``oracle/refshim`` provides a numpy-backed stand-in for the
    few dozen TF/GPflow/TFP calls the hot path makes, so the UNMODIFIED upstream sources are executed
    from where they lie and their outputs frozen in ``tests/golden/*.npz`` (generator committed).
  * Additionally the upstream Monte-Carlo tests are ported (``tests/test_oracle_mc.py``).
  * gpflow_sampling (pathwise sampler) has no source in the tree and no upstream test: that row's
    parity is UNPINNED; ``oracle/pathwise.py`` fixes the contract.
"""
