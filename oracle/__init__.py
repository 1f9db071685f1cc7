"""CPU oracle for the improved-diffusion hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is part of the product:
only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` /
``--impl reference`` legs of ``bench.py`` may import it, and there only as the
checker (or as the CPU arm being timed), never as the thing shipped.

Parity status: the reference repository has no tests, golden vectors or
known-answer fixtures of its own, so the oracle is
pinned the other way the task allows: ``oracle/gen_golden.py`` imports the
*unmodified* reference modules from ``/root/reference`` (behind import shims
for the absent ``pytorch_lightning`` / ``matplotlib``) in the build container,
runs them on seeded inputs and commits the outputs under ``tests/golden/``.
``tests/test_oracle_golden.py`` checks every oracle function against those
fixtures.  The learned-variance / L_hybrid extension has no reference
implementation: for that part parity is **unpinned** and the
oracle composes the reference's own ``normal_kl`` /
``discretized_gaussian_log_likelihood`` / ``q_posterior`` per Appendix C.
"""
