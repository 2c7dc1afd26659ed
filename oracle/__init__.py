"""CPU oracle for the SAC-learner / prioritized-replay hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is part of the product:
only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline``
/ ``--impl reference`` legs may import it, and only as the checker or the timed
CPU baseline.  The product path (``advanced-soft-actor-critic_b200/``) never
imports this package and fails loudly when its CUDA library is missing.

Parity status: the reference's own tests hold no golden vectors or assertions
for this path, so the oracle is pinned differentially:
``oracle/gen_golden.py`` imports the real reference from ``/root/reference`` in
the build container, runs it on seeded inputs with injected Gaussian noise and
writes ``tests/golden/*.npz``; ``tests/test_oracle_golden.py`` checks every
oracle function against those files.
"""
