"""TEST INFRASTRUCTURE ONLY — CPU oracle for the per-cell GP hot path.

Nothing in ``optimalinterpolation_b200`` imports this package.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs
may use it, and only as the checker / the timed CPU baseline.

Parity pin: the reference ships no tests, golden vectors or fixtures for this path
 — "parity unpinned" by the reference's own tests.  The oracle is
instead pinned against outputs of the reference's own functions (``SGPkernel``, ``SMLII``,
``GPR3D``, extracted verbatim at run time from /root/reference by
``oracle/reference_functions.py``) executed in the build container; those outputs are
committed under ``tests/golden/`` together with the script that made them
(``tests/golden/make_golden.py``).
"""
