"""TEST INFRASTRUCTURE ONLY -- CPU oracle for the TAGAN hot path.

This package restates, in plain CPU torch / numpy, the arithmetic of the reference's
per-snapshot attention-and-propagation core.  It is the checker:
only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import it.  Nothing under ``tagan_b200/`` (the product)
imports it, and the product has no CPU fallback.

Parity pinning: the reference ships no golden vectors, so the oracle
is pinned by *executing the reference modules themselves* in the build container
(``oracle/make_golden.py`` imports them from ``/root/reference``) and committing the
resulting input/output/gradient vectors under ``tests/golden/``.  ``tests/test_oracle_golden.py``
checks every restatement against those vectors on CPU.
"""
