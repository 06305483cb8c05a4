"""CPU oracle for the HermesPy channel hot path -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Everything under ``oracle/`` is a float64 numpy restatement of the reference algorithm
(each function cites the reference ``file:line`` it follows).  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs may
import it, and only as the checker / reported CPU baseline.  The product package
``hermespy_b200`` never imports ``oracle`` and has no CPU fallback.

Parity pinning: the restatement is checked (a) against the live reference in the build
container (``tests/test_oracle_vs_reference.py``, skipped when ``/root/reference`` is absent) and
(b) against golden vectors generated from the reference by ``oracle/make_golden.py`` and committed
under ``tests/golden/`` (``tests/test_oracle_golden.py``, runs everywhere).
"""
