"""CPU oracle for the PnP-ADMM hot path.  TEST INFRASTRUCTURE ONLY.

Nothing under ``oracle/`` is product code.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import it, and only as the checker.
The product path (``tfpnp_b200``) never imports this package and fails loudly
when its CUDA library is missing.
"""
