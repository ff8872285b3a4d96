"""CPU oracle for the render path.  TEST INFRASTRUCTURE ONLY (see oracle/README.md).

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference``
legs may import this package, and only as the checker.
"""
