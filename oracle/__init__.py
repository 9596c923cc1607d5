"""CPU oracle for the VeloCycle hot path -- TEST INFRASTRUCTURE ONLY.

Nothing under ``velocycle_b200/`` may import this package.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference``
legs use it, and only as the checker / the timed CPU baseline.
"""
