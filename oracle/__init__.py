"""CPU oracle for the tactile hot path -- TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference`` leg may
import this package; the product (``tacex_b200``) must never do so.
"""
