"""Workload generators in front of the hot path -- NOT product code and not counted as such.

``composer.py`` mirrors the reference's serial host-side constraint system (``src/lib.rs:100-1198``,
``src/permutation.rs:22-200``; SURVEY section 2 row 9 marks it out of scope): identical gate layouts are
needed to reproduce the reference's circuits, nothing here runs on the GPU.  ``synthetic.py`` holds the
seeded input generators of the benchmarks.  Pure numpy: importing this package never loads
``libzkp_b200.so``, so both arms of ``bench.py`` and the oracle-side tests can share the circuits.
"""
