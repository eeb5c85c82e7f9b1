"""oracle/ — TEST INFRASTRUCTURE ONLY.

CPU restatement of the rec_pangu ranking / multi-task hot path.
Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import this package, and only as the checker /
baseline — never from ``rec_pangu_b200`` (the product path has no CPU fallback).

Parity pinning: the reference ships no golden vectors for this path
(SURVEY.md §4, §8c).  The oracle is pinned against outputs of the reference's
own classes imported from ``/root/reference`` in the build container:
``tests/golden/make_golden.py`` generated ``tests/golden/*.npz`` and
``tests/test_oracle_golden.py`` checks every oracle function against them.
"""
from .restatement import *  # noqa: F401,F403
from .index_routing import *  # noqa: F401,F403
