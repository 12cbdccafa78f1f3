"""Seeded synthetic clouds for the tests: the generators live in hotrack_b200/synthetic.py
(bench.py and tools/ use the same ones)."""
from hotrack_b200.synthetic import *  # noqa: F401,F403
from hotrack_b200.synthetic import KINDS, make  # noqa: F401
