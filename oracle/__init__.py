"""oracle -- TEST INFRASTRUCTURE ONLY.

CPU checkers for the hot path: `ref` is the reference's own sorter compiled natively (oracle/_ref, git-ignored),
`port` is our C restatement (sort_oracle.c, raster_oracle.c).  Only tests/, __graft_entry__.smoke() and bench.py's
CPU-baseline legs may import this package; nothing under gaussiansplats3d_b200/ does.
"""
from .pyoracle import *  # noqa: F401,F403
