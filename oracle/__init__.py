"""CPU oracle for the B200 matcher/verifier -- TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this package.  PARITY UNPINNED (no reference golden vectors exist; see oracle_match.c).
"""
from .match import (  # noqa: F401
    match_bruteforce,
    match_guided,
    fast_match_pair,
    fast_match_pairs,
    fast_isa,
    acos_lut,
    build,
)
