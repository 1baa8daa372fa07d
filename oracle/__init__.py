"""ORACLE -- CPU restatement of the reference's hybridized-SBP solve path.

TEST INFRASTRUCTURE ONLY.  Only tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs may import this package, and only as the
checker / CPU baseline; the product (hybridsbp_b200/) never does.

Parity status: PARITY UNPINNED by golden vectors -- the reference ships none
and is written in a language that is not installed in the build container, so
it could not be executed to make any.  The restatement is pinned instead by the
reference's own identities, see tests/test_oracle_*.py.
"""
