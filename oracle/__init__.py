"""ORACLE -- test infrastructure only.

CPU restatements of the reference's depth-loss hot path (oracle_torch.py, rays_oracle.c) and the
recipe that compiles the UNMODIFIED reference headers into oracle/_ref/ (Makefile).  Only tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import or execute
anything in this directory -- as the checker, never as the thing measured or shipped.
Parity status: PINNED against the reference build (tests/test_oracle_pin.py, tests/golden/).
"""
