"""CPU oracle for the depth supervision / evaluation hot path.

TEST INFRASTRUCTURE ONLY. Importable from tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / `--impl reference` legs - never from the product package. Each function cites
the reference file:line it restates; parity is pinned by tests/golden/ vectors produced from the
reference's own code by oracle/gen_golden.py (see oracle/README.md).
"""
