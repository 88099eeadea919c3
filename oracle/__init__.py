"""ORACLE — test infrastructure only.

CPU restatement of the reference's ELIC_united compress/decompress/forward path, used as
the checker by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs.  Nothing under the product package imports this directory; the product path fails
loudly when its CUDA extension is missing.

Parity pin: every function here is checked (tests/test_oracle_*.py) against golden vectors
in tests/golden/ that were produced by the unmodified reference (oracle/make_golden.py,
run in the build container where /root/reference is mounted), and the rANS restatement is
additionally checked against the reference's own compiled coder in oracle/_ref/.
"""
