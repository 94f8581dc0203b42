"""CPU oracle for the trex_b200 hot path.  TEST INFRASTRUCTURE ONLY (see trex_oracle.c header):
only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs import it."""
