"""CPU oracle for the FAST Monte-Carlo hot path.  TEST INFRASTRUCTURE: only tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import it."""
