"""oracle/ — CPU restatement of the reference's FLUX DiT hot path.  TEST INFRASTRUCTURE ONLY:
importable from tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs; never from reflectionflow_b200/."""
