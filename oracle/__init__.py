"""CPU oracle for the dvd_b200 hot path — TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / `--impl reference` leg may
import this package. The product path (dynamic-video-depth_b200/) never does.
Parity status: PINNED — every function here is checked against the reference's own PyTorch code
executed in the authoring container (tests/test_oracle_vs_reference.py, skipped when
/root/reference is absent) and against the fixtures that run produced (tests/golden/).
"""
