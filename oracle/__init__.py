"""oracle/ -- TEST INFRASTRUCTURE (CPU restatement of the reference get() path).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this package. The product (ddstore_b200) never does.
"""
