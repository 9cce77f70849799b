"""CPU oracle for the dfdb_b200 hot path -- TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this package; the product (dataframedbs.jl_b200/) never does.
"""
