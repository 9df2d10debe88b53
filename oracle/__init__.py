"""CPU oracle for the particle-filter hot path — TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import
this package.  The product (sequential_monte_carlo_b200) never does.  See oracle/smc_oracle.c.
"""
