"""CPU oracle for the soft-DP alignment path -- TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs may import this package.  See softdp_oracle.c.
"""
