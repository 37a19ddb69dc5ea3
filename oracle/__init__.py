"""TEST INFRASTRUCTURE ONLY — CPU oracle for the CVT/RVD hot path.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs may import this package. The product (graphitethree_b200/) never does.
"""
