"""CPU oracle for the reference LM/Schur/PCG path -- TEST INFRASTRUCTURE.

Only tests/, __graft_entry__.smoke() and bench.py (cpu_baseline / --impl reference)
may import this package.  The product package graphite_b200 never does.
"""
