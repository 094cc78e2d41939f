"""CPU oracle package: test infrastructure only (see oracle/of_oracle.c, velocity_oracle.py).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this package. The product package never does.
"""
