"""CPU oracle -- TEST INFRASTRUCTURE ONLY (see oracle/mrs_oracle.c header).

May be imported only by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs.  The product package never imports it.
"""
from .oracle import Oracle, build_oracle, scale, spark_baseline_mae, max_threads, mean, std, combine  # noqa: F401
