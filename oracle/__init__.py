"""CPU oracle for the GTOP hot path — TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this package.  The product package (grad_traj_optimization_b200) never does.
PARITY UNPINNED upstream: see oracle/gtop_oracle.hpp.
"""
from .binding import (OracleMap, OracleOptimizer, OracleParams, TRAJ_METRICS, build, lib_path, nlopt_path,  # noqa: F401
                      traj_metrics)
