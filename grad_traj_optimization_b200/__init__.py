"""gtop-b200: B200-native inner loop of GTOP (EpicOne1/grad_traj_optimization).

The product is `libgtop_b200.so` (hand-written sm_100a CUDA behind the C ABI of include/gtop_c.h)
plus the ROS-free C++ host mirror in include/gtop/.  This Python package is only the plumbing the
tests and bench.py use to reach that ABI: `_capi` (ctypes), `api` (handles), `fixtures` (inputs).
There is no CPU fallback: `api` raises if the library is missing or no CUDA device is present.
"""
__version__ = "0.1.0"
