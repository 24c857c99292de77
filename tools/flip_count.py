#!/usr/bin/env python
"""Compares two result dumps of tools/kbench.py --dump (exact polynomial order vs an experimental one): how many
trajectories differ at all (every difference is a flipped float rounding of a sample position / velocity), and by how much."""
import sys

import numpy as np

a, b = np.load(sys.argv[1]), np.load(sys.argv[2])
dc = np.abs(a["cost"] - b["cost"]) / np.abs(a["cost"])
dg = np.abs(a["grad"] - b["grad"]).max(axis=1) / np.abs(a["grad"]).max(axis=1)
diff = (dc > 0) | (dg > 0)
print("trajectories %d, differing %d; max relative difference cost %.3e gradient %.3e; above 1e-9: %d" % (
    len(dc), int(diff.sum()), dc.max(), dg.max(), int(((dc > 1e-9) | (dg > 1e-9)).sum())))
