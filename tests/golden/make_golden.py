"""Regenerates tests/golden/*.npz from the CPU oracle (oracle/) — run from the repo root:

    python tests/golden/make_golden.py

The reference ships no golden vectors (parity unpinned upstream), so these pin the
oracle's own outputs on the two upstream fixtures; SURVEY §10's survey-computed numbers are checked
against them in tests/test_oracle.py.  Files are small (a few hundred kB) and committed.
"""
import hashlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from common import oracle_map_from_cfg, oracle_problem  # noqa: E402
from grad_traj_optimization_b200 import fixtures as fx  # noqa: E402
from oracle import nlopt_path, traj_metrics  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))


def window_case(cfg):
    """the box and the extra obstacle points of the windowed-update fixture (also used by the tests)"""
    wp = np.asarray(cfg["waypoints"], dtype=np.float64)
    lo, hi = wp.min(axis=0) - 0.7, wp.max(axis=0) + 0.4
    extra = wp[1:-1] + np.array([0.35, -0.2, 0.1])
    return lo, hi, extra


def main():
    ev, edt, sol = {}, {}, {}
    for name in ("text_input", "opti_node"):
        cfg = fx.load_cfg(name)
        om = oracle_map_from_cfg(cfg)
        d = om.distance()
        edt[name + "_grid"] = np.array(om.grid)
        edt[name + "_occupied"] = np.array(int(om.occupancy().sum()))
        edt[name + "_sha256"] = np.frombuffer(hashlib.sha256(d.tobytes()).digest(), dtype=np.uint8)
        edt[name + "_max"] = np.array(d.max())
        edt[name + "_slice_z10"] = d[::8, ::8, 10].copy()
        # windowed update (setUpdateRange, sdf_map.cpp:244-264): clear a box around the path, add obstacles, rebuild the box
        lo, hi, extra = window_case(cfg)
        ow = oracle_map_from_cfg(cfg)
        ow.reset_range(lo, hi), ow.set_occupancy(extra), ow.set_update_range(lo, hi), ow.update_esdf()
        dw = ow.distance()
        edt[name + "_window_sha256"] = np.frombuffer(hashlib.sha256(dw.tobytes()).digest(), dtype=np.uint8)
        edt[name + "_window_changed"] = np.array(int((dw != d).sum()))
        opt = oracle_problem(cfg, om, cfg["waypoints"])
        lb, ub = opt.bounds()
        X = fx.multi_starts(opt.x0(), lb, ub, 8, seed=21)
        ev[name + "_X"] = X
        ev[name + "_seg_time"] = opt.seg_time()
        for step in (1, 2):
            c, g, oob, _ = opt.eval_batch(X, step)
            assert oob == 0
            ev["%s_cost_step%d" % (name, step)] = c
            ev["%s_grad_step%d" % (name, step)] = g
        # post-hoc metrics of the same starts (PolynomialTraj as opti_node drives it, opti_node.cpp:104-144)
        ev[name + "_metrics"] = np.stack([traj_metrics(opt.coeff_from_x(x), opt.seg_time()) for x in X])
        if nlopt_path():
            for maxeval in (40, 100):
                o2 = oracle_problem(cfg, om, cfg["waypoints"])
                x, minf, res, nev = o2.solve(step=2, maxeval=maxeval)
                sol["%s_x_maxeval%d" % (name, maxeval)] = x
                sol["%s_minf_maxeval%d" % (name, maxeval)] = np.array(minf)
                sol["%s_coeff_maxeval%d" % (name, maxeval)] = o2.coeff()
    np.savez_compressed(os.path.join(OUT, "fixtures_eval.npz"), **ev)
    np.savez_compressed(os.path.join(OUT, "fixtures_edt.npz"), **edt)
    if sol:
        np.savez_compressed(os.path.join(OUT, "fixtures_solve.npz"), **sol)
    for f in sorted(os.listdir(OUT)):
        if f.endswith(".npz"):
            print(f, os.path.getsize(os.path.join(OUT, f)), "bytes")


if __name__ == "__main__":
    main()
