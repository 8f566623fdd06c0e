"""Freezes the oracle: final poses of small seeded workloads (rigid, kinematic chains, measured occlusion) as computed by
the strict oracle build today -> tests/golden/oracle_regression.npz. tests/test_oracle_golden.py compares against them so
that a later edit of the oracle (the checker of every GPU parity test) cannot drift unnoticed."""
import importlib
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle")); sys.path.insert(0, os.path.join(ROOT, "tests"))


def cases():
    synth = importlib.import_module("3dobjecttracking_b200.synth")
    from test_oracle_occlusion import occluded_workload
    yield "c2", synth.make_workload("c2", n_bodies=3, n_divides=3, seed=31)
    yield "c3", synth.make_workload("c3", n_bodies=3, n_divides=3, seed=32)
    yield "c5_projected", synth.make_chain_workload(n_chains=1, n_links=4, n_lines=100, n_points=100, n_divides=3, seed=33)
    yield "c5_constrained", synth.make_chain_workload(n_chains=1, n_links=4, n_lines=100, n_points=100, n_divides=3, seed=34,
                                                     variant="constrained")
    yield "c5_soft", synth.make_chain_workload(n_chains=1, n_links=3, n_lines=100, n_points=100, n_divides=3, seed=35,
                                               variant="constrained", soft=True)
    yield "occlusion", occluded_workload(n_bodies=2, seed=36)


def run(oracle, wl):
    t = oracle.OracleTracker(wl)
    t.start_modalities(0)
    for it in range(2):
        t.tracking_step(it)
        t.calculate_results(it)
    return t.get_poses()


if __name__ == "__main__":
    import oracle_py as oracle
    out = {name: run(oracle, wl) for name, wl in cases()}
    np.savez_compressed(os.path.join(HERE, "oracle_regression.npz"), **out)
    print({k: v.shape for k, v in out.items()})
