"""GPU parity: a seeded random sweep over the whole dispatch (tools/fuzz_parity.py) — sizes, horizons, batch widths,
per-knot stage-row patterns, Hessian modes, explicit D2, SOC and LTI flags — against the oracle and the refined truth.
Round 2 ran 120,000 cases of it (profiles/r2_fuzz_summary.txt); the suite keeps 250."""
import os
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools"))


@pytest.mark.parametrize("seed", [11, 12])
def test_random_sweep_of_the_dispatch(handle, oracle_mod, seed):
    import fuzz_parity as fz
    from oracle import dense_kkt
    from lqr_b200 import ops, problems
    fails, kernels = [], set()
    for case in range(125):
        rng = np.random.default_rng([seed, case])
        if rng.integers(0, 4) == 0:
            d = fz.run_riccati_case(rng, handle, oracle_mod, ops, problems, case)
        else:
            d = fz.run_kkt_case(rng, handle, oracle_mod, dense_kkt, ops, case)
        kernels.add(d.get("kernel", "?").split("<")[0])
        if "fail" in d:
            fails.append(d)
    assert not fails, fails[:3]
    assert {"kkt_coop", "kkt_tpi", "kkt_wp_dmma", "riccati_tpi", "riccati_dmma", "riccati_cta_dmma"} <= kernels, kernels
