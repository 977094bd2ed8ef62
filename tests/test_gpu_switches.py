"""The environment switches that select alternative code paths give the same results as the
defaults: every variant runs tests/_switch_probe.py in its own process (most switches are read once per process) and
the probe compares gradient / RK4 / flat-path results with the C oracle.  Tolerance 1e-12 relative."""
import json
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PROBE = os.path.join(ROOT, "tests", "_switch_probe.py")
TOL = 1e-12


@pytest.mark.timeout(300)
@pytest.mark.parametrize("env,expect", [
    ({}, {"rowtile": True}),                                              # defaults (reference for the variants)
    ({"GSG_ROWTILE": "0"}, {"rowtile": False}),                           # round-1 long-pole kernels on a big-item plan
    ({"GSG_RT_POOL": "1", "GSG_RT_GRID": "0", "GSG_RT_C": "2"}, {"rowtile": True}),   # one CTA per tile, high-priority pool, C = 2
    ({"GSG_RHS_SERIAL": "1", "GSG_NO_GRAPH": "1", "GSG_NO_PAIR": "1"}, {"rowtile": True}),   # sweep-by-sweep RHS, eager flat steps, no pair fusion
    ({"GSG_FLAT": "0", "GSG_LAP_NOSQ": "1"}, {"rowtile": True}),          # small index sets through the tiled kernels
])
def test_switch_variants_match_the_oracle(env, expect):
    e = dict(os.environ)
    e.update(env)
    out = subprocess.run([sys.executable, PROBE], env=e, capture_output=True, text=True, timeout=280)
    assert out.returncode == 0, out.stderr[-2000:]
    res = json.loads(out.stdout.strip().splitlines()[-1])
    print(env, res)
    assert res["rowtile"] == expect["rowtile"]
    for key, err in res["errors"].items():
        assert err <= TOL, (env, key, err)
