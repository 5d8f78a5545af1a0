"""The drop-in boundary proven from the reference side, in C++.

oracle/_ref/hadi_bridge_test (built by `make -C oracle bridge` where /root/reference exists; the binary travels to the
GPU box like the other oracle/_ref artefacts) is the reference's shipped driver test_calibration_european_multi_maturity
— unmodified sources, its own grid set-up, LM loop, normal equations and 5x5 solve — whose two solver entry points
(compute_jacobian_multi_maturity, compute_base_prices_multi_maturity) are interposed by definitions that forward to
libhadi.so through oracle/bridge/hadi_bridge.hpp.  What the driver prints must be the trajectory of the pure-reference
run (tests/golden/lm_multi_maturity.json)."""
import os
import re
import subprocess

import pytest

from conftest import ROOT, golden

BIN = os.path.join(ROOT, "oracle", "_ref", "hadi_bridge_test")


@pytest.mark.gpu
def test_reference_lm_driver_runs_on_libhadi_through_the_bridge(tmp_path):
    if not os.path.exists(BIN):
        pytest.skip("oracle/_ref/hadi_bridge_test not built (needs /root/reference at build time)")
    r = subprocess.run([BIN], cwd=tmp_path, capture_output=True, text=True, timeout=300)
    out = r.stdout
    assert r.returncode == 0, out[-2000:] + r.stderr[-2000:]
    G = golden("lm_multi_maturity.json")

    def grab(label):
        m = re.search(re.escape(label) + r"\s*=\s*([-+0-9.eE]+)", out)
        assert m, (label, out[-1500:])
        return float(m.group(1))

    got = [grab("κ"), grab("η"), grab("σ"), grab("ρ"), grab("v₀")]
    assert [repr(x) for x in got] == G["params"]
    assert repr(grab("final error")) == G["trajectory"][-1]["err"]
    assert int(grab("total iterations")) == G["iterations"]
    m = re.search(r"BRIDGE jacobian_calls (\d+) price_calls (\d+) hadi_kernel_launches (\d+) exact_reruns (\d+)", out)
    assert m, out[-500:]
    jac, price, launches, reruns = (int(x) for x in m.groups())
    assert jac == G["iterations"] and price >= G["iterations"] - 1 and launches == jac + price and reruns == 0


def test_bridge_header_is_what_integration_md_documents():
    """INTEGRATION.md points at the compiled bridge, not at a sketch."""
    txt = open(os.path.join(ROOT, "INTEGRATION.md")).read()
    assert "oracle/bridge/hadi_bridge.hpp" in txt and "oracle/bridge/bridge_main.cpp" in txt
