"""Committed golden vectors (tests/golden/*.npz, minted by tests/golden/make_golden.py from the outputs of the REFERENCE
ITSELF: oracle/_ref = /root/reference/*.cpp compiled against the godot-lite stand-in, see the generator's header).

CPU: the oracle must reproduce them bit for bit, and so must oracle/_ref wherever it is available.
GPU: the CUDA path, through the C ABI, must match them at the north-star tolerance
     (samples within 1e-5 relative or below -110 dBFS; routing and integer parameter fields bit-exact)."""
import glob
import importlib.util
import os

import numpy as np
import pytest

import scenarios as S

HERE = os.path.dirname(os.path.abspath(__file__))
_spec = importlib.util.spec_from_file_location("make_golden", os.path.join(HERE, "golden", "make_golden.py"))
mg = importlib.util.module_from_spec(_spec)
_spec.loader.exec_module(mg)

NAMES = sorted(mg.SCENARIOS)
GAIN_RTOL = 2e-6


def _load(name):
    return np.load(os.path.join(HERE, "golden", name + ".npz"))


def test_every_scenario_has_a_fixture():
    have = sorted(os.path.splitext(os.path.basename(p))[0] for p in glob.glob(os.path.join(HERE, "golden", "*.npz")))
    assert have == NAMES


@pytest.mark.parametrize("name", NAMES)
def test_oracle_reproduces_golden_bit_exact(orc, name):
    want = _load(name)
    got = mg.pack(mg.run_oracle(name))
    for k in got:
        assert np.array_equal(got[k], want[k]), f"{name}: {k} changed"


@pytest.mark.parametrize("name", NAMES)
def test_reference_reproduces_golden_bit_exact(name):
    from oracle import ref
    if not ref.available():
        pytest.skip("oracle/_ref not built and /root/reference not present")
    want = _load(name)
    assert "oracle/_ref" in str(want["minted_by"])
    got = mg.pack(mg.run_reference(name))
    for k in got:
        assert np.array_equal(got[k], want[k]), f"{name}: {k} changed"


@pytest.mark.gpu
@pytest.mark.parametrize("name", NAMES)
def test_cuda_matches_golden(gas, name):
    want = _load(name)
    sc = mg.scenario(name)
    with gas.Mixer(**S.config_of(sc)) as m:
        launches0 = m.kernel_launches
        got = mg.pack(S.run(m, sc))
        assert m.kernel_launches > launches0, "no CUDA kernel was launched"
    for k in ("params_update_parameters", "params_n_bus", "params_bus"):
        assert np.array_equal(got[k], want[k]), f"{name}: {k} differs"
    for k in ("params_mix_volumes", "params_bus_volumes", "params_pitch_scale", "params_linear_attenuation",
              "params_attenuation_filter_cutoff_hz", "state_prev_mix_volumes"):
        np.testing.assert_allclose(got[k], want[k], rtol=GAIN_RTOL, atol=1e-9, err_msg=f"{name}: {k}")
    for b in range(want["bus"].shape[0]):
        assert np.array_equal(S.routing(got["bus"][b]), S.routing(want["bus"][b])), f"{name} block {b}: routing differs"
        ok, worst, nbad = S.sample_close(got["bus"][b], want["bus"][b])
        assert ok, f"{name} block {b}: {nbad} samples out of tolerance, worst abs err {worst:.3e}"
    flagged = np.zeros(want["peaks"].shape[1], dtype=bool)
    if sc["want_peak_every"]:
        flagged[:: sc["want_peak_every"]] = True
    ok, worst, nbad = S.sample_close(got["peaks"][:, flagged], want["peaks"][:, flagged])
    assert ok, f"{name}: peaks differ (worst {worst:.3e})"
