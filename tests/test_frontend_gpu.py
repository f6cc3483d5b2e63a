"""GPU: the reference-shaped front end (`Simulation.load_data / set_engine / add_condition /
autorun`, README mini test of the reference, tests/00_mini_test.py) end to end on the device."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _write_star(tmp_path, seed=11, n=160, nins=2, kplan=1):
    from astroemperor_b200.synth import make_synthetic_rv
    d = tmp_path / "datafiles" / "synthstar" / "RV"
    d.mkdir(parents=True)
    for i, (t, rv, erv) in enumerate(make_synthetic_rv(seed=seed, n=n, nins=nins, kplan=kplan, span=300.0)):
        np.savetxt(d / f"synthstar_ins{i + 1}.vels", np.column_stack([t, rv, erv]), fmt="%.17g")
    return str(tmp_path) + os.sep


def test_simulation_autorun_finds_the_planet(tmp_path):
    from astroemperor_b200.frontend import Simulation
    sim = Simulation()
    sim.read_loc = _write_star(tmp_path)
    sim.load_data("synthstar")
    sim.set_engine("reddemcee")
    sim.engine_config["setup"] = [6, 128, 300, 2]
    sim.engine_config["progress"] = False
    sim.keplerian_parameterisation = 1
    sim.seed = 5
    sim.add_condition(["Period 1", "limits", [8, 20]])
    hist = sim.autorun(0, 1)
    assert [h["k"] for h in hist] == [0, 1]
    assert hist[1]["BIC"] < hist[0]["BIC"] - 5          # the 50 m/s, 12.3 d planet is overwhelming
    ch = sim.sampler.get_chain(discard=150, flat=True)[0]  # cold chain
    ll = sim.sampler.get_log_like(discard=150, flat=True)[0]
    best = ch[np.argmax(ll)]
    assert abs(best[0] - 12.3) < 0.1 and abs(best[1] - 50.0) < 5.0, best[:2]
    assert sim.sampler.get_smd().shape[1] == 5          # sampler.D_ was set like emp.py:595-602
    # the callables the parent re-imports for statistics (emp.py:818-828)
    eng = sim.sampler.engine
    model, err2 = eng.my_model(best)
    assert model.shape == (len(sim.data),) and np.all(err2 > 0)
    assert np.isclose(eng.my_likelihood(best), np.max(ll), rtol=1e-12)


def test_freeze_variant_stops_the_ladder_adaptation(tmp_path):
    """run_config adaptation_batches / adaptation_nsweeps select support/endit_freeze1.scr in the reference:
    `burnin` adaptation sweeps, sampler.select_adjustment('00'), then production sweeps on a frozen ladder."""
    from astroemperor_b200.frontend import Simulation
    sim = Simulation()
    sim.read_loc = _write_star(tmp_path)
    sim.load_data("synthstar")
    sim.set_engine("reddemcee")
    sim.engine_config["setup"] = [5, 64, 40, 1]
    sim.engine_config["progress"] = False
    sim.engine_config["adapt_tau"] = 20
    sim.run_config.update(burnin=15, adaptation_batches=1, adaptation_nsweeps=15)
    sim.seed = 3
    s = sim.run(1)
    b = s.get_betas()
    assert b.shape == (40, 5)
    assert np.any(b[14] != b[0])                      # the ladder moved during the adaptation phase
    assert np.all(b[15:] == b[14])                    # and is frozen afterwards
    assert s.get_chain().shape[1] == 40
    # the view EMPEROR's generated save section reads (emp.py:727-761)
    be = s.backend
    assert be.iteration == be[0].iteration == 40 and len(be) == 5
    assert be.tsw_history_bool and be.tsw_history.shape == (40, 4) and be.smd_history.shape == (40, 4)
    for t in (0, 4):
        assert np.array_equal(be[t].get_chain(), s.get_chain()[t]) and be[t].get_chain().shape == (40, 64, s.ndim)
        assert np.array_equal(be[t].get_log_like(), s.get_log_like()[t])
        assert np.array_equal(be[t].get_log_prob(), s.get_log_prob()[t])
        assert np.array_equal(be[t].get_betas(), b[:, t]) and be[t].accepted.shape == (64,)
    assert be[0].accepted.sum() == round(s.acceptance_fraction[0].sum() * 40)
    with pytest.raises(NotImplementedError):
        s.select_adjustment("11")


def test_unsupported_engine_and_missing_data():
    from astroemperor_b200.frontend import Simulation
    from astroemperor_b200.modelspec import UnsupportedModelError
    sim = Simulation()
    with pytest.raises(UnsupportedModelError):
        sim.set_engine("dynesty")
    sim.set_engine("reddemcee")
    with pytest.raises(RuntimeError):
        sim.build_model(1)
