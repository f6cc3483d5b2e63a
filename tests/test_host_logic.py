"""CPU tests of the host-side logic: draw protocol, front-end mirror vs the reference's own
model description, data loading, ladder sharding over a 2-rank gloo group."""
import json
import os

import numpy as np
import pytest

from conftest import REPO, load_golden


# ---------------------------------------------------------------- draws ----
@pytest.mark.parametrize("T,W,nsteps,threads", [(3, 10, 2, 1), (10, 512, 1, 4), (5, 2, 3, 2), (2, 100, 2, 8),
                                                (7, 1026, 1, 3)])
def test_native_draws_are_bit_identical_to_numpy_randomstate(built_lib, T, W, nsteps, threads):
    """csrc/emp_draws.cpp restates MT19937 + the legacy RandomState algorithms (random_sample, shuffle,
    permutation, randint): every array of a sweep equals what numpy.random.RandomState produces, sweep after
    sweep, for whole ladders and for the shards of a sharded one (pure host code: no GPU needed)."""
    from astroemperor_b200.draws import DrawStreams, draw_sweep
    a, b = DrawStreams(5, T, native=False), DrawStreams(5, T, native=True, n_threads=threads)
    for it in range(4):
        da, db = draw_sweep(a, W, 7, nsteps), draw_sweep(b, W, 7, nsteps)
        for f in da.FIELDS:
            assert np.array_equal(getattr(da, f), getattr(db, f)), (f, it)
    for sl in (slice(1, T, 2), slice(0, T, 2)):
        da = draw_sweep(a, W, 7, nsteps, temps=sl, swap_rows=range(T)[sl])
        db = draw_sweep(b, W, 7, nsteps, temps=sl, swap_rows=range(T)[sl])
        assert db.sharded_swap and all(np.array_equal(getattr(da, f), getattr(db, f)) for f in da.FIELDS)
    # draws written straight into caller-owned buffers (the sampler's pinned staging views)
    from astroemperor_b200.draws import sweep_shapes
    out = {f: np.full(shp, 7, dtype=dt) for f, shp, dt in sweep_shapes(T, W, nsteps, T - 1)}
    da, db = draw_sweep(a, W, 7, nsteps), draw_sweep(b, W, 7, nsteps, out=out)
    assert db.zz is out["zz"] and all(np.array_equal(getattr(da, f), out[f]) for f in da.FIELDS)


@pytest.mark.parametrize("T,W,nsteps,k,threads", [(3, 20, 2, 5, 3), (2, 100, 1, 64, 8), (1, 16, 1, 4, 2)])
def test_native_draws_of_a_chunk_equal_sweep_by_sweep(built_lib, T, W, nsteps, k, threads):
    """emp_draws_sweeps (the k sweeps of a chunk in one parallel region, sweep q at q * stride bytes) gives exactly
    the bytes of k successive emp_draws_sweep calls: a chunked run_mcmc consumes the same random streams."""
    from astroemperor_b200 import _lib
    from astroemperor_b200.draws import DrawStreams, sweep_shapes
    L = _lib.lib()
    shapes = sweep_shapes(T, W, nsteps, max(T - 1, 1))
    offs, tot = [], 0
    for f, shp, dt in shapes:
        offs.append(tot)
        tot += (int(np.prod(shp)) * np.dtype(dt).itemsize + 255) // 256 * 256

    def run(chunked):
        streams = DrawStreams(5, T, native=True, n_threads=threads)   # owns the C generator: keep it alive
        h = streams.handle()
        buf = np.zeros(k * tot, dtype=np.uint8)
        ts = np.arange(T, dtype=np.int32)
        rs = np.array([T + j for j in range(T - 1)], dtype=np.int32)
        p = {f: buf.ctypes.data + o for (f, _, _), o in zip(shapes, offs)}
        if chunked:
            _lib.check(L.emp_draws_sweeps(h, k, tot, ts.ctypes.data, T, W, nsteps, p["half_idx"], p["zz"], p["rint"],
                                          p["lnu"], rs.ctypes.data, len(rs), p["perm"], p["lnu_swap"]))
        else:
            for q in range(k):
                o = q * tot
                _lib.check(L.emp_draws_sweep(h, ts.ctypes.data, T, W, nsteps, p["half_idx"] + o, p["zz"] + o,
                                             p["rint"] + o, p["lnu"] + o, rs.ctypes.data, len(rs), p["perm"] + o,
                                             p["lnu_swap"] + o))
        return buf
    a, b = run(True), run(False)
    assert a.any() and np.array_equal(a, b)


@pytest.mark.timeout(120)
def test_native_draw_pool_survives_oversubscription(built_lib):
    """Tens of thousands of back-to-back tiny sweeps on a pool with four times more workers than this host has
    cores: workers wake up late, after the run they were woken for is over.  (The first pool shared its item counter
    between runs: a late worker could claim item 0 of the NEXT run with the previous run's cleared function pointer
    and drop it, and the caller waited for ever — seen once as a hung GPU test session.)"""
    from astroemperor_b200.draws import DrawStreams, draw_sweep
    st = DrawStreams(3, 2, native=True, n_threads=4 * (os.cpu_count() or 8))
    ref = DrawStreams(3, 2, native=False)
    for it in range(40000):
        d = draw_sweep(st, 4, 3, 1)
        if it % 8000 == 0:
            r = draw_sweep(ref, 4, 3, 1)
            for _ in range(7999 if it + 8000 <= 40000 else 0):
                r2 = draw_sweep(ref, 4, 3, 1)   # keep the NumPy streams in step
            assert np.array_equal(d.zz, r.zz) and np.array_equal(d.perm, r.perm)


def test_deterministic_exp_of_the_ladder_adaptation():
    """oracle exp_det (replayed operation by operation on the device, emp_pt.cuh) is within 1 ulp of exp; NumPy's own
    exp is not correctly rounded either, so "the reference's np.exp" is only defined to that level; over 2000
    adaptations the two ladders stay within 1e-13."""
    import mpmath as mp
    from oracle.pt_oracle import adapt_ladder, exp_det
    rng = np.random.default_rng(0)
    x = np.concatenate([rng.uniform(-1, 1, 3000), rng.uniform(-30, 30, 500), [0.0, 1e-300, -1e-9, 0.34657359, -0.34657359]])
    mp.mp.dps = 40
    truth = np.array([float(mp.exp(mp.mpf(float(v)))) for v in x])
    ulp = np.spacing(truth)
    assert np.max(np.abs(exp_det(x) - truth) / ulp) <= 1.0
    assert np.max(np.abs(np.exp(x) - truth) / ulp) <= 1.0
    assert exp_det(np.array([0.0]))[0] == 1.0
    b1 = b2 = np.geomspace(1.0, 1e-3, 12)
    for time in range(1, 2001):
        ratios = 0.3 + 0.05 * rng.standard_normal(11)
        b1 = adapt_ladder(b1, ratios, time, 1000, 1, exp=exp_det)
        b2 = adapt_ladder(b2, ratios, time, 1000, 1, exp=np.exp)
    assert np.max(np.abs(b1 - b2) / b2) < 1e-13


def test_draw_sweep_protocol():
    from astroemperor_b200.draws import DrawStreams, draw_sweep
    T, W, nd, ns = 3, 10, 4, 2
    d = draw_sweep(DrawStreams(1, T), W, nd, ns)
    assert d.half_idx.shape == (ns, T, 2, 5) and d.perm.shape == (T - 1, 2, W)
    for s in range(ns):
        for t in range(T):
            both = np.sort(np.concatenate([d.half_idx[s, t, 0], d.half_idx[s, t, 1]]))
            assert np.array_equal(both, np.arange(W))
            assert np.all(np.diff(d.half_idx[s, t, 0]) > 0)
    assert np.all((d.zz >= 0.5) & (d.zz <= 2.0))  # g(z) support [1/a, a]
    assert np.array_equal(d.factors, (nd - 1.0) * np.log(d.zz))
    assert np.all(d.rint >= 0) and np.all(d.rint < 5) and np.all(d.lnu <= 0)
    for j in range(T - 1):
        assert np.array_equal(d.perm[j, 0], np.arange(W))  # pairs are listed by their slot in the warmer row
        assert np.array_equal(np.sort(d.perm[j, 1]), np.arange(W))
    # ... which is a relabelling of the reference's draw order (permutation, permutation, uniform per pair)
    from astroemperor_b200.draws import relabel_swap_draws
    from oracle.pt_oracle import swap_sweep
    r = DrawStreams(1, T).swap_pair[1]
    ip, i1p, u = r.permutation(W), r.permutation(W), r.uniform(size=W)
    assert np.array_equal(d.perm[1, 1][ip], i1p) and np.array_equal(d.lnu_swap[1][ip], np.log(u))
    # and the swap sweep it drives is the same sweep: same counts, same plan, same state
    rng = np.random.RandomState(0)
    raw = [(q.permutation(W), q.permutation(W), q.uniform(size=W)) for q in DrawStreams(1, T).swap_pair]
    p_a, p_b = rng.normal(size=(T, W, 3)), None
    p_b = p_a.copy()
    ll = rng.normal(size=(T, W)) * 3
    betas = np.linspace(1, 0.2, T)
    perm_raw = np.array([[a, b] for a, b, _ in raw]).astype(np.int32)
    lnu_raw = np.log(np.array([c for _, _, c in raw]))
    na, src_a, _ = swap_sweep(p_a, ll.copy(), np.zeros((T, W)), betas, perm_raw, lnu_raw)
    nb, src_b, _ = swap_sweep(p_b, ll.copy(), np.zeros((T, W)), betas, d.perm, d.lnu_swap)
    assert np.array_equal(na, nb) and np.array_equal(src_a, src_b) and np.array_equal(p_a, p_b)
    # each temperature's stream is consumed in emcee's order: re-derive temperature 1 by hand
    r = DrawStreams(1, T).temp[1]
    inds = np.arange(W) % 2
    r.shuffle(inds)
    assert np.array_equal(d.half_idx[0, 1, 0], np.flatnonzero(inds == 0))
    assert np.array_equal(d.zz[0, 1, 0], ((2.0 - 1.0) * r.rand(5) + 1) ** 2.0 / 2.0)
    assert np.array_equal(d.rint[0, 1, 0], r.randint(5, size=(5,)))
    assert np.array_equal(d.lnu[0, 1, 0], np.log(r.rand(5)))
    d2 = draw_sweep(DrawStreams(1, T), W, nd, ns)
    assert np.array_equal(d.zz, d2.zz) and np.array_equal(d.perm, d2.perm)
    # a shard draws exactly its own temperatures' part and the same swap draws
    d3 = draw_sweep(DrawStreams(1, T), W, nd, ns, temps=slice(1, 3))
    assert np.array_equal(d3.zz, d.zz[:, 1:3]) and np.array_equal(d3.perm, d.perm)
    assert np.array_equal(d3.lnu_swap, d.lnu_swap)
    # a rank of a sharded ladder draws only the pair rows of its temperatures (per-pair streams); row T-1 pads
    d4 = draw_sweep(DrawStreams(1, T), W, nd, ns, temps=slice(1, T, 2), swap_rows=range(T)[1:T:2])
    d5 = draw_sweep(DrawStreams(1, T), W, nd, ns, temps=slice(0, T, 2), swap_rows=range(T)[0:T:2])
    assert np.array_equal(d4.perm[0], d.perm[1]) and np.array_equal(d5.perm[0], d.perm[0])
    assert np.array_equal(d5.lnu_swap[0], d.lnu_swap[0]) and not d5.perm[1].any() and not d5.lnu_swap[1].any()
    with pytest.raises(ValueError):
        draw_sweep(DrawStreams(1, 2), 7, 3, 1)


def chain_swap_plan(logl, betas, perm, lnu_swap):
    """NumPy restatement of csrc/emp_pt.cuh::pt_swap_plan_chain_kernel: the hot -> cold swap sweep as W independent
    chains.  Chain k starts in slot k of the hottest row carrying that walker (source index, logL); at pair j it moves
    from slot s of row j+1 to slot b = partner_j[s] of row j, decides the swap with the ORIGINAL occupant of (j, b)
    and carries whoever sits in (j, b) afterwards.  No chain reads anything another chain wrote."""
    T, W = logl.shape
    src = np.empty((T, W), dtype=np.int32)
    n_acc = np.zeros(max(T - 1, 0), dtype=np.int32)
    for k in range(W):                      # one device thread per chain
        s, hl, hs = k, logl[T - 1, k], (T - 1) * W + k
        for j in range(T - 2, -1, -1):
            assert perm[j, 0, s] == s       # pairs are listed by their slot in the warmer row
            b = perm[j, 1, s]
            lb, cold = logl[j, b], j * W + b
            acc = (betas[j] - betas[j + 1]) * (hl - lb) > lnu_swap[j, s]
            src[j + 1, s] = cold if acc else hs
            if not acc:
                hl, hs = lb, cold
            n_acc[j] += int(acc)
            s = b
        src[0, s] = hs
    return n_acc, src


@pytest.mark.parametrize("T,W,scale", [(2, 6, 1.0), (7, 40, 3.0), (16, 64, 0.3), (33, 10, 30.0)])
def test_swap_sweep_is_W_independent_chains(T, W, scale):
    """The claim behind the chain plan kernel: following each walker's path down the ladder gives
    the plan and the swap counts of the reference-order sweep (oracle/pt_oracle.py::swap_sweep, sequential over the
    pairs with whole rows exchanged) bit for bit, for any acceptance rate."""
    from astroemperor_b200.draws import DrawStreams, draw_sweep
    from oracle.pt_oracle import swap_sweep
    rng = np.random.RandomState(T * 1000 + W)
    d = draw_sweep(DrawStreams(9, T), W, 3, 1)
    logl = rng.normal(size=(T, W)) * scale - 50.0
    logl[rng.randint(T), rng.randint(W)] = -np.inf   # a walker outside the support never blocks a chain
    betas = np.geomspace(1.0, 1e-3, T)
    p = rng.normal(size=(T, W, 2))
    n_ref, src_ref, _ = swap_sweep(p.copy(), logl.copy(), np.zeros((T, W)), betas, d.perm, d.lnu_swap)
    n_chain, src_chain = chain_swap_plan(logl, betas, d.perm, d.lnu_swap)
    assert np.array_equal(n_ref, n_chain) and np.array_equal(src_ref, src_chain)
    assert 0 < n_ref.sum() < (T - 1) * W or T == 2   # both outcomes occur


def test_default_ladder_matches_the_ladders_the_reference_notebooks_print():
    """reddemcee's default ladder (EMPEROR passes betas=None, emp.py:2372) is ptemcee's default_beta_ladder.  The
    reference's notebooks print ladders whose FIXED ends pin the table: 'Beta Detail' ends in 5.057e-10 for the
    2-parameter model with 12 temperatures (tests/00_mini_test.ipynb) and in 2.478e-08 with 10
    (tests/01_51peg_basic.ipynb) — the hottest rung is not adapted —, and tests/quickstart.ipynb prints
    [1.0, 0.4002] for the 7-parameter model with 2 temperatures (no adaptation with fewer than 3)."""
    from astroemperor_b200.draws import _TSTEP, default_betas
    assert float("%.4g" % default_betas(2, 12)[-1]) == 5.057e-10
    assert float("%.4g" % default_betas(2, 10)[-1]) == 2.478e-08
    assert [float("%.4g" % b) for b in default_betas(7, 2)] == [1.0, 0.4002]
    assert len(_TSTEP) == 100 and np.all(np.diff(_TSTEP) < 0) and _TSTEP[-1] > 1.0
    big = default_betas(400, 5)   # beyond the table: 1 + 2 sqrt(ln 4 / ndim)
    assert np.isclose(big[1], 1.0 / (1.0 + 2.0 * np.sqrt(np.log(4.0)) / 20.0))
    assert default_betas(35, 32)[0] == 1.0 and np.all(np.diff(default_betas(35, 32)) < 0)


def test_initial_positions_follow_set_init():
    from astroemperor_b200.draws import initial_positions
    g, spec = load_golden("mini_51peg_k1_p1")
    p = initial_positions(np.random.RandomState(0), spec, 2, 64)
    fp = spec.free_params()
    assert p.shape == (2, 64, len(fp))
    # Period 1 has init_pos [4.1, 4.3]; Ecc_sin is a 'hou' parameter: range shrunk by 0.707
    assert p[..., 0].min() >= 4.1 and p[..., 0].max() <= 4.3
    assert np.abs(p[..., 3]).max() <= 0.707 + 1e-12
    for j, par in enumerate(fp):
        assert p[..., j].min() >= par.limits[0] - 1e-12 and p[..., j].max() <= par.limits[1] + 1e-12


# ------------------------------------------------- front-end mirror vs reference ----
MIRROR_CASES = {
    "c1_51peg_k1_p0": dict(kplan=1, parameterisation=0),
    "c1_51peg_k0": dict(kplan=0),
    "mini_51peg_k1_p1": dict(kplan=1, parameterisation=1, conditions=[
        ("Period 1", "limits", [3, 5]), ("Amplitude 1", "limits", [45, 60]), ("Offset 1", "limits", [-10., 10.]),
        ("Period 1", "init_pos", [4.1, 4.3]), ("Amplitude 1", "init_pos", [50, 60])]),
    "synth_k2_p2": dict(kplan=2, parameterisation=2), "synth_k2_p3": dict(kplan=2, parameterisation=3),
    "synth_k2_p4": dict(kplan=2, parameterisation=4), "synth_k2_p6": dict(kplan=2, parameterisation=6),
    "synth_k2_p7": dict(kplan=2, parameterisation=7),
    "synth_k1_p0_acc2_fixed": dict(kplan=1, acceleration=2, conditions=[("Eccentricity 1", "fixed", 0.1)]),
    "synth_k3_p1_ma1_perins": dict(kplan=3, parameterisation=1, moav={"order": 1, "global": False}),
    "synth_k1_p1_ma2_global": dict(kplan=1, parameterisation=1, moav={"order": 2, "global": True}),
    "synth_k1_p0_nojit": dict(kplan=1, jitter=False),
    "gj876_k2_p1": dict(kplan=2, parameterisation=1),
    "c4_synth5p_4ins_ma_global_n600": dict(kplan=5, moav={"order": 1, "global": True}),
    "synth_k1_p0_sinusoid": dict(kplan=1, sinusoid=1),
    "synth_k1_p1_magcycle_ma1_global": dict(kplan=1, parameterisation=1, moav={"order": 1, "global": True},
                                            magnetic_cycle=1, sinusoid=2),
    "synth_k1_p0_sai21": dict(kplan=1),
    "synth_k2_p1_sai03_ma1_global_sin": dict(kplan=2, parameterisation=1, moav={"order": 1, "global": True},
                                             sinusoid=1),
}


def _close(a, b):
    if isinstance(a, (list, tuple)):
        return len(a) == len(b) and all(_close(x, y) for x, y in zip(a, b))
    if isinstance(a, (int, float)) and isinstance(b, (int, float)):
        if np.isnan(a) and np.isnan(b):
            return True
        # limits are derived from the data; the reference computes them before its CSV
        # round trip (emp_model.py:337), which can move the data by 1 ulp
        return abs(a - b) <= 1e-13 * max(1.0, abs(a), abs(b))
    return a == b


@pytest.mark.parametrize("name", sorted(MIRROR_CASES))
def test_default_spec_matches_reference_model(name):
    """`default_spec` (our SmartSetter / block mirror) == the spec extracted from the real
    reference's ReddModel for the same user calls."""
    from astroemperor_b200.data import RVData
    from astroemperor_b200.frontend import default_spec
    g, ref = load_golden(name)
    data = RVData(g["t"], g["y"], g["yerr"], g["flag"], float(g["common_t"]), [f"i{i}" for i in range(ref.nins)])
    if "sai" in g.files:  # activity columns per instrument, as the loader reports them
        data.sai = g["sai"]
        data.cornums = [int(np.any(g["sai"][g["flag"] == i + 1] != 0, axis=0).sum()) for i in range(ref.nins)]
    mine = default_spec(data, **MIRROR_CASES[name])
    a, b = json.loads(ref.to_json()), json.loads(mine.to_json())
    assert len(a["blocks"]) == len(b["blocks"]) and a["nins"] == b["nins"]
    for ba, bb in zip(a["blocks"], b["blocks"]):
        for key in ba:
            if key == "params" or key == "additional":
                assert len(ba[key]) == len(bb[key])
                for pa, pb in zip(ba[key], bb[key]):
                    for f in pa:
                        assert _close(pa[f], pb[f]), (name, ba["type_"], pa.get("name"), f, pa[f], pb[f])
            else:
                assert ba[key] == bb[key], (name, key)
    # and the prior widths handed to the sampler (emp.py:595-602 sampler.D_)
    assert np.allclose(mine.prior_widths(), g["D_"], rtol=1e-13)


def test_load_rv_folder_matches_datawrapper(tmp_path):
    """File loading reproduces what the reference's DataWrapper produced for the same files
    (the golden case was generated from these exact synthetic tables)."""
    from astroemperor_b200.data import load_rv_folder
    from astroemperor_b200.synth import make_synthetic_rv
    g, _ = load_golden("c2_synth3p_2ins_n400")
    d = tmp_path / "datafiles" / "star" / "RV"
    d.mkdir(parents=True)
    for i, (t, rv, erv) in enumerate(make_synthetic_rv(seed=2, n=400, nins=2, kplan=3)):
        np.savetxt(d / f"star_ins{i + 1}.vels", np.column_stack([t, rv, erv]), fmt="%.17g")
    data = load_rv_folder(str(d) + os.sep)
    assert np.array_equal(data.flag, g["flag"])
    assert np.allclose(data.t, g["t"], rtol=0, atol=1e-9) and np.allclose(data.y, g["y"], rtol=1e-14, atol=1e-13)
    # the reference's temp_data.csv round trip (emp_model.py:337) can move values by 1 ulp
    assert np.allclose(data.yerr, g["yerr"], rtol=1e-15, atol=0) and data.common_t == float(g["common_t"])


def test_load_rv_folder_activity_columns_match_datawrapper(tmp_path):
    """switch_SA: the activity columns after eRV get the reference's per-file normalisation
    (qol_utils.py:88-93) and are zero outside their instrument; compared with the SAI{j}_ arrays the
    real generator's script loaded for the same files."""
    from astroemperor_b200.data import load_rv_folder
    from astroemperor_b200.synth import add_activity_columns, make_synthetic_rv
    g, spec = load_golden("synth_k1_p0_sai21")
    d = tmp_path / "datafiles" / "star" / "RV"
    d.mkdir(parents=True)
    files = add_activity_columns(make_synthetic_rv(seed=9, n=120, nins=2, kplan=1), [2, 1], 9)
    for i, (t, rv, erv, act) in enumerate(files):
        np.savetxt(d / f"star_ins{i + 1}.vels", np.column_stack([t, rv, erv, act]), fmt="%.17g")
    data = load_rv_folder(str(d) + os.sep, switch_SA=True)
    assert data.cornums == [2, 1] and data.sai.shape == g["sai"].shape
    assert np.array_equal(data.sai == 0, g["sai"] == 0)
    assert np.allclose(data.sai, g["sai"], rtol=1e-13, atol=1e-12)
    assert load_rv_folder(str(d) + os.sep).sai is None  # the default drops them (emp.py:2310-2312)
    cm = spec.compile()
    assert cm.sai_count == [2, 1] and cm.to_c().n_sai == 3


def test_unsupported_blocks_raise():
    from astroemperor_b200.modelspec import BlockSpec, ModelSpec, ParamSpec, UnsupportedModelError
    g, spec = load_golden("c1_51peg_k1_p0")
    spec.blocks.append(BlockSpec(type_="Sinusoid", params=[ParamSpec("Period", limits=[1, 2], prargs=0.0)]))
    with pytest.raises(UnsupportedModelError):
        spec.compile()


# ------------------------------------------------------------ ladder sharding ----
def _dist_worker(rank, world, port, T, W, C, seed, layout, q):
    import torch
    import torch.distributed as td
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    td.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from astroemperor_b200.dist import LadderShard
        from oracle.pt_oracle import swap_sweep
        rng = np.random.RandomState(seed)
        rows = rng.normal(size=(T, W, C))
        ll = rng.normal(size=(T, W)) * 5
        betas = np.linspace(1, 0.05, T)
        perm = np.stack([np.stack([rng.permutation(W), rng.permutation(W)]) for _ in range(T - 1)]).astype(np.int32)
        lnu = np.log(rng.uniform(size=(T - 1, W)))
        p = rows.copy()
        _, src, _ = swap_sweep(p, ll.copy(), np.zeros((T, W)), betas, perm, lnu)  # p is now the expected result
        sh = LadderShard(T, layout=layout)
        assert sh.world == world and sh.n_local == T // world
        local = torch.from_numpy(rows[sh.local_slice].reshape(-1, C).copy())
        # all-gather of the local logL reproduces the global array
        ga = sh.all_gather_rows(torch.from_numpy(ll[sh.local_slice].copy()))
        assert np.array_equal(ga.numpy(), ll)
        staged, src_local = sh.exchange_rows(torch.from_numpy(src), local, W)
        new = staged[src_local.long()].numpy().reshape(sh.n_local, W, C)
        ok = np.array_equal(new, p[sh.local_slice])
        n_remote = staged.shape[0] - local.shape[0]
        # the all-gather variant of the same step (sampler._apply_plan, exchange="allgather")
        (allr,) = sh.all_gather_flat(local)
        sg = torch.from_numpy(src)[sh.local_slice].reshape(-1).long()
        st, sw = sg // W, sg % W
        pos = sh.owner_of_temp(st) * (sh.n_local * W) + sh.local_of_temp(st) * W + sw
        ok = ok and np.array_equal(allr[pos].numpy().reshape(sh.n_local, W, C), p[sh.local_slice])
        # sharded swap draws: each rank draws the pair rows of its own temperatures (per-pair streams);
        # the all-gather restores the ladder order and equals what a single process draws
        from astroemperor_b200.draws import DrawStreams, draw_sweep
        full = draw_sweep(DrawStreams(seed, T), W, 4, 1)
        mine = draw_sweep(DrawStreams(seed, T), W, 4, 1, temps=sh.local_slice, swap_rows=range(T)[sh.local_slice])
        assert mine.perm.shape == (sh.n_local, 2, W) and mine.zz.shape[1] == sh.n_local
        gp = sh.all_gather_rows(torch.from_numpy(mine.perm))[: T - 1].numpy()
        gu = sh.all_gather_rows(torch.from_numpy(mine.lnu_swap))[: T - 1].numpy()
        ok = ok and np.array_equal(gp, full.perm) and np.array_equal(gu, full.lnu_swap)
        ok = ok and np.array_equal(mine.zz, full.zz[:, sh.local_slice])
        # the C generator draws the same shard (host code of the C-ABI library)
        mine_n = draw_sweep(DrawStreams(seed, T, native=True, n_threads=2), W, 4, 1, temps=sh.local_slice,
                            swap_rows=range(T)[sh.local_slice])
        ok = ok and all(np.array_equal(getattr(mine, f), getattr(mine_n, f)) for f in mine.FIELDS)
        # the product's sharded swap (pt_apply_plan_kernel): every rank's (p | logl | logp) block is addressed through
        # a table of base pointers, peer HBM (CUDA IPC) or, exchange='allgather', one all-gather of the blocks; the
        # kernel's owner / local-row arithmetic restated on the gathered blocks
        lp_ = rng.normal(size=(T, W))
        blk = np.concatenate([rows[sh.local_slice].ravel(), ll[sh.local_slice].ravel(), lp_[sh.local_slice].ravel()])
        (gathered,) = sh.all_gather_flat(torch.from_numpy(blk))
        gathered = gathered.numpy().reshape(world, -1)
        n_row = sh.n_local * W
        for tl in range(sh.n_local):
            tg = sh.temp_of(rank, tl)
            s_ = src[tg]
            st, sw = s_ // W, s_ % W
            owner, srow = sh.owner_of_temp(st), sh.local_of_temp(st) * W + sw
            got_p = np.stack([gathered[o, r * C:(r + 1) * C] for o, r in zip(owner, srow)])
            got_ll = np.array([gathered[o, n_row * C + r] for o, r in zip(owner, srow)])
            ll_exp = ll.copy()
            ok = ok and np.array_equal(got_p, p[tg])
        q.put((rank, ok, n_remote))
    finally:
        td.destroy_process_group()


@pytest.mark.parametrize("T,W,layout", [(4, 16, "contiguous"), (6, 10, "contiguous"), (4, 16, "strided"),
                                        (8, 10, "strided")])
def test_sharded_swap_exchange_two_ranks_gloo(T, W, layout):
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000) + T + (17 if layout == "strided" else 0)
    procs = [ctx.Process(target=_dist_worker, args=(r, 2, port, T, W, 5, 3, layout, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert all(ok for _, ok, _ in res), res
    assert sum(n for _, _, n in res) > 0  # some rows really crossed the shard edge


# ------------------------------------------------------------ post-run reductions ----
def test_evidence_estimators_on_an_analytic_ladder():
    """logL ~ N(mu(beta), .) with <logL>_beta = -1/(beta + 0.1): int_0^1 = -ln(11).  Trapezoid and
    PCHIP thermodynamic integration bracket the analytic value; PCHIP is the closer one on a coarse
    geometric ladder, and both report an error of the size of their miss."""
    from astroemperor_b200.postproc import evidence_ti
    rng = np.random.default_rng(0)
    betas = 2.0 ** (-np.arange(12.0))
    logl = np.stack([-1.0 / (b + 0.1) + 1e-3 * rng.normal(size=4000) for b in betas])
    exact = -np.log(11.0)
    z_tr, e_tr = evidence_ti(logl, betas)
    z_pc, e_pc = evidence_ti(logl, betas, pchip=True)
    assert abs(z_pc - exact) < abs(z_tr - exact) < 0.1
    assert abs(z_pc - exact) < 0.02 and e_tr > 0 and e_pc > 0
    # the ladder order does not matter
    p = rng.permutation(len(betas))
    assert np.isclose(evidence_ti(logl[p], betas[p], pchip=True)[0], z_pc)


def test_integrated_autocorrelation_time_of_an_ar1_chain():
    """emcee's estimator (what sampler.get_autocorr_time mirrors, emp.py:1375-1385) on AR(1) walkers:
    tau = (1 + rho) / (1 - rho)."""
    from astroemperor_b200.postproc import integrated_time
    rng = np.random.default_rng(2)
    n, W, rho = 20000, 16, np.array([0.5, 0.9])
    x = np.zeros((n, W, 2))
    eps = rng.normal(size=(n, W, 2))
    for i in range(1, n):
        x[i] = rho * x[i - 1] + eps[i]
    tau = integrated_time(x, c=5, tol=50, quiet=True)
    assert np.allclose(tau, (1 + rho) / (1 - rho), rtol=0.15), tau
    with pytest.raises(RuntimeError):  # a chain much shorter than tol * tau is refused unless quiet
        integrated_time(x[:200], c=5, tol=50, quiet=False)


def test_stepping_stone_matches_thermodynamic_integration_on_gaussian_rungs():
    """For logL | beta ~ N(m, s^2) independent of beta both estimators have closed forms:
    TI = m, SS = sum_i [db_i m + db_i^2 s^2 / 2]."""
    from astroemperor_b200.postproc import evidence_ss, evidence_ti
    rng = np.random.default_rng(3)
    betas = np.linspace(0.0, 1.0, 21)
    m, sd = -12.0, 0.8
    logl = m + sd * rng.normal(size=(len(betas), 20000))
    z_ti, _ = evidence_ti(logl, betas)
    z_ss, e_ss = evidence_ss(logl, betas)
    db = np.diff(betas)
    assert abs(z_ti - m) < 0.02
    assert abs(z_ss - (m + 0.5 * sd ** 2 * np.sum(db ** 2))) < 0.02 and e_ss < 0.02


def test_backend_round_trip_without_a_device(tmp_path):
    """save_backend -> load_backend: the stored run answers the parent's read-back calls (emp.py:777-807,
    1375-1447) from disk.  The sampler is stubbed: only its getters are used."""
    from astroemperor_b200.postproc import load_backend, save_backend
    rng = np.random.default_rng(4)
    T, n, W, nd = 3, 40, 8, 2
    betas = np.array([1.0, 0.4, 0.1])

    class Stub:
        _n_steps = n
        acceptance_fraction = rng.integers(4, 20, (T, W)) / n

        def __init__(self):
            self.chain = rng.normal(size=(T, n, W, nd))
            self.ll = -np.abs(rng.normal(size=(T, n, W))) * 3
            self.lpost = self.ll * betas[:, None, None] - 1.0
            self.bh = np.tile(betas, (n, 1))
            self.tsw, self.smd = rng.uniform(size=(n, T - 1)), rng.uniform(size=(n, T - 1))
            self.betas = betas

        def get_chain(self, discard=0): return self.chain[:, discard:]
        def get_log_like(self, discard=0): return self.ll[:, discard:]
        def get_log_prob(self, discard=0): return self.lpost[:, discard:]
        def get_betas(self, discard=0): return self.bh[discard:]
        def get_tsw(self, discard=0): return self.tsw[discard:]
        def get_smd(self, discard=0): return self.smd[discard:]

    s = Stub()
    path = save_backend(s, str(tmp_path / "run"))
    # the reference's file layout (emp.py:781-786): <name>.h5 + one <name>_<t>.h5 per temperature, HDF5
    assert path.endswith(".h5") and all(os.path.exists(str(tmp_path / f"run_{t}.h5")) for t in range(T))
    assert open(path, "rb").read(8) == b"\x89HDF\r\n\x1a\n"
    r = load_backend(str(tmp_path / "run"))
    assert (r.ntemps, r.iteration, r.nwalkers, r.ndim) == (T, n, W, nd)
    assert np.array_equal(r.get_chain(), s.chain) and np.array_equal(r.betas, betas)
    assert r.get_chain(discard=10, thin=3, flat=True).shape == (T, 10 * W, nd)
    assert np.array_equal(r.get_log_like(flat=True)[1], s.ll[1].reshape(-1))
    assert np.allclose(r.acceptance_fraction, s.acceptance_fraction)
    be = r.backend
    assert be.iteration == n and len(be) == T and np.array_equal(be[2].get_chain(), s.chain[2])
    assert np.array_equal(be[-1].get_betas(), s.bh[:, 2]) and be[0].accepted.shape == (W,)
    z, e = r.get_evidence_ti(discard=5)
    assert np.isfinite(z) and np.isfinite(r.get_evidence_ss()[0])
    assert r.get_autocorr_time(quiet=True).shape == (T, nd)
