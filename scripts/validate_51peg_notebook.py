"""End-to-end check of the device sampler against the statistics the REAL reddemcee run printed in the
reference's notebook tests/00_mini_test.ipynb (cell 7 output; setup [12, 500, 3000, 1] on 51Peg, reddemcee 0.9):

  k = 0  Offset + Jitter (2 parameters, Offset 1 limits [-10, 10]):
         Beta Detail  [1.0, 0.5573, 0.3317, 0.1952, 0.1176, 0.06762, 0.03774, 0.01973, 0.009203, 0.003676, 0.001009, 5.057e-10]
         Mean Acceptance Fraction [0.714, 0.712, 0.712, 0.712, 0.711, 0.709, 0.703, 0.695, 0.687, 0.680, 0.676, 0.657]
         evidence -1333.713 +- 1.558, maximum posterior -1333.617, maximum likelihood -1308.795
  k = 1  + one Keplerian (parameterisation 1; Period [3, 5], Amplitude [45, 60]; ladder carried over from k = 0,
         emp.py:789-791; Jitter 1 upper limit 40.573, the value EMPEROR derived from the k = 0 posterior):
         Mean Acceptance Fraction [0.186 ... 0.234], evidence -908.390 +- 1.425, maximum posterior -881.705,
         maximum likelihood -869.480

The sampler packages are not installable here, so this is the one place where the restated stretch move / swap
sweep / ladder adaptation meet numbers produced by the real thing.  Runs on one GPU in about a second per model;
writes profiles/notebook_51peg_validation.json.   usage: python scripts/validate_51peg_notebook.py [out.json]
"""
import json
import os
import sys

import numpy as np

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
sys.path.insert(0, os.path.join(REPO, "tests"))

NOTEBOOK = {
    "k0": dict(betas=[1.0, 0.5573, 0.3317, 0.1952, 0.1176, 0.06762, 0.03774, 0.01973, 0.009203, 0.003676, 0.001009,
                      5.057e-10],
               acceptance=[0.714, 0.712, 0.712, 0.712, 0.711, 0.709, 0.703, 0.695, 0.687, 0.680, 0.676, 0.657],
               smd=[1.465, 1.820, 2.154, 2.452, 2.698, 2.901, 3.098, 3.259, 3.400, 3.563, 3.770],
               evidence=(-1333.713, 1.558), max_posterior=-1333.617, max_likelihood=-1308.795),
    "k1": dict(betas=[1.0, 0.4913, 0.2435, 0.1253, 0.06376, 0.03226, 0.01682, 0.01111, 0.006404, 0.002565, 0.0007305,
                      5.057e-10],
               acceptance=[0.186, 0.181, 0.182, 0.188, 0.200, 0.217, 0.201, 0.155, 0.169, 0.195, 0.222, 0.234],
               smd=[0.814, 1.080, 1.419, 1.862, 2.498, 3.236, 4.043, 4.140, 4.019, 4.007, 4.072],
               evidence=(-908.390, 1.425), max_posterior=-881.705, max_likelihood=-869.480),
}
SETUP = (12, 500, 3000, 1)   # ntemps, nwalkers, nsweeps, nsteps


def specs():
    """The two models of the notebook on the 51Peg data (the arrays of the golden fixture = DataWrapper's output)."""
    from conftest import load_golden
    from astroemperor_b200.data import RVData
    from astroemperor_b200.frontend import default_spec
    g, _ = load_golden("c1_51peg_k0")
    data = RVData(t=g["t"], y=g["y"], yerr=g["yerr"], flag=g["flag"], common_t=0.0, labels=["LICK"])
    c0 = [["Offset 1", "limits", [-10.0, 10.0]]]
    c1 = c0 + [["Period 1", "limits", [3, 5]], ["Amplitude 1", "limits", [45, 60]],
               ["Period 1", "init_pos", [4.1, 4.3]], ["Amplitude 1", "init_pos", [50, 60]],
               ["Jitter 1", "limits", [1e-05, 40.57293916203054]]]
    return g, default_spec(data, 0, conditions=c0), default_spec(data, 1, parameterisation=1, conditions=c1)


def run(g, spec, betas, seed):
    from astroemperor_b200.engine import LikelihoodEngine
    from astroemperor_b200.sampler import PTSampler
    T, W, nsweeps, nsteps = SETUP
    eng = LikelihoodEngine(spec, g["t"], g["y"], g["yerr"], g["flag"])
    samp = PTSampler(W, eng.ndim, eng, ntemps=T, seed=seed, betas=betas)   # adapt_tau=1000, adapt_nu=1, mode 0
    samp.D_ = spec.prior_widths()
    samp.run_mcmc(samp.initial_positions(spec), nsweeps=nsweeps, nsteps=nsteps)
    disc = nsweeps * nsteps // 2
    ll, lp = samp.get_log_like(), samp.get_log_prior()
    z_ti, e_ti = samp.get_evidence_ti(discard=disc)
    z_tp, e_tp = samp.get_evidence_ti(discard=disc, pchip=True)
    z_ss, e_ss = samp.get_evidence_ss(discard=disc)
    out = dict(ndim=eng.ndim, betas_initial=[float(b) for b in samp._betas_initial],
               betas=[float(b) for b in samp.betas],
               acceptance=[float(a) for a in samp.acceptance_fraction.mean(axis=1)],
               swap_rate=[float(x) for x in samp.get_tsw(discard=disc).mean(axis=0)],
               smd=[float(x) for x in samp.get_smd(discard=disc).mean(axis=0)],
               evidence_ti=(z_ti, e_ti), evidence_ti_pchip=(z_tp, e_tp), evidence_ss=(z_ss, e_ss),
               max_likelihood=float(ll[0].max()), max_posterior=float((ll[0] + lp[0]).max()),
               sweeps=nsweeps, nan=int(eng.nan_count()))
    final = samp.betas.copy()
    del samp
    eng.close()
    return out, final


def compare(name, got):
    nb = NOTEBOOK[name]
    acc = np.array(got["acceptance"]) - np.array(nb["acceptance"])
    lb = np.log(np.array(got["betas"][1:-1])) - np.log(np.array(nb["betas"][1:-1]))
    ev = {k: got[k][0] - nb["evidence"][0] for k in ("evidence_ti", "evidence_ti_pchip", "evidence_ss")}
    return dict(acceptance_max_abs_diff=float(np.max(np.abs(acc))), acceptance_cold_diff=float(acc[0]),
                ladder_max_abs_dlog_beta=float(np.max(np.abs(lb))), evidence_minus_notebook=ev,
                notebook_evidence_sigma=nb["evidence"][1],
                max_likelihood_minus_notebook=got["max_likelihood"] - nb["max_likelihood"],
                max_posterior_minus_notebook=got["max_posterior"] - nb["max_posterior"])


def main():
    from astroemperor_b200.draws import default_betas
    g, s0, s1 = specs()
    out = {"setup": SETUP, "notebook": NOTEBOOK,
           "source": "/root/reference/tests/00_mini_test.ipynb cell 7 (output of a reddemcee 0.9 run on 24 cores)"}
    r0, final0 = run(g, s0, default_betas(2, SETUP[0]), seed=1234)
    out["k0"], out["k0_vs_notebook"] = r0, compare("k0", r0)
    r1, _ = run(g, s1, final0, seed=1235)   # EMPEROR hands the adapted ladder of k = 0 to the next model
    out["k1"], out["k1_vs_notebook"] = r1, compare("k1", r1)
    path = sys.argv[1] if len(sys.argv) > 1 else os.path.join(REPO, "profiles", "notebook_51peg_validation.json")
    with open(path, "w") as fh:
        json.dump(out, fh, indent=1)
    for k in ("k0", "k1"):
        print(k, json.dumps(out[k + "_vs_notebook"]))
        print("   acceptance", np.round(out[k]["acceptance"], 3).tolist())
        print("   betas     ", [float("%.4g" % b) for b in out[k]["betas"]])


if __name__ == "__main__":
    main()
