import faulthandler, sys, os
faulthandler.dump_traceback_later(45, exit=True)
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
from conftest import load_golden
from astroemperor_b200.engine import LikelihoodEngine
from astroemperor_b200.sampler import PTSampler
g, spec = load_golden("c1_51peg_k1_p0")
eng = LikelihoodEngine(spec, g["t"], g["y"], g["yerr"], g["flag"])
samp = PTSampler(100, eng.ndim, eng, ntemps=2, seed=21)
samp.D_ = spec.prior_widths()
p0 = samp.initial_positions(spec)
print("chunk len", samp._chunk_len(1), flush=True)
samp._init_state(p0)
lay = samp._chunk_layout(64, 1)
print("layout ok", flush=True)
samp._alloc_store(200); samp._alloc_hist(200)
samp._chunk_draw(lay, 0, 37)
print("draw ok", flush=True)
samp._chunk_launch(lay, 0, 37)
print("launch ok", flush=True)
import torch
torch.cuda.synchronize()
print("sync ok", samp.iteration, flush=True)
samp.run_mcmc(None, nsweeps=130, nsteps=1)
print("run ok", samp.iteration, samp.get_chain().shape, flush=True)
