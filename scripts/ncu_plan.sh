#!/bin/bash
# One full ncu capture of the swap-plan kernel at T=256, W=2048 (development aid).
python - <<'PY' > /dev/null 2>&1 &
PY
cat > /tmp/plan_one.py <<'PY'
import os, sys
sys.path.insert(0, os.getcwd()); sys.path.insert(0, os.path.join(os.getcwd(), "tests"))
import numpy as np, torch
from conftest import load_golden
from astroemperor_b200.engine import LikelihoodEngine
from astroemperor_b200.sampler import PTSampler
g, spec = load_golden("c1_51peg_k0")
eng = LikelihoodEngine(spec, g["t"], g["y"], g["yerr"], g["flag"])
T, W = 256, 2048
samp = PTSampler(W, eng.ndim, eng, ntemps=T, seed=3, store=None, betas=np.geomspace(1, 1e-3, T), graph=False)
samp.D_ = spec.prior_widths()
samp._init_state(samp.initial_positions(spec))
samp._alloc_hist(64)
for _ in range(6):
    samp.sweep_begin(samp.draw_staged(0))
torch.cuda.synchronize()
PY
ncu --set full --import-source on --clock-control none -k regex:pt_swap_plan -s 4 -c 1 -o gpurun_out/r02_plan -f python /tmp/plan_one.py > gpurun_out/r02_plan_ncu.log 2>&1
ncu -i gpurun_out/r02_plan.ncu-rep --page raw --csv > gpurun_out/r02_plan_raw.csv 2>/dev/null
ncu -i gpurun_out/r02_plan.ncu-rep --page source --csv > gpurun_out/r02_plan_source.csv 2>/dev/null
tail -3 gpurun_out/r02_plan_ncu.log
