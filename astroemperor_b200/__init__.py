"""astroemperor_b200 — B200-native hot path of EMPEROR (ReddTea/astroemperor).

Scope: batched multi-Keplerian RV model + Gaussian
log-likelihood + priors (+ Hipparcos-Gaia block) over every walker x
temperature of a reddemcee-style parallel-tempering ensemble, and the stretch
move / accept / swap step, as hand-written sm_100a CUDA behind a C-ABI
(include/emperor_b200.h).  There is no CPU fallback: every compute entry point
raises if `libemperor_b200.so` is missing or no GPU is present.
"""
from .modelspec import (ModelSpec, BlockSpec, ParamSpec, AdditionalPrior, CompiledModel,
                        UnsupportedModelError, spec_from_reddmodel)

__version__ = "0.1.0"

__all__ = ["ModelSpec", "BlockSpec", "ParamSpec", "AdditionalPrior", "CompiledModel",
           "UnsupportedModelError", "spec_from_reddmodel", "__version__"]
