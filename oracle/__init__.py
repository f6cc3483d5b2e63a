"""ORACLE — test infrastructure, not product code.

CPU restatement of the reference's hot path: the kepler.py
solver in C (`kepler_oracle.c`), the generated `my_model` / `my_likelihood` /
`my_prior` in NumPy (`rv_oracle.py`), the Hipparcos-Gaia block
(`am_oracle.py`) and one reddemcee-style parallel-tempering step with injected
random draws (`pt_oracle.py`).

Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline /
`--impl reference` legs may import this package.  The product package
`astroemperor_b200` never does.
"""
