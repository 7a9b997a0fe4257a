"""Shared parity-case table (TEST INFRASTRUCTURE ONLY): the same seeded cases are used by
``oracle/make_golden.py`` (reference → fixtures), ``tests/test_oracle.py`` (restatement vs
fixtures) and the ``-m gpu`` parity tests (CUDA path vs fixtures / oracle).

Case list follows SURVEY.md §8c: 0-dim and (B,) t; B=1 unpadded; padded batches with noise
in the pad region and pad in {1,2,>=3}; CFG branch with non-zero fake_*; n_mel in {80,128};
T in {1,2,3,127,128,129,300,1000}; plus fixed-grid ODE trajectories.
"""
from __future__ import annotations

WEIGHT_SEED = 0
CFG_SEED = 7

# name -> dict(kind, n_mel, T, lengths, seed, ...)
ESTIMATOR_CASES = {
    "call_scalar_t_padded":   dict(n_mel=80, T=300, lengths=[300, 251], seed=11),
    "call_batch_t":           dict(n_mel=80, T=200, lengths=[200, 137, 64], seed=12, t_per_sample=True),
    "call_T1":                dict(n_mel=80, T=1, lengths=[1], seed=13),
    "call_T2":                dict(n_mel=80, T=2, lengths=[2, 1], seed=14),
    "call_T3":                dict(n_mel=80, T=3, lengths=[3], seed=15),
    "call_T127":              dict(n_mel=80, T=127, lengths=[127], seed=16),
    "call_T128":              dict(n_mel=80, T=128, lengths=[128], seed=17),
    "call_T129":              dict(n_mel=80, T=129, lengths=[129, 128], seed=18),
    "call_pad0":              dict(n_mel=80, T=50, lengths=[50], seed=19),
    "call_pad1":              dict(n_mel=80, T=51, lengths=[50], seed=19),
    "call_pad2":              dict(n_mel=80, T=52, lengths=[50], seed=19),
    "call_pad3":              dict(n_mel=80, T=53, lengths=[50], seed=19),
    "call_pad10":             dict(n_mel=80, T=60, lengths=[50], seed=19),
    "call_mel128":            dict(n_mel=128, T=160, lengths=[160, 99], seed=20),
    "call_T1000":             dict(n_mel=80, T=1000, lengths=[1000], seed=21),
    "call_zero_len":          dict(n_mel=80, T=40, lengths=[40, 0], seed=22),
    "call_t0":                dict(n_mel=80, T=96, lengths=[96, 70], seed=23, t_value=0.0),
}

# full ODE solves through CFMDecoder.forward semantics
SOLVE_CASES = {
    "solve_euler10_cfg":      dict(n_mel=80, T=200, lengths=[200, 141], seed=31, steps=10, method="euler", cfg=3.0),
    "solve_euler10_nocfg":    dict(n_mel=80, T=150, lengths=[150, 150, 77], seed=32, steps=10, method="euler", cfg=None),
    "solve_midpoint4_cfg":    dict(n_mel=80, T=96, lengths=[96, 33], seed=33, steps=4, method="midpoint", cfg=2.0),
    "solve_rk4_3_nocfg":      dict(n_mel=80, T=64, lengths=[64], seed=34, steps=3, method="rk4", cfg=None),
    "solve_euler5_cfg_mel128": dict(n_mel=128, T=100, lengths=[100, 81], seed=35, steps=5, method="euler", cfg=3.0),
    "solve_dopri5fixed2_cfg":  dict(n_mel=80, T=80, lengths=[80, 51], seed=36, steps=2, method="dopri5_fixed", cfg=3.0),
}
