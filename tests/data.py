"""Seeded synthetic inputs for the parity tests and the bench."""
import json
import os

import numpy as np

import bssm_b200

HERE = os.path.dirname(os.path.abspath(__file__))


def poisson_series():
    return np.array(json.load(open(os.path.join(HERE, "golden", "poisson_series.json")))["y"], dtype=np.float64)


def exchange():
    return np.array(json.load(open(os.path.join(HERE, "golden", "exchange.json")))["y"], dtype=np.float64)


def svm_series(n, rho=0.98, sd_ar=0.2, sigma=1.0, gen_seed=2024):
    """alpha_1 ~ N(0, sd^2/(1-rho^2)), alpha_{t+1} = rho alpha_t + sd eta, y = sigma e^{alpha/2} eps"""
    rng = np.random.default_rng(gen_seed)
    a = np.empty(n)
    a[0] = rng.normal() * sd_ar / np.sqrt(1 - rho ** 2)
    for t in range(1, n):
        a[t] = rho * a[t - 1] + sd_ar * rng.normal()
    return sigma * np.exp(a / 2) * rng.normal(size=n)


def svm_model(n, rho=0.98, sd_ar=0.2, sigma=1.0, gen_seed=2024, missing=()):
    y = svm_series(n, rho, sd_ar, sigma, gen_seed)
    for t in missing:
        y[t] = np.nan
    return bssm_b200.svm(y, rho, sd_ar, sigma=sigma)


def poisson_bsm_model(n=100, missing=()):
    """the documented example of R/bootstrap_filter.R:78-82: bsm_ng on poisson_series"""
    y = poisson_series()[:n].copy()
    for t in missing:
        y[t] = np.nan
    return bssm_b200.bsm_ng(y, sd_level=0.1, sd_slope=0.01, P1=np.diag([1.0, 2.0]), distribution="poisson")
