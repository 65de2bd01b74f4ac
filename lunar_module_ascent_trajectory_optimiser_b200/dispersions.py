"""Synthetic parameter dispersions of BASELINE.json's configs.

Nominal = the reference's literals (LO:38, 50-52, 61-66, 70-71).  Draws come from
``torch.rand(B, 6, generator=manual_seed(seed))`` on the CPU with a fixed column order, so
the GPU run, the oracle and the CPU baseline all see identical inputs.
"""
from __future__ import annotations

import torch

from .api import AscentParams, G_ISP

NOMINAL_ISP = 15346.0 / (5.053 * G_ISP)   # Isp that reproduces the reference's M_dot literal


def nominal_params() -> AscentParams:
    return AscentParams()


def dispersed_params(B: int, seed: int = 11, columns=(0, 1, 2, 3, 4, 5)) -> AscentParams:
    """cfg 3: ``columns=(0,1,2,3)`` (thrust, Isp, initial mass, angular-accel limit);
    cfg 4/5: all six (adds target perilune/apolune).  Problem 0 of every batch is the
    nominal reference case so that each run carries its own known answer."""
    U = torch.rand(B, 6, dtype=torch.float64, generator=torch.Generator().manual_seed(seed))
    for c in range(6):
        if c not in columns:
            U[:, c] = 0.5
    U[0, :] = 0.5
    Ft = 15346.0 * (1 + 0.02 * (2 * U[:, 0] - 1))
    Isp = NOMINAL_ISP * (1 + 0.01 * (2 * U[:, 1] - 1))
    M_dot = Ft / (Isp * G_ISP)
    M0 = 4821.0 * (1 + 0.02 * (2 * U[:, 2] - 1))
    addm = 5e-4 * torch.pow(torch.tensor(2.0, dtype=torch.float64), 2 * U[:, 3] - 1)
    rp = 17703.0 * (1 + 0.10 * (2 * U[:, 4] - 1))
    ra = 88615.0 * (1 + 0.10 * (2 * U[:, 5] - 1))
    return AscentParams(Ft=Ft, M_dot=M_dot, M0=M0, angle_doubledot_max=addm, r_periapsis=rp, r_apoapsis=ra)
