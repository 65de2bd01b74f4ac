"""CPU oracle, part 3: the reference PDF's post-solve orbit check ("SECOND PART OF THE CODE",
PDF p.28-29, source lines 185-237).  TEST INFRASTRUCTURE ONLY.

The PDF coasts the final ascent state around the Moon with an explicit Euler scheme to see whether
the orbit closes:
    Gs = 6.67e-11 (sic), m_2 = 7.346e22                          src 186-188
    start = (f_x, f_y) = (-x*S, y*S + R0), velocity (-xdot*S, ydot*S)   src 189-199
    acceleration = Gs*m_2 * (-pos/|pos|) / |pos|^2                 src 202-212
    per step (delta_t = 0.001, t = 6600 s):                        src 225-236
        a  = acc(position)
        position += velocity * delta_t        (old velocity)
        velocity += a * delta_t
It only plots the path; the batched version also reports the extreme radii, which is what one reads
off the plot (achieved perilune / apolune).
"""
from __future__ import annotations

import numpy as np

GS_PDF = 6.67e-11      # PDF src 186 (two significant digits fewer than LO:50)
M2_PDF = 7.346e22      # PDF src 188


def coast(state: np.ndarray, nsteps: int, dt: float = 1e-3, gm: float = GS_PDF * M2_PDF):
    """state: [4, B] = x, y, vx, vy (SI, Moon-centred).  Returns (r_min, r_max, final state [4, B])."""
    x, y, vx, vy = (np.array(s, dtype=np.float64) for s in state)
    r2 = x * x + y * y
    r2min, r2max = r2.copy(), r2.copy()
    for _ in range(nsteps):
        r2 = x * x + y * y
        inv = 1.0 / np.sqrt(r2)
        k = gm * inv * inv * inv
        ax, ay = -k * x, -k * y
        x = x + vx * dt
        y = y + vy * dt
        vx = vx + ax * dt
        vy = vy + ay * dt
        r2n = x * x + y * y
        r2min = np.minimum(r2min, r2n)
        r2max = np.maximum(r2max, r2n)
    return np.sqrt(r2min), np.sqrt(r2max), np.stack([x, y, vx, vy])
