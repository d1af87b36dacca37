"""Host-side optimisers of the reference training loops.

These STAY on the host (np <= 285): in the Julia front end they are Flux's own
`update!(opt, p, grad)`; this numpy restatement exists so the Python mirror of the
scripts' epoch loop is runnable end to end.

  ADAM / ADAMW(eta, beta, decay) = Optimiser(ADAM, WeightDecay)  case1/case1.jl:18, robertson/rober_crnn.jl:19
  ExpDecay(eta, decay, step, clip) chained before ADAMW           case2/case2.jl:31-32
  NADAM(eta, beta)                                                case3/case3.jl:20
  gradient clipping by 2-norm                                     robertson/rober_crnn.jl:220-223
"""
from __future__ import annotations

import numpy as np


class ADAM:
    def __init__(self, eta=1e-3, beta=(0.9, 0.999), eps=1e-8):
        self.eta, self.beta, self.eps = eta, beta, eps
        self.m = self.v = None
        self.bp = [beta[0], beta[1]]   # running beta powers, as Flux keeps them

    def apply(self, p, g):
        if self.m is None:
            self.m = np.zeros_like(g); self.v = np.zeros_like(g)
        b1, b2 = self.beta
        self.m = b1 * self.m + (1 - b1) * g
        self.v = b2 * self.v + (1 - b2) * g * g
        d = self.m / (1 - self.bp[0]) / (np.sqrt(self.v / (1 - self.bp[1])) + self.eps) * self.eta
        self.bp = [self.bp[0] * b1, self.bp[1] * b2]
        return d


class NADAM(ADAM):
    def apply(self, p, g):
        if self.m is None:
            self.m = np.zeros_like(g); self.v = np.zeros_like(g)
        b1, b2 = self.beta
        self.m = b1 * self.m + (1 - b1) * g
        self.v = b2 * self.v + (1 - b2) * g * g
        b1p, b2p = self.bp
        d = (b1 * self.m / (1 - b1 * b1p) + (1 - b1) * g / (1 - b1p)) / (np.sqrt(self.v * b2 / (1 - b2p)) + self.eps) * self.eta
        self.bp = [b1p * b1, b2p * b2]
        return d


class WeightDecay:
    def __init__(self, wd=0.0):
        self.wd = wd

    def apply(self, p, g):
        return g + self.wd * p


class ExpDecay:
    """eta * decay^(floor(count/step)), floored at clip; multiplies the incoming gradient."""

    def __init__(self, eta=1e-3, decay=0.1, step=1000, clip=1e-4):
        self.eta, self.decay, self.step, self.clip = eta, decay, step, clip
        self.count = 0

    def apply(self, p, g):
        self.count += 1
        if self.count % self.step == 0 and self.eta > self.clip:
            self.eta = max(self.eta * self.decay, self.clip)
        return g * self.eta


class Optimiser:
    """Flux.Optimiser: a chain; update!(opt, p, g) does p .-= apply(...)."""

    def __init__(self, *chain):
        self.chain = list(chain)

    def apply(self, p, g):
        """Flux.Optimise.apply!(o::Optimiser, x, Δ): the chain applied in order - lets Optimisers nest, e.g. the literal
        `Flux.Optimiser(ExpDecay(...), ADAMW(...))` of case2/case2.jl:31-32 (ADAMW itself is an Optimiser)."""
        d = np.asarray(g, dtype=np.float64)
        for o in self.chain:
            d = o.apply(p, d)
        return d

    def update(self, p, g):
        p -= self.apply(p, g)
        return p


def ADAMW(eta=1e-3, beta=(0.9, 0.999), decay=0.0):
    return Optimiser(ADAM(eta, beta), WeightDecay(decay))


def clip_by_norm(g, gmax):
    """rober_crnn.jl:220-223."""
    n = float(np.linalg.norm(g))
    return (g / n * gmax, n) if n > gmax else (g, n)
