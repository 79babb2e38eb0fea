"""The compiled device registry that replaces Julia closures (north_star b).

A flux / source / condition function handed to FVMProblem must be one of these specs; anything
else (an arbitrary Python callable standing in for a Julia closure) is rejected with
UnsupportedClosureError before any kernel launch.  (x,y)-only coefficient functions may be plain
callables: they are tabulated on the host at setup (TabulatedDiffusion / TabulatedSource)."""
from dataclasses import dataclass
from typing import Callable, Sequence

# ids mirror include/fvmcuda.h
FLUX_DIFF_CONST, FLUX_DIFF_TABLE, FLUX_DIFF_POWER, FLUX_ADVDIFF, FLUX_KELLER_SEGEL = range(5)
SRC_ZERO, SRC_LINEAR, SRC_LOGISTIC, SRC_TABLE, SRC_GRAY_SCOTT, SRC_BRUSSELATOR, SRC_KELLER_SEGEL = range(7)
COND_CONST, COND_AFFINE_U, COND_EXP_SAT, COND_LINEAR_XY, COND_EXP_XYT = range(5)


# ---- diffusion / flux functions: q(x,y,t,alpha,beta,gamma,p) or D(x,y,t,u,p) ------------------
@dataclass(frozen=True)
class ConstantDiffusion:
    """D(x,y,t,u,p) = D   (q = -D grad u)"""
    D: float = 1.0


@dataclass(frozen=True)
class TabulatedDiffusion:
    """D(x,y,t,u,p) = fn(x, y): tabulated at the cv-edge midpoints and boundary quarter points."""
    fn: Callable


@dataclass(frozen=True)
class PowerDiffusion:
    """D(x,y,t,u,p) = D0 * u^(m-1)  (porous medium; use_abs: D0*|u|^(m-1))"""
    D0: float
    m: float
    use_abs: bool = False


@dataclass(frozen=True)
class AdvectionDiffusionFlux:
    """q = (nu_x u - D u_x, nu_y u - D u_y)"""
    D: float
    nu_x: float
    nu_y: float = 0.0


@dataclass(frozen=True)
class KellerSegelFlux:
    """species 0: chi(u) grad v - grad u, chi = c u / (1 + u^2); species 1: -D grad v
    (one spec for the whole FVMSystem; give it to every member problem)"""
    c: float
    D: float


# ---- sources S(x,y,t,u,p) ------------------------------------------------------------------------
@dataclass(frozen=True)
class ZeroSource:
    pass


@dataclass(frozen=True)
class LinearSource:
    """S = lam * u + mu"""
    lam: float
    mu: float = 0.0


@dataclass(frozen=True)
class LogisticSource:
    """S = lam * u * (1 - u)"""
    lam: float


@dataclass(frozen=True)
class TabulatedSource:
    """S = fn(x, y), tabulated per node"""
    fn: Callable


@dataclass(frozen=True)
class GrayScottSource:
    """(b (1-u) - u v^2, -d v + u v^2); one spec for the whole system"""
    b: float
    d: float


@dataclass(frozen=True)
class BrusselatorSource:
    """(u^2 v - 2u, -u^2 v + u)"""


@dataclass(frozen=True)
class KellerSegelSource:
    """(u (1-u), u - a v)"""
    a: float


# ---- boundary / internal condition functions a(x,y,t,u,p) ------------------------------------------
@dataclass(frozen=True)
class Const:
    c: float = 0.0


@dataclass(frozen=True)
class AffineU:
    """c0 + c1 * u"""
    c0: float
    c1: float


@dataclass(frozen=True)
class ExpSaturation:
    """c0 * (1 - exp(-t / tau))"""
    c0: float
    tau: float


@dataclass(frozen=True)
class LinearXY:
    """c0 + cx * x + cy * y"""
    c0: float
    cx: float
    cy: float


@dataclass(frozen=True)
class ExpXYT:
    """c0 * exp(cx * x + cy * y + ct * t)  (the Brusselator tutorial's boundary data)"""
    c0: float
    cx: float = 0.0
    cy: float = 0.0
    ct: float = 0.0


def cond_spec(fn):
    from ._lib import ERR_UNSUPPORTED, UnsupportedClosureError
    if isinstance(fn, Const):
        return COND_CONST, [fn.c]
    if isinstance(fn, AffineU):
        return COND_AFFINE_U, [fn.c0, fn.c1]
    if isinstance(fn, ExpSaturation):
        return COND_EXP_SAT, [fn.c0, fn.tau]
    if isinstance(fn, LinearXY):
        return COND_LINEAR_XY, [fn.c0, fn.cx, fn.cy]
    if isinstance(fn, ExpXYT):
        return COND_EXP_XYT, [fn.c0, fn.cx, fn.cy, fn.ct]
    if isinstance(fn, (int, float)):
        return COND_CONST, [float(fn)]
    raise UnsupportedClosureError(
        ERR_UNSUPPORTED,
        "condition function %r is not in the compiled device registry (Const, AffineU, ExpSaturation, LinearXY, ExpXYT); "
        "arbitrary closures cannot run on the GPU" % (fn,))
