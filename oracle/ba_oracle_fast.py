"""ORACLE (test infrastructure, NOT product code) -- LM driver over the C/OpenMP
hot loops of oracle/ba_oracle_c.c: the same Ceres semantics as
oracle/ceres_lm.py (which it is tested against), but with the Schur complement
formed explicitly and the reduced camera system factored densely (LAPACK
Cholesky through scipy) -- the structure of the reference's SPARSE_SCHUR path
(bundle_adjustment.cc:94-96; at these sizes the reduced system is dense).
Used for parity checks beyond numpy+splu sizes and as bench.py's CPU baseline
("port": the reference needs Ceres + COLMAP which cannot be built here).
PARITY UNPINNED, see oracle/ceres_lm.py.
"""
from __future__ import annotations

import ctypes as ct
import os
import subprocess
import time

import numpy as np
import scipy.linalg as sla

from .ba_oracle import BAOptions, quat_plus
from .ceres_lm import LMSummary

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def lib():
    global _LIB
    if _LIB is None:
        path = os.path.join(_HERE, "libba_oracle_c.so")
        if not os.path.exists(path):
            subprocess.check_call(["make", "-C", _HERE, "all"])
        L = ct.CDLL(path)
        L.ba_c_linearize.restype = ct.c_double
        L.ba_c_cost.restype = ct.c_double
        L.ba_c_num_threads.restype = ct.c_int
        _LIB = L
    return _LIB


def _p(a):
    return a.ctypes.data_as(ct.c_void_p)


def num_threads() -> int:
    return int(lib().ba_c_num_threads())


def solve_ba_fast(quat, trans, points, pt_obs_begin, obs_cam, obs_xy, cam_intr, intr_model, intr_params,
                  opts: BAOptions | None = None, cam_const_mask=None, fixed_num_iterations: int = 0, verbose=False):
    """Same contract as oracle.ba_oracle.solve_ba (intrinsics constant only).
    Returns (state dict, LMSummary); summary.times holds per-phase seconds."""
    o = opts or BAOptions()
    assert not o.optimize_intrinsics, "fast oracle: intrinsics are held constant"
    assert o.optimize_points, "fast oracle: points are variables"
    L = lib()
    C, P = len(quat), len(points)
    ptb = np.ascontiguousarray(pt_obs_begin, np.int64)
    oc = np.ascontiguousarray(obs_cam, np.int32)
    xy = np.ascontiguousarray(obs_xy, np.float64)
    ci = np.ascontiguousarray(cam_intr, np.int32)
    im = np.ascontiguousarray(intr_model, np.int32)
    intr = np.ascontiguousarray(intr_params, np.float64)
    N = len(oc)
    mask = np.zeros(C, np.uint8) if cam_const_mask is None else np.ascontiguousarray(cam_const_mask, np.uint8).copy()
    if not o.optimize_rotations:
        mask |= 1
    if not o.optimize_translation:
        mask |= 2
    q = np.ascontiguousarray(quat, np.float64).copy()
    q /= np.linalg.norm(q, axis=1, keepdims=True)
    t = np.ascontiguousarray(trans, np.float64).copy()
    X = np.ascontiguousarray(points, np.float64).copy()
    U = np.empty((C, 6, 6)); gc = np.empty((C, 6)); V = np.empty((P, 3, 3)); gp = np.empty((P, 3))
    W = np.empty((N, 18)); Vinv = np.empty((P, 3, 3)); S = np.empty((6 * C, 6 * C)); b = np.empty(6 * C)
    dp = np.empty((P, 3))
    mv, a = int(o.min_num_view_per_track), float(o.thres_loss_function)
    times = dict(linearize=0.0, schur=0.0, solve=0.0, backsub=0.0, cost=0.0)

    def linearize():
        t0 = time.perf_counter()
        c = L.ba_c_linearize(C, P, _p(ptb), _p(oc), _p(xy), _p(ci), _p(im), _p(intr), _p(q), _p(t), _p(X), _p(mask),
                             mv, ct.c_double(a), _p(U), _p(gc), _p(V), _p(gp), _p(W))
        times["linearize"] += time.perf_counter() - t0
        return c

    def cost_of(qq, tt, XX):
        t0 = time.perf_counter()
        c = L.ba_c_cost(C, P, _p(ptb), _p(oc), _p(xy), _p(ci), _p(im), _p(intr), _p(qq), _p(tt), _p(XX), mv,
                        ct.c_double(a))
        times["cost"] += time.perf_counter() - t0
        return c

    summ = LMSummary()
    cost = linearize()
    summ.initial_cost = cost
    summ.costs.append(cost)
    dU = np.einsum("cii->ci", U).copy()
    var_c = dU > 0                      # masked / unobserved dofs are not variables
    dV = np.einsum("pii->pi", V).copy()
    var_p = dV[:, 0] > 0
    js_c = np.where(var_c, 1.0 / (1.0 + np.sqrt(dU)), 1.0)
    js_p = 1.0 / (1.0 + np.sqrt(dV))
    radius, decrease, invalid, it = 1e4, 2.0, 0, 0
    fixed = fixed_num_iterations > 0
    max_it = fixed_num_iterations if fixed else o.max_num_iterations

    def gmax():
        return max(np.abs(gc[var_c]).max(initial=0.0), np.abs(gp[var_p]).max(initial=0.0))

    if not fixed and gmax() <= 1e-10:
        summ.termination = "gradient tolerance (initial)"
        max_it = 0
    while it < max_it:
        if radius < 1e-32:
            summ.termination = "min trust region radius"
            break
        it += 1
        dU = np.einsum("cii->ci", U)
        dV = np.einsum("pii->pi", V)
        Dc = np.where(var_c, np.clip(dU * js_c ** 2, 1e-6, 1e32) / (radius * js_c ** 2), 0.0)
        Dp = np.clip(dV * js_p ** 2, 1e-6, 1e32) / (radius * js_p ** 2)
        t0 = time.perf_counter()
        L.ba_c_schur(C, P, _p(ptb), _p(oc), mv, _p(U), _p(gc), _p(V), _p(gp), _p(W), _p(np.ascontiguousarray(Dc)),
                     _p(np.ascontiguousarray(Dp)), _p(S), _p(b), _p(Vinv))
        times["schur"] += time.perf_counter() - t0
        fix = ~var_c.ravel()
        if fix.any():
            idx = np.nonzero(fix)[0]
            S[idx, :] = 0; S[:, idx] = 0; S[idx, idx] = 1; b[idx] = 0
        t0 = time.perf_counter()
        try:
            cf = sla.cho_factor(S, lower=True, overwrite_a=False, check_finite=False)
            dc = sla.cho_solve(cf, b, check_finite=False)
        except sla.LinAlgError:
            dc = np.full(6 * C, np.nan)
        times["solve"] += time.perf_counter() - t0
        t0 = time.perf_counter()
        dc = np.ascontiguousarray(dc)
        L.ba_c_backsub(P, _p(ptb), _p(oc), mv, _p(Vinv), _p(gp), _p(W), _p(dc), _p(dp))
        times["backsub"] += time.perf_counter() - t0
        dcm = dc.reshape(C, 6)
        # exact solve: model_cost_change = -1/2 g.d + 1/2 d^T D d  (PCG residual = 0)
        mcc = 0.5 * (-(gc * dcm).sum() - (gp * dp).sum() + (Dc * dcm * dcm).sum() + (Dp * dp * dp).sum())
        if not np.isfinite(mcc) or mcc <= 0:
            invalid += 1
            if invalid >= 5:
                summ.termination = "too many invalid steps"; summ.usable = False
                break
            radius /= decrease; decrease *= 2
            continue
        invalid = 0
        qn = quat_plus(q, dcm[:, :3]); tn = t + dcm[:, 3:]; Xn = X + dp
        qn = np.ascontiguousarray(qn); tn = np.ascontiguousarray(tn); Xn = np.ascontiguousarray(Xn)
        cand = cost_of(qn, tn, Xn)
        if not fixed:
            rv = var_c[:, :3].any(1); tv = var_c[:, 3:].any(1)
            step = np.sqrt(((qn - q)[rv] ** 2).sum() + ((tn - t)[tv] ** 2).sum() + ((Xn - X)[var_p] ** 2).sum())
            xn = np.sqrt((q[rv] ** 2).sum() + (t[tv] ** 2).sum() + (X[var_p] ** 2).sum())
            if step <= 1e-8 * (xn + 1e-8):
                summ.termination = "parameter tolerance"
                break
            if abs(cost - cand) <= o.function_tolerance * cost:
                summ.termination = "function tolerance"
                break
        rel = (cost - cand) / mcc
        if verbose:
            print(f"  it {it}: cost {cost:.6e} -> {cand:.6e} rel {rel:.3f} radius {radius:.3e}")
        if rel > 1e-3:
            q, t, X = qn, tn, Xn
            cost = linearize()
            summ.num_successful_steps += 1
            summ.costs.append(cost)
            radius = min(1e16, radius / max(1.0 / 3.0, 1.0 - (2.0 * rel - 1.0) ** 3))
            decrease = 2.0
            if not fixed and gmax() <= 1e-10:
                summ.termination = "gradient tolerance"
                break
        else:
            radius /= decrease; decrease *= 2
    else:
        if not summ.termination:
            summ.termination = "max iterations"
    summ.iterations = it
    summ.final_cost = cost
    summ.times = times
    return dict(quat=q, trans=t, points=X, intr=intr), summ
