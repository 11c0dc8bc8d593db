"""ORACLE (test infrastructure, NOT product code) -- CPU restatement of the
Ceres trust-region Levenberg-Marquardt minimiser that the reference calls at
glomap/estimators/global_positioning.cc:83 and bundle_adjustment.cc:99
(`ceres::Solve` with SPARSE_SCHUR, glomap/estimators/optimization_base.h:18-23).

PARITY UNPINNED: Ceres is an un-vendored, un-pinned system dependency of the
reference (cmake/FindDependencies.cmake:4, vcpkg.json) and cannot be built in
this container; the semantics below are restated from the public Ceres 2.x
sources/documentation (trust_region_minimizer.cc, levenberg_marquardt_strategy.cc,
corrector.cc, loss_function.cc, line_search.cc) and are flagged
UPSTREAM-UNVERIFIED.  The reference holds no golden vectors for this path
(SURVEY.md 8(c)); the oracle is validated against synthetic ground truth using
the reference's own end-to-end thresholds (global_mapper_test.cc:84-86).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may import
this module.

Semantics restated:
  * cost = 1/2 sum rho(||r_i||^2); corrector with rho'' <= 0: residual and
    Jacobian rows scaled by sqrt(rho') (corrector.cc).
  * HuberLoss(a): rho(s) = s (s <= a^2) else 2 a sqrt(s) - a^2; ScaledLoss(a).
  * Jacobi scaling computed once at iteration 0: 1 / (1 + sqrt(colnorm^2)).
  * LM: D^2 = clamp(diag(J_s^T J_s), 1e-6, 1e32) / radius, initial radius 1e4,
    (J^T J + D^2) y = -J^T r solved exactly (the reference: Schur + CHOLMOD).
  * step quality = (cost - cost_new) / model_cost_change,
    model_cost_change = -(J d)^T (r + J d / 2); accepted if > 1e-3;
    accepted: radius /= max(1/3, 1 - (2 q - 1)^3); rejected: radius /= f, f *= 2.
  * termination checks in Ceres' order: parameter tolerance (1e-8), function
    tolerance (|dcost| <= ftol * cost), then accept/reject; gradient tolerance
    (1e-10, max-norm) and min radius (1e-32) after each iteration.
  * box bounds: Plus() projects; projected Armijo line search before the
    candidate evaluation (trust_region_minimizer.cc DoLineSearch).
"""
from __future__ import annotations

import dataclasses

import numpy as np
import scipy.sparse as sp
import scipy.sparse.linalg as spla


@dataclasses.dataclass
class LMOptions:
    max_num_iterations: int = 100          # optimization_base.h:20
    function_tolerance: float = 1e-5       # optimization_base.h:22
    gradient_tolerance: float = 1e-10      # Ceres default
    parameter_tolerance: float = 1e-8      # Ceres default
    initial_trust_region_radius: float = 1e4
    max_trust_region_radius: float = 1e16
    min_trust_region_radius: float = 1e-32
    min_relative_decrease: float = 1e-3
    min_lm_diagonal: float = 1e-6
    max_lm_diagonal: float = 1e32
    jacobi_scaling: bool = True
    max_num_consecutive_invalid_steps: int = 5
    max_num_line_search_step_size_iterations: int = 20
    verbose: bool = False


@dataclasses.dataclass
class LMSummary:
    iterations: int = 0            # LM iterations (successful + unsuccessful)
    num_successful_steps: int = 0
    initial_cost: float = 0.0
    final_cost: float = 0.0
    termination: str = ""
    usable: bool = True
    costs: list = dataclasses.field(default_factory=list)


def huber_rho(s: np.ndarray, a: float, scale: float = 1.0):
    """Ceres HuberLoss(a) optionally wrapped in ScaledLoss(scale): returns
    (rho, rho') evaluated at squared norms s."""
    b = a * a
    big = s > b
    r = np.sqrt(np.where(big, s, 1.0))
    rho0 = np.where(big, 2 * a * r - b, s)
    rho1 = np.where(big, np.maximum(np.finfo(float).tiny, a / r), 1.0)
    return scale * rho0, scale * rho1


def _interp_step(f0, g0, prev, cur, lo, hi):
    """Minimiser in [lo, hi] of the polynomial through (0, f0, g0), cur (and
    prev when given) -- value-only Armijo interpolation of line_search.cc."""
    x1, f1 = cur
    cands = [lo, hi]
    if prev is None:
        # quadratic: f0 + g0 x + c x^2
        c = (f1 - f0 - g0 * x1) / (x1 * x1)
        if c > 0:
            cands.append(-g0 / (2 * c))
        poly = lambda x: f0 + g0 * x + c * x * x
    else:
        x2, f2 = prev
        A = np.array([[x1 ** 3, x1 ** 2], [x2 ** 3, x2 ** 2]])
        rhs = np.array([f1 - f0 - g0 * x1, f2 - f0 - g0 * x2])
        try:
            a3, a2 = np.linalg.solve(A, rhs)
        except np.linalg.LinAlgError:
            a3, a2 = 0.0, (f1 - f0 - g0 * x1) / (x1 * x1)
        poly = lambda x: f0 + g0 * x + a2 * x * x + a3 * x ** 3
        roots = np.roots([3 * a3, 2 * a2, g0]) if abs(a3) > 0 else (np.array([-g0 / (2 * a2)]) if a2 != 0 else np.array([]))
        cands += [float(r.real) for r in np.atleast_1d(roots) if abs(r.imag) < 1e-14]
    cands = [min(max(c, lo), hi) for c in cands]
    return min(cands, key=poly)


def solve_lm(x0, evaluate, plus, opts: LMOptions, project=None, x_norm_fn=None):
    """Generic Ceres-style LM.

    evaluate(x, want_jacobian) -> (cost, r, J): robust cost 1/2 sum rho, the
      loss-corrected residual vector and (when asked) the corrected sparse
      Jacobian in the tangent space of the variable parameters.
    plus(x, delta) -> x (+) delta, INCLUDING the bound projection (Ceres
      ParameterBlock::Plus projects onto box constraints).
    project: callable used for the projected gradient norm when bounds exist.
    """
    summ = LMSummary()
    x = x0
    cost, r, J = evaluate(x, True)
    summ.initial_cost = cost
    summ.costs.append(cost)
    n = J.shape[1]
    if opts.jacobi_scaling:
        scale = 1.0 / (1.0 + np.sqrt(np.asarray(J.multiply(J).sum(axis=0)).ravel()))
    else:
        scale = np.ones(n)
    g = J.T @ r
    is_constrained = project is not None

    def grad_max_norm(xx, gg):
        if not is_constrained:
            return float(np.abs(gg).max()) if len(gg) else 0.0
        return float(np.abs(project(xx, -gg)).max())

    radius = opts.initial_trust_region_radius
    decrease_factor = 2.0
    invalid = 0
    x_norm = x_norm_fn(x) if x_norm_fn else 0.0
    if grad_max_norm(x, g) <= opts.gradient_tolerance:
        summ.termination = "gradient tolerance (initial)"
        summ.final_cost = cost
        return x, summ

    it = 0
    while True:
        if it >= opts.max_num_iterations:
            summ.termination = "max iterations"
            break
        if radius < opts.min_trust_region_radius:
            summ.termination = "min trust region radius"
            break
        it += 1
        Js = J @ sp.diags(scale)
        diag = np.asarray(Js.multiply(Js).sum(axis=0)).ravel()
        diag = np.minimum(np.maximum(diag, opts.min_lm_diagonal), opts.max_lm_diagonal)
        H = (Js.T @ Js + sp.diags(diag / radius)).tocsc()
        rhs = -(Js.T @ r)
        try:
            y = spla.splu(H).solve(rhs)
        except RuntimeError:
            y = np.full(n, np.nan)
        Jy = Js @ y
        model_cost_change = -float(Jy @ (r + 0.5 * Jy))
        if not np.all(np.isfinite(y)) or model_cost_change <= 0:
            invalid += 1
            if invalid >= opts.max_num_consecutive_invalid_steps:
                summ.termination = "too many invalid steps"
                summ.usable = False
                break
            radius /= decrease_factor
            decrease_factor *= 2
            continue
        invalid = 0
        delta = y * scale
        if is_constrained and opts.max_num_line_search_step_size_iterations > 0:
            # Projected Armijo line search (trust_region_minimizer.cc DoLineSearch).
            g0 = float(g @ delta)
            alpha, prev, cur = 1.0, None, None
            ok = False
            for _ in range(opts.max_num_line_search_step_size_iterations + 1):
                fa, _, _ = evaluate(plus(x, alpha * delta), False)
                if np.isfinite(fa) and fa <= cost + 1e-4 * g0 * alpha:
                    ok = True
                    break
                prev, cur = cur, (alpha, fa)
                if not np.isfinite(fa):
                    alpha *= 1e-3
                    prev = cur = None
                else:
                    alpha = _interp_step(cost, g0, prev, cur, 1e-3 * alpha, 0.6 * alpha)
                if alpha * np.abs(delta).max() < 1e-9:
                    break
            if ok:
                delta = delta * alpha
        x_cand = plus(x, delta)
        cand_cost, _, _ = evaluate(x_cand, False)
        # Parameter tolerance (step norm in the ambient space).
        if x_norm_fn is not None:
            step_norm = x_norm_fn(x_cand, x)
            if step_norm <= opts.parameter_tolerance * (x_norm + opts.parameter_tolerance):
                summ.termination = "parameter tolerance"
                break
        cost_change = cost - cand_cost
        if abs(cost_change) <= opts.function_tolerance * cost:
            summ.termination = "function tolerance"
            break
        rel = cost_change / model_cost_change
        if opts.verbose:
            print(f"  it {it}: cost {cost:.6e} -> {cand_cost:.6e} rel {rel:.3f} radius {radius:.3e}")
        if rel > opts.min_relative_decrease:
            x = x_cand
            cost, r, J = evaluate(x, True)
            g = J.T @ r
            x_norm = x_norm_fn(x) if x_norm_fn else 0.0
            summ.num_successful_steps += 1
            summ.costs.append(cost)
            radius = min(opts.max_trust_region_radius, radius / max(1.0 / 3.0, 1.0 - (2.0 * rel - 1.0) ** 3))
            decrease_factor = 2.0
            if grad_max_norm(x, g) <= opts.gradient_tolerance:
                summ.termination = "gradient tolerance"
                break
        else:
            radius /= decrease_factor
            decrease_factor *= 2
    summ.iterations = it
    summ.final_cost = cost
    return x, summ
