"""Second witness of the trust-region loop -- TEST INFRASTRUCTURE ONLY.

oracle/ba_oracle.c restates ceres::Solve (TrustRegionMinimizer + DoglegStrategy + DENSE_SCHUR) and the CUDA solver is
checked against it; nothing else checked the oracle's own state machine.  This file is a separately written NumPy version
of the same loop, written from the Ceres 1.14 documentation of the algorithm (solver options in
vins_estimator/src/estimator/estimator.cpp:3305-3315: DENSE_SCHUR, DOGLEG, max_num_iterations 8) and not from ba_oracle.c:

  * it works on the Jacobian, never on the normal equations: the Gauss-Newton step is the least-squares solution of the
    stacked system [J; sqrt(mu) D] y = [r; 0] (numpy.linalg.lstsq, an SVD), the Cauchy step length uses |J v|^2 and the
    model decrease uses J s directly;
  * no Schur complement: the landmarks stay in the system;
  * ambient-space norms (x_norm, step_norm, gradient_max_norm) are taken over the Python-side parameter arrays.

Only three things are shared with the C oracle, all of them factor-level and covered by their own finite-difference and
SciPy tests (tests/test_ba_oracle.py): gfo_ba_linearize (residuals + Jacobian at a point), gfo_ba_plus (the manifold
update) and gfo_ba_cost.  A mistake in the oracle's dogleg/trust-region bookkeeping therefore shows up as a trace mismatch.
"""
import numpy as np

from oracle import ba_oracle as O

_ARRAYS = ("para_pose", "para_speed_bias", "para_ex_pose", "para_feature", "para_td", "para_ex_wheel", "para_ix_wheel",
           "para_td_wheel", "para_plane_R", "para_plane_Z")


def _ambient(pb):
    return [np.array(getattr(pb, k), float).reshape(-1).copy() for k in _ARRAYS]


def _free_blocks(pb, n_cols):
    """Which ambient entries belong to non-constant parameter blocks: perturb a clone along every tangent direction and
    see what moves (block granularity: pose 7, speed-bias 9, extrinsics 7, plane rotation 4, scalars 1)."""
    q = pb.clone()
    a0 = _ambient(q)
    O.plus(q, np.full(n_cols, 1e-3))
    a1 = _ambient(q)
    sizes = dict(para_pose=7, para_speed_bias=9, para_ex_pose=7, para_feature=1, para_td=1, para_ex_wheel=7, para_ix_wheel=1,
                 para_td_wheel=1, para_plane_R=4, para_plane_Z=1)
    masks = []
    for k, u, v in zip(_ARRAYS, a0, a1):
        moved = (u != v).reshape(-1, sizes[k]).any(axis=1)
        masks.append(np.repeat(moved, sizes[k]))
    return masks


def _vec(pb, masks):
    return np.concatenate([a[m] for a, m in zip(_ambient(pb), masks)])


def solve(pb, function_tolerance=1e-6, gradient_tolerance=1e-10, parameter_tolerance=1e-8):
    """Runs the loop on pb (updated in place).  Returns dict(cost=[...], radius=[...], iterations, termination, successful)
    with the same per-iteration meaning as gf_ba_summary: entry k holds the accepted cost and the radius after iteration k."""
    r, J = O.linearize(pb)
    n = J.shape[1]
    masks = _free_blocks(pb, n)
    cost = O.cost(pb)                  # 1/2 sum rho(|r|^2): not 1/2 |r_corrected|^2 when a loss is active
    # Jacobi scaling, fixed at the first linearisation (Solver::Options::jacobi_scaling)
    col_scale = 1.0 / (1.0 + np.sqrt((J * J).sum(axis=0)))

    def gradient_max_norm(J_, r_):
        q = pb.clone()
        x0 = _vec(q, masks)
        O.plus(q, -(J_.T @ r_))
        return float(np.abs(x0 - _vec(q, masks)).max()) if x0.size else 0.0

    out = dict(cost=[cost], radius=[1e4], iterations=0, successful=0, termination="NO_CONVERGENCE")
    if n == 0 or gradient_max_norm(J, r) <= gradient_tolerance:
        out["termination"] = "CONVERGENCE_GRADIENT"
        return out
    Js = J * col_scale
    x_norm = float(np.linalg.norm(_vec(pb, masks)))
    radius, mu = 1e4, 1e-8
    fresh = True                       # the dogleg subspace (gradient, Cauchy length, Gauss-Newton point) must be recomputed
    bad_steps = 0
    it = 0
    while it < pb.max_num_iterations and radius >= 1e-32:
        it += 1
        ok = True
        if fresh:
            fresh = False
            D = np.sqrt(np.clip((Js * Js).sum(axis=0), 1e-6, 1e32))
            g = (Js.T @ r) / D                                     # gradient in the D-scaled variables
            Jv = Js @ (g / D)
            cauchy_len = float(g @ g) / float(Jv @ Jv)
            ok = False
            while mu < 1.0:
                A = np.vstack([Js, np.diag(D * np.sqrt(mu))])
                y, *_ = np.linalg.lstsq(A, np.concatenate([r, np.zeros(n)]), rcond=None)
                if np.all(np.isfinite(y)):
                    ok = True
                    break
                mu *= 10.0
            if ok:
                gn = -y * D                                        # Gauss-Newton point in the D-scaled variables
        if ok:
            gn_norm, g_norm = float(np.linalg.norm(gn)), float(np.linalg.norm(g))
            if gn_norm <= radius:
                s, s_norm = gn.copy(), gn_norm
            elif g_norm * cauchy_len >= radius:
                s, s_norm = -(radius / g_norm) * g, radius
            else:
                # walk from the Cauchy point a = -cauchy_len g towards gn until |a + beta (gn - a)| = radius
                a = -cauchy_len * g
                d = gn - a
                aa, ad, dd = float(a @ a), float(a @ d), float(d @ d)
                disc = np.sqrt(ad * ad + dd * (radius * radius - aa))
                beta = (disc - ad) / dd if ad <= 0 else (radius * radius - aa) / (disc + ad)
                s = a + beta * d
                s_norm = float(np.linalg.norm(s))
            s = s / D
            Jstep = Js @ s
            model_decrease = -float(Jstep @ (r + 0.5 * Jstep))
            ok = model_decrease > 0.0
        if not ok:
            bad_steps += 1
            out["cost"].append(cost); out["radius"].append(radius)
            if bad_steps >= 5:
                out["termination"] = "FAILURE"
                break
            mu *= 10.0
            fresh = True
            continue
        bad_steps = 0
        cand = pb.clone()
        x_before = _vec(pb, masks)
        O.plus(cand, s * col_scale)
        cand_cost = O.cost(cand)
        step_norm = float(np.linalg.norm(x_before - _vec(cand, masks)))
        if step_norm <= parameter_tolerance * (x_norm + parameter_tolerance):
            out["cost"].append(cost); out["radius"].append(radius); out["termination"] = "CONVERGENCE_PARAMETER"
            break
        if abs(cost - cand_cost) <= function_tolerance * cost:
            out["cost"].append(cost); out["radius"].append(radius); out["termination"] = "CONVERGENCE_FUNCTION"
            break
        rho = (cost - cand_cost) / model_decrease
        if rho > 1e-3:
            for k in _ARRAYS:
                getattr(pb, k)[...] = getattr(cand, k)
            r, J = O.linearize(pb)
            Js = J * col_scale
            cost = cand_cost
            x_norm = float(np.linalg.norm(_vec(pb, masks)))
            out["successful"] += 1
            if rho < 0.25:
                radius *= 0.5
            if rho > 0.75:
                radius = max(radius, 3.0 * s_norm)
            mu = max(1e-8, 2.0 * mu / 10.0)
            fresh = True
            out["cost"].append(cost); out["radius"].append(radius)
            if gradient_max_norm(J, r) <= gradient_tolerance:
                out["termination"] = "CONVERGENCE_GRADIENT"
                break
        else:
            radius *= 0.5
            out["cost"].append(cost); out["radius"].append(radius)
    out["iterations"] = it
    return out
