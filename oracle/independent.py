"""independent.py -- a SECOND, independent restatement of the hot path, in float64 numpy.

TEST INFRASTRUCTURE ONLY -- PARITY UNPINNED (see flame_oracle.h).  Written from the equations of
SURVEY.md Appendix A (NLTGV2-L1 Chambolle-Pock iteration on a graph) and Appendix B (per-feature
epipolar inverse-depth update), NOT from oracle/flame_oracle.c: different language, different
precision (float64), different evaluation strategy (whole-array scatter/gather instead of
per-element loops, dense candidate matrices instead of a sliding window).  The C oracle and the CUDA
kernels share one author and one expression order, so a shared misreading of the specification
would be invisible between them; it is not invisible against this file
(tests/test_oracle_independent.py).  Where the specification leaves a choice open, the choice
documented in DESIGN.md section 5 is restated here from that text.

What can NOT be checked by any file in this repository: whether robustrobotics/flame itself makes
the same choices (its source is absent, DESIGN.md "Parity unpinned").
"""
import numpy as np


# ------------------------------------------------------------------------------------------ Appendix A
def nltgv2_solve(pos, edges, alpha, beta, z, wt, iters, data_factor=0.15, step_x=0.001, step_q=125.0,
                 theta=0.25, x_min=0.0, x_max=10.0, state=None, history=None):
    """Chambolle-Pock iterations of

        min_{x,w}  sum_v lambda c_v |x_v - z_v|  +  sum_{e=(i,j)} || D_e(x, w) ||_1
        D_e = [ a_e (x_i - x_j - <w_i, p_i - p_j>),  b_e (w1_i - w1_j),  b_e (w2_i - w2_j) ]

    dual:   q <- proj_{|.|<=1}( q + sigma D(xbar, wbar) )
    primal: (x, w) <- (x, w) - tau D^T q ;  x <- prox_{tau lambda c |. - z|}(x) ;  x <- clip(x)
    extra:  (xbar, wbar) <- (x, w) + theta ((x, w) - (x, w)_old)

    state: dict(x, w [V,2], q [E,3], xb, wb) to continue from; None = cold start x = z, w = q = 0.
    history: optional list receiving (smoothness cost, data cost, saturated dual fraction) per iteration.
    Returns the state dict (float64)."""
    pos = np.asarray(pos, np.float64)
    i, j = np.asarray(edges, np.int64).T
    a, b = np.asarray(alpha, np.float64), np.asarray(beta, np.float64)
    z, c = np.asarray(z, np.float64), np.asarray(wt, np.float64)
    V, E = len(z), len(a)
    d = pos[i] - pos[j]
    if state is None:
        state = dict(x=z.copy(), w=np.zeros((V, 2)), q=np.zeros((E, 3)), xb=z.copy(), wb=np.zeros((V, 2)))
    x, w, q, xb, wb = (np.array(state[k], np.float64) for k in ("x", "w", "q", "xb", "wb"))

    def D(xx, ww):
        return np.stack([a * (xx[i] - xx[j] - np.einsum("ek,ek->e", ww[i], d)),
                         b * (ww[i, 0] - ww[j, 0]), b * (ww[i, 1] - ww[j, 1])], axis=1)

    for _ in range(iters):
        t = q + step_q * D(xb, wb)
        q = t / np.maximum(1.0, np.abs(t))
        # D^T q: row e of D touches x_i (+a), x_j (-a), w_i (-a d, +b), w_j (-b)
        gx = np.bincount(i, a * q[:, 0], V) - np.bincount(j, a * q[:, 0], V)
        gw = np.zeros((V, 2))
        for k in range(2):
            gw[:, k] = (np.bincount(i, -a * d[:, k] * q[:, 0] + b * q[:, 1 + k], V) - np.bincount(j, b * q[:, 1 + k], V))
        x_old, w_old = x, w
        xp = x - step_x * gx
        w = w - step_x * gw
        th = step_x * data_factor * c
        r = xp - z
        x = z + np.sign(r) * np.maximum(np.abs(r) - th, 0.0)   # soft threshold toward the data
        x = np.clip(x, x_min, x_max)
        xb = x + theta * (x - x_old)
        wb = w + theta * (w - w_old)
        if history is not None:
            history.append((float(np.abs(D(x, w)).sum()), float((data_factor * c * np.abs(x - z)).sum()),
                            float((np.abs(q) >= 0.999).mean())))
    return dict(x=x, w=w, q=q, xb=xb, wb=wb)


def costs(pos, edges, alpha, beta, z, wt, x, w, data_factor=0.15):
    """(smoothness, data) costs of Appendix A at (x, w)."""
    pos = np.asarray(pos, np.float64)
    i, j = np.asarray(edges, np.int64).T
    d = pos[i] - pos[j]
    x, w = np.asarray(x, np.float64), np.asarray(w, np.float64)
    k1 = alpha * (x[i] - x[j] - np.einsum("ek,ek->e", w[i], d))
    k2 = beta * (w[i, 0] - w[j, 0])
    k3 = beta * (w[i, 1] - w[j, 1])
    return float(np.abs(k1).sum() + np.abs(k2).sum() + np.abs(k3).sum()), float((data_factor * np.asarray(wt) * np.abs(x - z)).sum())


# ------------------------------------------------------------------------------------------ Appendix B
def quat_to_R(q):
    x, y, z, w = (float(v) for v in q)
    n = x * x + y * y + z * z + w * w
    s = 2.0 / n
    return np.array([[1 - s * (y * y + z * z), s * (x * y - w * z), s * (x * z + w * y)],
                     [s * (x * y + w * z), 1 - s * (x * x + z * z), s * (y * z - w * x)],
                     [s * (x * z - w * y), s * (y * z + w * x), 1 - s * (x * x + y * y)]])


def relative_geometry(K, pose_ref, pose_cmp):
    """A = K R Kinv, b = K t for T_cmp<-ref = T_cmp^-1 T_ref (poses are camera-in-world)."""
    K = np.asarray(K, np.float64).reshape(3, 3)
    Rr, Rc = quat_to_R(pose_ref[:4]), quat_to_R(pose_cmp[:4])
    tr, tc = np.asarray(pose_ref[4:7], np.float64), np.asarray(pose_cmp[4:7], np.float64)
    R = Rc.T @ Rr
    t = Rc.T @ (tr - tc)
    return K @ R @ np.linalg.inv(K), K @ t


def bilinear(img, x, y):
    """fp64 bilinear sample of a uint8 image at (x, y) arrays (caller keeps them inside)."""
    img = np.asarray(img, np.float64)
    x0, y0 = np.floor(x).astype(np.int64), np.floor(y).astype(np.int64)
    fx, fy = x - x0, y - y0
    i00, i10 = img[y0, x0], img[y0, x0 + 1]
    i01, i11 = img[y0 + 1, x0], img[y0 + 1, x0 + 1]
    top = i00 + fx * (i10 - i00)
    bot = i01 + fx * (i11 - i01)
    return top + fy * (bot - top)


def project_feature(A, b, u, xi):
    """Pixel in cmp of the ref pixel u at inverse depth xi: pi(A (u,1) + xi b); also the depth ratio."""
    p = A @ np.array([u[0], u[1], 1.0]) + xi * b
    return p[:2] / p[2], p[2]


def epipolar_measurement(img_ref, img_cmp, K, pose_ref, pose_cmp, u, mu, var, win=5, search_sigma=2.0,
                         idepth_min=0.0, idepth_max=10.0, max_search_px=64):
    """The photometric search of Appendix B for ONE feature, restated with a dense evaluation: every
    candidate position (1 px apart along the epipolar segment of mu +- k sigma) gets the mean squared
    difference between the reference patch (win samples, 1 px apart along the epipolar direction in
    ref) and the comparison samples.  Returns dict(n, costs, best, u_cmp, idepth) or None when there
    is nothing to search.  Thresholds / status logic are left to the C oracle: this restates the
    geometry, the sampling and the triangulation, which is where a misreading would hide."""
    A, b = relative_geometry(K, pose_ref, pose_cmp)
    sig = np.sqrt(var)
    lo, hi = max(mu - search_sigma * sig, idepth_min), min(mu + search_sigma * sig, idepth_max)
    p_lo, _ = project_feature(A, b, u, lo)
    p_hi, _ = project_feature(A, b, u, hi)
    p_mu, _ = project_feature(A, b, u, mu)
    seg = p_hi - p_lo
    length = float(np.hypot(*seg))
    if length < 1e-9:
        return None
    l = seg / length
    # epipolar direction in ref: image of the cmp camera centre seen from u
    Ainv_b = np.linalg.solve(A, b)   # A^-1 b = K R^T t direction; epipole in ref = -A^-1 b (homogeneous)
    e = -Ainv_b
    if abs(e[2]) > 1e-12:
        dir_ref = np.array([u[0], u[1]]) - e[:2] / e[2]
        if e[2] < 0:
            dir_ref = -dir_ref
    else:
        dir_ref = -e[:2]
    # orient the ref direction so that it maps onto +l in cmp: move u by one pixel along dir_ref
    dir_ref = dir_ref / np.hypot(*dir_ref)
    q1, _ = project_feature(A, b, (u[0] + dir_ref[0], u[1] + dir_ref[1]), mu)
    if np.dot(q1 - p_mu, l) < 0:
        dir_ref = -dir_ref
    half = win // 2
    offs = np.arange(-half, half + 1, dtype=np.float64)
    ref_patch = bilinear(img_ref, u[0] + offs * dir_ref[0], u[1] + offs * dir_ref[1])
    n = int(min(max_search_px, np.floor(length) + 1))
    H, W = np.asarray(img_cmp).shape
    start = p_mu - 0.5 * (n - 1) * l if n * 1.0 < length + 1 else p_lo
    cand = start[None, :] + np.arange(n)[:, None] * l[None, :]
    samp = cand[:, None, :] + offs[None, :, None] * l[None, None, :]
    inside = (samp[..., 0] >= 0) & (samp[..., 1] >= 0) & (samp[..., 0] < W - 1) & (samp[..., 1] < H - 1)
    ok = inside.all(axis=1)
    costs = np.full(n, np.inf)
    if ok.any():
        vals = bilinear(img_cmp, samp[ok][..., 0], samp[ok][..., 1])
        costs[ok] = ((vals - ref_patch[None, :]) ** 2).mean(axis=1)
    if not np.isfinite(costs).any():
        return None
    best = int(np.argmin(costs))
    sub = 0.0
    if 0 < best < n - 1 and np.isfinite(costs[best - 1]) and np.isfinite(costs[best + 1]):
        den = costs[best - 1] - 2 * costs[best] + costs[best + 1]
        if den > 1e-12:
            sub = 0.5 * (costs[best - 1] - costs[best + 1]) / den
    u_cmp = cand[best] + sub * l
    # triangulate along the dominant axis of the epipolar line: p(xi) = (P0 + xi b), x = p.x / p.z
    P0 = A @ np.array([u[0], u[1], 1.0])
    ax = 0 if abs(l[0]) >= abs(l[1]) else 1
    idepth = (u_cmp[ax] * P0[2] - P0[ax]) / (b[ax] - u_cmp[ax] * b[2])
    return dict(n=n, costs=costs, best=best, u_cmp=u_cmp, idepth=float(idepth), dir_ref=dir_ref, l=l)


def gaussian_fuse(mu, var, mu_m, var_m):
    """Product of two Gaussians (Appendix B step 5)."""
    return (var_m * mu + var * mu_m) / (var + var_m), var * var_m / (var + var_m)
