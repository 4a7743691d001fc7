"""TEST INFRASTRUCTURE ONLY.  Oracle for the global Gauss-Newton (VSLAM/backend/src/gn_kernels.cu): the reference's kernels
and host loop restated in PyTorch — residuals, Huber weights, Jacobians (gn_kernels.cu:924-1095 rays, :1346-1500 calib),
``apply_Sim3_adj_inv`` (:267-290), the 14x14 block layout and the index bookkeeping of ``SparseBlock`` (:71-113), a double
Cholesky solve (the reference: Eigen SimplicialLLT in double, :136-160), ``expSim3`` / ``retrSim3`` (:316-412) and the
termination rule of the host loop (:1141-1229).  Everything except the solve runs in float32 like the kernels.
PARITY UNPINNED by the reference: its GN extension needs Eigen, which is not vendored, so it cannot be built here; this is a
restatement from the source."""
import torch

EPS = 1e-6


def quat_comp(qi, qj):
    x = qi[..., 3] * qj[..., 0] + qi[..., 0] * qj[..., 3] + qi[..., 1] * qj[..., 2] - qi[..., 2] * qj[..., 1]
    y = qi[..., 3] * qj[..., 1] - qi[..., 0] * qj[..., 2] + qi[..., 1] * qj[..., 3] + qi[..., 2] * qj[..., 0]
    z = qi[..., 3] * qj[..., 2] + qi[..., 0] * qj[..., 1] - qi[..., 1] * qj[..., 0] + qi[..., 2] * qj[..., 3]
    w = qi[..., 3] * qj[..., 3] - qi[..., 0] * qj[..., 0] - qi[..., 1] * qj[..., 1] - qi[..., 2] * qj[..., 2]
    return torch.stack([x, y, z, w], -1)


def act_so3(q, X):
    qv = q[..., :3].expand(X.shape)
    uv = 2.0 * torch.cross(qv, X, dim=-1)
    return X + q[..., 3:4] * uv + torch.cross(qv, uv, dim=-1)


def rel_sim3(ti, qi, si, tj, qj, sj):
    si_inv = 1.0 / si
    qi_inv = qi * torch.tensor([-1.0, -1.0, -1.0, 1.0], dtype=qi.dtype)
    return act_so3(qi_inv, tj - ti) * si_inv, quat_comp(qi_inv, qj), si_inv * sj


def adj_inv(t, q, s, X):
    """Row vectors X [..., 7] times the inverse adjoint (tau, omega, s)."""
    s_inv = 1.0 / s
    Ra = act_so3(q, X[..., 0:3])
    Y0 = s_inv * Ra
    Y1 = act_so3(q, X[..., 3:6]) + s_inv * torch.cross(t.expand(Ra.shape), Ra, dim=-1)
    Y2 = X[..., 6:7] + s_inv * (t * Ra).sum(-1, keepdim=True)
    return torch.cat([Y0, Y1, Y2], -1)


def huber(r):
    a = r.abs()
    return torch.where(a < 1.345, torch.ones_like(a), 1.345 / a)


def _rows_rays(P, Xi, valid, q, sigma_ray, sigma_dist):
    n1i = Xi.norm(dim=-1, keepdim=True)
    n2j = (P * P).sum(-1, keepdim=True)
    n1j = n2j.sqrt()
    r = P / n1j
    err = torch.cat([r - Xi / n1i, n1j - n1i], -1)                       # [n,4]
    sw_a = torch.where(valid, q.sqrt() / sigma_ray, torch.zeros_like(q))
    sw_b = torch.where(valid, q.sqrt() / sigma_dist, torch.zeros_like(q))
    sw = torch.stack([sw_a, sw_a, sw_a, sw_b], -1)
    w = huber(sw * err) * sw * sw
    n3 = 1.0 / (n1j * n2j)
    eye = torch.eye(3, dtype=P.dtype)
    dr = eye / n1j[..., None] - P[..., :, None] * P[..., None, :] * n3[..., None]   # [n,3,3]
    z = torch.zeros_like(r[..., 0])
    rot = torch.stack([torch.stack([z, r[..., 2], -r[..., 1]], -1), torch.stack([-r[..., 2], z, r[..., 0]], -1),
                       torch.stack([r[..., 1], -r[..., 0], z], -1)], -2)            # rows: (0,z,-y), (-z,0,x), (y,-x,0)
    J3 = torch.cat([dr, rot, torch.zeros_like(r)[..., None]], -1)                   # [n,3,7]
    J4 = torch.cat([r, torch.zeros_like(r), n1j], -1)[..., None, :]                 # [n,1,7]
    return err, w, torch.cat([J3, J4], -2)


def _rows_calib(P, Xi, ind, valid, q, K, height, width, pixel_border, z_eps, sigma_pixel, sigma_depth):
    fx, fy, cx, cy = K[0, 0], K[1, 1], K[0, 2], K[1, 2]
    u_t, v_t = (ind % width).to(P.dtype), (ind // width).to(P.dtype)
    vz = (P[..., 2] > z_eps) & (Xi[..., 2] > z_eps)
    zj_inv = torch.where(vz, 1.0 / P[..., 2], torch.zeros_like(P[..., 2]))
    zj_log = torch.where(vz, P[..., 2].clamp_min(1e-30).log(), torch.zeros_like(zj_inv))
    zi_log = torch.where(vz, Xi[..., 2].clamp_min(1e-30).log(), torch.zeros_like(zj_inv))
    xz, yz = P[..., 0] * zj_inv, P[..., 1] * zj_inv
    u, v = fx * xz + cx, fy * yz + cy
    vu = (u > pixel_border) & (u < width - 1 - pixel_border)
    vv = (v > pixel_border) & (v < height - 1 - pixel_border)
    valid = valid & vu & vv & vz
    err = torch.stack([u - u_t, v - v_t, zj_log - zi_log], -1)
    sw_a = torch.where(valid, q.sqrt() / sigma_pixel, torch.zeros_like(q))
    sw_b = torch.where(valid, q.sqrt() / sigma_depth, torch.zeros_like(q))
    sw = torch.stack([sw_a, sw_a, sw_b], -1)
    w = huber(sw * err) * sw * sw
    z = torch.zeros_like(xz)
    o = torch.ones_like(xz)
    J = torch.stack([torch.stack([fx * zj_inv, z, -fx * xz * zj_inv, -fx * xz * yz, fx * (1 + xz * xz), -fx * yz, z], -1),
                     torch.stack([z, fy * zj_inv, -fy * yz * zj_inv, -fy * (1 + yz * yz), fy * xz * yz, fy * xz, z], -1),
                     torch.stack([z, z, zj_inv, yz, -xz, z, o], -1)], -2)
    return err, w, J


def exp_so3(phi):
    th2 = (phi * phi).sum()
    if th2 < EPS:
        th4 = th2 * th2
        imag = 0.5 - th2 / 48.0 + th4 / 3840.0
        real = 1.0 - th2 / 8.0 + th4 / 384.0
    else:
        th = th2.sqrt()
        imag = torch.sin(0.5 * th) / th
        real = torch.cos(0.5 * th)
    return torch.cat([imag * phi, real.reshape(1)])


def exp_sim3(xi):
    tau, phi, sigma = xi[0:3], xi[3:6], xi[6]
    scale = torch.exp(sigma)
    q = exp_so3(phi)
    th2 = (phi * phi).sum()
    th = th2.sqrt()
    one = torch.tensor(1.0, dtype=xi.dtype)
    if sigma.abs() < EPS:
        C = one
        if th.abs() < EPS:
            A, B = 0.5 * one, one / 6.0
        else:
            A, B = (1 - torch.cos(th)) / th2, (th - torch.sin(th)) / (th2 * th)
    else:
        C = (scale - 1) / sigma
        if th.abs() < EPS:
            s2 = sigma * sigma
            A = ((sigma - 1) * scale + 1) / s2
            B = (scale * 0.5 * s2 + scale - 1 - sigma * scale) / (s2 * sigma)
        else:
            a, b, c = scale * torch.sin(th), scale * torch.cos(th), th2 + sigma * sigma
            A = (a * sigma + (1 - b) * th) / (th * c)
            B = (C - ((b - 1) * sigma + a * th) / c) / th2
    c1 = torch.linalg.cross(phi, tau)
    c2 = torch.linalg.cross(phi, c1)
    return C * tau + A * c1 + B * c2, q, scale


def gauss_newton(mode, Twc, Xs, Cs, ii, jj, idx_ii2jj, valid_match, Q, sigma_a, sigma_b, C_thresh, Q_thresh, max_iter,
                 delta_thresh, K=None, height=0, width=0, pixel_border=0, z_eps=0.0):
    """mode 'rays' | 'calib'.  Twc [K,8] float32 is updated in place; returns (dx_last [K-1,7], iterations run)."""
    num_fix = 1
    E, n = idx_ii2jj.shape[0], Xs.shape[1]
    unique = torch.unique(torch.cat([ii, jj]), sorted=True)
    ie, je = torch.searchsorted(unique, ii), torch.searchsorted(unique, jj)
    Kp = Xs.shape[0]
    D = 7 * (Kp - num_fix)
    Cs2, Q2, vm2 = Cs.reshape(Kp, n), Q.reshape(E, n), valid_match.reshape(E, n)
    dx, its = torch.zeros(Kp - num_fix, 7), 0
    for _ in range(max_iter):
        H = torch.zeros(D, D, dtype=torch.float64)
        b = torch.zeros(D, dtype=torch.float64)
        for e in range(E):
            ix, jx = int(ie[e]), int(je[e])
            ti, qi, si = Twc[ix, 0:3], Twc[ix, 3:7], Twc[ix, 7:8]
            tj, qj, sj = Twc[jx, 0:3], Twc[jx, 3:7], Twc[jx, 7:8]
            tij, qij, sij = rel_sim3(ti, qi, si, tj, qj, sj)
            vm = vm2[e]
            ind = torch.where(vm, idx_ii2jj[e], torch.zeros_like(idx_ii2jj[e]))
            Xi, Xj = Xs[ix][ind], Xs[jx]
            P = act_so3(qij, Xj) * sij + tij
            q = Q2[e]
            valid = vm & (q > Q_thresh) & (Cs2[ix][ind] > C_thresh) & (Cs2[jx] > C_thresh)
            if mode == "rays":
                err, w, Jrow = _rows_rays(P, Xi, valid, q, sigma_a, sigma_b)
            else:
                err, w, Jrow = _rows_calib(P, Xi, ind, valid, q, K, height, width, pixel_border, z_eps, sigma_a, sigma_b)
            Jj = adj_inv(ti, qi, si, Jrow)                                  # [n,R,7]
            Jx = torch.cat([-Jj, Jj], -1)                                    # Ji = -Jj
            Hb = torch.einsum("nr,nra,nrb->ab", w, Jx, Jx).double()          # fp32 products, like the kernel
            g = torch.einsum("nr,nr,nra->a", w, err, Jx).double()
            io, jo = ix - num_fix, jx - num_fix
            for (a, ra) in ((io, 0), (jo, 7)):
                if a < 0:
                    continue
                b[a * 7:a * 7 + 7] += g[ra:ra + 7]
                for (c, rc) in ((io, 0), (jo, 7)):
                    if c >= 0:
                        H[a * 7:a * 7 + 7, c * 7:c * 7 + 7] += Hb[ra:ra + 7, rc:rc + 7]
        try:
            L = torch.linalg.cholesky(H)
            x = torch.cholesky_solve(b[:, None], L)[:, 0]
            dx = (-x).float().reshape(-1, 7)
        except Exception:  # noqa: BLE001  (SimplicialLLT failure path returns zeros)
            dx = torch.zeros(Kp - num_fix, 7)
        for k in range(num_fix, Kp):
            dt, dq, ds = exp_sim3(dx[k - num_fix])
            t, q, s = Twc[k, 0:3].clone(), Twc[k, 3:7].clone(), Twc[k, 7].clone()
            Twc[k, 3:7] = quat_comp(dq, q)
            Twc[k, 0:3] = act_so3(dq, t) * ds + dt
            Twc[k, 7] = ds * s
        its += 1
        if float(dx.norm()) < delta_thresh:
            break
    return dx, its
