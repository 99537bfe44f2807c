"""TEST INFRASTRUCTURE -- second, independent restatement of the reference MPPI step in NumPy float32.

Purpose: cross-check oracle/mppi_oracle.c (two restatements written separately from the same Warp
kernels must agree), and serve as BASELINE.md's "B1" vectorised CPU baseline.  Vectorised over the K
samples, sequential over the T steps, float32 throughout (every intermediate is forced to np.float32 so
NumPy never promotes).  Transcendentals are NumPy's float32 sin/cos/exp (libm-grade), so agreement with
the C oracle's MATH_LIBM mode is to rounding, not bitwise.

File:line cites are relative to /root/reference/thesis_master/warp_implementation/.
Parity status: unpinned by the reference's own tests (none exist for this path); see oracle/README.md.
Only tests/ and bench.py's CPU-baseline leg may import this module.
"""
from __future__ import annotations

import numpy as np

f32 = np.float32


def _F(x):
    return np.asarray(x, dtype=np.float32)


DEFAULTS = dict(
    dt=0.045, u1_min=-1.0, u1_max=1.0, u2_min=-1.0, u2_max=1.0, v_min=0.0, v_max=2.0, w_min=-1.0, w_max=1.0,
    lam=0.3, r_wheels=1.2, filt_k=3.5, filt_a=0.96, opt_k=3.0, opt_a=0.92, wheel_offset=0.2,
    cw_path=100.5, cw_slope=50.5, cw_speed=0.5, cw_obs=25.0, lethal_thresh=0.99, lethal_penalty=100000.0,
    near_goal_cut=2.0, speed_eps=0.0001, pf_eps=1e-6, pf_near_gain=10.0, slope_eps=1e-6, slope_gain=5.0,
    # optional critics (weight 0 = off): the reference's dormant ones, then the roll / pitch / effort extensions
    cw_orient=0.0, cw_slope_path=0.0, cw_goal_angle=0.0, goal_angle_radius=0.5, cw_roll=0.0, cw_pitch=0.0,
    cw_effort=0.0,
)


class P:
    """Parameter bag with every field as np.float32 (ints stay ints)."""

    def __init__(self, K, T, proj=3, input_model=0, **kw):
        d = dict(DEFAULTS)
        d.update(kw)
        self.input_model = int(input_model)      # 0 wheel inputs + filter, 1 velocity space (sampling_warp.py:10-48)
        d.setdefault("horizon", d["dt"] * d["v_max"] * T)         # MPPI_isaac.py:440
        d.setdefault("target_speed", d["v_max"])                  # MPPI_isaac.py:619
        self.K, self.T, self.proj = int(K), int(T), int(proj)
        for k, v in d.items():
            setattr(self, k, f32(v))


def clamp(x, lo, hi):
    return np.minimum(np.maximum(x, lo), hi)          # wp.clamp


def sample_inputs(p, nom1, nom2, s1, s2, eps1, eps2):
    """sampling_warp.py:54-92: u[k,t] = clamp(nom[t+1] + sigma*eps) (last step reuses nom[T-1])."""
    T = p.T
    src = np.minimum(np.arange(T) + 1, T - 1)
    lo1, hi1, lo2, hi2 = ((p.v_min, p.v_max, p.w_min, p.w_max) if getattr(p, "input_model", 0) == 1
                          else (p.u1_min, p.u1_max, p.u2_min, p.u2_max))
    u1 = clamp(_F(nom1)[src][None, :] + f32(s1) * _F(eps1), lo1, hi1)
    u2 = clamp(_F(nom2)[src][None, :] + f32(s2) * _F(eps2), lo2, hi2)
    return _F(u1), _F(u2)


def inputs_to_velocities(p, u1, u2, wl, wr, k, a):
    """sampling_warp.py:96-138 (u arrays [N, T])."""
    u1, u2 = _F(u1), _F(u2)
    N, T = u1.shape
    k, a = f32(k), f32(a)
    one_m_a = f32(1.0) - a
    l = np.full(N, wl, dtype=np.float32)
    r = np.full(N, wr, dtype=np.float32)
    v = np.zeros((N, T), np.float32)
    w = np.zeros((N, T), np.float32)
    for t in range(T):
        l = l * a + u1[:, t] * k * one_m_a
        r = r * a + u2[:, t] * k * one_m_a
        v[:, t] = clamp((l + r) / f32(2.0), p.v_min, p.v_max)
        w[:, t] = clamp((-l + r) / p.r_wheels, p.w_min, p.w_max)
    return v, w


class Terrain:
    def __init__(self, dem, half_width, costmap):
        self.Z = np.ascontiguousarray(dem, np.float32).reshape(-1)
        self.gs = int(dem.shape[0])
        self.hw = f32(half_width)
        self.res = f32(2.0 * half_width / self.gs)                # MPPI_isaac.py:265
        self.cm = np.ascontiguousarray(costmap, np.float32).reshape(-1)
        self.cms = int(costmap.shape[0])
        self.cres = f32(2.0 * half_width / self.cms)              # MPPI_isaac.py:272


def cell_index(ter, x, y):
    """projection_warp.py:39-40 with x_min = y_min = -half_width (MPPI_isaac.py:584-585)."""
    x_min = -ter.hw
    y_min = -ter.hw
    i = np.trunc((x - x_min) / ter.res).astype(np.int32)
    j = -np.trunc((y + y_min) / ter.res).astype(np.int32)
    return i, j


def corners(ter, x, y):
    """projection_warp.py:8-48 -> q00, q01, q10, q11, (i, j)."""
    i, j = cell_index(ter, x, y)
    ci = np.clip(i, 0, ter.gs - 2).astype(np.int64)
    cj = np.clip(j, 0, ter.gs - 2).astype(np.int64)
    gs = ter.gs
    return (ter.Z[cj * gs + ci], ter.Z[cj * gs + ci + 1], ter.Z[(cj + 1) * gs + ci], ter.Z[(cj + 1) * gs + ci + 1],
            i, j)


def bilinear(x, y, q00, q01, q10, q11, res):
    """projection_warp.py:70-100."""
    xn = x / res
    yn = y / res
    x2 = xn - np.trunc(xn)
    y2 = yn - np.trunc(yn)
    one = f32(1.0)
    return (one - x2) * (one - y2) * q00 + x2 * (one - y2) * q10 + (one - x2) * y2 * q01 + x2 * y2 * q11


def normal_on_grid(q00, q01, q10, q11, res):
    """projection_warp.py:129-151."""
    vx = -res / f32(2.0) * (q01 - q00 - q10 + q11)
    vy = -res / f32(2.0) * (q10 - q00 - q01 + q11)
    vz = np.full_like(vx, res * res)
    norm = np.sqrt(vx * vx + vy * vy + vz * vz)
    return np.stack([vx / norm, vy / norm, vz / norm], axis=-1)


def dot(a, b):
    return a[..., 0] * b[..., 0] + a[..., 1] * b[..., 1] + a[..., 2] * b[..., 2]


def cross(a, b):
    return np.stack([a[..., 1] * b[..., 2] - a[..., 2] * b[..., 1],
                     a[..., 2] * b[..., 0] - a[..., 0] * b[..., 2],
                     a[..., 0] * b[..., 1] - a[..., 1] * b[..., 0]], axis=-1)


def tangent(n, prev):
    """projection_warp.py:168-190."""
    d = dot(prev, n)
    proj = prev - d[..., None] * n
    return proj / np.sqrt(dot(proj, proj))[..., None]


def update_position(x, y, h, v, dt):
    """projection_warp.py:207-223."""
    h = h / np.sqrt(dot(h, h))[..., None]
    disp = h * v[..., None] * dt
    return x + disp[..., 0], y + disp[..., 1]


def update_orientation(h, w, n, dt):
    """projection_warp.py:225-248 (Rodrigues)."""
    h = h / np.sqrt(dot(h, h))[..., None]
    ang = w * dt
    c = np.cos(ang).astype(np.float32)
    s = np.sin(ang).astype(np.float32)
    r = h * c[..., None] + cross(n, h) * s[..., None] + n * dot(n, h)[..., None] * (f32(1.0) - c)[..., None]
    return r / np.sqrt(dot(r, r))[..., None]


def update_orientation_2d(h, w, dt):
    """projection_warp.py:251-275."""
    th = w * dt
    c = np.cos(th).astype(np.float32)
    s = np.sin(th).astype(np.float32)
    nx = c * h[..., 0] - s * h[..., 1]
    ny = s * h[..., 0] + c * h[..., 1]
    norm = np.sqrt(nx * nx + ny * ny)
    pos = norm > 0
    nx = np.where(pos, nx / np.where(pos, norm, f32(1)), nx)
    ny = np.where(pos, ny / np.where(pos, norm, f32(1)), ny)
    return np.stack([nx, ny, np.zeros_like(nx)], axis=-1)


def rollout(p, ter, x0, y0, heading0, v, w):
    """projection_warp.py:284-350 (3-D) / :353-382 (2-D) for N samples at once."""
    N, T = v.shape
    x = np.full(N, x0, np.float32)
    y = np.full(N, y0, np.float32)
    h0 = np.tile(_F(heading0)[None, :], (N, 1))
    traj = np.zeros((N, T, 3), np.float32)
    head = np.zeros((N, T, 3), np.float32)
    lw = np.zeros((N, T, 3), np.float32)
    rw = np.zeros((N, T, 3), np.float32)
    dem_ij = np.zeros((N, T, 2), np.int32)
    lw_ij = np.zeros((N, T, 2), np.int32)
    rw_ij = np.zeros((N, T, 2), np.int32)
    gs = ter.gs
    if p.proj == 3:
        q00, q01, q10, q11, _, _ = corners(ter, x, y)
        n = normal_on_grid(q00, q01, q10, q11, ter.res)
        prev = tangent(n, h0)
        for t in range(T):
            x, y = update_position(x, y, prev, v[:, t], p.dt)
            q00, q01, q10, q11, i, j = corners(ter, x, y)
            hgt = bilinear(x, y, q00, q01, q10, q11, ter.res)
            n = normal_on_grid(q00, q01, q10, q11, ter.res)
            prev = tangent(n, prev)
            cur = update_orientation(prev, w[:, t], n, p.dt)
            head[:, t] = cur
            traj[:, t, 0], traj[:, t, 1], traj[:, t, 2] = x, y, hgt
            dem_ij[:, t, 0], dem_ij[:, t, 1] = i, j
            right = p.wheel_offset * cross(n, cur)
            for sign, arr, ij in ((f32(1), lw, lw_ij), (f32(-1), rw, rw_ij)):
                xw = x + right[:, 0] if sign > 0 else x - right[:, 0]
                yw = y + right[:, 1] if sign > 0 else y - right[:, 1]
                wi, wj = cell_index(ter, xw, yw)
                ij[:, t, 0], ij[:, t, 1] = wi, wj
                ci = np.clip(wi, 0, gs - 1).astype(np.int64)
                cj = np.clip(wj, 0, gs - 1).astype(np.int64)
                arr[:, t, 0], arr[:, t, 1], arr[:, t, 2] = xw, yw, ter.Z[cj * gs + ci]
            prev = cur
    else:
        prev = h0
        for t in range(T):
            x, y = update_position(x, y, prev, v[:, t], p.dt)
            cur = update_orientation_2d(prev, w[:, t], p.dt)
            head[:, t] = cur
            q00, q01, q10, q11, i, j = corners(ter, x, y)
            hgt = bilinear(x, y, q00, q01, q10, q11, ter.res)
            traj[:, t, 0], traj[:, t, 1], traj[:, t, 2] = x, y, hgt
            dem_ij[:, t, 0], dem_ij[:, t, 1] = i, j
            prev = cur
    return dict(traj=traj, heading=head, lw=lw, rw=rw, dem_ij=dem_ij, lw_ij=lw_ij, rw_ij=rw_ij)


def path_follow(p, x, y, gx, gy, traj):
    """critics_warp.py:85-127."""
    x, y, gx, gy = f32(x), f32(y), f32(gx), f32(gy)
    xd, yd = gx - x, gy - y
    dist = np.sqrt(xd * xd + yd * yd)
    if dist > p.horizon:
        igx = x + xd * p.horizon / (dist + p.pf_eps)
        igy = y + yd * p.horizon / (dist + p.pf_eps)
        last = traj[:, -1]
        cost = (last[:, 0] - igx) * (last[:, 0] - igx) + (last[:, 1] - igy) * (last[:, 1] - igy)
        return cost * (f32(1.0) + f32(2.0) * p.horizon / dist)
    cost = np.zeros(traj.shape[0], np.float32)
    for t in range(traj.shape[1] - 1):
        cost = cost + p.pf_near_gain * (np.abs(traj[:, t, 0] - gx) + np.abs(traj[:, t, 1] - gy))
    return cost


def avoid_slope_wheels(p, lw, rw):
    """critics_warp.py:168-218."""
    N, T, _ = lw.shape
    total = np.zeros(N, np.float32)
    one = f32(1.0)
    for i in range(0, T - 3, 2):
        out = []
        for arr in (lw, rw):
            c, pv = arr[:, i + 2], arr[:, i]
            dz = c[:, 2] - pv[:, 2]
            d = np.sqrt((c[:, 0] - pv[:, 0]) * (c[:, 0] - pv[:, 0]) + (c[:, 1] - pv[:, 1]) * (c[:, 1] - pv[:, 1]))
            ratio = np.abs(dz / (d + p.slope_eps))
            out.append((one + p.slope_gain * ratio) * (one + p.slope_gain * ratio))
        total = total + np.where(out[0] > out[1], out[0], out[1])
    return total


def maximise_speed(p, x, y, gx, gy, v):
    """critics_warp.py:269-300."""
    xd, yd = f32(gx) - f32(x), f32(gy) - f32(y)
    dist = np.sqrt(xd * xd + yd * yd)
    if dist < p.near_goal_cut:
        return np.zeros(v.shape[0], np.float32)
    acc = np.zeros(v.shape[0], np.float32)
    for t in range(v.shape[1]):
        acc = acc + (p.target_speed - v[:, t]) / (v[:, t] + p.speed_eps)
    return acc


def avoid_obstacle(p, ter, traj):
    """critics_warp.py:220-267."""
    N, T, _ = traj.shape
    acc = np.zeros(N, np.float32)
    cm_ij = np.zeros((N, T, 2), np.int32)
    for t in range(T):
        ix = np.trunc((traj[:, t, 0] + ter.hw) / ter.cres).astype(np.int32)
        iy = np.trunc((-traj[:, t, 1] + ter.hw) / ter.cres).astype(np.int32)
        cm_ij[:, t, 0], cm_ij[:, t, 1] = ix, iy
        c = ter.cm[np.clip(ix, 0, ter.cms - 1).astype(np.int64) + ter.cms * np.clip(iy, 0, ter.cms - 1).astype(np.int64)]
        acc = acc + np.where(c > p.lethal_thresh, p.lethal_penalty, f32(0.0))
        acc = acc + c
    return acc, cm_ij


def path_orientation(p, x, y, gx, gy, traj):
    """critics_warp.py:44-83."""
    xd, yd = f32(gx) - f32(x), f32(gy) - f32(y)
    xd2 = traj[:, -1, 0] - traj[:, -2, 0]
    yd2 = traj[:, -1, 1] - traj[:, -2, 1]
    sp = xd * xd2 + yd * yd2
    return np.where(sp <= 0, -sp / (np.abs(xd) + np.abs(yd)), f32(0.0)).astype(np.float32)


def avoid_slope_path(p, traj):
    """critics_warp.py:131-166."""
    N, T, _ = traj.shape
    total = np.zeros(N, np.float32)
    one = f32(1.0)
    for i in range(0, T - 3, 2):
        c, pv = traj[:, i + 2], traj[:, i]
        dz = c[:, 2] - pv[:, 2]
        d = np.sqrt((c[:, 0] - pv[:, 0]) * (c[:, 0] - pv[:, 0]) + (c[:, 1] - pv[:, 1]) * (c[:, 1] - pv[:, 1]))
        ratio = np.abs(dz / (d + p.slope_eps))
        total = total + (one + p.slope_gain * ratio) * (one + p.slope_gain * ratio)
    return total


def goal_angle(p, x, y, gx, gy, gtheta, traj):
    """critics_warp.py:5-41 (libm atan; the C oracle's MATH_DET mode uses the specified atan instead)."""
    x, y, gx, gy = f32(x), f32(y), f32(gx), f32(gy)
    dist = np.sqrt((x - gx) * (x - gx) + (y - gy) * (y - gy))
    if not dist < p.goal_angle_radius:
        return np.zeros(traj.shape[0], np.float32)
    with np.errstate(divide="ignore", invalid="ignore"):
        q = (traj[:, -1, 1] - traj[:, -2, 1]) / (traj[:, -1, 0] - traj[:, -2, 0])
    return np.abs(np.arctan(q).astype(np.float32) - f32(gtheta))


def roll_pitch_effort(p, lw, rw, heading, u1, u2):
    """Extensions (no reference counterpart): see include/mppi_b200.h MppiParams.cw_roll / cw_pitch / cw_effort."""
    N, T, _ = lw.shape
    roll = np.zeros(N, np.float32)
    pitch = np.zeros(N, np.float32)
    effort = np.zeros(N, np.float32)
    track = f32(2.0) * p.wheel_offset
    for t in range(0, T, 2):
        r = (lw[:, t, 2] - rw[:, t, 2]) / track
        roll = roll + r * r
        pitch = pitch + heading[:, t, 2] * heading[:, t, 2]
    for t in range(T):
        effort = effort + (u1[:, t] * u1[:, t] + u2[:, t] * u2[:, t])
    return roll, pitch, effort


def mppi_step(p: P, dem, half_width, costmap, state: dict, nom1, nom2, eps1, eps2, want_dump=True):
    """One full MPPI step (MPPI_isaac.py:505-720).  Returns a dict of every intermediate."""
    ter = Terrain(dem, half_width, costmap)
    st = {k: f32(v) for k, v in state.items()}
    u1, u2 = sample_inputs(p, nom1, nom2, st["sigma1"], st["sigma2"], eps1, eps2)
    if p.input_model == 1:
        v, w = u1, u2                                             # velocity-space samples are (v, w) themselves
    else:
        v, w = inputs_to_velocities(p, u1, u2, st["wheel_l"], st["wheel_r"], p.filt_k, p.filt_a)
    h0 = np.array([st["hx"], st["hy"], st["hz"]], np.float32)
    ro = rollout(p, ter, st["x"], st["y"], h0, v, w)
    c_path = path_follow(p, st["x"], st["y"], st["goal_x"], st["goal_y"], ro["traj"])
    c_slope = avoid_slope_wheels(p, ro["lw"], ro["rw"])
    c_speed = maximise_speed(p, st["x"], st["y"], st["goal_x"], st["goal_y"], v)
    c_obs, cm_ij = avoid_obstacle(p, ter, ro["traj"])
    x_orient = path_orientation(p, st["x"], st["y"], st["goal_x"], st["goal_y"], ro["traj"])
    x_slope = avoid_slope_path(p, ro["traj"])
    x_angle = goal_angle(p, st["x"], st["y"], st["goal_x"], st["goal_y"], st.get("goal_theta", 0.0), ro["traj"])
    x_roll, x_pitch, x_effort = roll_pitch_effort(p, ro["lw"], ro["rw"], ro["heading"], u1, u2)
    cost = np.zeros(p.K, np.float32)                              # critics_warp.py:324-329 (+ optional terms)
    if p.cw_orient != 0:
        cost = cost + p.cw_orient * x_orient
    cost = cost + p.cw_path * c_path
    if p.cw_slope_path != 0:
        cost = cost + p.cw_slope_path * x_slope
    cost = cost + p.cw_slope * c_slope
    cost = cost + p.cw_speed * c_speed
    cost = cost + p.cw_obs * c_obs
    for wgt, val in ((p.cw_goal_angle, x_angle), (p.cw_roll, x_roll), (p.cw_pitch, x_pitch), (p.cw_effort, x_effort)):
        if wgt != 0:
            cost = cost + wgt * val
    # update, critics_warp.py:338-376 (race-free intent, old_files/run_mppi.py:222-226)
    ceff = np.where(np.isnan(cost), np.float32(np.inf), cost)     # a NaN cost gets zero weight (product rule, stats[4])
    m = ceff.min()
    argmin = int(np.argmin(ceff))
    with np.errstate(over="ignore"):
        wts = np.exp(-(ceff - m) / p.lam).astype(np.float32)
    S = f32(0.0)
    for k in range(p.K):                                          # sequential fp32 atomic_add order 0..K-1
        S = f32(S + wts[k])
    nz = np.nonzero(wts)[0]
    n1 = np.zeros(p.T, np.float32)
    n2 = np.zeros(p.T, np.float32)
    for k in nz:                                                  # zero weights add exactly +0
        n1 = n1 + wts[k] * u1[k] / S
        n2 = n2 + wts[k] * u2[k] / S
    if not S > 0:                                                 # no valid sample: the product keeps the nominal
        n1, n2 = np.asarray(nom1, np.float32).copy(), np.asarray(nom2, np.float32).copy()
    if p.input_model == 1:
        ov, ow = n1[None, :].copy(), n2[None, :].copy()
    else:
        ov, ow = inputs_to_velocities(p, n1[None, :], n2[None, :], st["wheel_l"], st["wheel_r"], p.opt_k, p.opt_a)
    p3 = P(1, p.T, proj=3)
    p3.__dict__.update({k: v_ for k, v_ in p.__dict__.items() if k not in ("K", "proj")})
    p3.K, p3.proj = 1, 3
    sim = rollout(p3, ter, st["x"], st["y"], h0, ov, ow)
    out = dict(u1=u1, u2=u2, v=v, w=w, critics=np.stack([c_path, c_slope, c_speed, c_obs], axis=1), cost=cost,
               critics_ext=np.stack([x_orient, x_slope, x_angle, x_roll, x_pitch, x_effort], axis=1),
               cm_ij=cm_ij, weights=wts, min_cost=float(m), argmin=argmin, weights_sum=float(S), nominal1=n1,
               nominal2=n2, opt_v=ov[0], opt_w=ow[0], sim_traj=sim["traj"][0], sim_heading=sim["heading"][0])
    out.update(ro)
    return out
