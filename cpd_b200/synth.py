"""Synthetic Waymo-shaped LiDAR sweeps.

A spinning multi-ring sensor at (0, 0, 2.1 m) is ray-cast against a ground plane,
~40 car-sized boxes and ~8 wall slabs (more occluders starve the far field); hits get 1 cm noise, are clipped to the CPD
range (tools/cfgs/dataset_configs/waymo_unsupervised/waymo_unsupervised_cproto.yaml:118)
and randomly permuted like the train-time shuffle (data_processor.py:105-126).
Point features follow the wire format of waymo_unsupervised_dataset.py:137-144:
[x, y, z, intensity in (0,1), elongation/time = 0], float32.
"""
import numpy as np

PC_RANGE = np.array([-75.2, -75.2, -2.0, 75.2, 75.2, 4.0], dtype=np.float32)
VOXEL_SIZE = np.array([0.1, 0.1, 0.15], dtype=np.float32)
SENSOR_Z = 2.1


def _scene(rng, n_cars, n_walls):
    lo, hi = [], []
    for _ in range(n_cars):
        r = rng.uniform(4.0, 72.0)
        a = rng.uniform(0, 2 * np.pi)
        cx, cy = r * np.cos(a), r * np.sin(a)
        dx, dy, dz = rng.uniform(3.8, 5.2), rng.uniform(1.7, 2.1), rng.uniform(1.4, 2.0)
        if rng.random() < 0.5:
            dx, dy = dy, dx
        lo.append([cx - dx / 2, cy - dy / 2, 0.0])
        hi.append([cx + dx / 2, cy + dy / 2, dz])
    for _ in range(n_walls):
        r = rng.uniform(12.0, 70.0)
        a = rng.uniform(0, 2 * np.pi)
        cx, cy = r * np.cos(a), r * np.sin(a)
        ln, th, ht = rng.uniform(8.0, 30.0), rng.uniform(0.3, 0.8), rng.uniform(2.5, 6.0)
        if rng.random() < 0.5:
            ln, th = th, ln
        lo.append([cx - ln / 2, cy - th / 2, 0.0])
        hi.append([cx + ln / 2, cy + th / 2, ht])
    return np.asarray(lo, np.float64), np.asarray(hi, np.float64)


def _cast(dirs, lo, hi):
    """nearest hit distance of rays from the sensor along unit `dirs` (R,3); inf = miss."""
    o = np.array([0.0, 0.0, SENSOR_Z])
    t_best = np.full(dirs.shape[0], np.inf)
    dz = dirs[:, 2]
    down = dz < -1e-6
    t_best[down] = -SENSOR_Z / dz[down]
    with np.errstate(divide="ignore", invalid="ignore"):
        inv = 1.0 / dirs
        for s in range(0, dirs.shape[0], 32768):
            iv = inv[s:s + 32768, None, :]
            t1 = (lo[None] - o) * iv
            t2 = (hi[None] - o) * iv
            tmin = np.minimum(t1, t2).max(axis=2)
            tmax = np.maximum(t1, t2).min(axis=2)
            hit = (tmax >= np.maximum(tmin, 0.0)) & (tmin > 0.0)
            t = np.where(hit, tmin, np.inf).min(axis=1)
            t_best[s:s + 32768] = np.minimum(t_best[s:s + 32768], t)
    return t_best


def synth_scan(n_points, seed, rings=64, n_cars=40, n_walls=8, surfaces=1.0, el_jitter=1.2e-3, ring_pow=0.8):
    """One sweep of exactly `n_points` in-range points, shape (n_points, 5) float32.

    `rings`/`surfaces` raise the number of distinct occupied cells (the dense-scene
    stress config needs >=192 rings to reach ~200 k voxels)."""
    rng = np.random.default_rng(seed)
    lo, hi = _scene(rng, int(n_cars * surfaces), int(n_walls * surfaces))
    # ring elevations: 80 % of the rings are spaced so their ground-hit radii cover
    # 6.6 m .. 75 m (dense near the horizon, like a real spinning lidar), the rest look
    # slightly up/level and only return from objects.
    n_down = max(1, int(round(rings * 0.8)))
    radii = 6.6 + (74.0 - 6.6) * (np.arange(n_down) / max(n_down - 1, 1)) ** ring_pow
    elev = np.concatenate([-np.arctan2(SENSOR_Z, radii), np.deg2rad(np.linspace(-1.2, 2.4, rings - n_down))])
    oversample = 1.35
    for _ in range(6):
        n_az = int(np.ceil(oversample * n_points / rings))
        az = (np.arange(n_az) + rng.uniform(0, 1)) * (2 * np.pi / n_az)
        el = np.repeat(elev, n_az) + rng.normal(0, el_jitter, rings * n_az)
        azs = np.tile(az, rings)
        dirs = np.stack([np.cos(el) * np.cos(azs), np.cos(el) * np.sin(azs), np.sin(el)], axis=1)
        t = _cast(dirs, lo, hi)
        ok = np.isfinite(t) & (t < 110.0)
        pts = dirs[ok] * t[ok, None]
        pts[:, 2] += SENSOR_Z
        pts += rng.normal(0, 0.01, pts.shape)
        inr = ((pts[:, 0] > PC_RANGE[0]) & (pts[:, 0] < PC_RANGE[3]) & (pts[:, 1] > PC_RANGE[1]) &
               (pts[:, 1] < PC_RANGE[4]) & (pts[:, 2] > PC_RANGE[2]) & (pts[:, 2] < PC_RANGE[5]))
        pts = pts[inr]
        if pts.shape[0] >= n_points:
            break
        oversample *= 1.1 * n_points / max(pts.shape[0], 1)
    assert pts.shape[0] >= n_points, "scene too sparse for the requested point count"
    sel = rng.permutation(pts.shape[0])[:n_points]
    out = np.zeros((n_points, 5), np.float32)
    out[:, :3] = pts[sel].astype(np.float32)
    out[:, 3] = rng.uniform(0, 1, n_points).astype(np.float32)
    return out


def synth_gt_boxes(n_boxes, seed, num_class=3):
    """(n_boxes, 8) float32 [x, y, z, dx, dy, dz, heading, class in 1..num_class]."""
    rng = np.random.default_rng(seed + 7919)
    b = np.zeros((n_boxes, 8), np.float32)
    r = rng.uniform(3.0, 70.0, n_boxes)
    a = rng.uniform(0, 2 * np.pi, n_boxes)
    b[:, 0], b[:, 1] = r * np.cos(a), r * np.sin(a)
    b[:, 2] = rng.uniform(0.5, 1.2, n_boxes)
    b[:, 3] = rng.uniform(0.6, 5.2, n_boxes)
    b[:, 4] = rng.uniform(0.6, 2.2, n_boxes)
    b[:, 5] = rng.uniform(1.2, 2.2, n_boxes)
    b[:, 6] = rng.uniform(-np.pi, np.pi, n_boxes)
    b[:, 7] = rng.integers(1, num_class + 1, n_boxes)
    return b


def synth_nms_boxes(n, seed, clusters=None):
    """Detection-like boxes (n,7) + scores (n,): clusters of near-duplicates so that
    NMS at IoU 0.3-0.8 has real work."""
    rng = np.random.default_rng(seed)
    clusters = clusters or max(1, n // 6)
    cx = rng.uniform(-70, 70, clusters)
    cy = rng.uniform(-70, 70, clusters)
    cd = np.stack([rng.uniform(3.5, 5.0, clusters), rng.uniform(1.6, 2.1, clusters), rng.uniform(1.4, 1.9, clusters)], 1)
    ch = rng.uniform(-np.pi, np.pi, clusters)
    which = rng.integers(0, clusters, n)
    b = np.zeros((n, 7), np.float32)
    b[:, 0] = cx[which] + rng.normal(0, 0.25, n)
    b[:, 1] = cy[which] + rng.normal(0, 0.25, n)
    b[:, 2] = rng.uniform(0.5, 1.2, n)
    b[:, 3:6] = cd[which] * rng.uniform(0.9, 1.1, (n, 3))
    b[:, 6] = ch[which] + rng.normal(0, 0.08, n)
    scores = rng.uniform(0.1, 1.0, n).astype(np.float32)
    return b, scores


def count_voxels(points, pc_range=PC_RANGE, voxel_size=VOXEL_SIZE):
    """Number of distinct occupied cells of one sweep (same fp32 cell arithmetic as the voxelizer)."""
    p = points[:, :3].astype(np.float32)
    cell = np.floor((p - pc_range[:3]) / voxel_size).astype(np.int64)
    grid = np.round((pc_range[3:] - pc_range[:3]) / voxel_size).astype(np.int64)
    ok = ((cell >= 0) & (cell < grid)).all(1)
    cell = cell[ok]
    return int(np.unique((cell[:, 2] * grid[1] + cell[:, 1]) * grid[0] + cell[:, 0]).size)


def synth_dense_scan(n_points=300000, seed=0, target_voxels=200000, tol=0.05):
    """BASELINE configs[4], the dense-scene stress sweep: `n_points` points occupying target_voxels +- tol active voxels.
    A sparse scene (few occluders) with strong elevation jitter spreads the rings over the far field; the ring spacing
    exponent is then bisected until the voxel count lands in the window (the count depends on the random scene; a scene
    whose occluders starve it is thinned out once).  Returns (points (n, 5) float32, voxel count)."""
    best = None
    for n_cars, n_walls in ((12, 2), (4, 0)):
        lo, hi = 0.3, 2.4
        for _ in range(8):
            rp = 0.5 * (lo + hi)
            pts = synth_scan(n_points, seed, rings=64, n_cars=n_cars, n_walls=n_walls, el_jitter=8e-3, ring_pow=rp)
            m = count_voxels(pts)
            if best is None or abs(m - target_voxels) < abs(best[1] - target_voxels):
                best = (pts, m)
            if abs(m - target_voxels) <= tol * target_voxels:
                return best
            if m < target_voxels:            # the count falls as the exponent grows (rings crowd towards the sensor)
                hi = rp
            else:
                lo = rp
    return best
