"""Synthetic inputs of the benchmark / parity workloads (SURVEY.md §8d "scene v0" and the pose correspondence sets).
Pure numpy, deterministic per frame index. Units: millimetres / pixels / radians."""
import numpy as np

from . import abi

# world <- camera axis change of the reference: x-forward/y-left/z-up <- x-right/y-down/z-forward
# (camera_transformation.cpp:11-17; numerically [[0,0,1],[-1,0,0],[0,-1,0]])
C_CAM_TO_WORLD = np.array([[0.0, 0.0, 1.0], [-1.0, 0.0, 0.0], [0.0, -1.0, 0.0]])


def intrinsics(scale=1):
    return (550.0 * scale, 550.0 * scale, 320.0 * scale, 240.0 * scale)


def scene_v0_depth(frame_index, width=640, height=480, sigma=1.0, zero_prob=0.02):
    """One synthetic depth frame: four planes + a vertical cylinder, N(0, sigma) range noise on valid pixels and 2 %
    invalid pixels. Returns float32 [H, W]. RNG = default_rng(frame_index): normal((H,W)) then random((H,W))."""
    scale = width / 640.0
    fx, fy, cx, cy = intrinsics(scale)
    u = np.arange(width, dtype=np.float64)
    v = np.arange(height, dtype=np.float64)
    dx = ((u - cx) / fx)[None, :]
    dy = ((v - cy) / fy)[:, None]
    dx = np.broadcast_to(dx, (height, width))
    dy = np.broadcast_to(dy, (height, width))
    planes = [((0.0, 0.0, -1.0), 2500.0), ((0.0, -1.0, -0.05), 900.0), ((1.0, 0.0, -0.2), 1400.0),
              ((-1.0, 0.0, -0.3), 1600.0)]
    depth = np.full((height, width), np.inf)
    for n, d0 in planes:
        n = np.asarray(n, dtype=np.float64)
        n = n / np.linalg.norm(n)
        nr = n[0] * dx + n[1] * dy + n[2]
        with np.errstate(divide="ignore", invalid="ignore"):
            t = -d0 / nr
        ok = (nr < 0) & (t > 0)
        depth = np.where(ok & (t < depth), t, depth)
    # vertical cylinder (axis parallel to y) through (x, z) = (300, 1800), radius 250: near root
    a = dx * dx + 1.0
    b = -2.0 * (300.0 * dx + 1800.0)
    c = 300.0 ** 2 + 1800.0 ** 2 - 250.0 ** 2
    disc = b * b - 4 * a * c
    with np.errstate(invalid="ignore"):
        t = (-b - np.sqrt(disc)) / (2 * a)
    ok = (disc > 0) & (t > 0)
    depth = np.where(ok & (t < depth), t, depth)
    valid = np.isfinite(depth)
    depth = np.where(valid, depth, 0.0)
    rng = np.random.default_rng(frame_index)
    noise = rng.normal(0.0, sigma, (height, width))
    drop = rng.random((height, width)) < zero_prob
    depth = np.where(valid, depth + noise, 0.0)
    depth = np.where(drop, 0.0, depth)
    return depth.astype(np.float32)


def scene_v0_batch(first_frame, batch, width=640, height=480):
    return np.stack([scene_v0_depth(first_frame + i, width, height) for i in range(batch)])


def random_scene_depth(seed, width=640, height=480):
    """A randomly furnished room for parity sweeps: a back wall, 2-5 further random planes, 0-3 cylinders with random
    axes and 0-1 sphere; range noise N(0, sigma) with sigma in [0.5, 3] mm and 0-6 % invalid pixels, all drawn from
    default_rng(seed). Returns float32 [H, W]. The geometry is arbitrary on purpose (slanted cylinders, small regions,
    planes that nearly merge): it is there to walk the region-growing / cylinder / merge branches, not to look real."""
    rng = np.random.default_rng(1_000_003 * 7 + seed)
    scale = width / 640.0
    fx, fy, cx, cy = intrinsics(scale)
    u = np.arange(width, dtype=np.float64)
    v = np.arange(height, dtype=np.float64)
    dx = np.broadcast_to(((u - cx) / fx)[None, :], (height, width))
    dy = np.broadcast_to(((v - cy) / fy)[:, None], (height, width))
    depth = np.full((height, width), np.inf)

    def take(t, ok):
        nonlocal depth
        depth = np.where(ok & (t > 200.0) & (t < depth), t, depth)

    planes = [(np.array([0.0, 0.0, -1.0]) + rng.normal(0, 0.08, 3), rng.uniform(2500.0, 4500.0))]
    for _ in range(int(rng.integers(2, 6))):
        n = rng.normal(0, 1, 3)
        n[2] = -abs(n[2]) - 0.15
        planes.append((n, rng.uniform(700.0, 2500.0)))
    for n, d0 in planes:
        n = n / np.linalg.norm(n)
        nr = n[0] * dx + n[1] * dy + n[2]
        with np.errstate(divide="ignore", invalid="ignore"):
            t = -d0 / nr
        take(t, nr < 0)
    for _ in range(int(rng.integers(0, 4))):
        # |(o + t r - c) - ((o + t r - c) . a) a|^2 = R^2 with o = 0
        c = np.array([rng.uniform(-900, 900), rng.uniform(-600, 600), rng.uniform(1200, 2600)])
        a = rng.normal(0, 1, 3)
        a /= np.linalg.norm(a)
        R = rng.uniform(120.0, 450.0)
        ra = dx * a[0] + dy * a[1] + a[2]
        ca = float(c @ a)
        A = (dx * dx + dy * dy + 1.0) - ra * ra
        B = -2.0 * ((dx * c[0] + dy * c[1] + c[2]) - ra * ca)
        C = float(c @ c) - ca * ca - R * R
        disc = B * B - 4 * A * C
        with np.errstate(invalid="ignore", divide="ignore"):
            t = (-B - np.sqrt(disc)) / (2 * A)
        take(t, disc > 0)
    if rng.random() < 0.5:
        c = np.array([rng.uniform(-700, 700), rng.uniform(-500, 500), rng.uniform(1200, 2400)])
        R = rng.uniform(150.0, 400.0)
        A = dx * dx + dy * dy + 1.0
        B = -2.0 * (dx * c[0] + dy * c[1] + c[2])
        C = float(c @ c) - R * R
        disc = B * B - 4 * A * C
        with np.errstate(invalid="ignore"):
            t = (-B - np.sqrt(disc)) / (2 * A)
        take(t, disc > 0)
    valid = np.isfinite(depth)
    sigma = rng.uniform(0.5, 3.0)
    zero_prob = rng.uniform(0.0, 0.06)
    noise = rng.normal(0.0, sigma, (height, width))
    drop = rng.random((height, width)) < zero_prob
    depth = np.where(valid & ~drop, depth + noise, 0.0)
    return depth.astype(np.float32)


def random_scene_batch(first_seed, batch, width=640, height=480):
    return np.stack([random_scene_depth(first_seed + i, width, height) for i in range(batch)])


# ---- pose helpers (numpy restatement used only to BUILD synthetic correspondences) --------------------
def quat_from_euler(yaw, pitch, roll):
    """AngleAxis(roll,X)*AngleAxis(pitch,Y)*AngleAxis(yaw,Z) as (w,x,y,z) (angle_utils.cpp:6-11)."""
    def q(axis, ang):
        s = np.sin(ang / 2)
        return np.array([np.cos(ang / 2), *(s * np.asarray(axis, dtype=np.float64))])

    def mul(a, b):
        return np.array([a[0] * b[0] - a[1] * b[1] - a[2] * b[2] - a[3] * b[3],
                         a[0] * b[1] + a[1] * b[0] + a[2] * b[3] - a[3] * b[2],
                         a[0] * b[2] + a[2] * b[0] + a[3] * b[1] - a[1] * b[3],
                         a[0] * b[3] + a[3] * b[0] + a[1] * b[2] - a[2] * b[1]])
    return mul(mul(q((1, 0, 0), roll), q((0, 1, 0), pitch)), q((0, 0, 1), yaw))


def quat_to_rot(q):
    w, x, y, z = q
    return np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)],
                     [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
                     [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)]])


def camera_to_world(pose7):
    """4x4 camera->world of a pose (x y z qw qx qy qz): C * [R t; 0 1] (camera_transformation.cpp:19-23)."""
    T = np.eye(4)
    T[:3, :3] = quat_to_rot(pose7[3:7])
    T[:3, 3] = pose7[:3]
    Cm = np.eye(4)
    Cm[:3, :3] = C_CAM_TO_WORLD
    return Cm @ T


def pose_correspondences(frame_index, n_points=300, n_planes=20, outlier_frac=0.1, scale=1, guess=0.9, n_points2d=0,
                         position=(10.0, 10.0, 10.0), euler_deg=(45.0, -45.0, 20.0), pixel_sigma=0.5):
    """Matched features of one frame: returns (true_pose7, guess_pose7, matches[n_points+n_planes(+n_points2d)]).
    The last outlier_frac of each kind are outliers built like the reference's tests build them. n_points2d > 0 appends
    inverse-depth features (RS_FEAT_POINT2D, the "line" residual), first observed from the true camera position."""
    rng = np.random.default_rng(1000 + frame_index)
    fx, fy, cx, cy = intrinsics(scale)
    d2r = np.pi / 180.0
    yaw, pitch, roll = euler_deg[0] * d2r, euler_deg[1] * d2r, euler_deg[2] * d2r
    position = np.asarray(position, dtype=np.float64)
    true_pose = np.concatenate([position, quat_from_euler(yaw, pitch, roll)])
    guess_pose = np.concatenate([position * guess, quat_from_euler(yaw * guess, pitch * guess, roll * guess)])
    c2w = camera_to_world(true_pose)
    w2c = np.linalg.inv(c2w)
    m = np.zeros((n_points + n_planes + n_points2d,), dtype=abi.match_dtype)

    # points: uniform in a 2 m x 2 m x (1..3 m) frustum in front of the true pose
    pc = np.stack([rng.uniform(-1000, 1000, n_points), rng.uniform(-1000, 1000, n_points), rng.uniform(1000, 3000, n_points)], 1)
    pw = (c2w[:3, :3] @ pc.T).T + c2w[:3, 3]
    uv = np.stack([fx * pc[:, 0] / pc[:, 2] + cx, fy * pc[:, 1] / pc[:, 2] + cy], 1) + rng.normal(0, pixel_sigma, (n_points, 2))
    n_out_pt = int(round(n_points * outlier_frac))
    if n_out_pt:
        uv[n_points - n_out_pt:, 0] = rng.uniform(0, 640 * scale, n_out_pt)
        uv[n_points - n_out_pt:, 1] = rng.uniform(0, 480 * scale, n_out_pt)
    m["type"][:n_points] = abi.RS_FEAT_POINT
    m["obs"][:n_points, :2] = uv
    m["map"][:n_points, :3] = pw
    m["sigma"][:n_points, :3] = 5.0

    # planes: random world planes transformed exactly to the camera frame (+ U(-5,5) mm on the observed d)
    nw = rng.normal(size=(n_planes, 3))
    nw /= np.linalg.norm(nw, axis=1, keepdims=True)
    dw = rng.uniform(500, 3000, n_planes)
    Rc, tc = c2w[:3, :3], c2w[:3, 3]
    ncam = (Rc.T @ nw.T).T
    dcam = nw @ tc + dw
    dcam = dcam + rng.uniform(-5, 5, n_planes)
    n_out_pl = int(round(n_planes * outlier_frac))
    if n_out_pl:
        r = rng.normal(size=(n_out_pl, 3))
        ncam[n_planes - n_out_pl:] = r / np.linalg.norm(r, axis=1, keepdims=True)
        dcam[n_planes - n_out_pl:] = rng.uniform(-100, 100, n_out_pl)
    sl = slice(n_points, n_points + n_planes)
    m["type"][sl] = abi.RS_FEAT_PLANE
    m["obs"][sl, :3] = ncam
    m["obs"][sl, 3] = dcam
    m["map"][sl, :3] = nw
    m["map"][sl, 3] = dw
    m["sigma"][sl] = np.array([0.01, 0.01, 0.01, 1.0])
    if n_points2d:
        # inverse-depth points: bearing (theta, phi) from the first observation towards a world point, inverse depth
        # 1 / distance; two thirds with a small inverse-depth sigma (both depth estimates collapse on the far point,
        # inverse_depth_coordinates.cpp:142-154), one third with a large one (the furthest estimate falls behind the
        # first observation, so the screen line really is a line)
        q = np.stack([rng.uniform(-1000, 1000, n_points2d), rng.uniform(-1000, 1000, n_points2d),
                      rng.uniform(1000, 3000, n_points2d)], 1)
        qw = (c2w[:3, :3] @ q.T).T + c2w[:3, 3]
        origin = c2w[:3, 3]
        v = qw - origin
        dist = np.linalg.norm(v, axis=1)
        theta = np.arccos(v[:, 2] / dist)
        phi = np.arctan2(v[:, 1], v[:, 0])
        uv2 = np.stack([fx * q[:, 0] / q[:, 2] + cx, fy * q[:, 1] / q[:, 2] + cy], 1) + rng.normal(0, 0.5, (n_points2d, 2))
        n_out = int(round(n_points2d * outlier_frac))
        if n_out:
            uv2[n_points2d - n_out:, 0] = rng.uniform(0, 640 * scale, n_out)
            uv2[n_points2d - n_out:, 1] = rng.uniform(0, 480 * scale, n_out)
        sl2 = slice(n_points + n_planes, n_points + n_planes + n_points2d)
        m["type"][sl2] = abi.RS_FEAT_POINT2D
        m["obs"][sl2, :2] = uv2
        m["obs"][sl2, 2] = theta
        m["obs"][sl2, 3] = phi
        m["map"][sl2, :3] = origin
        m["map"][sl2, 3] = 1.0 / dist
        sig_d = np.where(np.arange(n_points2d) % 3 == 2, 1e-6, 1e-8)
        m["sigma"][sl2, 0] = sig_d
        m["sigma"][sl2, 1] = 0.002
        m["sigma"][sl2, 2] = 0.002
    del w2c
    return true_pose, guess_pose, m


def random_pose_problem(index):
    """A random pose problem for parity sweeps: 5-300 points, 0-20 planes, sometimes 1-12 inverse-depth points, 0-45 %
    outliers, true pose up to 2 m / 60 degrees per axis away, initial guess 0.5-1.0 x truth, pixel noise 0.2-1.5 px. About
    a third are 'hard' (0-24 points, 0-4 planes, 30-90 % outliers): rejected frames, early exits, thin consensus sets.
    Returns (true_pose7, guess_pose7, matches); at most 336 matches."""
    rng = np.random.default_rng(77_000 + index)
    hard = rng.random() < 0.35
    kw = dict(n_points=int(rng.integers(0, 25)) if hard else int(rng.integers(5, 301)),
              n_planes=int(rng.integers(0, 5)) if hard else int(rng.integers(0, 21)),
              n_points2d=int(rng.integers(0, 13)) if rng.random() < 0.3 else 0,
              outlier_frac=float(rng.uniform(0.3, 0.9)) if hard else float(rng.uniform(0, 0.45)),
              guess=float(rng.uniform(0.5, 1.0)),
              position=rng.uniform(-2000, 2000, 3), euler_deg=rng.uniform(-60, 60, 3),
              pixel_sigma=float(rng.uniform(0.2, 1.5)))
    return pose_correspondences(index, **kw)


def pose_batch(first_frame, batch, max_matches, **kw):
    cur = np.zeros((batch, 7))
    truth = np.zeros((batch, 7))
    matches = np.zeros((batch, max_matches), dtype=abi.match_dtype)
    n = np.zeros((batch,), dtype=np.int32)
    for b in range(batch):
        t, g, m = pose_correspondences(first_frame + b, **kw)
        truth[b], cur[b] = t, g
        matches[b, :len(m)] = m
        n[b] = len(m)
    return truth, cur, matches, n


def star_polygon(rng, n, r_min=200.0, r_max=900.0, center=(0.0, 0.0), clockwise=False, closed=False):
    """A simple (star-shaped, generally concave) ring of n vertices around `center`, mm."""
    while True:   # every angular gap below pi: the ring is star-shaped about `center`, hence simple
        ang = np.sort(rng.uniform(0.0, 2 * np.pi, n))
        if np.max(np.diff(np.concatenate([ang, ang[:1] + 2 * np.pi]))) < 0.9 * np.pi:
            break
    ang += np.arange(n) * 1e-9   # no two vertices on one ray
    r = rng.uniform(r_min, r_max, n)
    ring = np.stack([center[0] + r * np.cos(ang), center[1] + r * np.sin(ang)], axis=1)
    if clockwise:
        ring = ring[::-1]
    if closed:
        ring = np.concatenate([ring, ring[:1]])
    return np.ascontiguousarray(ring)


def _plane_frame(normal):
    """x / y axes spanning the plane of `normal` (the role of utils::get_plane_coordinate_system, polygon.cpp:74-115)."""
    r = np.array([1.0, 0.0, 0.0]) if abs(normal[0]) < 0.9 else np.array([0.0, 1.0, 0.0])
    x = np.cross(normal, r)
    x /= np.linalg.norm(x)
    y = np.cross(normal, x)
    return x, y / np.linalg.norm(y)


def plane_match_problem(seed, n_frames=4, n_det=6, n_extra_map=2, max_vertices=24):
    """Synthetic input of rs_plane_match: per frame a camera pose, n_det detected planes (camera frame) with star-shaped
    boundary polygons, and a local map holding a perturbed world-space copy of most detections plus unrelated planes.
    Returns (w2c [F,4,4], det, det_first, det_xy, map, map_first, map_xy, det_matched)."""
    rng = np.random.default_rng(seed)
    det, mp, det_xy, map_xy, det_first, map_first, w2cs, matched = [], [], [], [], [0], [0], [], []

    def add(lst, xy_list, normal, d, center, xa, ya, ring):
        rec = np.zeros((), dtype=abi.polygon_plane_dtype)
        rec["normal"], rec["d"], rec["center"], rec["x_axis"], rec["y_axis"] = normal, d, center, xa, ya
        rec["first_vertex"], rec["n_vertices"] = sum(len(r) for r in xy_list), len(ring)
        lst.append(rec)
        xy_list.append(ring)

    for f in range(n_frames):
        q = quat_from_euler(*rng.uniform(-0.6, 0.6, 3))
        pose = np.concatenate([rng.uniform(-1500, 1500, 3), q])
        c2w = camera_to_world(pose)
        w2c = np.linalg.inv(c2w)
        w2cs.append(w2c)
        R, t = c2w[:3, :3], c2w[:3, 3]
        nd = int(rng.integers(max(1, n_det - 2), n_det + 1))
        for k in range(nd):
            n = rng.normal(size=3)
            n /= np.linalg.norm(n)
            center = rng.uniform(-800, 800, 3) + np.array([0, 0, 2500.0])
            d = -float(n @ center)
            xa, ya = _plane_frame(n)
            nv = int(rng.integers(3, max_vertices + 1))
            ring = star_polygon(rng, nv, clockwise=bool(rng.integers(2)), closed=bool(rng.integers(2)))
            add(det, det_xy, n, d, center, xa, ya, ring)
            matched.append(rng.random() < 0.1)
            if rng.random() < 0.8:   # the map's view of this plane: shifted / rotated in-plane / slightly tilted, other polygon
                tilt = rng.normal(scale=0.05, size=3)
                nm = n + tilt
                nm /= np.linalg.norm(nm)
                cm = center + rng.uniform(-250, 250, 3)
                cm -= nm * (nm @ cm + d + rng.uniform(-60, 60))   # onto the (offset) plane
                xm, ym = _plane_frame(nm)
                a = rng.uniform(0, 2 * np.pi)
                xm, ym = np.cos(a) * xm + np.sin(a) * ym, -np.sin(a) * xm + np.cos(a) * ym
                ring_m = star_polygon(rng, int(rng.integers(3, max_vertices + 1)), r_min=150.0, r_max=1100.0,
                                      clockwise=bool(rng.integers(2)), closed=bool(rng.integers(2)))
                nw, cw = R @ nm, R @ cm + t
                add(mp, map_xy, nw, -float(nw @ cw), cw, R @ xm, R @ ym, ring_m)
        for k in range(n_extra_map):
            n = rng.normal(size=3)
            n /= np.linalg.norm(n)
            cw = rng.uniform(-3000, 3000, 3)
            xa, ya = _plane_frame(n)
            add(mp, map_xy, n, -float(n @ cw), cw, xa, ya, star_polygon(rng, int(rng.integers(3, max_vertices + 1))))
        det_first.append(len(det))
        map_first.append(len(mp))
    return (np.stack(w2cs), np.array(det, dtype=abi.polygon_plane_dtype), np.array(det_first, np.int32), np.concatenate(det_xy),
            np.array(mp, dtype=abi.polygon_plane_dtype), np.array(map_first, np.int32), np.concatenate(map_xy),
            np.array(matched, np.uint8))
