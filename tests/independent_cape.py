"""TEST INFRASTRUCTURE - a SECOND, independent restatement of the reference's CAPE path, used to pin the C++ oracle.

Written from the reference sources (file:line cited per function), NOT from oracle/cape.cpp, with different building
blocks wherever the reference calls a third party:
  * numpy.linalg.eigh (LAPACK)             where the reference calls Eigen::SelfAdjointEigenSolver<Matrix3d>
  * the real OpenCV (cv2.erode / cv2.dilate) where the reference calls cv::erode / cv::dilate
  * numpy's MT19937 (init_genrand seeding)  where the reference draws from std::mt19937 through uniform_real_distribution
  * literal recursion for region_growing    where oracle and CUDA kernel argue "reachability closure"
Plain Python loops: slow (about a second per 640x480 frame), which is fine for a checker. Label grids must come out
identical to the oracle's; real-valued outputs agree to the last bits of the two eigen-solvers (tests/test_independent_cape.py).

This pins the restatement against a second reading of the sources; it is NOT the reference binary. The reference itself
cannot be compiled in this image (no Eigen / OpenCV C++ headers); oracle/_ref/README.md holds the recipe that closes the pin
on a box that has them."""
import math
import sys

import cv2
import numpy as np

DBL_MAX = sys.float_info.max
DBL_EPS = sys.float_info.epsilon
f32 = np.float32


# ---- parameters.hpp:16-18, 68-86 --------------------------------------------------------------------------------------
DEPTH_SIGMA_ERROR, DEPTH_SIGMA_MULTIPLIER, DEPTH_SIGMA_MARGIN = 2.73, 0.74, -0.53
MIN_PLANE_SEED_PROPORTION = 0.8 / 100.0
MIN_CELL_ACTIVATED_PROPORTION = 0.65 / 100.0
MIN_ZERO_DEPTH_PROPORTION = f32(0.7)
MAX_PLANE_ANGLE_FOR_MERGE_D = f32(18.0)
MAX_PLANE_DISTANCE_FOR_MERGE_MM = f32(50.0)
CYL_SQRT_MAX_DISTANCE = f32(0.04)
CYL_MIN_SCORE = f32(75)
CYL_INLIER_PROPORTIONS = f32(0.33)
CYL_PROBABILITY_OF_SUCCESS = f32(0.8)


def depth_quantization(depth):
    """utils::get_depth_quantization, covariances.cpp:12-19."""
    sigma_error = DEPTH_SIGMA_ERROR * ((1.0 / 1000.0) * (1.0 / 1000.0))
    sigma_mult = DEPTH_SIGMA_MULTIPLIER / 1000.0
    return max(DEPTH_SIGMA_MARGIN + sigma_mult * depth + sigma_error * (depth * depth), 0.5)


class StdMt19937Uniform:
    """utils::Random::get_random_double on a fresh thread (random.hpp:17-31): std::mt19937(seed) through
    std::uniform_real_distribution<double>(0, 1). libstdc++'s generate_canonical<double, 53> takes two 32-bit draws,
    (first + second * 2^32) / 2^64. numpy's RandomState(seed) seeds MT19937 with the same init_genrand(seed)."""

    def __init__(self, seed):
        self._rs = np.random.RandomState(int(seed) & 0xFFFFFFFF)

    def _raw(self):
        return int(self._rs.randint(0, 1 << 32, dtype=np.uint64))

    def uniform(self):
        lo = self._raw()
        hi = self._raw()
        v = (float(lo) + float(hi) * 4294967296.0) / 18446744073709551616.0
        if v >= 1.0:
            v = math.nextafter(1.0, 0.0)
        return v

    def random_uint(self, max_value):
        """get_random_uint(0, maxValue), random.hpp:55-61."""
        return int(math.floor(self.uniform() * (max_value - 0)))


def eigen_inverse3(K):
    """Eigen's Matrix3d::inverse() (cofactors, determinant from the first column, one reciprocal)."""
    def cof(i, j):
        i1, i2, j1, j2 = (i + 1) % 3, (i + 2) % 3, (j + 1) % 3, (j + 2) % 3
        return K[i1][j1] * K[i2][j2] - K[i1][j2] * K[i2][j1]
    c00, c10, c20 = cof(0, 0), cof(1, 0), cof(2, 0)
    det = (c00 * K[0][0] + c10 * K[1][0]) + c20 * K[2][0]
    inv = 1.0 / det
    out = np.zeros((3, 3))
    out[0, 0], out[0, 1], out[0, 2] = c00 * inv, c10 * inv, c20 * inv
    out[1, 0], out[1, 1], out[1, 2] = cof(0, 1) * inv, cof(1, 1) * inv, cof(2, 1) * inv
    out[2, 0], out[2, 1], out[2, 2] = cof(0, 2) * inv, cof(1, 2) * inv, cof(2, 2) * inv
    return out


def screen_to_camera_factors(K4, W, H):
    """transform_screen_to_camera (point_coordinates.cpp:79-83): (K^-1 * (u, v, 1)).head<2>(), coefficient by coefficient
    as Eigen evaluates a 3x3 * 3x1 product: (a0 u + a1 v) + a2."""
    fx, fy, cx, cy = K4
    Kinv = eigen_inverse3([[fx, 0.0, cx], [0.0, fy, cy], [0.0, 0.0, 1.0]])
    u = np.arange(W, dtype=np.float64)
    v = np.arange(H, dtype=np.float64)
    # the cross terms are exact zeros for a skew-free camera: x depends on u only, y on v only
    camx = (Kinv[0, 0] * u + Kinv[0, 1] * 0.0) + Kinv[0, 2] * 1.0
    camy = (Kinv[1, 0] * 0.0 + Kinv[1, 1] * v) + Kinv[1, 2] * 1.0
    return camx, camy


class PlaneSegment:
    """Plane_Segment (plane_segment.hpp:118-140)."""

    def __init__(self):
        self.clear()

    def clear(self):
        """clear_plane_parameters, plane_segment.cpp:286-308."""
        self.planar = False
        self.count = 0
        self.score = 0.0
        self.mse = DBL_MAX
        self.centroid = np.zeros(3)
        self.normal = np.zeros(3)
        self.d = 0.0
        self.S = np.zeros(9)     # Sx Sy Sz Sxs Sys Szs Sxy Syz Szx

    def copy(self):
        o = PlaneSegment()
        o.planar, o.count, o.score, o.mse, o.d = self.planar, self.count, self.score, self.mse, self.d
        o.centroid, o.normal, o.S = self.centroid.copy(), self.normal.copy(), self.S.copy()
        return o

    def expand(self, other):
        """expand_segment, plane_segment.cpp:170-191."""
        self.S = self.S + other.S
        self.count += other.count

    def fit_plane(self):
        """fit_plane + get_point_cloud_Huygen_covariance, plane_segment.cpp:205-284."""
        self.planar = False
        o = 1.0 / float(self.count)
        Sx, Sy, Sz, Sxs, Sys, Szs, Sxy, Syz, Szx = (float(v) for v in self.S)
        self.centroid = np.array([Sx * o, Sy * o, Sz * o])
        xx = max(0.0, Sxs - (Sx * Sx) * o)
        yy = max(0.0, Sys - (Sy * Sy) * o)
        zz = max(0.0, Szs - (Sz * Sz) * o)
        xy = Sxy - Sx * Sy * o
        xz = Szx - Sx * Sz * o
        yz = Syz - Sy * Sz * o
        A = np.array([[xx, xy, xz], [xy, yy, yz], [xz, yz, zz]])
        det = (A[0, 0] * (A[1, 1] * A[2, 2] - A[1, 2] * A[2, 1]) - A[0, 1] * (A[1, 0] * A[2, 2] - A[1, 2] * A[2, 0])
               + A[0, 2] * (A[1, 0] * A[2, 1] - A[1, 1] * A[2, 0]))
        if abs(det - 0.0) <= DBL_EPS:       # utils::double_equal(det, 0): previous parametrisation / MSE are kept
            return
        w, v = np.linalg.eigh(A)             # ascending, like SelfAdjointEigenSolver
        lam = np.abs(w)
        n = v[:, 0] / np.linalg.norm(v[:, 0])
        d = -float(n @ self.centroid)
        if d <= 0:
            n, d = -n, -d
        n = n / np.linalg.norm(n)            # PlaneCoordinates' constructor normalises again (plane_coordinates.hpp:23)
        self.normal, self.d = n, d
        self.mse = float(lam[0]) * o
        self.score = float(lam[1]) / max(float(lam[0]), 1e-6)
        self.planar = True

    def can_be_merged(self, p, max_match_distance):
        """plane_segment.cpp:322-326."""
        maximum_merge_angle = math.cos(float(MAX_PLANE_ANGLE_FOR_MERGE_D) * math.pi / 180.0)
        cos_angle = float(self.normal @ p.normal)
        dist = float(self.normal @ p.centroid) + self.d
        return cos_angle > maximum_merge_angle and abs(dist) < max_match_distance


def is_continuous(pixel_depth, last):
    """plane_segment.cpp:44-61. Returns (ok, new last)."""
    if pixel_depth > 0:
        if abs(f32(pixel_depth) - f32(last)) <= 4.0 * depth_quantization(float(pixel_depth)):
            return True, pixel_depth
        return False, last
    return True, last


def init_plane_segment(seg, cx_, cy_, cz_, cell):
    """Plane_Segment::init_plane_segment, plane_segment.cpp:102-168. cx_/cy_/cz_: the cell's cloud rows (float32, raster
    order inside the cell), i.e. depthCloudArray.block(offset, k, P, 1)."""
    seg.clear()
    P = cell * cell
    # horizontal scan through the middle row (:83-100)
    start = int(cell * (cell / 2.0))
    last = max(cz_[start], cz_[start + 1])
    if last <= 0:
        return
    for i in range(start + 1, start + cell):
        ok, last = is_continuous(cz_[i], last)
        if not ok:
            return
    # vertical scan through the middle column (:63-81)
    start = cell // 2
    end = P - start
    last = max(cz_[start], cz_[start + cell])
    if last <= 0:
        return
    for i in range(start + cell, end, cell):
        ok, last = is_continuous(cz_[i], last)
        if not ok:
            return
    valid = cz_ > 0
    if int(valid.sum()) < P // 2:
        return
    x, y, z = cx_[valid], cy_[valid], cz_[valid]          # float32, raster order
    seg.count = int(valid.sum())

    def acc(values32):                                     # `double += float`, one pixel after the other
        return float(np.cumsum(values32.astype(np.float64))[-1]) if len(values32) else 0.0
    seg.S = np.array([acc(x), acc(y), acc(z), acc(x * x), acc(y * y), acc(z * z), acc(x * y), acc(y * z), acc(x * z)])
    min_zero_point_count = int(math.floor(float(f32(P) * MIN_ZERO_DEPTH_PROPORTION)))     # plane_segment.hpp:33-34
    if seg.count < min_zero_point_count:
        return
    seg.fit_plane()
    q = depth_quantization(float(seg.centroid[2]))
    seg.planar = bool(seg.mse <= q * q)


class Histogram:
    """Histogram<Size>, histogram.hpp:21-128 (Size = depthMapPatchSize_px, primitive_detection.hpp:199)."""

    def __init__(self, size):
        self.size = size
        self.hist = [0] * (size * size)
        self.bins = []

    def init(self, points, unassigned):
        self.bins = [-1] * len(points)
        for i, (theta, phi) in enumerate(points):
            if unassigned[i]:
                xq = int(math.floor((self.size - 1) * (theta - 0.0) / (math.pi - 0.0)))
                yq = 0
                if xq > 0:
                    yq = int(math.floor((self.size - 1) * (phi - (-math.pi)) / (math.pi - (-math.pi))))
                b = yq * self.size + xq
                self.bins[i] = b
                self.hist[b] += 1

    def most_frequent(self):
        best, occ = -1, 0
        for i, h in enumerate(self.hist):
            if h > occ:
                best, occ = i, h
        if best < 0:
            return []
        return [i for i, b in enumerate(self.bins) if b == best]

    def remove_point(self, i):
        if self.hist[self.bins[i]] != 0:
            self.hist[self.bins[i]] -= 1
        self.bins[i] = 1                 # sic (histogram.hpp:112)


class CylinderSegment:
    """Cylinder_Segment(planeGrid, isActivatedMask, cellActivatedCount), cylinder_segment.cpp:35-222."""

    def __init__(self, grid, activated, count, rng):
        self.n_cells = count
        self.local2global = [i for i in range(len(grid)) if activated[i]]
        self.axis = np.zeros(3)
        self.radius, self.centers, self.inliers, self.mse = [], [], [], []
        self.pca_score = 0.0
        m = count
        N = np.array([grid[i].normal for i in self.local2global]).T          # 3 x m
        C = np.array([grid[i].centroid for i in self.local2global]).T
        NN = np.concatenate([N, -N], axis=1)                                   # [Normals -Normals]
        cov = (NN @ NN.T) / float(NN.shape[1] - 1)
        w, v = np.linalg.eigh(cov)
        self.pca_score = float(w[2] / w[0]) if w[0] != 0 else math.inf
        if self.pca_score < float(CYL_MIN_SCORE):
            return
        axis = v[:, 0].copy()
        self.axis = axis
        cdt = axis @ C
        PC = C - np.outer(axis, cdt)
        ndt = axis @ N
        PN = N - np.outer(axis, ndt)
        PN = PN / np.linalg.norm(PN, axis=0)
        max_iterations = int(f32(math.log(f32(1.0) - CYL_PROBABILITY_OF_SUCCESS)) / f32(math.log(f32(1.0) - CYL_INLIER_PROPORTIONS ** f32(3.0))))
        assert max_iterations == 43, max_iterations
        left = m
        ids_mask = [True] * m
        ids_left = list(range(m))
        minimum_cell_activated = int(MIN_CELL_ACTIVATED_PROPORTION * float(len(grid)))
        while left > minimum_cell_activated and left > 0.1 * m:
            final = self._ransac(max_iterations, ids_left, PN, PC, ids_mask, rng)
            k = len(final)
            if k < 6:
                break
            is_inlier = [False] * m
            for i in final:
                is_inlier[i] = True
            b = 0.0
            sn, sc = np.zeros(3), np.zeros(3)
            ids_left = []
            for i in range(m):
                if is_inlier[i]:
                    ids_mask[i] = False
                    left -= 1
                    sn = sn + PN[:, i]
                    sc = sc + PC[:, i]
                    b += float((PN[:, i] * PC[:, i]).sum())
                elif ids_mask[i]:
                    ids_left.append(i)
            one_over_k2 = 1.0 / float(k * k)
            a = 1 - float(sn @ sn) * one_over_k2
            b /= float(k)
            b -= float(sn @ sc) * one_over_k2
            radius = b / a
            center = (sc - radius * sn) / k
            if radius < 0:
                radius = -radius
            self.radius.append(radius)
            self.centers.append(center)
            self.inliers.append(is_inlier)
            P1, P2 = center, center + axis
            P1P2 = float(np.linalg.norm(P2 - P1))
            mse = 0.0
            for i in range(m):
                if is_inlier[i]:
                    P3 = C[:, i]
                    dist = float(np.linalg.norm(np.cross(P2 - P1, P3 - P2))) / P1P2 - radius
                    mse += dist * dist
            self.mse.append(mse / float(k))

    def _ransac(self, max_iterations, ids_left, PN, PC, ids_mask, rng):
        """run_ransac_loop, cylinder_segment.cpp:227-322. Returns the final inlier indexes."""
        if len(ids_left) < 3:
            return []
        n_left = len(ids_left)
        accepted = int(math.floor(0.9 * n_left))
        max_d = float(CYL_SQRT_MAX_DISTANCE)
        min_dist = float(CYL_SQRT_MAX_DISTANCE * f32(n_left))
        final = []
        m = PN.shape[1]
        for _ in range(max_iterations):
            i1 = ids_left[rng.random_uint(n_left)]
            i2 = ids_left[rng.random_uint(n_left)]
            i3 = ids_left[rng.random_uint(n_left)]
            n1, n2, n3 = PN[:, i1], PN[:, i2], PN[:, i3]
            c1, c2, c3 = PC[:, i1], PC[:, i2], PC[:, i3]
            sn = (n1 + n2) + n3
            sc = (c1 + c2) + c3
            a = 1.0 - float(sn @ sn) / 9.0
            b = float(((n1 * c1) + (n2 * c2) + (n3 * c3)).sum()) / 3.0 - (float(sn @ sc) / 9.0)
            with np.errstate(divide="ignore", invalid="ignore"):
                radius = np.float64(b) / np.float64(a)
                inv_r2 = np.float64(1.0) / (radius * radius)
                center = (sc - radius * sn) / 3.0
                inl = []
                dist = 0.0
                for i in range(m):
                    if not ids_mask[i]:
                        continue
                    v = (PC[:, i] - radius * PN[:, i]) - center
                    distance = float(v @ v) * inv_r2
                    if distance < max_d:
                        dist += float(distance)
                        inl.append(i)
                    else:
                        dist += max_d
            if dist < min_dist:
                min_dist = dist
                final, inl = inl, final              # swap: `inl` now holds the PREVIOUS best set
                if len(inl) > accepted:              # the early stop tests that previous set (:304-313)
                    break
        return final


def find_primitives(depth, cell=20, K=(550.0, 550.0, 320.0, 240.0), seed=0):
    """Depth_Map_Transformation::get_organized_cloud_array + Primitive_Detection::find_primitives
    (depth_map_transformation.cpp:89-142, primitive_detection.cpp:119-166) on one frame. Returns a dict of outputs laid
    out like the C-ABI's (tests compare it field by field with the oracle)."""
    sys.setrecursionlimit(max(sys.getrecursionlimit(), 100000))
    depth = np.asarray(depth, dtype=np.float32)
    H, W = depth.shape
    hc, vc = W // cell, H // cell
    Nc = hc * vc
    P = cell * cell
    camx, camy = screen_to_camera_factors(K, W, H)
    zd = depth.astype(np.float64)
    valid = depth > 0
    X = np.where(valid, (zd * camx[None, :]).astype(np.float32), f32(0))     # cloud rows are zero for invalid pixels
    Y = np.where(valid, (zd * camy[:, None]).astype(np.float32), f32(0))
    Z = np.where(valid, depth, f32(0))

    # ---- init_planar_cell_fitting (primitive_detection.cpp:187-237) ----
    sin_merge = f32(math.sin(f32(float(MAX_PLANE_ANGLE_FOR_MERGE_D) * math.pi / 180.0)))
    sin_merge = np.sin(f32(float(MAX_PLANE_ANGLE_FOR_MERGE_D) * math.pi / 180.0), dtype=np.float32)   # sinf
    grid, tols = [], []
    for cid in range(Nc):
        r, c = divmod(cid, hc)
        sl = (slice(r * cell, (r + 1) * cell), slice(c * cell, (c + 1) * cell))
        cx_, cy_, cz_ = X[sl].reshape(-1), Y[sl].reshape(-1), Z[sl].reshape(-1)
        seg = PlaneSegment()
        init_plane_segment(seg, cx_, cy_, cz_, cell)
        grid.append(seg)
        if seg.planar:
            dx, dy, dz = cx_[P - 1] - cx_[0], cy_[P - 1] - cy_[0], cz_[P - 1] - cz_[0]
            diameter = np.sqrt((dx * dx + dy * dy) + dz * dz, dtype=np.float32)
            tols.append(min(MAX_PLANE_DISTANCE_FOR_MERGE_MM, diameter * sin_merge * np.sqrt(f32(seg.count), dtype=np.float32)))
        else:
            tols.append(f32(0))

    # ---- init_histogram (:239-265) ----
    unassigned = [False] * Nc
    pts = [(0.0, 0.0)] * Nc
    remaining = 0
    for cid, seg in enumerate(grid):
        if seg.planar:
            pts[cid] = (math.acos(-seg.normal[2]), math.atan2(seg.normal[0], seg.normal[1]))
            remaining += 1
            unassigned[cid] = True
    histogram = Histogram(cell)
    histogram.init(pts, unassigned)

    grid_plane = np.zeros((vc, hc), dtype=np.int32)
    grid_cyl = np.zeros((vc, hc), dtype=np.int32)
    plane_segments, cylinder_segments, cylinder2region = [], [], []
    cyl_assigned, cyl_plane_mse = [], []
    rng = StdMt19937Uniform(seed)       # thread_local engine of the per-frame std::async thread: restarts at the seed
    n_seeds = 0

    def region_growing(x, y, plane_to_expand, activated):
        """:778-818, literally (recursive, left / right / up / down)."""
        index = x + hc * y
        if index >= Nc:
            return
        if (not unassigned[index]) or activated[index]:
            return
        patch = grid[index]
        if plane_to_expand.can_be_merged(patch, float(tols[index])):
            activated[index] = True
            if x > 0:
                region_growing(x - 1, y, patch, activated)
            if x < hc - 1:
                region_growing(x + 1, y, patch, activated)
            if y > 0:
                region_growing(x, y - 1, patch, activated)
            if y < vc - 1:
                region_growing(x, y + 1, patch, activated)

    # ---- grow_planes_and_cylinders (:267-311) ----
    untried = remaining
    while untried > 0:
        cands = histogram.most_frequent()
        if len(cands) < int(MIN_PLANE_SEED_PROPORTION * Nc):
            break
        seed_id, min_mse = 0, DBL_MAX
        for cnd in cands:
            if grid[cnd].mse >= min_mse:
                continue
            seed_id, min_mse = cnd, grid[cnd].mse
            if min_mse <= 0:
                break
        if min_mse >= DBL_MAX:
            break
        n_seeds += 1
        # ---- grow_plane_segment_at_seed (:313-389) ----
        plane_to_grow = grid[seed_id]
        if not plane_to_grow.planar:
            continue
        new_seg = plane_to_grow.copy()
        y0, x0 = divmod(seed_id, hc)
        activated = [False] * Nc
        region_growing(x0, y0, new_seg, activated)
        cnt, fitable = 0, False
        for i in range(Nc):
            if activated[i] and grid[i].planar:
                new_seg.expand(grid[i])
                cnt += 1
                histogram.remove_point(i)
                unassigned[i] = False
                untried -= 1
                fitable = True
        if (not fitable) or cnt < int(MIN_CELL_ACTIVATED_PROPORTION * Nc):
            histogram.remove_point(seed_id)
            continue
        new_seg.fit_plane()
        if not new_seg.planar:
            continue
        if new_seg.score > 100:
            plane_segments.append(new_seg)
            k = len(plane_segments)
            for i in range(Nc):
                if activated[i]:
                    grid_plane[i // hc, i % hc] = k
        elif cnt > 5:
            # ---- cylinder_fitting (:476-501) ----
            cyl = CylinderSegment(grid, activated, cnt, rng)
            cylinder_segments.append(cyl)
            region = len(cylinder_segments) - 1
            cyl_assigned.append([0] * len(cyl.radius))
            cyl_plane_mse.append([DBL_MAX] * len(cyl.radius))
            for s in range(len(cyl.radius)):
                merged = PlaneSegment()
                fit = False
                for col in range(cnt):
                    if cyl.inliers[s][col] and grid[cyl.local2global[col]].planar:
                        merged.expand(grid[cyl.local2global[col]])
                        fit = True
                if not fit:
                    continue
                merged.fit_plane()
                cyl_plane_mse[region][s] = merged.mse
                # add_cylinder_to_features (:438-474)
                if merged.mse < cyl.mse[s]:
                    plane_segments.append(merged)
                    k = len(plane_segments)
                    for col in range(cnt):
                        if cyl.inliers[s][col]:
                            g = cyl.local2global[col]
                            grid_plane[g // hc, g % hc] = k
                    cyl_assigned[region][s] = -k
                else:
                    cylinder2region.append((region, s))
                    k = len(cylinder2region)
                    for col in range(cnt):
                        if cyl.inliers[s][col]:
                            g = cyl.local2global[col]
                            grid_cyl[g // hc, g % hc] = k
                    cyl_assigned[region][s] = k

    # ---- merge_planes + get_connected_components_matrix (:503-560, 736-776) ----
    n_planes = len(plane_segments)
    conn = np.zeros((n_planes, n_planes), dtype=bool)
    for row in range(vc - 1):
        for col in range(hc - 1):
            pid = grid_plane[row, col]
            if pid <= 0:
                continue
            nxt, below = grid_plane[row, col + 1], grid_plane[row + 1, col]
            if nxt > 0 and pid != nxt:
                conn[pid - 1, nxt - 1] = conn[nxt - 1, pid - 1] = True
            if below > 0 and pid != below:
                conn[pid - 1, below - 1] = conn[below - 1, pid - 1] = True
    labels = list(range(n_planes))
    for row in range(n_planes):
        expanded = False
        pid = labels[row]
        target = plane_segments[pid]
        if not target.planar:
            continue
        for col in range(row + 1, n_planes):
            if not conn[row, col]:
                continue
            other = plane_segments[col]
            if not other.planar:
                continue
            if target.can_be_merged(other, float(MAX_PLANE_DISTANCE_FOR_MERGE_MM)):
                target.expand(other)
                labels[col] = pid
                expanded = True
            else:
                conn[row, col] = conn[col, row] = False
        if expanded:
            target.fit_plane()

    # ---- add_planes_to_primitives + compute_plane_segment_boundary (:562-703) ----
    cross = np.array([[0, 1, 0], [1, 1, 1], [0, 1, 0]], dtype=np.uint8)
    square = np.ones((3, 3), dtype=np.uint8)
    plane_labels = np.zeros((vc, hc), dtype=np.int32)
    boundary, plane_out = [], []
    side = int(np.sqrt(f32(P), dtype=np.float32))
    for k in range(n_planes):
        seg = plane_segments[k]
        rec = dict(merge_label=labels[k], planar=int(seg.planar), is_final=int(labels[k] == k and seg.planar), count=seg.count,
                   S=seg.S.copy(), centroid=seg.centroid.copy(), normal=seg.normal.copy(), d=seg.d, mse=seg.mse, score=seg.score,
                   n_boundary=0, boundary_offset=len(boundary))
        plane_out.append(rec)
        if not rec["is_final"]:
            continue
        mask = np.zeros((vc, hc), dtype=np.uint8)
        for j in range(k, n_planes):
            if labels[j] == labels[k]:
                mask[grid_plane == (j + 1)] = 1
        plane_labels[mask > 0] = k + 1
        max_boundary_distance = 3 * math.sqrt(seg.mse)
        eroded = cv2.erode(mask, cross, anchor=(-1, -1), iterations=1, borderType=cv2.BORDER_CONSTANT, borderValue=0)
        dilated = cv2.dilate(mask, square)
        result = cv2.subtract(dilated, eroded)          # cv::Mat - cv::Mat on uchar saturates
        for row in range(vc):
            for col in range(hc):
                if result[row, col] <= 0:
                    continue
                cxp, cyp = int(col * side + side // 2), int(row * side + side // 2)
                dpt = float(depth[cyp, cxp])
                if dpt > 0:
                    p = np.array([dpt * camx[cxp], dpt * camy[cyp], dpt])
                    if abs(float(seg.normal @ p) + seg.d) < max_boundary_distance:
                        boundary.append(p)
                        rec["n_boundary"] += 1

    # ---- add_cylinders_to_primitives (:705-734) ----
    cyl_kept = []
    for ci in range(len(cylinder2region)):
        mask = (grid_cyl == (ci + 1)).astype(np.uint8)
        mask = cv2.dilate(mask, cross)
        mask = cv2.erode(mask, cross)
        er = cv2.erode(mask, cross)
        mn, mx = float(er.min()), float(er.max())
        cyl_kept.append(not (mx <= 0 or mn >= mx))

    return dict(hc=hc, vc=vc, grid=grid, tols=np.array(tols, dtype=np.float32), hist_bins=list(histogram.bins),
                plane_grid=grid_plane.reshape(-1), plane_labels=plane_labels.reshape(-1), cyl_labels=grid_cyl.reshape(-1),
                planes=plane_out, boundary=np.array(boundary).reshape(-1, 3), cylinders=cylinder_segments,
                cyl_assigned=cyl_assigned, cyl_plane_mse=cyl_plane_mse, cyl_kept=cyl_kept, cylinder2region=cylinder2region,
                n_seeds=n_seeds, n_planar=sum(1 for s in grid if s.planar))
