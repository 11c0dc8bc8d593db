"""Deterministic synthetic scenes of the shapes BASELINE.json names.

The reference synthesises its test data with ``colmap::SynthesizeDataset``
(glomap/controllers/global_mapper_test.cc:58-64), which is not vendored; these
generators produce the same kind of world (cameras looking at a point cloud,
pixel observations with optional noise, a view graph with relative rotations)
directly in the flat SoA layout that crosses the C ABI (include/b200sfm.h).

Shapes:
  config 1: ring of 100 cameras, 500 relative poses        -> make_ring_relposes
  config 2: 1k cams / 200k points / 2M observations         -> make_scene(1000, 200000)
  config 4: 10k cams / 2M points / 20M observations         -> make_scene(10000, 2000000)
  config 5: 100k-camera lattice view graph, 5M edges        -> make_lattice_view_graph
"""
from __future__ import annotations

import dataclasses

import numpy as np

from . import geometry as geo

# COLMAP camera model ids (colmap/sensor/models.h; un-vendored, public enum).
SIMPLE_PINHOLE, PINHOLE, SIMPLE_RADIAL, RADIAL = 0, 1, 2, 3
INTR_STRIDE = 12  # doubles reserved per intrinsics block across the C ABI
MODEL_NUM_PARAMS = {SIMPLE_PINHOLE: 3, PINHOLE: 4, SIMPLE_RADIAL: 4, RADIAL: 5}


@dataclasses.dataclass
class Scene:
    """Flat BA/GP problem: CSR by point (track) over observations."""
    quat: np.ndarray          # [C,4] xyzw cam_from_world
    trans: np.ndarray         # [C,3]
    points: np.ndarray        # [P,3]
    pt_obs_begin: np.ndarray  # [P+1] int64
    obs_cam: np.ndarray       # [N] int32
    obs_xy: np.ndarray        # [N,2] pixels
    cam_intr: np.ndarray      # [C] int32 -> intrinsics block
    intr_model: np.ndarray    # [K] int32
    intr_params: np.ndarray   # [K, INTR_STRIDE]

    @property
    def C(self):
        return len(self.quat)

    @property
    def P(self):
        return len(self.points)

    @property
    def N(self):
        return len(self.obs_cam)

    def copy(self):
        return Scene(*[np.array(getattr(self, f.name), copy=True) for f in dataclasses.fields(self)])


def project(model: int, params: np.ndarray, Xc: np.ndarray) -> np.ndarray:
    """Pixel projection of camera-frame points for the supported COLMAP models
    (SIMPLE_PINHOLE f,cx,cy | PINHOLE fx,fy,cx,cy | SIMPLE_RADIAL f,cx,cy,k |
    RADIAL f,cx,cy,k1,k2)."""
    u = Xc[..., 0] / Xc[..., 2]
    v = Xc[..., 1] / Xc[..., 2]
    if model == SIMPLE_PINHOLE:
        f, cx, cy = params[:3]
        return np.stack([f * u + cx, f * v + cy], -1)
    if model == PINHOLE:
        fx, fy, cx, cy = params[:4]
        return np.stack([fx * u + cx, fy * v + cy], -1)
    r2 = u * u + v * v
    if model == SIMPLE_RADIAL:
        f, cx, cy, k = params[:4]
        d = 1 + k * r2
    elif model == RADIAL:
        f, cx, cy, k1, k2 = params[:5]
        d = 1 + k1 * r2 + k2 * r2 * r2
    else:
        raise ValueError(f"unsupported camera model {model}")
    return np.stack([f * u * d + cx, f * v * d + cy], -1)


def _look_at_rotations(centers: np.ndarray, rng, jitter_deg: float) -> np.ndarray:
    """cam_from_world rotations with the optical axis (+z) toward the origin,
    plus a random rotation of up to ``jitter_deg`` degrees."""
    z = -centers / np.linalg.norm(centers, axis=1, keepdims=True)
    up = np.tile(np.array([0.0, 1.0, 0.0]), (len(centers), 1))
    bad = np.abs((z * up).sum(1)) > 0.99
    up[bad] = np.array([1.0, 0.0, 0.0])
    x = np.cross(up, z)
    x /= np.linalg.norm(x, axis=1, keepdims=True)
    y = np.cross(z, x)
    R = np.stack([x, y, z], axis=1)  # rows = camera axes in world coords
    if jitter_deg > 0:
        w = rng.normal(size=(len(centers), 3))
        w /= np.linalg.norm(w, axis=1, keepdims=True)
        w *= np.radians(jitter_deg) * rng.uniform(0, 1, size=(len(centers), 1))
        R = geo.so3_exp(w) @ R
    return R


def make_cameras(C: int, seed: int = 1, jitter_deg: float = 10.0):
    """Camera poses only (identical on every rank)."""
    rng = np.random.default_rng([seed, 0])
    d = rng.normal(size=(C, 3))
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    centers = d * rng.uniform(8, 12, size=(C, 1))
    R = _look_at_rotations(centers, rng, jitter_deg)
    t = -np.einsum("nij,nj->ni", R, centers)
    return R, t


def make_intrinsics(C: int, model: int, focal: float, image_size: int, num_intrinsics: int):
    K = num_intrinsics
    intr_model = np.full(K, model, dtype=np.int32)
    intr_params = np.zeros((K, INTR_STRIDE))
    half = image_size / 2
    for k in range(K):
        # a few shared blocks: 2 % steps around `focal`; many (per-image) blocks: bounded +-10 % variation
        f = focal * (1 + 0.02 * (k - (K - 1) / 2)) if K <= 8 else focal * (1 + 0.1 * np.sin(1.7 * k))
        if model == SIMPLE_PINHOLE:
            intr_params[k, :3] = [f, half, half]
        elif model == PINHOLE:
            intr_params[k, :4] = [f, f * 1.01, half, half]
        elif model == SIMPLE_RADIAL:
            intr_params[k, :4] = [f, half, half, 0.02]
        elif model == RADIAL:
            intr_params[k, :5] = [f, half, half, 0.02, -0.005]
        else:
            raise ValueError(f"unsupported camera model {model}")
    cam_intr = (np.arange(C) % K).astype(np.int32)
    return cam_intr, intr_model, intr_params


def make_scene(C: int, P: int, mean_track_len: float = 10.0, seed: int = 1, pixel_sigma: float = 0.0,
               model: int = SIMPLE_PINHOLE, focal: float = 1000.0, image_size: int = 1000,
               num_intrinsics: int = 1, candidates_mult: int = 3, ragged: bool = True,
               jitter_deg: float = 10.0, chunk: int = 100_000, point_range: tuple[int, int] | None = None) -> Scene:
    """Cameras on a shell r in [8,12] looking at the origin (+-jitter), points
    uniform in a ball of radius 3, each point observed by the cameras (of a
    random candidate set) with the smallest off-axis angle.  Track lengths are
    3 + Poisson(mean-3) when ``ragged`` else constant.

    Points are generated in chunks of ``chunk`` with one RNG stream per chunk,
    so ``point_range=(a, b)`` (multiples of ``chunk``) yields exactly the
    points [a, b) of the full scene -- how the multi-GPU bench shards."""
    R, t = make_cameras(C, seed, jitter_deg)
    cam_intr, intr_model, intr_params = make_intrinsics(C, model, focal, image_size, num_intrinsics)
    K = num_intrinsics
    a, b = point_range if point_range is not None else (0, P)
    assert a % chunk == 0 and (b % chunk == 0 or b == P), "point_range must align with chunk"
    M = int(min(C, max(int(candidates_mult * mean_track_len), 8)))
    R32, t32 = R.astype(np.float32), t.astype(np.float32)
    tan_half = 0.48 * image_size / focal

    pts_chunks, obs_cam_chunks, obs_xy_chunks, len_chunks = [], [], [], []
    for s in range(a, b, chunk):
        e = min(P, s + chunk)
        n = e - s
        rng = np.random.default_rng([seed, 1 + s // chunk])
        pts = rng.normal(size=(n, 3))
        pts /= np.linalg.norm(pts, axis=1, keepdims=True)
        pts *= 3.0 * rng.uniform(0, 1, size=(n, 1)) ** (1 / 3)
        if ragged:
            lens = 3 + rng.poisson(max(mean_track_len - 3, 0), size=n)
        else:
            lens = np.full(n, int(round(mean_track_len)))
        lens = np.minimum(lens, M).astype(np.int64)
        if M >= C:
            cand = np.tile(np.arange(C), (n, 1))
        else:
            cand = rng.integers(0, C, size=(n, M))
            cand.sort(axis=1)
        dup = np.zeros_like(cand, dtype=bool)
        dup[:, 1:] = cand[:, 1:] == cand[:, :-1]
        # candidate scoring in float32 (only the ranking matters)
        Xc = np.einsum("nmij,nj->nmi", R32[cand], pts.astype(np.float32)) + t32[cand]
        z = Xc[..., 2]
        cosang = z / np.linalg.norm(Xc, axis=-1)
        with np.errstate(divide="ignore", invalid="ignore"):
            u = Xc[..., 0] / z
            v = Xc[..., 1] / z
        vis = (z > 0.1) & (np.abs(u) < tan_half) & (np.abs(v) < tan_half) & ~dup
        score = np.where(vis, -cosang, np.inf)
        order = np.argsort(score, axis=1, kind="stable")
        ln = np.minimum(lens, vis.sum(1))
        take = np.arange(M)[None, :] < ln[:, None]
        sel_cam = np.take_along_axis(cand, order, axis=1)[take]
        pidx = np.repeat(np.arange(n), ln)
        Xs = np.einsum("nij,nj->ni", R[sel_cam], pts[pidx]) + t[sel_cam]
        xy = np.empty((len(sel_cam), 2))
        ci = cam_intr[sel_cam]
        for k in range(K):
            mk = ci == k if K > 1 else slice(None)
            xy[mk] = project(int(intr_model[k]), intr_params[k], Xs[mk])
        if pixel_sigma > 0:
            xy += rng.normal(scale=pixel_sigma, size=xy.shape)
        pts_chunks.append(pts)
        obs_cam_chunks.append(sel_cam.astype(np.int32))
        obs_xy_chunks.append(xy)
        len_chunks.append(ln)
    points = np.concatenate(pts_chunks)
    obs_cam = np.concatenate(obs_cam_chunks)
    obs_xy = np.concatenate(obs_xy_chunks)
    final_lens = np.concatenate(len_chunks)
    pt_obs_begin = np.zeros(len(points) + 1, dtype=np.int64)
    np.cumsum(final_lens, out=pt_obs_begin[1:])
    quat = geo.rotmat_to_quat_xyzw_fast(R)
    return Scene(quat, t, points, pt_obs_begin, obs_cam, obs_xy, cam_intr, intr_model, intr_params)


def perturb_scene(scene: Scene, rot_deg: float = 0.5, center_frac: float = 0.01, point_frac: float = 0.01,
                  seed: int = 2, extent: float = 10.0, chunk: int = 100_000, point_offset: int = 0) -> Scene:
    """BA initial state: ground truth perturbed by ``rot_deg`` degrees,
    ``center_frac``*extent camera-centre noise, ``point_frac``*3 point noise
    (SURVEY.md 8(d) config 4).  Point noise uses one RNG stream per chunk of
    the global point index (``point_offset`` = first global index of a shard)."""
    rng = np.random.default_rng([seed, 0])
    out = scene.copy()
    R = geo.quat_xyzw_to_rotmat(scene.quat)
    c = geo.centers_from_pose(R, scene.trans)
    w = rng.normal(size=(scene.C, 3)) * np.radians(rot_deg) / np.sqrt(3)
    Rn = geo.so3_exp(w) @ R
    cn = c + rng.normal(size=c.shape) * center_frac * extent / np.sqrt(3)
    out.quat = geo.rotmat_to_quat_xyzw_fast(Rn)
    out.trans = -np.einsum("nij,nj->ni", Rn, cn)
    assert point_offset % chunk == 0
    for s in range(0, scene.P, chunk):
        e = min(scene.P, s + chunk)
        prng = np.random.default_rng([seed, 1 + (point_offset + s) // chunk])
        out.points[s:e] = scene.points[s:e] + prng.normal(size=(e - s, 3)) * point_frac * 3.0 / np.sqrt(3)
    return out


@dataclasses.dataclass
class RigScene:
    """Flat BA/GP problem with KNOWN camera rigs (b200sfm_ba_problem_create_rig): the pose unknowns are
    the F frames (rig_from_world); image (f, s) = frame f seen through sensor s, whose cam_from_rig and
    intrinsics block are constants of the sensor (glomap/scene/frame.h, colmap::Rig)."""
    quat: np.ndarray          # [F,4] xyzw rig_from_world
    trans: np.ndarray         # [F,3]
    points: np.ndarray        # [P,3]
    pt_obs_begin: np.ndarray  # [P+1] int64
    obs_frame: np.ndarray     # [N] int32
    obs_sensor: np.ndarray    # [N] uint16
    obs_xy: np.ndarray        # [N,2] pixels
    sensor_quat: np.ndarray   # [S,4] xyzw cam_from_rig
    sensor_trans: np.ndarray  # [S,3]
    sensor_intr: np.ndarray   # [S] int32 -> intrinsics block
    intr_model: np.ndarray    # [K] int32
    intr_params: np.ndarray   # [K, INTR_STRIDE]

    @property
    def C(self):
        return len(self.quat)

    @property
    def F(self):
        return len(self.quat)

    @property
    def S(self):
        return len(self.sensor_quat)

    @property
    def P(self):
        return len(self.points)

    @property
    def N(self):
        return len(self.obs_frame)

    def copy(self):
        return RigScene(*[np.array(getattr(self, f.name), copy=True) for f in dataclasses.fields(self)])

    def image_poses(self):
        """cam_from_world of all F*S images, image id = f * S + s."""
        Rf = geo.quat_xyzw_to_rotmat(self.quat)
        Rs = geo.quat_xyzw_to_rotmat(self.sensor_quat)
        R = np.einsum("sij,fjk->fsik", Rs, Rf).reshape(-1, 3, 3)
        t = (np.einsum("sij,fj->fsi", Rs, self.trans) + self.sensor_trans[None]).reshape(-1, 3)
        return R, t

    def images_scene(self) -> Scene:
        """The same observations as a trivial-frame Scene over the F*S images (poses composed)."""
        R, t = self.image_poses()
        obs_cam = (self.obs_frame.astype(np.int64) * self.S + self.obs_sensor).astype(np.int32)
        cam_intr = np.tile(self.sensor_intr, self.F).astype(np.int32)
        return Scene(geo.rotmat_to_quat_xyzw_fast(R), t, self.points.copy(), self.pt_obs_begin.copy(), obs_cam,
                     self.obs_xy.copy(), cam_intr, self.intr_model.copy(), self.intr_params.copy())

    def rig_dict(self):
        """The ``rig`` argument of oracle.ba_oracle (per-image arrays, image id = f * S + s)."""
        # img_sensor / sensor_q / sensor_t are read only with optimize_rig_poses (sensor 0 = reference sensor: constant)
        img_sensor = np.tile(np.where(np.arange(self.S) == 0, -1, np.arange(self.S)), self.F)
        return dict(obs_img=self.obs_frame.astype(np.int64) * self.S + self.obs_sensor,
                    img_q=np.tile(self.sensor_quat, (self.F, 1)), img_t=np.tile(self.sensor_trans, (self.F, 1)),
                    img_intr=np.tile(self.sensor_intr, self.F), img_sensor=img_sensor,
                    sensor_q=self.sensor_quat.copy(), sensor_t=self.sensor_trans.copy())


def make_rig_scene(F: int, S: int, P: int, mean_track_len: float = 8.0, seed: int = 1, pixel_sigma: float = 0.0,
                   model: int = SIMPLE_PINHOLE, focal: float = 1000.0, image_size: int = 1000,
                   shared_intrinsics: bool = False, sensor_rot_deg: float = 20.0, sensor_offset: float = 0.4) -> RigScene:
    """F rigs of S cameras looking at a ball of points.  Sensor 0 is the reference sensor (identity
    cam_from_rig); the others are rotated by up to ``sensor_rot_deg`` and shifted by ``sensor_offset``.
    Every sensor has its own intrinsics block unless ``shared_intrinsics``.  Small sizes only (tests)."""
    rng = np.random.default_rng([seed, 77])
    Rf, tf = make_cameras(F, seed=seed, jitter_deg=5.0)
    w = rng.normal(size=(S, 3))
    w = w / np.linalg.norm(w, axis=1, keepdims=True) * np.radians(sensor_rot_deg) * rng.uniform(0.3, 1.0, size=(S, 1))
    w[0] = 0.0
    Rs = geo.so3_exp(w)
    ts = rng.normal(size=(S, 3))
    ts = ts / np.linalg.norm(ts, axis=1, keepdims=True) * sensor_offset
    ts[0] = 0.0
    K = 1 if shared_intrinsics else S
    _, intr_model, intr_params = make_intrinsics(S, model, focal, image_size, K)
    sensor_intr = (np.arange(S) % K).astype(np.int32)
    scene = RigScene(geo.rotmat_to_quat_xyzw_fast(Rf), tf, np.zeros((P, 3)), np.zeros(P + 1, np.int64),
                     np.zeros(0, np.int32), np.zeros(0, np.uint16), np.zeros((0, 2)), geo.rotmat_to_quat_xyzw_fast(Rs), ts,
                     sensor_intr, intr_model, intr_params)
    Ri, ti = scene.image_poses()
    d = rng.normal(size=(P, 3))
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    points = d * 3.0 * rng.uniform(0, 1, size=(P, 1)) ** (1 / 3)
    lens = 3 + rng.poisson(max(mean_track_len - 3, 0.0), size=P)
    fr, se, xy, begin = [], [], [], [0]
    lim = 0.5 * image_size / focal * 1.2
    for p in range(P):
        Xc = np.einsum("nij,j->ni", Ri, points[p]) + ti
        u, v = Xc[:, 0] / Xc[:, 2], Xc[:, 1] / Xc[:, 2]
        ok = np.nonzero((Xc[:, 2] > 0.5) & (np.abs(u) < lim) & (np.abs(v) < lim))[0]
        pick = rng.permutation(ok)[: lens[p]]
        pick.sort()
        for img in pick:
            f, sidx = divmod(int(img), S)
            k = sensor_intr[sidx]
            px = project(int(intr_model[k]), intr_params[k], Xc[img])
            fr.append(f); se.append(sidx); xy.append(px)
        begin.append(len(fr))
    scene.points = points
    scene.pt_obs_begin = np.asarray(begin, np.int64)
    scene.obs_frame = np.asarray(fr, np.int32)
    scene.obs_sensor = np.asarray(se, np.uint16)
    scene.obs_xy = np.asarray(xy, np.float64).reshape(-1, 2)
    if pixel_sigma > 0:
        scene.obs_xy = scene.obs_xy + rng.normal(size=scene.obs_xy.shape) * pixel_sigma
    return scene


def perturb_rig_scene(scene: RigScene, rot_deg: float = 0.5, center_frac: float = 0.01, point_frac: float = 0.01,
                      seed: int = 2, extent: float = 10.0) -> RigScene:
    """BA initial state for a rig scene: frame poses and points perturbed like ``perturb_scene``."""
    rng = np.random.default_rng([seed, 5])
    out = scene.copy()
    R = geo.quat_xyzw_to_rotmat(scene.quat)
    c = geo.centers_from_pose(R, scene.trans)
    w = rng.normal(size=(scene.F, 3)) * np.radians(rot_deg) / np.sqrt(3)
    Rn = geo.so3_exp(w) @ R
    cn = c + rng.normal(size=c.shape) * center_frac * extent / np.sqrt(3)
    out.quat = geo.rotmat_to_quat_xyzw_fast(Rn)
    out.trans = -np.einsum("nij,nj->ni", Rn, cn)
    out.points = scene.points + rng.normal(size=scene.points.shape) * point_frac * 3.0 / np.sqrt(3)
    return out


def bearings_from_scene(scene: Scene) -> np.ndarray:
    """Unit bearing of each observation in the camera frame -- what the
    reference keeps in ``Image::features_undist`` (glomap/scene/image.h:31,
    processors/image_undistorter.cc).  Exact inverse for the pinhole models;
    radial models are inverted by fixed-point iteration."""
    out = np.empty((scene.N, 3))
    ci = scene.cam_intr[scene.obs_cam]
    for k in range(len(scene.intr_model)):
        mk = ci == k
        if not mk.any():
            continue
        m = int(scene.intr_model[k])
        p = scene.intr_params[k]
        xy = scene.obs_xy[mk]
        if m == SIMPLE_PINHOLE:
            u, v = (xy[:, 0] - p[1]) / p[0], (xy[:, 1] - p[2]) / p[0]
        elif m == PINHOLE:
            u, v = (xy[:, 0] - p[2]) / p[0], (xy[:, 1] - p[3]) / p[1]
        else:
            ud, vd = (xy[:, 0] - p[1]) / p[0], (xy[:, 1] - p[2]) / p[0]
            u, v = ud.copy(), vd.copy()
            for _ in range(50):
                r2 = u * u + v * v
                dd = 1 + p[3] * r2 + (p[4] * r2 * r2 if m == RADIAL else 0.0)
                u, v = ud / dd, vd / dd
        b = np.stack([u, v, np.ones_like(u)], 1)
        out[mk] = b / np.linalg.norm(b, axis=1, keepdims=True)
    return out


# ---------------------------------------------------------------------------
# View graphs for rotation averaging
# ---------------------------------------------------------------------------
@dataclasses.dataclass
class ViewGraph:
    """Flat view graph: edge e relates images (i, j) with R_rel = R_j R_i^T
    (``cam2_from_cam1`` of the reference, glomap/scene/image_pair.h)."""
    n_images: int
    ei: np.ndarray      # [E] int32 image index 1
    ej: np.ndarray      # [E] int32 image index 2
    R_rel: np.ndarray   # [E,3,3]
    weight: np.ndarray  # [E]
    R_gt: np.ndarray    # [n,3,3] ground-truth cam_from_world rotations

    @property
    def E(self):
        return len(self.ei)


def _noisy_relative(R_gt, ei, ej, rng, noise_deg, outlier_ratio):
    R_rel = R_gt[ej] @ np.swapaxes(R_gt[ei], -1, -2)
    E = len(ei)
    if noise_deg > 0:
        w = rng.normal(size=(E, 3)) * np.radians(noise_deg) / np.sqrt(3)
        R_rel = geo.so3_exp(w) @ R_rel
    if outlier_ratio > 0:
        out = rng.uniform(size=E) < outlier_ratio
        w = rng.normal(size=(int(out.sum()), 3))
        w /= np.linalg.norm(w, axis=1, keepdims=True)
        w *= rng.uniform(0, np.pi, size=(len(w), 1))
        R_rel[out] = geo.so3_exp(w)
    return R_rel


def make_ring_view_graph(n: int = 100, k: int = 5, seed: int = 1, noise_deg: float = 0.0,
                         outlier_ratio: float = 0.0) -> ViewGraph:
    """Config 1: n cameras on a ring of radius 10 looking at the centre,
    edges (i, i+d mod n) for d = 1..k."""
    rng = np.random.default_rng(seed)
    ang = 2 * np.pi * np.arange(n) / n
    centers = np.stack([10 * np.sin(ang), np.zeros(n), 10 * np.cos(ang)], 1)
    R_gt = _look_at_rotations(centers, rng, 0.0)
    ei = np.repeat(np.arange(n), k)
    ej = (ei + np.tile(np.arange(1, k + 1), n)) % n
    R_rel = _noisy_relative(R_gt, ei, ej, rng, noise_deg, outlier_ratio)
    return ViewGraph(n, ei.astype(np.int32), ej.astype(np.int32), R_rel, np.ones(len(ei)), R_gt)


def make_random_view_graph(n: int, avg_degree: float, seed: int = 1, noise_deg: float = 0.0,
                           outlier_ratio: float = 0.0) -> ViewGraph:
    """Random rotations, a spanning path for connectivity plus random edges."""
    rng = np.random.default_rng(seed)
    w = rng.normal(size=(n, 3))
    w /= np.linalg.norm(w, axis=1, keepdims=True)
    w *= rng.uniform(0, np.pi, size=(n, 1))
    R_gt = geo.so3_exp(w)
    perm = rng.permutation(n)
    pairs = {(min(a, b), max(a, b)) for a, b in zip(perm[:-1], perm[1:])}
    target = int(n * avg_degree / 2)
    while len(pairs) < target:
        a = rng.integers(0, n, size=target)
        b = rng.integers(0, n, size=target)
        for x, y in zip(a, b):
            if x != y:
                pairs.add((min(x, y), max(x, y)))
            if len(pairs) >= target:
                break
    pr = np.array(sorted(pairs), dtype=np.int64)
    ei, ej = pr[:, 0], pr[:, 1]
    R_rel = _noisy_relative(R_gt, ei, ej, rng, noise_deg, outlier_ratio)
    return ViewGraph(n, ei.astype(np.int32), ej.astype(np.int32), R_rel, np.ones(len(ei)), R_gt)


def make_lattice_view_graph(n: int = 100_000, neighbours: int = 50, seed: int = 1, noise_deg: float = 2.0,
                            outlier_ratio: float = 0.05) -> ViewGraph:
    """Config 5: cameras on a 2-D lattice, each linked to its ``neighbours``
    nearest lattice neighbours (half of them stored, i<j)."""
    rng = np.random.default_rng(seed)
    side = int(np.ceil(np.sqrt(n)))
    gx, gy = np.divmod(np.arange(n), side)
    w = rng.normal(size=(n, 3)) * 0.5
    R_gt = geo.so3_exp(w)
    rad = 1
    while (2 * rad + 1) ** 2 - 1 < neighbours:
        rad += 1
    offs = [(dx, dy) for dx in range(-rad, rad + 1) for dy in range(-rad, rad + 1) if (dx, dy) > (0, 0)]
    offs.sort(key=lambda o: o[0] * o[0] + o[1] * o[1])
    offs = offs[: neighbours // 2]
    ei_l, ej_l = [], []
    for dx, dy in offs:
        nx, ny = gx + dx, gy + dy
        j = nx * side + ny
        ok = (nx >= 0) & (nx < side) & (ny >= 0) & (ny < side) & (j < n)
        ei_l.append(np.arange(n)[ok])
        ej_l.append(j[ok])
    ei = np.concatenate(ei_l)
    ej = np.concatenate(ej_l)
    R_rel = _noisy_relative(R_gt, ei, ej, rng, noise_deg, outlier_ratio)
    return ViewGraph(n, ei.astype(np.int32), ej.astype(np.int32), R_rel, np.ones(len(ei)), R_gt)


def view_graph_from_scene(scene: Scene, min_shared: int = 30, seed: int = 3, noise_deg: float = 0.0,
                          outlier_ratio: float = 0.0) -> ViewGraph:
    """Camera pairs sharing >= ``min_shared`` points, with relative rotations
    from the scene's rotations (config 2's view graph)."""
    import scipy.sparse as sp
    rng = np.random.default_rng(seed)
    pt_of_obs = np.repeat(np.arange(scene.P), np.diff(scene.pt_obs_begin))
    V = sp.csr_matrix((np.ones(scene.N, dtype=np.int32), (scene.obs_cam, pt_of_obs)), shape=(scene.C, scene.P))
    cov = sp.triu(V @ V.T, k=1).tocoo()
    keep = cov.data >= min_shared
    ei, ej = cov.row[keep], cov.col[keep]
    R_gt = geo.quat_xyzw_to_rotmat(scene.quat)
    R_rel = _noisy_relative(R_gt, ei, ej, rng, noise_deg, outlier_ratio)
    return ViewGraph(scene.C, ei.astype(np.int32), ej.astype(np.int32), R_rel, cov.data[keep].astype(np.float64), R_gt)


# ---------------------------------------------------------------------------
# Text formats of `glomap rotation_averager` (docs/rotation_averager.md:43-69)
# ---------------------------------------------------------------------------
def write_relpose_file(path: str, vg: ViewGraph, names=None) -> None:
    """IMAGE_NAME_1 IMAGE_NAME_2 QW QX QY QZ TX TY TZ (glomap/io/pose_io.cc:36-75)."""
    names = names or [f"img{i:04d}" for i in range(vg.n_images)]
    q = geo.rotmat_to_quat_xyzw_fast(vg.R_rel)
    with open(path, "w") as f:
        for e in range(vg.E):
            f.write(f"{names[vg.ei[e]]} {names[vg.ej[e]]} {q[e,3]:.17g} {q[e,0]:.17g} {q[e,1]:.17g} {q[e,2]:.17g} 1 0 0\n")


def read_relpose_file(path: str) -> tuple[ViewGraph, list[str]]:
    """Parser with the reference's id assignment: images numbered in order of
    first appearance (glomap/io/pose_io.cc:46-59)."""
    names, idx, ei, ej, qs = [], {}, [], [], []
    with open(path) as f:
        for line in f:
            tok = line.rstrip("\n").split(" ")
            if len(tok) < 9:
                continue
            for nm in tok[:2]:
                if nm not in idx:
                    idx[nm] = len(names)
                    names.append(nm)
            ei.append(idx[tok[0]])
            ej.append(idx[tok[1]])
            qw, qx, qy, qz = (float(x) for x in tok[2:6])
            qs.append([qx, qy, qz, qw])
    R_rel = geo.quat_xyzw_to_rotmat(np.array(qs))
    n = len(names)
    vg = ViewGraph(n, np.array(ei, np.int32), np.array(ej, np.int32), R_rel, np.ones(len(ei)), np.tile(np.eye(3), (n, 1, 1)))
    return vg, names


def write_global_rotation_file(path: str, names, R: np.ndarray) -> None:
    """IMAGE_NAME QW QX QY QZ, default ostream precision (pose_io.cc:182-200)."""
    q = geo.rotmat_to_quat_xyzw_fast(R)
    with open(path, "w") as f:
        for i, nm in enumerate(names):
            f.write(f"{nm} {q[i,3]:.6g} {q[i,0]:.6g} {q[i,1]:.6g} {q[i,2]:.6g}\n")


# ---------------------------------------------------------------------------
# Flat binary problem file (glomap_b200/host/b200sfm_cli.cc)
# ---------------------------------------------------------------------------
def write_flat_problem(path: str, scene: Scene, bearings: np.ndarray | None = None) -> None:
    """int64 {C,P,N,K}; int64 pt_obs_begin[P+1]; int32 obs_cam[N]; f64 obs_xy[2N]; f64 bearings[3N];
    int32 cam_intr[C]; int32 intr_model[K]; f64 intr[K*12]; f64 quat[4C]; f64 trans[3C]; f64 points[3P]."""
    b = bearings_from_scene(scene) if bearings is None else bearings
    with open(path, "wb") as f:
        np.array([scene.C, scene.P, scene.N, len(scene.intr_model)], np.int64).tofile(f)
        for arr, dt in ((scene.pt_obs_begin, np.int64), (scene.obs_cam, np.int32), (scene.obs_xy, np.float64),
                        (b, np.float64), (scene.cam_intr, np.int32), (scene.intr_model, np.int32),
                        (scene.intr_params, np.float64), (scene.quat, np.float64), (scene.trans, np.float64),
                        (scene.points, np.float64)):
            np.ascontiguousarray(arr, dtype=dt).tofile(f)


def read_flat_problem(path: str) -> Scene:
    with open(path, "rb") as f:
        C, P, N, K = (int(x) for x in np.fromfile(f, np.int64, 4))
        ptb = np.fromfile(f, np.int64, P + 1)
        cam = np.fromfile(f, np.int32, N)
        xy = np.fromfile(f, np.float64, 2 * N).reshape(N, 2)
        np.fromfile(f, np.float64, 3 * N)
        ci = np.fromfile(f, np.int32, C)
        im = np.fromfile(f, np.int32, K)
        intr = np.fromfile(f, np.float64, K * INTR_STRIDE).reshape(K, INTR_STRIDE)
        quat = np.fromfile(f, np.float64, 4 * C).reshape(C, 4)
        trans = np.fromfile(f, np.float64, 3 * C).reshape(C, 3)
        pts = np.fromfile(f, np.float64, 3 * P).reshape(P, 3)
    return Scene(quat, trans, pts, ptb, cam, xy, ci, im, intr)
