"""Deterministic synthetic multi-view scenes for parity tests and benchmarks.

The reference ships no sample data and its input loader (InuputInitialization, reference
APD.cpp:1045-1495) needs OpenCV plus external mono-depth files, so the harness synthesises exactly what
that loader leaves in the APD object: float grey images (reference view + S source views, all W x H),
cameras in the reference's `Camera` layout (main.h:58-67), an initial (world normal, depth) plane map,
and the edge / label / radius priors of SupportInitialization (APD.cpp:1615-1668).

Scene: a closed piecewise-planar room (floor, ceiling, two side walls, back wall) with one
fronto-parallel box face in front of the back wall, ray-cast per view (so occlusions are real), textured
with band-limited noise; one side wall is nearly textureless to provoke WEAK pixels.
Everything is a pure function of (W, H, S, seed).
"""
from __future__ import annotations

import dataclasses
import os
from concurrent.futures import ThreadPoolExecutor

import numpy as np

CAMERA_DTYPE = np.dtype(
    [("K", "<f4", (9,)), ("R", "<f4", (9,)), ("t", "<f4", (3,)), ("c", "<f4", (3,)),
     ("height", "<i4"), ("width", "<i4"), ("depth_min", "<f4"), ("depth_max", "<f4")]
)
assert CAMERA_DTYPE.itemsize == 112

SEED_IMAGE, SEED_CAMERA, SEED_INIT, SEED_RNG = 20250101, 20250102, 20250103, 0x5EED


@dataclasses.dataclass
class Scene:
    width: int
    height: int
    num_src: int
    images: np.ndarray      # [(1+S), H, W] float32
    depths: np.ndarray      # [(1+S), H, W] float32 ground-truth z-depth per view
    cameras: np.ndarray     # [(1+S)] CAMERA_DTYPE
    planes_init: np.ndarray  # [H, W, 4] float32 (world normal, noisy depth) — FIRST_INIT input
    planes_true: np.ndarray  # [H, W, 4] float32 (world normal, true depth)
    plane_id: np.ndarray    # [H, W] int32
    edge: np.ndarray        # [H, W] uint8 0/255
    label: np.ndarray       # [H, W] int32
    depth_min: float        # PatchMatchParams.depth_min = 0.6 * cam.depth_min (APD.cpp:1109)
    depth_max: float        # PatchMatchParams.depth_max = 1.2 * cam.depth_max (APD.cpp:1110)


def _rot_yx(yaw: float, pitch: float) -> np.ndarray:
    cy, sy, cp, sp = np.cos(yaw), np.sin(yaw), np.cos(pitch), np.sin(pitch)
    ry = np.array([[cy, 0, sy], [0, 1, 0], [-sy, 0, cy]])
    rx = np.array([[1, 0, 0], [0, cp, -sp], [0, sp, cp]])
    return rx @ ry


def _blur_periodic(a: np.ndarray, sigma: float) -> np.ndarray:
    """Gaussian blur with wrap-around, via FFT (keeps the tile seamlessly periodic)."""
    n0, n1 = a.shape
    f0 = np.fft.fftfreq(n0)[:, None]
    f1 = np.fft.rfftfreq(n1)[None, :]
    g = np.exp(-2.0 * (np.pi ** 2) * (sigma ** 2) * (f0 ** 2 + f1 ** 2))
    return np.fft.irfft2(np.fft.rfft2(a) * g, s=a.shape)


def _make_texture(rng: np.random.Generator, size: int, sigma: float) -> np.ndarray:
    t = _blur_periodic(rng.uniform(0.0, 255.0, (size, size)), sigma)
    if sigma < 10:
        t = (t - t.mean()) / (t.std() + 1e-9) * 55.0 + 127.0  # restore contrast lost to the blur
    else:
        t = (t - t.mean()) / (t.std() + 1e-9) * 2.0 + 127.0  # nearly constant: +-2 grey levels
    return np.clip(t, 0.0, 255.0)


# Planes of the room in WORLD coordinates: (point, normal (towards the room interior), u axis, v axis, bounds or None)
def _room():
    P = []
    P.append(dict(p0=(0, 1.2, 0), n=(0, -1, 0), eu=(1, 0, 0), ev=(0, 0, 1), bounds=None))    # 0 floor
    P.append(dict(p0=(0, -1.5, 0), n=(0, 1, 0), eu=(1, 0, 0), ev=(0, 0, 1), bounds=None))    # 1 ceiling
    P.append(dict(p0=(-2.5, 0, 0), n=(1, 0, 0), eu=(0, 0, 1), ev=(0, 1, 0), bounds=None))    # 2 left wall
    P.append(dict(p0=(2.5, 0, 0), n=(-1, 0, 0), eu=(0, 0, 1), ev=(0, 1, 0), bounds=None))    # 3 right wall (textureless)
    P.append(dict(p0=(0, 0, 8.0), n=(0, 0, -1), eu=(1, 0, 0), ev=(0, 1, 0), bounds=None))    # 4 back wall
    P.append(dict(p0=(0, 0, 4.0), n=(0, 0, -1), eu=(1, 0, 0), ev=(0, 1, 0), bounds=(-0.8, 0.9, -0.4, 1.2)))  # 5 box face
    for d in P:
        for k in ("p0", "n", "eu", "ev"):
            d[k] = np.asarray(d[k], dtype=np.float64)
    return P


def _bilinear_periodic(tex: np.ndarray, u: np.ndarray, v: np.ndarray) -> np.ndarray:
    n = tex.shape[0]
    u0 = np.floor(u); v0 = np.floor(v)
    fu = u - u0; fv = v - v0
    iu = u0.astype(np.int64) % n; iv = v0.astype(np.int64) % n
    iu1 = (iu + 1) % n; iv1 = (iv + 1) % n
    return ((1 - fu) * (1 - fv) * tex[iv, iu] + fu * (1 - fv) * tex[iv, iu1]
            + (1 - fu) * fv * tex[iv1, iu] + fu * fv * tex[iv1, iu1])


def _render(K, R, c, W, H, planes, textures, texels_per_m, rows_per_chunk=64):
    """Ray-cast one view. Returns image, z-depth, plane id, camera-frame normal per pixel."""
    img = np.zeros((H, W), np.float32); dep = np.zeros((H, W), np.float32)
    pid = np.full((H, W), -1, np.int32)
    fx, fy, cx, cy = K[0, 0], K[1, 1], K[0, 2], K[1, 2]
    xs = (np.arange(W, dtype=np.float64) - cx) / fx
    Rt = R.T
    def chunk(y0):
        y1 = min(H, y0 + rows_per_chunk)
        ys = (np.arange(y0, y1, dtype=np.float64) - cy) / fy
        dx, dy = np.meshgrid(xs, ys)
        dcam = np.stack([dx, dy, np.ones_like(dx)], -1)     # z component 1 -> ray parameter == z-depth
        dw = dcam @ Rt.T                                      # world direction = R^T d
        best_t = np.full(dx.shape, np.inf)
        best_id = np.full(dx.shape, -1, np.int32)
        best_val = np.zeros(dx.shape)
        for i, pl in enumerate(planes):
            denom = dw @ pl["n"]
            with np.errstate(divide="ignore", invalid="ignore"):
                t = ((pl["p0"] - c) @ pl["n"]) / denom
            ok = (denom < 0) & (t > 1e-6) & (t < best_t)
            if not ok.any():
                continue
            Xw = c + dw * np.where(ok, t, 0.0)[..., None]
            u = (Xw - pl["p0"]) @ pl["eu"]; v = (Xw - pl["p0"]) @ pl["ev"]
            if pl["bounds"] is not None:
                b = pl["bounds"]
                ok &= (u >= b[0]) & (u <= b[1]) & (v >= b[2]) & (v <= b[3])
            val = _bilinear_periodic(textures[i], u * texels_per_m + 17.0 * i, v * texels_per_m + 31.0 * i)
            best_val = np.where(ok, val, best_val)
            best_id = np.where(ok, i, best_id)
            best_t = np.where(ok, t, best_t)
        img[y0:y1] = best_val.astype(np.float32)
        dep[y0:y1] = np.where(np.isfinite(best_t), best_t, 0.0).astype(np.float32)
        pid[y0:y1] = best_id

    # row chunks are independent (each writes its own rows): rendered on a thread pool, numpy releases the GIL
    starts = list(range(0, H, rows_per_chunk))
    workers = min(len(starts), os.cpu_count() or 1)
    if workers > 1:
        with ThreadPoolExecutor(workers) as ex:
            list(ex.map(chunk, starts))
    else:
        for y0 in starts:
            chunk(y0)
    return img, dep, pid


def make_camera(K, R, c, W, H, dmin, dmax) -> np.ndarray:
    cam = np.zeros((), CAMERA_DTYPE)
    cam["K"] = K.reshape(9).astype(np.float32)
    cam["R"] = R.reshape(9).astype(np.float32)
    cam["t"] = (-R @ c).astype(np.float32)
    cam["c"] = c.astype(np.float32)   # the reference computes the centre in double and stores f32 (APD.cpp:676)
    cam["height"] = H; cam["width"] = W
    cam["depth_min"] = dmin; cam["depth_max"] = dmax
    return cam


def _first_init_planes(depth, K, R_ref, rng_init):
    """FIRST_INIT input: noisy depth (10 % of pixels invalid = 0), normals by finite differences of that depth
    map, flipped towards the camera and rotated to the world frame (as APD.cpp:1365-1422 does)."""
    H, W = depth.shape
    noisy = depth.astype(np.float64) * (1.0 + rng_init.normal(0.0, 0.05, (H, W)))
    noisy[rng_init.uniform(size=(H, W)) < 0.10] = 0.0
    xs = (np.arange(W) - K[0, 2]) / K[0, 0]; ys = (np.arange(H) - K[1, 2]) / K[1, 1]
    X = np.stack([xs[None, :] * noisy, ys[:, None] * noisy, noisy], -1)
    dPdx = np.zeros_like(X); dPdy = np.zeros_like(X)
    dPdx[:, :-1] = X[:, 1:] - X[:, :-1]
    dPdy[:-1, :] = X[1:, :] - X[:-1, :]
    nrm = np.cross(dPdx, dPdy)
    ln = np.linalg.norm(nrm, axis=-1, keepdims=True)
    nrm = np.where(ln > 0, nrm / np.maximum(ln, 1e-30), np.array([0.0, 0.0, -1.0]))
    flip = (nrm * X).sum(-1) > 0
    nrm[flip] *= -1.0
    nrm_world = nrm @ R_ref  # R^T n, row-vector form
    nrm_world[0, :] = 0; nrm_world[-1, :] = 0; nrm_world[:, 0] = 0; nrm_world[:, -1] = 0  # border left unset by the reference loop
    return np.concatenate([nrm_world, noisy[..., None]], -1).astype(np.float32)


def _edge_label(pid):
    """priors: edges = plane-id boundaries; labels = plane id + 1, 0 on boundaries"""
    H, W = pid.shape
    edge = np.zeros((H, W), np.uint8)
    edge[:, :-1] |= (pid[:, :-1] != pid[:, 1:]).astype(np.uint8)
    edge[:-1, :] |= (pid[:-1, :] != pid[1:, :]).astype(np.uint8)
    edge *= 255
    label = (pid + 1).astype(np.int32)
    label[edge > 0] = 0
    return edge, label


def make_scene(width: int, height: int, num_src: int, seed: int = 0, quantize: bool = True) -> Scene:
    """Build the scene. `quantize` rounds grey levels to integers like an 8-bit image would be."""
    W, H, S = int(width), int(height), int(num_src)
    rng_img = np.random.default_rng(SEED_IMAGE + seed)
    rng_cam = np.random.default_rng(SEED_CAMERA + seed)
    rng_init = np.random.default_rng(SEED_INIT + seed)
    planes = _room()
    fx = 0.55 * W
    K = np.array([[fx, 0, W / 2.0], [0, fx, H / 2.0], [0, 0, 1.0]])
    texels_per_m = fx / 8.0
    tex_size = 2048
    textures = [_make_texture(rng_img, tex_size, 40.0 if i == 3 else 1.5) for i in range(len(planes))]

    c_ref = np.array([0.10, -0.05, 0.20])
    R_ref = _rot_yx(np.deg2rad(3.0), np.deg2rad(-2.0))
    cams_KRc = [(K, R_ref, c_ref)]
    centre = np.array([0.0, 0.0, 5.0])
    for i in range(S):
        ang = 2 * np.pi * (i + 0.25 * rng_cam.uniform()) / max(S, 1)
        rad = rng_cam.uniform(0.25, 0.6)
        c = c_ref + np.array([rad * np.cos(ang), 0.6 * rad * np.sin(ang), rng_cam.uniform(-0.05, 0.05)])
        to = centre - c
        yaw = np.arctan2(to[0], to[2]) + np.deg2rad(rng_cam.uniform(-1, 1))
        pitch = -np.arctan2(to[1], np.hypot(to[0], to[2])) + np.deg2rad(rng_cam.uniform(-1, 1))
        cams_KRc.append((K, _rot_yx(yaw, pitch), c))

    cam_dmin, cam_dmax = 1.5, 12.0
    images = np.zeros((S + 1, H, W), np.float32)
    depths = np.zeros((S + 1, H, W), np.float32)
    cameras = np.zeros((S + 1,), CAMERA_DTYPE)
    pid0 = None
    for i, (Ki, Ri, ci) in enumerate(cams_KRc):
        img, dep, pid = _render(Ki, Ri, ci, W, H, planes, textures, texels_per_m)
        images[i] = np.rint(img) if quantize else img
        depths[i] = dep
        cameras[i] = make_camera(Ki, Ri, ci, W, H, cam_dmin, cam_dmax)
        if i == 0:
            pid0 = pid

    # ground-truth plane map of the reference view: (world normal facing the camera, depth)
    planes_true = np.zeros((H, W, 4), np.float32)
    for i, pl in enumerate(planes):
        m = pid0 == i
        planes_true[m, 0:3] = pl["n"].astype(np.float32)
    planes_true[..., 3] = depths[0]

    planes_init = _first_init_planes(depths[0], K, R_ref, rng_init)

    edge, label = _edge_label(pid0)

    return Scene(W, H, S, images, depths, cameras, planes_init, planes_true, pid0, edge, label,
                 0.6 * cam_dmin, 1.2 * cam_dmax)


# ---------------------------------------------------------------------------------------------------------------
# multi-view, multi-level scenes for the in-memory schedule (dvp_scene_*)
def level_size(full_w: int, full_h: int, scale: int):
    """round(cols * (1 / scale)) in float32, half away from zero (APD.cpp:1119-1123)."""
    f = np.float32(1.0) / np.float32(scale)
    return int(np.floor(np.float32(full_w) * f + np.float32(0.5))), int(np.floor(np.float32(full_h) * f + np.float32(0.5)))


def level_camera(cam_full: np.ndarray, full_w: int, full_h: int, w: int, h: int, scale: int) -> np.ndarray:
    """Intrinsics scaled by the achieved ratios, float32 (APD.cpp:1125-1140)."""
    cam = np.array(cam_full, dtype=CAMERA_DTYPE, copy=True).reshape(())
    if scale != 1:
        sx = np.float32(w) / np.float32(full_w); sy = np.float32(h) / np.float32(full_h)
        K = cam["K"].copy()
        K[0] *= sx; K[2] *= sx; K[4] *= sy; K[5] *= sy
        cam["K"] = K
    cam["width"] = w; cam["height"] = h
    return cam


@dataclasses.dataclass
class MultiView:
    full_w: int
    full_h: int
    num_levels: int
    cameras: np.ndarray        # [V] CAMERA_DTYPE, full resolution
    src_views: list            # [V] list of source view indices
    levels: list               # [level][view] dict(w, h, image, edge, label, depth, camera)
    planes_init: list          # [V] [h0][w0][4] FIRST_INIT prior at level 0
    depth_min: float
    depth_max: float


def make_multiview(full_w: int, full_h: int, num_views: int, num_levels: int, seed: int = 0, num_src: int | None = None) -> MultiView:
    """The room of make_scene seen from `num_views` cameras; every view is rendered directly at each pyramid level
    (scale 2^num_levels ... 2) with the level's intrinsics — standing in for the reference's cv::resize pyramid."""
    V = int(num_views)
    S = V - 1 if num_src is None else int(num_src)
    rng_img = np.random.default_rng(SEED_IMAGE + seed)
    rng_cam = np.random.default_rng(SEED_CAMERA + seed)
    rng_init = np.random.default_rng(SEED_INIT + seed)
    planes = _room()
    fx = 0.55 * full_w
    K = np.array([[fx, 0, full_w / 2.0], [0, fx, full_h / 2.0], [0, 0, 1.0]])
    textures = [_make_texture(rng_img, 2048, 40.0 if i == 3 else 1.5) for i in range(len(planes))]
    c_ref = np.array([0.10, -0.05, 0.20])
    KRc = [(K, _rot_yx(np.deg2rad(3.0), np.deg2rad(-2.0)), c_ref)]
    centre = np.array([0.0, 0.0, 5.0])
    for i in range(V - 1):
        ang = 2 * np.pi * (i + 0.25 * rng_cam.uniform()) / max(V - 1, 1)
        rad = rng_cam.uniform(0.25, 0.6)
        c = c_ref + np.array([rad * np.cos(ang), 0.6 * rad * np.sin(ang), rng_cam.uniform(-0.05, 0.05)])
        to = centre - c
        yaw = np.arctan2(to[0], to[2]) + np.deg2rad(rng_cam.uniform(-1, 1))
        pitch = -np.arctan2(to[1], np.hypot(to[0], to[2])) + np.deg2rad(rng_cam.uniform(-1, 1))
        KRc.append((K, _rot_yx(yaw, pitch), c))
    cam_dmin, cam_dmax = 1.5, 12.0
    cameras = np.zeros((V,), CAMERA_DTYPE)
    for v, (Kv, Rv, cv_) in enumerate(KRc):
        cameras[v] = make_camera(Kv, Rv, cv_, full_w, full_h, cam_dmin, cam_dmax)
    src_views = [[(v + 1 + k) % V for k in range(S)] for v in range(V)]
    levels, planes_init = [], []
    for level in range(num_levels):
        scale = 1 << (num_levels - level)
        w, h = level_size(full_w, full_h, scale)
        per_view = []
        for v, (Kv, Rv, cv_) in enumerate(KRc):
            cam = level_camera(cameras[v], full_w, full_h, w, h, scale)
            Kl = cam["K"].astype(np.float64).reshape(3, 3)
            img, dep, pid = _render(Kl, Rv, cv_, w, h, planes, textures, Kl[0, 0] / 8.0)
            edge, label = _edge_label(pid)
            per_view.append(dict(w=w, h=h, image=np.rint(img).astype(np.float32), edge=edge, label=label, depth=dep, camera=cam))
            if level == 0:
                planes_init.append(_first_init_planes(dep, Kl, Rv, rng_init))
        levels.append(per_view)
    return MultiView(full_w, full_h, num_levels, cameras, src_views, levels, planes_init, 0.6 * cam_dmin, 1.2 * cam_dmax)


def make_fusion_views(mv: MultiView, levels, seed: int = 0, depth_noise: float = 0.001, normal_noise: float = 0.015) -> list:
    """What RunFusion holds after its loading loop (reference APD.cpp:1841-1873), synthesised from a MultiView: per view
    a depth map (true depth with multiplicative noise, 3 % invalid, 2 % gross outliers), world normals (finite
    differences of the true depth, perturbed), a 3-channel uint8 image, a weak map (the textureless wall is WEAK) and
    the source list.  `levels` is one pyramid level for all views or one per view (views of different sizes make many
    pixels claim the same source cell, which is what the fusion's visiting order decides)."""
    V = len(mv.cameras)
    levels = [levels] * V if np.isscalar(levels) else list(levels)
    rng = np.random.default_rng(20250104 + seed)
    out = []
    for v in range(V):
        lv = mv.levels[levels[v]][v]
        cam = np.array(lv["camera"], dtype=CAMERA_DTYPE).reshape(())
        H, W = lv["h"], lv["w"]
        true = lv["depth"].astype(np.float64)
        K = cam["K"].astype(np.float64).reshape(3, 3); R = cam["R"].astype(np.float64).reshape(3, 3)
        xs = (np.arange(W) - K[0, 2]) / K[0, 0]; ys = (np.arange(H) - K[1, 2]) / K[1, 1]
        X = np.stack([xs[None, :] * true, ys[:, None] * true, true], -1)
        dx = np.zeros_like(X); dy = np.zeros_like(X)
        dx[:, :-1] = X[:, 1:] - X[:, :-1]; dx[:, -1] = dx[:, -2]
        dy[:-1, :] = X[1:, :] - X[:-1, :]; dy[-1, :] = dy[-2, :]
        nrm = np.cross(dx, dy)
        nrm /= np.maximum(np.linalg.norm(nrm, axis=-1, keepdims=True), 1e-30)
        nrm[(nrm * X).sum(-1) > 0] *= -1.0
        nrm = nrm @ R + rng.normal(0.0, normal_noise, (H, W, 3))
        nrm /= np.maximum(np.linalg.norm(nrm, axis=-1, keepdims=True), 1e-30)
        depth = true * (1.0 + rng.normal(0.0, depth_noise, (H, W)))
        u = rng.uniform(size=(H, W))
        depth[u < 0.03] = 0.0
        depth[(u >= 0.03) & (u < 0.05)] *= 1.05
        grey = np.clip(lv["image"], 0, 255)
        image = np.stack([grey, np.clip(grey * 0.9 + 10, 0, 255), np.clip(255 - grey, 0, 255)], -1).astype(np.uint8)
        weak = np.where(lv["label"] == 4, 0, 1).astype(np.uint8)   # plane 3 (label 4) is the textureless wall; WEAK = 0
        out.append(dict(camera=cam, depth=depth.astype(np.float32), normal=nrm.astype(np.float32), image=image, weak=weak,
                        src_views=list(mv.src_views[v])))
    return out
