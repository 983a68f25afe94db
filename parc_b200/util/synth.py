"""Seeded synthetic clips and heightfields of the reference's shapes (the dataset is not public).

Clip: 265 frames x 34 DoF (root pos 3, root exp-map 3, 28 joint DoFs), 15 contact flags, 30 fps --
README.md:94-101 of the reference.  Terrains: procedural boxes / stairs on a 0.4 m grid like
PARC/kin_gen_default.yaml:8-11, :98-121.  Everything is generated on the host with numpy so the oracle
and the CUDA path are fed bit-identical inputs; nothing here is on the hot path.
"""
from __future__ import annotations

from typing import Optional, Tuple

import numpy as np

NUM_FRAMES = 265
FPS = 30.0


def dof_limits(char_model) -> Tuple[np.ndarray, np.ndarray]:
    lo = char_model._lower_dof_limits.detach().cpu().numpy().astype(np.float64)
    hi = char_model._upper_dof_limits.detach().cpu().numpy().astype(np.float64)
    return lo, hi


def box_terrain(rng: np.random.Generator, dim_x=16, dim_y=16, num_boxes=10, h_range=(-2.0, 2.0),
                len_range=(2, 6)) -> np.ndarray:
    """Flat ground with random axis-aligned boxes (raised or sunken)."""
    hf = np.zeros((dim_x, dim_y), dtype=np.float32)
    for _ in range(num_boxes):
        lx, ly = rng.integers(len_range[0], len_range[1] + 1, size=2)
        x0 = rng.integers(0, max(1, dim_x - lx + 1))
        y0 = rng.integers(0, max(1, dim_y - ly + 1))
        hf[x0:x0 + lx, y0:y0 + ly] = np.float32(rng.uniform(*h_range))
    return hf


def stairs_terrain(rng: np.random.Generator, dim_x=16, dim_y=16, num_stairs=4, step_range=(0.15, 0.25),
                   thickness_range=(2, 8)) -> np.ndarray:
    """A run of steps along x, each `thickness` cells deep."""
    hf = np.zeros((dim_x, dim_y), dtype=np.float32)
    x = int(rng.integers(0, max(1, dim_x // 4)))
    h = 0.0
    for _ in range(num_stairs):
        t = int(rng.integers(thickness_range[0], thickness_range[1] + 1))
        h += float(rng.uniform(*step_range))
        hf[x:, :] = np.float32(h)
        x += t
        if x >= dim_x:
            break
    return hf


def rolling_terrain(rng: np.random.Generator, dim_x: int, dim_y: int, num_boxes: int, h_range=(-1.0, 1.5),
                    len_range=(3, 12)) -> np.ndarray:
    """Large tracker-style global heightfield: many boxes scattered over a big grid."""
    hf = np.zeros((dim_x, dim_y), dtype=np.float32)
    lx = rng.integers(len_range[0], len_range[1] + 1, size=num_boxes)
    ly = rng.integers(len_range[0], len_range[1] + 1, size=num_boxes)
    x0 = rng.integers(0, dim_x, size=num_boxes)
    y0 = rng.integers(0, dim_y, size=num_boxes)
    hh = rng.uniform(h_range[0], h_range[1], size=num_boxes).astype(np.float32)
    for i in range(num_boxes):
        hf[x0[i]:x0[i] + lx[i], y0[i]:y0[i] + ly[i]] = hh[i]
    return hf


def nearest_height(hf: np.ndarray, min_xy, dxdy, xy: np.ndarray) -> np.ndarray:
    g = np.rint((xy - np.asarray(min_xy)) / np.asarray(dxdy)).astype(np.int64)
    g[..., 0] = np.clip(g[..., 0], 0, hf.shape[0] - 1)
    g[..., 1] = np.clip(g[..., 1], 0, hf.shape[1] - 1)
    return hf[g[..., 0], g[..., 1]]


def synth_clips(char_model, num_clips: int, seed: int = 1234, num_frames: int = NUM_FRAMES, fps: float = FPS,
                hf: Optional[np.ndarray] = None, min_xy=(0.0, 0.0), dxdy=(0.4, 0.4),
                area: Optional[Tuple[float, float, float, float]] = None,
                contact_p: float = 0.3) -> Tuple[np.ndarray, np.ndarray]:
    """-> frames [M, F, 6+D] f32, contacts [M, F, J] f32 (0/1).

    Root xy follows a smooth path (<= ~3 m/s) inside `area` (x0, y0, x1, y1); root z rides 0.9 +- 0.05 m
    above the local terrain height; the root exp-map has norm in (1e-3, 0.5] (mostly yaw); joint DoFs are
    band-limited sinusoids inside the joint limits and never exactly zero (the reference's exp-map
    gradient is NaN at 0)."""
    rng = np.random.default_rng(seed)
    lo, hi = dof_limits(char_model)
    D = lo.shape[0]
    J = char_model.get_num_joints()
    M, F = num_clips, num_frames
    t = np.arange(F, dtype=np.float64)[None, :] / fps                      # [1,F]

    if area is None:
        if hf is not None:
            area = (min_xy[0] + 1.0, min_xy[1] + 1.0, min_xy[0] + (hf.shape[0] - 1) * dxdy[0] - 1.0,
                    min_xy[1] + (hf.shape[1] - 1) * dxdy[1] - 1.0)
        else:
            area = (0.0, 0.0, 6.0, 6.0)
    start = np.stack([rng.uniform(area[0], area[2], size=M), rng.uniform(area[1], area[3], size=M)], axis=-1)
    speed = rng.uniform(0.3, 3.0, size=(M, 1))
    yaw0 = rng.uniform(-np.pi, np.pi, size=(M, 1))
    yaw_rate = rng.uniform(-0.6, 0.6, size=(M, 1))
    yaw = yaw0 + yaw_rate * t                                              # [M,F]
    vel = speed[..., None] * np.stack([np.cos(yaw), np.sin(yaw)], axis=-1)  # [M,F,2]
    xy = start[:, None, :] + np.cumsum(vel, axis=1) / fps
    xy[..., 0] = np.clip(xy[..., 0], area[0], area[2])
    xy[..., 1] = np.clip(xy[..., 1], area[1], area[3])
    ground = nearest_height(hf, min_xy, dxdy, xy) if hf is not None else 0.0
    z = ground + 0.9 + 0.05 * np.sin(2 * np.pi * rng.uniform(0.5, 2.0, size=(M, 1)) * t + rng.uniform(0, 6.28, size=(M, 1)))

    # root exp-map: axis near +z (tilted a little), angle in (1e-3, 0.5]
    tilt = 0.15 * np.stack([np.sin(1.3 * t + rng.uniform(0, 6.28, size=(M, 1))),
                            np.cos(0.7 * t + rng.uniform(0, 6.28, size=(M, 1)))], axis=-1)
    axis = np.concatenate([tilt, np.ones((M, F, 1))], axis=-1)
    axis /= np.linalg.norm(axis, axis=-1, keepdims=True)
    ang = 0.2505 + 0.2495 * np.sin(yaw_rate * 2.0 * t + yaw0)               # in (1e-3, 0.5]
    root_exp = axis * ang[..., None]

    # joint DoFs: mid + amp * sin, strictly inside limits, nudged away from 0
    mid = 0.5 * (lo + hi)
    amp = 0.35 * (hi - lo)
    freq = rng.uniform(0.2, 1.5, size=(M, 1, D))
    ph = rng.uniform(0, 2 * np.pi, size=(M, 1, D))
    dofs = mid + amp * np.sin(2 * np.pi * freq * t[..., None] + ph)
    tiny = np.abs(dofs) < 1e-3
    dofs = np.where(tiny, np.where(dofs >= 0, 1e-3, -1e-3), dofs)

    frames = np.concatenate([xy, z[..., None], root_exp, dofs], axis=-1).astype(np.float32)
    contacts = (rng.uniform(size=(M, F, J)) < contact_p).astype(np.float32)
    return frames, contacts


def synth_motion_samples(char_model, batch: int, frames: int, hf: np.ndarray, min_xy, dxdy, seed: int = 99):
    """Inputs of the kin-gen ranking loss (config 3): pose quantities as the MDM would emit them.
    -> dict of float32 arrays root_pos [B,F,3], root_exp [B,F,3], joint_dof [B,F,D], contacts [B,F,J]
    with contacts in [-0.05, 1] (small negatives included, cf. motion_optimization.py:233-234) and root z
    perturbed +-0.1 m so that some body points penetrate."""
    fr, ct = synth_clips(char_model, batch, seed=seed, num_frames=frames, hf=hf, min_xy=min_xy, dxdy=dxdy)
    rng = np.random.default_rng(seed + 7)
    fr[..., 2] += rng.uniform(-0.1, 0.1, size=fr.shape[:2]).astype(np.float32)
    soft = rng.uniform(-0.05, 1.0, size=ct.shape).astype(np.float32)
    contacts = np.where(ct > 0.5, soft, np.float32(0.0)).astype(np.float32)
    return {"root_pos": fr[..., 0:3].copy(), "root_exp": fr[..., 3:6].copy(), "joint_dof": fr[..., 6:].copy(),
            "contacts": contacts}
