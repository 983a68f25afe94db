import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")
ASSET = os.path.join(ROOT, "parc_b200", "assets", "humanoid.xml")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def golden(name):
    return np.load(os.path.join(GOLDEN, name), allow_pickle=False)


@pytest.fixture(scope="session")
def oracle_model():
    from oracle import parc_oracle as O
    return O.CharModel.from_npz(os.path.join(GOLDEN, "humanoid_model.npz"))


@pytest.fixture(scope="session")
def cpu_model():
    from parc_b200.anim.kin_char_model import KinCharModel
    m = KinCharModel("cpu")
    m.load_char_file(ASSET)
    return m


@pytest.fixture(scope="session")
def gpu_model():
    from parc_b200.anim.kin_char_model import KinCharModel
    m = KinCharModel("cuda:0")
    m.load_char_file(ASSET)
    return m


def lib_clips_from_golden():
    """The 3-clip library the golden query vectors were generated on (oracle/make_golden.py)."""
    from oracle import parc_oracle as O
    civ = golden("clip_civilization.npz")
    tea = golden("clip_teaser_terrain.npz")
    return [O.Clip(civ["frames"], civ["contacts"], 30.0, O.CLAMP, 1.0),
            O.Clip(tea["frames"], tea["contacts"], 30.0, O.WRAP, 2.0),
            O.Clip(civ["frames"][:40], civ["contacts"][:40], 60.0, O.WRAP, 0.5)]


def write_clip_library(tmpdir, clips):
    """Write oracle Clip objects as the reference's pkl + yaml format; returns the yaml path."""
    import pickle
    lines = ["motions:"]
    for i, c in enumerate(clips):
        p = os.path.join(str(tmpdir), f"clip_{i:05d}.pkl")
        with open(p, "wb") as f:
            pickle.dump({"frames": np.asarray(c.frames, np.float32), "contacts": np.asarray(c.contacts, np.float32),
                         "fps": float(c.fps), "loop_mode": "WRAP" if c.loop_mode == 1 else "CLAMP"}, f)
        lines += [f"- file: {p}", f"  weight: {c.weight}"]
    y = os.path.join(str(tmpdir), "lib.yaml")
    with open(y, "w") as f:
        f.write("\n".join(lines) + "\n")
    return y


def assert_close(actual, expected, rtol=1e-5, atol=1e-6, what=""):
    """|a - e| <= atol + rtol * |e| elementwise (north_star: 1e-5 relative, fp32)."""
    a = torch.as_tensor(actual).detach().cpu().double()
    e = torch.as_tensor(expected).detach().cpu().double()
    assert a.shape == e.shape, f"{what}: shape {tuple(a.shape)} vs {tuple(e.shape)}"
    err = (a - e).abs()
    tol = atol + rtol * e.abs()
    bad = err > tol
    if bad.any():
        i = torch.argmax(err - tol)
        raise AssertionError(f"{what}: {int(bad.sum())}/{bad.numel()} elements out of tolerance; worst |err|="
                             f"{err.flatten()[i].item():.3e} at expected={e.flatten()[i].item():.6g}")


def assert_close_normwise(actual, expected, rtol=1e-5, what=""):
    """max|a - e| <= rtol * max|e| -- for gradients, whose small entries are differences of large terms."""
    a = torch.as_tensor(actual).detach().cpu().double()
    e = torch.as_tensor(expected).detach().cpu().double()
    assert a.shape == e.shape, f"{what}: shape {tuple(a.shape)} vs {tuple(e.shape)}"
    scale = max(e.abs().max().item(), 1e-30)
    err = (a - e).abs().max().item()
    assert err <= rtol * scale, f"{what}: max|err|={err:.3e} > {rtol:g} * max|expected|={scale:.3e}"


def make_random_tree_model(num_bodies=21, seed=4):
    """A synthetic character with more than 15 bodies (exercises the one-character-per-warp path): random
    tree, mixed joint types, NON-identity local rotations.  Returns (oracle CharModel, dict of plain lists for
    parc_b200.ops.make_char_model)."""
    from oracle import parc_oracle as O
    rng = np.random.default_rng(seed)
    parents = [-1] + [int(rng.integers(max(0, b - 4), b)) for b in range(1, num_bodies)]
    types = [O.ROOT] + [int(rng.choice([O.HINGE, O.SPHERICAL, O.SPHERICAL, O.FIXED])) for _ in range(1, num_bodies)]
    lt = rng.uniform(-0.3, 0.3, size=(num_bodies, 3)).astype(np.float32)
    lt[0] = 0
    lr = rng.normal(size=(num_bodies, 4)).astype(np.float32)
    lr /= np.linalg.norm(lr, axis=1, keepdims=True)
    lr[0] = [0, 0, 0, 1]
    axes = np.zeros((num_bodies, 3), np.float32)
    dof_idx, dof_dim, d = [], [], 0
    for b in range(num_bodies):
        dim = 1 if types[b] == O.HINGE else (3 if types[b] == O.SPHERICAL else 0)
        if types[b] == O.HINGE:
            a = rng.normal(size=3)
            axes[b] = (a / np.linalg.norm(a)).astype(np.float32)
        dof_idx.append(d)
        dof_dim.append(dim)
        d += dim
    om = O.CharModel(body_names=[f"b{i}" for i in range(num_bodies)], parents=parents, local_trans=torch.tensor(lt),
                     local_rot=torch.tensor(lr), joint_type=types, joint_axis=torch.tensor(axes), dof_idx=dof_idx,
                     dof_dim=dof_dim, body_points=[])
    plain = dict(parents=parents, local_trans=lt.tolist(), local_rot=lr.tolist(), joint_types=types,
                 joint_axes=axes.tolist(), dof_idx=dof_idx)
    return om, plain
