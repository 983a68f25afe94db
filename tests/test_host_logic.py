"""CPU-side checks: the C-ABI library loads and exports every declared symbol, rejects bad arguments
before touching the GPU, and the host logic (MJCF parser, table building, point sampling, synthetic
generators) reproduces the reference's golden values.  No compute kernels are launched here."""
import ctypes as C
import os
import re

import numpy as np
import pytest
import torch

from conftest import ROOT, golden, lib_clips_from_golden, write_clip_library


def T(x):
    return torch.as_tensor(np.asarray(x))


# ------------------------------------------------------------------ C ABI
def declared_symbols():
    src = open(os.path.join(ROOT, "include", "parc_b200.h")).read()
    return sorted(set(re.findall(r"\b(parc_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    from parc_b200 import _lib
    lib = _lib.load()
    names = declared_symbols()
    assert len(names) >= 17
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/parc_b200.h but not exported"
        assert n in _lib.SIGNATURES, f"{n} has no ctypes signature"
    assert set(_lib.SIGNATURES) == set(names)
    assert lib.parc_abi_version() == 3
    assert lib.parc_error_string(-1).decode() == "a required pointer is NULL"


def test_struct_sizes_match_header():
    from parc_b200 import _lib
    assert C.sizeof(_lib.ParcClipMeta) == 32
    assert C.sizeof(_lib.ParcRowLayout) == 32
    assert C.sizeof(_lib.ParcCharModel) == 16 + 4 * 24 * 4 + 24 * (3 + 4 + 3) * 4
    assert C.sizeof(_lib.ParcMotionTables) == 48 and C.sizeof(_lib.ParcFrameOut) == 80
    assert C.sizeof(_lib.ParcHeightfield) == 32 and C.sizeof(_lib.ParcTerrainBatch) == 80


def test_row_layout_and_model_validation(cpu_model):
    from parc_b200 import _lib, ops
    m = cpu_model.c_model()
    lay = ops.row_layout(m)
    assert (lay.pose_slots, lay.contact_slot, lay.vel_slot, lay.vel_slots, lay.row_floats) == (20, 16, 20, 9, 120)
    assert m.max_depth == 4 and list(m.depth)[:15] == [0, 1, 2, 2, 3, 4, 2, 3, 4, 1, 2, 3, 1, 2, 3]
    bad = _lib.ParcCharModel.from_buffer_copy(m)
    bad.parent[3] = 7                               # parent after child
    assert _lib.load().parc_validate_model(C.byref(bad)) == -3
    bad = _lib.ParcCharModel.from_buffer_copy(m)
    bad.num_bodies = 25
    assert _lib.load().parc_validate_model(C.byref(bad)) == -3


def test_row_layout_of_a_large_model():
    from conftest import make_random_tree_model
    from parc_b200 import ops
    om, plain = make_random_tree_model(21)
    m = ops.make_char_model(**plain)
    lay = ops.row_layout(m)
    assert m.num_bodies == 21 and m.dof_size == om.dof_size and m.max_depth >= 2
    assert lay.contact_slot == 22 and lay.pose_slots == 22 + 6 and lay.vel_slots == 2 + (om.dof_size + 3) // 4
    assert lay.row_floats % 8 == 0 and lay.row_floats >= (lay.pose_slots + lay.vel_slots) * 4
    _, plain25 = make_random_tree_model(25)
    from parc_b200._lib import ParcLibraryError
    with pytest.raises(ParcLibraryError):
        ops.make_char_model(**plain25)


def test_argument_errors_are_returned_not_thrown(cpu_model):
    """NULL / negative-size / misaligned arguments come back as negative codes; nothing is launched."""
    from parc_b200 import _lib
    lib = _lib.load()
    m = cpu_model.c_model()
    assert lib.parc_motion_query(None, None, None, 0, C.byref(m), None, None, None, None, None, None) == -1
    tb = _lib.ParcMotionTables()
    tb.rows, tb.clips, tb.total_frames, tb.num_clips, tb.row_floats, tb.tree = 256, 512, 10, 1, 120, 1024
    fo = _lib.ParcFrameOut()
    assert lib.parc_motion_query(C.byref(tb), 64, 64, -5, C.byref(m), C.byref(fo), None, None, None, None, None) == -2
    tb.row_floats = 116
    assert lib.parc_motion_query(C.byref(tb), 64, 64, 0, C.byref(m), C.byref(fo), None, None, None, None, None) == -5
    tb.row_floats, tb.rows = 120, 260                # rows not 16-byte aligned
    assert lib.parc_motion_query(C.byref(tb), 64, 64, 0, C.byref(m), C.byref(fo), None, None, None, None, None) == -4
    assert lib.parc_fk_fwd(None, None, None, 4, C.byref(m), None, None, None) == -1
    assert lib.parc_hf_sample(None, None, 4, None, None, None) == -1
    assert lib.parc_points_hf_sdf(None, 1, 1, None, 1, None, None, None) == -1
    assert lib.parc_exp_map_to_quat_fwd(64, -1, 64, None) == -2
    tb.rows, tb.tree = 256, None                     # the device-resident tree is required
    assert lib.parc_motion_query(C.byref(tb), 64, 64, 0, C.byref(m), C.byref(fo), None, None, None, None, None) == -1
    tb.tree = 1028
    assert lib.parc_motion_query(C.byref(tb), 64, 64, 0, C.byref(m), C.byref(fo), None, None, None, None, None) == -4
    host_tree = (C.c_uint8 * _lib.PARC_TREE_BYTES)()
    assert lib.parc_tree_from_model(C.byref(m), host_tree) == 0 and lib.parc_tree_from_model(C.byref(m), None) == -1
    words = np.frombuffer(bytes(host_tree), np.int32)
    assert words[0] == m.num_bodies and words[1] == m.dof_size and words[2] == m.max_depth
    assert list(words[4:4 + m.num_bodies]) == list(m.parent)[:m.num_bodies]
    # n == 0 is a valid no-op
    tb.rows, tb.tree = 256, 1024
    assert lib.parc_motion_query(C.byref(tb), 64, 64, 0, C.byref(m), C.byref(fo), None, None, None, None, None) == 0


def test_ops_refuse_cpu_tensors(cpu_model):
    from parc_b200 import ops
    from parc_b200._lib import ParcLibraryError
    with pytest.raises(ParcLibraryError):
        cpu_model.forward_kinematics(torch.zeros(2, 3), torch.zeros(2, 4), torch.zeros(2, 14, 4))
    with pytest.raises(ParcLibraryError):
        cpu_model.dof_to_rot(torch.zeros(2, 28))
    with pytest.raises(ParcLibraryError):
        ops.exp_map_to_quat(torch.zeros(2, 3))


def test_product_never_imports_the_oracle():
    import subprocess
    import sys
    code = ("import sys; sys.path.insert(0, %r); import parc_b200.anim.motion_lib, parc_b200.tools.procgen.mdm_path, "
            "parc_b200.tools.motion_opt.motion_optimization, parc_b200.envs.ig_parkour.mgdm_dm_util, parc_b200.sharding, "
            "parc_b200.envs.ig_parkour.step_assembly, parc_b200.envs.ig_char_env, parc_b200.anim.packed_format, "
            "parc_b200.zmotion_editing_tools.motion_edit_lib; "
            "assert not any(m == 'oracle' or m.startswith('oracle.') for m in sys.modules), 'oracle imported'" % ROOT)
    subprocess.run([sys.executable, "-c", code], check=True)
    # statically: outside tests/, oracle/ itself, smoke() in __graft_entry__.py and the CPU legs of bench.py, no
    # python file of the repository names the oracle package in an import
    pat = re.compile(r"^\s*(from\s+oracle\b|import\s+oracle\b)", re.M)
    offenders = []
    for dirpath, dirnames, files in os.walk(ROOT):
        dirnames[:] = [d for d in dirnames if d not in (".git", "tests", "oracle", "gpurun_out", "__pycache__", "baseline")]
        for f in files:
            if f.endswith(".py") and os.path.join(dirpath, f) not in (os.path.join(ROOT, "bench.py"),
                                                                       os.path.join(ROOT, "__graft_entry__.py")):
                if pat.search(open(os.path.join(dirpath, f)).read()):
                    offenders.append(os.path.relpath(os.path.join(dirpath, f), ROOT))
    assert not offenders, offenders
    for dirpath, _, files in os.walk(os.path.join(ROOT, "parc_b200")):
        for f in files:
            if f.endswith(".py"):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, re.M), f"{f} imports the oracle"


# ------------------------------------------------------------------ character model
def test_mjcf_parser_matches_reference_model(cpu_model):
    g = golden("humanoid_model.npz")
    assert cpu_model.get_body_names() == [str(s) for s in g["body_names"]]
    assert cpu_model._parent_indices.tolist() == g["parents"].tolist()
    assert torch.equal(cpu_model._local_translation, T(g["local_translation"]))
    assert torch.equal(cpu_model._local_rotation, T(g["local_rotation"]))
    assert [j.joint_type.value for j in cpu_model._joints] == g["joint_type"].tolist()
    assert [j.dof_idx for j in cpu_model._joints] == g["dof_idx"].tolist()
    assert cpu_model.get_dof_size() == 28 and cpu_model.get_num_joints() == 15
    assert torch.equal(cpu_model._lower_dof_limits, T(g["lower_dof_limits"]))
    assert torch.equal(cpu_model._upper_dof_limits, T(g["upper_dof_limits"]))
    for j, jt in enumerate(cpu_model._joints):
        if jt.axis is not None:
            assert torch.equal(jt.axis, T(g["joint_axis"][j]))
    assert cpu_model.get_body_id("left_foot") == 14 and cpu_model.get_joint_id("torso") == 0
    rows = []
    for b in range(15):
        for ge in cpu_model.get_geoms(b):
            d = np.zeros(3, np.float32)
            dd = ge._dims.numpy().reshape(-1)
            d[:dd.shape[0]] = dd
            rows.append([b, ge._shape_type.value, *ge._offset.tolist(), *d.tolist(), -1.0 if ge._radius is None else ge._radius])
    assert np.array_equal(np.array(rows, np.float32), g["geoms"])


def test_body_point_samples_match_reference(cpu_model):
    from parc_b200.util import geom_util
    g = golden("humanoid_model.npz")
    pts = geom_util.get_char_point_samples(cpu_model)
    assert [p.shape[0] for p in pts] == g["body_point_counts"].tolist() and sum(p.shape[0] for p in pts) == 304
    assert torch.equal(torch.cat(pts), T(g["body_points"]))
    mp = geom_util.get_minimal_char_point_samples(cpu_model)
    assert torch.equal(torch.cat(mp), T(g["min_body_points"]))
    # foot sole = first 18 points at z = offset_z - half_z (motion_optimization.py:320 relies on it)
    assert torch.allclose(pts[11][:18, 2], torch.full((18,), -0.05))


def test_templates_match_reference():
    from parc_b200.util import geom_util
    g = golden("obs_golden.npz")
    tm = geom_util.get_xy_points_cone(torch.zeros(2), 0.05, 2, 60, 3, 3, 0.26179938779)
    assert tm.shape == (441, 2) and torch.equal(tm, T(g["tmpl"]))
    gt = geom_util.get_xy_grid_points(torch.zeros(2), 0.2, 0.2, 15, 15, 15, 15)
    assert gt.shape == (31, 31, 2) and torch.equal(gt, T(g["grid_tmpl"]))


def test_host_quaternion_helpers_match_reference(cpu_model):
    from parc_b200.util import torch_util
    civ, g = golden("clip_civilization.npz"), golden("dof_golden.npz")
    assert torch.equal(torch_util.exp_map_to_quat(T(civ["frames"][:, 3:6])), T(g["root_quat"]))
    assert torch.equal(cpu_model.host_dof_to_rot(T(civ["frames"][:, 6:])), T(g["joint_rot"]))
    assert torch.equal(cpu_model.rot_to_dof(T(g["joint_rot"])), T(g["dof_back"]))
    og = golden("obs_golden.npz")
    assert torch.equal(torch_util.calc_heading(T(og["root_quat"])), T(og["heading"]))
    # the encoding the policy observations use, against the reference's compute_char_obs output (global frame:
    # columns 0:6 are quat_to_tan_norm(root_rot), 12:96 the joints')
    ts = golden("tracker_step_golden.npz")
    assert torch.equal(torch_util.quat_to_tan_norm(T(ts["root_rot"])), T(ts["char_obs_g1_h0"])[:, 0:6])
    assert torch.equal(torch_util.quat_to_tan_norm(T(ts["joint_rot"])).reshape(-1, 84), T(ts["char_obs_g1_h0"])[:, 12:96])
    q = torch_util.quat_unit(T(ts["root_rot"]))
    ident = torch_util.quat_multiply(q, torch_util.quat_inv(q))
    assert torch.allclose(ident, torch.tensor([0.0, 0.0, 0.0, 1.0]).expand_as(ident), atol=1e-6)
    assert torch.allclose(torch_util.quat_abs(q), torch.ones(q.shape[0]), atol=1e-6)
    hq = torch_util.heading_to_quat(torch_util.calc_heading(q))
    assert torch.equal(hq, torch_util.calc_heading_quat(q))


# ------------------------------------------------------------------ MotionLib loading
def test_motion_file_loader_builds_reference_tables(cpu_model, tmp_path):
    from parc_b200.anim.motion_lib import MotionLib
    g = golden("tables_golden.npz")
    lib = MotionLib(write_clip_library(tmp_path, lib_clips_from_golden()), cpu_model, "cpu", init_type="motion_file",
                    contact_info=True)
    assert lib.num_motions() == 3 and lib._motion_num_frames.tolist() == [254, 58, 40]
    for k, a in (("root_rot", lib._frame_root_rot), ("joint_rot", lib._frame_joint_rot), ("root_vel", lib._frame_root_vel),
                 ("root_ang_vel", lib._frame_root_ang_vel), ("dof_vel", lib._frame_dof_vel), ("lengths", lib._motion_lengths),
                 ("start_idx", lib._motion_start_idx), ("root_pos_delta", lib._motion_root_pos_delta),
                 ("weights", lib._motion_weights)):
        assert torch.equal(a, T(g[k])), k
    assert lib._frame_contacts.shape == (352, 15) and lib._motion_frames.shape == (352, 34)
    assert lib.get_motion_names() == ["clip_00000", "clip_00001", "clip_00002"]
    assert lib.get_motion_loop_mode(torch.tensor([0, 1])).tolist() == [0, 1]
    ph = lib.calc_motion_phase(torch.tensor([0, 1, 1]), torch.tensor([100.0, 2.85, -0.5]))
    assert ph[0] == 1.0 and 0 <= ph[1] < 1 and 0 <= ph[2] < 1
    assert lib._packed is None                       # no CUDA device: tables exist, queries do not
    from parc_b200._lib import ParcLibraryError
    with pytest.raises(ParcLibraryError):
        lib.calc_motion_frame(torch.tensor([0]), torch.tensor([0.0]))


def test_motion_frames_loader_reproduces_reference_quirk(cpu_model):
    from parc_b200.anim.motion_lib import LoopMode, MotionLib
    g = golden("tables_motion_frames_golden.npz")
    lib = MotionLib(T(g["frames"]), cpu_model, "cpu", init_type="motion_frames", loop_mode=LoopMode.CLAMP, fps=30,
                    contact_info=True, contacts=T(g["contacts"]))
    assert torch.equal(lib._frame_dof_vel, T(g["dof_vel"]))       # fps passed as dt: 900x too small
    assert torch.equal(lib._frame_root_ang_vel, T(g["root_ang_vel"]))
    assert torch.equal(lib._motion_lengths, T(g["lengths"]))
    with pytest.raises(ValueError):
        MotionLib("x", cpu_model, "cpu", init_type="diffusion_file")


def test_reference_pickles_with_terrain_load(cpu_model, tmp_path):
    """A clip pickle that names the reference's SubTerrain class resolves to ours."""
    import pickle
    import sys
    import types
    from parc_b200.anim.motion_lib import load_clip_file
    from parc_b200.util.terrain_util import SubTerrain
    civ = golden("clip_civilization.npz")
    fake_pkg, fake_mod = types.ModuleType("util"), types.ModuleType("util.terrain_util")

    class _Ref:
        pass
    _Ref.__name__ = _Ref.__qualname__ = "SubTerrain"
    _Ref.__module__ = "util.terrain_util"
    fake_mod.SubTerrain = _Ref
    saved = {k: sys.modules.get(k) for k in ("util", "util.terrain_util")}
    sys.modules["util"], sys.modules["util.terrain_util"] = fake_pkg, fake_mod
    try:
        t = _Ref()
        t.terrain_name, t.hf, t.dims = "t", civ["hf"], np.array([50, 50])
        t.min_point, t.dxdy, t.hf_mask = civ["min_point"], civ["dxdy"], np.zeros((50, 50), bool)
        p = tmp_path / "c.pkl"
        with open(p, "wb") as f:
            pickle.dump({"frames": civ["frames"], "contacts": civ["contacts"], "fps": 30.0, "loop_mode": "CLAMP",
                         "terrain": t}, f)
    finally:
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v
    clip = load_clip_file(str(p))
    terr = clip["terrain"]
    assert isinstance(terr, SubTerrain)
    terr.update_old()
    terr.to_torch("cpu")
    assert terr.hf.shape == (50, 50) and terr.hf_maxmin.shape == (50, 50, 2) and terr.dims.tolist() == [50, 50]


# ------------------------------------------------------------------ synthetic inputs
def test_synthetic_generators_are_seeded_and_in_spec(cpu_model):
    from parc_b200.util import synth
    rng = np.random.default_rng(0)
    hf = synth.box_terrain(rng)
    assert hf.shape == (16, 16) and hf.dtype == np.float32
    st = synth.stairs_terrain(np.random.default_rng(1))
    assert (np.diff(st[:, 0]) >= 0).all() and st.max() > 0
    f1, c1 = synth.synth_clips(cpu_model, 3, seed=5, hf=hf)
    f2, c2 = synth.synth_clips(cpu_model, 3, seed=5, hf=hf)
    assert np.array_equal(f1, f2) and np.array_equal(c1, c2)
    assert f1.shape == (3, 265, 34) and c1.shape == (3, 265, 15) and f1.dtype == np.float32
    lo, hi = synth.dof_limits(cpu_model)
    assert (f1[..., 6:] >= lo - 1e-6).all() and (f1[..., 6:] <= hi + 1e-6).all()
    assert (np.abs(f1[..., 6:]) >= 1e-3 - 1e-9).all()                      # never exactly zero (F8d)
    n = np.linalg.norm(f1[..., 3:6], axis=-1)
    assert (n > 1e-3).all() and (n <= 0.5 + 1e-6).all()
    speed = np.linalg.norm(np.diff(f1[..., 0:2], axis=1), axis=-1) * 30.0
    assert speed.max() <= 3.0 + 1e-3
    assert set(np.unique(c1)) <= {0.0, 1.0}
    s = synth.synth_motion_samples(cpu_model, 2, 10, hf, (0.0, 0.0), (0.4, 0.4))
    assert s["contacts"].min() < 0 or s["contacts"].min() == 0
    assert s["root_pos"].shape == (2, 10, 3) and s["joint_dof"].shape == (2, 10, 28)


def test_packed_container_round_trip_and_rejections(tmp_path):
    """anim/packed_format.py: the .parcpack container (header + 64-byte-aligned raw arrays) without a GPU."""
    from parc_b200.anim import packed_format as pf
    rng = np.random.default_rng(0)
    arrays = {"rows": rng.normal(size=(37, 120)).astype(np.float32), "num_frames": np.array([30, 7], np.int64),
              "loop_modes": np.array([0, 1], np.int32), "mask": (rng.uniform(size=(5, 3)) < 0.5).astype(np.uint8),
              "empty": np.zeros((0, 4), np.float32)}
    meta = {"num_clips": 2, "names": ["a", "b"], "terrains": [None, {"name": "t", "dxdy": [0.4, 0.4]}]}
    path = str(tmp_path / "x.parcpack")
    pf.write_container(path, meta, arrays)
    raw = open(path, "rb").read()
    assert raw[:8] == pf.MAGIC and not os.path.exists(path + ".tmp")
    m2, a2 = pf.read_container(path)
    assert m2 == meta and set(a2) == set(arrays)
    head = __import__("json").loads(raw[16:16 + int(np.frombuffer(raw[12:16], "<u4")[0])])
    for k, v in arrays.items():
        assert a2[k].dtype == v.dtype and a2[k].shape == v.shape and np.array_equal(a2[k], v), k
        assert head["arrays"][k]["offset"] % 64 == 0
    for bad, why in ((b"NOTAPACK" + raw[8:], "magic"), (raw[:8] + np.array([99], "<u4").tobytes() + raw[12:], "version"),
                     (raw[:len(raw) // 2], "truncated"), (raw[:10], "short")):
        p = str(tmp_path / f"bad_{why}.parcpack")
        open(p, "wb").write(bad)
        with pytest.raises(ValueError):
            pf.read_container(p)


def test_peer_gather_entry_points_reject_bad_arguments():
    """parc_peer_push / parc_peer_barrier validate before any launch (csrc/peer_gather.cu)."""
    from parc_b200 import _lib
    lib = _lib.load()
    sig = _lib.ParcPeerSignals()
    assert lib.parc_peer_barrier(None, 0, None) == -1
    assert lib.parc_peer_barrier(C.byref(sig), 0, None) == -2                  # world < 1
    sig.world, sig.num_slots = 2, 4
    assert lib.parc_peer_barrier(C.byref(sig), 0, None) == -1                  # no local slots / epochs
    sig.local_signal, sig.epoch = 4096, 8192
    assert lib.parc_peer_barrier(C.byref(sig), 0, None) == -1                  # neither multicast nor peer addresses
    assert lib.parc_peer_barrier(C.byref(sig), 4, None) == -2                  # slot beyond num_slots
    assert lib.parc_peer_barrier(C.byref(sig), -1, None) == -2
    sig.world = _lib.PARC_MAX_PEERS + 1
    assert lib.parc_peer_barrier(C.byref(sig), 0, None) == -2
    sig.world = 2
    sig.peer_signal[0], sig.peer_signal[1] = 4096, 16384
    sig.local_signal = 4100
    assert lib.parc_peer_barrier(C.byref(sig), 0, None) == -4                  # 8-byte counters
    sig.local_signal = 4096
    seg = (_lib.ParcPeerSegment * 2)()
    assert lib.parc_peer_push(None, 1, C.byref(sig), 4, None) == -1
    assert lib.parc_peer_push(seg, 0, C.byref(sig), 4, None) == -2
    assert lib.parc_peer_push(seg, _lib.PARC_MAX_PUSH_SEGMENTS + 1, C.byref(sig), 4, None) == -2
    assert lib.parc_peer_push(seg, 1, C.byref(sig), 8, None) == -2             # more blocks than signal slots
    seg[0].bytes = 64
    assert lib.parc_peer_push(seg, 1, C.byref(sig), 4, None) == -1             # no source
    seg[0].src = 4096
    assert lib.parc_peer_push(seg, 1, C.byref(sig), 4, None) == -1             # no destination for rank 0
    seg[0].dst_peer[0], seg[0].dst_peer[1] = 8192, 12288
    seg[0].bytes = 66
    assert lib.parc_peer_push(seg, 1, C.byref(sig), 4, None) == -2             # not a multiple of 4 bytes
    seg[0].bytes, seg[0].src = 64, 4098
    assert lib.parc_peer_push(seg, 1, C.byref(sig), 4, None) == -4             # misaligned for fp32
    assert C.sizeof(_lib.ParcPeerSignals) == 8 + 8 * _lib.PARC_MAX_PEERS + 8 + 8 + 8 + 8 + 8 + 8
    sig.timeout_ns = -5
    assert lib.parc_peer_barrier(C.byref(sig), 0, None) == -2
    sig.timeout_ns = 0
    assert C.sizeof(_lib.ParcPeerSegment) == 8 + 8 + 8 * _lib.PARC_MAX_PEERS + 8


def test_step_and_loader_entry_points_reject_bad_arguments(cpu_model):
    """Argument validation of the §8(f)-3 / §8(f)-4 entry points (no launch happens on any of these paths)."""
    from parc_b200 import _lib
    lib = _lib.load()
    m = cpu_model.c_model()
    st = _lib.ParcCharState()
    st.env_stride = 1
    # empty batches are no-ops even with NULL pointers
    assert lib.parc_char_obs(C.byref(st), 0, 14, 28, 4, 0, 0, None, 0, None) == 0
    assert lib.parc_tar_obs(None, None, None, None, None, None, 0, 6, 14, 4, 0, 0, 6, None, 0, None, 0, None) == 0
    assert lib.parc_char_obs(C.byref(st), 8, 14, 28, 4, 0, 0, None, 0, None) == -1          # NULL state arrays
    assert lib.parc_char_obs(None, 8, 14, 28, 4, 0, 0, 64, 0, None) == -1
    assert lib.parc_char_obs(C.byref(st), -1, 14, 28, 4, 0, 0, 64, 0, None) == -2
    for f in ("root_pos", "root_rot", "root_vel", "root_ang_vel", "joint_rot", "dof_vel", "key_pos"):
        setattr(st, f, 4096)
    assert lib.parc_char_obs(C.byref(st), 8, 14, 28, 4, 0, 0, 4096, 100, None) == -2        # out_stride < row width
    st.root_rot = 4100
    assert lib.parc_char_obs(C.byref(st), 8, 14, 28, 4, 0, 0, 4096, 0, None) == -4          # misaligned quaternions
    st.root_rot, st.env_stride = 4096, 0
    assert lib.parc_char_obs(C.byref(st), 8, 14, 28, 4, 0, 0, 4096, 0, None) == -2          # env_stride < 1
    st.env_stride = 1
    # the reference cannot form the key-body reward term without key bodies; neither can we
    assert lib.parc_deepmimic_reward(C.byref(st), C.byref(st), 8, 14, 28, 0, 4096, 4096, 1, 1, 4096, None) == -2
    # targets: the env stride must cover the steps
    assert lib.parc_tar_obs(4096, 4096, 4096, 4096, 4096, 4096, 8, 6, 14, 4, 0, 0, 5, None, 0, 4096, 0, None) == -2
    spec = _lib.ParcDoneSpec()
    assert lib.parc_done(None, 4096, None, 4096, None, None, None, None, None, None, 0, 1, 8, 15, 4096, None, None) == -1
    assert lib.parc_done(C.byref(spec), 4096, None, 4096, None, None, None, None, None, None, 0, 1, 8, 40, 4096, None, None) == -2
    spec.enable_early_termination, spec.has_contact_bodies = 1, 1
    # fall test requested without contact forces / without any height source
    assert lib.parc_done(C.byref(spec), 4096, None, 4096, None, None, None, None, None, None, 0, 1, 8, 15, 4096, None, None) == -1
    assert lib.parc_done(C.byref(spec), 4096, None, 4096, None, None, 4096, None, None, None, 0, 1, 8, 15, 4096, None, None) == -1
    # loader
    assert lib.parc_build_tables(None, 0, 34, None, None, None, None, None, None, 0, C.byref(m), None, None) == 0
    assert lib.parc_build_tables(4096, 10, 20, None, 4096, 4096, 4096, 4096, 4096, 1, C.byref(m), 4096, None) == -2   # stride < 6+D
    assert lib.parc_build_tables(4096, 10, 34, None, 4096, 4096, 4096, 4096, 4096, 1, C.byref(m), None, None) == -1
    assert lib.parc_build_tables(4096, 10, 34, None, 4096, 4096, 4096, 4096, 4096, 1, C.byref(m), 4100, None) == -4
    assert lib.parc_build_tables(4096, 10, 34, None, 4096, 4096, 4096, 4096, 4096, 0, C.byref(m), 4096, None) == -2
    # heightmap observation needs a heading or a root rotation
    hf = _lib.ParcHeightfield()
    hf.hf, hf.dim_x, hf.dim_y, hf.dx, hf.dy = 4096, 4, 4, 0.4, 0.4
    ob = _lib.ParcObsSpec()
    ob.tmpl_xy, ob.num_points, ob.relative = 4096, 8, 1
    assert lib.parc_hf_obs(C.byref(hf), C.byref(ob), 4096, 3, None, None, None, 0, 4, 4096, 0, None) == -1
    assert lib.parc_hf_obs(C.byref(hf), C.byref(ob), 4096, 3, 4096, None, 4096, 2, 4, 4096, 0, None) == -2   # offset needs z
    assert lib.parc_hf_obs(C.byref(hf), C.byref(ob), 4096, 3, 4096, None, None, 0, 4, 4096, 4, None) == -2    # out_stride < P


def test_header_is_plain_c_and_links_from_c(tmp_path):
    """The drop-in boundary is a C ABI: include/parc_b200.h must compile as C99 (no C++-isms) and a plain C program
    must link against the shared library and call a host-only entry point."""
    import shutil
    import subprocess
    if shutil.which("gcc") is None:
        pytest.skip("no gcc")
    from parc_b200 import _lib
    _lib.load()
    src = tmp_path / "abi.c"
    src.write_text('#include "parc_b200.h"\n'
                   "int main(void) {\n"
                   "  ParcCharModel m; ParcRowLayout lay; ParcSimStep s; ParcMotionTables t;\n"
                   "  (void)s; (void)t; m.num_bodies = 0;\n"
                   "  if (parc_abi_version() != PARC_ABI_VERSION) return 1;\n"
                   "  if (parc_row_layout(&m, &lay) != PARC_E_MODEL) return 2;   /* 0 bodies: rejected, nothing launched */\n"
                   "  if (sizeof(ParcClipMeta) != 32 || PARC_TREE_BYTES != 880) return 3;\n"
                   "  return 0;\n}\n")
    libdir = os.path.join(ROOT, "parc_b200")
    exe = str(tmp_path / "abi")
    subprocess.run(["gcc", "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", "-I", os.path.join(ROOT, "include"),
                    str(src), "-L", libdir, "-lparc_b200", "-Wl,-rpath," + libdir, "-o", exe], check=True)
    assert subprocess.run([exe]).returncode == 0


def test_mask_index_lists_to_bit_rows():
    """The reference keeps per-frame [n,2] cell lists (util/terrain_util.py:1951-1997); the device wants bit rows."""
    from parc_b200.diffusion.mdm_heightfield_contact_motion_sampler import mask_inds_to_bits
    inds = [torch.tensor([[0, 0], [1, 2], [3, 9], [3, 9]]), torch.zeros(0, 2, dtype=torch.long), torch.tensor([[4, 1]])]
    bits = mask_inds_to_bits(inds, dim_y=10, words=2)
    assert bits.shape == (3, 2) and bits.dtype == np.uint32
    want0 = np.zeros(64, dtype=bool)
    want0[[0, 12, 39]] = True
    got0 = ((bits[0][:, None] >> np.arange(32, dtype=np.uint32)) & 1).astype(bool).reshape(-1)
    assert np.array_equal(got0, want0) and bits[1].sum() == 0 and bits[2][1] == (1 << (41 - 32))
