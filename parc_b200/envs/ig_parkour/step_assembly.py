"""The kinematic side of one control step of the tracking environment, as a fixed launch sequence.

The reference spreads this over `DMEnv._update_ref_motion` (envs/ig_parkour/dm_env.py:570-595),
`DMEnv.compute_tar_obs` (:686-718), `IGParkourEnv._refresh_obs_hfs` / `_compute_obs` / `_update_reward` /
`_update_done` (envs/ig_parkour/ig_parkour_env.py:636-655, :1053-1248, :1270-1312, :1250-1268) -- several
hundred eager torch ops per step.  `TrackerStep` performs the same computations with the operators of this
package, reading the simulator's state tensors in place:

  1 launch   reference frame + the S future target frames + FK of all of them   (parc_motion_query_steps,
             with the per-env terrain placement of `_move_to_motion_terrain` applied inside)
  1          DoF -> joint rotations of the simulated character                  (parc_dof_to_rot_fwd)
  1          ray heightmap around the simulated character                       (parc_hf_obs, heading from the root
                                                                                 rotation and env offset inside)
  1          proprioceptive observation                                         (parc_char_obs)
  1          target observation, written straight into the policy observation   (parc_tar_obs)
  1          DeepMimic reward terms                                             (parc_deepmimic_reward)
  1          episode flags incl. termination heights                            (parc_done)
  + 2 small copies (target / character contact flags).  The observation operators write their blocks of the
  policy-observation row in place (row-strided outputs), so there is no concatenation pass.

With `fused=True` (default) the simulated character's share -- DoF conversion, proprioceptive observation, reward
terms, episode flags and both contact-flag blocks -- is ONE launch (`parc_sim_step`, same device code as the
stand-alone kernels, joint rotations never leave registers).  With `split_sim=True` (default) that launch is cut in two
where it starts to need the reference frame: the first half runs beside the query on the side branch, only reward terms
and episode flags remain behind it (5 launches, a shorter critical path); `split_sim=False` keeps the single launch (4).  With `fuse_tar_obs=True`
the target observation is written by the query kernel itself from the registers that hold the targets
(ParcTarObsSpec; 3 launches) -- an opt-in, because it measured slower than the separate launch, which runs beside
`parc_sim_step` on a parallel branch.

Every output buffer is allocated once, so the sequence can be captured in a CUDA graph (`capture()`), after
which a step is one graph launch.  The Isaac Gym classes themselves (simulation, resets, actors) are out of
scope; this is the piece of them that sits on the kinematic-query path.
"""
from __future__ import annotations

from typing import Optional, Sequence

import torch

from ... import ops


class TrackerStep:
    def __init__(self, mlib, terrain, num_envs: int, timestep: float, tar_obs_steps: Sequence[float],
                 key_body_ids: Sequence[int], ray_xy_points: torch.Tensor, *, joint_err_w: torch.Tensor,
                 pose_termination_dist: torch.Tensor, contact_body_ids: Sequence[int] = (), global_obs: bool = False,
                 root_height_obs: bool = False, track_root: bool = True, track_root_h: bool = True,
                 pose_termination: bool = True, enable_early_termination: bool = True,
                 termination_height: float = 0.15, episode_length: float = 10.0,
                 root_pos_termination_dist: float = 0.6, root_rot_termination_angle: float = 1.309,
                 min_obs_h: float = -3.0, max_obs_h: float = 3.0, fused: bool = True, fuse_tar_obs: bool = False,
                 query_variant: int = 0, split_sim: bool = True):
        dev = mlib._device if hasattr(mlib, "_device") else ray_xy_points.device
        self.device = torch.device(dev)
        self.mlib, self.kcm, self.terrain = mlib, mlib._kin_char_model, terrain
        self.n = int(num_envs)
        J, D = self.kcm.get_num_joints(), self.kcm.get_dof_size()
        self.key_body_ids = torch.as_tensor(list(key_body_ids), dtype=torch.int32, device=self.device)
        self.contact_body_ids = tuple(int(i) for i in contact_body_ids)
        self.ray_xy_points = ray_xy_points.to(self.device, torch.float32).contiguous()
        self.cfg = dict(global_obs=global_obs, root_height_obs=root_height_obs, track_root=track_root,
                        track_root_h=track_root_h, pose_termination=pose_termination,
                        enable_early_termination=enable_early_termination, termination_height=termination_height,
                        episode_length=episode_length, root_pos_termination_dist=root_pos_termination_dist,
                        root_rot_termination_angle=root_rot_termination_angle, min_obs_h=min_obs_h, max_obs_h=max_obs_h)
        self.joint_err_w = joint_err_w.to(self.device, torch.float32).contiguous()
        # per-DoF weights from the per-joint ones (ig_parkour_env.py:1573-1591)
        self.dof_err_w = torch.zeros(D, dtype=torch.float32, device=self.device)
        for j in range(1, J):
            dim = self.kcm.get_joint_dof_dim(j)
            if dim > 0:
                i = self.kcm.get_joint_dof_idx(j)
                self.dof_err_w[i:i + dim] = self.joint_err_w[j - 1]
        self.pose_termination_dist = pose_termination_dist.to(self.device, torch.float32).contiguous()
        # inputs the caller updates in place between steps
        self.motion_ids = torch.zeros(self.n, dtype=torch.int64, device=self.device)
        self.motion_times = torch.zeros(self.n, dtype=torch.float32, device=self.device)
        self.motion_xy_offset = torch.zeros(self.n, 2, dtype=torch.float32, device=self.device)
        steps = torch.as_tensor(list(tar_obs_steps), dtype=torch.float32)
        self.S = int(steps.shape[0])
        # fetch_tar_obs_data forms timestep * tar_obs_steps in fp32 (mgdm_dm_util.py:289); step 0 = the reference frame
        self.time_offsets = torch.cat([torch.zeros(1), timestep * steps]).to(self.device)
        # split_sim=True (with fused): the simulated character's launch is cut where it starts to need the reference
        # frame -- DoF conversion + observation run BESIDE the query on the side branch, reward terms + episode flags
        # after it -- which takes ~40 % of that launch off the step's critical path
        self.split_sim = bool(split_sim) and bool(fused)
        self.query_variant = int(query_variant)          # tuning: instantiation of the query kernel (0 = by batch size)
        self._plan = mlib.make_query_plan(self.motion_ids, self.motion_times, want_fk=True,
                                          time_offsets=self.time_offsets, root_xy_offset=self.motion_xy_offset,
                                          variant=self.query_variant)
        K = int(self.key_body_ids.shape[0])
        self.char_w = (1 if root_height_obs else 0) + 12 + 6 * (J - 1) + D + 3 * K
        self.tar_w = 9 + 6 * (J - 1) + 3 * K
        # the policy-observation row, written in place by the operators (no concatenation pass):
        # char_obs | tar_obs | tar_contacts | [char_contacts] | ray heightmap   (ig_parkour_env.py:1178-1215)
        self.P = int(self.ray_xy_points.shape[0])
        self._contacts = bool(getattr(mlib, "_contact_info", False))
        self._obs_bufs = {}
        self._graph = None
        # fused=True: the simulated character's share of the step (DoF conversion, observation block, reward, done,
        # contact blocks) is ONE launch (parc_sim_step) instead of four launches and two copies
        self.fused = bool(fused)
        # fuse_tar_obs=True (with fused): the query kernel itself writes the target observation of steps 1..S from its
        # registers (ParcTarObsSpec) -- 3 launches per step, no re-read of the target frames.  Off by default: measured
        # slower than the separate launch (58.9 vs 55.1 us per 4096-env step), which overlaps parc_sim_step.
        self.fuse_tar_obs = bool(fuse_tar_obs) and self.fused
        self._sim_plan, self._sim_key = None, None
        self._side = torch.cuda.Stream(self.device)
        self._query_done = torch.cuda.Event()
        D = self.kcm.get_dof_size()
        self._reward = torch.empty(self.n, 5, dtype=torch.float32, device=self.device)
        self._done = torch.empty(self.n, dtype=torch.int32, device=self.device)
        self._joint_rot = torch.empty(self.n, J - 1, 4, dtype=torch.float32, device=self.device)

    def _obs_layout(self, with_char_contacts: bool):
        J = self.kcm.get_num_joints()
        cols, c = {}, 0
        for name, w in (("char", self.char_w), ("tar", self.S * self.tar_w),
                        ("tar_contacts", self.S * J if self._contacts else 0),
                        ("char_contacts", J if (self._contacts and with_char_contacts) else 0), ("ray", self.P)):
            cols[name] = (c, c + w)
            c += w
        key = bool(with_char_contacts)
        if key not in self._obs_bufs:
            self._obs_bufs[key] = torch.empty(self.n, c, dtype=torch.float32, device=self.device)
        return self._obs_bufs[key], cols

    # ---- views of the query output: [n, S+1, ...] rows, step 0 = reference frame, 1.. = targets -------------
    def _views(self, out):
        n, S1 = self.n, self.S + 1
        v = {k: t.view(n, S1, *t.shape[1:]) for k, t in out.items()}
        ref = {k: t[:, 0] for k, t in v.items()}
        tar = {k: t[:, 1:] for k, t in v.items()}
        return ref, tar

    def step(self, root_pos, root_rot, root_vel, root_ang_vel, dof_pos, dof_vel, body_pos, contact_forces, time_buf,
             env_offsets, char_contacts: Optional[torch.Tensor] = None) -> dict:
        """Simulator state in (all [n, ...] fp32 CUDA, read in place) -> dict with
        obs [n, W] (char_obs | tar_obs | tar_contacts | [char_contacts] | ray heightmap, the order of
        ig_parkour_env.py:1178-1215), reward_terms [n,5], done [n] int32, and the reference frame views
        (ref_root_pos, ref_root_rot, ref_root_vel, ref_root_ang_vel, ref_joint_rot, ref_dof_vel, ref_body_pos,
        ref_contacts)."""
        c = self.cfg
        obs, col = self._obs_layout(char_contacts is not None)
        blk = lambda name: obs[:, col[name][0]:col[name][1]]
        if self.fused:
            ref, tar = self._views(self._plan.out)           # buffers are fixed; the query is launched inside
            return self._step_fused(ref, tar, obs, blk, root_pos, root_rot, root_vel, root_ang_vel, dof_pos, dof_vel,
                                    body_pos, contact_forces, time_buf, env_offsets, char_contacts)
        ref, tar = self._views(self._plan.launch())
        joint_rot = self.kcm.dof_to_rot(dof_pos)
        # ray heightmap around the SIMULATED character (ig_parkour_env.py:636-646, mgdm_dm_util.py:158-179)
        # heading = calc_heading(root_rot) and the env-local -> terrain shift are taken inside the launch
        ray_hfs = ops.hf_obs(self.terrain.hf_desc(), self.ray_xy_points, root_pos, None, relative=True,
                             min_h=c["min_obs_h"], max_h=c["max_obs_h"], root_rot=root_rot, root_offset=env_offsets,
                             out=blk("ray"))
        char_obs = ops.char_obs(root_pos, root_rot, root_vel, root_ang_vel, joint_rot, dof_vel, body_pos,
                                c["global_obs"], c["root_height_obs"], key_body_ids=self.key_body_ids, out=blk("char"))
        tar_obs = ops.tar_obs(root_pos, root_rot, tar["root_pos"], tar["root_rot"], tar["joint_rot"], tar["body_pos"],
                              c["global_obs"], False, key_body_ids=self.key_body_ids, out=blk("tar"))
        if self._contacts:
            J = self.kcm.get_num_joints()
            blk("tar_contacts").view(self.n, self.S, J).copy_(tar["contacts"])
            if char_contacts is not None:
                blk("char_contacts").copy_(char_contacts)
        reward_terms = ops.deepmimic_reward(
            (root_pos, root_rot, root_vel, root_ang_vel, joint_rot, dof_vel, body_pos),
            (ref["root_pos"], ref["root_rot"], ref["root_vel"], ref["root_ang_vel"], ref["joint_rot"], ref["dof_vel"],
             ref["body_pos"]), self.joint_err_w, self.dof_err_w, c["track_root_h"], c["track_root"],
            key_body_ids=self.key_body_ids)
        done = ops.done_flags(time_buf, c["episode_length"], root_rot, body_pos, ref["root_rot"], ref["body_pos"],
                              contact_forces, self.contact_body_ids, c["pose_termination"], self.pose_termination_dist,
                              c["enable_early_termination"], c["track_root"], c["root_pos_termination_dist"],
                              c["root_rot_termination_angle"], hf=self.terrain.hf_desc(), env_offsets=env_offsets,
                              termination_height=c["termination_height"])
        res = dict(obs=obs, reward_terms=reward_terms, done=done, char_obs=char_obs,
                   tar_obs=tar_obs.view(self.n, self.S, self.tar_w), ray_hfs=ray_hfs)
        res.update({"ref_" + k: t for k, t in ref.items()})
        return res

    def _step_fused(self, ref, tar, obs, blk, root_pos, root_rot, root_vel, root_ang_vel, dof_pos, dof_vel, body_pos,
                    contact_forces, time_buf, env_offsets, char_contacts):
        """4 launches: query (already done by the caller), ray heightmap, target observation, parc_sim_step -- each a
        prebuilt argument list over the caller's (persistent) state tensors, rebuilt only if a tensor is replaced."""
        c = self.cfg
        state = (root_pos, root_rot, root_vel, root_ang_vel, dof_pos, dof_vel, body_pos, contact_forces, time_buf,
                 env_offsets, char_contacts)
        key = tuple(None if t is None else (t.data_ptr(), tuple(t.shape), tuple(t.stride())) for t in state) \
            + (obs.data_ptr(),)
        if self._sim_plan is None or key != self._sim_key:
            ray = ops.hf_obs(self.terrain.hf_desc(), self.ray_xy_points, root_pos, None, relative=True,
                             min_h=c["min_obs_h"], max_h=c["max_obs_h"], root_rot=root_rot, root_offset=env_offsets,
                             out=blk("ray"), plan=True)
            if self.fuse_tar_obs:
                # same buffers as self._plan (its `out` dict is reused), plus the target-observation block
                tarp = self.mlib.make_query_plan(
                    self.motion_ids, self.motion_times, want_fk=True, time_offsets=self.time_offsets,
                    root_xy_offset=self.motion_xy_offset, out=self._plan.out, variant=self.query_variant,
                    tar_obs=dict(sim_root_pos=root_pos, sim_root_rot=root_rot, key_body_ids=self.key_body_ids,
                                 out=blk("tar"), global_obs=c["global_obs"], global_tar_root_h=False))
            else:
                tarp = ops.tar_obs(root_pos, root_rot, tar["root_pos"], tar["root_rot"], tar["joint_rot"], tar["body_pos"],
                                   c["global_obs"], False, key_body_ids=self.key_body_ids, out=blk("tar"), plan=True)
            sim = dict(root_pos=root_pos, root_rot=root_rot, root_vel=root_vel, root_ang_vel=root_ang_vel,
                       dof_pos=dof_pos, dof_vel=dof_vel, body_pos=body_pos, contact_force=contact_forces, time=time_buf,
                       env_offsets=env_offsets, char_contacts=char_contacts)
            refd = dict(ref)
            out = dict(char_obs=blk("char"), reward=self._reward, done=self._done, joint_rot=self._joint_rot)
            if self._contacts:
                refd["tar_contacts"] = tar["contacts"]
                out["tar_contacts"] = blk("tar_contacts")
                if char_contacts is not None:
                    out["char_contacts"] = blk("char_contacts")
            cfg = dict(c, pose_termination_dist=self.pose_termination_dist)
            mk = lambda phase: ops.SimStepPlan(self.kcm.c_model(), sim, refd, self.key_body_ids, self.joint_err_w,
                                               self.dof_err_w, self.terrain.hf_desc(), out, cfg=cfg,
                                               contact_body_ids=self.contact_body_ids, phase=phase)
            simp = (mk(1), mk(2)) if self.split_sim else mk(0)
            self._sim_plan, self._sim_key = (ray, tarp, simp), key
            res = dict(obs=obs, reward_terms=self._reward, done=self._done, char_obs=blk("char"),
                       tar_obs=blk("tar").view(self.n, self.S, self.tar_w), ray_hfs=blk("ray"), joint_rot=self._joint_rot)
            res.update({"ref_" + k: t for k, t in ref.items()})
            self._fused_result = res
        # Two branches: the ray heightmap needs only the simulator state, and once the query has finished the target
        # observation and the character's step are independent of each other.  (Under CUDA-graph capture the side
        # stream becomes a parallel branch of the graph.)
        ray, tarp, simp = self._sim_plan
        cur, side = torch.cuda.current_stream(self.device), self._side
        side.wait_stream(cur)
        ray.launch(side.cuda_stream)
        pre, post = simp if self.split_sim else (None, simp)
        if pre is not None:
            pre.launch(side.cuda_stream)                 # beside the query: needs the simulator state only
        if self.fuse_tar_obs:
            tarp.launch(cur.cuda_stream)                 # query + FK + target observation in one launch
            if pre is not None:
                cur.wait_stream(side)
            post.launch(cur.cuda_stream)
            cur.wait_stream(side)
            return self._fused_result
        self._plan.launch(cur.cuda_stream)
        self._query_done.record(cur)
        side.wait_event(self._query_done)
        post.launch(side.cuda_stream)                    # behind `pre` on the same stream, and behind the query
        tarp.launch(cur.cuda_stream)
        cur.wait_stream(side)
        return self._fused_result

    def capture(self, *state) -> "torch.cuda.CUDAGraph":
        """Capture `step(*state)` over the given (persistent) simulator tensors into a CUDA graph; afterwards
        `replay()` re-runs the whole step as one graph launch and `self.result` holds the same output tensors."""
        side = torch.cuda.Stream(self.device)
        side.wait_stream(torch.cuda.current_stream(self.device))
        with torch.cuda.stream(side):
            for _ in range(2):
                self.step(*state)
        torch.cuda.current_stream(self.device).wait_stream(side)
        self._graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self._graph):
            self.result = self.step(*state)
        return self._graph

    def replay(self) -> dict:
        self._graph.replay()
        return self.result
