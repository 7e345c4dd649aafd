"""RocketVecEnv: a stable-baselines3-compatible vectorised environment whose `step` is ONE kernel launch.

Replaces `DummyVecEnv([lambda: Monitor(TimeLimit(Rocket6DOF(**cfg)))])` (main_6DOF.py:21-22, :71; main.py:65-70, :105):
TimeLimit truncation, Monitor's per-episode return/length and DummyVecEnv's auto-reset + `terminal_observation`
are fused into the step kernel.  API = SB3 1.6 `VecEnv` (reset / step_async / step_wait / step / close / seed /
get_attr / set_attr / env_method / env_is_wrapped / render / get_images) plus a tensor-native fast path:

    obs, reward, done = venv.step_tensor(actions_cuda)      # no host sync
    obs, reward, done = venv.step_host(actions_pinned)      # pinned host buffers in/out, one H2D + one D2H per step
    h = venv.step_host_async(actions_pinned); ...; obs, reward, done = h.wait()     # the same, pipelined over `depth` slots
"""
import ctypes as C
import time

import numpy as np
import torch

from . import spaces
from .config import DEFAULT_DISCRETE_TABLE, EnvSpec
from . import _native as N
from .engine import RocketBatch, _ptr

try:  # pragma: no cover - SB3 is absent from the build image
    from stable_baselines3.common.vec_env.base_vec_env import VecEnv as _VecEnvBase
except Exception:  # noqa: BLE001
    class _VecEnvBase:
        def __init__(self, num_envs, observation_space, action_space):
            self.num_envs, self.observation_space, self.action_space = num_envs, observation_space, action_space

        def step(self, actions):
            self.step_async(actions)
            return self.step_wait()

        @property
        def unwrapped(self):
            return self


class _LazyInfos(list):
    """list[dict] of length num_envs built without O(N) Python work: all non-terminal envs share one empty dict."""


class RocketVecEnv(_VecEnvBase):
    def __init__(self, num_envs, dof=6, max_episode_steps=None, device="cuda:0", discrete_actions=False, discrete_table=None,
                 reward_annealing=False, math_mode="fast", env_offset=0, trace_indices=None, trace_capacity=4096, precision="f64",
                 gaudet_obs=False, **env_kwargs):
        table = (discrete_table or DEFAULT_DISCRETE_TABLE) if discrete_actions else None
        self.spec = EnvSpec(dof, max_episode_steps=max_episode_steps or 0, reward_mode=1 if reward_annealing else 0,
                            discrete_table=table, obs_mode="gaudet" if gaudet_obs else "state", **env_kwargs)
        self.batch = RocketBatch(self.spec, num_envs, device=device, math_mode=math_mode, env_offset=env_offset,
                                 debug_outputs=trace_indices is not None, precision=precision)
        self.trace = TrajectoryTrace(self.batch, trace_indices, trace_capacity) if trace_indices is not None else None
        obs_space = spaces.Box(low=-1, high=1, shape=(self.spec.nobs,))
        act_space = spaces.Discrete(len(table)) if discrete_actions else spaces.Box(low=-1, high=1, shape=(self.spec.nact,))
        super().__init__(num_envs, obs_space, act_space)
        self.discrete = bool(discrete_actions)
        self.device = self.batch.device
        n, no = num_envs, self.spec.nobs
        # pinned host mirrors for the numpy-facing API (one packed D2H per step)
        self._h_actions = torch.zeros((n,) if self.discrete else (n, self.spec.nact),
                                      dtype=torch.int64 if self.discrete else torch.float32).pin_memory()
        self._d_actions = torch.zeros_like(self._h_actions, device=self.device)
        self._h_pack = torch.zeros_like(self.batch.out_pack, device="cpu").pin_memory()      # obs | reward | done, one D2H per step
        self._h_obs = self._h_pack[:n * no * 4].view(torch.float32).view(n, no)
        self._h_rew = self._h_pack[n * no * 4:n * no * 4 + n * 4].view(torch.float32)
        self._h_done = self._h_pack[n * no * 4 + n * 4:]
        self._ot_host = None          # RlrStepOut with obs | reward | done in mapped pinned host memory (built on first use)
        self._pipe = None             # host-buffer pipeline (rlr_pipe_*), built by the first step_host_async
        self._fin_pack = None         # device staging of the finished episodes' info fields (step_wait)
        self._t_start = time.time()
        self._pending = False
        self.metadata = {"render.modes": []}

    # ---------------------------------------------------------------- tensor-native fast path
    def reset_tensor(self):
        obs = self.batch.reset()
        if self.trace is not None:
            self.trace.start()
        return obs

    def step_tensor(self, actions, chained=False):
        """Tensor-native step (no host synchronisation).  chained=True: open-loop chaining of consecutive steps (6DOF, f64; see
        RocketBatch.step_tensor) -- for action sources that do not need the previous step's results; not with a trajectory trace."""
        if chained and self.trace is not None:
            raise ValueError("chained stepping cannot be combined with a trajectory trace (the trace reads every step's outputs)")
        out = self.batch.step_tensor(actions, chained=chained)
        if self.trace is not None:
            self.trace.record(actions)
        return out

    def rollout_tensor(self, actions, obs_buf=None, reward_buf=None, done_buf=None, trunc_buf=None):
        return self.batch.rollout(actions, obs_buf, reward_buf, done_buf, trunc_buf)

    def step_host(self, actions_host, zero_copy=None):
        """actions: pinned host tensor (or numpy array).  Returns pinned host tensors (obs, reward, done) -- views of persistent
        buffers.  The host -> device transfer of the actions and the device -> host transfer of the results are inside this call.

        zero_copy (default: on for continuous actions without a trajectory trace): the kernel reads the actions from, and writes
        obs | reward | done to, MAPPED pinned host memory, so the 4.8 MB of PCIe traffic per step of 65 536 envs overlaps the
        integration instead of following it (measured 146 -> 108 us per step).  Otherwise: one H2D copy, the step, one packed D2H."""
        if not torch.is_tensor(actions_host):
            actions_host = torch.as_tensor(np.ascontiguousarray(actions_host))
        b = self.batch
        if zero_copy is None:
            zero_copy = not self.discrete and self.trace is None
        if zero_copy:
            a = actions_host.reshape(self._h_actions.shape)
            if not (a.dtype == torch.float32 and a.is_contiguous() and a.is_pinned()):
                self._h_actions.copy_(a)
                a = self._h_actions
            if self._ot_host is None:
                self._ot_host = b.step_out_struct(self._h_obs, self._h_rew, self._h_done)
            b.step_raw(a.data_ptr(), self._ot_host)
            torch.cuda.current_stream(self.device).synchronize()
            return self._h_obs, self._h_rew, self._h_done
        self._d_actions.copy_(actions_host.reshape(self._d_actions.shape), non_blocking=True)
        self.step_tensor(self._d_actions)
        self._h_pack.copy_(b.out_pack, non_blocking=True)
        torch.cuda.current_stream(self.device).synchronize()
        return self._h_obs, self._h_rew, self._h_done

    # ---------------------------------------------------------------- pipelined host path
    def step_host_async(self, actions_host, depth=2):
        """Non-blocking `step_host`: enqueues H2D(actions) -> step kernel -> D2H(obs | reward | done) through the C ABI's host
        pipeline (rlr_pipe_submit, one ctypes call) and returns a handle; `handle.wait()` blocks until this step's results are in
        pinned host memory and returns (obs, reward, done) views of the slot's buffers -- valid until the slot is reused `depth`
        submits later.  Consecutive submits keep their order on the device (step t + 1 integrates from step t's state) while the
        PCIe transfers of step t overlap the integration of step t + 1: for open-loop action sources (random / scripted / action
        repeat) or two environment groups stepped alternately.  `actions_host` must be pinned and stay untouched until wait()."""
        if not torch.is_tensor(actions_host):
            actions_host = torch.as_tensor(np.ascontiguousarray(actions_host))
        a = actions_host.reshape(self._h_actions.shape)
        if not (a.dtype == self._h_actions.dtype and a.is_contiguous() and a.is_pinned()):
            raise ValueError("step_host_async needs a contiguous pinned host tensor of the action dtype (it is read asynchronously)")
        b = self.batch
        if self._pipe is None:
            self._pipe = _HostPipe(self, depth)
        return self._pipe.submit(a)

    def h2d_bytes_per_step(self):
        return self._h_actions.numel() * self._h_actions.element_size()

    def d2h_bytes_per_step(self):
        return self._h_pack.numel()

    # ---------------------------------------------------------------- SB3 VecEnv API
    def reset(self):
        return self.reset_tensor().cpu().numpy()

    def step_async(self, actions):
        a = np.asarray(actions)
        if self.discrete:
            self._h_actions.copy_(torch.as_tensor(a.astype(np.int64).reshape(-1)))
        else:
            self._h_actions.copy_(torch.as_tensor(a.astype(np.float32).reshape(self.num_envs, -1)))
        self._pending = True

    def _fin_buffers(self):
        """Device queue of the step's finished episodes (RlrStepOut.fin_count / fin_records) + its pinned host mirror.  One D2H copy
        of the counter and the first `head` records follows the step; only a step that finishes more episodes than that needs a
        second copy."""
        if self._fin_pack is None:
            b, n = self.batch, self.num_envs
            rec = N.FIN_HEADER + self.spec.nobs
            head = min(n, max(64, n // 16))
            d = torch.zeros(1 + n * rec, dtype=torch.float64, device=self.device)          # [count as u32 in word 0 | records]
            h = torch.zeros(1 + head * rec, dtype=torch.float64).pin_memory()
            out = b.step_out_struct(self._h_obs, self._h_rew, self._h_done)                  # obs | reward | done in mapped pinned memory
            out.fin_count = d.data_ptr()
            out.fin_records = d.data_ptr() + 8
            self._fin_pack = (d, h, out, rec, head)
        return self._fin_pack

    def step_wait(self):
        assert self._pending, "step_wait() without step_async()"
        self._pending = False
        b, n, no = self.batch, self.num_envs, self.spec.nobs
        d_fin, h_fin, out, rec, head = self._fin_buffers()
        with torch.cuda.device(self.device):
            d_fin[:1].zero_()
            kind = N.ACT_DISCRETE if self.discrete else N.ACT_NORMALIZED
            if self.trace is None:
                b.step_raw(self._h_actions.data_ptr(), out, action_kind=kind)                # the kernel reads / writes mapped host memory
            else:                                                                            # the trace reads the batch's own outputs
                self._d_actions.copy_(self._h_actions, non_blocking=True)
                out_t = b._ot.__class__.from_buffer_copy(b._ot)
                out_t.fin_count, out_t.fin_records = out.fin_count, out.fin_records
                b.step_raw(self._d_actions.data_ptr(), out_t, action_kind=kind)
                self.trace.record(self._d_actions)
                self._h_pack.copy_(b.out_pack, non_blocking=True)
            h_fin.copy_(d_fin[:h_fin.numel()], non_blocking=True)
            torch.cuda.current_stream(self.device).synchronize()
        obs_np, rew_np, done_np = self._h_obs.numpy().copy(), self._h_rew.numpy().copy(), self._h_done.numpy().astype(bool)
        empty = {}                    # one fresh dict per step, shared by this step's non-terminal environments only
        infos = _LazyInfos([empty] * n)
        count = int(h_fin[:1].view(torch.int32)[0])
        if count:
            if count > head:          # rare: more finished episodes than the head copy carries
                recs = d_fin[1:1 + count * rec].cpu().numpy().reshape(count, rec)
            else:
                recs = h_fin[1:1 + count * rec].numpy().reshape(count, rec)
            term = recs[:, N.FIN_HEADER:].astype(np.float32)       # exact: the values were float32
            now = round(time.time() - self._t_start, 6)
            for j in range(count):
                r = recs[j]
                fl = int(r[4])
                infos[int(r[0])] = {"terminal_observation": term[j], "TimeLimit.truncated": bool(r[3]),
                                    "episode": {"r": round(float(r[1]), 6), "l": int(r[2]), "t": now},
                                    "bounds_violation": bool(fl & 1), "landed": bool(fl & 2), "is_done": True}
        return obs_np, rew_np, done_np, infos

    def close(self):
        pass

    def seed(self, seed=None):
        seed = 0 if seed is None else int(seed)
        self.batch.reseed(seed)            # re-keys the reset RNG and restarts the episode numbering: seed(s) + reset() reproduces
        return [seed + i for i in range(self.num_envs)]

    def _indices(self, indices):
        if indices is None:
            return range(self.num_envs)
        if isinstance(indices, int):
            return [indices]
        return indices

    def get_attr(self, attr_name, indices=None):
        src = self if hasattr(self, attr_name) else self.spec
        return [getattr(src, attr_name) for _ in self._indices(indices)]

    def set_attr(self, attr_name, value, indices=None):
        setattr(self, attr_name, value)

    def env_method(self, method_name, *args, indices=None, **kwargs):
        return [getattr(self, method_name)(*args, **kwargs) for _ in self._indices(indices)]

    def env_is_wrapped(self, wrapper_class, indices=None):
        return [False for _ in self._indices(indices)]

    def get_images(self):
        raise NotImplementedError("rendering is outside the hot path this package replaces")

    def render(self, mode="human"):
        raise NotImplementedError("rendering is outside the hot path this package replaces")

    # reference-side names reached through get_attr / wrappers
    @property
    def reward_coefficients(self):
        return self.spec.reward_coeff

    def set_reward_annealing(self, on=True):
        """Second training phase of main_6DOF.py:95-103 / main.py:130-175 (RewardAnnealing) without rebuilding the envs."""
        self.spec = self.spec.with_(reward_mode=1 if on else 0)
        self.batch.spec = self.spec

    def episode_stats(self, reset=False):
        return self.batch.read_stats(reset=reset)


class _HostPipe:
    """`depth` slots of (device actions, device obs|reward|done pack, pinned host pack) around the C ABI's rlr_pipe_*."""

    def __init__(self, venv, depth):
        b = venv.batch
        self.venv, self.depth, self.turn = venv, int(depth), 0
        self.lib = b.lib
        n, no = b.n, b.spec.nobs
        self.kind = N.ACT_DISCRETE if venv.discrete else N.ACT_NORMALIZED
        self.handle = C.c_void_p()
        with torch.cuda.device(b.device):
            N.check(self.lib.rlr_pipe_create(C.byref(self.handle), self.depth))
        self.slots = []
        for _ in range(self.depth):
            d_act = torch.zeros_like(venv._h_actions, device=b.device)
            d_pack = torch.zeros_like(b.out_pack)
            h_pack = torch.zeros_like(b.out_pack, device="cpu").pin_memory()
            obs = d_pack[:n * no * 4].view(torch.float32).view(n, no)
            rew = d_pack[n * no * 4:n * no * 4 + n * 4].view(torch.float32)
            done = d_pack[n * no * 4 + n * 4:]
            out = b.step_out_struct(obs, rew, done)
            views = (h_pack[:n * no * 4].view(torch.float32).view(n, no), h_pack[n * no * 4:n * no * 4 + n * 4].view(torch.float32),
                     h_pack[n * no * 4 + n * 4:])
            self.slots.append((d_act, d_pack, h_pack, out, views))

    def submit(self, a):
        b = self.venv.batch
        s = self.turn % self.depth
        self.turn += 1
        d_act, d_pack, h_pack, out, views = self.slots[s]
        prev = torch.cuda.current_device()
        if prev != b._dev_index:
            torch.cuda.set_device(b._dev_index)
        b._chain_seq = 0              # (a step outside the batch's open-loop chain)
        rc = self.lib.rlr_pipe_submit(self.handle, s, b._params_ref, b.precision, b.math_mode, b._st, a.data_ptr(), d_act.data_ptr(),
                                      a.numel() * a.element_size(), self.kind, out, d_pack.data_ptr(), h_pack.data_ptr(), h_pack.numel(),
                                      b.n, b.env_offset, b.seed, torch.cuda.current_stream(b._dev_index).cuda_stream)
        if prev != b._dev_index:
            torch.cuda.set_device(prev)
        if rc:
            N.check(rc)
        b.launches += 1
        return HostStepHandle(self, s, views, a)

    def __del__(self):
        try:
            self.lib.rlr_pipe_destroy(self.handle)
        except Exception:  # noqa: BLE001 - interpreter shutdown
            pass


class HostStepHandle:
    def __init__(self, pipe, slot, views, actions):
        self._pipe, self._slot, self._views, self._actions = pipe, slot, views, actions      # (keeps the pinned actions alive)

    def wait(self):
        rc = self._pipe.lib.rlr_pipe_wait(self._pipe.handle, self._slot)
        if rc:
            N.check(rc)
        return self._views


class TrajectoryTrace:
    """Device-side trajectory history of a few selected environments of a batch -- the vectorised counterpart of what the
    reference keeps per environment in `SIM.states / SIM.actions / SIM.times` and `vtarg_history` (simulator.py:28-31, :201-206,
    :246-248; rocket_env.py:245, :1012) and reads back through `states_to_dataframe`, `actions_to_dataframe`, `vtarg_to_dataframe`
    and `used_mass` (rocket_env.py:1016-1034) for EpisodeAnalyzer / RecordVideoFigure (wrappers.py:189-235).

    Nothing is added to the step kernel's hot path beyond the optional `y_pre_reset` / `vtarg` outputs: after every step the m
    selected columns are gathered into ring buffers `[capacity, m, ...]` on the device (no host synchronisation); the host reads
    them only when an analysis asks for a dataframe.  Row 0 of an episode is its initial condition (action 0, as simulator.py:30).
    """

    def __init__(self, batch, indices, capacity=4096):
        self.batch = batch
        self.idx = torch.as_tensor(list(indices), dtype=torch.long, device=batch.device)
        if self.idx.numel() == 0 or int(self.idx.min()) < 0 or int(self.idx.max()) >= batch.n:
            raise ValueError("trace_indices must be non-empty environment indices of this batch")
        m, nd, na, dev = self.idx.numel(), batch.spec.ndim, batch.spec.nact, batch.device
        self.capacity = int(capacity)
        self.states = torch.zeros((self.capacity, m, nd), dtype=torch.float64, device=dev)       # (f32 states are widened on copy)
        self.actions = torch.zeros((self.capacity, m, na), dtype=torch.float32, device=dev)
        self.post = torch.zeros((self.capacity, m, nd), dtype=torch.float64, device=dev)      # state after auto-reset (next episode's IC)
        self.vtarg = torch.zeros((self.capacity, m, 4), dtype=torch.float64, device=dev)
        self.reward = torch.zeros((self.capacity, m), dtype=torch.float64, device=dev)
        self.step_in_episode = torch.zeros((self.capacity, m), dtype=torch.int32, device=dev)      # 0 = initial condition row
        self.episode = torch.zeros((self.capacity, m), dtype=torch.int32, device=dev)
        self.count = 0            # rows written so far (monotonic; row r lives at r % capacity)

    def _put(self, states, actions, vtarg, reward, k, episode):
        r = self.count % self.capacity
        self.states[r].copy_(states)
        self.post[r].copy_(self.batch.y[:, self.idx].T)
        self.actions[r].copy_(actions)
        self.vtarg[r].copy_(vtarg)
        self.reward[r].copy_(reward)
        self.step_in_episode[r].copy_(k)
        self.episode[r].copy_(episode)
        self.count += 1

    def start(self):
        """Called after a reset of the whole batch: row 0 = the initial conditions."""
        b, m = self.batch, self.idx.numel()
        self.count = 0
        z = torch.zeros((m,), dtype=torch.float64, device=b.device)
        self._put(b.y[:, self.idx].T, torch.zeros((m, b.spec.nact), device=b.device), torch.zeros((m, 4), dtype=torch.float64, device=b.device),
                  z, torch.zeros((m,), dtype=torch.int32, device=b.device), b.episode[self.idx])

    def record(self, actions):
        """Called after every step: the post-step (pre-reset) state, the action taken, vtarg / t_go, the reward, and the state after
        the fused auto-reset (`post`), which is the next episode's initial condition for the environments that just finished."""
        b, idx = self.batch, self.idx
        a = actions[idx].to(torch.float32).reshape(idx.numel(), -1)
        done = b.done[idx].bool()
        # k was already reset for finished environments: their last step index comes from fin_len
        k = torch.where(done, b.fin_len[idx], b.k[idx])
        ep = b.episode[idx] - done.to(torch.int32)
        self._put(b.y_pre_reset[:, idx].T, a, b.vtarg[:, idx].T, b.reward64[idx], k, ep)

    def to_host(self):
        """dict of numpy arrays in chronological order (the last `capacity` rows)."""
        n = min(self.count, self.capacity)
        order = [(self.count - n + j) % self.capacity for j in range(n)]
        sel = torch.as_tensor(order, dtype=torch.long, device=self.batch.device)
        return {name: getattr(self, name)[sel].cpu().numpy()
                for name in ("states", "post", "actions", "vtarg", "reward", "step_in_episode", "episode")}

    def last_episode(self, j=0):
        """Rows of the most recent (possibly running) episode of traced environment j: (states, actions, vtarg, times)."""
        h = self.to_host()
        ep = h["episode"][:, j]
        rows = np.nonzero(ep == ep[-1])[0]
        states, actions, vtarg, k = h["states"][rows, j], h["actions"][rows, j], h["vtarg"][rows, j], h["step_in_episode"][rows, j]
        if k[0] != 0 and rows[0] > 0:          # the episode began with an auto-reset: its IC is the `post` state of the previous row
            states = np.concatenate([h["post"][rows[0] - 1, j][None], states])
            actions = np.concatenate([np.zeros_like(actions[:1]), actions])
            vtarg = np.concatenate([np.zeros_like(vtarg[:1]), vtarg])
            k = np.concatenate([[0], k])
        t = k.astype(np.float64) * self.batch.spec.timestep
        return states, actions, vtarg, np.round(t, 3)

    def states_to_dataframe(self, j=0, names=None):
        import pandas as pd
        return pd.DataFrame(self.last_episode(j)[0], columns=names)

    def actions_to_dataframe(self, j=0, names=None):
        import pandas as pd
        return pd.DataFrame(self.last_episode(j)[1], columns=names)

    def vtarg_to_dataframe(self, j=0):
        import pandas as pd
        nv = 3 if self.batch.spec.dof == 6 else 2
        return pd.DataFrame(self.last_episode(j)[2][1:, :nv], columns=["v_x", "v_y", "v_z"][:nv])

    def used_mass(self, j=0):
        s = self.last_episode(j)[0]
        return s[0, -1] - s[-1, -1]
