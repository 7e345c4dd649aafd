"""RocketBatch: N environments resident on one GPU, stepped by the sm_100a kernels through the C ABI.

PyTorch owns every buffer (structure-of-arrays state, row-major obs/action tensors); the library gets
raw `data_ptr()`s and the current CUDA stream -- zero-copy, no hidden synchronisation.  This module is
the tensor-native API (`step_tensor`, `rollout`); the gym / SB3 facades in envs.py and vec_env.py sit on top.
"""
import ctypes as C

import torch

from . import _native as N
from .config import EnvSpec


def _ptr(t):
    return None if t is None else C.c_void_p(t.data_ptr())


class RocketBatch:
    def __init__(self, spec: EnvSpec, num_envs: int, device="cuda:0", seed=None, env_offset=0, math_mode="fast", debug_outputs=False,
                 precision="f64", stats=None):
        self.lib = N.load()                      # raises if the CUDA extension is not built
        if not torch.cuda.is_available():
            raise RuntimeError("rl_rocket_b200 needs a CUDA device (there is no CPU fallback)")
        self.spec = spec                         # (property: also caches the byref of the params POD)
        self.n = int(num_envs)
        self.device = torch.device(device)
        self.seed = int(spec.seed if seed is None else seed)
        self.env_offset = int(env_offset)
        self.math_mode = {"exact": N.MATH_EXACT, "fast": N.MATH_FAST}[math_mode]
        # precision of the integrated state: "f64" (reference parity <= 1e-9) or "f32" (6DOF, fast math; <= 1e-4, see rocket_f32.cuh)
        self.precision = {"f64": N.F64, "f32": N.F32}[precision]
        if self.precision == N.F32 and (spec.dof != 6 or self.math_mode != N.MATH_FAST):
            raise ValueError("precision='f32' is implemented for the 6DOF environment with math_mode='fast'")
        ydt = torch.float64 if self.precision == N.F64 else torch.float32
        n, nd, no, dev = self.n, spec.ndim, spec.nobs, self.device
        # persistent state (SoA)
        self.y = torch.zeros((nd, n), dtype=ydt, device=dev)
        self.k = torch.zeros(n, dtype=torch.int32, device=dev)
        self.episode = torch.zeros(n, dtype=torch.int32, device=dev)
        self.v0 = torch.zeros(n, dtype=torch.float32, device=dev)
        self.ep_ret = torch.zeros(n, dtype=torch.float64, device=dev)
        # per-step outputs.  obs | reward | done are views of ONE packed buffer so that the host-facing API needs a single D2H copy
        self.out_pack = torch.zeros(n * no * 4 + n * 4 + n, dtype=torch.uint8, device=dev)
        self.obs = self.out_pack[:n * no * 4].view(torch.float32).view(n, no)
        self.reward = self.out_pack[n * no * 4:n * no * 4 + n * 4].view(torch.float32)
        self.done = self.out_pack[n * no * 4 + n * 4:]
        self.reward64 = torch.zeros(n, dtype=torch.float64, device=dev)
        self.truncated = torch.zeros(n, dtype=torch.uint8, device=dev)
        self.terminal_obs = torch.zeros((n, no), dtype=torch.float32, device=dev)
        self.fin_ret = torch.zeros(n, dtype=torch.float64, device=dev)
        self.fin_len = torch.zeros(n, dtype=torch.int32, device=dev)
        self.flags = torch.zeros(n, dtype=torch.uint8, device=dev)
        # episode statistics record; several batches of one rank (groups of environments stepped alternately) may share one
        if stats is not None and (stats.dtype != torch.int64 or stats.numel() != N.NSTATS or not stats.is_cuda or not stats.is_contiguous()):
            raise ValueError(f"stats must be a contiguous CUDA int64[{N.NSTATS}] tensor on the batch's device")
        self.stats = torch.zeros(N.NSTATS, dtype=torch.int64, device=dev) if stats is None else stats
        self.debug = bool(debug_outputs)
        self.natt = torch.zeros(n, dtype=torch.int32, device=dev) if self.debug else None
        self.y_pre_reset = torch.zeros((nd, n), dtype=ydt, device=dev) if self.debug else None
        self.reward_terms = torch.zeros((6, n), dtype=torch.float64, device=dev) if self.debug else None
        self.vtarg = torch.zeros((4, n), dtype=torch.float64, device=dev) if self.debug else None
        self.launches = 0
        # open-loop chaining (step_tensor(..., chained=True) -> rlr_step_chained): one sequence word per 64-environment block
        self._block_seq = torch.zeros(max(1, (n + 63) // 64), dtype=torch.int32, device=dev)
        self._chain_seq = 0          # seq_publish of the last chained launch; 0 = the chain is broken (the next chained step waits for everything)
        self._started = False        # True once reset() has drawn initial conditions: later explicit resets move to the NEXT episode key
        # the buffers above are persistent, so the pointer structs handed to the C ABI are built once
        self._st = N.RlrState(_ptr(self.y), _ptr(self.k), _ptr(self.episode), _ptr(self.v0), _ptr(self.ep_ret))
        self._ot = N.RlrStepOut(_ptr(self.obs), _ptr(self.reward), _ptr(self.reward64), _ptr(self.done), _ptr(self.truncated),
                                _ptr(self.terminal_obs), _ptr(self.fin_ret), _ptr(self.fin_len), _ptr(self.flags), _ptr(self.natt),
                                _ptr(self.y_pre_reset), _ptr(self.reward_terms), _ptr(self.vtarg), _ptr(self.stats))
        self._dev_index = self.device.index if self.device.index is not None else torch.cuda.current_device()
        self._act_shape = (n, spec.nact)
        self._step_fn = self.lib.rlr_step

    # -- raw structs -------------------------------------------------------------------------
    @property
    def spec(self):
        return self._spec

    @spec.setter
    def spec(self, value):
        self._spec = value
        self._params_ref = C.byref(value.params)

    def _state(self):
        return self._st

    def _out(self):
        return self._ot

    def _stream(self):
        return torch.cuda.current_stream(self._dev_index).cuda_stream

    def _check_actions(self, actions, leading=()):
        spec = self.spec
        if actions.device != self.device:
            raise ValueError("actions must live on the environment's device")
        if actions.dtype == torch.int64:
            kind, shape = N.ACT_DISCRETE, (*leading, self.n)
        elif actions.dtype == torch.float32:
            kind, shape = None, (*leading, self.n, spec.nact)
        else:
            raise TypeError("actions must be float32 (continuous) or int64 (discrete)")
        if tuple(actions.shape) != shape and not (spec.nact == 1 and tuple(actions.shape) == (*leading, self.n)):
            raise ValueError(f"actions must have shape {shape}, got {tuple(actions.shape)}")
        if not actions.is_contiguous():
            actions = actions.contiguous()
        return actions, kind

    # -- API ----------------------------------------------------------------------------------
    def reseed(self, seed):
        """gym's `seed()`: re-key the reset RNG and restart the episode numbering, so that seed(s) + reset() reproduces."""
        self.seed = int(seed)
        self.episode.zero_()
        self._started = False
        self._chain_seq = 0

    def reset(self, mask=None, ic=None):
        """(Re)start environments; returns the obs tensor [n, nobs] (device).  `ic`: optional f32 [n, ndim] initial conditions.

        Initial conditions are Philox draws keyed by (seed, global env index, episode index).  The first reset after construction
        (or reseed) uses each environment's current episode index; every later explicit reset first advances the index of the
        environments it restarts, so that -- like `init_space.sample()` in the reference (rocket_env.py:142, :672) -- it draws
        fresh initial conditions instead of replaying the running episode's."""
        with torch.cuda.device(self.device):
            if ic is not None:
                ic = torch.as_tensor(ic, dtype=torch.float32, device=self.device).contiguous()
                assert tuple(ic.shape) == (self.n, self.spec.ndim)
            if mask is not None:
                mask = torch.as_tensor(mask, device=self.device).to(torch.uint8).contiguous()
            if self._started:
                if mask is None:
                    self.episode += 1
                else:
                    self.episode += (mask != 0).to(torch.int32)
            self._started = True
            self._chain_seq = 0
            N.check(self.lib.rlr_reset(self._params_ref, self.precision, self._state(), _ptr(mask), _ptr(ic), _ptr(self.obs),
                                       self.n, self.env_offset, self.seed, self._stream()))
            self.launches += 1
        return self.obs

    def step_tensor(self, actions, physical=False, ic_inject=None, chained=False):
        """One step for all envs.  actions: cuda f32 [n, nact] (or i64 [n] discrete).  Returns (obs, reward, done) device tensors
        (views of persistent buffers, overwritten by the next call).

        chained=True (6DOF, f64): open-loop chaining -- consecutive chained steps of this batch are ordered block by block instead of
        launch by launch, so the tail of one step overlaps the body of the next (rlr_step_chained).  Only for action sources that do
        not depend on the previous step's results (random / scripted / repeated actions, several batches stepped in turn): `actions`
        must not be produced by work still pending on the stream, and nothing else may touch this batch's state or outputs on the
        stream between two chained steps.  Any other method of the batch ends the chain; the results are identical either way."""
        if chained:
            return self._step_chained(actions, physical)
        self._chain_seq = 0
        if (actions.dtype is torch.float32 and tuple(actions.shape) == self._act_shape and actions.is_contiguous()
                and actions.device == self.device and ic_inject is None):
            # hot path: no per-call allocation, one ctypes call (the caller's current device is restored when it was another GPU)
            prev = torch.cuda.current_device()
            if prev != self._dev_index:
                torch.cuda.set_device(self._dev_index)
            rc = self._step_fn(self._params_ref, self.precision, self.math_mode, self._st, actions.data_ptr(),
                               N.ACT_PHYSICAL if physical else N.ACT_NORMALIZED, None, self._ot, self.n, self.env_offset, self.seed,
                               torch.cuda.current_stream(self._dev_index).cuda_stream)
            if prev != self._dev_index:
                torch.cuda.set_device(prev)
            if rc:
                N.check(rc)
            self.launches += 1
            return self.obs, self.reward, self.done
        actions, kind = self._check_actions(actions)
        if kind is None:
            kind = N.ACT_PHYSICAL if physical else N.ACT_NORMALIZED
        with torch.cuda.device(self.device):
            if ic_inject is not None:
                ic_inject = torch.as_tensor(ic_inject, dtype=torch.float32, device=self.device).contiguous()
            N.check(self.lib.rlr_step(self._params_ref, self.precision, self.math_mode, self._state(), _ptr(actions), kind,
                                      _ptr(ic_inject), self._out(), self.n, self.env_offset, self.seed, self._stream()))
            self.launches += 1
        return self.obs, self.reward, self.done

    def _step_chained(self, actions, physical):
        if self.spec.dof != 6 or self.precision != N.F64:
            raise ValueError("chained stepping is implemented for the 6DOF environment with precision='f64'")
        if not (actions.dtype is torch.float32 and tuple(actions.shape) == self._act_shape and actions.is_contiguous()
                and actions.device == self.device):
            raise ValueError(f"chained stepping needs contiguous float32 actions of shape {self._act_shape} on {self.device}")
        prev = torch.cuda.current_device()
        if prev != self._dev_index:
            torch.cuda.set_device(self._dev_index)
        wait = self._chain_seq
        publish = wait + 1 if wait < 0xFFFFFFF0 else 1
        rc = self.lib.rlr_step_chained(self._params_ref, self.precision, self.math_mode, self._st, actions.data_ptr(),
                                       N.ACT_PHYSICAL if physical else N.ACT_NORMALIZED, None, self._ot, self.n, self.env_offset, self.seed,
                                       self._block_seq.data_ptr(), wait, publish, torch.cuda.current_stream(self._dev_index).cuda_stream)
        if prev != self._dev_index:
            torch.cuda.set_device(prev)
        if rc:
            N.check(rc)
        self._chain_seq = publish
        self.launches += 1
        return self.obs, self.reward, self.done

    def step_out_struct(self, obs, reward, done):
        """An RlrStepOut whose obs / reward / done point at caller-owned buffers (e.g. mapped pinned HOST memory: the kernel then
        writes the step's results straight across PCIe, see RocketVecEnv.step_host); every other output stays on the device.
        NB: a step through this struct does not refresh the batch's own `obs` / `reward` / `reward64` / `done` tensors."""
        return N.RlrStepOut(_ptr(obs), _ptr(reward), None, _ptr(done), _ptr(self.truncated), _ptr(self.terminal_obs), _ptr(self.fin_ret),
                            _ptr(self.fin_len), _ptr(self.flags), None, None, None, None, _ptr(self.stats))

    def step_raw(self, actions_ptr, out_struct, stream=None, begin=0, count=None, action_kind=N.ACT_NORMALIZED):
        """One step with a raw action pointer (device, or mapped pinned host memory) and a custom output struct, optionally for the
        sub-range [begin, begin + count) of the environments only and on an explicit CUDA stream (sub-batch pipelining: several
        ranges on several streams, see RocketVecEnv.step_host_async)."""
        prev = torch.cuda.current_device()
        if prev != self._dev_index:
            torch.cuda.set_device(self._dev_index)
        if stream is None:
            stream = torch.cuda.current_stream(self._dev_index).cuda_stream
        if count is None:
            count = self.n - begin
        self._chain_seq = 0
        rc = self.lib.rlr_step_range(self._params_ref, self.precision, self.math_mode, self._st, actions_ptr, action_kind, None,
                                     out_struct, self.n, begin, count, self.env_offset, self.seed, stream)
        if prev != self._dev_index:
            torch.cuda.set_device(prev)
        if rc:
            N.check(rc)
        self.launches += 1

    def rollout(self, actions, obs_buf=None, reward_buf=None, done_buf=None, trunc_buf=None, physical=False):
        """T fused steps in ONE launch.  actions: cuda f32 [T, n, nact] (or i64 [T, n]); optional [T, n, ...] output buffers."""
        T = actions.shape[0]
        self._chain_seq = 0
        actions, kind = self._check_actions(actions, leading=(T,))
        if kind is None:
            kind = N.ACT_PHYSICAL if physical else N.ACT_NORMALIZED
        with torch.cuda.device(self.device):
            N.check(self.lib.rlr_rollout(self._params_ref, self.precision, self.math_mode, self._state(), _ptr(actions), kind, T,
                                         _ptr(obs_buf), _ptr(reward_buf), _ptr(done_buf), _ptr(trunc_buf), _ptr(self.obs),
                                         _ptr(self.stats), self.n, self.env_offset, self.seed, self._stream()))
            self.launches += 1
        return self.obs

    def read_stats(self, reset=False):
        """Episode statistics accumulated on the device by warp-shuffle reductions (host sync)."""
        v = self.stats.cpu().tolist()
        if reset:
            self.stats.zero_()
        d = dict(zip(N.STAT_NAMES, v))
        d["ret_sum"] = d.pop("ret_sum_fx") / 1048576.0
        return d


def fma_peak(precision="f64", device="cuda:0", blocks=None, threads=256, iters=4096, repeats=5):
    """Measured FMA roofline (Tflop/s) of the device: the denominator for roofline.frac (MEASURED_PEAKS.json has no ALU figure)."""
    lib = N.load()
    dev = torch.device(device)
    with torch.cuda.device(dev):
        props = torch.cuda.get_device_properties(dev)
        blocks = blocks or props.multi_processor_count * 8
        sink = torch.zeros(8, dtype=torch.float64, device=dev)
        flops = C.c_double(0)
        prec = N.F64 if precision == "f64" else N.F32
        stream = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
        best = 0.0
        for r in range(repeats + 2):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            N.check(lib.rlr_fma_peak(prec, blocks, threads, iters, _ptr(sink), C.byref(flops), stream))
            e1.record()
            e1.synchronize()
            if r >= 2:
                best = max(best, flops.value / (e0.elapsed_time(e1) * 1e-3) / 1e12)
    return best
