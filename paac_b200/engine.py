"""RolloutEngine: the device-resident half of PAACLearner.train() (paac.py:99-168).

It owns every device buffer one learner needs for N environments and t_max steps and issues the C-ABI
calls in the order the reference's loop implies:

    for t in range(T):   act(t)      -> paacb_policy_forward(+sample)          paac.py:105-112
                         observe(t)  -> paacb_preprocess_u8 / state upload     emulator_runner.py:24-31, paac.py:119-123
    update(lr)           -> bootstrap forward (paac.py:140-142), training forward, paacb_returns_loss_grad
                            (paac.py:144-149 + the loss graph), paacb_backward, [NCCL allreduce],
                            paacb_clip_rmsprop (actor_learner.py:54-70)

All launches go to the current torch stream without host synchronisation, so ``update`` (and ``act`` +
``observe`` when frames are device-resident) can be captured in a CUDA graph (``capture_update``).
PyTorch is used for memory, streams and torch.distributed only.

Scheduling of the training forward (``train_forward=``).  PAAC's training batch is the concatenation of the t_max
acting batches under unchanged parameters (paac.py:92,112,151), so its forward can be scheduled three ways with
bit-identical results (tests/test_gpu_learner.py):
    'batched'   the reference's schedule: one forward over T*N samples inside update()               (default)
    'stepwise'  the SAME work issued per step with ``train_forward_step(t)`` (paacb_policy_forward_at), e.g. while the
                emulators run and the frames cross PCIe; update() issues whatever steps are still missing
    'reuse'     act(t) writes its activations straight into the training workspace and update() runs no training
                forward at all (the acting forward already IS that computation); values[t] aliases v
"""
import ctypes as C

import numpy as np
import torch

from . import _lib

STATE_SHAPE = (84, 84, 4)
FRAME_SLOT_SHAPE = (4, 2, 210, 160)


class RolloutEngine(object):

    def __init__(self, network, n_envs, t_max, gamma=0.99, rho=0.99, eps=0.1, momentum=0.0,
                 clip_norm=3.0, clip_norm_type='global', seed=3, process_group=None, world_size=1,
                 train_forward='batched', overlap_allreduce=True):
        self.net = network
        self.lib = network._lib
        self.ctx = network.ctx
        self.dev = network.torch_device
        self.N, self.T, self.A = int(n_envs), int(t_max), int(network.num_actions)
        self.B = self.N * self.T
        self.gamma, self.rho, self.eps, self.momentum = float(gamma), float(rho), float(eps), float(momentum)
        self.clip_norm = float(clip_norm)
        if clip_norm_type == 'global':
            self.clip_type = _lib.CLIP_GLOBAL
        elif clip_norm_type == 'ignore':
            self.clip_type = _lib.CLIP_IGNORE
        elif clip_norm_type == 'local':
            raise Exception("clip_norm_type 'local' is broken in the reference (actor_learner.py:62-63 iterates "
                            "(grad, var) tuples); use 'global' or 'ignore'")
        else:
            raise Exception('Norm type not recognized')          # actor_learner.py:67
        self.beta = float(network.entropy_regularisation_strength)
        self.group, self.world = process_group, int(world_size)
        self.overlap_allreduce = bool(overlap_allreduce)
        self.tail_off = int(self.lib.paacb_grad_tail_offset(self.ctx))
        self._pending = []

        d, N, T, A, B = self.dev, self.N, self.T, self.A, self.B
        f32 = dict(dtype=torch.float32, device=d)
        self.states = torch.zeros((T + 1, N) + STATE_SHAPE, dtype=torch.uint8, device=d)
        self.actions = torch.zeros((T, N), dtype=torch.int32, device=d)
        self.onehot = torch.zeros((N, A), **f32)
        self.values = torch.zeros((T, N), **f32)
        self.rewards = torch.zeros((T, N), **f32)
        self.over = torch.zeros((T, N), **f32)
        self.uniforms = torch.zeros((T, N), **f32)
        self.pi_act = torch.zeros((N, A), **f32)
        self.boot_v = torch.zeros((N,), **f32)
        self.boot_pi = torch.zeros((N, A), **f32)
        self.pi = torch.zeros((B, A), **f32)
        self.v = torch.zeros((B,), **f32)
        self.y = torch.zeros((B,), **f32)
        self.adv = torch.zeros((B,), **f32)
        self.dlogits = torch.zeros((B, A), **f32)
        self.dv = torch.zeros((B,), **f32)
        self.loss = torch.zeros((1,), **f32)
        self.norm = torch.zeros((1,), **f32)
        P = network.param_count
        self.grads = torch.zeros((P,), **f32)
        self.ms = torch.ones((P,), **f32)            # rms slot <- ones, momentum slot <- zeros
        self.mom = torch.zeros((P,), **f32)
        self.act_ws = torch.empty((network.workspace_floats(N),), **f32)
        self.fwd_ws = torch.empty((network.workspace_floats(B),), **f32)
        self.bwd_ws = torch.empty((int(self.lib.paacb_backward_workspace_floats(self.ctx, B)),), **f32)
        self.opt_ws = torch.empty((int(self.lib.paacb_optimizer_workspace_floats(self.ctx)),), **f32)
        self.gen = torch.Generator(device=d)
        self.gen.manual_seed(int(seed))
        self._slice_ws = {}        # forward workspaces of environment slices (act(t, lo, hi))
        self._stepped = set()      # (t, lo, hi) slices of the training forward already issued ('stepwise')
        self._booted = []          # [lo, hi) slices whose bootstrap value V(s_T) was already issued (bootstrap())
        self._values_own = self.values
        self.set_train_forward(train_forward)

    # ---- helpers ----------------------------------------------------------------------------------
    def set_train_forward(self, mode):
        """Choose the schedule of the training forward (see the module docstring); call between updates only."""
        if mode not in ('batched', 'stepwise', 'reuse'):
            raise ValueError("train_forward must be 'batched', 'stepwise' or 'reuse'")
        self.train_forward = mode
        self._stepped.clear()
        # 'reuse': the acting value IS the training value
        self.values = self.v.view(self.T, self.N) if mode == 'reuse' else self._values_own

    def _stream(self):
        return C.c_void_p(torch.cuda.current_stream(self.dev).cuda_stream)

    def draw_uniforms(self):
        """One Philox draw per update for all T*N sampling decisions."""
        self.uniforms.uniform_(0.0, 1.0, generator=self.gen)
        # uniform_ on fp32 can return exactly 1.0 only by rounding; keep u in [0, 1)
        self.uniforms.clamp_(max=float(np.nextafter(np.float32(1.0), np.float32(0.0))))

    # ---- rollout ----------------------------------------------------------------------------------
    def act(self, t, lo=0, hi=None):
        """Forward on states[t] + categorical sampling; fills actions[t], values[t], onehot, pi_act.
        [lo, hi) restricts the call to a contiguous slice of the environments (the runner protocol hands every worker
        a contiguous slice, runners.py:17-18): slices can be issued on different streams so that the next slice's
        forward overlaps this slice's frame ingestion."""
        if hi is None:
            hi = self.N
        if self.train_forward == 'reuse':
            first = t * self.N + lo
            self.net.forward(self.states[t, lo:hi], self.pi[first:first + hi - lo], self.v[first:first + hi - lo], self.fwd_ws,
                             uniforms=self.uniforms[t, lo:hi], actions=self.actions[t, lo:hi], onehot=self.onehot[lo:hi],
                             ws_capacity=self.B, ws_first=first)
            return
        if lo == 0 and hi == self.N:
            ws = self.act_ws
        else:
            ws = self._slice_ws.get((lo, hi))
            if ws is None:
                ws = self._slice_ws[(lo, hi)] = torch.empty((self.net.workspace_floats(hi - lo),), dtype=torch.float32,
                                                            device=self.dev)
        self.net.forward(self.states[t, lo:hi], self.pi_act[lo:hi], self.values[t, lo:hi], ws,
                         uniforms=self.uniforms[t, lo:hi], actions=self.actions[t, lo:hi], onehot=self.onehot[lo:hi])

    def observe_frames(self, t, frames_ptr, pairs_per_env, reset_u8, rewards, over, lo=0, hi=None):
        """Raw-frame protocol: states[t+1] <- preprocess(frames | states[t]); rewards/over: device or pinned host tensors.
        frames_ptr addresses the slot of environment `lo`; it may point into pinned, mapped HOST memory (the kernel then
        reads the rows it needs zero-copy over PCIe)."""
        if hi is None:
            hi = self.N
        p = _lib.ptr
        _lib.check(self.lib.paacb_preprocess_u8(self.ctx, C.c_void_p(frames_ptr), int(pairs_per_env),
                                                p(reset_u8[lo:hi]) if reset_u8 is not None else None,
                                                p(self.states[t, lo:hi]), p(self.states[t + 1, lo:hi]), hi - lo,
                                                self._stream()), 'paacb_preprocess_u8')
        self.rewards[t, lo:hi].copy_(rewards[lo:hi] if rewards.shape[0] == self.N else rewards, non_blocking=True)
        self.over[t, lo:hi].copy_(over[lo:hi] if over.shape[0] == self.N else over, non_blocking=True)

    def observe_states(self, t, states, rewards, over):
        """Classic protocol: the environments produced stacked 84x84x4 observations themselves."""
        self.states[t + 1].copy_(states, non_blocking=True)
        self.rewards[t].copy_(rewards, non_blocking=True)
        self.over[t].copy_(over, non_blocking=True)

    def _ws_for(self, lo, hi):
        if lo == 0 and hi == self.N:
            return self.act_ws
        ws = self._slice_ws.get((lo, hi))
        if ws is None:
            ws = self._slice_ws[(lo, hi)] = torch.empty((self.net.workspace_floats(hi - lo),), dtype=torch.float32,
                                                        device=self.dev)
        return ws

    def bootstrap(self, lo=0, hi=None):
        """V(s_T) of a slice of the environments (paac.py:140-142), as soon as ITS last frames are in: the end-to-end loop
        issues it per slice so that it overlaps the frame ingestion of the other slices.  update() computes the bootstrap
        of whatever slices were not issued."""
        if hi is None:
            hi = self.N
        self.net.forward(self.states[self.T, lo:hi], self.boot_pi[lo:hi], self.boot_v[lo:hi], self._ws_for(lo, hi))
        self._booted.append((lo, hi))

    def train_forward_step(self, t, lo=0, hi=None):
        """'stepwise': the training forward of the samples (t, lo..hi), written into their place in the batch workspace."""
        if self.train_forward != 'stepwise':
            raise RuntimeError("train_forward_step() needs RolloutEngine(train_forward='stepwise')")
        if hi is None:
            hi = self.N
        first = t * self.N + lo
        self.net.forward(self.states[t, lo:hi], self.pi[first:first + hi - lo], self.v[first:first + hi - lo], self.fwd_ws,
                         ws_capacity=self.B, ws_first=first)
        self._stepped.add((t, lo, hi))

    def _finish_stepwise(self):
        """Issue the steps of the training forward the caller has not issued itself (whole steps only)."""
        for t in range(self.T):
            covered = sorted((lo, hi) for (tt, lo, hi) in self._stepped if tt == t)
            pos = 0
            for lo, hi in covered:
                if lo > pos:
                    self.train_forward_step(t, pos, lo)
                pos = max(pos, hi)
            if pos < self.N:
                self.train_forward_step(t, pos, self.N)
        self._stepped.clear()

    # ---- update -----------------------------------------------------------------------------------
    def forward_backward(self):
        """Bootstrap forward, training forward, returns + loss gradient, backward.  Leaves dL/dparams in grads."""
        T, N, B = self.T, self.N, self.B
        p = _lib.ptr
        st = self._stream()
        pos = 0                                                                                   # paac.py:140-142
        for lo, hi in sorted(self._booted):
            if lo > pos:
                self.bootstrap(pos, lo)
            pos = max(pos, hi)
        if pos < N:
            self.bootstrap(pos, N)
        self._booted = []
        flat_states = self.states[:T].view((B,) + STATE_SHAPE)                                    # paac.py:151
        if self.train_forward == 'batched':
            self.net.forward(flat_states, self.pi, self.v, self.fwd_ws)
        elif self.train_forward == 'stepwise':
            self._finish_stepwise()
        _lib.check(self.lib.paacb_returns_loss_grad(
            self.ctx, p(self.rewards), p(self.over), p(self.values), p(self.boot_v), p(self.actions), p(self.pi),
            p(self.v), T, N, C.c_double(self.gamma), C.c_float(self.beta), p(self.y), p(self.adv), p(self.dlogits),
            p(self.dv), p(self.loss), st), 'paacb_returns_loss_grad')
        if self.world > 1 and self.overlap_allreduce:
            # multi-GPU: the tail of the flat gradient (hidden fc layer + heads, 95 % of the parameters) is final after the
            # first part of the backward; its all-reduce runs on NCCL's stream under the conv layers' weight gradients
            for part in (_lib.BWD_TAIL, _lib.BWD_HEAD):
                _lib.check(self.lib.paacb_backward_part(self.ctx, p(self.net.params), p(flat_states), B, p(self.fwd_ws),
                                                        p(self.dlogits), p(self.dv), p(self.bwd_ws), p(self.grads), part, st),
                           'paacb_backward_part')
                if part == _lib.BWD_TAIL:
                    self._pending = [torch.distributed.all_reduce(self.grads[self.tail_off:], op=torch.distributed.ReduceOp.SUM,
                                                                  group=self.group, async_op=True)]
            return
        _lib.check(self.lib.paacb_backward(self.ctx, p(self.net.params), p(flat_states), B, p(self.fwd_ws),
                                           p(self.dlogits), p(self.dv), p(self.bwd_ws), p(self.grads), st),
                   'paacb_backward')

    def allreduce(self):
        if self.world > 1:
            if self._pending:
                self._pending.append(torch.distributed.all_reduce(self.grads[:self.tail_off], op=torch.distributed.ReduceOp.SUM,
                                                                  group=self.group, async_op=True))
                for w in self._pending:
                    w.wait()             # stream-level wait: the optimizer kernel is ordered after both reductions
                self._pending = []
            else:
                torch.distributed.all_reduce(self.grads, op=torch.distributed.ReduceOp.SUM, group=self.group)

    def apply(self, lr):
        p = _lib.ptr
        _lib.check(self.lib.paacb_clip_rmsprop(
            self.ctx, p(self.net.params), p(self.ms), p(self.mom), p(self.grads), C.c_float(1.0 / self.world),
            C.c_float(lr), C.c_float(self.rho), C.c_float(self.eps), C.c_float(self.momentum),
            C.c_float(self.clip_norm), self.clip_type, p(self.norm), p(self.opt_ws), self._stream()),
            'paacb_clip_rmsprop')

    def roll(self):
        """The last state of this rollout is the first state of the next (paac.py:99-112 reuse shared_states)."""
        self.states[0].copy_(self.states[self.T], non_blocking=True)

    def update(self, lr):
        self.forward_backward()
        self.allreduce()
        self.apply(lr)
        self.roll()
