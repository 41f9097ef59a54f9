"""Oracle (TEST INFRASTRUCTURE ONLY): the reference's hot path restated for the host CPU, timed as a baseline.

"Restated reference CPU baseline (not TF1)": TensorFlow 1 and ALE cannot be installed here (no network), so
bench.py times this port of one PAAC update cycle (paac.py:99-168) on the box's host cores:
  T x [ forward on N stacked states + categorical sampling (paac.py:105-112),
        synthetic raw frames -> np.amax + nearest resize + 4-frame ring (atari_emulator.py:69-75, environment.py:58-75) ]
  + bootstrap forward, float64 n-step returns (paac.py:140-149),
  + loss / autograd gradients (policy_v_network.py, actor_learner.py:44), global-norm clip, RMSProp (actor_learner.py:54-70).
torch-CPU fp32 with torch.set_num_threads(cores); preprocessing in NumPy as the reference's workers do.
It is never on the product path; it is only ever the checker or the timed baseline.
"""
import os
import time

import numpy as np
import torch

from . import network, update, preprocess


class CpuPaac(object):
    def __init__(self, arch='NATURE', n_envs=32, t_max=5, num_actions=6, seed=3, cores=None):
        self.cores = int(cores or os.cpu_count() or 1)
        torch.set_num_threads(self.cores)
        self.arch, self.N, self.T, self.A = arch, n_envs, t_max, num_actions
        self.params = network.init_params(arch, num_actions, seed)
        self.specs = network.param_specs(arch, num_actions)
        self.ms = {n: np.ones(s, np.float32) for n, s, _ in self.specs}
        self.mom = {n: np.zeros(s, np.float32) for n, s, _ in self.specs}
        self.rng = np.random.RandomState(seed)
        self.row, self.col = preprocess.pillow_nearest_tables()
        self.states = self.rng.randint(0, 256, (n_envs, 84, 84, 4)).astype(np.uint8)
        self.global_step = 0

    def cycle(self):
        N, T, A = self.N, self.T, self.A
        states = np.zeros((T, N, 84, 84, 4), np.uint8)
        actions = np.zeros((T, N), np.int64)
        values = np.zeros((T, N), np.float32)
        rewards = np.zeros((T, N)); over = np.zeros((T, N))
        with torch.no_grad():
            for t in range(T):
                out = network.forward(self.params, self.states, self.arch)
                actions[t] = update.sample_actions(out['pi'].numpy(), self.rng.random_sample(N).astype(np.float32))
                values[t] = out['v'].numpy()
                states[t] = self.states
                frames = self.rng.randint(0, 256, (N, 1, 2, 210, 160), dtype=np.uint8)
                u = self.rng.random_sample(N)
                rewards[t] = np.where(u < 0.05, -1.0, np.where(u > 0.95, 1.0, 0.0))
                over[t] = self.rng.random_sample(N) < 0.01
                self.states = preprocess.step_states(self.states, frames, np.zeros(N, np.uint8), self.row, self.col)
                self.global_step += N
            boot = network.forward(self.params, self.states, self.arch)['v'].numpy()
        y, adv = update.nstep_returns(rewards, over, values, boot, 0.99)
        loss, grads, _ = network.loss_and_grads(self.params, states.reshape(T * N, 84, 84, 4), actions.reshape(-1),
                                                adv.reshape(-1), y.reshape(-1), np.float32(0.02), self.arch, A)
        names = [n for n, _, _ in self.specs]
        clipped, _ = update.clip_by_global_norm([grads[n] for n in names], 3.0)
        lr = update.get_lr(self.global_step, 0.0224, 80000000)
        for n, g in zip(names, clipped):
            self.params[n], self.ms[n], self.mom[n] = update.rmsprop_apply(self.params[n], self.ms[n], self.mom[n], g,
                                                                           lr, 0.99, 0.1)
        return loss


def time_cycles(arch, n_envs, t_max, num_actions, steps, warmup, cores=None, max_seconds=None):
    """Returns (env_steps_per_s, ms_per_cycle, cycles_timed, cores)."""
    m = CpuPaac(arch, n_envs, t_max, num_actions, cores=cores)
    for _ in range(warmup):
        m.cycle()
    t0 = time.perf_counter()
    done = 0
    for _ in range(steps):
        m.cycle()
        done += 1
        if max_seconds is not None and time.perf_counter() - t0 > max_seconds:
            break
    dt = time.perf_counter() - t0
    return done * n_envs * t_max / dt, 1e3 * dt / done, done, m.cores
