"""PAACLearner: mirror of paac.py:13-187 -- the PAAC loop, act -> step envs -> n-step returns -> train step.

Same constructor, ``train()``, ``cleanup()`` and static ``choose_next_actions(network, num_actions, states,
session)``.  What the reference does with T+2 ``session.run`` calls, three per-env Python loops and NumPy per
update happens here on the GPU through the C ABI (RolloutEngine); the host keeps only the worker fan-out /
fan-in and the episode statistics.

Environments that implement the raw-frame protocol (BaseEnvironment.next_raw) write their two pooled raw
frames into shared, pinned+mapped memory and the GPU does max-pool + resize + stacking; others hand over
stacked 84x84x4 observations exactly as in the reference.
"""
import logging
import time

import numpy as np
import torch

from .actor_learner import ActorLearner
from .emulator_runner import EmulatorRunner, RawFrameEmulatorRunner
from .engine import FRAME_SLOT_SHAPE
from .runners import Runners
from .session import forward_numpy


class MappedArray(object):
    """A host array the GPU can address (page-locked + mapped, Runners.pin): just the device-side address and the shape,
    which is all the C ABI needs."""

    def __init__(self, dev_ptr, shape):
        self._ptr, self.shape = int(dev_ptr), tuple(shape)

    def data_ptr(self):
        return self._ptr


class PAACLearner(ActorLearner):
    def __init__(self, network_creator, environment_creator, args):
        super(PAACLearner, self).__init__(network_creator, environment_creator, args)
        self.workers = args.emulator_workers
        self.raw_frames = bool(getattr(args, 'raw_frames', True))
        self.train_forward = getattr(args, 'train_forward', 'reuse')
        graphs = str(getattr(args, 'graphs', 'auto')).lower()
        # CUDA graphs pay where a cycle is launch-bound (the reference's default: 32 environments, train.py:95)
        self.use_graphs = (graphs == 'true') or (graphs == 'auto' and self.local_emulator_counts <= 1024)
        self.stats = None
        if getattr(args, 'summaries', False) and self.rank == 0:
            from .logger_utils import StatsWriter
            self.stats = StatsWriter(self.debugging_folder, every=getattr(args, 'summary_interval', 100))
        self.last_loss = None
        self.last_norm = None
        self.steps_per_second = None

    @staticmethod
    def choose_next_actions(network, num_actions, states, session):
        """paac.py:18-29: forward + categorical sampling; returns (one_hot [N,A], v [N], pi [N,A])."""
        # (the static method keeps the reference's host-side randomness; train() samples inside the heads kernel)
        uniforms = np.random.random_sample(len(states)).astype(np.float32)
        uniforms = np.minimum(uniforms, np.nextafter(np.float32(1.0), np.float32(0.0)))
        network_output_pi, network_output_v, action_indices = forward_numpy(network, states, uniforms)
        new_actions = np.eye(num_actions)[action_indices]
        return new_actions, network_output_v, network_output_pi

    def __choose_next_actions(self, states):
        return PAACLearner.choose_next_actions(self.network, self.num_actions, states, self.session)

    def train(self):
        """
        Main actor learner loop for parallel advantage actor critic learning.
        """
        self.global_step = self.init_network()

        logging.debug("Starting training at Step {}".format(self.global_step))
        counter = 0
        global_step_start = self.global_step
        total_rewards = []

        eng = self.engine
        N, T, A = self.local_emulator_counts, self.max_local_steps, self.num_actions
        dev = eng.dev
        raw = self.raw_frames and all(getattr(e, 'supports_raw_frames', False) for e in self.emulators)

        # state (or raw frame slots), reward, episode_over, action -- positional, as paac.py:73-77
        if raw:
            first = np.zeros((N,) + FRAME_SLOT_SHAPE, dtype=np.uint8)
            for i, emulator in enumerate(self.emulators):
                emulator.get_initial_state_raw(first[i])
            runner_cls = RawFrameEmulatorRunner
        else:
            first = np.asarray([emulator.get_initial_state() for emulator in self.emulators], dtype=np.uint8)
            runner_cls = EmulatorRunner
        variables = [first,
                     (np.zeros(N, dtype=np.float32)),
                     (np.asarray([False] * N, dtype=np.float32)),
                     (np.zeros((N, A), dtype=np.float32))]

        self.runners = Runners(runner_cls, self.emulators, self.workers, variables)
        self.runners.start()
        shared_states, shared_rewards, shared_episode_over, shared_actions = self.runners.get_shared_variables()
        # All four shared arrays are page-locked and mapped into the GPU's address space: the kernels read the frames, rewards
        # and episode-over flags the workers wrote, and write the sampled one-hot actions, in place -- no staging copies.
        states_dev_ptr = self.runners.pin(0)
        rewards_dev = MappedArray(self.runners.pin(1), shared_rewards.shape)
        over_dev = MappedArray(self.runners.pin(2), shared_episode_over.shape)
        actions_dev = MappedArray(self.runners.pin(3), shared_actions.shape)
        main_stream = torch.cuda.current_stream(dev)

        if raw:      # initial stack: every env is "reset" -> four fresh planes
            all_reset = torch.ones(N, dtype=torch.uint8, device=dev)
            self._preprocess_into(0, states_dev_ptr, all_reset)
        else:
            eng.state(0).copy_(torch.from_numpy(shared_states))
        torch.cuda.synchronize(dev)

        emulator_steps = np.zeros(N, dtype=np.int64)
        total_episode_rewards = np.zeros(N, dtype=np.float64)
        start_time = time.time()

        # The training batch is the concatenation of the acting batches under unchanged parameters (paac.py:92,112,151).
        # 'reuse' (default): the acting forward of step t writes its activations straight into the training workspace and the
        # update runs no second forward -- the same bits as the reference's schedule (tests/test_gpu_tc.py::
        # test_training_forward_schedules_are_bit_identical).  'stepwise': a separate training forward per step on a side
        # stream while the emulators run.  'batched': the reference's schedule, one training forward inside the update.
        eng.set_train_forward(self.train_forward)
        side_stream = torch.cuda.Stream(dev) if self.train_forward == 'stepwise' else None
        if self.use_graphs:
            eng.enable_graphs()        # small batches are launch-bound: act(t) and update() become one graph launch each

        while self.global_step < self.max_global_steps:
            loop_start_time = time.time()
            for t in range(T):
                # paac.py:105-108: forward + sampling; the heads kernel writes np.eye(A)[action] into the shared action array
                eng.act(t, onehot_out=actions_dev)
                main_stream.synchronize()                               # the workers may read the actions now

                # Start updating all environments with next_actions
                self.runners.update_environments()
                if side_stream is not None:
                    side_stream.wait_stream(main_stream)                # states[t] are complete
                    with torch.cuda.stream(side_stream):
                        eng.train_forward_step(t)                       # paac.py:151-161's forward, step t's share
                self.runners.wait_updated()
                # Done updating all environments, have new states, rewards and is_over

                if raw:     # K1 + rewards[t] / over[t] bookkeeping in one launch; episode_over doubles as the reset flag
                    eng.observe_frames(t, states_dev_ptr, 4, None, rewards_dev, over_dev, over_is_reset=True)
                else:
                    eng.observe_states(t, torch.from_numpy(shared_states), torch.from_numpy(shared_rewards),
                                       torch.from_numpy(shared_episode_over))

                # episode statistics (paac.py:121-138), vectorised; reward clipping happens on the GPU
                total_episode_rewards += shared_rewards
                emulator_steps += 1
                self.global_step += self.emulator_counts
                for e in np.nonzero(shared_episode_over)[0]:
                    total_rewards.append(total_episode_rewards[e])
                    if self.stats is not None:
                        self.stats.episode(self.global_step, total_episode_rewards[e], emulator_steps[e])
                    total_episode_rewards[e] = 0
                    emulator_steps[e] = 0

            if side_stream is not None:
                main_stream.wait_stream(side_stream)
            lr = self.get_lr()
            if self.stats is not None and self.stats.wants_gradients(counter):
                eng.forward_backward(); eng.allreduce()
                self.stats.gradients(self.global_step, eng, lr)         # actor_learner.py:85-87, opt-in
                eng.apply(lr); eng.roll()
            else:
                eng.update(lr)                                          # paac.py:140-165
            self.last_loss, self.last_norm = eng.loss, eng.norm

            counter += 1
            if counter % (2048 / self.emulator_counts) == 0:
                curr_time = time.time()
                global_steps = self.global_step
                last_ten = 0.0 if len(total_rewards) < 1 else np.mean(total_rewards[-10:])
                logging.info("Ran {} steps, at {} steps/s ({} steps/s avg), last 10 rewards avg {}"
                             .format(global_steps,
                                     self.max_local_steps * self.emulator_counts / (curr_time - loop_start_time),
                                     (global_steps - global_step_start) / (curr_time - start_time),
                                     last_ten))
            self.save_vars()

        torch.cuda.synchronize(dev)
        self.steps_per_second = (self.global_step - global_step_start) / max(time.time() - start_time, 1e-9)
        self.cleanup()

    def _preprocess_into(self, slot, frames_ptr, reset_u8):
        import ctypes as C
        from . import _lib
        eng = self.engine
        _lib.check(eng.lib.paacb_preprocess_u8(eng.ctx, C.c_void_p(frames_ptr), 4, _lib.ptr(reset_u8),
                                               _lib.ptr(eng.state(slot)), _lib.ptr(eng.state(slot)), eng.N,
                                               eng._stream()), 'paacb_preprocess_u8')

    def cleanup(self):
        super(PAACLearner, self).cleanup()
        if getattr(self, 'runners', None) is not None:
            self.runners.stop()
            for r in self.runners.runners:
                r.join(5)
