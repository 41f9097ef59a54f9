"""SyntheticEmulator: an environment that emits seeded raw 210x160 luminance frames.

Used where ALE is not installed (this image) and for the synthetic-frame benchmark configurations
(BASELINE.json configs 3 and 5).  It has the cadence of the reference's AtariEmulator (atari_emulator.py:77-106): one
``next`` = one action repeat whose last two frames are pooled, reward summed over the repeat, reset = four action repeats
of action 0 -- only the frame source differs.  It speaks the raw-frame protocol; the classic ``next`` /
``get_initial_state`` come from RawFrameEnvironment.  Seeding mirrors atari_emulator.py:18: ``random_seed * (actor_id + 1)``.
"""
import numpy as np

from .environment import RawFrameEnvironment, STACK, PAIR, FRAME_SHAPE


class SyntheticEmulator(RawFrameEnvironment):

    def __init__(self, actor_id, args):
        self.actor_id = actor_id
        self.num_actions = int(getattr(args, 'synthetic_actions', 6))
        self.legal_actions = np.arange(self.num_actions, dtype=np.int32)
        self.rng = np.random.RandomState((int(args.random_seed) * (actor_id + 1)) % (2 ** 31))
        self.p_terminal = float(getattr(args, 'synthetic_p_terminal', 0.01))
        self._terminal = False

    def get_legal_actions(self):
        return self.legal_actions

    def get_noop(self):
        return [1.0, 0.0]

    def _emulate(self, a, out_pair):
        """One action repeat: fills out_pair (uint8[2,210,160]) with the last two frames, returns the summed reward."""
        out_pair[...] = self.rng.randint(0, 256, size=(PAIR,) + FRAME_SHAPE, dtype=np.uint8)
        u = self.rng.random_sample()
        reward = -1.0 if u < 0.05 else (1.0 if u > 0.95 else 0.0)
        self._terminal = self.rng.random_sample() < self.p_terminal
        return reward

    def next_raw(self, action, out_pairs):
        reward = self._emulate(int(np.argmax(action)), out_pairs[0])
        return reward, self._terminal

    def get_initial_state_raw(self, out_pairs):
        for k in range(STACK):
            self._emulate(0, out_pairs[k])
        self._terminal = False
