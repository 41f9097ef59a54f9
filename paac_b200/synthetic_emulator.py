"""SyntheticEmulator: a BaseEnvironment that emits seeded raw 210x160 luminance frames.

Used where ALE is not installed (this image) and for the synthetic-frame benchmark configurations
(BASELINE.json configs 3 and 5).  It follows AtariEmulator's structure step for step
(atari_emulator.py:15-118): action repeat 4 with the last two frames pooled, 4-frame observation stack,
reset = four action-repeats of action 0, reward summed over the repeat -- only the frame source differs.
Seeding mirrors atari_emulator.py:18: ``random_seed * (actor_id + 1)``.
"""
import numpy as np

from .environment import BaseEnvironment, FramePool, ObservationPool
from .resize_tables import ROW, COL

IMG_SIZE_X = 84
IMG_SIZE_Y = 84
NR_IMAGES = 4
ACTION_REPEAT = 4
FRAMES_IN_POOL = 2
SCREEN_H, SCREEN_W = 210, 160


class SyntheticEmulator(BaseEnvironment):
    supports_raw_frames = True

    def __init__(self, actor_id, args):
        self.actor_id = actor_id
        self.num_actions = int(getattr(args, 'synthetic_actions', 6))
        self.legal_actions = np.arange(self.num_actions, dtype=np.int32)
        self.rng = np.random.RandomState((int(args.random_seed) * (actor_id + 1)) % (2 ** 31))
        self.p_terminal = float(getattr(args, 'synthetic_p_terminal', 0.01))
        self.observation_pool = ObservationPool(np.zeros((IMG_SIZE_X, IMG_SIZE_Y, NR_IMAGES), dtype=np.uint8))
        self.frame_pool = FramePool(np.empty((FRAMES_IN_POOL, SCREEN_H, SCREEN_W), dtype=np.uint8),
                                    self.__process_frame_pool)
        self._terminal = False

    def get_legal_actions(self):
        return self.legal_actions

    def get_noop(self):
        return [1.0, 0.0]

    # ---- frame source -------------------------------------------------------------------------------
    def _emulate(self, a, out_pair):
        """One action repeat: fills out_pair (uint8[2,210,160]) with the last two frames, returns reward."""
        out_pair[...] = self.rng.randint(0, 256, size=(FRAMES_IN_POOL, SCREEN_H, SCREEN_W), dtype=np.uint8)
        u = self.rng.random_sample()
        reward = -1.0 if u < 0.05 else (1.0 if u > 0.95 else 0.0)
        self._terminal = self.rng.random_sample() < self.p_terminal
        return reward

    # ---- classic protocol (atari_emulator.py:69-106) -------------------------------------------------
    def __process_frame_pool(self, frame_pool):
        img = np.amax(frame_pool, axis=0)
        return img[ROW[:, None], COL[None, :]].astype(np.uint8)

    def __action_repeat(self, a):
        return self._emulate(a, self.frame_pool.frame_pool)

    def get_initial_state(self):
        for _ in range(NR_IMAGES):
            self.__action_repeat(0)
            self.observation_pool.new_observation(self.frame_pool.get_processed_frame())
        self._terminal = False
        return self.observation_pool.get_pooled_observations()

    def next(self, action):
        reward = self.__action_repeat(int(np.argmax(action)))
        self.observation_pool.new_observation(self.frame_pool.get_processed_frame())
        return self.observation_pool.get_pooled_observations(), reward, self._terminal

    # ---- raw-frame protocol --------------------------------------------------------------------------
    def next_raw(self, action, out_pairs):
        reward = self._emulate(int(np.argmax(action)), out_pairs[0])
        return reward, self._terminal

    def get_initial_state_raw(self, out_pairs):
        for k in range(NR_IMAGES):
            self._emulate(0, out_pairs[k])
        self._terminal = False
