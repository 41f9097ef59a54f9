"""AtariEmulator: the Arcade Learning Environment behind the raw-frame protocol.

Stands in for the reference's atari_emulator.py:15-118.  ALE is a host-side C++ dependency that stays on the CPU
(north_star: "host ALE runners"); it is NOT installed in this image, so importing this module raises ImportError there
and EnvironmentCreator reports that plainly.

What is kept from the reference is the emulation contract, because parity depends on it: ALE settings
(atari_emulator.py:17-23: seed = random_seed * (actor_id + 1), no sticky actions, frame_skip 1, no colour averaging),
the minimal action set, action repeat 4 with the LAST TWO luminance frames kept (:77-86), reset = reset_game + up to 30
random no-ops + four repeats of action 0 (:60-67, :88-96), terminal = game over, or a lost life with
``single_life_episodes`` (:108-115).  What is gone is every NumPy / PIL operation on pixels: ``getScreenGrayscale``
writes straight into the frame-pair slot the runner hands in (pinned, mapped shared memory) and the GPU does max-pool,
resize and stacking.  The classic ``next`` / ``get_initial_state`` are derived by RawFrameEnvironment.
"""
import random

import numpy as np
from ale_python_interface import ALEInterface   # noqa: E402  (absent here -> ImportError, by design)

from .environment import RawFrameEnvironment, STACK, PAIR

ACTION_REPEAT = 4
MAX_START_WAIT = 30


class AtariEmulator(RawFrameEnvironment):

    def __init__(self, actor_id, args):
        ale = ALEInterface()
        ale.setInt(b"random_seed", args.random_seed * (actor_id + 1))
        ale.setFloat(b"repeat_action_probability", 0.0)
        ale.setInt(b"frame_skip", 1)
        ale.setBool(b"color_averaging", False)
        ale.loadROM(str.encode(args.rom_path + "/" + args.game + ".bin"))
        self.ale = ale
        self.legal_actions = ale.getMinimalActionSet()
        width, height = ale.getScreenDims()
        self.random_start = args.random_start
        self.single_life_episodes = args.single_life_episodes
        self.lives = ale.lives()
        self._rgb = np.zeros((height, width, 3), dtype=np.uint8) if args.visualize else None

    def get_legal_actions(self):
        return self.legal_actions

    def get_noop(self):
        return [1.0, 0.0]

    def _repeat(self, action_index, pair):
        """ACTION_REPEAT emulator steps of one action; the last PAIR screens land in pair[0], pair[1]."""
        code = self.legal_actions[action_index]
        reward = 0
        for step in range(ACTION_REPEAT):
            reward += self.ale.act(code)
            keep = step - (ACTION_REPEAT - PAIR)
            if keep >= 0:
                self.ale.getScreenGrayscale(pair[keep][..., None])       # (210, 160, 1) view of the shared slot: no copy
                if self._rgb is not None:
                    self.ale.getScreenRGB(self._rgb)
                    self.on_new_frame(self._rgb)
        return reward

    def _terminal(self):
        lost_life = self.single_life_episodes and self.lives > self.ale.lives()
        return self.ale.game_over() or lost_life

    def next_raw(self, action, out_pairs):
        reward = self._repeat(int(np.argmax(action)), out_pairs[0])
        terminal = self._terminal()
        self.lives = self.ale.lives()
        return reward, terminal

    def get_initial_state_raw(self, out_pairs):
        self.ale.reset_game()
        self.lives = self.ale.lives()
        if self.random_start:
            for _ in range(random.randint(0, MAX_START_WAIT)):
                self.ale.act(self.legal_actions[0])
        for k in range(STACK):
            self._repeat(0, out_pairs[k])
        if self._terminal():
            raise Exception('This should never happen.')          # atari_emulator.py:95
