"""AtariEmulator: mirror of the reference's atari_emulator.py:15-118 over the Arcade Learning Environment.

ALE is a host-side C++ dependency that stays on the CPU (north_star: "host ALE runners").  It is NOT
installed in this image, so this module raises ImportError on import there; EnvironmentCreator reports
that plainly.  Besides the classic ``next`` / ``get_initial_state`` (NumPy max-pool + nearest resize + stack,
as upstream), it implements the raw-frame protocol: ``getScreenGrayscale`` writes straight into the shared,
pinned+mapped frame slots and the GPU does the rest (paacb_preprocess_u8).
"""
import random

import numpy as np
from ale_python_interface import ALEInterface   # noqa: E402  (absent here -> ImportError, by design)

from .environment import BaseEnvironment, FramePool, ObservationPool
from .resize_tables import ROW, COL

IMG_SIZE_X = 84
IMG_SIZE_Y = 84
NR_IMAGES = 4
ACTION_REPEAT = 4
MAX_START_WAIT = 30
FRAMES_IN_POOL = 2


class AtariEmulator(BaseEnvironment):
    supports_raw_frames = True

    def __init__(self, actor_id, args):
        self.ale = ALEInterface()
        self.ale.setInt(b"random_seed", args.random_seed * (actor_id + 1))
        self.ale.setFloat(b"repeat_action_probability", 0.0)
        self.ale.setInt(b"frame_skip", 1)
        self.ale.setBool(b"color_averaging", False)
        full_rom_path = args.rom_path + "/" + args.game + ".bin"
        self.ale.loadROM(str.encode(full_rom_path))
        self.legal_actions = self.ale.getMinimalActionSet()
        self.screen_width, self.screen_height = self.ale.getScreenDims()
        self.lives = self.ale.lives()

        self.random_start = args.random_start
        self.single_life_episodes = args.single_life_episodes
        self.call_on_new_frame = args.visualize

        self.observation_pool = ObservationPool(np.zeros((IMG_SIZE_X, IMG_SIZE_Y, NR_IMAGES), dtype=np.uint8))
        self.rgb_screen = np.zeros((self.screen_height, self.screen_width, 3), dtype=np.uint8)
        self.gray_screen = np.zeros((self.screen_height, self.screen_width, 1), dtype=np.uint8)
        self.frame_pool = FramePool(np.empty((2, self.screen_height, self.screen_width), dtype=np.uint8),
                                    self.__process_frame_pool)

    def get_legal_actions(self):
        return self.legal_actions

    def __get_screen_image(self):
        self.ale.getScreenGrayscale(self.gray_screen)
        if self.call_on_new_frame:
            self.ale.getScreenRGB(self.rgb_screen)
            self.on_new_frame(self.rgb_screen)
        return np.squeeze(self.gray_screen)

    def on_new_frame(self, frame):
        pass

    def __new_game(self):
        self.ale.reset_game()
        self.lives = self.ale.lives()
        if self.random_start:
            wait = random.randint(0, MAX_START_WAIT)
            for _ in range(wait):
                self.ale.act(self.legal_actions[0])

    def __process_frame_pool(self, frame_pool):
        img = np.amax(frame_pool, axis=0)
        return img[ROW[:, None], COL[None, :]].astype(np.uint8)

    def __action_repeat(self, a, times=ACTION_REPEAT, sink=None):
        """Repeat the action; the last FRAMES_IN_POOL frames go to ``sink`` (frame pool or raw slot)."""
        reward = 0
        for _ in range(times - FRAMES_IN_POOL):
            reward += self.ale.act(self.legal_actions[a])
        for i in range(FRAMES_IN_POOL):
            reward += self.ale.act(self.legal_actions[a])
            if sink is None:
                self.frame_pool.new_frame(self.__get_screen_image())
            else:
                sink[i] = self.__get_screen_image()
        return reward

    def get_initial_state(self):
        self.__new_game()
        for _ in range(NR_IMAGES):
            self.__action_repeat(0)
            self.observation_pool.new_observation(self.frame_pool.get_processed_frame())
        if self.__is_terminal():
            raise Exception('This should never happen.')
        return self.observation_pool.get_pooled_observations()

    def next(self, action):
        reward = self.__action_repeat(np.argmax(action))
        self.observation_pool.new_observation(self.frame_pool.get_processed_frame())
        terminal = self.__is_terminal()
        self.lives = self.ale.lives()
        return self.observation_pool.get_pooled_observations(), reward, terminal

    # ---- raw-frame protocol --------------------------------------------------------------------------
    def next_raw(self, action, out_pairs):
        reward = self.__action_repeat(np.argmax(action), sink=out_pairs[0])
        terminal = self.__is_terminal()
        self.lives = self.ale.lives()
        return reward, terminal

    def get_initial_state_raw(self, out_pairs):
        self.__new_game()
        for k in range(NR_IMAGES):
            self.__action_repeat(0, sink=out_pairs[k])
        if self.__is_terminal():
            raise Exception('This should never happen.')

    def __is_terminal(self):
        if self.single_life_episodes:
            return self.__is_over() or (self.lives > self.ale.lives())
        return self.__is_over()

    def __is_over(self):
        return self.ale.game_over()

    def get_noop(self):
        return [1.0, 0.0]
