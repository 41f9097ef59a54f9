"""Environment plugin API: mirror of the reference's environment.py:4-75.

``BaseEnvironment`` is the contract an environment author implements (same five methods, same
semantics).  Two optional additions let an environment take part in the raw-frame protocol, in which
the worker only writes the two pooled raw luminance frames and the GPU does max-pool + resize + stacking
(paacb_preprocess_u8): ``next_raw`` and ``get_initial_state_raw``.

``FramePool`` / ``ObservationPool`` are the host-side observation helpers for environments that stay on
the classic protocol (they hand 84x84x4 observations to the learner, as in the reference).
"""
import numpy as np

from .resize_tables import ROW, COL

STACK = 4                      # frames per observation (atari_emulator.py:11 NR_IMAGES)
FRAME_SHAPE = (210, 160)       # ALE luminance screen
PAIR = 2                       # frames max-pooled per observation plane (atari_emulator.py:14 FRAMES_IN_POOL)


class BaseEnvironment(object):
    def get_initial_state(self):
        """
        Sets the environment to its initial state.
        :return: the initial state
        """
        raise NotImplementedError()

    def next(self, action):
        """
        Appies the current action to the environment.
        :param action: one hot vector.
        :return: (observation, reward, is_terminal) tuple
        """
        raise NotImplementedError()

    def get_legal_actions(self):
        """
        Get the set of indices of legal actions
        :return: a numpy array of the indices of legal actions
        """
        raise NotImplementedError()

    def get_noop(self):
        """
        Gets the no-op action, to be used with self.next
        :return: the action
        """
        raise NotImplementedError()

    def on_new_frame(self, frame):
        """
        Called whenever a new frame is available.
        :param frame: raw frame
        """
        pass

    # ---- raw-frame protocol (optional) ------------------------------------------------------------
    supports_raw_frames = False

    def next_raw(self, action, out_pairs):
        """Step and write the two pooled raw frames into out_pairs[0] (uint8[2,210,160]).
        :return: (reward, is_terminal)"""
        raise NotImplementedError()

    def get_initial_state_raw(self, out_pairs):
        """Reset and write the four action-repeat frame pairs of the initial state, oldest first, into
        out_pairs[0..3] (uint8[4,2,210,160])."""
        raise NotImplementedError()


class FramePool(object):
    """environment.py:42-55: ring of the last frames, reduced by ``operation`` on demand."""

    def __init__(self, frame_pool, operation):
        self.frame_pool = frame_pool
        self.frame_pool_index = 0
        self.frames_in_pool = frame_pool.shape[0]
        self.operation = operation

    def new_frame(self, frame):
        self.frame_pool[self.frame_pool_index] = frame
        self.frame_pool_index = (self.frame_pool_index + 1) % self.frames_in_pool

    def get_processed_frame(self):
        return self.operation(self.frame_pool)


class ObservationPool(object):
    """environment.py:58-75: ring over the last axis; reads come back oldest-first."""

    def __init__(self, observation_pool):
        self.observation_pool = observation_pool
        self.pool_size = observation_pool.shape[-1]
        self.current_observation_index = 0

    def new_observation(self, observation):
        self.observation_pool[..., self.current_observation_index] = observation
        self.current_observation_index = (self.current_observation_index + 1) % self.pool_size

    def get_pooled_observations(self):
        return np.roll(self.observation_pool, -self.current_observation_index, axis=-1).copy()


class RawFrameEnvironment(BaseEnvironment):
    """Base class for environments that speak the raw-frame protocol natively: a subclass implements ``next_raw`` and
    ``get_initial_state_raw`` (it only ever produces raw 210x160 luminance frame pairs) and the GPU turns them into
    observations (paacb_preprocess_u8).  The classic methods of the plugin contract are DERIVED here for callers that
    want 84x84x4 observations on the host (the evaluation loop, ``--raw_frames False``): the same arithmetic in NumPy --
    element-wise max of the pair (atari_emulator.py:72), nearest resize through the Pillow index tables (:73), a 4-deep
    stack with the oldest plane first (environment.py:58-75)."""
    supports_raw_frames = True

    @staticmethod
    def _plane(pair):
        return np.maximum(pair[0], pair[1])[ROW[:, None], COL[None, :]]

    def get_initial_state(self):
        pairs = np.empty((STACK, PAIR) + FRAME_SHAPE, dtype=np.uint8)
        self.get_initial_state_raw(pairs)
        self._stack = np.stack([self._plane(p) for p in pairs], axis=-1)
        self._pair = pairs[:1].copy()
        return self._stack.copy()

    def next(self, action):
        reward, terminal = self.next_raw(action, self._pair)
        self._stack = np.concatenate([self._stack[..., 1:], self._plane(self._pair[0])[..., None]], axis=-1)
        return self._stack.copy(), reward, terminal
