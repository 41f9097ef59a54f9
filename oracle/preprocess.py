"""Oracle (TEST INFRASTRUCTURE ONLY): Atari observation pipeline on the CPU.

Restates, in NumPy:
  * ``AtariEmulator.__process_frame_pool``  (atari_emulator.py:69-75):
      ``img = np.amax(frame_pool, axis=0)``; ``imresize(img, (84, 84), 'nearest')``
  * ``FramePool``                           (environment.py:42-55)
  * ``ObservationPool``                     (environment.py:58-75)
  * the reset rule of ``EmulatorRunner._run`` (emulator_runner.py:24-29) combined
    with ``AtariEmulator.get_initial_state`` (atari_emulator.py:88-96): on a
    terminal step the state handed to the learner is a fresh stack of four
    pooled frames.

``scipy.misc.imresize(img, (84, 84), interp='nearest')`` for uint8 input is
``PIL.Image.fromarray(img).resize((84, 84), NEAREST)`` (scipy <1.3 source:
``toimage`` does no byte scaling for uint8, then ``im.resize(size, resample=0)``).
The index tables are recovered from the installed Pillow by resizing index
images, see ``pillow_nearest_tables``.
"""
import numpy as np

SCREEN_H, SCREEN_W = 210, 160      # ale.getScreenDims(), atari_emulator.py:28
IMG = 84                           # IMG_SIZE_X / IMG_SIZE_Y, atari_emulator.py:7-8
NR_IMAGES = 4                      # atari_emulator.py:9
FRAMES_IN_POOL = 2                 # atari_emulator.py:12


def pillow_nearest_tables(src_h=SCREEN_H, src_w=SCREEN_W, dst=IMG):
    """(ROW[dst], COL[dst]) int32 such that Pillow NEAREST gives out[y, x] = img[ROW[y], COL[x]]."""
    from PIL import Image
    # index images must fit uint8: encode the index in two planes (hi, lo)
    rows = np.repeat(np.arange(src_h, dtype=np.int32)[:, None], src_w, axis=1)
    cols = np.repeat(np.arange(src_w, dtype=np.int32)[None, :], src_h, axis=0)

    def rs(a):
        return np.asarray(Image.fromarray(a.astype(np.uint8)).resize((dst, dst), Image.NEAREST)).astype(np.int32)

    r = rs(rows >> 4) * 16 + rs(rows & 15)
    c = rs(cols >> 4) * 16 + rs(cols & 15)
    assert (r == r[:, :1]).all() and (c == c[:1, :]).all()
    return r[:, 0].copy(), c[0, :].copy()


def resize_nearest(img, row_tab, col_tab):
    """imresize(img, (84,84), 'nearest') via index tables (atari_emulator.py:73)."""
    return img[np.asarray(row_tab)[:, None], np.asarray(col_tab)[None, :]]


def process_frame_pool(frame_pool, row_tab, col_tab):
    """atari_emulator.py:69-75 -- amax over the pooled frames, nearest resize, uint8."""
    img = np.amax(frame_pool, axis=0)
    return resize_nearest(img, row_tab, col_tab).astype(np.uint8)


class ObservationRing(object):
    """environment.py:58-75 restated without fancy indexing (same results)."""

    def __init__(self):
        self.pool = np.zeros((IMG, IMG, NR_IMAGES), np.uint8)
        self.idx = 0

    def new_observation(self, plane):                     # environment.py:66-68
        self.pool[:, :, self.idx] = plane
        self.idx = (self.idx + 1) % NR_IMAGES

    def get_pooled_observations(self):                    # environment.py:70-71
        order = [(self.idx + k) % NR_IMAGES for k in range(NR_IMAGES)]
        return self.pool[:, :, order].copy()


def step_states(prev_states, frames, reset, row_tab, col_tab):
    """One env-step of the observation pipeline for N envs, in the raw-frame protocol.

    prev_states uint8[N,84,84,4]   stack handed out after the previous step (oldest = channel 0)
    frames      uint8[N,4,2,210,160] raw luminance frame pairs.  Slot 0 holds this step's
                pair when reset[n] == 0.  When reset[n] != 0 the four slots hold the four
                action-repeat pairs of get_initial_state() (atari_emulator.py:88-96), oldest first.
    reset       uint8[N]
    returns     uint8[N,84,84,4]

    Non-reset: ObservationPool.new_observation + get_pooled_observations == drop channel 0,
    append the new plane as channel 3 (environment.py:66-71).
    Reset: the ring holds exactly the four new planes, oldest first (a full turn of the ring).
    """
    n = prev_states.shape[0]
    out = np.empty_like(prev_states)
    for i in range(n):
        if reset[i]:
            for k in range(NR_IMAGES):
                out[i, :, :, k] = process_frame_pool(frames[i, k], row_tab, col_tab)
        else:
            out[i, :, :, :3] = prev_states[i, :, :, 1:]
            out[i, :, :, 3] = process_frame_pool(frames[i, 0], row_tab, col_tab)
    return out
