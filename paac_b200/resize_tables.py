"""Nearest-neighbour index tables for 210x160 -> 84x84, out[y, x] = img[ROW[y], COL[x]].

These are what Pillow's Image.resize((84, 84), NEAREST) -- the call behind the reference's
scipy.misc.imresize(img, (84, 84), interp='nearest') (atari_emulator.py:73) -- selects.  ROW follows
floor((y + 0.5) * 2.5); COL follows floor((x + 0.5) * 160 / 84) except x = 52 -> 99 and x = 73 -> 139
(float accumulation inside Pillow).  tests/test_preprocess_oracle.py pins them to the installed Pillow and to
the golden vectors; libpaacb.so carries the same defaults (csrc/api.cu) and accepts overrides through
paacb_set_resize_tables.
"""
import numpy as np

ROW = np.asarray([((2 * y + 1) * 5) // 4 for y in range(84)], dtype=np.int32)
COL = np.asarray([(2 * x + 1) * 160 // (2 * 84) for x in range(84)], dtype=np.int32)
COL[52] = 99
COL[73] = 139
