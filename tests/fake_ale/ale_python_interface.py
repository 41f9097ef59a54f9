"""TEST INFRASTRUCTURE: a deterministic stand-in for the Arcade Learning Environment's Python binding.

ALE (a C++ emulator + ROMs) cannot be installed here.  This module has the same import name and the part of the
``ALEInterface`` API that arjunchandra/paac's atari_emulator.py calls, so that BOTH the reference's AtariEmulator (run
unchanged from /root/reference by oracle/make_golden.py) and paac_b200's AtariEmulator can be executed against the same
"game" and compared frame for frame.  It is a scripted game, not an emulator: the screen is a function of (seed, frame
counter, actions so far) -- a dark 210x160 luminance screen with a few moving sprites, so that the two-frame max-pool, the
nearest resize and the stack order all matter -- rewards, lives and game-over follow the frame counter.
"""
import numpy as np


class ALEInterface(object):
    WIDTH, HEIGHT = 160, 210

    def __init__(self):
        self.settings = {}
        self.rom = None
        self.reset_game()

    # ---- settings / ROM ----------------------------------------------------------------------------
    def setInt(self, key, value):
        self.settings[bytes(key)] = int(value)

    def setFloat(self, key, value):
        self.settings[bytes(key)] = float(value)

    def setBool(self, key, value):
        self.settings[bytes(key)] = bool(value)

    def loadROM(self, path):
        self.rom = bytes(path)
        self.reset_game()

    def getMinimalActionSet(self):
        return np.array([0, 1, 3, 4], dtype=np.int32)          # NOOP, FIRE, RIGHT, LEFT (Breakout's minimal set)

    def getScreenDims(self):
        return self.WIDTH, self.HEIGHT

    # ---- game ------------------------------------------------------------------------------------------
    def reset_game(self):
        self.frame = 0
        self.episode = getattr(self, 'episode', -1) + 1
        self.paddle = 80
        self._lives = 3
        self.acted = 0            # checksum of the actions of this episode (the screen depends on it)

    def lives(self):
        return self._lives

    def game_over(self):
        return self._lives <= 0

    def act(self, action):
        action = int(action)
        self.frame += 1
        self.acted = (self.acted * 31 + action + 1) % 1000003
        if action == 3:
            self.paddle = min(self.paddle + 3, 150)
        elif action == 4:
            self.paddle = max(self.paddle - 3, 2)
        if self.frame % 61 == 0:
            self._lives -= 1
        return (action + 1) if self.frame % 13 == 0 else 0

    # ---- screen ------------------------------------------------------------------------------------------
    def _screen(self):
        seed = self.settings.get(b'random_seed', 0)
        s = np.zeros((self.HEIGHT, self.WIDTH), dtype=np.uint8)
        f = self.frame
        # a ball that moves every frame (consecutive frames differ: the two-frame max-pool has something to pool)
        by, bx = (7 * f + 3 * seed) % 200, (11 * f + self.acted) % 150
        s[by:by + 4, bx:bx + 3] = 200 + (f % 50)
        # a flickering row of bricks: present on even frames only (what the max-pool is there for)
        if f % 2 == 0:
            s[57:63, 8:152:4] = 90 + (self.episode % 7) * 10
        # the paddle
        s[189:193, self.paddle - 2:self.paddle + 8] = 143
        # score digits depend on the action history
        s[5:14, 20 + (self.acted % 40):24 + (self.acted % 40)] = 236
        # single bright pixels on rows / columns next to the nearest-resize sampling points
        s[(f * 5) % 210, (f * 3) % 160] = 255
        s[209, 159] = f % 256
        return s

    def getScreenGrayscale(self, buf):
        buf[...] = self._screen().reshape(buf.shape)

    def getScreenRGB(self, buf):
        buf[...] = self._screen()[..., None]
