"""EmulatorRunner: mirror of emulator_runner.py:4-33 (one worker process stepping its slice of envs).

Protocol (unchanged): ``variables = [states, rewards, episode_over, actions]`` are views of shared memory
owned by the learner; the worker blocks on ``queue.get()``, ``None`` means exit, anything else means "step
every emulator once with ``variables[-1][i]``", then it writes the results in place and puts ``True`` on
``barrier``.  On a terminal step the state written is the NEW episode's initial state (:26-27).

``RawFrameEmulatorRunner`` is the raw-frame variant of the same protocol: ``variables[0]`` is the frame
slot array uint8[n, 4, 2, 210, 160] instead of stacked observations; slot 0 receives this step's frame pair,
and on a terminal step all four slots receive the new episode's initial pairs.  ``episode_over`` doubles as
the reset flag the GPU preprocessing kernel consumes.
"""
from multiprocessing import Process


class EmulatorRunner(Process):

    def __init__(self, id, emulators, variables, queue, barrier):
        super(EmulatorRunner, self).__init__()
        self.id = id
        self.emulators = emulators
        self.variables = variables
        self.queue = queue
        self.barrier = barrier

    def run(self):
        super(EmulatorRunner, self).run()
        self._run()

    def _run(self):
        count = 0
        while True:
            instruction = self.queue.get()
            if instruction is None:
                break
            for i, (emulator, action) in enumerate(zip(self.emulators, self.variables[-1])):
                new_s, reward, episode_over = emulator.next(action)
                if episode_over:
                    self.variables[0][i] = emulator.get_initial_state()
                else:
                    self.variables[0][i] = new_s
                self.variables[1][i] = reward
                self.variables[2][i] = episode_over
            count += 1
            self.barrier.put(True)


class RawFrameEmulatorRunner(EmulatorRunner):

    def _run(self):
        count = 0
        while True:
            instruction = self.queue.get()
            if instruction is None:
                break
            for i, (emulator, action) in enumerate(zip(self.emulators, self.variables[-1])):
                reward, episode_over = emulator.next_raw(action, self.variables[0][i])
                if episode_over:
                    emulator.get_initial_state_raw(self.variables[0][i])
                self.variables[1][i] = reward
                self.variables[2][i] = episode_over
            count += 1
            self.barrier.put(True)
