"""Runners: mirror of runners.py:7-50 (shared buffers, fan-out / fan-in over worker processes).

Same constructor and methods.  Differences, both deliberate and documented:
  * uint8 arrays are shared as true uint8 (the reference's NUMPY_TO_C_DTYPE maps np.uint8 -> c_uint, a
    4x wider buffer holding the same values);
  * ``pin()`` page-locks the shared buffers and maps them into the GPU address space
    (paacb_host_register), so the learner's kernels read what the workers wrote without a staging copy.
"""
import ctypes as C
from ctypes import c_ubyte, c_float, c_double, c_int32
from multiprocessing import Queue
from multiprocessing.sharedctypes import RawArray

import numpy as np


class Runners(object):

    NUMPY_TO_C_DTYPE = {np.float32: c_float, np.float64: c_double, np.uint8: c_ubyte, np.int32: c_int32}

    def __init__(self, EmulatorRunner, emulators, workers, variables):
        self.variables = [self._get_shared(var) for var in variables]
        self.workers = workers
        self.queues = [Queue() for _ in range(workers)]
        self.barrier = Queue()
        self._pinned = []

        self.runners = [EmulatorRunner(i, emulators, vars, self.queues[i], self.barrier) for i, (emulators, vars) in
                        enumerate(zip(np.split(emulators, workers), zip(*[np.split(var, workers) for var in self.variables])))]

    def _get_shared(self, array):
        """
        Returns a RawArray backed numpy array that can be shared between processes.
        :param array: the array to be shared
        :return: the RawArray backed numpy array
        """
        dtype = self.NUMPY_TO_C_DTYPE[array.dtype.type]
        shape = array.shape
        shared = RawArray(dtype, array.size)
        view = np.frombuffer(shared, dtype).reshape(shape)
        view[...] = array
        return view

    def start(self):
        for r in self.runners:
            r.start()

    def stop(self):
        for queue in self.queues:
            queue.put(None)
        self.unpin()

    def get_shared_variables(self):
        return self.variables

    def update_environments(self):
        for queue in self.queues:
            queue.put(True)

    def wait_updated(self):
        for wd in range(self.workers):
            self.barrier.get()

    # ---- B200 addition: pinned + mapped shared buffers -------------------------------------------------
    def pin(self, index):
        """Page-lock variables[index] and return a CUDA uint8/float tensor aliasing it (zero-copy)."""
        import torch
        from . import _lib
        lib = _lib.load()
        arr = self.variables[index]
        dptr = C.c_void_p()
        _lib.check(lib.paacb_host_register(C.c_void_p(arr.ctypes.data), arr.nbytes, C.byref(dptr)),
                   'paacb_host_register')
        self._pinned.append(arr.ctypes.data)
        return dptr.value

    def unpin(self):
        if not self._pinned:
            return
        from . import _lib
        lib = _lib.load()
        for p in self._pinned:
            lib.paacb_host_unregister(C.c_void_p(p))
        self._pinned = []
