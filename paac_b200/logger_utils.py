"""args.json persistence: mirror of logger_utils.py:8-20, and the reference's summaries (logger_utils.py:23-33,
actor_learner.py:85-87, paac.py:130-135) as an OPT-IN JSONL stream (``train.py --summaries True``): same tags
(``summaries/raw_gradients/{mean,stddev,max,min}``, ``summaries/clipped_gradients/...``, ``global_norm``, ``rl/reward``,
``rl/episode_length``), one JSON object per line in ``<debugging_folder>/summaries.jsonl`` -- no TensorBoard dependency.
The reference computes them on every update inside session.run; here a record costs one extra reduction pass
(paacb_grad_stats) and is taken every ``every`` updates only."""
import ctypes as C
import json
import math
import os


def load_args(path):
    if path is None:
        return {}
    with open(path, 'r') as f:
        return json.load(f)


def save_args(args, folder, file_name='args.json'):
    args = vars(args)
    if not os.path.exists(folder):
        os.makedirs(folder)
    with open(os.path.join(folder, file_name), 'w') as f:
        return json.dump(args, f)


class StatsWriter(object):

    def __init__(self, folder, every=100, file_name='summaries.jsonl'):
        os.makedirs(folder, exist_ok=True)
        self.path = os.path.join(folder, file_name)
        self.f = open(self.path, 'a')
        self.every = max(1, int(every))
        self._out = None

    def _write(self, step, tag, value):
        self.f.write(json.dumps({'step': int(step), 'tag': tag, 'value': float(value)}) + '\n')

    def episode(self, step, reward, length):
        """paac.py:130-135"""
        self._write(step, 'rl/reward', reward)
        self._write(step, 'rl/episode_length', length)
        self.f.flush()

    def wants_gradients(self, update_index):
        return update_index % self.every == 0

    def gradients(self, step, engine, lr):
        """variable_summaries(flat_raw_gradients), (flat_clipped_gradients), global_norm -- actor_learner.py:85-87.
        Call between RolloutEngine.allreduce() and apply(): engine.grads holds the (summed) raw gradient."""
        import torch
        from . import _lib
        if self._out is None:
            self._out = torch.zeros(4, dtype=torch.float64, device=engine.dev)
        _lib.check(engine.lib.paacb_grad_stats(engine.ctx, _lib.ptr(engine.grads), C.c_float(1.0 / engine.world),
                                               _lib.ptr(engine.opt_ws), _lib.ptr(self._out), engine._stream()),
                   'paacb_grad_stats')
        s, q, mx, mn = [float(x) for x in self._out.cpu().tolist()]
        n = float(engine.grads.numel())
        mean = s / n
        std = math.sqrt(max(q / n - mean * mean, 0.0))
        norm = math.sqrt(q)
        scale = 1.0
        if engine.clip_type == _lib.CLIP_GLOBAL:
            scale = engine.clip_norm * min(1.0 / norm if norm > 0 else float('inf'), 1.0 / engine.clip_norm)
        for name, k in (('raw_gradients', 1.0), ('clipped_gradients', scale)):
            self._write(step, 'summaries/%s/mean' % name, mean * k)
            self._write(step, 'summaries/%s/stddev' % name, std * k)
            self._write(step, 'summaries/%s/max' % name, mx * k)
            self._write(step, 'summaries/%s/min' % name, mn * k)
        self._write(step, 'global_norm', norm)
        self._write(step, 'learning_rate', lr)
        self.f.flush()
        return dict(mean=mean, stddev=std, max=mx, min=mn, norm=norm, clip_scale=scale)

    def close(self):
        self.f.close()
