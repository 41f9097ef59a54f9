"""Session / Saver shims so reference-style call sites keep working.

The reference drives its TF graph through ``session.run(fetches, feed_dict)`` (paac.py:20-23, 140-142,
163-165) and checkpoints through ``tf.train.Saver`` (actor_learner.py:79-93).  ``Session.run`` here
understands exactly the fetches that appear on the hot path -- the network's ``output_layer_v`` /
``output_layer_pi`` given ``{network.input_ph: states}`` -- and dispatches them to the C-ABI forward.
``Saver`` stores the flat parameter buffer with the TF variable names via torch.save.
"""
import glob
import os

import numpy as np
import torch


class Session(object):

    def __init__(self, config=None):
        self.config = config
        self.closed = False

    def run(self, fetches, feed_dict=None):
        single = not isinstance(fetches, (list, tuple))
        flist = [fetches] if single else list(fetches)
        feed_dict = feed_dict or {}
        net = flist[0].owner
        if any(f.kind not in ('pi', 'v') for f in flist):
            raise NotImplementedError('Session.run supports the hot-path fetches output_layer_pi / output_layer_v; '
                                      'training runs through PAACLearner.train()')
        states = feed_dict[net.input_ph]
        pi, v = forward_numpy(net, states)
        out = [pi if f.kind == 'pi' else v for f in flist]
        return out[0] if single else out

    def close(self):
        self.closed = True


def forward_numpy(net, states, uniforms=None):
    """states: uint8 [b,84,84,4] host array (or CUDA tensor) -> (pi [b,A], v [b]) numpy (+ actions if uniforms)."""
    dev = net.torch_device
    if torch.is_tensor(states):
        st = states.to(dev)
    else:
        st = torch.from_numpy(np.ascontiguousarray(np.asarray(states, dtype=np.uint8))).to(dev)
    b = st.shape[0]
    pi = torch.empty((b, net.num_actions), dtype=torch.float32, device=dev)
    v = torch.empty((b,), dtype=torch.float32, device=dev)
    ws = torch.empty((net.workspace_floats(b),), dtype=torch.float32, device=dev)
    if uniforms is None:
        net.forward(st, pi, v, ws)
        return pi.cpu().numpy(), v.cpu().numpy()
    u = torch.as_tensor(np.asarray(uniforms, np.float32)).to(dev)
    act = torch.empty((b,), dtype=torch.int32, device=dev)
    net.forward(st, pi, v, ws, uniforms=u, actions=act)
    return pi.cpu().numpy(), v.cpu().numpy(), act.cpu().numpy()


class Saver(object):
    """tf.train.Saver stand-in: ``<folder>-<step>.pt`` files, newest wins, ``max_to_keep`` honoured."""

    def __init__(self, get_state, set_state, max_to_keep=5, name='Saver'):
        self.get_state, self.set_state, self.max_to_keep, self.name = get_state, set_state, max_to_keep, name

    @staticmethod
    def latest_checkpoint(folder):
        files = glob.glob(os.path.join(folder, '-*.pt'))
        if not files:
            return None
        return max(files, key=lambda f: int(os.path.basename(f)[1:].split('.')[0]))

    def save(self, session, folder, global_step):
        os.makedirs(folder, exist_ok=True)
        path = os.path.join(folder, '-%d.pt' % int(global_step))
        torch.save(self.get_state(), path)
        files = sorted(glob.glob(os.path.join(folder, '-*.pt')), key=lambda f: int(os.path.basename(f)[1:].split('.')[0]))
        for old in files[:-self.max_to_keep]:
            os.remove(old)
        return path

    def restore(self, session, path):
        self.set_state(torch.load(path, map_location='cpu'))
