"""Session / Saver shims so reference-style call sites keep working.

The reference drives its TF graph through ``session.run(fetches, feed_dict)`` (paac.py:20-23, 140-142,
163-165) and checkpoints through ``tf.train.Saver`` (actor_learner.py:79-93).  ``Session.run`` here
understands exactly the fetches that appear on the hot path -- the network's ``output_layer_v`` /
``output_layer_pi`` given ``{network.input_ph: states}`` -- and dispatches them to the C-ABI forward.
``Saver`` stores the variables as a TensorFlow-1 tensor bundle under the reference's names (tf_bundle.py).
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
    """tf.train.Saver stand-in (actor_learner.py:79-93).  Checkpoints are TensorFlow-1 tensor bundles
    (``<folder>/-<step>.index`` + ``.data-00000-of-00001`` + the ``checkpoint`` state file, tf_bundle.py) holding the
    variables under the reference's names -- ``<name>_1/conv1_weights`` ... for the trunk, ``<name>_2/actor_output_*`` /
    ``critic_output_*`` for the heads, ``.../OptimizerVariables[_1]`` for the RMSProp slots (the network
    bundle carries them as well, like the reference's ``tf.train.Saver()``) -- so the folders are interchangeable with the
    reference's ``pretrained/*`` layout in both directions.  ``max_to_keep`` honoured; ``.pt`` files
    written by earlier versions are still restored."""

    def __init__(self, get_state, set_state, max_to_keep=5, name='Saver', scope='local_learning', optional=None):
        self.get_state, self.set_state, self.max_to_keep, self.name, self.scope = get_state, set_state, max_to_keep, name, scope
        self.optional = optional or (lambda key: False)      # keys a checkpoint may lack (restore passes what it finds)

    def _tf_name(self, key):
        base = key.split('/')[0]
        head = base.startswith('actor_output') or base.startswith('critic_output')
        return '%s_%d/%s' % (self.scope, 2 if head else 1, key)

    @staticmethod
    def _step_of(path):
        return int(os.path.basename(path)[1:].split('.')[0])

    @staticmethod
    def latest_checkpoint(folder):
        state = os.path.join(folder, 'checkpoint')
        if os.path.exists(state):
            for line in open(state):
                if line.startswith('model_checkpoint_path:'):
                    prefix = os.path.join(folder, line.split('"')[1])
                    if os.path.exists(prefix + '.index'):
                        return prefix
        files = [f[:-len('.index')] for f in glob.glob(os.path.join(folder, '-*.index'))]
        files += glob.glob(os.path.join(folder, '-*.pt'))
        if not files:
            return None
        return max(files, key=Saver._step_of)

    def save(self, session, folder, global_step):
        from . import tf_bundle
        os.makedirs(folder, exist_ok=True)
        prefix = os.path.join(folder, '-%d' % int(global_step))
        tf_bundle.write_bundle(prefix, {self._tf_name(k): v.numpy() for k, v in self.get_state().items()})
        tf_bundle.write_checkpoint_state(folder, '-%d' % int(global_step))
        kept = sorted((f[:-len('.index')] for f in glob.glob(os.path.join(folder, '-*.index'))), key=Saver._step_of)
        for old in kept[:-self.max_to_keep]:
            for ext in ('.index', '.data-00000-of-00001'):
                if os.path.exists(old + ext):
                    os.remove(old + ext)
        return prefix

    def restore(self, session, path):
        if path.endswith('.pt'):
            self.set_state(torch.load(path, map_location='cpu'))
            return
        from . import tf_bundle
        tensors = tf_bundle.read_bundle(path)
        wanted = self.get_state().keys()
        by_key = {name.split('/', 1)[1]: torch.from_numpy(a) for name, a in tensors.items() if '/' in name}
        missing = [k for k in wanted if k not in by_key and not self.optional(k)]
        if missing:
            raise KeyError('checkpoint %s lacks variables %s' % (path, missing[:4]))
        self.set_state({k: by_key[k] for k in wanted if k in by_key})
