"""Oracle (TEST INFRASTRUCTURE ONLY): interpreter for the reference's shipped TF-1.0.1 training graph.

The reference's model/loss/gradient/clip/RMSProp arithmetic is defined by a TensorFlow graph that
``actor_learner.py:31-70`` + ``networks.py`` + ``policy_v_network.py`` build at start-up.  TensorFlow is not
installable here, but the reference ships that very graph, serialized by its own ``tf.train.Saver``
(actor_learner.py:79,92), as ``pretrained/<game>/checkpoints/-80000000.meta`` (TF 1.0.1, NIPS arch).
This module evaluates that GraphDef node by node in NumPy fp32 (convolutions via torch-CPU):
the forward pass, the loss, TF's own autodiff subgraph (``gradients/...``), the
``clip_by_global_norm`` subgraph and the ten ``ApplyRMSProp`` nodes, exactly as wired in the file.
What is restated here is only the per-op arithmetic (documented TF op definitions); the graph
structure, op order and every constant come from the reference artefact.

It is used (a) by ``oracle/make_golden.py`` to emit ``tests/golden/tf_graph_*.npz`` and (b) by CPU tests that
check ``oracle/network.py`` + ``oracle/update.py`` against it.  It reads ``/root/reference`` and therefore never
runs on the GPU box; only its committed outputs travel.
"""
import numpy as np
import torch
import torch.nn.functional as F

_F32 = np.float32


def load_meta(path):
    from tensorboard.compat.proto import meta_graph_pb2
    m = meta_graph_pb2.MetaGraphDef()
    with open(path, 'rb') as f:
        m.ParseFromString(f.read())
    return m


def _const(node):
    from tensorboard.util import tensor_util
    return np.asarray(tensor_util.make_ndarray(node.attr['value'].tensor))


def _nhwc_conv(x, w, strides):
    xt = torch.from_numpy(np.ascontiguousarray(x)).permute(0, 3, 1, 2)
    wt = torch.from_numpy(np.ascontiguousarray(w)).permute(3, 2, 0, 1)
    y = F.conv2d(xt, wt, stride=(strides[1], strides[2]))
    return y.permute(0, 2, 3, 1).contiguous().numpy()


def _nhwc_conv_bwd_input(input_sizes, w, dy, strides):
    n, h, wd, c = [int(v) for v in input_sizes]
    wt = torch.from_numpy(np.ascontiguousarray(w)).permute(3, 2, 0, 1).contiguous()
    dyt = torch.from_numpy(np.ascontiguousarray(dy)).permute(0, 3, 1, 2).contiguous()
    dx = torch.nn.grad.conv2d_input((n, c, h, wd), wt, dyt, stride=(strides[1], strides[2]))
    return dx.permute(0, 2, 3, 1).contiguous().numpy()


def _nhwc_conv_bwd_filter(x, filter_sizes, dy, strides):
    kh, kw, ci, co = [int(v) for v in filter_sizes]
    xt = torch.from_numpy(np.ascontiguousarray(x)).permute(0, 3, 1, 2).contiguous()
    dyt = torch.from_numpy(np.ascontiguousarray(dy)).permute(0, 3, 1, 2).contiguous()
    dw = torch.nn.grad.conv2d_weight(xt, (co, ci, kh, kw), dyt, stride=(strides[1], strides[2]))
    return dw.permute(2, 3, 1, 0).contiguous().numpy()


class GraphInterpreter(object):
    """Lazy evaluator over a TF GraphDef.  ``variables``: dict VariableV2-name -> ndarray (mutated by
    ``run_train_step``); ``feeds``: dict placeholder-name -> ndarray."""

    def __init__(self, meta):
        self.graph = meta.graph_def
        self.nodes = {n.name: n for n in self.graph.node}
        self.ops_used = set()

    # ---- naming helpers -------------------------------------------------------------------------
    def variable_names(self):
        return [n.name for n in self.graph.node if n.op == 'VariableV2']

    def variable_shape(self, name):
        return tuple(d.size for d in self.nodes[name].attr['shape'].shape.dim)

    def slot_initial_value(self, slot_name):
        """Value the graph's initializer assigns to an optimizer slot (Const ones / zeros, App. B)."""
        for n in self.graph.node:
            if n.op == 'Assign' and n.input[0] == slot_name:
                src = self.nodes[n.input[1]]
                assert src.op == 'Const', (slot_name, src.op)
                return np.broadcast_to(_const(src), self.variable_shape(slot_name)).astype(np.float32).copy()
        raise KeyError(slot_name)

    # ---- evaluation -----------------------------------------------------------------------------
    def run(self, fetches, feeds, variables):
        self._memo, self._feeds, self._vars = {}, feeds, variables
        return [self._eval(f) for f in fetches]

    def _eval(self, ref):
        if ref.startswith('^'):
            return None
        name, _, port = ref.partition(':')
        port = int(port) if port else 0
        if name not in self._memo:
            self._memo[name] = self._compute(self.nodes[name])
        out = self._memo[name]
        return out[port] if isinstance(out, tuple) else out

    def _compute(self, n):
        op = n.op
        self.ops_used.add(op)
        ins = [i for i in n.input if not i.startswith('^')]
        if op == 'Placeholder':
            return np.asarray(self._feeds[n.name])
        if op == 'VariableV2':
            return self._vars[n.name]
        if op == 'Const':
            return _const(n)
        a = [self._eval(i) for i in ins]
        if op in ('Identity', 'StopGradient'):
            return a[0]
        if op == 'NoOp':
            return None
        if op == 'Cast':
            return a[0].astype({1: np.float32, 3: np.int32, 4: np.uint8, 9: np.int64, 10: bool}[n.attr['DstT'].type])
        if op == 'Add':
            return a[0] + a[1]
        if op == 'AddN':
            s = a[0]
            for t in a[1:]:
                s = s + t
            return s
        if op == 'Sub':
            return a[0] - a[1]
        if op == 'Mul':
            return a[0] * a[1]
        if op == 'RealDiv':
            return a[0] / a[1]
        if op == 'FloorDiv':
            return np.floor_divide(a[0], a[1])
        if op == 'FloorMod':
            return np.mod(a[0], a[1])
        if op == 'Neg':
            return -a[0]
        if op == 'Reciprocal':
            return (_F32(1.0) / a[0]).astype(a[0].dtype)
        if op == 'Maximum':
            return np.maximum(a[0], a[1])
        if op == 'Minimum':
            return np.minimum(a[0], a[1])
        if op == 'Greater':
            return a[0] > a[1]
        if op == 'Select':
            return np.where(a[0], a[1], a[2])
        if op == 'ZerosLike':
            return np.zeros_like(a[0])
        if op == 'Log':
            with np.errstate(divide='ignore', invalid='ignore'):
                return np.log(a[0])
        if op == 'Sqrt':
            return np.sqrt(a[0])
        if op == 'Pow':
            return np.power(a[0], a[1])
        if op == 'Relu':
            return np.maximum(a[0], _F32(0))
        if op == 'ReluGrad':                                   # (gradients, features)
            return np.where(a[1] > 0, a[0], _F32(0)).astype(np.float32)
        if op == 'Softmax':
            z = a[0] - a[0].max(axis=-1, keepdims=True)
            e = np.exp(z)
            return (e / e.sum(axis=-1, keepdims=True, dtype=np.float32)).astype(np.float32)
        if op == 'L2Loss':
            return _F32(np.sum(a[0] * a[0], dtype=np.float32) / _F32(2))
        if op == 'Shape':
            return np.asarray(np.shape(a[0]), np.int32)
        if op == 'Reshape':
            return np.reshape(a[0], [int(v) for v in np.atleast_1d(a[1])])
        if op == 'Fill':
            return np.full([int(v) for v in np.atleast_1d(a[0])], a[1], dtype=np.asarray(a[1]).dtype)
        if op == 'Range':
            return np.arange(int(a[0]), int(a[1]), int(a[2]), dtype=np.int32)
        if op == 'Prod':
            return np.prod(a[0], axis=tuple(int(v) for v in np.atleast_1d(a[1])), keepdims=bool(n.attr['keep_dims'].b)).astype(a[0].dtype)
        if op in ('Sum', 'Mean'):
            axes = tuple(int(v) for v in np.atleast_1d(a[1]))
            fn = np.sum if op == 'Sum' else np.mean
            if len(axes) == 0:
                return a[0]
            return fn(a[0], axis=axes, keepdims=bool(n.attr['keep_dims'].b), dtype=a[0].dtype).astype(a[0].dtype)
        if op == 'Tile':
            return np.tile(a[0], [int(v) for v in np.atleast_1d(a[1])])
        if op == 'Pack':
            return np.stack(a, axis=int(n.attr['axis'].i))
        if op == 'DynamicStitch':
            k = int(n.attr['N'].i)
            idx, data = a[:k], a[k:]
            size = max(int(np.max(i)) for i in idx) + 1
            out = np.zeros(size, dtype=np.asarray(data[0]).dtype)
            for i, d in zip(idx, data):
                out[np.asarray(i).reshape(-1)] = np.broadcast_to(d, np.shape(i)).reshape(-1)
            return out
        if op == 'BroadcastGradientArgs':
            s0, s1 = list(a[0]), list(a[1])
            r = max(len(s0), len(s1))
            p0 = [1] * (r - len(s0)) + s0
            p1 = [1] * (r - len(s1)) + s1
            r0 = [i for i in range(r) if p0[i] == 1 and (p1[i] != 1 or i < r - len(s0))]
            r1 = [i for i in range(r) if p1[i] == 1 and (p0[i] != 1 or i < r - len(s1))]
            return (np.asarray(r0, np.int32), np.asarray(r1, np.int32))
        if op == 'MatMul':
            x = a[0].T if n.attr['transpose_a'].b else a[0]
            y = a[1].T if n.attr['transpose_b'].b else a[1]
            return (torch.from_numpy(np.ascontiguousarray(x)) @ torch.from_numpy(np.ascontiguousarray(y))).numpy()
        if op == 'Conv2D':
            assert n.attr['padding'].s == b'VALID' and n.attr['data_format'].s == b'NHWC'
            return _nhwc_conv(a[0], a[1], list(n.attr['strides'].list.i))
        if op == 'Conv2DBackpropInput':
            return _nhwc_conv_bwd_input(a[0], a[1], a[2], list(n.attr['strides'].list.i))
        if op == 'Conv2DBackpropFilter':
            return _nhwc_conv_bwd_filter(a[0], a[1], a[2], list(n.attr['strides'].list.i))
        raise NotImplementedError(op + ' (' + n.name + ')')

    # ---- the train step -------------------------------------------------------------------------
    def run_train_step(self, feeds, variables, extra_fetches=()):
        """Evaluate every ApplyRMSProp node (what ``session.run(train_step)`` does, paac.py:163-165).
        ``variables`` is updated in place after all gradients are evaluated (TF reads then assigns;
        each variable is read only by its own apply op after the gradient graph has run)."""
        applies = [n for n in self.graph.node if n.op == 'ApplyRMSProp']
        self.run([], feeds, variables)
        extras = [self._eval(f) for f in extra_fetches]
        pending = []
        for n in applies:
            var_n, ms_n, mom_n = n.input[0], n.input[1], n.input[2]
            lr, rho, momentum, eps, grad = [self._eval(i) for i in n.input[3:8]]
            f = np.float32
            ms = variables[ms_n] + (grad * grad - variables[ms_n]) * (f(1) - f(rho))
            mom = variables[mom_n] * f(momentum) + (grad * f(lr)) / np.sqrt(ms + f(eps))
            var = variables[var_n] - mom
            pending.append((var_n, var.astype(f), ms_n, ms.astype(f), mom_n, mom.astype(f)))
        for var_n, var, ms_n, ms, mom_n, mom in pending:
            variables[var_n], variables[ms_n], variables[mom_n] = var, ms, mom
        self.ops_used.add('ApplyRMSProp')
        return extras
