"""Per-layer / per-tensor check of the bf16x3 path against the fp64 oracle (run on a B200, under `timeout`)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tests'))
import numpy as np, torch
import gpu_util as G
from oracle import network
from util import rel_err


def where_bad(a, r, tag):
    scale = np.abs(r).max()
    bad = np.argwhere(np.abs(a - r) > 1e-3 * scale)
    if len(bad) == 0:
        return
    print('    %s: %d of %d elements off (scale %.3g, nan %d); first %s' % (tag, len(bad), a.size, scale, int(np.isnan(a).sum()), bad[:6].tolist()))
    for ax in range(a.ndim):
        u = np.unique(bad[:, ax])
        print('      axis %d: %d distinct indices of %d: %s' % (ax, len(u), a.shape[ax], u[:24].tolist()))
    i = tuple(bad[0])
    print('      got %r want %r' % (float(a[i]), float(r[i])))


def fwd(arch, A, b):
    net = G.make_net(arch, A, seed=3, math='bf16x3')
    rng = np.random.RandomState(b)
    states = rng.randint(0, 256, (b, 84, 84, 4)).astype(np.uint8)
    out = G.forward(net, states)
    acts = G.layer_acts(net, out['ws'], b)
    params = network.unflatten_params(net.get_params(), arch, A)
    orc = network.forward(params, states, arch, dtype=torch.float64, keep=True)
    oacts = [a.numpy() for a in orc['acts']] + [orc['h'].numpy()]
    line = 'FWD %s A=%d b=%d:' % (arch, A, b)
    for i, (a, o) in enumerate(zip(acts, oacts)):
        line += ' L%d=%.2e' % (i, rel_err(a, o))
    line += ' pi=%.2e v=%.2e' % (rel_err(out['pi'].cpu().numpy(), orc['pi'].numpy()), rel_err(out['v'].cpu().numpy(), orc['v'].numpy()))
    print(line, flush=True)
    for i, (a, o) in enumerate(zip(acts, oacts)):
        if not rel_err(a, o) < 1e-4:
            where_bad(a, o, 'L%d' % i)
            break


def bwd(arch, A, b):
    net = G.make_net(arch, A, seed=11, math='bf16x3')
    params = network.unflatten_params(net.get_params(), arch, A)
    rng = np.random.RandomState(b + A)
    states = rng.randint(0, 256, (b, 84, 84, 4)).astype(np.uint8)
    acts = rng.randint(0, A, b)
    adv = rng.randn(b).astype(np.float32); tgt = rng.randn(b).astype(np.float32)
    f = G.forward(net, states)
    masks = [x > 0 for x in G.layer_acts(net, f['ws'], b)]
    g64, dzs, f64 = network.masked_loss_and_grads(params, states, acts, adv, tgt, 0.02, arch, A, masks)
    _, dz, dv = network.closed_form_head_grads(f64['logits'], f64['v'], acts, adv, tgt, np.float32(0.02))
    flat, bws = G.backward(net, f, dz, dv)
    got = network.unflatten_params(flat, arch, A)
    line = 'BWD %s A=%d b=%d:' % (arch, A, b)
    for name, _, _ in network.param_specs(arch, A):
        line += ' %s=%.1e' % (name.replace('_weights', '_w').replace('_biases', '_b').replace('_output', ''), rel_err(got[name], g64[name]))
    gdz = G.layer_acts(net, bws, b)
    line += ' | dz:'
    for i, d in enumerate(dzs):
        line += ' L%d=%.1e' % (i, rel_err(gdz[i].reshape(d.shape), d))
    print(line, flush=True)
    for i, d in enumerate(dzs):
        if not rel_err(gdz[i].reshape(d.shape), d) < 1e-4:
            where_bad(gdz[i].reshape(d.shape), d, 'dZ L%d' % i)
    for name, _, _ in network.param_specs(arch, A):
        if not rel_err(got[name], g64[name]) < 1e-4:
            where_bad(np.asarray(got[name]), np.asarray(g64[name]), name)


which = sys.argv[1] if len(sys.argv) > 1 else 'all'
if which in ('fwd', 'all'):
    for arch, A, b in (('NATURE', 6, 1), ('NATURE', 6, 3), ('NATURE', 6, 160), ('NATURE', 18, 517), ('NATURE', 6, 1111)):
        fwd(arch, A, b)
if which in ('bwd', 'all'):
    for arch, A, b in (('NATURE', 6, 5), ('NATURE', 6, 160), ('NATURE', 4, 700), ('NATURE', 6, 1111)):
        bwd(arch, A, b)
print('bf16_check done')
