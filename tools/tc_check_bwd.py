"""Per-tensor check of the tcgen05 backward against fp64 autograd and the SIMT path (run on a B200, under `timeout`)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tests'))
import numpy as np, torch
import gpu_util as G
from oracle import network
from util import rel_err

def run(arch, A, b, math):
    net = G.make_net(arch, A, seed=11)
    params = network.unflatten_params(net.get_params(), arch, A)
    rng = np.random.RandomState(b + A)
    states = rng.randint(0, 256, (b, 84, 84, 4)).astype(np.uint8)
    acts = rng.randint(0, A, b)
    adv = rng.randn(b).astype(np.float32); tgt = rng.randn(b).astype(np.float32)
    fwd = G.forward(net, states)
    masks = [x > 0 for x in G.layer_acts(net, fwd['ws'], b)]
    g64, dzs, f64 = network.masked_loss_and_grads(params, states, acts, adv, tgt, 0.02, arch, A, masks)
    _, dz, dv = network.closed_form_head_grads(f64['logits'], f64['v'], acts, adv, tgt, np.float32(0.02))
    simt, bws0 = G.backward(net, fwd, dz, dv)
    net.set_math(math)
    # same activations (the fp32 forward workspace) so that both backward paths see identical ReLU masks
    flat, bws1 = G.backward(net, fwd, dz, dv)
    got = network.unflatten_params(flat, arch, A); ref = network.unflatten_params(simt, arch, A)
    line = '%s A=%d b=%d %s:' % (arch, A, b, math)
    for name, _, _ in network.param_specs(arch, A):
        line += ' %s=%.1e(simt %.1e)' % (name.replace('_weights', '_w').replace('_biases', '_b').replace('_output', ''), rel_err(got[name], g64[name]), rel_err(ref[name], g64[name]))
    for tag, bws in (('simt', bws0), ('tc', bws1)):
        off = 0
        line += ' | dz(%s):' % tag
        for i, d in enumerate(dzs):
            n = d.size
            line += ' L%d=%.1e' % (i, rel_err(bws[off:off + n].cpu().numpy().reshape(d.shape), d))
            off += n
    print(line, flush=True)

for math in ('tf32x3', 'tf32'):
    for arch, A, b in (('NATURE', 6, 160), ('NIPS', 4, 97), ('NATURE', 4, 700), ('NATURE', 6, 1111)):
        run(arch, A, b, math)
print('tc_check_bwd done')
