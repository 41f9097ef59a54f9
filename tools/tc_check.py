"""Per-layer check of the tcgen05 path against the SIMT fp32 path and the oracle (run on a B200, under `timeout`)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tests'))
import numpy as np, torch
import gpu_util as G
from oracle import network
from util import rel_err

def run(arch, A, b, math):
    net = G.make_net(arch, A, seed=3)
    rng = np.random.RandomState(b)
    states = rng.randint(0, 256, (b, 84, 84, 4)).astype(np.uint8)
    ref = G.forward(net, states)
    ref_acts = G.layer_acts(net, ref['ws'], b)
    net.set_math(math)
    out = G.forward(net, states)
    acts = G.layer_acts(net, out['ws'], b)
    params = network.unflatten_params(net.get_params(), arch, A)
    orc = network.forward(params, states, arch, dtype=torch.float64, keep=True)
    oacts = [a.numpy() for a in orc['acts']] + [orc['h'].numpy()]
    line = '%s A=%d b=%d %s:' % (arch, A, b, math)
    for i, (a, r, o) in enumerate(zip(acts, ref_acts, oacts)):
        line += ' L%d vs_simt=%.2e vs_f64=%.2e (simt_vs_f64 %.2e)' % (i, rel_err(a, r), rel_err(a, o), rel_err(r, o))
    line += ' pi=%.2e v=%.2e' % (rel_err(out['pi'].cpu().numpy(), orc['pi'].numpy()), rel_err(out['v'].cpu().numpy(), orc['v'].numpy()))
    print(line, flush=True)
    if '--dump' in sys.argv:
        a, r = acts[0], ref_acts[0]
        bad = np.argwhere(np.abs(a - r) > 1e-3 * np.abs(r).max())
        print('  L0 mismatches:', len(bad), 'of', a.size, 'first:', bad[:8].tolist())
        print('  got', a.reshape(-1)[:8], '\n  ref', r.reshape(-1)[:8])

for math in ('tf32x3', 'tf32'):
    for arch, A, b in (('NATURE', 6, 3), ('NATURE', 6, 160), ('NIPS', 4, 33), ('NATURE', 6, 1111)):
        run(arch, A, b, math)
print('tc_check done')
