"""Where do the forward workspaces of PAACB_CONV3_PACKED=0 / 1 differ?  (development helper)"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import torch
import gpu_util as G

# per-sample element counts of conv1, conv2, conv3, fc (Nature): float offsets of the layers' regions = cumulative * cap
E = [20 * 20 * 32, 9 * 9 * 64, 7 * 7 * 64, 512]
for b, cap, first in ((1, 6, 3), (1, 1, 0), (2, 2, 0), (37, 37, 0), (130, 135, 3)):
    gen = torch.Generator(device='cuda'); gen.manual_seed(100 + b)
    states = torch.randint(0, 256, (b, 84, 84, 4), dtype=torch.uint8, device='cuda', generator=gen)
    out = []
    for knob in ('0', '1'):
        os.environ['PAACB_CONV3_PACKED'] = knob
        net = G.make_net('NATURE', 6, seed=21, math='bf16x3')
        pi = torch.full((b, 6), -1.0, device='cuda'); v = torch.full((b,), -1.0, device='cuda')
        ws = torch.zeros((net.workspace_floats(cap),), device='cuda')
        net.forward(states, pi, v, ws, ws_capacity=cap, ws_first=first)
        torch.cuda.synchronize()
        out.append((ws.view(torch.int16).clone(), pi, v))
    w0, w1 = out[0][0], out[1][0]
    off = 0
    print('b=%d cap=%d first=%d: pi equal %s v equal %s' % (b, cap, first, torch.equal(out[0][1], out[1][1]), torch.equal(out[0][2], out[1][2])))
    for li, e in enumerate(E):
        for plane, name in ((0, 'hi'), (1, 'lo')):
            lo = (off * cap * 2) + plane * e * cap       # int16 index: region base (floats -> 2 int16) + plane
            a, c = w0[lo:lo + e * cap], w1[lo:lo + e * cap]
            nd = (a != c).nonzero().flatten()
            msg = ''
            if nd.numel():
                i = int(nd[0]); s_ = i // e; r = i % e
                fa = torch.tensor([int(a[i]) << 16], dtype=torch.int32).view(torch.float32).item()
                fc = torch.tensor([int(c[i]) << 16], dtype=torch.int32).view(torch.float32).item()
                msg = ' first diff: sample slot %d elem %d (pos %d ch %d): %g vs %g' % (s_, r, r // (64 if li in (1, 2) else (32 if li == 0 else 512)), r % (64 if li in (1, 2) else (32 if li == 0 else 512)), fa, fc)
            print('  layer %d %s: %d of %d differ%s' % (li + 1, name, nd.numel(), e * cap, msg))
        base = off * cap * 2
        def val(w):
            hi = (w[base:base + e * cap].to(torch.int32) << 16).view(torch.float32)
            lo = (w[base + e * cap:base + 2 * e * cap].to(torch.int32) << 16).view(torch.float32)
            return hi.double() + lo.double()
        v0, v1 = val(w0), val(w1)
        d = (v0 - v1).abs()
        i = int(d.argmax())
        print('  layer %d values: max |diff| %.3e at elem %d (values %.9g vs %.9g), max |value| %.3e, rel-to-scale %.2e, mean |diff| %.2e' % (
            li + 1, d.max().item(), i % e, v0[i].item(), v1[i].item(), v0.abs().max().item(), d.max().item() / v0.abs().max().item(), d.mean().item()))
        off += e
