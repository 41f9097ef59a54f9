"""One training forward + backward of the bf16x3 path at batch 20,480 (twice: the first pass is warm-up), for ncu."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tests'))
import numpy as np, torch
import gpu_util as G
from paac_b200 import _lib

b = int(sys.argv[1]) if len(sys.argv) > 1 else 20480
math = sys.argv[2] if len(sys.argv) > 2 else 'bf16x3'
net = G.make_net('NATURE', 6, seed=3, math=math)
A = 6
st = torch.randint(0, 256, (b, 84, 84, 4), dtype=torch.uint8, device='cuda')
pi = torch.empty((b, A), device='cuda'); v = torch.empty((b,), device='cuda')
ws = torch.zeros((net.workspace_floats(b),), device='cuda')
bws = torch.zeros((int(net._lib.paacb_backward_workspace_floats(net.ctx, b)),), device='cuda')
grads = torch.zeros((net.param_count,), device='cuda')
dl = torch.randn((b, A), device='cuda') * 1e-4; dv = torch.randn((b,), device='cuda') * 1e-4
p = _lib.ptr
import ctypes as C
ne = min(b, 4096)
frames = torch.randint(0, 256, (ne, 1, 2, 210, 160), dtype=torch.uint8, device='cuda')
nxt = torch.empty((ne, 84, 84, 4), dtype=torch.uint8, device='cuda')
for it in range(2):
    if it == 1:
        torch.cuda.synchronize()
        torch.cuda.profiler.start()          # ncu --profile-from-start off: capture exactly the second pass
    _lib.check(net._lib.paacb_preprocess_u8(net.ctx, p(frames), 1, None, p(st[:ne]), p(nxt), ne,
                                            C.c_void_p(torch.cuda.current_stream().cuda_stream)), 'k1')
    net.forward(st, pi, v, ws)
    _lib.check(net._lib.paacb_backward(net.ctx, p(net.params), p(st), b, p(ws), p(dl), p(dv), p(bws), p(grads),
                                       C.c_void_p(torch.cuda.current_stream().cuda_stream)), 'bwd')
    torch.cuda.synchronize()
torch.cuda.profiler.stop()
print('profile_step done', float(grads.abs().sum()))
