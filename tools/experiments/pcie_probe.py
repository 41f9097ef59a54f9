"""How fast can the 84 + 84 selected rows of N raw frame pairs cross PCIe?  (development experiment)
  a) plain pinned H2D memcpy of a contiguous buffer of the same size                (DMA ceiling)
  b) cudaMemcpy2DAsync x2: rows 1 + 5k and 3 + 5k of every frame (width 160 B, pitch 800 B)
  c) the zero-copy preprocessing kernel as shipped (grid 96)
"""
import ctypes as C, sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from cuda.bindings import runtime as rt
from paac_b200 import _lib
from paac_b200.policy_v_network import NaturePolicyVNetwork

N = 4096
dev = torch.device('cuda', 0)
host = torch.randint(0, 256, (N, 1, 2, 210, 160), dtype=torch.uint8).pin_memory()
sel_bytes = N * 2 * 84 * 160
hc = torch.randint(0, 256, (sel_bytes,), dtype=torch.uint8).pin_memory()
dc = torch.empty((sel_bytes,), dtype=torch.uint8, device=dev)
d2 = torch.empty((N * 2 * 84, 160), dtype=torch.uint8, device=dev)
st = torch.cuda.current_stream(dev)

def timeit(fn, k=10):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(k):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / k

ms = timeit(lambda: dc.copy_(hc, non_blocking=True))
print('a) contiguous pinned H2D  %.1f MB: %.3f ms  %.1f GB/s' % (sel_bytes / 1e6, ms, sel_bytes / ms / 1e6))

def copy2d():
    rows = N * 2 * 42
    for j, r0 in enumerate((1, 3)):
        err, = rt.cudaMemcpy2DAsync(d2.data_ptr() + j * 160, 320, host.data_ptr() + r0 * 160, 800, 160, rows,
                                    rt.cudaMemcpyKind.cudaMemcpyHostToDevice, st.cuda_stream)
        assert err == rt.cudaError_t.cudaSuccess, err
ms = timeit(copy2d)
print('b) memcpy2D rows 1+5k, 3+5k: %.3f ms  %.1f GB/s' % (ms, sel_bytes / ms / 1e6))

conf = dict(name='x', num_actions=6, clip_norm=3.0, clip_norm_type='global', device='/gpu:0',
            entropy_regularisation_strength=0.02, seed=3, math='bf16x3')
net = NaturePolicyVNetwork(conf)
prev = torch.zeros((N, 84, 84, 4), dtype=torch.uint8, device=dev)
nxt = torch.zeros_like(prev)
def k1():
    _lib.check(net._lib.paacb_preprocess_u8(net.ctx, C.c_void_p(host.data_ptr()), 1, None, _lib.ptr(prev), _lib.ptr(nxt), N,
                                            C.c_void_p(st.cuda_stream)), 'k1')
ms = timeit(k1)
print('c) zero-copy K1 (grid 96): %.3f ms  %.1f GB/s over PCIe' % (ms, sel_bytes / ms / 1e6))
for pipe, g in ((1, 32), (1, 148), (1, 592), (2, 8), (2, 16), (2, 32), (2, 64), (2, 148)):
    # the knobs are read when a context is created: a fresh network per setting
    os.environ['PAACB_K1_PIPE'] = str(pipe)
    os.environ['PAACB_K1_HOST_GRID'] = str(g)
    os.environ['PAACB_K1_PIPE_HOST_GRID'] = str(g)
    net = NaturePolicyVNetwork(conf)
    try:
        ms = timeit(k1)
        print('   %s grid %d: %.3f ms  %.1f GB/s' % ('zero-copy K1 (CTA per env)' if pipe == 1 else 'zero-copy K1 through the TMA pipeline', g, ms, sel_bytes / ms / 1e6))
        if pipe == 2:
            ref = nxt.clone()
            os.environ['PAACB_K1_PIPE'] = '1'
            net = NaturePolicyVNetwork(conf); k1(); torch.cuda.synchronize()
            print('      same bytes as the CTA-per-env kernel:', bool(torch.equal(ref, nxt)))
    except Exception as e:
        print('   pipe=%d grid %d failed: %s' % (pipe, g, str(e)[:200]))
        break
