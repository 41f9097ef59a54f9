"""Does the acting forward get cheaper when its activations stay in L2?  A 4096-sample forward writes 210 MB of conv1 output
(two bf16 planes) and reads it back in conv2: more than the 126 MB L2.  Here the same forward runs as C chunks of 4096 / C
samples through ONE chunk-sized workspace (the activations of an acting / bootstrap forward are never read again), and the
per-kernel device time (CUDA events, paacb_profile_*) is compared with the single launch.
    python tools/experiments/chunked_forward.py > gpurun_out/chunked_forward.json
"""
import ctypes as C
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch

from paac_b200 import _lib
from paac_b200.policy_v_network import NaturePolicyVNetwork, NIPSPolicyVNetwork


def profile(net):
    out = {}
    for s in range(net._lib.paacb_profile_slots()):
        name = C.create_string_buffer(64)
        ms, cnt = C.c_double(), C.c_int64()
        net._lib.paacb_profile_read(net.ctx, s, name, 64, C.byref(ms), C.byref(cnt))
        if cnt.value > 0:
            out[name.value.decode()] = ms.value
    return out


def main():
    N, A, reps = 4096, 6, 20
    rows = []
    for arch in ('NATURE', 'NIPS'):
        conf = dict(name='x', num_actions=A, clip_norm=3.0, clip_norm_type='global', device='/gpu:0',
                    entropy_regularisation_strength=0.02, seed=3, math='bf16x3')
        net = (NaturePolicyVNetwork if arch == 'NATURE' else NIPSPolicyVNetwork)(conf)
        gen = torch.Generator(device='cuda'); gen.manual_seed(1)
        pool = [torch.randint(0, 256, (N, 84, 84, 4), dtype=torch.uint8, device='cuda', generator=gen) for _ in range(6)]
        pi = torch.empty((N, A), device='cuda'); v = torch.empty((N,), device='cuda')
        for chunks in (1, 2, 4, 8):
            n = N // chunks
            ws = torch.empty((net.workspace_floats(n),), device='cuda')

            def run(k):
                st = pool[k % len(pool)]
                for c in range(chunks):
                    net.forward(st[c * n:(c + 1) * n], pi[c * n:(c + 1) * n], v[c * n:(c + 1) * n], ws)
            for k in range(3):
                run(k)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for k in range(reps):
                run(k)
            e1.record(); torch.cuda.synchronize()
            total = e0.elapsed_time(e1) / reps
            _lib.check(net._lib.paacb_profile_enable(net.ctx, 1)); _lib.check(net._lib.paacb_profile_reset(net.ctx))
            for k in range(reps):
                run(k)
            torch.cuda.synchronize()
            prof = {k: round(ms / reps, 4) for k, ms in profile(net).items()}
            _lib.check(net._lib.paacb_profile_enable(net.ctx, 0))
            rows.append({'arch': arch, 'chunks': chunks, 'samples_per_chunk': n, 'ms_per_forward': round(total, 4), 'kernels_ms': prof,
                         'workspace_mb': round(ws.numel() * 4 / 1e6, 1)})
            print(rows[-1], file=sys.stderr)
    print(json.dumps({'what': 'one 4096-sample forward (bf16x3) as C chunks through one chunk-sized workspace', 'rows': rows}, indent=1))


if __name__ == '__main__':
    main()
