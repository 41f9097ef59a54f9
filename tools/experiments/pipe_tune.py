"""Role sizes of the layer-pipelined forward (csrc/tc2_pipe.cu): device time of one forward at 4096 and 20,480 samples for a
list of (conv1, conv2, conv3) CTA splits, against the layer-by-layer forward.
    python tools/experiments/pipe_tune.py [NATURE|NIPS] > gpurun_out/pipe_tune.json
"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch

from paac_b200.policy_v_network import NaturePolicyVNetwork, NIPSPolicyVNetwork


def time_forward(net, pool, pi, v, ws, reps):
    for k in range(3):
        net.forward(pool[k % len(pool)], pi, v, ws)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for k in range(reps):
        net.forward(pool[k % len(pool)], pi, v, ws)
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def main():
    archs = [a for a in sys.argv[1:] if a in ('NATURE', 'NIPS')] or ['NATURE', 'NIPS']
    A = 6
    out = []
    for arch in archs:
        conf = dict(name='x', num_actions=A, clip_norm=3.0, clip_norm_type='global', device='/gpu:0',
                    entropy_regularisation_strength=0.02, seed=3, math='bf16x3')
        net = (NaturePolicyVNetwork if arch == 'NATURE' else NIPSPolicyVNetwork)(conf)
        if arch == 'NATURE':
            splits = [(36, 44, 38), (30, 46, 40), (40, 44, 36), (44, 42, 34), (34, 48, 40), (38, 46, 38), (42, 46, 36), (48, 42, 32),
                      (36, 40, 36), (32, 42, 36)]
        else:
            splits = [(53, 50, 0), (46, 54, 0), (60, 48, 0), (40, 56, 0), (66, 44, 0), (50, 58, 0), (56, 56, 0)]
        for N in (4096, 20480):
            gen = torch.Generator(device='cuda'); gen.manual_seed(1)
            pool = [torch.randint(0, 256, (N, 84, 84, 4), dtype=torch.uint8, device='cuda', generator=gen) for _ in range(3 if N > 8192 else 6)]
            pi = torch.empty((N, A), device='cuda'); v = torch.empty((N,), device='cuda')
            ws = torch.empty((net.workspace_floats(N),), device='cuda')
            reps = 20 if N <= 8192 else 8
            net.set_forward_pipeline(False)
            base = time_forward(net, pool, pi, v, ws, reps)
            row = {'arch': arch, 'samples': N, 'layer_by_layer_ms': round(base, 4), 'pipelined_ms': {}}
            for sp in splits:
                net.set_forward_pipeline(True, sp)
                row['pipelined_ms']['%d/%d/%d' % sp] = round(time_forward(net, pool, pi, v, ws, reps), 4)
            row['errors'] = net.forward_pipeline_errors()
            out.append(row)
            print(row, file=sys.stderr)
            del pool, ws
            torch.cuda.empty_cache()
    print(json.dumps({'what': 'device ms per forward (bf16x3): layer by layer vs layer-pipelined with the given CTA split conv1/conv2/conv3 '
                              '(fc takes the rest of the 148 SMs)', 'rows': out}, indent=1))


if __name__ == '__main__':
    main()
