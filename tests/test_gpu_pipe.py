"""GPU: the layer-pipelined forward (csrc/tc2_pipe.cu: conv1 ... hidden fc as one persistent kernel with the SMs partitioned
between the layers) against the layer-by-layer forward of the same library: every activation plane of the workspace, pi, v
and the sampled actions must be the SAME BITS (each sample is computed by the same instruction sequence in both), for ragged
and tile-aligned batches, for slices of a larger workspace, for unbalanced role sizes, and after the parameters change."""
import numpy as np
import pytest
import torch

import gpu_util as G

pytestmark = pytest.mark.gpu


def run(net, states, cap=None, first=0):
    b = states.shape[0]
    cap = b if cap is None else cap
    A = net.num_actions
    pi = torch.full((b, A), -1.0, device='cuda'); v = torch.full((b,), -1.0, device='cuda')
    ws = torch.zeros((net.workspace_floats(cap),), device='cuda')
    net.forward(states, pi, v, ws, ws_capacity=cap, ws_first=first)
    torch.cuda.synchronize()
    return ws.view(torch.int32).clone(), pi, v


@pytest.mark.parametrize('arch,A,b', [('NATURE', 6, 1), ('NATURE', 6, 37), ('NATURE', 18, 128), ('NATURE', 6, 700), ('NATURE', 6, 4096),
                                      ('NIPS', 4, 3), ('NIPS', 6, 129), ('NIPS', 6, 1111), ('NIPS', 6, 4096)])
def test_pipelined_forward_is_bit_identical(arch, A, b):
    net = G.make_net(arch, A, seed=11, math='bf16x3')
    gen = torch.Generator(device='cuda'); gen.manual_seed(b)
    states = torch.randint(0, 256, (b, 84, 84, 4), dtype=torch.uint8, device='cuda', generator=gen)
    net.set_forward_pipeline(False)
    ws0, pi0, v0 = run(net, states)
    launches0 = net.launch_count()
    net.set_forward_pipeline(True)
    ws1, pi1, v1 = run(net, states)
    assert net.launch_count() - launches0 == 2, 'the pipelined forward is the pipe kernel + the heads kernel'
    assert net.forward_pipeline_errors() == 0
    assert torch.equal(ws0, ws1) and torch.equal(pi0, pi1) and torch.equal(v0, v1)
    assert torch.isfinite(pi1).all() and (pi1 >= 0).all()


@pytest.mark.parametrize('ctas', [(1, 1, 1), (100, 20, 20), (10, 100, 10), (5, 5, 130)])
def test_any_role_split_gives_the_same_bits(ctas):
    """Starved producers, starved consumers: the hand-off is a correctness protocol, not a schedule."""
    net = G.make_net('NATURE', 6, seed=12, math='bf16x3')
    gen = torch.Generator(device='cuda'); gen.manual_seed(5)
    states = torch.randint(0, 256, (1000, 84, 84, 4), dtype=torch.uint8, device='cuda', generator=gen)
    net.set_forward_pipeline(False)
    ref = run(net, states)
    net.set_forward_pipeline(True, ctas)
    got = run(net, states)
    assert net.forward_pipeline_errors() == 0
    for a, b in zip(ref, got):
        assert torch.equal(a, b)


def test_slices_of_a_workspace_and_changed_parameters():
    net = G.make_net('NATURE', 6, seed=13, math='bf16x3')
    gen = torch.Generator(device='cuda'); gen.manual_seed(6)
    states = torch.randint(0, 256, (600, 84, 84, 4), dtype=torch.uint8, device='cuda', generator=gen)
    A = net.num_actions

    def sliced(enable):
        net.set_forward_pipeline(enable)
        pi = torch.zeros((600, A), device='cuda'); v = torch.zeros((600,), device='cuda')
        ws = torch.zeros((net.workspace_floats(600),), device='cuda')
        for lo, hi in ((0, 250), (250, 251), (251, 600)):          # paacb_policy_forward_at: ragged slices of one workspace
            net.forward(states[lo:hi], pi[lo:hi], v[lo:hi], ws, ws_capacity=600, ws_first=lo)
        torch.cuda.synchronize()
        return ws.view(torch.int32).clone(), pi, v
    for a, b in zip(sliced(False), sliced(True)):
        assert torch.equal(a, b)
    whole = run(net, states)
    for a, b in zip(whole, sliced(True)):
        assert torch.equal(a, b)
    # new parameters: the cached operand images are refreshed for both schedules
    net.set_params(net.get_params() * np.float32(0.5))
    r0 = sliced(False); r1 = sliced(True)
    for a, b in zip(r0, r1):
        assert torch.equal(a, b)
    assert not torch.equal(r1[2], whole[2])
    assert net.forward_pipeline_errors() == 0


def test_two_streams_get_their_own_counters():
    net = G.make_net('NIPS', 6, seed=14, math='bf16x3')
    gen = torch.Generator(device='cuda'); gen.manual_seed(7)
    sa = torch.randint(0, 256, (2000, 84, 84, 4), dtype=torch.uint8, device='cuda', generator=gen)
    sb = torch.randint(0, 256, (1500, 84, 84, 4), dtype=torch.uint8, device='cuda', generator=gen)
    net.set_forward_pipeline(False)
    ra, rb = run(net, sa), run(net, sb)
    net.set_forward_pipeline(True)
    A = net.num_actions
    outs = []
    streams = [torch.cuda.Stream(), torch.cuda.Stream()]
    for st, s in zip(streams, (sa, sb)):
        b = s.shape[0]
        pi = torch.zeros((b, A), device='cuda'); v = torch.zeros((b,), device='cuda')
        ws = torch.zeros((net.workspace_floats(b),), device='cuda')
        outs.append((ws, pi, v))
    torch.cuda.synchronize()
    for rep in range(3):
        for st, s, (ws, pi, v) in zip(streams, (sa, sb), outs):
            with torch.cuda.stream(st):
                net.forward(s, pi, v, ws)
    torch.cuda.synchronize()
    assert net.forward_pipeline_errors() == 0
    for ref, (ws, pi, v) in zip((ra, rb), outs):
        assert torch.equal(ref[0], ws.view(torch.int32)) and torch.equal(ref[1], pi) and torch.equal(ref[2], v)
