"""Helpers for the -m gpu parity tests: thin wrappers that call the C ABI (through ctypes) on CUDA tensors."""
import ctypes as C

import numpy as np
import torch

from paac_b200 import _lib
from paac_b200.policy_v_network import NaturePolicyVNetwork, NIPSPolicyVNetwork


def make_net(arch, A, seed=3, math='fp32', beta=0.02):
    conf = dict(name='local_learning', num_actions=A, clip_norm=3.0, clip_norm_type='global', device='/gpu:0',
                entropy_regularisation_strength=beta, seed=seed, math=math)
    return (NIPSPolicyVNetwork if arch.upper() == 'NIPS' else NaturePolicyVNetwork)(conf)


def dev(x, dtype=None):
    t = torch.as_tensor(np.ascontiguousarray(x))
    if dtype is not None:
        t = t.to(dtype)
    return t.cuda()


def stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def preprocess(net, frames, pairs, reset, prev, out=None):
    f = dev(frames)
    r = dev(reset) if reset is not None else None
    p = dev(prev)
    o = torch.empty_like(p) if out is None else out
    n = p.shape[0]
    _lib.check(net._lib.paacb_preprocess_u8(net.ctx, _lib.ptr(f), pairs, _lib.ptr(r), _lib.ptr(p), _lib.ptr(o), n, stream()),
               'paacb_preprocess_u8')
    torch.cuda.synchronize()
    return o


def forward(net, states, uniforms=None):
    st = dev(states)
    b = st.shape[0]
    A = net.num_actions
    pi = torch.empty((b, A), dtype=torch.float32, device='cuda')
    v = torch.empty((b,), dtype=torch.float32, device='cuda')
    ws = torch.zeros((max(net.workspace_floats(b), 4),), dtype=torch.float32, device='cuda')
    out = dict(pi=pi, v=v, ws=ws, states=st)
    if uniforms is not None:
        out['u'] = dev(uniforms, torch.float32)
        out['actions'] = torch.full((b,), -1, dtype=torch.int32, device='cuda')
        out['onehot'] = torch.full((b, A), -1.0, dtype=torch.float32, device='cuda')
        net.forward(st, pi, v, ws, uniforms=out['u'], actions=out['actions'], onehot=out['onehot'])
    else:
        net.forward(st, pi, v, ws)
    torch.cuda.synchronize()
    return out


def layer_acts(net, ws, b):
    """Split the forward workspace into per-layer activations [b, ...] (conv NHWC, then hidden fc)."""
    return [t.cpu().numpy() for t in net.layer_tensors(ws, b)]


def returns_loss_grad(net, rewards, over, values, boot, actions, pi, v, gamma, beta):
    T, N = rewards.shape
    A = net.num_actions
    B = T * N
    f = lambda a: dev(np.asarray(a, np.float32))
    d = dict(rewards=f(rewards), over=f(over), values=f(values), boot=f(boot), actions=dev(np.asarray(actions, np.int32)),
             pi=f(pi), v=f(v))
    o = dict(y=torch.empty(B, device='cuda'), adv=torch.empty(B, device='cuda'), dlogits=torch.empty((B, A), device='cuda'),
             dv=torch.empty(B, device='cuda'), loss=torch.full((1,), 123.0, device='cuda'))
    p = _lib.ptr
    _lib.check(net._lib.paacb_returns_loss_grad(net.ctx, p(d['rewards']), p(d['over']), p(d['values']), p(d['boot']),
                                                p(d['actions']), p(d['pi']), p(d['v']), T, N, C.c_double(gamma),
                                                C.c_float(beta), p(o['y']), p(o['adv']), p(o['dlogits']), p(o['dv']),
                                                p(o['loss']), stream()), 'paacb_returns_loss_grad')
    torch.cuda.synchronize()
    return {k: t.cpu().numpy() for k, t in o.items()}


def backward(net, fwd, dlogits, dv):
    b = fwd['states'].shape[0]
    bws = torch.zeros((max(int(net._lib.paacb_backward_workspace_floats(net.ctx, b)), 4),), dtype=torch.float32, device='cuda')
    grads = torch.full((net.param_count,), 7.0, dtype=torch.float32, device='cuda')     # must be overwritten
    dl, dvv = dev(np.asarray(dlogits, np.float32)), dev(np.asarray(dv, np.float32))
    p = _lib.ptr
    _lib.check(net._lib.paacb_backward(net.ctx, p(net.params), p(fwd['states']), b, p(fwd['ws']), p(dl), p(dvv), p(bws),
                                       p(grads), stream()), 'paacb_backward')
    torch.cuda.synchronize()
    return grads.cpu().numpy(), bws


def clip_rmsprop(net, params, ms, mom, grads, gscale, lr, rho, eps, momentum, clip, clip_type):
    t = [dev(np.asarray(a, np.float32)) for a in (params, ms, mom, grads)]
    norm = torch.zeros(1, device='cuda')
    ws = torch.zeros((int(net._lib.paacb_optimizer_workspace_floats(net.ctx)),), dtype=torch.float32, device='cuda')
    p = _lib.ptr
    # the context's P is fixed; callers pass vectors of exactly net.param_count
    _lib.check(net._lib.paacb_clip_rmsprop(net.ctx, p(t[0]), p(t[1]), p(t[2]), p(t[3]), C.c_float(gscale), C.c_float(lr),
                                           C.c_float(rho), C.c_float(eps), C.c_float(momentum), C.c_float(clip),
                                           clip_type, p(norm), p(ws), stream()), 'paacb_clip_rmsprop')
    torch.cuda.synchronize()
    return t[0].cpu().numpy(), t[1].cpu().numpy(), t[2].cpu().numpy(), float(norm.item())


def fp64_masked_grads_cuda(arch, A, params_flat, flat_states, acts, dlogits, dv, chunk=1024):
    """The fp64 arbiter at sizes the CPU oracle cannot reach: oracle.network.masked_loss_and_grads restated with torch ops ON
    THE GPU (float64 conv2d / matmul), chunk by chunk.  The backward is linear in (dlogits, dv), so the gradient of the loss
    is the vector-Jacobian product of (logits, v) with the head gradients the implementation itself was given; every ReLU
    is a multiplication with the implementation's own 0/1 mask (acts[i] > 0), exactly as in the CPU oracle.
    acts: per-layer activations [B, ...] NHWC + hidden [B, F] (CUDA tensors); returns dict name -> float64 numpy."""
    from oracle import network
    import torch.nn.functional as F
    names = [n for n, _, _ in network.param_specs(arch, A)]
    P = {n: torch.as_tensor(v).cuda().double().requires_grad_(True)
         for n, v in network.unflatten_params(np.asarray(params_flat), arch, A).items()}
    a = network.ARCH[arch.upper()]
    total = {n: torch.zeros_like(P[n]) for n in names}
    B = flat_states.shape[0]
    scale = torch.tensor(float(network.INPUT_SCALE), dtype=torch.float64, device='cuda')
    for lo in range(0, B, chunk):
        hi = min(B, lo + chunk)
        x = (flat_states[lo:hi].double() * scale).permute(0, 3, 1, 2)
        for li, (name, k, cin, cout, stride) in enumerate(a['convs']):
            z = F.conv2d(x, P[name + '_weights'].permute(3, 2, 0, 1), P[name + '_biases'], stride=stride)
            x = z * (acts[li][lo:hi] > 0).double().permute(0, 3, 1, 2)
        flat = x.permute(0, 2, 3, 1).reshape(hi - lo, -1)
        fname = a['fc'][0]
        h = (flat @ P[fname + '_weights'] + P[fname + '_biases']) * (acts[len(a['convs'])][lo:hi] > 0).double()
        logits = h @ P['actor_output_weights'] + P['actor_output_biases']
        v = (h @ P['critic_output_weights'] + P['critic_output_biases']).reshape(-1)
        gs = torch.autograd.grad([logits, v], [P[n] for n in names], [dlogits[lo:hi].double(), dv[lo:hi].double()])
        for n, g in zip(names, gs):
            total[n] += g
    return {n: total[n].cpu().numpy() for n in names}
