"""Oracle (TEST INFRASTRUCTURE ONLY): NIPS / Nature policy-value network, loss and gradients.

torch-CPU restatement of the TF1 graph the reference builds:
  * ``Network.__init__``      networks.py:102-120   (uint8 NHWC input, cast f32, * (1/255))
  * ``conv2d`` / ``fc`` / ``softmax`` helpers and "torch" init  networks.py:12-89
  * ``NIPSNetwork``           networks.py:138-151
  * ``NatureNetwork``         networks.py:154-169
  * ``PolicyVNetwork``        policy_v_network.py:6-57  (heads, entropy, A2C loss, x5 scaling)

Layouts at the interface are the reference's: states NHWC uint8, conv weights HWIO
``[kh, kw, cin, cout]`` (networks.py:13), fc weights ``[in, out]`` (networks.py:51), flatten
order (h, w, c) (networks.py:6-9).  ``dtype`` selects fp32 (like-for-like) or fp64 (arbiter).
"""
import numpy as np
import torch
import torch.nn.functional as F

INPUT_SCALE = np.float32(1.0 / 255.0)         # networks.py:115; f32 constant 0.003921568859368563 in the .meta
LOG_EPS = np.float32(1e-30)                   # policy_v_network.py:29
LOSS_SCALING = 5.0                            # networks.py:112
CRITIC_SCALE = 0.25                           # policy_v_network.py:53

# (name, kernel, cin, cout, stride) -- networks.py:145-149 / 161-167
ARCH = {
    'NIPS': dict(convs=[('conv1', 8, 4, 16, 4), ('conv2', 4, 16, 32, 2)], fc=('fc3', 2592, 256)),
    'NATURE': dict(convs=[('conv1', 8, 4, 32, 4), ('conv2', 4, 32, 64, 2), ('conv3', 3, 64, 64, 1)],
                   fc=('fc4', 3136, 512)),
}


def param_specs(arch, num_actions):
    """[(tf_variable_name, shape, fan_in)] in TF variable-creation order."""
    a = ARCH[arch.upper()]
    specs = []
    for name, k, cin, cout, _ in a['convs']:
        fan = k * k * cin                                     # networks.py:31-34,44
        specs.append((name + '_weights', (k, k, cin, cout), fan))
        specs.append((name + '_biases', (cout,), fan))
    fname, fin, fout = a['fc']
    specs.append((fname + '_weights', (fin, fout), fin))      # networks.py:69-70
    specs.append((fname + '_biases', (fout,), fin))           # networks.py:79
    specs.append(('actor_output_weights', (fout, num_actions), fout))
    specs.append(('actor_output_biases', (num_actions,), fout))
    specs.append(('critic_output_weights', (fout, 1), fout))
    specs.append(('critic_output_biases', (1,), fout))
    return specs


def param_count(arch, num_actions):
    return int(sum(int(np.prod(s)) for _, s, _ in param_specs(arch, num_actions)))


def init_params(arch, num_actions, seed):
    """U(-d, d), d = 1/sqrt(fan_in) for weights AND biases (networks.py:24-46, 63-81).  fp32 numpy dict."""
    rng = np.random.RandomState(seed)
    out = {}
    for name, shape, fan in param_specs(arch, num_actions):
        d = 1.0 / np.sqrt(fan)
        out[name] = rng.uniform(-d, d, size=shape).astype(np.float32)
    return out


def flatten_params(params, arch, num_actions):
    return np.concatenate([np.asarray(params[n], np.float32).reshape(-1) for n, _, _ in param_specs(arch, num_actions)])


def unflatten_params(flat, arch, num_actions):
    out, o = {}, 0
    for n, s, _ in param_specs(arch, num_actions):
        k = int(np.prod(s))
        out[n] = np.asarray(flat[o:o + k]).reshape(s).copy()
        o += k
    assert o == len(flat)
    return out


def _t(x, dtype):
    return torch.as_tensor(np.asarray(x), dtype=dtype)


def forward(params, states_u8, arch, dtype=torch.float32, keep=False):
    """states_u8 uint8[b,84,84,4] -> dict(pi[b,A], v[b], logits, h, activations).

    params: dict name -> numpy or torch tensors (torch tensors keep autograd).
    """
    a = ARCH[arch.upper()]
    P = {k: (v if torch.is_tensor(v) else _t(v, dtype)) for k, v in params.items()}
    x = torch.as_tensor(np.asarray(states_u8)).to(dtype)
    # networks.py:115: scalar_mul(1/255, cast(input, f32)); the scale is the f32 constant.
    x = x * torch.tensor(float(INPUT_SCALE), dtype=dtype)
    x = x.permute(0, 3, 1, 2)                                       # NHWC -> NCHW for torch
    acts = []
    for name, k, cin, cout, stride in a['convs']:
        w = P[name + '_weights'].permute(3, 2, 0, 1)                # HWIO -> OIHW
        x = F.relu(F.conv2d(x, w, P[name + '_biases'], stride=stride))   # VALID; networks.py:17-20
        acts.append(x.permute(0, 2, 3, 1))
    flat = x.permute(0, 2, 3, 1).reshape(x.shape[0], -1)            # (h, w, c) order; networks.py:6-9
    fname = a['fc'][0]
    h = F.relu(flat @ P[fname + '_weights'] + P[fname + '_biases'])     # networks.py:55-58
    logits = h @ P['actor_output_weights'] + P['actor_output_biases']   # networks.py:88
    pi = torch.softmax(logits, dim=1)
    v = (h @ P['critic_output_weights'] + P['critic_output_biases']).reshape(-1)   # policy_v_network.py:26,37
    out = dict(pi=pi, v=v, logits=logits, h=h)
    if keep:
        out['acts'] = acts
    return out


def a2c_loss(pi, v, onehot, adv, target, beta):
    """policy_v_network.py:29-57.  All torch tensors of one dtype."""
    dtype = pi.dtype
    logpi = torch.log(pi + torch.tensor(float(LOG_EPS), dtype=dtype))           # :29-30
    entropy = torch.sum(-1.0 * (pi * logpi), dim=1)                              # :33-35
    critic = target - v                                                          # :40
    log_sel = torch.sum(logpi * onehot, dim=1)                                   # :42-44
    actor_mean = torch.mean(-1.0 * (log_sel * adv + beta * entropy))             # :46-51
    critic_mean = torch.mean(CRITIC_SCALE * critic.pow(2))                       # :53
    return LOSS_SCALING * (actor_mean + critic_mean)                             # :57


def loss_and_grads(params, states_u8, actions, adv, target, beta, arch, num_actions, dtype=torch.float32):
    """Autograd of the reference loss w.r.t. every variable (actor_learner.py:44).

    actions int[b]; adv, target float[b] are fed constants (placeholders, no grad).
    Returns (loss float, grads dict name->numpy, fwd dict of numpy).
    """
    P = {n: _t(params[n], dtype).clone().requires_grad_(True) for n, _, _ in param_specs(arch, num_actions)}
    fwd = forward(P, states_u8, arch, dtype)
    onehot = F.one_hot(torch.as_tensor(np.asarray(actions), dtype=torch.long), num_actions).to(dtype)
    loss = a2c_loss(fwd['pi'], fwd['v'], onehot, _t(adv, dtype), _t(target, dtype), float(beta))
    names = [n for n, _, _ in param_specs(arch, num_actions)]
    gs = torch.autograd.grad(loss, [P[n] for n in names])
    grads = {n: g.detach().numpy() for n, g in zip(names, gs)}
    fwd_np = {k: t.detach().numpy() for k, t in fwd.items() if torch.is_tensor(t)}
    return float(loss.detach()), grads, fwd_np


def closed_form_head_grads(logits, v, actions, adv, target, beta):
    """SURVEY App. C closed forms (float64 numpy): returns (loss, dlogits[b,A], dv[b])."""
    z = np.asarray(logits, np.float64)
    b, A = z.shape
    z = z - z.max(axis=1, keepdims=True)
    pi = np.exp(z)
    pi /= pi.sum(axis=1, keepdims=True)
    eps = float(LOG_EPS)
    logpi = np.log(pi + eps)
    H = -(pi * logpi).sum(axis=1)
    onehot = np.eye(A)[np.asarray(actions)]
    adv = np.asarray(adv, np.float64)
    target = np.asarray(target, np.float64)
    v = np.asarray(v, np.float64)
    loss = (LOSS_SCALING / b) * np.sum(-((logpi * onehot).sum(1) * adv + beta * H) + CRITIC_SCALE * (target - v) ** 2)
    # exact-epsilon form: dL/dpi_j = -adv*onehot_j/(pi_j+eps) + beta*(logpi_j + pi_j/(pi_j+eps)); then softmax Jacobian
    dpi = -adv[:, None] * onehot / (pi + eps) + beta * (logpi + pi / (pi + eps))
    dz = pi * (dpi - (dpi * pi).sum(axis=1, keepdims=True))
    dlogits = (LOSS_SCALING / b) * dz
    dv = (LOSS_SCALING * CRITIC_SCALE * 2.0 / b) * (v - target)
    return loss, dlogits, dv


def masked_loss_and_grads(params, states_u8, actions, adv, target, beta, arch, num_actions, masks, dtype=torch.float64):
    """Like loss_and_grads, but every ReLU is replaced by multiplication with a GIVEN 0/1 mask (``masks``: one array
    per conv layer + the hidden fc, NHWC / [b, F]).  Used to check a backward implementation like-for-like: with the
    masks taken from the implementation's own forward activations, a pre-activation that rounds to the other side of
    zero in fp32 cannot flip a ReLU between the two sides of the comparison (TF's ReluGrad uses its own activations too).
    Returns (grads dict, dz list [conv layers..., hidden fc] as numpy, NHWC)."""
    names = [n for n, _, _ in param_specs(arch, num_actions)]
    P = {n: _t(params[n], dtype).clone().requires_grad_(True) for n in names}
    a = ARCH[arch.upper()]
    x = torch.as_tensor(np.asarray(states_u8)).to(dtype) * torch.tensor(float(INPUT_SCALE), dtype=dtype)
    x = x.permute(0, 3, 1, 2)
    zs = []
    for li, (name, k, cin, cout, stride) in enumerate(a['convs']):
        w = P[name + '_weights'].permute(3, 2, 0, 1)
        z = F.conv2d(x, w, P[name + '_biases'], stride=stride)
        z.retain_grad()
        zs.append(z)
        x = z * torch.as_tensor(np.asarray(masks[li])).to(dtype).permute(0, 3, 1, 2)
    flat = x.permute(0, 2, 3, 1).reshape(x.shape[0], -1)
    fname = a['fc'][0]
    zh = flat @ P[fname + '_weights'] + P[fname + '_biases']
    zh.retain_grad()
    h = zh * torch.as_tensor(np.asarray(masks[len(a['convs'])])).to(dtype)
    logits = h @ P['actor_output_weights'] + P['actor_output_biases']
    pi = torch.softmax(logits, dim=1)
    v = (h @ P['critic_output_weights'] + P['critic_output_biases']).reshape(-1)
    onehot = F.one_hot(torch.as_tensor(np.asarray(actions), dtype=torch.long), num_actions).to(dtype)
    loss = a2c_loss(pi, v, onehot, _t(adv, dtype), _t(target, dtype), float(beta))
    loss.backward()
    grads = {n: P[n].grad.numpy() for n in names}
    dz = [z.grad.permute(0, 2, 3, 1).contiguous().numpy() for z in zs] + [zh.grad.numpy()]
    return grads, dz, dict(logits=logits.detach().numpy(), v=v.detach().numpy())
