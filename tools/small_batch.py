"""The reference's DEFAULT size on a B200 (BASELINE.json configs[0], [1], [3]: 32 environments per GPU, train.py:95-96): a PAAC
cycle is ~60 small launches and launch latency is the whole cost, so the engine replays act(t) / update() as CUDA graphs.

    python tools/small_batch.py [--out profiles/r02_small_batch.json] [--envs 32] [--cycles 200] [--no_learner]

Per architecture (NIPS = cfg1's network, Nature = cfg2 / cfg4's): ms per update and per whole cycle (device-resident raw
frames, event-timed after warm-up), eager and graph-replayed, and env-steps/s; then the PRODUCT loop --
PAACLearner.train() with its Runners / worker processes on the synthetic game (cfg1's shape: 32 emulators, 8 workers) --
as wall-clock steps/s.  bench.py copies the summary into its `variants`.
"""
import argparse
import json
import os
import sys
import tempfile
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch


def engine_numbers(arch, envs, cycles, graphs, math='auto', T=5, A=6, train_forward='batched'):
    from paac_b200.engine import RolloutEngine
    from paac_b200.policy_v_network import NaturePolicyVNetwork, NIPSPolicyVNetwork
    dev = torch.device('cuda', torch.cuda.current_device())
    conf = dict(name='local_learning', num_actions=A, clip_norm=3.0, clip_norm_type='global', device='/gpu:%d' % dev.index,
                entropy_regularisation_strength=0.02, seed=3, math=math)
    net = (NaturePolicyVNetwork if arch == 'NATURE' else NIPSPolicyVNetwork)(conf)
    eng = RolloutEngine(net, envs, T, seed=3, train_forward=train_forward)
    gen = torch.Generator(device=dev); gen.manual_seed(3)
    pool = [torch.randint(0, 256, (envs, 1, 2, 210, 160), dtype=torch.uint8, device=dev, generator=gen) for _ in range(8)]
    u = torch.rand((T, envs), device=dev, generator=gen)
    rewards = torch.where(u < 0.05, -1.0, torch.where(u > 0.95, 1.0, 0.0)).float()
    over = (torch.rand((T, envs), device=dev, generator=gen) < 0.01).float()
    eng.state(0).copy_(torch.randint(0, 256, eng.state(0).shape, dtype=torch.uint8, device=dev, generator=gen))
    if graphs:
        eng.enable_graphs()
    k = [0]
    upd = [torch.cuda.Event(enable_timing=True) for _ in range(2 * cycles)]

    def cycle(i=None):
        for t in range(T):
            eng.act(t)
            eng.observe_frames(t, pool[k[0] % 8].data_ptr(), 1, None, rewards[t], over[t])
            k[0] += 1
        if i is not None:
            upd[2 * i].record()
        eng.update(0.0224)
        if i is not None:
            upd[2 * i + 1].record()

    for _ in range(10):
        cycle()
    torch.cuda.synchronize()
    launches0 = net.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    w0 = time.perf_counter()
    e0.record()
    for i in range(cycles):
        cycle(i)
    e1.record()
    torch.cuda.synchronize()
    wall = time.perf_counter() - w0
    ms = e0.elapsed_time(e1) / cycles
    upd_ms = sum(upd[2 * i].elapsed_time(upd[2 * i + 1]) for i in range(cycles)) / cycles
    return {'arch': arch, 'envs': envs, 'math': net.math, 'graphs': bool(graphs), 'train_forward': train_forward,
            'ms_per_cycle': ms, 'update_ms': upd_ms, 'env_steps_per_s': envs * T / (ms / 1e3), 'wall_ms_per_cycle': 1e3 * wall / cycles,
            'kernel_launches_issued_per_cycle': (net.launch_count() - launches0) / cycles, 'loss': float(eng.loss.item())}


def learner_numbers(arch, envs, workers, updates, graphs='auto', train_forward='reuse'):
    from paac_b200 import train
    from paac_b200.paac import PAACLearner
    folder = tempfile.mkdtemp(prefix='paacb_small_')
    argv = ['-g', 'synthetic', '-d', '/gpu:%d' % torch.cuda.current_device(), '--arch', arch, '-ec', str(envs), '-ew', str(workers),
            '--max_global_steps', str(envs * 5 * updates), '-df', folder + '/', '--graphs', graphs, '--train_forward', train_forward]
    args = train.get_arg_parser().parse_args(argv)
    nc, ec = train.get_network_and_environment_creator(args)
    learner = PAACLearner(nc, ec, args)
    learner.train()
    return {'arch': arch, 'envs': envs, 'workers': workers, 'updates': updates, 'graphs': graphs, 'train_forward': train_forward,
            'steps_per_s_wall': learner.steps_per_second,
            'what': 'PAACLearner.train() end to end: %d worker processes step the synthetic emulator (raw 210x160 frames into '
                    'pinned, mapped shared memory), the GPU preprocesses, acts and updates; wall clock incl. worker IPC' % workers}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--out', default='')
    ap.add_argument('--envs', type=int, default=32)
    ap.add_argument('--cycles', type=int, default=200)
    ap.add_argument('--no_learner', action='store_true')
    a = ap.parse_args()
    out = {'what': 'PAAC at the reference default size (%d environments per GPU, t_max 5, 6 actions) on one B200' % a.envs,
           'engine': [], 'learner': []}
    for arch in ('NIPS', 'NATURE'):
        for graphs in (False, True):
            out['engine'].append(engine_numbers(arch, a.envs, a.cycles, graphs))
        out['engine'].append(engine_numbers(arch, a.envs, a.cycles, True, train_forward='reuse'))
    if not a.no_learner:
        for arch in ('NIPS', 'NATURE'):
            out['learner'].append(learner_numbers(arch, a.envs, 8, 300))
        out['learner'].append(learner_numbers('NATURE', a.envs, 8, 300, graphs='false', train_forward='batched'))
    text = json.dumps(out, indent=1)
    if a.out:
        with open(a.out, 'w') as f:
            f.write(text + '\n')
    print(text)


if __name__ == '__main__':
    main()
