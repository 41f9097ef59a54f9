"""Oracle (TEST INFRASTRUCTURE ONLY): generate tests/golden/*.npz by RUNNING THE REFERENCE ITSELF.

Run in the build container (needs /root/reference; never on the GPU box):

    python -m oracle.make_golden

What is executed from the reference, unmodified, by import:
  * environment.FramePool / ObservationPool        (environment.py:42-75)
  * runners.Runners + emulator_runner.EmulatorRunner (runners.py:7-50, emulator_runner.py:4-33)
  * the serialized TF-1.0.1 training graph pretrained/breakout/checkpoints/-80000000.meta,
    evaluated by oracle/tf_graph.py
and, standing in for scipy.misc.imresize(interp='nearest') (removed from scipy), Pillow's
Image.resize(NEAREST), which is what imresize called.

Inputs are regenerated from seeds by the tests (np.random.RandomState is a frozen legacy stream), so the
fixtures hold only the seeds and the reference outputs.
"""
import os
import sys
import json
import numpy as np

REF = '/root/reference'
OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tests', 'golden')


# ---- shared seeded input generators (imported by the tests too) ------------------------------------
def gen_frames(seed, n_envs, n_pairs):
    """uint8[n_envs, n_pairs, 2, 210, 160]: mix of i.i.d. noise and structured sprites."""
    rng = np.random.RandomState(seed)
    f = rng.randint(0, 256, size=(n_envs, n_pairs, 2, 210, 160)).astype(np.uint8)
    # structured half: mostly-dark screens with a few bright rectangles (so max-pool/resize index bugs show)
    for e in range(n_envs):
        for p in range(0, n_pairs, 2):
            g = np.zeros((2, 210, 160), np.uint8)
            for k in range(6):
                y, x = rng.randint(0, 200), rng.randint(0, 150)
                g[k % 2, y:y + rng.randint(1, 10), x:x + rng.randint(1, 10)] = rng.randint(1, 256)
            f[e, p] = g
    return f


class ScriptedEnv(object):
    """A deterministic BaseEnvironment used to drive the reference Runners/EmulatorRunner.
    State = f(env id, step count, last action); terminal every ``period`` steps."""

    def __init__(self, i, num_actions, period):
        self.i, self.A, self.period, self.t = i, num_actions, period, 0

    def _obs(self, a):
        base = (self.i * 37 + self.t * 11 + a * 5) % 256
        return ((np.arange(84 * 84 * 4, dtype=np.int64) * 7 + base) % 256).astype(np.uint8).reshape(84, 84, 4)

    def get_initial_state(self):
        self.t = 0
        return self._obs(0)

    def next(self, action):
        a = int(np.argmax(action))
        self.t += 1
        return self._obs(a), float((a + self.i) % 3 - 1) * 2.0, (self.t % self.period) == 0

    def get_legal_actions(self):
        return np.arange(self.A)

    def get_noop(self):
        return [1.0] + [0.0] * (self.A - 1)


def scripted_actions(seed, steps, n, A):
    rng = np.random.RandomState(seed)
    return rng.randint(0, A, size=(steps, n))


TF_SCOPE = {'conv1': 'local_learning_1', 'conv2': 'local_learning_1', 'fc3': 'local_learning_1',
            'actor': 'local_learning_2', 'critic': 'local_learning_2'}


def gen_train_batch(seed, b, A):
    rng = np.random.RandomState(seed)
    states = rng.randint(0, 256, (b, 84, 84, 4)).astype(np.uint8)
    acts = rng.randint(0, A, b).astype(np.int32)
    adv = rng.randn(b).astype(np.float32)
    tgt = rng.randn(b).astype(np.float32)
    return states, acts, adv, tgt


def main():
    sys.path.insert(0, REF)
    os.makedirs(OUT, exist_ok=True)
    import PIL
    from PIL import Image
    import environment as ref_env                      # the reference module, unmodified
    from oracle import preprocess as opre

    # ---------------- 1. resize tables + preprocessing through the reference classes --------------
    row_tab, col_tab = opre.pillow_nearest_tables()

    def ref_process(frame_pool):                        # atari_emulator.py:69-75 with imresize -> Pillow
        img = np.amax(frame_pool, axis=0)
        img = np.asarray(Image.fromarray(img).resize((84, 84), Image.NEAREST))
        return img.astype(np.uint8)

    seed, n_envs, steps = 1234, 3, 7
    reset_at = {(0, 3), (2, 5), (1, 1), (1, 2)}         # (env, step) pairs whose step ends an episode
    n_pairs = 4 + steps * 4                             # worst case: every step consumes 4 pairs
    frames = gen_frames(seed, n_envs, n_pairs)
    states = np.zeros((steps + 1, n_envs, 84, 84, 4), np.uint8)
    used = np.zeros((steps + 1, n_envs, 4), np.int32) - 1   # which pair indices fed each step (slot order)
    resets = np.zeros((steps + 1, n_envs), np.uint8)
    for e in range(n_envs):
        cursor = 0
        fp = ref_env.FramePool(np.empty((2, 210, 160), np.uint8), ref_process)
        op = ref_env.ObservationPool(np.zeros((84, 84, 4), np.uint8))

        def action_repeat():                            # atari_emulator.py:77-86 (last two frames pooled)
            nonlocal cursor
            fp.new_frame(frames[e, cursor, 0]); fp.new_frame(frames[e, cursor, 1])
            cursor += 1
            return cursor - 1

        def initial_state(t):                           # atari_emulator.py:88-96
            for k in range(4):
                used[t, e, k] = action_repeat()
                op.new_observation(fp.get_processed_frame())
            return op.get_pooled_observations()

        states[0, e] = initial_state(0); resets[0, e] = 1
        for t in range(1, steps + 1):
            # AtariEmulator.next (atari_emulator.py:98-106)
            p = action_repeat()
            op.new_observation(fp.get_processed_frame())
            s = op.get_pooled_observations()
            if (e, t) in reset_at:                      # emulator_runner.py:26-27
                s = initial_state(t); resets[t, e] = 1
            else:
                used[t, e, 0] = p
            states[t, e] = s
    np.savez_compressed(os.path.join(OUT, 'preprocess_golden.npz'), seed=seed, n_envs=n_envs, steps=steps,
                        n_pairs=n_pairs, used=used, resets=resets, states=states, row_tab=row_tab, col_tab=col_tab,
                        pillow=PIL.__version__)

    # ---------------- 2. Runners / EmulatorRunner protocol ----------------------------------------
    from runners import Runners as RefRunners
    from emulator_runner import EmulatorRunner as RefEmulatorRunner
    n, W, A, period, rsteps = 8, 4, 5, 3, 7
    emus = np.asarray([ScriptedEnv(i, A, period) for i in range(n)])
    variables = [np.asarray([e.get_initial_state() for e in emus], dtype=np.uint8),
                 np.zeros(n, dtype=np.float32), np.asarray([False] * n, dtype=np.float32),
                 np.zeros((n, A), dtype=np.float32)]
    rr = RefRunners(RefEmulatorRunner, emus, W, variables)
    rr.start()
    sh_states, sh_rew, sh_over, sh_act = rr.get_shared_variables()
    acts = scripted_actions(99, rsteps, n, A)
    rec_s, rec_r, rec_o = [np.array(sh_states, dtype=np.uint8)], [], []
    for t in range(rsteps):
        sh_act[:] = np.eye(A, dtype=np.float32)[acts[t]]
        rr.update_environments(); rr.wait_updated()
        rec_s.append(np.array(sh_states, dtype=np.uint8)); rec_r.append(np.array(sh_rew)); rec_o.append(np.array(sh_over))
    rr.stop()
    for r in rr.runners:
        r.join(5)
    np.savez_compressed(os.path.join(OUT, 'runner_golden.npz'), n=n, W=W, A=A, period=period, steps=rsteps,
                        act_seed=99, states=np.asarray(rec_s), rewards=np.asarray(rec_r), over=np.asarray(rec_o),
                        shared_state_dtype=str(sh_states.dtype))

    # ---------------- 3. the shipped TF training graph ---------------------------------------------
    from oracle import tf_graph, network
    meta = tf_graph.load_meta(os.path.join(REF, 'pretrained/breakout/checkpoints/-80000000.meta'))
    gi = tf_graph.GraphInterpreter(meta)
    A, b, wseed, bseed, lr = 4, 12, 3, 77, np.float32(0.0224)
    params = network.init_params('NIPS', A, wseed)
    variables = {}
    for nme, shp, _ in network.param_specs('NIPS', A):
        full = TF_SCOPE[nme.split('_')[0]] + '/' + nme
        assert gi.variable_shape(full) == tuple(shp)
        variables[full] = params[nme].copy()
        variables[full + '/OptimizerVariables'] = gi.slot_initial_value(full + '/OptimizerVariables')
        variables[full + '/OptimizerVariables_1'] = gi.slot_initial_value(full + '/OptimizerVariables_1')
    consts = {}
    for nm in ['local_learning/scalar', 'local_learning_2/Const', 'local_learning_2/Mul_4/x',
               'OptimizerVariables/decay', 'OptimizerVariables/momentum', 'OptimizerVariables/epsilon']:
        if nm in gi.nodes:
            consts[nm] = float(tf_graph._const(gi.nodes[nm]))
    out = dict(A=A, b=b, wseed=wseed, bseed=bseed, lr=lr, consts=json.dumps(consts),
               tf_version=meta.meta_info_def.tensorflow_version,
               slot_ms_init=float(variables['local_learning_1/conv1_weights/OptimizerVariables'].flat[0]),
               slot_mom_init=float(variables['local_learning_1/conv1_weights/OptimizerVariables_1'].flat[0]))
    for step in range(2):
        states, acts_b, adv, tgt = gen_train_batch(bseed + step, b, A)
        feeds = {'local_learning/input': states, 'local_learning/selected_action': np.eye(A, dtype=np.float32)[acts_b],
                 'local_learning_2/target': tgt, 'local_learning_2/advantage': adv, 'Placeholder': lr}
        grad_refs = [n.input[7] for n in gi.graph.node if n.op == 'ApplyRMSProp']     # clipped grads, TF var order
        var_refs = [n.input[0] for n in gi.graph.node if n.op == 'ApplyRMSProp']
        raw_refs = [n.input[0] for n in gi.graph.node if n.op == 'L2Loss']            # unclipped grads
        fetch = ['local_learning_2/actor_output_policy', 'local_learning_2/Reshape', 'local_learning_2/mul_1',
                 'global_norm/global_norm'] + grad_refs + raw_refs
        ex = gi.run_train_step(feeds, variables, fetch)
        out['pi%d' % step], out['v%d' % step] = ex[0], ex[1]
        out['loss%d' % step], out['norm%d' % step] = np.float32(ex[2]), np.float32(ex[3])
        k = len(grad_refs)
        out['clipped_sumsq%d' % step] = np.asarray([np.sum(np.square(g.astype(np.float64))) for g in ex[4:4 + k]])
        out['raw_sumsq%d' % step] = np.asarray([np.sum(np.square(g.astype(np.float64))) for g in ex[4 + k:]])
        out['raw_grad_head%d' % step] = np.concatenate([g.reshape(-1)[:64] for g in ex[4 + k:]])
        flat = np.concatenate([variables[v].reshape(-1) for v in var_refs])
        out['var_sample%d' % step] = flat[::997].copy()
        out['var_sum%d' % step] = np.asarray([np.sum(variables[v].astype(np.float64)) for v in var_refs])
        out['ms_sum%d' % step] = np.asarray([np.sum(variables[v + '/OptimizerVariables'].astype(np.float64)) for v in var_refs])
    out['var_order'] = json.dumps(var_refs)
    out['ops_used'] = json.dumps(sorted(gi.ops_used))
    np.savez_compressed(os.path.join(OUT, 'tf_graph_nips.npz'), **out)

    # ---------------- 4. graph pins for every shipped game (constants / shapes only) ---------------
    pins = {}
    for game in sorted(os.listdir(os.path.join(REF, 'pretrained'))):
        ck = os.path.join(REF, 'pretrained', game, 'checkpoints')
        metas = [f for f in os.listdir(ck) if f.endswith('.meta')] if os.path.isdir(ck) else []
        if not metas:
            continue
        g = tf_graph.GraphInterpreter(tf_graph.load_meta(os.path.join(ck, metas[0])))
        shapes = {v: list(g.variable_shape(v)) for v in g.variable_names() if 'Optimizer' not in v}
        with open(os.path.join(REF, 'pretrained', game, 'args.json')) as f:
            args = json.load(f)
        pins[game] = dict(shapes=shapes, args=args,
                          input_scale=float(tf_graph._const(g.nodes['local_learning/scalar'])))
    with open(os.path.join(OUT, 'tf_graph_pins.json'), 'w') as f:
        json.dump(pins, f, indent=1, sort_keys=True)

    # ---------------- 5. the shipped checkpoint-bundle index tables (TF 1.0.1 Saver output, ~1.3 KB each) -----------
    # data artefacts, copied verbatim: tests/test_tf_bundle.py parses them and re-serialises them byte for byte
    import shutil
    for game in sorted(os.listdir(os.path.join(REF, 'pretrained'))):
        ck = os.path.join(REF, 'pretrained', game, 'checkpoints')
        idx = [f for f in os.listdir(ck) if f.endswith('.index')] if os.path.isdir(ck) else []
        if idx:
            shutil.copyfile(os.path.join(ck, idx[0]), os.path.join(OUT, 'tf_index_%s.index' % game))
    print('golden written to', OUT, {f: os.path.getsize(os.path.join(OUT, f)) for f in sorted(os.listdir(OUT))})


if __name__ == '__main__':
    main()
