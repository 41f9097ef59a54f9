"""GPU: PAACLearner.train() end to end on the synthetic environment (host workers + shared, pinned+mapped buffers),
both runner protocols."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _args(tmp, raw, arch='NATURE'):
    from paac_b200 import train
    a = train.get_arg_parser().parse_args(['-g', 'synthetic', '-d', '/gpu:0', '--arch', arch, '-ec', '8', '-ew', '2',
                                           '--max_global_steps', str(8 * 5 * 3), '-df', str(tmp) + '/',
                                           '--raw_frames', 'True' if raw else 'False'])
    a.synthetic_p_terminal = 0.1
    return a


@pytest.mark.timeout(300)
def test_train_raw_and_classic_protocols_agree(tmp_path):
    from paac_b200 import train
    from paac_b200.paac import PAACLearner
    results = []
    for raw in (True, False):
        args = _args(tmp_path / ('raw' if raw else 'classic'), raw)
        net_creator, env_creator = train.get_network_and_environment_creator(args)
        learner = PAACLearner(net_creator, env_creator, args)
        learner.train()
        assert learner.global_step == 8 * 5 * 3
        loss = float(learner.last_loss.item())
        assert np.isfinite(loss) and np.isfinite(float(learner.last_norm.item()))
        results.append((learner.engine.get_states().cpu().numpy(), learner.network.get_params(), loss))
        assert learner.network.launch_count() > 0
    # same seeds, same sampled actions => the GPU-preprocessed states equal the host-preprocessed ones, bit for bit
    assert np.array_equal(results[0][0], results[1][0])
    # weights: the weight-gradient kernels accumulate with fp32 atomics (summation order varies run to run)
    assert np.max(np.abs(results[0][1] - results[1][1])) <= 1e-5 * np.max(np.abs(results[1][1]))


@pytest.mark.timeout(300)
def test_checkpoint_resume(tmp_path):
    from paac_b200 import train
    from paac_b200.paac import PAACLearner
    args = _args(tmp_path, True, 'NIPS')
    nc, ec = train.get_network_and_environment_creator(args)
    l1 = PAACLearner(nc, ec, args)
    l1.train()
    w = l1.network.get_params()
    args.max_global_steps = 8 * 5 * 4
    nc, ec = train.get_network_and_environment_creator(args)
    l2 = PAACLearner(nc, ec, args)
    step0 = l2.init_network()
    assert step0 == 8 * 5 * 3 and np.array_equal(l2.network.get_params(), w)
    assert abs(float(l2.engine.ms.mean().item()) - 1.0) > 0          # optimizer slots restored, no longer all ones


@pytest.mark.timeout(300)
def test_evaluation_loop_restores_tf_bundle_checkpoint(tmp_path):
    """test.py mirror: train a little, checkpoint (TensorFlow-1 bundle layout under <df>/checkpoints), then the evaluation
    loop restores the newest checkpoint from args.json + folder alone and plays episodes to the end."""
    import os
    from paac_b200 import logger_utils, tf_bundle, train
    from paac_b200 import test as evaluation
    from paac_b200.paac import PAACLearner
    args = _args(tmp_path, True)
    logger_utils.save_args(args, args.debugging_folder)
    nc, ec = train.get_network_and_environment_creator(args)
    learner = PAACLearner(nc, ec, args)
    learner.train()
    learner.cleanup()
    w = learner.network.get_params()
    ck = os.path.join(args.debugging_folder, 'checkpoints')
    assert os.path.exists(os.path.join(ck, 'checkpoint')) and os.path.exists(os.path.join(ck, '-120.index'))
    _, entries = tf_bundle.read_index(os.path.join(ck, '-120.index'))
    assert entries['local_learning_1/conv1_weights'].shape == (8, 8, 4, 32)
    assert entries['local_learning_2/actor_output_weights'].shape == (512, 6)
    eargs = evaluation.get_arg_parser().parse_args(['-f', args.debugging_folder, '-tc', '3', '-np', '2'])
    rewards = evaluation.evaluate(eargs)
    assert rewards.shape == (3,) and np.all(np.isfinite(rewards))


@pytest.mark.timeout(300)
@pytest.mark.parametrize('arch,single_life', [('NIPS', False), ('NATURE', True)])
def test_train_on_the_atari_emulator_both_protocols_agree(tmp_path, arch, single_life):
    """cfg1 / cfg2's code path with the reference's emulator class in the loop: EnvironmentCreator -> AtariEmulator (on the
    scripted ALE stand-in of tests/fake_ale: real ALE is not installable here) -> 2 forked worker processes -> shared, pinned +
    mapped buffers -> K1 / forward / update on the GPU.  The raw-frame protocol (workers write raw frame pairs, the GPU
    preprocesses) and the classic protocol (workers hand over 84x84x4 observations) must see the same states.  The learning
    rate is 0: this game's frames depend on the actions, and two runs whose weights differ by the rounding of their fp32
    atomics would eventually sample a different action; with frozen weights both runs play the same 80 steps per emulator
    (through game-over and lost-life resets) and every state must match bit for bit."""
    import os
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), 'fake_ale'))
    from paac_b200 import train
    from paac_b200.paac import PAACLearner
    results = []
    for raw in (True, False):
        a = train.get_arg_parser().parse_args(['-g', 'fake_breakout', '--rom_path', str(tmp_path), '-d', '/gpu:0', '--arch', arch, '-ec', '8',
                                               '-ew', '2', '--max_global_steps', str(8 * 5 * 16), '-df', str(tmp_path / ('r%d' % raw)) + '/',
                                               '--raw_frames', 'True' if raw else 'False', '--single_life_episodes', str(single_life), '-lr', '0.0',
                                               '--random_start', 'False'])
        # (no random no-op starts: Python re-seeds the global `random` stream of a forked worker from OS entropy, so the no-op
        # counts of the resets inside the workers differ from run to run -- in the reference as well)
        nc, ec = train.get_network_and_environment_creator(a)
        assert ec.num_actions == 4
        learner = PAACLearner(nc, ec, a)
        learner.train()
        assert learner.global_step == 8 * 5 * 16
        results.append((learner.engine.get_states().cpu().numpy(), learner.network.get_params(), float(learner.last_loss.item())))
    assert np.array_equal(results[0][0], results[1][0])
    assert results[0][0].max() > 0 and np.isfinite(results[0][2])
    assert np.array_equal(results[0][1], results[1][1]) and abs(results[0][2] - results[1][2]) <= 1e-5 * max(1.0, abs(results[1][2]))
