"""CPU, world_size 2 over gloo: the multi-GPU scheme of paac_b200/parallel.py -- shard the environments, allreduce-SUM the
flat gradient, scale by 1/world before the global-norm clip -- reproduces the single-learner update on the concatenated
batch, and leaves every rank with identical parameters.  Gradients come from the CPU oracle (this is a test of the
host-side parallel logic; the GPU kernels are covered by the -m gpu tests)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.multiprocessing as mp

from oracle import network, update
from paac_b200 import parallel

ARCH, A, N_TOTAL, T = 'NIPS', 4, 8, 2


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _batch(seed=5):
    rng = np.random.RandomState(seed)
    B = N_TOTAL * T
    states = rng.randint(0, 256, (T, N_TOTAL, 84, 84, 4)).astype(np.uint8)
    acts = rng.randint(0, A, (T, N_TOTAL))
    adv = rng.randn(T, N_TOTAL).astype(np.float32)
    tgt = rng.randn(T, N_TOTAL).astype(np.float32)
    return states, acts, adv, tgt


def _apply(params, flat_grad, specs):
    grads = network.unflatten_params(flat_grad, ARCH, A)
    clipped, norm = update.clip_by_global_norm([grads[n] for n, _, _ in specs], 3.0)
    out = {}
    for (n, s, _), g in zip(specs, clipped):
        out[n], _, _ = update.rmsprop_apply(params[n], np.ones(s, np.float32), np.zeros(s, np.float32), g, 0.0224, 0.99, 0.1)
    return network.flatten_params(out, ARCH, A), float(norm)


def _worker(rank, world, port, ret):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    torch.set_num_threads(2)
    r, w, _ = parallel.init_from_env('gloo')
    assert (r, w) == (rank, world)
    lo, hi = parallel.shard_range(rank, world, N_TOTAL)
    states, acts, adv, tgt = _batch()
    params = network.init_params(ARCH, A, 3)
    specs = network.param_specs(ARCH, A)
    b = (hi - lo) * T
    _, g, _ = network.loss_and_grads(params, states[:, lo:hi].reshape(b, 84, 84, 4), acts[:, lo:hi].reshape(-1),
                                     adv[:, lo:hi].reshape(-1), tgt[:, lo:hi].reshape(-1), np.float32(0.02), ARCH, A)
    flat = torch.from_numpy(network.flatten_params(g, ARCH, A).copy())
    parallel.allreduce_mean_grads(flat, world)
    new, norm = _apply(params, flat.numpy() * np.float32(1.0 / world), specs)       # 1/world BEFORE the clip
    ret[rank] = (new, norm)
    torch.distributed.barrier()
    torch.distributed.destroy_process_group()


@pytest.mark.timeout(300)
def test_two_rank_update_equals_single_learner():
    world = 2
    port = _free_port()
    with mp.Manager() as mgr:
        ret = mgr.dict()
        mp.spawn(_worker, args=(world, port, ret), nprocs=world, join=True)
        results = [ret[r] for r in range(world)]
    states, acts, adv, tgt = _batch()
    params = network.init_params(ARCH, A, 3)
    specs = network.param_specs(ARCH, A)
    B = N_TOTAL * T
    _, g, _ = network.loss_and_grads(params, states.reshape(B, 84, 84, 4), acts.reshape(-1), adv.reshape(-1),
                                     tgt.reshape(-1), np.float32(0.02), ARCH, A)
    want, want_norm = _apply(params, network.flatten_params(g, ARCH, A), specs)
    p0 = network.flatten_params(params, ARCH, A)
    assert np.array_equal(results[0][0], results[1][0])                      # ranks stay bit-identical
    for new, norm in results:
        assert abs(norm - want_norm) <= 1e-5 * want_norm
        assert np.max(np.abs((new - p0) - (want - p0))) <= 1e-4 * np.max(np.abs(want - p0))


def test_shard_range():
    assert parallel.shard_range(0, 8, 256) == (0, 32) and parallel.shard_range(7, 8, 256) == (224, 256)
    assert parallel.shard_range(0, 1, 32) == (0, 32)
    with pytest.raises(ValueError):
        parallel.shard_range(0, 3, 32)
