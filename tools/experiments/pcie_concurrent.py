"""Do the GPUs of one box share their host links?  Run under torchrun with G ranks (G = 1, 2, 4, 8): after a barrier every
rank times, CONCURRENTLY with the others, (a) a plain pinned host->device copy and (b) the zero-copy preprocessing kernel
reading 4096 raw frame pairs from its own pinned buffer, and rank 0 prints the per-rank and aggregate GB/s.
    python -m torch.distributed.run --nnodes=1 --nproc-per-node G --master-addr 127.0.0.1 --master-port 29561 tools/experiments/pcie_concurrent.py
"""
import ctypes as C
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch

from paac_b200 import _lib
from paac_b200.policy_v_network import NaturePolicyVNetwork


def main():
    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    if world > 1:
        torch.distributed.init_process_group('nccl', device_id=dev)
    N = 4096
    host = torch.randint(0, 256, (N, 1, 2, 210, 160), dtype=torch.uint8).pin_memory()
    sel = N * 2 * 84 * 160
    hc = torch.randint(0, 256, (sel,), dtype=torch.uint8).pin_memory()
    dc = torch.empty((sel,), dtype=torch.uint8, device=dev)
    conf = dict(name='x', num_actions=6, clip_norm=3.0, clip_norm_type='global', device='/gpu:%d' % local,
                entropy_regularisation_strength=0.02, seed=3, math='bf16x3')
    net = NaturePolicyVNetwork(conf)
    prev = torch.zeros((N, 84, 84, 4), dtype=torch.uint8, device=dev)
    nxt = torch.zeros_like(prev)
    st = torch.cuda.current_stream(dev)

    def k1():
        _lib.check(net._lib.paacb_preprocess_u8(net.ctx, C.c_void_p(host.data_ptr()), 1, None, _lib.ptr(prev), _lib.ptr(nxt), N,
                                                C.c_void_p(st.cuda_stream)), 'k1')

    def timeit(fn, k=20):
        fn(); torch.cuda.synchronize()
        if world > 1:
            torch.distributed.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(k):
            fn()
        e1.record(); torch.cuda.synchronize()
        return sel / (e0.elapsed_time(e1) / k) / 1e6

    res = torch.tensor([timeit(lambda: dc.copy_(hc, non_blocking=True)), timeit(k1)], device=dev)
    allr = [torch.zeros_like(res) for _ in range(world)]
    if world > 1:
        torch.distributed.all_gather(allr, res)
    else:
        allr = [res]
    if rank == 0:
        rows = [[float(x) for x in r.tolist()] for r in allr]
        try:
            import pynvml
            pynvml.nvmlInit()
            aff = [list(pynvml.nvmlDeviceGetCpuAffinity(pynvml.nvmlDeviceGetHandleByIndex(i), 4)) for i in range(world)]
        except Exception:
            aff = None
        print(json.dumps({'world': world, 'bytes_per_transfer': sel,
                          'per_rank_gbs': [{'pinned_h2d_copy': r[0], 'zero_copy_k1': r[1]} for r in rows],
                          'aggregate_gbs': {'pinned_h2d_copy': sum(r[0] for r in rows), 'zero_copy_k1': sum(r[1] for r in rows)},
                          'nvml_cpu_affinity_masks': aff}), flush=True)
    if world > 1:
        torch.distributed.barrier()
        torch.distributed.destroy_process_group()


if __name__ == '__main__':
    main()
