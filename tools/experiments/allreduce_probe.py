"""How long does the update's gradient all-reduce take by itself on this box, and what does NCCL run for it?  (torchrun, G ranks)
Times torch.distributed.all_reduce(SUM) of the flat fp32 gradient (Nature: 1,686,693 floats = 6.75 MB; its fc + heads tail and its
conv head as the engine splits them) with CUDA events, queued back to back, for the default communicator and for
communicators limited to a few CTAs (ProcessGroupNCCL.Options.config.max_ctas) -- the persistent weight-gradient kernels leave
paacb_set_sm_reserve SMs free, and a collective that wants more CTAs than that waits for a kernel boundary.
NCCL_DEBUG=INFO NCCL_DEBUG_SUBSYS=TUNING,INIT shows the algorithm / protocol (rank 0's lines are kept)."""
import json, os, sys, statistics
import torch
import torch.distributed as dist

rank, world, local = int(os.environ['RANK']), int(os.environ['WORLD_SIZE']), int(os.environ['LOCAL_RANK'])
torch.cuda.set_device(local)
dev = torch.device('cuda', local)
dist.init_process_group('nccl', device_id=dev)
P, HEAD = 1686693, 77920           # Nature, 6 actions: all parameters; conv1..conv3 weights + biases (the head of the flat buffer)
g = torch.randn((P,), device=dev)


def time_ar(t, group, iters=40):
    for _ in range(5):
        dist.all_reduce(t, group=group)
    torch.cuda.synchronize(); dist.barrier()
    evs = []
    for _ in range(iters):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); dist.all_reduce(t, group=group); e1.record()
        evs.append((e0, e1))
    torch.cuda.synchronize()
    ms = [a.elapsed_time(b) for a, b in evs]
    t_med = torch.tensor([statistics.median(ms)], device=dev)
    dist.all_reduce(t_med, op=dist.ReduceOp.MAX)
    return float(t_med) * 1e3


rows = []
groups = [('default', None)]
for ctas in (4, 8, 16):
    try:
        o = dist.ProcessGroupNCCL.Options()
        o.config.max_ctas = ctas
        o.config.min_ctas = 1
        groups.append(('max_ctas=%d' % ctas, dist.new_group(ranks=list(range(world)), pg_options=o)))
    except Exception as e:
        if rank == 0:
            print('max_ctas=%d unavailable: %s' % (ctas, str(e)[:120]), file=sys.stderr)
for name, grp in groups:
    for what, t in (('whole gradient 6.75 MB', g), ('fc + heads tail 6.44 MB', g[HEAD:]), ('conv head 0.31 MB', g[:HEAD]), ('one float', g[:1])):
        us = time_ar(t, grp)
        rows.append({'communicator': name, 'buffer': what, 'bytes': t.numel() * 4, 'median_us_max_over_ranks': round(us, 1),
                     'algbw_gbs': round(t.numel() * 4 / us / 1e3, 1)})
if rank == 0:
    print(json.dumps({'world': world, 'nccl': '.'.join(map(str, torch.cuda.nccl.version())), 'rows': rows}, indent=1))
dist.destroy_process_group()
