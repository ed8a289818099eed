"""Host<->device copy bandwidth of the GPU box (pinned memory): each direction alone and both at once.
The end-to-end bench line moves 20 B in and 16 B out per particle-step; this is its roofline."""
import json
import sys
import torch

n = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
h_in = torch.empty(n * 20, dtype=torch.uint8).pin_memory()
h_out = torch.empty(n * 16, dtype=torch.uint8).pin_memory()
d_in = torch.empty(n * 20, dtype=torch.uint8, device='cuda')
d_out = torch.empty(n * 16, dtype=torch.uint8, device='cuda')
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()


def timed(fn, reps=10):
    fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    for s in (s1, s2):
        torch.cuda.current_stream().wait_stream(s)
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps


def h2d():
    with torch.cuda.stream(s1):
        d_in.copy_(h_in, non_blocking=True)


def d2h():
    with torch.cuda.stream(s2):
        h_out.copy_(d_out, non_blocking=True)


def both():
    h2d()
    d2h()


res = {}
t = timed(h2d); res['h2d_alone_GBps'] = n * 20 / t / 1e6
t = timed(d2h); res['d2h_alone_GBps'] = n * 16 / t / 1e6
t = timed(both); res['both_ms_per_step'] = t
res['both_h2d_GBps'] = n * 20 / t / 1e6
res['both_d2h_GBps'] = n * 16 / t / 1e6
res['e2e_ceiling_particle_steps_per_s'] = n / (t * 1e-3)
print(json.dumps(res))
