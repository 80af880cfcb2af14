"""Experiment: does running the iteration as two concurrent half-batches (24 + 24 views on two streams) beat one 48-view
chain?  Two independent graph-replayed optimisers stand in for the two view groups (each also repeats the vertex work
and its own update, so this is a pessimistic bound for a real split)."""
import os, sys, warnings
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
warnings.filterwarnings("ignore")
import torch
from fmhr_b200 import synth
from fmhr_b200.ham import HamOptimizer
from fmhr_b200.render import render_views

dev = torch.device("cuda", 0)
scene = synth.build_scene("interhand_48x512x334", lambda *a: render_views(*a, device=dev))
n = scene["imgs"].shape[0]


def make(sel):
    c = lambda k, dt=torch.float32, s=None: torch.tensor(scene[k] if s is None else scene[k][s], dtype=dt, device=dev)
    return HamOptimizer(c("vertices"), c("faces", torch.int32), c("imgs", s=sel), c("masks", s=sel), c("valid_masks", s=sel),
                        c("w2cs", s=sel), c("projs", s=sel), c("sh_coeffs", s=sel), c("albedo"), scene["conf"],
                        use_graphs=True, n_views_global=n)


def timed(fn, k=300, warm=50):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(k):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / k


full = make(list(range(n)))
vf = torch.arange(n, dtype=torch.int32, device=dev)
print("one chain, 48 views: %.4f ms" % timed(lambda: full.step_phase_b(vf)))
for groups in (2, 3, 4):
    size = n // groups
    opts = [make(list(range(g * size, (g + 1) * size))) for g in range(groups)]
    vs = torch.arange(size, dtype=torch.int32, device=dev)
    streams = [torch.cuda.Stream(device=dev) for _ in range(groups)]
    for o in opts:
        o.step_phase_b(vs)  # capture on the default stream first
    torch.cuda.synchronize()

    def step():
        cur = torch.cuda.current_stream()
        for o, s in zip(opts, streams):
            s.wait_stream(cur)
            with torch.cuda.stream(s):
                o.step_phase_b(vs)
        for s in streams:
            cur.wait_stream(s)

    print("%d concurrent chains of %d views: %.4f ms" % (groups, size, timed(step)))
    one = timed(lambda: opts[0].step_phase_b(vs))
    print("   (a single %d-view chain alone: %.4f ms)" % (size, one))
