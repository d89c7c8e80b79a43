"""Is the fused step capturable in a CUDA graph, and what does a replay cost compared with eager launches?"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from r3d_b200 import ops
B, T, C = 64, 512, 512
dev = torch.device("cuda:0"); dt = torch.bfloat16
torch.manual_seed(0)
step = ops.FuserStep(B, T, C, dt, dev)
c = torch.arange(C, device=dev, dtype=torch.float32)
bufs = [torch.stack([(torch.randn(B, T, C, device=dev).relu_() * (1 + c / C)).to(dt),
                     (torch.randn(B, T, C, device=dev).relu_() * (2 - c / C)).to(dt)]) for _ in range(4)]
gs = [torch.randn(B, T, 2, C, device=dev).to(dt) for _ in range(4)]
static_in, static_g = bufs[0].clone(), gs[0].clone()
for i in range(3): step(static_in, static_g)
torch.cuda.synchronize()
ref_er = step.er.clone(); ref_dg = step.dgrad.clone(); ref_out = step.out.clone()
def timed(fn, n=10):
    torch.cuda.synchronize(); e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(n): fn(i)
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
def eager(i):
    static_in.copy_(bufs[i % 4]); static_g.copy_(gs[i % 4]); step(static_in, static_g)
print("eager ms/step", round(timed(eager), 3))
g = torch.cuda.CUDAGraph()
s = torch.cuda.Stream()
s.wait_stream(torch.cuda.current_stream())
with torch.cuda.stream(s):
    step(static_in, static_g)
torch.cuda.current_stream().wait_stream(s)
torch.cuda.synchronize()
t0 = time.perf_counter()
with torch.cuda.graph(g):
    step(static_in, static_g)
print("capture s", round(time.perf_counter() - t0, 3))
static_in.copy_(bufs[0]); static_g.copy_(gs[0]); g.replay(); torch.cuda.synchronize()
print("graph == eager:", torch.equal(step.er, ref_er), torch.equal(step.dgrad, ref_dg), torch.equal(step.out, ref_out),
      float((step.er - ref_er).abs().max()))
def replay(i):
    static_in.copy_(bufs[i % 4]); static_g.copy_(gs[i % 4]); g.replay()
print("graph ms/step", round(timed(replay), 3))
