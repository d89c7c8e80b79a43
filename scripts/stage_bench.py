"""Isolated timing of the streaming stages at the headline shape: N back-to-back launches between two CUDA events,
rotating over 4 input sets (4 x 134 MB > L2) so every launch reads HBM.  Prints us/launch and % of the measured HBM peak.
Usage: python scripts/stage_bench.py [row_chunk_mult]"""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from r3d_b200 import _lib
from r3d_b200.ops import _p, _dt, _stream, check

if len(sys.argv) > 1:
    _lib.set_option("row_chunk_mult", float(sys.argv[1]))
pk = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
B, T, C = 64, 512, 512
dev = torch.device("cuda:0")
L = _lib.lib()
rows = B * T
sets = [(torch.randn(B, T, C, device=dev).relu_().bfloat16(), torch.randn(B, T, C, device=dev).relu_().bfloat16(),
         torch.randn(B, T, 2, C, device=dev).bfloat16()) for _ in range(4)]
ws = torch.empty(L.r3d_score_workspace_floats(rows, C), dtype=torch.float32, device=dev)
out = torch.empty(B, T, 2, C, dtype=torch.bfloat16, device=dev)
d_r = torch.empty(B, T, C, dtype=torch.bfloat16, device=dev)
d_d = torch.empty_like(d_r)
k = C // 4
idx = torch.stack([torch.randperm(C, device=dev)[:k], torch.randperm(C, device=dev)[:k]]).contiguous()
N = 40


def timeit(fn):
    for i in range(4):
        fn(i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(N):
        fn(i)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / N * 1e3


def score(i):
    r, d, _ = sets[i % 4]
    check(L.r3d_channel_score_partial(_p(r), _p(d), rows, C, _dt(r), _p(ws), _stream()))


def fwd(i):
    r, d, _ = sets[i % 4]
    check(L.r3d_exchange_fwd(_p(r), _p(d), _p(idx[0]), _p(idx[1]), k, None, None, 0, _p(out), rows, C, _dt(r), _stream()))


def bwd(i):
    _, _, g = sets[i % 4]
    check(L.r3d_exchange_bwd(_p(g), None, None, _p(idx[0]), _p(idx[1]), k, None, None, None, 0, _p(d_r), _p(d_d), None,
                             rows, C, _dt(g), _stream()))


N_el = B * T * C
for name, fn, nbytes in (("score_partial", score, 2 * N_el * 2), ("exchange_fwd", fwd, 4 * N_el * 2),
                         ("exchange_bwd", bwd, 4 * N_el * 2)):
    us = timeit(fn)
    print(f"{name:14s} {us:7.2f} us  {nbytes / us / 1e3:8.1f} GB/s  {nbytes / us / 1e3 / pk * 100:5.1f} % of measured HBM peak")
