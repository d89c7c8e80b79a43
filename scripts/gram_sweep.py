"""Gram kernel alone over the effective-rank sweep shapes (BASELINE.json configs[3]): time, TFLOP/s against the
measured bf16 peak, and GB/s of algorithmic traffic (read X once + write G once) against the measured HBM peak.
Usage: python scripts/gram_sweep.py [out.json]"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from r3d_b200 import ops

SHAPES = [(128, 512, 512), (64, 1024, 512), (256, 256, 256), (256, 128, 512), (256, 4096, 128), (32, 1024, 1024),
          (16, 2048, 1024), (8, 2048, 2048), (16, 4096, 2048), (16, 2048, 4096)]


def main():
    pk = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))
    dev = torch.device("cuda:0")
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    rows = []
    for B, T, C in SHAPES:
        x = torch.randn(B, T, C, device=dev).relu_().to(torch.bfloat16)
        n, m = min(T, C), max(T, C)
        for _ in range(3):
            G = ops.gram(x, ops.GRAM_TCGEN05)
        ts = []
        for _ in range(10):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); G = ops.gram(x, ops.GRAM_TCGEN05); e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        ts.sort()
        ms = ts[len(ts) // 2]
        flops = 2.0 * n * n * m * B
        bytes_ = B * (T * C * 2 + n * n * 4)
        ref = torch.einsum("btc,bsc->bts", x[:1].float(), x[:1].float()) if T < C else \
            torch.einsum("btc,btd->bcd", x[:1].float(), x[:1].float())
        err = ((G[:1] - ref).abs().max() / ref.abs().max()).item()
        rows.append(dict(B=B, T=T, C=C, n=n, ms=ms, tflops=flops / ms / 1e9, frac_tensor=flops / ms / 1e9 / pk["bf16_tflops"],
                         gbs=bytes_ / ms / 1e6, frac_hbm=bytes_ / ms / 1e6 / pk["hbm_gbs"], rel_err=err))
        print(rows[-1], flush=True)
    if len(sys.argv) > 1:
        json.dump(rows, open(sys.argv[1], "w"), indent=1)


if __name__ == "__main__":
    main()
