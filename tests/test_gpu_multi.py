"""Multi-GPU parity on real hardware: N NCCL ranks on one box, one process per GPU, the batch sharded by
clip.  Each rank runs the CUDA path on its shard; the parent compares against the single-process oracle
  * on the CONCATENATED batch for score_scope='global' (selected channels bit-exact, fused output, parameter gradients
    after GradBucket.allreduce, batch-mean effective rank from the packed statistic), and
  * PER SHARD for score_scope='local' (what nn.DataParallel does in the reference, main_utkinects.py:129).
Skipped when fewer than two GPUs are visible (`gpurun --gpus 2 -- python -m pytest tests/test_gpu_multi.py -m gpu`)."""
import os
import socket
import tempfile

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle import erank_oracle as EO             # noqa: E402
from oracle import fuser_oracle as O              # noqa: E402
from oracle.torch_port import PortCMFuser         # noqa: E402

B, T, C, HEADS = 6, 24, 64, 4


def _inputs():
    g = torch.Generator().manual_seed(77)
    c = torch.arange(C, dtype=torch.float32)
    pr, pd = torch.randperm(C, generator=g), torch.randperm(C, generator=g)
    rgb = (torch.relu(torch.randn(B, T, C, generator=g)) * (1 + c / C))[:, :, pr].contiguous()
    dep = (torch.relu(torch.randn(B, T, C, generator=g)) * (2 - c / C))[:, :, pd].contiguous()
    gy = torch.randn(B, T, C, generator=g)
    gst = torch.randn(B, T, 2, C, generator=g)
    return rgb, dep, gy, gst


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, sd_path, out_dir):
    import torch.distributed as dist
    import r3d_b200
    from r3d_b200 import dist as D, ops
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    os.environ.setdefault("NCCL_DEBUG", "WARN")
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    rgb, dep, gy, gst = _inputs()
    lo, hi = D.shard_bounds(B, world, rank)
    sd = torch.load(sd_path)
    res = {}
    for scope in ("global", "local"):
        f = r3d_b200.CMFuser(C, depth=1, num_heads=HEADS, score_scope=scope)
        f.load_state_dict(sd)
        f = f.to(dev).eval()
        r = rgb[lo:hi].to(dev).requires_grad_(True)
        d = dep[lo:hi].to(dev).requires_grad_(True)
        y = f({"rgb": r, "depth": d}, "test")
        y.backward(gy[lo:hi].to(dev))
        bucket = D.GradBucket(f.parameters())
        bucket.allreduce(average=False)          # sum over ranks = gradient of the whole batch
        res[scope] = {
            "idx_r": f.last_indices[0].cpu().numpy(), "idx_d": f.last_indices[1].cpu().numpy(),
            "y": y.detach().cpu().numpy(), "grad_rgb": r.grad.cpu().numpy(), "grad_dep": d.grad.cpu().numpy(),
            "param_grads": {n: p.grad.cpu().numpy() for n, p in f.named_parameters() if p.grad is not None},
        }
    # the fused step of bench.py on this shard: packed statistic all-reduced over NCCL (erank mean of the global batch)
    step = ops.FuserStep(hi - lo, T, C, torch.float32, dev)
    buf = torch.stack([rgb[lo:hi], dep[lo:hi]]).to(dev)
    out, er, dgrad = step(buf, gst[lo:hi].to(dev).contiguous())
    torch.cuda.synchronize()
    packed = step.packed.cpu().numpy()
    res["step"] = {"idx": step.idx.cpu().numpy(), "out": out.cpu().numpy(), "er": er.cpu().numpy(),
                   "dgrad": dgrad.cpu().numpy(), "erank_sum": packed[2 * C], "rows": packed[2 * C + 1]}
    torch.save(res, os.path.join(out_dir, f"rank{rank}.pt"))
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2])
def test_nccl_ranks_match_single_process_oracle(world):
    if not torch.cuda.is_available() or torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs")
    import torch.multiprocessing as mp
    from r3d_b200 import dist as D
    torch.manual_seed(0)
    ref = PortCMFuser(C, depth=1, num_heads=HEADS, variant="tokenfusion").eval()
    rgb, dep, gy, gst = _inputs()
    with tempfile.TemporaryDirectory() as tmp:
        sd_path = os.path.join(tmp, "sd.pt")
        torch.save(ref.state_dict(), sd_path)
        mp.spawn(_worker, args=(world, _free_port(), sd_path, tmp), nprocs=world, join=True)
        outs = [torch.load(os.path.join(tmp, f"rank{r}.pt"), weights_only=False) for r in range(world)]

    def oracle_run(r_, d_, g_):
        ref.zero_grad()
        r = r_.clone().requires_grad_(True)
        d = d_.clone().requires_grad_(True)
        y = ref({"rgb": r, "depth": d}, "test")
        y.backward(g_)
        return (y.detach().numpy(), r.grad.numpy(), d.grad.numpy(), [i.numpy() for i in ref.last_indices],
                {n: p.grad.numpy().copy() for n, p in ref.named_parameters() if p.grad is not None})

    # ---- global scope: the oracle on the concatenated batch
    y, gr, gd, idx, pg = oracle_run(rgb, dep, gy)
    for rank, o in enumerate(outs):
        lo, hi = D.shard_bounds(B, world, rank)
        g = o["global"]
        np.testing.assert_array_equal(g["idx_r"], idx[0])            # bit-exact, identical on every rank
        np.testing.assert_array_equal(g["idx_d"], idx[1])
        np.testing.assert_allclose(g["y"], y[lo:hi], rtol=1e-4, atol=1e-5)
        np.testing.assert_allclose(g["grad_rgb"], gr[lo:hi], rtol=1e-3, atol=2e-5)
        np.testing.assert_allclose(g["grad_dep"], gd[lo:hi], rtol=1e-3, atol=2e-5)
        for n, v in pg.items():
            if n in g["param_grads"]:
                np.testing.assert_allclose(g["param_grads"][n], v, rtol=1e-3, atol=1e-4 * max(1.0, np.abs(v).max()),
                                           err_msg=f"rank {rank} {n}")
    # every rank holds the same reduced gradients
    for n in outs[0]["global"]["param_grads"]:
        np.testing.assert_array_equal(outs[0]["global"]["param_grads"][n], outs[1]["global"]["param_grads"][n])

    # ---- local scope: the oracle per shard (nn.DataParallel semantics); parameter gradients still sum over shards
    pg_sum = None
    for rank, o in enumerate(outs):
        lo, hi = D.shard_bounds(B, world, rank)
        y, gr, gd, idx, pg = oracle_run(rgb[lo:hi], dep[lo:hi], gy[lo:hi])
        l = o["local"]
        np.testing.assert_array_equal(l["idx_r"], idx[0])
        np.testing.assert_array_equal(l["idx_d"], idx[1])
        np.testing.assert_allclose(l["y"], y, rtol=1e-4, atol=1e-5)
        np.testing.assert_allclose(l["grad_rgb"], gr, rtol=1e-3, atol=2e-5)
        pg_sum = pg if pg_sum is None else {n: pg_sum[n] + v for n, v in pg.items()}
    for n, v in pg_sum.items():
        if n in outs[0]["local"]["param_grads"]:
            np.testing.assert_allclose(outs[0]["local"]["param_grads"][n], v, rtol=1e-3,
                                       atol=1e-4 * max(1.0, np.abs(v).max()), err_msg=n)

    # ---- the fused step: selection from the all-reduced packed statistic, exchange, erank mean of the global batch
    rn, dn = rgb.numpy(), dep.numpy()
    st_ref, ir, idd = O.token_fusion("tokenfusion", rn, dn, "test", return_indices=True)
    er_ref = np.concatenate([EO.erank(rn), EO.erank(dn)])
    for rank, o in enumerate(outs):
        lo, hi = D.shard_bounds(B, world, rank)
        s = o["step"]
        np.testing.assert_array_equal(s["idx"][0], ir)
        np.testing.assert_array_equal(s["idx"][1], idd)
        np.testing.assert_array_equal(s["out"], st_ref[lo:hi])
        np.testing.assert_allclose(s["er"], np.concatenate([EO.erank(rn[lo:hi]), EO.erank(dn[lo:hi])]), rtol=1e-4)
        assert s["rows"] == B * T
        np.testing.assert_allclose(s["erank_sum"] / (2 * B), er_ref.mean(), rtol=1e-4)
