#!/usr/bin/env python
"""bench.py -- fused clips/sec of the R3D token-fuser + effective-rank hot path.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

Workload (BASELINE.json configs[1], the config the metric is quoted on): per GPU,
B=64 clips of RGB+depth features, T=512 tokens, C=512 channels, bf16.  One "step" is
one TRAINING pass of the fuser path over that batch (r3d_b200.ops.FuserTrainStep):
    erank(rgb), erank(depth)  [Gram -> block Jacobi -> refinement -> entropy/exp]
    CMFuser.forward: channel score -> [all-reduce] -> bottom-k -> exchange/stack -> Block (LayerNorm, V / proj /
                     MLP GEMMs on tcgen05) -> LayerNorm -> mean over the two modality tokens            (forward)
    backward of an upstream gradient through CMFuser (input + parameter gradients), d(mean erank)/dX
    accumulated, and with N > 1 the all-reduce of the fuser's parameter gradients                    (backward)
N > 1: one process per GPU (torchrun), batch-sharded: weak scaling by default (B=64 per GPU), strong scaling with
--global-batch G (configs[2]: G=512).  Collectives: one all-reduce of the packed (2C + 2)-float score / erank
statistic in the forward, one all-reduce of the flat fp32 parameter gradients (13.7 MB at C=512) in the backward,
issued on a side stream beside the effective-rank backward.

`value`  : clips/s with inputs resident in HBM (CUDA events, max over ranks).
`e2e`    : same metric with each step's inputs copied from pinned host memory and the
           step's erank statistic read back, inside the timed region.
`--impl reference`: the reference's CPU path on the host cores: the reference's own CMFuser (staged under oracle/_ref by
           oracle/build_ref.py; the torch-CPU port of oracle/torch_port.py when that directory is absent) forward +
           backward, plus the svdvals erank restatement (the reference has no effective-rank code) forward + backward.
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "fused clips/sec (eff-rank+token-fuse fwd/bwd)"
UNIT = "clips/s"
B, T, C = 64, 512, 512
NSETS = 4   # rotating input sets: 4 x (67 MB inputs + 67 MB upstream grad) >> 126 MB L2


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--cpu-sample", type=int, default=16, help="clips per CPU-baseline step")
    ap.add_argument("--gram", default="auto", choices=["auto", "tcgen05", "simt"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--opt", action="append", default=[], help="library tuning knob key=value (r3d_set_option)")
    ap.add_argument("--global-batch", type=int, default=0, help="strong scaling: total clips per step, split over the ranks")
    ap.add_argument("--no-yardstick", action="store_true", help="skip the same-GPU library eigensolver yardstick")
    ap.add_argument("--no-extras", action="store_true",
                    help="timed steps only: no isolated stages, full-fuser comparison, yardsticks, C-ABI or CPU legs "
                         "(for ncu launch lists)")
    return ap.parse_args()


def synth_host(seed, b, dtype):
    import torch
    g = torch.Generator().manual_seed(seed)
    c = torch.arange(C, dtype=torch.float32)
    rgb = torch.relu(torch.randn(b, T, C, generator=g)) * (1 + c / C)
    dep = torch.relu(torch.randn(b, T, C, generator=g)) * (2 - c / C)
    return torch.stack([rgb, dep]).to(dtype)      # (2, b, T, C)


# ------------------------------------------------------------------------------------
# CPU leg: the reference's own algorithm on the host cores (oracle port; "kind": "port")
# ------------------------------------------------------------------------------------
def cpu_step_factory(sample):
    """-> (step, kind): the reference's own CMFuser when oracle/_ref (or /root/reference) is present, else the port."""
    import torch
    from oracle.torch_port import PortCMFuser, erank_torch
    from oracle import ref_loader
    torch.set_num_threads(os.cpu_count() or 1)
    torch.manual_seed(0)
    cls = ref_loader.load("tokenfusion")
    if cls is not None:
        fuser, kind = cls(dim=C, depth=1, num_heads=8), "reference"
    else:
        fuser, kind = PortCMFuser(C, depth=1, num_heads=8, variant="tokenfusion"), "port"
    fuser.train()
    fuser.embd_drop.p = 0.0                                     # dropout off on both arms (parity, determinism)
    buf = synth_host(1234, sample, torch.bfloat16).float()     # fp32 on the bf16-rounded inputs
    gy = torch.randn(sample, T, C, generator=torch.Generator().manual_seed(4321))

    def step():
        rgb = buf[0].clone().requires_grad_(True)
        dep = buf[1].clone().requires_grad_(True)
        fuser.zero_grad(set_to_none=True)
        y = fuser({"rgb": rgb, "depth": dep}, "test")
        er = torch.cat([erank_torch(rgb), erank_torch(dep)])
        torch.autograd.backward([y, er], [gy, torch.full_like(er, 1.0 / er.numel())])
        return float(er.detach().mean())

    return step, kind


def cpu_gram_route_ms(sample):
    """Extra yardstick: the effective-rank forward through the cheaper CPU route (fp32 Gram + eigvalsh, BASELINE.md
    section 2 measured it 2.6x faster than svdvals at this shape) -- ms per clip pair (rgb + depth)."""
    import torch
    x = synth_host(1234, sample, torch.bfloat16).float().reshape(2 * sample, T, C)
    t0 = time.perf_counter()
    lam = torch.linalg.eigvalsh(x.transpose(1, 2) @ x)
    sg = lam.clamp_min(0).sqrt()
    p = sg / sg.sum(-1, keepdim=True)
    _ = torch.exp(-(p * torch.log(p.clamp_min(1e-30))).sum(-1))
    return (time.perf_counter() - t0) * 1e3 / sample


def run_cpu(steps, warmup, sample):
    step, kind = cpu_step_factory(sample)
    for _ in range(warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(steps):
        step()
    dt = time.perf_counter() - t0
    return sample * steps / dt, dt / steps * 1e3, kind


CPU_SAMPLE_TEXT = ("{n} clips/step of the same shape (T=512, C=512): the reference's CMFuser forward + backward "
                   "({kind}; train mode, dropout 0, eval-branch score) + the svdvals effective-rank restatement forward "
                   "+ backward for both modalities, fp32 on the bf16-rounded inputs")


def main_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import torch
    sample = args.cpu_sample
    val, ms, kind = run_cpu(args.steps, args.warmup, sample)
    cores = torch.get_num_threads()
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"DARai RGB+depth fuser training step fwd+bwd, T={T}, C={C} (BASELINE.json configs[1]); CPU "
                               f"step = {sample} clips of that shape", "B_per_step": sample, "cpu_clips_per_step": sample,
                   "T": T, "C": C},
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": kind,
                         "sample": CPU_SAMPLE_TEXT.format(n=sample, kind="oracle/_ref: unmodified reference module"
                                                          if kind == "reference" else "torch-CPU port") +
                                   f" x {args.steps} steps",
                         "erank_part": "restatement (the reference has no effective-rank code, SURVEY.md F1)"},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


# ------------------------------------------------------------------------------------
class ClockSampler:
    """SM clock and throttle reasons of one GPU DURING the timed region: an NVML polling thread (10 ms period, so even a
    250 ms region yields >20 samples); falls back to `nvidia-smi -lms` when pynvml is missing."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
    BITS = {"sw_power_cap": 0x4, "hw_slowdown": 0x8, "sw_thermal_slowdown": 0x20, "hw_thermal_slowdown": 0x40}

    def __init__(self, gpu_index):
        self.idx = gpu_index
        self.p = None
        self.f = None
        self.thread = None
        self.stop_flag = False
        self.sm, self.mx, self.reasons = [], [], set()
        # CUDA_VISIBLE_DEVICES remapping: NVML wants the physical index
        vis = os.environ.get("CUDA_VISIBLE_DEVICES")
        self.phys = gpu_index
        if vis:
            try:
                self.phys = int(vis.split(",")[gpu_index])
            except Exception:
                self.phys = gpu_index

    def _poll(self):
        import pynvml
        h = pynvml.nvmlDeviceGetHandleByIndex(self.phys)
        try:
            mx = pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM)
        except Exception:
            mx = None
        while not self.stop_flag:
            try:
                self.sm.append(float(pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM)))
                if mx is not None:
                    self.mx.append(float(mx))
                r = pynvml.nvmlDeviceGetCurrentClocksThrottleReasons(h)
                for nm, bit in self.BITS.items():
                    if r & bit:
                        self.reasons.add(nm)
            except Exception:
                pass
            time.sleep(0.01)

    def start(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            self.thread = threading.Thread(target=self._poll, daemon=True)
            self.thread.start()
            return
        except Exception:
            self.thread = None
        try:
            self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
            self.p = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-i",
                                       str(self.phys), "-lms", "100"], stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.thread is not None:
            self.stop_flag = True
            self.thread.join(timeout=2)
            sm, mx, reasons = list(self.sm), list(self.mx), set(self.reasons)
        elif self.p is not None:
            self.p.terminate()
            try:
                self.p.wait(timeout=5)
            except Exception:
                self.p.kill()
            self.f.flush()
            self.f.seek(0)
            sm, mx, reasons = [], [], set()
            names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
            for ln in self.f.read().splitlines():
                parts = [x.strip() for x in ln.split(",")]
                if len(parts) < 9:
                    continue
                try:
                    sm.append(float(parts[1])); mx.append(float(parts[2]))
                except ValueError:
                    continue
                for nm, v in zip(names, parts[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(nm)
            self.f.close()
            try:
                os.unlink(self.f.name)
            except OSError:
                pass
        else:
            return out
        if sm:
            sm.sort()
            out.update(sm_mhz=sm[len(sm) // 2], sm_max_mhz=(max(mx) if mx else None), reasons=sorted(reasons),
                       samples=len(sm))
        return out


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return {"hbm_gbs": d["hbm_gbs"], "bf16_tflops": d["bf16_tflops"], "bf16_tflops_sustained":
                d.get("bf16_tflops_sustained", d["bf16_tflops"]), "source": "measured"}
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "source": "fallback"}


def stage_table(prof, steps, world, panel_tiles=None, B=B):
    """Per-stage algorithmic bytes / flops per launch -> achieved and fraction of peak.
    The Jacobi launch sequence is fixed and launches after convergence exit at once, so for the panel passes the
    algorithmic bytes of the region are counted by the kernel itself (tiles actually processed x 64 KB) and divided
    by the stage's total time -- not launches x one full pass."""
    pk = peaks()
    es = 2                       # bf16 inputs
    N = B * T * C                # elements per modality per rank (B = clips per rank and step)
    n = min(T, C); m = max(T, C); nb2 = 2 * B
    npad = ((n + 63) // 64) * 64
    alg = {
        "score_partial": ("hbm", 2 * N * es),
        "exchange_fwd": ("hbm", 4 * N * es),
        "exchange_bwd": ("hbm", 4 * N * es),
        "gram": ("tensor", 2.0 * n * n * m * nb2),
        "refine_y": ("tensor", 2.0 * n * n * m * nb2),
        "bwd_gemm": ("tensor", 2.0 * n * n * m * nb2),
        # one panel pass: read + write one np x np fp32 matrix per sample (G pass 1, G pass 2, V update)
        "jacobi_update": ("hbm", nb2 * npad * npad * 4 * 2),
        "jacobi_vupdate": ("hbm", nb2 * npad * npad * 4 * 2),
        "sigma": ("hbm", nb2 * (n * m + n * n) * 4),
    }
    out = {}
    for name, rec in prof.items():
        launches = max(rec["launches"], 1)
        avg_ms = rec["ms"] / launches
        e = {"ms_per_step": rec["ms"] / steps, "launches_per_step": rec["launches"] / steps, "avg_launch_us": avg_ms * 1e3}
        if panel_tiles and name in ("jacobi_update", "jacobi_vupdate") and rec["ms"] > 0:
            tiles = panel_tiles[0 if name == "jacobi_update" else 1]
            nbytes = tiles * 65536.0                       # 128 x 64 fp32 read + written per tile
            ach = nbytes / (rec["ms"] * 1e-3) / 1e9
            e.update(bound="hbm", achieved=ach, unit="GB/s", peak=pk["hbm_gbs"], frac=ach / pk["hbm_gbs"],
                     tiles_per_step=tiles / steps, algorithmic_mb_per_step=nbytes / steps / 1e6,
                     full_pass_equivalents_per_step=nbytes / steps / alg[name][1])
        elif name in alg and avg_ms > 0:
            bound, work = alg[name]
            if bound == "hbm":
                ach = work / (avg_ms * 1e-3) / 1e9
                e.update(bound="hbm", achieved=ach, unit="GB/s", peak=pk["hbm_gbs"], frac=ach / pk["hbm_gbs"])
            else:
                ach = work / (avg_ms * 1e-3) / 1e12
                e.update(bound="tensor", achieved=ach, unit="TFLOP/s", peak=pk["bf16_tflops_sustained"],
                         frac=ach / pk["bf16_tflops_sustained"])
        out[name] = e
    return out, pk


def isolated_stage_numbers(dev_in, dev_g, pk, n=40):
    """Kernel-only time of the streaming stages: n back-to-back launches between two CUDA events (the per-stage events
    of the table above include ~8 us of launch/event latency, which is most of a 20 us kernel), rotating over the
    NSETS input sets so that every launch reads HBM.  Fractions are of the measured copy bandwidth."""
    import torch
    from r3d_b200 import _lib
    from r3d_b200.ops import _p, _dt, _stream, check
    L = _lib.lib()
    dev = dev_in[0].device
    rows, k = B * T, C // 4
    ws = torch.empty(L.r3d_score_workspace_floats(rows, C), dtype=torch.float32, device=dev)
    out = torch.empty(B, T, 2, C, dtype=dev_in[0].dtype, device=dev)
    d_r = torch.empty(B, T, C, dtype=dev_in[0].dtype, device=dev)
    d_d = torch.empty_like(d_r)
    idx = torch.stack([torch.randperm(C, device=dev)[:k], torch.randperm(C, device=dev)[:k]]).contiguous()

    def score(i):
        x = dev_in[i % NSETS]
        check(L.r3d_channel_score_partial(_p(x[0]), _p(x[1]), rows, C, _dt(x), _p(ws), _stream()))

    def fwd(i):
        x = dev_in[i % NSETS]
        check(L.r3d_exchange_fwd(_p(x[0]), _p(x[1]), _p(idx[0]), _p(idx[1]), k, None, None, 0, _p(out), rows, C, _dt(x),
                                 _stream()))

    def bwd(i):
        g = dev_g[i % NSETS]
        check(L.r3d_exchange_bwd(_p(g), None, None, _p(idx[0]), _p(idx[1]), k, None, None, None, 0, _p(d_r), _p(d_d),
                                 None, rows, C, _dt(g), _stream()))

    def ref_sum(i):                       # library yardstick for a read-only pass over the same bytes
        dev_in[i % NSETS].sum(dtype=torch.float32)

    N_el, es = B * T * C, 2
    res = {}
    # one real Jacobi panel pass ALONE (test hook: V pass + the two G passes of one round on scratch buffers, all
    # tasks rotating), timed by the library's stage events: what the kernel reaches without the V/G contention and
    # the already-converged launches of the real step
    try:
        nbm, npad = 2 * B, ((min(T, C) + 127) // 128) * 128
        nt = npad // 64
        Gd = torch.randn(nbm, npad, npad, device=dev)
        Hd, Vd = torch.empty_like(Gd), torch.randn_like(Gd)
        Qd = torch.eye(64, device=dev).repeat(nbm, nt, 1, 1).contiguous()
        scratch = torch.zeros(nbm * 32 + nbm * nt, dtype=torch.int32, device=dev)
        for i in range(2):
            check(L.r3d_debug_panel_round(_p(Gd), _p(Hd), _p(Vd), _p(Qd), nbm, npad, 1 + i, _p(scratch), _stream()))
        _lib.profile_enable(True)
        _lib.profile_read(reset=True)
        for i in range(8):
            check(L.r3d_debug_panel_round(_p(Gd), _p(Hd), _p(Vd), _p(Qd), nbm, npad, 1 + i % 7, _p(scratch), _stream()))
        pr = _lib.profile_read(reset=True)
        _lib.profile_enable(False)
        ms = pr["jacobi_update"]["ms"] + pr["jacobi_vupdate"]["ms"]
        nl = pr["jacobi_update"]["launches"] + pr["jacobi_vupdate"]["launches"]
        us = ms / nl * 1e3
        gbs = nbm * npad * npad * 8 / us / 1e3
        res["jacobi_panel_pass_alone"] = {"avg_launch_us": us, "achieved": gbs, "unit": "GB/s", "peak": pk["hbm_gbs"],
                                          "frac": gbs / pk["hbm_gbs"]}
        del Gd, Hd, Vd, Qd
    except Exception as ex:   # pragma: no cover
        res["jacobi_panel_pass_alone"] = {"error": repr(ex)[:200]}
    for name, fn, nbytes in (("score_partial", score, 2 * N_el * es), ("exchange_fwd", fwd, 4 * N_el * es),
                             ("exchange_bwd", bwd, 4 * N_el * es), ("torch_sum_same_bytes", ref_sum, 2 * N_el * es)):
        for i in range(4):
            fn(i)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(n):
            fn(i)
        e1.record()
        torch.cuda.synchronize()
        us = e0.elapsed_time(e1) / n * 1e3
        gbs = nbytes / us / 1e3
        res[name] = {"avg_launch_us": us, "achieved": gbs, "unit": "GB/s", "peak": pk["hbm_gbs"],
                     "frac": gbs / pk["hbm_gbs"]}
    return res


def full_fuser_numbers(dev, dtype, steps=10):
    """Whole CMFuser forward+backward (token fusion + Block + LN + mean) at the headline shape: this repo's module
    vs the reference's own op sequence (oracle/torch_port.py: full qkv GEMM, 2x2 masked softmax, clone + index_put
    + stack) run eagerly on the same GPU.  Reported as an extra; the Block still uses library GEMMs."""
    import torch
    import r3d_b200
    from oracle.torch_port import PortCMFuser
    out = {}
    buf = synth_host(99, B, dtype).to(dev)
    gy = torch.randn(B, T, C, device=dev, dtype=dtype)
    torch.manual_seed(0)
    ref = PortCMFuser(C, depth=1, num_heads=8, variant="tokenfusion").to(dev).to(dtype).train()
    ours = r3d_b200.CMFuser(C, depth=1, num_heads=8).to(dev).to(dtype).train()
    ours.load_state_dict(ref.state_dict())
    ref.embd_drop.p = ours.embd_drop.p = 0.0

    r = buf[0].clone().requires_grad_(True)
    d = buf[1].clone().requires_grad_(True)

    def run(mod):
        r.grad = None
        d.grad = None
        for p_ in mod.parameters():
            p_.grad = None
        y = mod({"rgb": r, "depth": d}, "test")
        y.backward(gy)
        return y

    for name, mod in (("ours", ours), ("reference_ops_eager_gpu", ref)):
        try:
            for _ in range(3):
                y = run(mod)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(steps):
                y = run(mod)
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / steps
            out[name] = {"ms_per_step": ms, "clips_per_s": B / ms * 1e3}
        except Exception as ex:   # pragma: no cover
            out[name] = {"error": repr(ex)[:200]}
    return out


def c_abi_e2e(host_in, Bl, dtype, steps):
    """The hot path the reference-facing C entry covers -- erank of both modalities, score -> bottom-k -> exchange,
    exchange backward + erank gradient; NOT the Block, which lives in the nn.Module -- through ONE C-ABI call per step
    with HOST buffers (r3d_fuser_step_host: H2D, kernels, D2H, stream sync inside the call), pinned memory."""
    import torch
    from r3d_b200 import ops
    gst = torch.randn(Bl, T, 2, C, generator=torch.Generator().manual_seed(7)).to(dtype).pin_memory()
    out = (torch.empty(Bl, T, 2, C, dtype=dtype).pin_memory(), torch.empty(2 * Bl, dtype=torch.float32).pin_memory(),
           torch.empty(Bl, T, C, dtype=dtype).pin_memory(), torch.empty(Bl, T, C, dtype=dtype).pin_memory(),
           torch.empty(2, C // 4, dtype=torch.int64).pin_memory())
    for i in range(2):
        ops.fuser_step_host(host_in[i % NSETS][0], host_in[i % NSETS][1], gst, out=out)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for i in range(steps):
        ops.fuser_step_host(host_in[i % NSETS][0], host_in[i % NSETS][1], gst, out=out)
    dt_s = time.perf_counter() - t0                       # the call synchronises its stream: wall clock is device-complete
    es = 2
    return {"value": Bl * steps / dt_s, "unit": UNIT, "ms_per_step": dt_s / steps * 1e3,
            "h2d_bytes_per_step": int(4 * Bl * T * C * es), "d2h_bytes_per_step": int(4 * Bl * T * C * es + 2 * Bl * 4),
            "entry": "r3d_fuser_step_host (include/r3d_b200.h)",
            "note": "token-fuse + effective-rank fwd/bwd only (the Block is nn.Module-side); copies not overlapped with "
                    "compute -- one synchronous call per step"}


def library_yardstick(buf, dev):
    """The same eigen / singular-value problem on the same GPU through the libraries (cuSOLVER behind torch.linalg):
    the 2B = 128 samples of 512 x 512 of one step.  Forward only (no gradient), one warm-up + one timed call each --
    answers "does the hand-written eigensolver beat the library", independent of the CPU baseline."""
    import torch
    x = buf.reshape(-1, T, C).float()
    out = {"batch": int(x.shape[0]), "shape": [T, C], "dtype": "fp32"}

    def t1(fn):
        fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1)

    try:
        G = x.transpose(1, 2) @ x
        out["torch_linalg_eigh_on_gram_ms"] = t1(lambda: torch.linalg.eigh(G))
        out["torch_linalg_eigvalsh_on_gram_ms"] = t1(lambda: torch.linalg.eigvalsh(G))
        out["torch_linalg_svdvals_ms"] = t1(lambda: torch.linalg.svdvals(x))
    except Exception as ex:   # pragma: no cover
        out["error"] = repr(ex)[:200]
    return out


def main_ours(args):
    import torch
    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py --impl ours needs a CUDA device (no CPU fallback)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        if os.environ.get("NCCL_DEBUG", "VERSION").upper() in ("VERSION", "WARN"):
            # NCCL prints its version banner on stdout at every level >= VERSION (WARN included); stdout must carry
            # the single JSON line only.  An explicit INFO/TRACE request is respected.
            os.environ["NCCL_DEBUG"] = "NONE"
        dist.init_process_group("nccl", device_id=dev)
    import r3d_b200
    from r3d_b200 import _lib, ops

    for kv in args.opt:
        k_, v_ = kv.split("=")
        _lib.set_option(k_, float(v_))
    dtype = torch.bfloat16
    gram_impl = {"auto": ops.GRAM_TCGEN05, "tcgen05": ops.GRAM_TCGEN05, "simt": ops.GRAM_SIMT}[args.gram]
    Bl = B                                        # clips per rank and step
    if args.global_batch:
        if args.global_batch % world:
            raise SystemExit("--global-batch must be a multiple of the number of ranks")
        Bl = args.global_batch // world
    torch.manual_seed(0)                          # identical (replicated) fuser weights on every rank
    fuser = r3d_b200.CMFuser(C, depth=1, num_heads=8, score_scope="global").to(dev).to(dtype).train()
    fuser.embd_drop.p = 0.0                       # dropout off on both arms (parity, determinism)
    step = ops.FuserTrainStep(fuser, Bl, T, C, dtype, dev, gram_impl=gram_impl)
    host_in = [synth_host(1234 + rank * 100 + i, Bl, dtype).pin_memory() for i in range(NSETS)]
    dev_in = [h.to(dev) for h in host_in]
    gg = torch.Generator(device=dev).manual_seed(4321 + rank)
    dev_g = [torch.randn(Bl, T, C, generator=gg, device=dev).to(dtype) for _ in range(NSETS)]
    stage_buf = torch.empty_like(dev_in[0])
    er_host = torch.empty(2 * Bl, dtype=torch.float32).pin_memory()

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    def timed(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(steps):
            fn(i)
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item())

    def resident(i):
        step(dev_in[i % NSETS], dev_g[i % NSETS])

    # End to end the way a training loop with a prefetching loader runs it: the pinned-host -> device copy of step
    # i+1 is issued on a copy stream while step i computes (two staging buffers); every step's input still crosses
    # PCIe inside the timed region and every step's statistic is read back to the host inside it.  The read of step i
    # is waited for after step i+1 has been enqueued (the usual one-step lag of a logged loss), so the host stays one
    # step ahead of the device; the last step's read is waited for before the region ends.
    copy_stream = torch.cuda.Stream(device=dev)
    stage = [stage_buf, torch.empty_like(stage_buf)]
    ev_copied = [torch.cuda.Event(), torch.cuda.Event()]
    ev_consumed = [torch.cuda.Event(), torch.cuda.Event()]
    e2e_state = {"steps": 0}
    er_host2 = [er_host, torch.empty_like(er_host).pin_memory()]
    ev_read = [torch.cuda.Event(), torch.cuda.Event()]

    def issue_copy(i):
        s_ = i % 2
        with torch.cuda.stream(copy_stream):
            copy_stream.wait_event(ev_consumed[s_])                     # the step that last read this buffer is done
            stage[s_].copy_(host_in[i % NSETS], non_blocking=True)      # H2D of step i's inputs
            ev_copied[s_].record(copy_stream)

    def end_to_end(i):
        main = torch.cuda.current_stream()
        if i == 0:
            issue_copy(0)
        if i + 1 < e2e_state["steps"]:
            issue_copy(i + 1)
        main.wait_event(ev_copied[i % 2])
        _, er, _, _ = step(stage[i % 2], dev_g[i % NSETS])
        ev_consumed[i % 2].record(main)
        er_host2[i % 2].copy_(er, non_blocking=True)                    # D2H of the step's statistic
        ev_read[i % 2].record(main)
        if i > 0:
            ev_read[(i - 1) % 2].synchronize()                          # step i-1's statistic is on the host now
            e2e_state["seen"] = float(er_host2[(i - 1) % 2][0])
        if i + 1 == e2e_state["steps"]:
            ev_read[i % 2].synchronize()                                # the last step: no successor to hide behind
            e2e_state["seen"] = float(er_host2[i % 2][0])

    for i in range(max(args.warmup, 3)):
        resident(i)
    # ---- timed region: K steps, no per-stage events (they cost ~4 % of a step that makes ~1000 launches)
    _lib.launch_count(reset=True)
    clocks = ClockSampler(local) if rank == 0 else None
    if clocks:
        clocks.start()
    ms = timed(resident, args.steps)
    launches = _lib.launch_count(reset=True)
    clk = clocks.stop() if clocks else None
    # ---- the same K steps again with CUDA events around every stage on the launching stream -> stage table, roofline.
    # In the timed region above the Jacobi iterations of the two batch chunks and the V update run on separate streams
    # and share the SMs; events around a stage would then time the co-scheduled neighbours as well.  This pass turns
    # the side streams off (one chunk, V update on the main stream): kernels are serialised, a stage's time is its
    # kernels' own duration -- what the roofline fraction is about.  Options the user set with --opt are restored.
    user_opts = dict(kv.split("=", 1) for kv in args.opt)
    _lib.set_option("jacobi_chunks", 1)
    _lib.set_option("jacobi_overlap_v", 0)
    resident(0)
    _lib.profile_enable(True)
    _lib.profile_read(reset=True)
    _lib.panel_tiles(reset=True)
    ms_prof = timed(resident, args.steps)
    prof = _lib.profile_read(reset=True)
    tiles = _lib.panel_tiles(reset=True)
    _lib.profile_enable(False)
    _lib.set_option("jacobi_chunks", float(user_opts.get("jacobi_chunks", 2)))
    _lib.set_option("jacobi_overlap_v", float(user_opts.get("jacobi_overlap_v", 1)))
    resident(0)
    _lib.launch_count(reset=True)
    sweeps = step.sweeps.abs().float().mean().item()
    not_converged = int((step.sweeps < 0).sum().item())
    er_mean = step.er.mean().item()
    # ---- end to end
    e2e_state["steps"] = 2
    for i in range(2):
        end_to_end(i)
    e2e_state["steps"] = args.steps
    ms_e2e = timed(end_to_end, args.steps)

    # the gradient all-reduce alone (N > 1): K back-to-back reductions of the flat fp32 bucket, max over ranks
    ar = None
    if world > 1:
        ar_ms = timed(lambda i: step.bucket.allreduce(average=True), args.steps) / args.steps
        ar = {"ms": ar_ms, "bytes": int(step.bucket.flat.numel() * 4),
              "note": "flat fp32 parameter-gradient bucket: copy in + NCCL all-reduce + copy out, timed alone; inside the "
                      "step it runs on a side stream beside the effective-rank backward"}
    value = world * Bl * args.steps / (ms * 1e-3)
    e2e = world * Bl * args.steps / (ms_e2e * 1e-3)
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    stages, pk = stage_table(prof, args.steps, world, tiles, Bl)
    isolated = None
    if world == 1 and not args.global_batch and not args.no_extras:
        gst = [torch.randn(Bl, T, 2, C, generator=gg, device=dev).to(dtype) for _ in range(NSETS)]
        isolated = isolated_stage_numbers(dev_in, gst, pk)
        del gst
    largest = max(stages.items(), key=lambda kv: kv[1]["ms_per_step"])[0]
    # the roofline object is for the dominant kernel that HAS a roofline (HBM- or tensor-bound); the inner Jacobi
    # solver is bound by instruction issue and is reported under `stages` only
    dname, d = max(((k, v) for k, v in stages.items() if v.get("frac") is not None), key=lambda kv: kv[1]["ms_per_step"])
    traffic = None
    tp = os.path.join(ROOT, "profiles", "r02_traffic.json")
    if os.path.exists(tp):
        with open(tp) as f:
            traffic = json.load(f).get(dname)
    roofline = {"kernel": dname, "bound": d.get("bound", "hbm"), "achieved": d.get("achieved"),
                "peak": d.get("peak"), "unit": d.get("unit", "GB/s"), "frac": d.get("frac"),
                "peak_source": pk["source"], "traffic": traffic,
                "bytes_basis": "kernel traffic: tiles the kernel actually processed x 64 KB (32 KB read + 32 KB written "
                               "of fp32 working matrix) / the stage's total time, not an algorithmic lower bound"
                               if dname.startswith("jacobi") else "algorithmic bytes / flops of the stage (DESIGN.md section 4)",
                "share_of_step": d["ms_per_step"] / (ms_prof / args.steps), "largest_stage_by_time": largest}
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
        "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "strong" if args.global_batch else "weak",
        "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
        "config": {"workload": ("DARai RGB+depth fuser training step fwd+bwd bf16 B=64 T=512 C=512 per GPU (BASELINE.json "
                                "configs[1])" if not args.global_batch else
                                f"NTU RGB+D-shaped fuser training step, global batch {args.global_batch} sharded over the ranks "
                                "(BASELINE.json configs[2]; T=512, C=512 bf16 -- the config names no T/C)") +
                               ": erank(rgb, depth) fwd + CMFuser fwd (score/bottom-k/exchange + Block + LN + token mean) + "
                               "CMFuser bwd (input and parameter gradients) + erank bwd + parameter-gradient all-reduce",
                   "B_per_gpu": Bl, "T": T, "C": C, "global_batch": Bl * world, "parallelism": f"dp{world}",
                   "dropout": "p=0 on both arms",
                   "l2": f"{NSETS} rotating input+grad sets ({NSETS * 3 * Bl * T * C * 2 / 1e6:.0f} MB) > 126 MB L2",
                   "jacobi_sweeps_mean": sweeps, "jacobi_not_converged": not_converged, "erank_mean": er_mean, "gram": args.gram},
        "e2e": {"value": e2e, "unit": UNIT, "ms_per_step": ms_e2e / args.steps,
                "h2d_bytes_per_step": int(2 * Bl * T * C * 2), "d2h_bytes_per_step": int(2 * Bl * 4)},
        "gpu_launches": int(launches),
        "clocks": clk,
        "roofline": roofline,
        "ms_per_step_with_stage_events": ms_prof / args.steps,
        "stage_timing": "separate pass with the side streams off (jacobi_chunks=1, jacobi_overlap_v=0) and CUDA events around "
                        "every stage: kernels serialised on one stream, a stage's time is its own kernels' duration; the timed "
                        "region (`value`) runs two batch chunks and the V update on concurrent streams",
        "stages": stages,
    }
    if ar:
        line["grad_allreduce"] = ar
    if world == 1 and not args.global_batch and not args.no_extras:
        line["e2e_c_abi"] = c_abi_e2e(host_in, Bl, dtype, args.steps)
    if isolated:
        line["stages_isolated"] = isolated
    if world == 1 and not args.global_batch and not args.no_extras:
        line["full_fuser_fwd_bwd"] = full_fuser_numbers(dev, dtype)
        if not args.no_yardstick:
            line["gpu_library_yardstick"] = library_yardstick(dev_in[0], dev)
    if world == 1 and not args.no_cpu_baseline and not args.no_extras:
        val, cms, kind = run_cpu(4, 1, args.cpu_sample)
        line["config"]["cpu_clips_per_step"] = args.cpu_sample
        line["cpu_baseline"] = {"value": val, "unit": UNIT, "cores": torch.get_num_threads(), "kind": kind,
                                "sample": CPU_SAMPLE_TEXT.format(n=args.cpu_sample,
                                                                 kind="oracle/_ref: unmodified reference module"
                                                                 if kind == "reference" else "torch-CPU port") +
                                          " x 4 steps (1 warm-up)",
                                "erank_part": "restatement (the reference has no effective-rank code, SURVEY.md F1)",
                                "erank_fwd_gram_eigvalsh_ms_per_clip": cpu_gram_route_ms(min(args.cpu_sample, 8))}
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    a = parse()
    if a.impl == "reference":
        main_reference(a)
    else:
        main_ours(a)
