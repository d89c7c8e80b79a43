"""CPU-only tests (-m "not gpu"): the C-ABI library loads and exports every symbol the header declares,
the host-side module mirrors the reference interface, CPU tensors are refused, and the N > 1 host logic
agrees with the single-process oracle under a 2-rank gloo group."""
import os
import re
import socket

import numpy as np
import pytest
import torch

from conftest import ROOT, golden_files
from oracle import fuser_oracle as O


def _header_symbols():
    txt = open(os.path.join(ROOT, "include", "r3d_b200.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(r3d_[a-z0-9_]+)\s*\(", txt)))


def test_library_exports_every_header_symbol():
    import ctypes
    from r3d_b200 import _lib
    assert os.path.exists(_lib.LIB_PATH), "build the library first: python -m r3d_b200.csrc.build"
    h = ctypes.CDLL(_lib.LIB_PATH)
    syms = _header_symbols()
    assert len(syms) >= 25
    for s in syms:
        assert hasattr(h, s), f"{s} declared in include/r3d_b200.h but not exported"
    # and the ctypes binding covers the same set (no drift between header and host side)
    assert set(_lib.SIGNATURES) == set(syms), set(_lib.SIGNATURES) ^ set(syms)
    L = _lib.lib()
    assert L.r3d_abi_version() == 2
    assert L.r3d_score_workspace_floats(2048, 512) > 0          # size queries need no GPU
    assert L.r3d_erank_workspace_bytes(2, 64, 128, 0) > 0
    assert L.r3d_profile_num_stages() > 10


def test_abi_argument_errors_without_gpu():
    from r3d_b200 import _lib
    L = _lib.lib()
    rc = L.r3d_bottomk(None, 1, 16, 4, None, None)
    assert rc != 0 and b"null" in L.r3d_last_error()
    rc = L.r3d_exchange_fwd(None, None, None, None, 0, None, None, 0, None, 4, 16, 0, None)
    assert rc != 0
    rc = L.r3d_set_option(b"no_such_option", 1.0)
    assert rc != 0 and b"unknown option" in L.r3d_last_error()


@pytest.mark.parametrize("variant", ["tokenfusion", "vary", "batchnorm", "safuser"])
def test_state_dict_names_match_reference(variant):
    import r3d_b200
    path = [p for p in golden_files(variant) if "C16" in p][0]
    z = np.load(path)
    ref_names = sorted(k[3:] for k in z.files if k.startswith("sd/"))
    f = r3d_b200.CMFuser(16, depth=1, num_heads=4, variant=variant)
    assert sorted(f.state_dict().keys()) == ref_names
    f.load_state_dict({k: torch.from_numpy(z["sd/" + k]) for k in ref_names}, strict=True)
    for k in ref_names:
        assert tuple(f.state_dict()[k].shape) == tuple(z["sd/" + k].shape)


def test_reference_constructor_and_api_surface():
    import inspect
    import r3d_b200
    sig = inspect.signature(r3d_b200.CMFuser.__init__)
    assert list(sig.parameters)[:6] == ["self", "dim", "depth", "num_heads", "mlp_ratio", "qkv_bias"]
    assert sig.parameters["depth"].default == 1 and sig.parameters["num_heads"].default == 4
    f = r3d_b200.CMFuser(32)
    assert f.k_for(512) == 128 and r3d_b200.CMFuser(32, variant="batchnorm").k_for(512) == 51
    assert r3d_b200.CMFuser(8, variant="batchnorm").k_for(8) == 0          # int(C * 0.1) == 0 is legal
    m = r3d_b200.CMFuser.generate_cross_attention_mask(2)
    assert torch.isinf(m[0, 0]) and m[0, 1] == 0
    with pytest.raises(ValueError):
        r3d_b200.CMFuser(8, variant="nope")


def test_cpu_tensors_are_refused():
    import r3d_b200
    f = r3d_b200.CMFuser(16)
    x = torch.randn(2, 3, 16)
    with pytest.raises(r3d_b200.R3DError):
        f.token_fusion(x, x, "test")
    with pytest.raises(r3d_b200.R3DError):
        r3d_b200.ops.channel_score(x, x)
    with pytest.raises(r3d_b200.R3DError):
        r3d_b200.ops.erank(x)
    with pytest.raises(r3d_b200.R3DError):
        r3d_b200.ops.bottomk(torch.randn(4, 8), 2)


def test_block_closed_form_equals_masked_attention():
    """The CUDA path's Block evaluates the 2x2 masked attention in closed form; check it on CPU
    against the oracle's unsimplified arithmetic."""
    import r3d_b200
    torch.manual_seed(0)
    C, H = 32, 4
    blk = r3d_b200.Block(C, H)
    x = torch.randn(7, 2, C)
    y, _ = blk(x)
    p = {"b." + k: v.detach().numpy() for k, v in blk.state_dict().items()}
    ref, attn = O.block_forward(x.numpy().astype(np.float64), p, "b.", H)
    np.testing.assert_allclose(y.detach().numpy(), ref, rtol=1e-4, atol=1e-5)
    assert set(np.unique(attn)) == {0.0, 1.0}


def test_shard_bounds_cover_batch():
    from r3d_b200.dist import shard_bounds
    for total in (1, 7, 64, 513):
        for world in (1, 2, 3, 8):
            spans = [shard_bounds(total, world, r) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == total
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            assert max(h - l for l, h in spans) - min(h - l for l, h in spans) <= 1


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _gloo_worker(rank, world, port, B, T, C, k, q):
    import torch.distributed as dist
    from r3d_b200 import dist as D
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    g = torch.Generator().manual_seed(1234)
    c = torch.arange(C, dtype=torch.float32)
    rgb = torch.relu(torch.randn(B, T, C, generator=g)) * (1 + c / C)
    dep = torch.relu(torch.randn(B, T, C, generator=g)) * (2 - c / C)
    lo, hi = D.shard_bounds(B, world, rank)
    # local partial statistics (on a GPU these come from r3d_channel_score_partial / r3d_erank_fwd)
    sums = torch.stack([rgb[lo:hi].abs().sum(dim=(0, 1)), dep[lo:hi].abs().sum(dim=(0, 1))])
    er_local = torch.arange(lo, hi, dtype=torch.float32).sum()          # stand-in per-sample statistic
    packed = D.allreduce_statistics(D.pack_statistics(sums, er_local, (hi - lo) * T))
    score, er_mean = D.unpack_statistics(packed, B)
    idx = D.global_bottomk_indices(score, k)
    # gradient bucket: rank-dependent grads, one unused parameter
    p1 = torch.nn.Parameter(torch.zeros(5))
    p2 = torch.nn.Parameter(torch.zeros(3))
    p3 = torch.nn.Parameter(torch.zeros(2))
    p1.grad = torch.full((5,), float(rank + 1))
    p2.grad = None
    p3.grad = torch.full((2,), 10.0 * (rank + 1))
    D.GradBucket([p1, p2, p3]).allreduce(average=True)
    q.put((rank, score.numpy(), float(er_mean), idx.numpy(), p1.grad.numpy(), p2.grad.numpy(), p3.grad.numpy()))
    dist.destroy_process_group()


def test_two_rank_gloo_matches_single_process_oracle():
    import torch.multiprocessing as mp
    B, T, C = 6, 9, 64
    k = C // 4
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_gloo_worker, args=(r, 2, port, B, T, C, k, q)) for r in range(2)]
    for p in procs:
        p.start()
    outs = sorted([q.get(timeout=120) for _ in procs], key=lambda t: t[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    g = torch.Generator().manual_seed(1234)
    c = torch.arange(C, dtype=torch.float32)
    rgb = (torch.relu(torch.randn(B, T, C, generator=g)) * (1 + c / C)).numpy()
    dep = (torch.relu(torch.randn(B, T, C, generator=g)) * (2 - c / C)).numpy()
    ref_score = np.stack([O.channel_score(rgb), O.channel_score(dep)])
    ref_idx = np.stack([O.bottomk(ref_score[0], k), O.bottomk(ref_score[1], k)])
    for rank, score, er_mean, idx, g1, g2, g3 in outs:
        np.testing.assert_allclose(score, ref_score, rtol=1e-5)            # global scope == concatenated batch
        np.testing.assert_array_equal(idx, ref_idx)                        # identical selection on every rank
        assert abs(er_mean - np.arange(B).mean()) < 1e-6
        np.testing.assert_allclose(g1, 1.5)
        np.testing.assert_allclose(g2, 0.0)
        np.testing.assert_allclose(g3, 15.0)
    np.testing.assert_array_equal(outs[0][3], outs[1][3])


# ------------------------------------------------------------------ bench.py contract (CPU-runnable parts)
def test_bench_reference_arm_prints_one_json_line():
    """`bench.py --impl reference` (the reference's CPU algorithm on the host cores) prints exactly one JSON line with
    the keys the driver reads."""
    import json, subprocess, sys
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1",
                          "--warmup", "0", "--cpu-sample", "1"], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [ln for ln in out.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "clips/s" and d["higher_is_better"] is True
    assert d["vs_baseline"] is None and d["value"] > 0 and d["steps"] == 1
    assert set(d["cpu_baseline"]) >= {"value", "unit", "cores", "kind", "sample"} and d["cpu_baseline"]["kind"] in ("port", "reference")
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0
    assert "workload" in d["config"]


def test_bench_ours_refuses_without_gpu():
    """The product arm has no CPU fallback: without a CUDA device it exits with an error instead of timing the oracle."""
    import subprocess, sys
    import torch
    if torch.cuda.is_available():
        pytest.skip("needs a box without a GPU")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "1", "--warmup", "1"],
                         capture_output=True, text=True, timeout=300, cwd=ROOT)
    assert out.returncode != 0
    assert "CUDA" in (out.stderr + out.stdout)


def test_header_is_plain_c_and_links(tmp_path):
    """The drop-in boundary is a C ABI: include/r3d_b200.h must compile as C99 (no C++ or torch types) and a C program
    must link against the shared library and call it (no compute: there is no GPU here)."""
    import shutil, subprocess
    if shutil.which("gcc") is None:
        pytest.skip("no gcc")
    src = tmp_path / "abi.c"
    src.write_text('#include "r3d_b200.h"\n#include <stdio.h>\n'
                   'int main(void) {\n'
                   '  if (r3d_abi_version() < 1) return 1;\n'
                   '  /* a NULL-pointer call must fail cleanly with a message, not crash */\n'
                   '  if (r3d_bottomk(NULL, 1, 8, 2, NULL, NULL) == 0) return 2;\n'
                   '  const char* e = r3d_last_error();\n'
                   '  if (!e || !e[0]) return 3;\n'
                   '  printf("%s\\n", e);\n  return 0;\n}\n')
    libdir = os.path.join(ROOT, "r3d_b200", "csrc")
    exe = tmp_path / "abi"
    subprocess.run(["gcc", "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", "-I", os.path.join(ROOT, "include"),
                    str(src), "-o", str(exe), "-L", libdir, "-lr3d_b200", f"-Wl,-rpath,{libdir}"], check=True)
    out = subprocess.run([str(exe)], capture_output=True, text=True)
    assert out.returncode == 0, (out.returncode, out.stdout, out.stderr)
    assert "null" in out.stdout.lower()


# ------------------------------------------------------------------ eigensolver schedule (host logic, no GPU)
def _round_plan(nb):
    from r3d_b200 import _lib
    out = np.zeros(2 * 512, np.int32)
    n = _lib.lib().r3d_debug_round_plan(nb, out.ctypes.data, 512)
    return [tuple(int(v) for v in out[2 * i:2 * i + 2]) for i in range(n)]


@pytest.mark.parametrize("nb", [8, 16, 32, 64, 128, 256])
def test_round_plan_covers_every_mask_exactly_once(nb):
    """The rounds of a sweep are the XOR matchings i <-> i ^ mask; a super-round {a, b, a ^ b} keeps three of them inside
    4-block cosets (csrc/erank_kernels.cu: spread_plan).  Every mask 1..nb-1 must occur exactly once per sweep -- then every
    block pair meets exactly once, as in the circle method."""
    plan = _round_plan(nb)
    masks = []
    for a, b in plan:
        assert 0 < a < nb and 0 <= b < nb
        masks += [a] if b == 0 else [a, b, a ^ b]
        if b:
            assert a != b and (a ^ b) not in (0, a, b)
    assert sorted(masks) == list(range(1, nb))
    k = nb.bit_length() - 1
    if k % 2 == 0:                       # GF(2^k) over GF(4): a full spread, no single rounds left over
        assert all(b != 0 for _, b in plan) and len(plan) == (nb - 1) // 3
    assert plan[0][1] != 0               # the first entry of a sweep carries the generic (within-block) round


@pytest.mark.parametrize("nb", [2, 4, 6, 12, 24])
def test_round_plan_falls_back_to_circle_method(nb):
    assert _round_plan(nb) == []


@pytest.mark.parametrize("nb", [8, 16, 32, 64])
def test_chain_plan_matches_the_xor_pairing(nb):
    """Tile bookkeeping of the chained V update (csrc/jacobi_tc.cu: chain_plan) against the pairing the inner solver uses
    (rr_pair with a negative round code): the cosets partition the blocks; in round k the accumulator holds the coset's
    blocks as [pair 0: I, J | pair 1: I, J] with J = I ^ mask_k, I the block with the mask's top bit clear, and the task
    index is I with that bit removed -- the row of the round's Q^T buffer the inner solver wrote for (I, J)."""
    from r3d_b200 import _lib
    L = _lib.lib()
    for a, b in _round_plan(nb):
        if b == 0:
            continue
        seen = []
        for g in range(nb // 4):
            o = np.zeros(22, np.int32)
            assert L.r3d_debug_chain_plan(a, b, g, o.ctypes.data) == 0
            blk = [int(v) for v in o[:4]]
            assert sorted(blk) == sorted({blk[0], blk[0] ^ a, blk[0] ^ b, blk[0] ^ a ^ b}) and len(set(blk)) == 4
            seen += blk
            for k, mask in enumerate((a, b, a ^ b)):
                sig = [int(v) for v in o[4 + 4 * k:8 + 4 * k]]
                assert sorted(sig) == [0, 1, 2, 3]
                hb = mask.bit_length() - 1
                for p in range(2):
                    I, J = blk[sig[2 * p]], blk[sig[2 * p + 1]]
                    assert J == I ^ mask and (I >> hb) & 1 == 0
                    task = ((I >> (hb + 1)) << hb) | (I & ((1 << hb) - 1))
                    assert int(o[16 + 2 * k + p]) == task
        assert sorted(seen) == list(range(nb))


def test_round_plan_drives_a_convergent_block_jacobi():
    """The sweep plan the library emits, used by a plain numpy block Jacobi (blocks of 32 columns, every block pair of a
    round diagonalised exactly with eigh, rotations accumulated): off-diagonal mass must vanish within a few sweeps and
    the spectrum / eigenvectors must equal numpy's -- i.e. the XOR / super-round ordering is a complete Jacobi ordering,
    and applying V <- V Q1 Q2 Q3 once per super-round (what the chained kernel does) equals three separate updates."""
    rng = np.random.default_rng(3)
    nb, jb = 8, 32
    n = nb * jb
    A = rng.standard_normal((n, 2 * n))
    G = A @ A.T
    plan = _round_plan(nb)
    V = np.eye(n)
    V_chained = np.eye(n)

    def round_q(G, mask):
        Q = np.eye(n)
        hb = mask.bit_length() - 1
        for t in range(nb // 2):
            i = ((t >> hb) << (hb + 1)) | (t & ((1 << hb) - 1))
            j = i ^ mask
            ix = np.r_[i * jb:(i + 1) * jb, j * jb:(j + 1) * jb]
            _, q = np.linalg.eigh(G[np.ix_(ix, ix)])
            Q[np.ix_(ix, ix)] = q
        return Q

    off0 = np.linalg.norm(G - np.diag(np.diag(G)))
    for sweep in range(6):
        for a, b in plan:
            qs = []
            for mask in ([a] if b == 0 else [a, b, a ^ b]):
                Q = round_q(G, mask)
                G = Q.T @ G @ Q
                V = V @ Q
                qs.append(Q)
            P = qs[0] if len(qs) == 1 else qs[0] @ qs[1] @ qs[2]
            V_chained = V_chained @ P                     # one pass per super-round
    off = np.linalg.norm(G - np.diag(np.diag(G)))
    assert off < 1e-10 * off0
    np.testing.assert_allclose(V, V_chained, atol=1e-12)
    lam = np.linalg.eigvalsh(A @ A.T)
    np.testing.assert_allclose(np.sort(np.diag(G)), lam, rtol=1e-10)
    np.testing.assert_allclose(V.T @ (A @ A.T) @ V, np.diag(np.diag(G)), atol=1e-8 * lam[-1])
