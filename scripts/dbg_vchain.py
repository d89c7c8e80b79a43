"""Chained V update (jacobi_schedule = 2) against float64: V <- V Q1 Q2 Q3 for three XOR rounds."""
import sys, numpy as np, torch
sys.path.insert(0, '.')
from r3d_b200 import _lib
from r3d_b200.ops import _p, _stream
L = _lib.lib()


def xor_pair(mask, t):
    hb = mask.bit_length() - 1
    a = ((t >> hb) << (hb + 1)) | (t & ((1 << hb) - 1))
    return a, a ^ mask


def run(B, npad, ga, gb, seed=0, identity=False):
    rng = np.random.default_rng(seed)
    nb, nt = npad // 32, npad // 64
    V = rng.standard_normal((B, npad, npad)).astype(np.float32)
    Q = rng.standard_normal((3, B, nt, 64, 64)).astype(np.float32) / 8
    if identity:
        Q[:] = np.eye(64, dtype=np.float32)
    dev = torch.device('cuda')
    Vb = np.ascontiguousarray(V.reshape(B, npad, npad // 32, 32).transpose(0, 2, 1, 3))
    Vd = torch.from_numpy(Vb).to(dev)
    Qd = torch.from_numpy(np.ascontiguousarray(Q.transpose(0, 1, 2, 4, 3))).to(dev)      # Q^T per task
    scratch = torch.zeros(B * 32 + 3 * B * nt, dtype=torch.int32, device=dev)
    _lib.check(L.r3d_debug_vchain(_p(Vd), _p(Qd), B, npad, ga, gb, _p(scratch), _stream()))
    torch.cuda.synchronize()
    ref = V.astype(np.float64)
    for k, mask in enumerate((ga, gb, ga ^ gb)):
        Qf = np.zeros((B, npad, npad))
        for b in range(B):
            for t in range(nt):
                I, J = xor_pair(mask, t)
                ix = np.concatenate([np.arange(I * 32, I * 32 + 32), np.arange(J * 32, J * 32 + 32)])
                Qf[b][np.ix_(ix, ix)] = Q[k, b, t]
        ref = ref @ Qf
    got = Vd.cpu().numpy().reshape(B, npad // 32, npad, 32).transpose(0, 2, 1, 3).reshape(B, npad, npad)
    err = np.abs(got - ref).max() / np.abs(ref).max()
    print(f"B={B} np={npad} masks=({ga},{gb},{ga ^ gb}) identity={identity}: rel err {err:.3e}")
    return err


if __name__ == "__main__":
    worst = 0.0
    worst = max(worst, run(1, 256, 1, 2, identity=True))
    for (B, npad, ga, gb) in ((1, 256, 1, 2), (2, 256, 3, 5), (3, 512, 1, 6), (2, 512, 9, 14), (5, 512, 15, 4), (1, 1024, 21, 10)):
        worst = max(worst, run(B, npad, ga, gb))
    print("worst", worst)
    sys.exit(0 if worst < 1e-5 else 1)
