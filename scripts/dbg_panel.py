import sys, numpy as np, torch
sys.path.insert(0, '.')
from r3d_b200 import _lib
from r3d_b200.ops import _p, _stream
L = _lib.lib()
def rr_pair(m, r, t):
    if m == 2: return 0, 1
    if t == 0: x, y = r, m - 1
    else: x, y = (r + t) % (m - 1), (r - t + (m - 1)) % (m - 1)
    return min(x, y), max(x, y)
def run(B, npad, rnd, seed=0, simpleQ=False, debug=0, pattern=False):
    _lib.set_option("panel_debug", debug)
    _lib.set_option("panel_sym", SYM)
    rng = np.random.default_rng(seed)
    nb, nt = npad // 32, npad // 64
    if pattern:
        G = (np.arange(npad)[:, None] + np.arange(npad)[None, :] / 1024.0).astype(np.float32)[None].repeat(B, 0)
        V = G.copy() + 1000
    else:
        A = rng.standard_normal((B, npad, npad)).astype(np.float32)
        G = (A + A.transpose(0, 2, 1)) / 2
        V = rng.standard_normal((B, npad, npad)).astype(np.float32)
    Q = np.tile(np.eye(64, dtype=np.float32), (B, nt, 1, 1)) if simpleQ else rng.standard_normal((B, nt, 64, 64)).astype(np.float32) / 8
    dev = torch.device('cuda')
    def to_blocked(M):   # (B, np, np) row-major -> [B][np/32][np][32]
        return np.ascontiguousarray(M.reshape(B, npad, npad // 32, 32).transpose(0, 2, 1, 3))
    def from_blocked(t):
        return t.cpu().numpy().reshape(B, npad // 32, npad, 32).transpose(0, 2, 1, 3).reshape(B, npad, npad)
    G_rm, V_rm = G, V
    G, V = to_blocked(G), to_blocked(V)
    Gd, Vd, Qd = torch.from_numpy(G).to(dev), torch.from_numpy(V).to(dev), torch.from_numpy(np.ascontiguousarray(Q.transpose(0, 1, 3, 2))).to(dev)
    Hd = torch.full_like(Gd, -7.0)
    G, V = G_rm, V_rm
    scratch = torch.zeros(B * 32 + B * nt, dtype=torch.int32, device=dev)
    _lib.check(L.r3d_debug_panel_round(_p(Gd), _p(Hd), _p(Vd), _p(Qd), B, npad, rnd, _p(scratch), _stream()))
    torch.cuda.synchronize()
    Qf = np.zeros((B, npad, npad), np.float64)
    for b in range(B):
        for t in range(nt):
            I, J = rr_pair(nb, rnd, t)
            ix = np.concatenate([np.arange(I * 32, I * 32 + 32), np.arange(J * 32, J * 32 + 32)])
            Qf[b][np.ix_(ix, ix)] = Q[b, t]
    G64, V64 = G.astype(np.float64), V.astype(np.float64)
    Href = (G64 @ Qf).transpose(0, 2, 1)
    Gref = Qf.transpose(0, 2, 1) @ G64 @ Qf
    Vref = V64 @ Qf
    print(f"--- B={B} np={npad} round={rnd} simpleQ={simpleQ} debug={debug} pattern={pattern}")
    for name, got, ref in (('H', Hd, Href), ('G', Gd, Gref), ('V', Vd, Vref)):
        g = from_blocked(got)
        err = np.abs(g - ref).max() / np.abs(ref).max()
        print(f"  {name}: rel err {err:.3e} got absmax {np.abs(g).max():.3f} ref absmax {np.abs(ref).max():.3f}")
    if pattern:
        print("  V out [0,:4,:6]:\n", Vd.cpu().numpy()[0, :4, :6])
        print("  V out [0,64:66,32:38]:\n", Vd.cpu().numpy()[0, 64:66, 32:38])
SYM = 1
which = sys.argv[1] if len(sys.argv) > 1 else "a"
if which == "sym":      # one-pass symmetric G update (jacobi_sym.cu): H is not written, G and V are
    for (B, npad, rnd) in ((1, 128, 0), (1, 256, 1), (2, 256, 2), (3, 512, 4), (5, 384, 3)):
        run(B, npad, rnd)
    sys.exit(0)
SYM = 0
if which == "a":
    run(1, 128, 0, simpleQ=True, debug=1 | 4, pattern=True)  # MMA hi*hi only on raw data
    run(1, 128, 0, simpleQ=False, debug=1)
    run(1, 128, 0, simpleQ=False, debug=0)
    run(2, 256, 1, simpleQ=False, debug=0)
elif which == "b":
    run(3, 512, 4)
elif which == "c":
    cap = int(sys.argv[2])
    _lib.set_option("panel_grid_cap", cap)
    run(1, 128, 0, simpleQ=False, debug=0)
