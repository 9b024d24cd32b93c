"""Diagnose cdn_pw_slice_tf32x3 with structured inputs: identity weights (out = x reveals pixel / channel permutations)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch, ctypes as C
from codenet_b200 import _lib
L = _lib.load()
L.cdn_set_debug_flags(int(os.environ.get('PT_FLAGS', '0')))
ptr = lambda t: C.c_void_p(t.data_ptr()) if t is not None else None
st = lambda: C.c_void_p(torch.cuda.current_stream().cuda_stream)
def run(x, w, Co, Cc, ppi, B):
    tx, tw = torch.from_numpy(x).cuda(), torch.from_numpy(w).cuda()
    n = int(L.cdn_pw_tf32x3_packed_floats(Co, Cc))
    packed = torch.empty(n, device="cuda")
    _lib.check(L.cdn_pw_tf32x3_pack(ptr(tw), Co, Cc, ptr(packed), st()))
    out = torch.full((B, Co, ppi), -77.0, device="cuda")
    _lib.check(L.cdn_pw_slice_tf32x3(ptr(tx), Cc, 0, Cc, ptr(packed), None, ptr(out), Co, 0, 1, Co, 0, B, ppi, st()))
    torch.cuda.synchronize()
    return out.cpu().numpy(), None, None
for Cc in (16, 32):
    Co, ppi, B = 16, 256, 1
    x = (np.arange(Cc)[None, :, None] * 1000 + np.arange(ppi)[None, None, :]).astype(np.float32) * np.ones((B, 1, 1), np.float32)
    w = np.zeros((Co, Cc), np.float32); w[np.arange(Co), np.arange(Co)] = 1
    got, hi, lo = run(x, w, Co, Cc, ppi, B)
    print("C=%d identity: max err" % Cc, np.abs(got - x[:, :Co]).max())
    if np.abs(got - x[:, :Co]).max() > 0:
        for co in (0, 1, 5):
            print(" co", co, got[0, co, :40].tolist())
            print(" co", co, got[0, co, 120:140].tolist())
# random small
rng = np.random.default_rng(0)
for (Cc, Co, ppi, B) in ((16, 16, 256, 1), (32, 16, 256, 1), (48, 128, 256, 2), (122, 122, 512, 2), (244, 244, 256, 2)):
    x = rng.normal(0, 1, (B, Cc, ppi)).astype(np.float32)
    w = rng.normal(0, 1, (Co, Cc)).astype(np.float32)
    got, hi, lo = run(x, w, Co, Cc, ppi, B)
    ref = np.einsum("oc,bcp->bop", w.astype(np.float64), x.astype(np.float64))
    e = np.abs(got - ref)
    print("rand C=%d Co=%d ppi=%d B=%d: max err %.3g  rel l2 %.3g; worst at" % (Cc, Co, ppi, B, e.max(), np.sqrt((e**2).sum()/(ref**2).sum())), np.unravel_index(e.argmax(), e.shape))
    bad = e > 1e-3
    if bad.any():
        print("  bad fraction %.4f; bad channels %s; bad px (first 20) %s" % (bad.mean(), np.unique(np.nonzero(bad)[1])[:20], np.unique(np.nonzero(bad)[2])[:20]))
