import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import numpy as np, torch, ctypes as C
from codenet_b200 import _lib
L = _lib.load()
ptr = lambda t: C.c_void_p(t.data_ptr()) if t is not None else None
st = lambda: C.c_void_p(torch.cuda.current_stream().cuda_stream)
Cc, Co, ppi, B = 16, 16, 256, 1
x = (np.arange(Cc)[None, :, None] * 1000 + np.arange(ppi)[None, None, :]).astype(np.float32)
w = np.zeros((Co, Cc), np.float32); w[np.arange(Co), np.arange(Co)] = 1; w[3, 5] = 0.333
tx, tw = torch.from_numpy(x).cuda(), torch.from_numpy(w).cuda()
n = int(L.cdn_pw_tf32x3_packed_floats(Co, Cc))
hi, lo = torch.empty(n, device="cuda"), torch.empty(n, device="cuda")
_lib.check(L.cdn_pw_tf32x3_pack(ptr(tw), Co, Cc, ptr(hi), ptr(lo), st()))
print("packed n", n, "hi nz", int((hi != 0).sum()), "hi diag", hi.cpu().numpy().reshape(-1, 16)[:4, :6].tolist(), "lo[3,5]", float(lo.view(-1, 16)[3, 5]))
L.cdn_pw_tf32x3_debug.argtypes = [C.c_void_p]; L.cdn_pw_tf32x3_debug.restype = None
for variant in (0, 2):
    L.cdn_set_debug_flags(variant << 8)
    dbg = torch.full((49152 // 4 + 128 * 16,), -5.0, device="cuda")
    L.cdn_pw_tf32x3_debug(ptr(dbg))
    out = torch.full((B, Co, ppi), -77.0, device="cuda")
    _lib.check(L.cdn_pw_slice_tf32x3(ptr(tx), Cc, 0, Cc, ptr(hi), ptr(lo), None, ptr(out), Co, 0, 1, Co, 0, B, ppi, st()))
    torch.cuda.synchronize()
    acc = dbg.cpu().numpy()[12288:].reshape(128, 16)
    print("variant", variant, "acc nz", int((acc != 0).sum()), "lanes 0,1,40:", acc[0, :8].tolist(), acc[1, :8].tolist(), acc[40, :8].tolist(), acc[40, 8:].tolist())
