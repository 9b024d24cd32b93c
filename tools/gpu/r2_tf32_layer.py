"""Time cdn_pw_slice_tf32x3 on the float model's layer shapes (batch 128, 512x512 input geometry); CUDA events, L2-sized rotation."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import numpy as np, torch, ctypes as C
from codenet_b200 import _lib
L = _lib.load()
ptr = lambda t: C.c_void_p(t.data_ptr()) if t is not None else None
st = lambda: C.c_void_p(torch.cuda.current_stream().cuda_stream)
only = sys.argv[1] if len(sys.argv) > 1 else None
B = 128
shapes = [("l1.pw", 122, 122, 4096), ("l2.pw", 244, 244, 1024), ("l3.pw", 488, 488, 256), ("layer4", 976, 2153, 256), ("up0.ch", 2153, 256, 256),
          ("l2.0.pw1", 244, 244, 4096), ("hm.pw1", 64, 64, 16384), ("hm.out", 64, 80, 16384), ("heads.pw1", 64, 192, 16384), ("l1.0.pw1", 24, 122, 16384)]
for name, Cc, Co, ppi in shapes:
    if only and name != only: continue
    x = torch.randn(B, Cc, ppi, device="cuda"); w = torch.randn(Co, Cc, device="cuda") / Cc ** 0.5
    n = int(L.cdn_pw_tf32x3_packed_floats(Co, Cc))
    packed = torch.empty(n, device="cuda")
    _lib.check(L.cdn_pw_tf32x3_pack(ptr(w), Co, Cc, ptr(packed), st()))
    out = torch.empty(B, Co, ppi, device="cuda")
    for flags in (0, 1 << 31):                   # bit 31: two stages per accumulation chunk (A/B)
        L.cdn_set_debug_flags(flags)
        f = lambda: _lib.check(L.cdn_pw_slice_tf32x3(ptr(x), Cc, 0, Cc, ptr(packed), None, ptr(out), Co, 0, 1, Co, 1, B, ppi, st()))
        for _ in range(3): f()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10): f()
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 10
        fl = 2.0 * B * ppi * Cc * Co
        by = 4.0 * B * ppi * (Cc + Co)
        print("%-9s flags=%x chunk=%d  %.3f ms  %.1f TFLOP/s (x3 MMA: %.0f)  %.0f GB/s" % (name, flags >> 8 & 15, 2 if flags >> 31 else 1, ms, fl / ms / 1e9, 3 * fl / ms / 1e9, by / ms / 1e6))
    if not only:
        L.cdn_set_debug_flags(0)
        f2 = lambda: _lib.check(L.cdn_pw_slice_f32(ptr(x), Cc, 0, Cc, ptr(w), None, ptr(out), Co, 0, 1, Co, 1, B, ppi, st()))
        for _ in range(2): f2()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5): f2()
        e1.record(); torch.cuda.synchronize()
        print("%-9s SIMT     %.3f ms" % (name, e0.elapsed_time(e1) / 5))
