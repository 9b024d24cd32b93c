"""CPU interpreter of a codenet_b200 Plan (numpy).  TEST TOOL: it executes the *descriptors* the engine hands to
the CUDA kernels (chunk tables, physical weight order, virtual upsampling, pad values) with plain integer
arithmetic, so the host-side compiler can be checked against the oracle without a GPU."""
import numpy as np

F = np.float64


def _rq(acc, M, B, lo):
    t = acc.astype(F) * M
    t = t + B
    return np.clip(np.rint(t), lo, 127).astype(np.int64)


def run_plan(plan, images):
    return _run(plan, images, {})


def _run(plan, images, T):
    B = images.shape[0] if images is not None else next(iter(T.values())).shape[0]
    heads = None
    for op in plan.ops:
        a = op.a
        if op.kind == "stem":
            x = images.astype(F)
            st = a["stride"]
            Ho, Wo = (a["H"] - 1) // st + 1, (a["W"] - 1) // st + 1
            p = np.zeros((B, 3, a["H"] + 2, a["W"] + 2), F); p[:, :, 1:-1, 1:-1] = x
            acc = np.zeros((B, Ho, Wo, a["C"]), F)
            w = a["wq"].astype(F)
            for c in range(3):
                for i in range(3):
                    for j in range(3):
                        acc += p[:, c, i:i + st * Ho:st, j:j + st * Wo:st][..., None] * w[:, c, i, j]
            q = _rq_f(acc, a["M"], a["B"], a["lo"])
            if a["pool"]:
                pp = np.full((B, Ho + 2, Wo + 2, a["C"]), -128, np.int64); pp[:, 1:-1, 1:-1] = q
                Hp, Wp = (Ho - 1) // 2 + 1, (Wo - 1) // 2 + 1
                q = np.full((B, Hp, Wp, a["C"]), -128, np.int64)
                for i in range(3):
                    for j in range(3):
                        q = np.maximum(q, pp[:, i:i + 2 * Hp:2, j:j + 2 * Wp:2])
            t = plan.tensors[a["out_t"]]
            buf = np.zeros((B, t.H, t.W, t.pitch), np.int64); buf[..., :a["C"]] = q
            T[t.id] = buf
        elif op.kind in ("dw", "deform"):
            tin, tout = plan.tensors[a["in_t"]], plan.tensors[a["out_t"]]
            x = T[tin.id]
            if a["in_shift"]:
                x = np.repeat(np.repeat(x, 2, axis=1), 2, axis=2)
            H, W = x.shape[1:3]
            C = a["C"]
            x = x[..., :C]
            w = a["wq"].astype(np.int64)
            zx = a["zx"]
            if op.kind == "dw":
                st = a["stride"]
                Ho, Wo = (H - 1) // st + 1, (W - 1) // st + 1
                p = np.full((B, H + 2, W + 2, C), -zx, np.int64); p[:, 1:-1, 1:-1] = x
                acc = np.zeros((B, Ho, Wo, C), np.int64)
                for i in range(3):
                    for j in range(3):
                        acc += p[:, i:i + st * Ho:st, j:j + st * Wo:st] * w[:, i * 3 + j]
                acc += zx * w.sum(1)
                q = _rq(acc, a["M"], a["B"], a["lo"])
            else:
                ws = a["ws"].astype(np.int64)
                acc_s = (x * ws).sum(-1) + zx * ws.sum()
                u = acc_s.astype(F) * F(a["Ms"]) + F(a["bs"])
                u = np.clip(u, F(-a["bound"] + 1), F(a["bound"]))
                qs = np.rint(F(a["ss"]) * u - F(a["zs"]))
                s = (qs + F(a["zs"])) / F(a["ss"])
                assert a["mode"] == 0, "plan_sim only interprets integer offsets"
                s = np.rint(s).astype(np.int64)
                hh, ww = np.arange(H).reshape(1, H, 1), np.arange(W).reshape(1, 1, W)
                bi = np.arange(B).reshape(B, 1, 1)
                acc = np.zeros((B, H, W, C), np.int64)
                for i in range(3):
                    for j in range(3):
                        hi, wi = hh + (i - 1) * s, ww + (j - 1) * s
                        ok = (hi >= 0) & (hi < H) & (wi >= 0) & (wi < W)
                        v = x[bi, np.clip(hi, 0, H - 1), np.clip(wi, 0, W - 1)]
                        v = np.where(ok[..., None], v, -zx)
                        acc += v * w[:, i * 3 + j]
                acc += zx * w.sum(1)
                q = _rq(acc, a["M"], a["B"], a["lo"])
            buf = np.zeros((B, tout.H, tout.W, tout.pitch), np.int64); buf[..., :C] = q
            T[tout.id] = buf
        elif op.kind == "pw":
            tin = plan.tensors[a["in_t"]]
            x = T[tin.id]
            Bn, H, W, _ = x.shape
            K, k_off = a["K"], a["k_off"]
            A = x[..., k_off:k_off + K].reshape(-1, K).astype(F)
            w = a["wq"].astype(F)
            acc = np.rint(A @ w.T).astype(np.int64) + a["zx"] * a["wq"].astype(np.int64).sum(1)
            if a["n_f32"]:
                n = a["n_f32"]
                y = acc[:, :n].astype(F) * a["Mf"] + a["bf"]
                heads = y.reshape(Bn, H, W, n).transpose(0, 3, 1, 2)
                continue
            q = _rq(acc, a["M"], a["B"], a["lo"])
            tout = plan.tensors[a["out_t"]]
            out = np.zeros((Bn * H * W, tout.pitch), np.int64)
            pas = T[a["pass_t"]].reshape(Bn * H * W, -1) if a["pass_t"] >= 0 else None
            for col, cnt, poff, dst in a["chunks"]:
                if cnt == 0:
                    continue
                if poff < 0:
                    out[:, dst:dst + cnt] = q[:, col:col + cnt]
                else:
                    out[:, dst:dst + 2 * cnt:2] = pas[:, poff:poff + cnt]
                    out[:, dst + 1:dst + 2 * cnt:2] = q[:, col:col + cnt]
            T[tout.id] = out.reshape(Bn, H, W, tout.pitch)
    return T, heads


def _rq_f(accf, M, B, lo):
    t = accf * M
    t = t + B
    return np.clip(np.rint(t), lo, 127).astype(np.int64)


def logical(plan, T, label):
    t = plan.tensors[plan.taps[label]]
    return T[t.id][..., t.phys(np.arange(t.C))].transpose(0, 3, 1, 2)


def run_plan_seeded(plan, seed):
    """Interpret a plan whose first tensors are given (no stem): returns (tensor dict, heads)."""
    import types
    ops = [o for o in plan.ops if o.kind != "stem"]
    shadow = types.SimpleNamespace(ops=ops, tensors=plan.tensors, taps=getattr(plan, "taps", {}))
    return _run(shadow, None, dict(seed))
