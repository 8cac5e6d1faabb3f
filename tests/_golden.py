"""Iterates the golden fixtures recorded from the reference (tests/golden/make_golden.py)."""
import os
import re

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))


def load(device):
    path = os.path.join(HERE, "golden", f"ref_{device}.npz")
    if not os.path.exists(path):
        return None
    return np.load(path)


def get(z, key):
    """returns (array, dt) for a key stored as key__<dt>."""
    for dt in ("f32", "f16", "bf16"):
        k = f"{key}__{dt}"
        if k in z.files:
            return z[k], dt
    raise KeyError(key)


_Q = re.compile(r"^(q|sq|qs)_(.+)_(f32|f16|bf16)_m(\d+)_b(\d+)(?:_(\d+):(\d+))?__(f32|f16|bf16)$")
_S = re.compile(r"^s_(.+)_(f32|f16|bf16)_(\d+):(\d+)__(f32|f16|bf16)$")


def quant_cases(z):
    """yields dict(kind, name, dt, m, B, N, M, x, y) for every q/sq/qs/s fixture."""
    for k in z.files:
        mt = _Q.match(k)
        if mt:
            kind, name, dt, m, B, N, M, odt = mt.groups()
            x, _ = get(z, f"in_{name}_{dt}")
            yield dict(kind=kind, name=name, dt=dt, m=int(m), B=int(B), N=int(N or 2), M=int(M or 4), x=x, y=z[k],
                       odt=odt, key=k)
            continue
        mt = _S.match(k)
        if mt:
            name, dt, N, M, odt = mt.groups()
            x, _ = get(z, f"in_{name}_{dt}")
            yield dict(kind="s", name=name, dt=dt, m=0, B=0, N=int(N), M=int(M), x=x, y=z[k], odt=odt, key=k)


def bits(a):
    a = np.ascontiguousarray(a)
    if a.dtype == np.float32:
        return a.view(np.uint32)
    if a.dtype == np.float16:
        return a.view(np.uint16)
    return a


def is_nan(a, dt):
    if dt == "bf16":
        return (a & 0x7fff) > 0x7f80
    return np.isnan(a)


def mismatches(a, b, dt):
    """bit mismatches, counting NaN==NaN (any payload) as equal."""
    ba, bb = bits(a), bits(b)
    neq = ba != bb
    if neq.any():
        neq &= ~(is_nan(a, dt) & is_nan(b, dt))
    return int(neq.sum())
