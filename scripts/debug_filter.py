import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
import oracle
from paris_b200 import capi
from cases import both_det
port = oracle.Port()
ctx = capi.Context(0)
for n_row, n_col in [(96, 96), (100, 37), (256, 64), (1024, 16), (2048, 8)]:
    odet, det = both_det(n_row, n_col, l_px=0.2)
    rng = np.random.default_rng(1)
    p = rng.standard_normal((n_col, n_row)).astype(np.float32)
    ref = port.filter(port.weight(p, odet), odet)
    f = ctx.filter_create(capi.filter_size(n_row), 0.2)
    slot_bytes, pitch = capi.stack_slot_bytes(n_row, n_col)
    for layout in (0, 1):
        if layout == 1 and n_row <= 64:
            continue
        d = ctx.dev_alloc(p.nbytes); ctx.proj_h2d(p, d, n_row, n_col)
        st = ctx.stack_alloc(n_row, n_col, 2)
        ctx.filter_to_stack(d, det, f, st, 1, layout)
        out = np.empty((n_row, pitch), np.float32)
        ctx.proj_d2h(st + slot_bytes, out, pitch, n_row)
        if layout == 1:
            o2 = np.empty_like(out); o2[:, 0::2] = out[:, :pitch // 2]; o2[:, 1::2] = out[:, pitch // 2:]; out = o2
        got = out[:, :n_col].T
        err = np.abs(got - ref).max() / np.abs(ref).max()
        print(n_row, n_col, "layout", layout, "rel err", err)
        ctx.dev_free(d); ctx.stack_free(st)
    ctx.filter_destroy(f)
