import numpy as np, sys, time
sys.path.insert(0, "/root/repo")
import grape_vector_db_b200 as gv
from grape_vector_db_b200 import synth
def bf16(x):
    u = np.ascontiguousarray(x, dtype=np.float32).view(np.uint32).astype(np.uint64)
    r = ((u + 0x7FFF + ((u >> 16) & 1)) >> 16) << 16
    return r.astype(np.uint32).view(np.float32)
for dim, n, nq in ((768, 5000, 200), (128, 3000, 70), (256, 700, 130)):
    rows = synth.lowrank_rows(0, n, dim); qs = synth.lowrank_queries(0, nq, dim)
    with gv.GpuIndex(dim) as idx:
        idx.add(rows)
        d = idx.approx_dot(qs)
    ref = bf16(qs).astype(np.float64) @ bf16(rows).astype(np.float64).T
    exact = qs.astype(np.float64) @ rows.astype(np.float64).T
    scale = np.linalg.norm(qs.astype(np.float64), axis=1)[:, None] * np.linalg.norm(rows.astype(np.float64), axis=1)[None, :]
    print(dim, n, nq, "max |gpu - bf16 ref| / (|q||r|):", float(np.max(np.abs(d - ref) / scale)), " max |gpu - exact| / (|q||r|):", float(np.max(np.abs(d - exact) / scale)))
