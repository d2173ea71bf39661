#!/usr/bin/env python
"""Reads the clock stamps written by the INT8 GLM kernel under B200NUTS_I8_TRACE=<file> (CTA 0 of one launch) and prints, per row-tile
iteration, when each role did what (cycles relative to the first stamp).
MMA thread:  a0 start of issue_a, a1 after a_empty / q_full, a2 after the first X plane arrived, a3 first product issued + committed,
             b4 start of issue_b, b5 after dm_full, b6 second product issued + committed
epilogue (warp 2 lane 0): e0 before a_full, e1 after a_full, e2 Eta read + a_empty arrived, e3 link + digits done, e4 after dm_empty,
             e5 dm stored + dm_full arrived
producer:    p0..p5 issue time of the six X plane copies of the tile"""
import sys
import numpy as np

K = 96
t = np.fromfile(sys.argv[1], dtype=np.int64).reshape(-1, K, 8)
t0 = t[t > 0].min()
n = int(sys.argv[2]) if len(sys.argv) > 2 else 30
rel = lambda v: (v - t0) if v > 0 else -1
print(" it |   a0     a1     a2     a3 |   b4     b5     b6 |   e0     e1     e2     e3     e4     e5 | p0 .. p5")
for it in range(n):
    m, e, p = t[0, it], t[1, it], t[2, it]
    print(f"{it:3d} | " + " ".join(f"{rel(m[k]):6d}" for k in range(4)) + " | " + " ".join(f"{rel(m[k]):6d}" for k in (4, 5, 6)) + " | " +
          " ".join(f"{rel(e[k]):6d}" for k in range(6)) + " | " + " ".join(f"{rel(p[k]):6d}" for k in range(6)))
e = t[1]
valid = [it for it in range(1, K) if e[it, 1] > 0 and e[it - 1, 1] > 0]
if valid:
    per = np.array([e[it, 1] - e[it - 1, 1] for it in valid])
    print("tile period (e1 to e1): median", int(np.median(per)), "mean", int(per.mean()))
    print("epilogue: wait a_full", int(np.median([e[i, 1] - e[i, 0] for i in valid])), " read Eta", int(np.median([e[i, 2] - e[i, 1] for i in valid])),
          " link+digits", int(np.median([e[i, 3] - e[i, 2] for i in valid])), " wait dm_empty", int(np.median([e[i, 4] - e[i, 3] for i in valid])),
          " store", int(np.median([e[i, 5] - e[i, 4] for i in valid])))
    m = t[0]
    print("MMA thread: A waits a_empty/q", int(np.median([m[i, 1] - m[i, 0] for i in valid])), " first X plane", int(np.median([m[i, 2] - m[i, 1] for i in valid])),
          " issue A", int(np.median([m[i, 3] - m[i, 2] for i in valid])), " B wait dm_full", int(np.median([m[i, 5] - m[i, 4] for i in valid])),
          " issue B", int(np.median([m[i, 6] - m[i, 5] for i in valid])))

for role, name in ((1, "warp 2 (sub-partition 2)"), (3, "warp 5 (sub-partition 1, beside the MMA thread)")):
    if t.shape[0] > role and valid:
        e = t[role]
        f = lambda a, b: int(np.median([e[i, a] - e[i, b] for i in valid if e[i, a] > 0 and e[i, b] > 0] or [0]))
        print(f"{name}: wait a_full {f(1, 0)}  read Eta {f(2, 1)}  recombine {f(6, 2)}  link g0 {f(7, 6)}  digits+wait g0 {f(4, 7)}  group 1 + stores {f(5, 4)}  next tile {int(np.median([e[i + 1, 0] - e[i, 5] for i in valid[:-1]]))}")
