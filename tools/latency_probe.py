"""p50 / p99 of qppvm_solve_one over a tick sequence (configs[4]); QPPVM_RESIDENT=0 for the launch-per-tick path."""
import sys, time, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from qppvm_b200 import api, gen
from qppvm_b200.layout import CONFIGS, layout
n = int(sys.argv[1]) if len(sys.argv) > 1 else 10000
desc = CONFIGS[4]["desc"]; L = layout(desc)
st0 = gen.generate_states(desc, 1, 77)[0]
rng = np.random.default_rng(77); na = desc.n_a
states = np.repeat(st0[None], n, axis=0)
states[:, :na] += np.cumsum(rng.normal(0.0, 2e-3, (n, na)), axis=0)
states[:, na:2 * na] += np.cumsum(rng.normal(0.0, 5e-3, (n, na)), axis=0)
recs = gen.records_from_states(desc, states)
s = api.Solver(desc); o = np.empty(L.out_doubles)
for i in range(300): s.solve_one(recs[i], o)
for mode in ("hot", "cold"):
    lat = np.empty(n)
    for i in range(n):
        if mode == "cold": s.reset_warm()
        t0 = time.perf_counter(); s.solve_one(recs[i], o); lat[i] = time.perf_counter() - t0
    print(mode, "resident" if os.environ.get("QPPVM_RESIDENT", "1") != "0" else "launch", "p50 %.1f p99 %.1f max %.1f us" % (np.percentile(lat, 50) * 1e6, np.percentile(lat, 99) * 1e6, lat.max() * 1e6))

if os.environ.get("QPPVM_RESIDENT", "1") != "0":
    acc = np.zeros(6); m = 300
    for i in range(m):
        s.solve_one(recs[i], o)
        st = s.tick_stamps().astype(np.int64)
        acc += np.diff(st)
    print("stages (us): copy %.1f prepare %.1f handoff %.1f solve %.1f handoff %.1f certify+publish %.1f | chain %.1f" % (*(acc / m / 1e3), acc.sum() / m / 1e3))
