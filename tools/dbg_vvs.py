import sys, os
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "tests")); sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
import numpy as np
import oracle_lib
from memilio_b200 import capi, flat, synthetic
import test_gpu_vvs as t
p = t._vvs_variant("rkf45")
with capi.GraphSimulation(p) as sim:
    for k in range(10):
        sim.advance(0.5 * (k + 1))
        st = sim.state(); cnt = sim.counters(); dts = sim.node_dt()
        ref = oracle_lib.graph_simulate(p, 0, 0.5 * (k + 1))
        d = np.abs(st[0] - ref["state"]); rel = d / (np.abs(ref["state"]) + 1e-9)
        i = np.unravel_index(rel.argmax(), rel.shape)
        print(k, "maxrel", rel.max(), "at", i, "val", ref["state"][i], "cnt eq", np.array_equal(cnt[0][:, :3], ref["counters"][:, :3]),
              "dt rel", (np.abs(dts[0] - ref["dt"]) / ref["dt"]).max())
