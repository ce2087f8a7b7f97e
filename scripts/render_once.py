#!/usr/bin/env python
"""One render of a BASELINE workload class through the public API (for ncu captures: no warm-up loops, no timers)."""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mitsuba2_b200 import abi, api, scenes  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--workload", default="cbox", choices=["cbox", "veach", "grid", "matball"])
ap.add_argument("--res", type=int, default=1024)
ap.add_argument("--spp", type=int, default=32)
ap.add_argument("--renders", type=int, default=1)
ap.add_argument("--lbvh", action="store_true")
ap.add_argument("--flags", type=int, default=abi.FLAG_SINGLE_SET)
a = ap.parse_args()
H = api.Host()
sb = {"cbox": lambda: scenes.cbox(H, a.res, a.res), "veach": lambda: scenes.veach_mis(H, a.res, a.res * 9 // 16),
      "grid": lambda: scenes.instanced_grid(H, a.res, a.res), "matball": lambda: scenes.matball(H, a.res, a.res * 9 // 16)}[a.workload]()
if a.lbvh:
    sb.build_flags = abi.BUILD_DEVICE_LBVH
scene = api.Scene(sb)
integ = api.PathIntegrator(max_depth=6, rr_depth=5, samples_per_pass=16, block_size=32)
for _ in range(a.renders):
    integ.render_device(scene, a.spp, flags=a.flags)
print({k: integ.stats[k] for k in ("n_samples", "n_segments", "n_iterations", "n_launches", "ms_total")})
