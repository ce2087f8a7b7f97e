#!/usr/bin/env python
"""SURVEY 8(f) rank 2 / VERDICT item 6: the flattened C5 lattice (1000 x 9800 = 9.8 M triangles in one mesh): device LBVH
build time vs the host SAH build, and the film rendered with either tree (must be identical). Prints one JSON line."""
import ctypes as C
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mitsuba2_b200 import abi, api, scenes  # noqa: E402

lattice = int(sys.argv[1]) if len(sys.argv) > 1 else 10
res = 512
out = {}
films = {}
for name, flag in (("device_lbvh", abi.BUILD_DEVICE_LBVH), ("host_sah", abi.BUILD_HOST_SAH)):
    sb = scenes.instanced_grid(api.Host(), res, res, lattice=lattice, flatten=True)
    sb.build_flags = flag
    t0 = time.perf_counter(); scene = api.Scene(sb); t_create = time.perf_counter() - t0
    integ = api.PathIntegrator(max_depth=6, rr_depth=5, block_size=32)
    rp = integ.params(8, flags=abi.FLAG_COUNT_BVH)
    film = np.empty((res, res, 5), np.float32); st = abi.B2Stats()
    assert api.lib().b2_render(scene._h, C.byref(rp), film.ctypes.data, C.byref(st)) == 0
    s = st.as_dict(); rays = s["n_segments"] + s["n_shadow_rays"]
    films[name] = film
    out[name] = dict(triangles=sb.n_triangles, scene_create_s=t_create, ms_build=s["ms_build"], ms_upload=s["ms_upload"], bvh_nodes=s["bvh_nodes"],
                     bvh_depth=s["bvh_depth"], render_ms=s["ms_total"], nodes_per_ray=s["n_nodes_visited"] / rays, tris_per_ray=s["n_tris_tested"] / rays)
    scene.close()
out["films_identical"] = bool(np.array_equal(films["device_lbvh"].view(np.uint32), films["host_sah"].view(np.uint32)))
print(json.dumps(out))
