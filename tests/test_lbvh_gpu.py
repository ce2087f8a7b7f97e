"""Device BVH build (b2_lbvh.cu: LBVH -> BVH4) vs the host binned-SAH builder: any valid tree gives the same closest
hits (lexicographic (t, -prim) minimum), so the films must be IDENTICAL bit for bit — the way the reference
validates its kd-tree against brute force (src/librender/tests/test_kdtrees.py:26-101)."""
import ctypes as C

import numpy as np
import pytest

from mitsuba2_b200 import abi, scenes

pytestmark = pytest.mark.gpu


def _render(b2, sb, spp=4, flags=0):
    sc = b2.Scene(sb)
    rp = abi.render_params(max_depth=6, rr_depth=5, spp=spp, block_size=32, flags=flags)
    out = np.empty((sc.height, sc.width, 5), np.float32)
    st = abi.B2Stats()
    assert b2.lib().b2_render(sc._h, C.byref(rp), out.ctypes.data, C.byref(st)) == 0, b2.lib().b2_last_error()
    return out, st.as_dict(), sc


@pytest.mark.parametrize("name", ["cbox", "matball", "veach", "grid"])
def test_lbvh_film_identical_to_host_build(b2, name):
    H = b2.Host()
    make = {"cbox": lambda: scenes.cbox(H, 64, 64), "matball": lambda: scenes.matball(H, 96, 54, rfilter="box", level=3),
            "veach": lambda: scenes.veach_mis(H, 96, 54), "grid": lambda: scenes.instanced_grid(H, 96, 64, grid=21, lattice=4)}[name]
    a, sa, _ = _render(b2, make())
    sb = make(); sb.build_flags = abi.BUILD_DEVICE_LBVH
    b, sbs, _ = _render(b2, sb)
    assert sa["n_segments"] == sbs["n_segments"] and sa["n_shadow_rays"] == sbs["n_shadow_rays"]
    assert np.array_equal(a.view(np.uint32), b.view(np.uint32))
    assert sbs["bvh_depth"] >= 1 and sbs["bvh_nodes"] >= 1


def test_lbvh_intersect_vs_bruteforce(b2, oracle):
    sb = scenes.matball(b2.Host(), 96, 54, rfilter="box", level=3); sb.build_flags = abi.BUILD_DEVICE_LBVH
    sc = b2.Scene(sb)
    rng = np.random.default_rng(11)
    n = 20000
    o = rng.uniform(-3, 3, (n, 3)).astype(np.float32); o[:, 1] = np.abs(o[:, 1]) + 0.01
    d = rng.normal(size=(n, 3)).astype(np.float32); d /= np.linalg.norm(d, axis=1, keepdims=True).astype(np.float32)
    ro = np.ascontiguousarray(np.concatenate([o, np.full((n, 1), 1e-4, np.float32)], 1), np.float32)
    rd = np.ascontiguousarray(np.concatenate([d, np.full((n, 1), np.inf, np.float32)], 1), np.float32)
    hit = sc.ray_intersect_preliminary(ro, rd)
    desc = sb.desc()
    hit2 = np.zeros_like(hit)
    oracle.lib().b2o_stage_intersect(C.byref(desc), n, ro.ctypes.data, rd.ctypes.data, hit2.ctypes.data, None, 1)
    assert (np.isfinite(hit2[:, 0])).sum() > 2000
    assert np.array_equal(hit.view(np.uint32), hit2.view(np.uint32))


def test_lbvh_million_triangles(b2):
    """A flattened lattice (125 instances x 9800 triangles = 1.2 M triangles in ONE mesh): device build in milliseconds,
    same film as the host SAH build; the BVH counters show what the cheaper tree costs in traversal."""
    H = b2.Host()
    a, sa, _ = _render(b2, scenes.instanced_grid(H, 128, 128, lattice=5, flatten=True), spp=2, flags=abi.FLAG_COUNT_BVH)
    sb = scenes.instanced_grid(H, 128, 128, lattice=5, flatten=True); sb.build_flags = abi.BUILD_DEVICE_LBVH
    b, sbs, _ = _render(b2, sb, spp=2, flags=abi.FLAG_COUNT_BVH)
    assert np.array_equal(a.view(np.uint32), b.view(np.uint32))
    assert sbs["bvh_tris"] == 125 * 9800 + 2
    print("LBVH 1.2M: device build %.2f ms (host SAH %.0f ms), depth %d vs %d, nodes/ray %.1f vs %.1f" % (
        sbs["ms_build"], sa["ms_build"], sbs["bvh_depth"], sa["bvh_depth"],
        sbs["n_nodes_visited"] / (sbs["n_segments"] + sbs["n_shadow_rays"]), sa["n_nodes_visited"] / (sa["n_segments"] + sa["n_shadow_rays"])))
    assert sbs["ms_build"] < 50.0


def test_scene_update_equals_fresh_create(b2):
    """b2_scene_update (Scene::parameters_changed -> accel_parameters_changed, scene.cpp:242-260): moved geometry, then a
    completely different scene (instanced, other film size), behind ONE handle; each film is bit-identical to the film of a
    freshly created scene, and the device rebuild is reported in the stats."""
    H = b2.Host()
    frames = [lambda: scenes.cbox(H, 64, 64, lift=0.0), lambda: scenes.cbox(H, 64, 64, lift=40.0),
              lambda: scenes.instanced_grid(H, 96, 64, grid=21, lattice=4), lambda: scenes.cbox(H, 48, 80, lift=10.0)]
    first = frames[0](); first.build_flags = abi.BUILD_DEVICE_LBVH
    sc = b2.Scene(first)
    rp = abi.render_params(max_depth=6, rr_depth=5, spp=8, block_size=32)
    prev = None
    for k, make in enumerate(frames):
        sb = make(); sb.build_flags = abi.BUILD_DEVICE_LBVH
        if k:
            sc.update(sb)
        out = np.empty((sc.height, sc.width, 5), np.float32); st = abi.B2Stats()
        assert b2.lib().b2_render(sc._h, C.byref(rp), out.ctypes.data, C.byref(st)) == 0, b2.lib().b2_last_error()
        ref, _, fresh = _render(b2, make(), spp=8)           # host-built tree, fresh handle: same film (tie rule is tree-independent)
        fresh.close()
        assert np.array_equal(out.view(np.uint32), ref.view(np.uint32)), k
        if prev is not None and prev.shape == out.shape:
            assert not np.array_equal(prev, out)             # the geometry did change
        prev = out
    # a failing update leaves the handle usable with its old contents
    bad = frames[3](); d = bad.desc(); d.meshes[0].bsdf = 99
    assert b2.lib().b2_scene_update(sc._h, C.byref(d)) == 1
    out2 = np.empty_like(prev)
    assert b2.lib().b2_render(sc._h, C.byref(rp), out2.ctypes.data, None) == 0
    assert np.array_equal(out2.view(np.uint32), prev.view(np.uint32))
