"""CPU-only: the C-ABI library (mitsuba2_b200/libb2pt.so) loads, exports every entry point include/b2pt.h declares, its
structs have the layout the ctypes binding assumes, and its HOST helpers (sensor / film / spiral / BVH build — no GPU
work) agree with the oracle's restatements of the reference's host code. No compute entry point is called without a GPU,
except to check that it fails loudly."""
import ctypes as C
import os
import re
import subprocess

import numpy as np
import pytest

from mitsuba2_b200 import abi, api, scenes

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "b2pt.h")


def _bytes(x):
    return bytes(memoryview(x)) if not isinstance(x, np.ndarray) else x.tobytes()


def test_exports_every_declared_symbol():
    names = re.findall(r"B2_API\s+[\w\s\*]+?\b(b2_\w+)\s*\(", open(HEADER).read())
    assert len(names) >= 20, names
    L = api.lib()
    missing = [n for n in names if not hasattr(L, n)]
    assert not missing, missing


def test_struct_layouts_match_the_header(tmp_path):
    structs = ["B2Bsdf", "B2Texture", "B2Emitter", "B2Mesh", "B2Shape", "B2Instance", "B2Sensor", "B2Film", "B2SceneDesc",
               "B2RenderParams", "B2Stats"]
    src = tmp_path / "sz.c"
    src.write_text('#include "b2pt.h"\n#include <stdio.h>\nint main(void) {\n' +
                   "".join(f'  printf("{s} %zu\\n", sizeof({s}));\n' for s in structs) + "  return 0;\n}\n")
    exe = tmp_path / "sz"
    subprocess.run(["gcc", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)], check=True)
    out = dict(l.split() for l in subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.splitlines())
    for s in structs:
        assert int(out[s]) == C.sizeof(getattr(abi, s)), (s, out[s], C.sizeof(getattr(abi, s)))


def test_flag_constants_match_the_header(tmp_path):
    """abi.FLAG_* / flag_yield_threshold / flag_blocks mirror the B2_FLAG_* macros of include/b2pt.h (evaluated by the C compiler)."""
    names = ["B2_FLAG_STAGE_TIMING", "B2_FLAG_SINGLE_SET", "B2_FLAG_COUNT_BVH", "B2_FLAG_PERSISTENT_TRACE", "B2_FLAG_RESUME_TRACE",
             "B2_FLAG_YIELD_THRESHOLD(20)", "B2_FLAG_TRACE_BLOCK(2)", "B2_FLAG_SHADE_BLOCK(3)"]
    src = tmp_path / "fl.c"
    src.write_text('#include "b2pt.h"\n#include <stdio.h>\nint main(void) {\n' +
                   "".join(f'  printf("%u\\n", (unsigned) ({n}));\n' for n in names) + "  return 0;\n}\n")
    exe = tmp_path / "fl"
    subprocess.run(["gcc", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)], check=True)
    got = [int(x) for x in subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.split()]
    want = [abi.FLAG_STAGE_TIMING, abi.FLAG_SINGLE_SET, abi.FLAG_COUNT_BVH, abi.FLAG_PERSISTENT_TRACE, abi.FLAG_RESUME_TRACE,
            abi.flag_yield_threshold(20), abi.flag_blocks(trace_sel=2), abi.flag_blocks(shade_sel=3)]
    assert got == want, (got, want)
    # the fields do not overlap
    assert abi.flag_yield_threshold(31) & (abi.flag_blocks(3, 3) | 0xff) == 0 and abi.flag_blocks(3, 0) & abi.flag_blocks(0, 3) == 0


def test_host_helpers_match_oracle(oracle):
    H, O = api.Host(), oracle.OracleHost()
    assert _bytes(H.look_at((278, 273, -800), (278, 273, -799), (0, 1, 0))) == _bytes(O.look_at((278, 273, -800), (278, 273, -799), (0, 1, 0)))
    assert _bytes(H.look_at((3.7, 2.6, 3.5), (0.2, 0.8, 0.0), (0.1, 1, 0.2))) == _bytes(O.look_at((3.7, 2.6, 3.5), (0.2, 0.8, 0.0), (0.1, 1, 0.2)))
    for axis in ("x", "y", "diagonal", "smaller", "larger"):
        for w, h in ((64, 64), (96, 54), (48, 80)):
            assert np.float32(H.parse_fov(39.3077, axis, w, h)) == np.float32(O.parse_fov(39.3077, axis, w, h)), (axis, w, h)
    for w, h in ((64, 64), (80, 48)):
        assert _bytes(H.film_box(w, h)) == _bytes(O.film_box(w, h))
        assert _bytes(H.film_gaussian(w, h, 0.5)) == _bytes(O.film_gaussian(w, h, 0.5))
    film = H.film_gaussian(96, 54, 0.5)
    film.crop_offset[:] = (8, 4); film.crop_size[:] = (40, 30)
    tw = H.look_at((3.7, 2.6, 3.5), (0.2, 0.8, 0.0), (0, 1, 0))
    assert _bytes(H.perspective_sensor(tw, 28.0, 0.1, 100.0, film)) == _bytes(O.perspective_sensor(tw, 28.0, 0.1, 100.0, film))


def test_spiral_matches_oracle(oracle):
    L = oracle.lib()
    for w, h, bs in ((64, 64, 32), (80, 48, 32), (100, 37, 16), (33, 33, 8), (1, 1, 32)):
        mine = api.Host().spiral(w, h, bs)
        n = C.c_uint32()
        L.b2o_spiral(w, h, bs, C.byref(n), None)
        ref = np.zeros((n.value, 2), np.int32)
        L.b2o_spiral(w, h, bs, C.byref(n), ref.ctypes.data_as(C.POINTER(C.c_int32)))
        assert mine.shape == ref.shape and np.array_equal(mine, ref), (w, h, bs)
        # every block exactly once
        assert len({tuple(b) for b in mine}) == ((w + bs - 1) // bs) * ((h + bs - 1) // bs)


def test_host_bvh_is_a_valid_tree():
    """b2_host_build_bvh4: every triangle appears in exactly one leaf and lies inside all boxes on its path."""
    L = api.lib()
    L.b2_host_build_bvh4.argtypes = [C.c_void_p, C.c_uint32, C.c_void_p, C.c_uint32, C.POINTER(C.c_uint32), C.c_void_p]
    rng = np.random.default_rng(3)
    for n in (1, 2, 7, 38, 500):
        tri = (rng.normal(size=(n, 1, 3)) * 5 + rng.normal(size=(n, 3, 3)) * 0.3).astype(np.float32)
        nodes = np.zeros((2 * n + 8, 32), np.float32); order = np.zeros(n, np.uint32); nn = C.c_uint32()
        assert L.b2_host_build_bvh4(tri.ctypes.data, n, nodes.ctypes.data, len(nodes), C.byref(nn), order.ctypes.data) == 0
        assert sorted(order.tolist()) == list(range(n))
        seen = np.zeros(n, np.int32)
        codes = nodes.view(np.int32)

        def walk(node, boxes):
            for k in range(4):
                c = int(codes[node, 24 + k])
                if c == 0:
                    continue
                box = (nodes[node, [0 + k, 4 + k, 8 + k]], nodes[node, [12 + k, 16 + k, 20 + k]])
                if c > 0:
                    walk(c, boxes + [box])
                else:
                    u = c & 0xffffffff; first, count = u & 0x07ffffff, (u >> 27) & 0xf
                    assert 1 <= count <= 2
                    for s in range(first, first + count):
                        t = tri[order[s]]; seen[order[s]] += 1
                        for lo, hi in boxes + [box]:
                            assert (t >= lo).all() and (t <= hi).all()
        walk(0, [])
        assert (seen == 1).all()


def test_compute_entry_points_fail_loudly_without_a_gpu():
    if api.device_count() > 0:
        pytest.skip("a GPU is present")
    sb = scenes.cbox(api.Host(), 16, 16)
    with pytest.raises(api.B2Error) as e:
        api.Scene(sb)
    assert e.value.code == 4 and "no CUDA device" in str(e.value)        # B2_ERR_NO_DEVICE: no CPU fallback


def test_scene_builders_reject_bad_input():
    sb = abi.SceneBuffers()
    d = sb.add_dielectric(1.5, 1.0)
    with pytest.raises(ValueError):
        sb.make_twosided(d)                                               # twosided.cpp:90-91
    with pytest.raises(ValueError):                                       # direct.cpp:79-84, 93-94
        api.DirectIntegrator(shading_samples=2, emitter_samples=1)
    with pytest.raises(ValueError):
        api.DirectIntegrator(emitter_samples=0, bsdf_samples=0)
    rp = api.DirectIntegrator(emitter_samples=2, bsdf_samples=3).params(8)
    assert (rp.direct_emitter_samples, rp.direct_bsdf_samples, rp.max_depth) == (2, 3, 2)
    rp = api.DirectIntegrator().params(8)                                 # shading_samples = 1 == path with max_depth 2
    assert (rp.direct_emitter_samples, rp.direct_bsdf_samples, rp.max_depth) == (0, 0, 2)


def test_parallel_bvh_build_is_deterministic():
    """The host BVH build runs its top levels on several threads; the tree must not depend on the thread count
    (B2_BVH_THREADS=1 forces the serial path)."""
    import sys
    code = (
        "import sys, ctypes as C, hashlib; sys.path.insert(0, %r)\n"
        "import numpy as np\n"
        "from mitsuba2_b200 import api\n"
        "L = api.lib(); L.b2_host_build_bvh4.argtypes = [C.c_void_p, C.c_uint32, C.c_void_p, C.c_uint32, C.POINTER(C.c_uint32), C.c_void_p]\n"
        "rng = np.random.default_rng(1); n = 80000\n"
        "c = rng.normal(size=(n, 3)); c /= np.linalg.norm(c, axis=1, keepdims=True)\n"
        "tri = (c[:, None, :] + 0.01 * rng.normal(size=(n, 3, 3))).astype(np.float32)\n"
        "nodes = np.zeros((2 * n + 8, 32), np.float32); order = np.zeros(n, np.uint32); nn = C.c_uint32()\n"
        "assert L.b2_host_build_bvh4(tri.ctypes.data, n, nodes.ctypes.data, len(nodes), C.byref(nn), order.ctypes.data) == 0\n"
        "print(nn.value, hashlib.md5(nodes[:nn.value].tobytes() + order.tobytes()).hexdigest())\n" % ROOT)
    outs = []
    for threads in ("1", "4", "16"):
        env = dict(os.environ); env["B2_BVH_THREADS"] = threads
        r = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True)
        assert r.returncode == 0, r.stderr
        outs.append(r.stdout.strip())
    assert outs[0] == outs[1] == outs[2], outs


def _create_error(sb):
    with pytest.raises(api.B2Error) as e:
        api.Scene(sb)
    return e.value


def test_scene_description_is_validated_like_the_reference():
    """b2_scene_create validates the description before it touches the device, with the reference's error wording
    where it has one (no GPU needed: malformed input never reaches CUDA)."""
    H = api.Host()
    sb = scenes.cbox(H, 16, 16); sb.meshes[1]["bsdf"] = 99
    assert _create_error(sb).code == 1 and "bad bsdf index" in str(_create_error(sb))
    sb = scenes.cbox(H, 16, 16); sb.meshes[2]["faces"] = sb.meshes[2]["faces"].copy(); sb.meshes[2]["faces"][0, 0] = 10_000
    assert "face index out of range" in str(_create_error(sb))
    # two environment emitters (scene.cpp:48-49)
    sb = scenes.cbox(H, 16, 16); sb.add_constant_emitter((1, 1, 1)); sb.add_constant_emitter((1, 1, 1))
    assert "Only one environment emitter can be specified per scene." in str(_create_error(sb))
    # an emitter inside a shapegroup (shapegroup.cpp:24-25)
    sb = scenes.cbox(H, 16, 16); g = sb.add_shapegroup([0]); sb.add_instance(g, np.eye(4, dtype=np.float32))
    assert "Instancing of emitters is not supported" in str(_create_error(sb))
    # analytic shapes
    sb = scenes.cbox(H, 16, 16); sb.add_sphere((0, 0, 0), 1.0, bsdf=77)
    assert "shape 0: bad bsdf index" in str(_create_error(sb))
    sb = scenes.cbox(H, 16, 16); sb.add_sphere((0, 0, 0), -1.0)
    assert "sphere radius must be positive" in str(_create_error(sb))
    # BSDF records
    sb = scenes.cbox(H, 16, 16); sb.bsdfs[0].type = 42
    assert "bsdf 0: unsupported type 42" in str(_create_error(sb))
    sb = scenes.cbox(H, 16, 16, materials={"box": lambda s: s.add_roughconductor(sample_visible=False)})
    assert "sample_visible = false is not supported" in str(_create_error(sb))
    # rough dielectric, spot and directional emitters have a CUDA path (round 2): a valid description passes validation and
    # only then fails for lack of a device (code 4) on a CPU box
    sb = scenes.cbox(H, 16, 16, materials={"box": lambda s: s.add_roughdielectric(0.1, "ggx", 1.5, 1.0)})
    assert _create_error(sb).code == 4
    sb = scenes.cbox(H, 16, 16, materials={"box": lambda s: s.add_roughdielectric(0.1, "ggx", 1.5, 1.0)})
    [b for b in sb.bsdfs if b.type == abi.BSDF_ROUGHDIELECTRIC][0].distribution = 7
    assert "invalid distribution" in str(_create_error(sb))
    sb = scenes.cbox(H, 16, 16); sb.add_spot_light(np.eye(4), (1, 1, 1))
    assert _create_error(sb).code == 4
    sb = scenes.cbox(H, 16, 16); sb.add_directional_light(np.eye(4), (1, 1, 1))
    assert _create_error(sb).code == 4
    sb = scenes.cbox(H, 16, 16); sb.add_point_light((0, 0, 0), (1, 1, 1)); sb.emitters[-1].type = 9
    assert "emitters are supported" in str(_create_error(sb))
    # shapegroups must cover the non-top-level meshes exactly
    sb = scenes.cbox(H, 16, 16, instances=[np.eye(4, dtype=np.float32)])
    d = sb.desc(); d.n_top_meshes -= 1; h = C.c_void_p()
    assert api.lib().b2_scene_create(C.byref(d), C.byref(h)) == 1 and b"cover the meshes" in api.lib().b2_last_error()
    # film limits
    sb = scenes.cbox(H, 16, 16); sb.film.crop_size[0] = 0
    assert "film crop size" in str(_create_error(sb))


def test_c_record_builders_match_the_python_ones():
    """b2_host_bsdf_plastic / b2_host_shape_sphere / b2_host_shape_rectangle (for hosts that do not link the reference)
    produce the records mitsuba2_b200.abi builds in numpy — the ones the golden scenes are pinned with."""
    L = api.lib()
    L.b2_host_bsdf_plastic.argtypes = [C.c_float, C.c_float, C.POINTER(C.c_float), C.c_float, C.POINTER(C.c_float), C.c_float, C.c_int,
                                       C.POINTER(abi.B2Bsdf)]
    L.b2_host_shape_sphere.argtypes = [C.POINTER(C.c_float), C.c_float, C.c_int, C.c_int, C.POINTER(abi.B2Shape)]
    L.b2_host_shape_rectangle.argtypes = [C.POINTER(C.c_float), C.c_int, C.c_int, C.POINTER(abi.B2Shape)]
    for kw in (dict(diffuse_reflectance=(0.1, 0.27, 0.36), int_ior=1.9, ext_ior=1.000277),
               dict(diffuse_reflectance=(0.7, 0.6, 0.5), int_ior=1.49, ext_ior=1.000277, specular_reflectance=(0.8, 0.9, 1.0)),
               dict(diffuse_reflectance=(0.8, 0.3, 0.2), int_ior=1.49, ext_ior=1.000277, nonlinear=True)):
        sb = abi.SceneBuffers(); ref = sb.bsdfs[sb.add_plastic(**kw)]
        d = kw["diffuse_reflectance"]; s = kw.get("specular_reflectance")
        out = abi.B2Bsdf()
        assert L.b2_host_bsdf_plastic(kw["int_ior"], kw["ext_ior"], (C.c_float * 3)(*d), abi._rgb_mean(d),
                                      (C.c_float * 3)(*s) if s else None, abi._rgb_mean(s) if s else 1.0, int(kw.get("nonlinear", False)),
                                      C.byref(out)) == 0
        assert _bytes(out) == _bytes(ref), kw
    sb = abi.SceneBuffers(); sb.add_diffuse()
    for center, radius, flip in (((186.0, 64.0, 169.0), 64.0, False), ((1.5, -2.25, 0.125), 0.5, True)):
        idx = sb.add_sphere(center, radius, 0, flip); ref = sb.shapes[idx]
        out = abi.B2Shape()
        assert L.b2_host_shape_sphere((C.c_float * 3)(*center), radius, int(flip), 0, C.byref(out)) == 0
        assert _bytes(out) == _bytes(ref)
    for M, flip in (([[64, 0, 0, 278], [0, 0, -1, 548], [0, 32, 0, 280], [0, 0, 0, 1]], False),
                    ([[32, 0, 0, 100], [0, 64, 0, 64], [0, 0, 1, 400], [0, 0, 0, 1]], True),
                    ([[0.3, -1.2, 0.5, 1.0], [0.9, 0.1, -0.4, 2.0], [0.2, 0.7, 1.1, -3.0], [0, 0, 0, 1]], False)):
        idx = sb.add_rectangle(M, 0, flip); ref = sb.shapes[idx]
        out = abi.B2Shape()
        flat = (C.c_float * 16)(*[float(x) for row in M for x in row])
        assert L.b2_host_shape_rectangle(flat, int(flip), 0, C.byref(out)) == 0
        assert (out.type, out.bsdf, out.emitter, out.flip_normals) == (ref.type, ref.bsdf, ref.emitter, ref.flip_normals)
        assert np.array_equal(np.array(out.to_world[:]), np.array(ref.to_world[:]))
        # (the inverse is computed in double and rounded once on both sides, by different formulas: last-ulp agreement)
        assert np.allclose(np.array(out.to_object[:]), np.array(ref.to_object[:]), rtol=3e-7, atol=1e-9), (M, flip)


def _bvh4_depth(nodes, n_nodes):
    """levels of a BVH4 in the 128-byte node layout (child codes = words 24..27: > 0 inner node index)"""
    codes = nodes[:n_nodes, 24:28].view(np.int32)
    level = np.array([0]); depth = 0
    while len(level):
        depth += 1
        c = codes[level].ravel()
        level = c[c > 0]
    return depth


def test_host_bvh_depth_is_bounded():
    """A 1 M-triangle multi-scale soup (triangle sizes over five decades, clustered): the binned-SAH builder alone goes 30+
    levels deep on such input; the level budget (median splits once it runs short) keeps every tree within
    B2_BVH_MAX_DEPTH = 42 levels, which is what the traversal kernels' 128-entry stacks are sized for — no mesh is refused
    for its depth. Also: every triangle is referenced exactly once."""
    L = api.lib()
    L.b2_host_build_bvh4.argtypes = [C.c_void_p, C.c_uint32, C.c_void_p, C.c_uint32, C.POINTER(C.c_uint32), C.c_void_p]
    rng = np.random.default_rng(7); n = 1_000_000
    centres = rng.normal(size=(64, 3)) * 50
    c = centres[rng.integers(0, 64, n)] + rng.normal(size=(n, 3)) * np.exp(rng.uniform(-6, 2, (n, 1)))
    size = np.exp(rng.uniform(-9, 2, (n, 1, 1)))
    tri = (c[:, None, :] + size * rng.normal(size=(n, 3, 3))).astype(np.float32)
    tri[:1000] = tri[0]                                  # a thousand identical triangles (all centroids coincide)
    nodes = np.zeros((n + 8, 32), np.float32); order = np.zeros(n, np.uint32); nn = C.c_uint32()
    assert L.b2_host_build_bvh4(tri.ctypes.data, n, nodes.ctypes.data, len(nodes), C.byref(nn), order.ctypes.data) == 0
    assert np.array_equal(np.sort(order), np.arange(n, dtype=np.uint32))
    d = _bvh4_depth(nodes, nn.value)
    assert 8 <= d <= 42, d
    # a long thin strip (SAH splits it one triangle at a time)
    x = np.arange(200_000, dtype=np.float64)
    strip = np.stack([np.stack([x * x * 1e-6, 0 * x, 0 * x], 1), np.stack([x * x * 1e-6 + 1e-3, 0 * x + 1, 0 * x], 1),
                      np.stack([(x + 1) ** 2 * 1e-6, 0 * x, 0 * x + 1e-3], 1)], 1).astype(np.float32)
    assert L.b2_host_build_bvh4(strip.ctypes.data, len(strip), nodes.ctypes.data, len(nodes), C.byref(nn), order.ctypes.data) == 0
    assert _bvh4_depth(nodes, nn.value) <= 42
