"""GPU parity tests proper: the CUDA wavefront path (through the C-ABI) vs the CPU oracle on
identical inputs. Integer work (RNG state, hit primitive) is compared bit for bit; the film is
compared bit for bit where the accumulation order is deterministic (box filter, one pass) and
within 2e-6 rel-L2 where fp32 atomics reorder the sum (several passes / gaussian taps).
Tolerance stated by north_star: 1e-4 rel-L2."""
import ctypes as C

import numpy as np
import pytest

from mitsuba2_b200 import abi, scenes
from tests import golden_cases as G

pytestmark = pytest.mark.gpu


def _oracle_render(oracle, sb, rp):
    img, st = oracle.render(sb, rp)
    return img, st


def _gpu_render(b2, sb, rp):
    sc = b2.Scene(sb)
    out = np.empty((sc.height, sc.width, 5), np.float32)
    st = abi.B2Stats()
    rc = b2.lib().b2_render(sc._h, C.byref(rp), out.ctypes.data, C.byref(st))
    assert rc == 0, b2.lib().b2_last_error()
    return out, st.as_dict(), sc


def test_device_present(b2):
    assert b2.device_count() >= 1


def test_raygen_bitexact(b2, oracle):
    sb = scenes.cbox(b2.Host(), 96, 64)
    sc = b2.Scene(sb)
    rng = np.random.default_rng(1)
    n = 5000
    px = rng.integers(0, 96, n).astype(np.uint32); py = rng.integers(0, 64, n).astype(np.uint32)
    so = rng.integers(0, 2**40, n).astype(np.uint64)
    ro, rd, state, pos = sc.sample_camera_rays(px, py, so, base_seed=77)
    d = sb.desc()
    ro2 = np.zeros_like(ro); rd2 = np.zeros_like(rd); st2 = np.zeros_like(state); pos2 = np.zeros_like(pos)
    oracle.lib().b2o_stage_raygen(C.byref(d), n, px.ctypes.data, py.ctypes.data, so.ctypes.data, 77,
                                  ro2.ctypes.data, rd2.ctypes.data, st2.ctypes.data, pos2.ctypes.data)
    assert np.array_equal(state, st2)
    assert np.array_equal(ro.view(np.uint32), ro2.view(np.uint32))
    assert np.array_equal(rd.view(np.uint32), rd2.view(np.uint32))
    assert np.array_equal(pos.view(np.uint32), pos2.view(np.uint32))


def _random_rays(sb, n, seed):
    rng = np.random.default_rng(seed)
    o = rng.uniform(0, 550, (n, 3)).astype(np.float32)
    d = rng.normal(size=(n, 3)).astype(np.float32)
    d /= np.linalg.norm(d, axis=1, keepdims=True).astype(np.float32)
    ro = np.concatenate([o, np.full((n, 1), 1e-3, np.float32)], 1)
    rd = np.concatenate([d, np.full((n, 1), np.inf, np.float32)], 1)
    return np.ascontiguousarray(ro, np.float32), np.ascontiguousarray(rd, np.float32)


@pytest.mark.parametrize("name", ["cbox", "cbox_64_spp16_shapes", "cbox_64_spp16_textures"])
def test_intersect_vs_bruteforce(b2, oracle, name):
    """BVH4 traversal == brute force over all triangles (the reference validates its kd-tree the
    same way: src/librender/tests/test_kdtrees.py:26-101) — also with analytic rectangles / spheres in the
    top-level BVH and with texture coordinates (uv-derived dp_du in the SurfaceInteraction)."""
    sb = scenes.cbox(b2.Host(), 64, 64) if name == "cbox" else G.load(name, b2.Host())[0]
    sc = b2.Scene(sb)
    ro, rd = _random_rays(sb, 20000, 3)
    hit = sc.ray_intersect_preliminary(ro, rd)
    d = sb.desc()
    hit2 = np.zeros_like(hit)
    oracle.lib().b2o_stage_intersect(C.byref(d), len(ro), ro.ctypes.data, rd.ctypes.data, hit2.ctypes.data, None, 1)
    assert np.array_equal(hit.view(np.uint32), hit2.view(np.uint32))
    # shadow rays with finite maxt
    rd[:, 3] = 300.0
    occ = sc.ray_test(ro, rd)
    occ2 = np.zeros(len(ro), np.uint8)
    oracle.lib().b2o_stage_occluded(C.byref(d), len(ro), ro.ctypes.data, rd.ctypes.data, occ2.ctypes.data, 1)
    assert np.array_equal(occ, occ2.astype(bool))
    # SI fill
    si = sc.compute_surface_interaction(ro, rd, hit)
    si2 = np.zeros_like(si)
    oracle.lib().b2o_stage_si_fill(C.byref(d), len(ro), ro.ctypes.data, rd.ctypes.data, hit2.ctypes.data, None, si2.ctypes.data)
    assert np.array_equal(si.view(np.uint32), si2.view(np.uint32))


def test_instances_intersect_vs_bruteforce(b2, oracle):
    """Two-level BVH (instances: scaled, rotated, mirrored copies) == brute force over every
    instance x triangle, including the instance index and the transformed SurfaceInteraction."""
    sb, rp, _ = G.load("cbox_64_spp16_instances", b2.Host())
    sc = b2.Scene(sb)
    ro, rd = _random_rays(sb, 20000, 5)
    n = len(ro)
    hit = np.zeros((n, 4), np.float32); inst = np.zeros(n, np.int32)
    assert b2.lib().b2_stage_intersect(sc._h, n, ro.ctypes.data, rd.ctypes.data, hit.ctypes.data, inst.ctypes.data) == 0
    d = sb.desc()
    hit2 = np.zeros_like(hit); inst2 = np.zeros_like(inst)
    oracle.lib().b2o_stage_intersect(C.byref(d), n, ro.ctypes.data, rd.ctypes.data, hit2.ctypes.data, inst2.ctypes.data, 1)
    assert (inst2 >= 0).sum() > 500
    assert np.array_equal(inst, inst2)
    assert np.array_equal(hit.view(np.uint32), hit2.view(np.uint32))
    rd[:, 3] = 250.0
    occ = sc.ray_test(ro, rd)
    occ2 = np.zeros(n, np.uint8)
    oracle.lib().b2o_stage_occluded(C.byref(d), n, ro.ctypes.data, rd.ctypes.data, occ2.ctypes.data, 1)
    assert np.array_equal(occ, occ2.astype(bool))
    si = np.zeros((n, 16), np.float32); si2 = np.zeros_like(si)
    assert b2.lib().b2_stage_si_fill(sc._h, n, ro.ctypes.data, rd.ctypes.data, hit.ctypes.data, inst.ctypes.data, si.ctypes.data) == 0
    oracle.lib().b2o_stage_si_fill(C.byref(d), n, ro.ctypes.data, rd.ctypes.data, hit2.ctypes.data, inst2.ctypes.data, si2.ctypes.data)
    assert np.array_equal(si.view(np.uint32), si2.view(np.uint32))


@pytest.mark.parametrize("name", ["cbox_64_spp16", "cbox_80x48_spp8", "cbox_64_spp16_depth3_rr2",
                                  "cbox_64_spp16_bs16_seed7", "cbox_64_spp16_dielectric_lift",
                                  "cbox_64_spp16_ggx", "cbox_64_spp16_beckmann", "cbox_64_spp16_two_emitters",
                                  "cbox_64_spp16_instances", "cbox_64_spp16_env", "cbox_64_spp16_env_area",
                                  "cbox_64_spp16_env_ggx", "veach_96x54_spp16", "cbox_64_spp16_conductor", "cbox_64_spp16_plastic",
                                  "cbox_64_spp16_textures", "cbox_64_spp16_crop", "cbox_64_spp8_depthinf_rr3",
                                  "cbox_64_spp16_depth1", "cbox_64_spp16_depth2", "cbox_48_spp1", "cbox_64_spp16_shapes",
                                  "cbox_64_spp16_twosided", "cbox_64_spp16_direct", "cbox_64_spp16_stratified",
                                  "cbox_64_spp16_multijitter", "cbox_64_spp8_multijitter_nojitter",
                                  "cbox_64_spp16_shapes_textured", "cbox_64_spp16_ldsampler",
                                  "cbox_64_spp16_constant", "cbox_64_spp16_constant_area",
                                  "cbox_64_spp16_point", "cbox_64_spp16_point_area",
                                  "cbox_64_spp16_roughglass_beckmann", "cbox_64_spp16_roughglass_ggx",
                                  "cbox_64_spp16_spot", "cbox_64_spp16_directional",
                                  "cbox_64_spp25_orthogonal", "cbox_64_spp9_orthogonal_nojitter",
                                  "cbox_64_spp8_direct_e2_b3", "cbox_64_spp8_direct_e3_b0", "cbox_64_spp8_direct_e0_b2"])
def test_render_bitexact_single_pass(b2, oracle, name):
    """Box filter + one pass: one lane owns a pixel and adds its samples in order, exactly like
    render_block -> the whole XYZAW film must match the oracle bit for bit."""
    sb, rp, ref = G.load(name, b2.Host())
    gpu, gst, sc = _gpu_render(b2, sb, rp)
    cpu, cst = _oracle_render(oracle, sb, rp)
    assert gst["n_samples"] == cst["n_samples"] == sb.film.crop_size[0] * sb.film.crop_size[1] * rp.spp
    assert gst["n_segments"] == cst["n_segments"]
    assert gst["n_shadow_rays"] == cst["n_shadow_rays"]
    assert gst["n_invalid"] == cst["n_invalid"]
    # Bit for bit — except that a sample within RayEpsilon of a pixel edge lands on the NEIGHBOURING pixel (box radius
    # 0.5 + RayEpsilon, box.cpp:24-29), i.e. on a film value another lane is accumulating: there the order of two fp32
    # additions is not fixed, which can move that value by an ulp. Allow a handful of such values, nothing else.
    ga, ca = gpu.view(np.int32).astype(np.int64).ravel(), cpu.view(np.int32).astype(np.int64).ravel()
    bad = np.nonzero(ga != ca)[0]
    assert len(bad) <= 4 and (len(bad) == 0 or np.abs(ga[bad] - ca[bad]).max() <= 2), \
        f"{len(bad)} film values differ; rel_l2={oracle.rel_l2(gpu, cpu):.3e}"
    # and against the reference's own image (flip-robust statistics)
    s = G.parity_stats(sc.develop(gpu), ref)
    assert s["median_rel"] < 2e-6 and s["flipped_pixels"] <= max(20, s["pixels"] // 300), s


@pytest.mark.parametrize("name", ["cbox_64_spp16_pass1", "cbox_64_spp16_pass4", "cbox_64_spp16_gauss",
                                  "cbox_64_spp16_stratified_nojitter_pass4"])
def test_render_atomics_order(b2, oracle, name):
    """Several passes / gaussian taps: fp32 atomics add in a different order than the CPU."""
    sb, rp, ref = G.load(name, b2.Host())
    gpu, gst, sc = _gpu_render(b2, sb, rp)
    cpu, cst = _oracle_render(oracle, sb, rp)
    assert gst["n_segments"] == cst["n_segments"] and gst["n_shadow_rays"] == cst["n_shadow_rays"]
    assert oracle.rel_l2(gpu, cpu) < 2e-6


def test_instanced_grid_bitexact(b2, oracle):
    """BASELINE configs[4] in miniature: a displaced-grid mesh instantiated on a lattice with
    arbitrary rotations (two-level BVH, un-normalised object-space rays): film bit-exact vs oracle."""
    sb = scenes.instanced_grid(b2.Host(), 96, 64, grid=21, lattice=4)
    rp = abi.render_params(max_depth=6, rr_depth=5, spp=8, block_size=32)
    gpu, gst, sc = _gpu_render(b2, sb, rp)
    cpu, cst = _oracle_render(oracle, sb, rp)
    assert gst["n_segments"] == cst["n_segments"] and gst["n_shadow_rays"] == cst["n_shadow_rays"]
    assert int((gpu.view(np.uint32) != cpu.view(np.uint32)).sum()) == 0


def test_matball_bitexact(b2, oracle):
    """BASELINE configs[2] class in miniature: GGX rough-conductor ball (BVH in global memory, smooth normals) on a
    checkerboard-textured plane with texture coordinates, environment-map lighting only: film bit-exact vs oracle."""
    sb = scenes.matball(b2.Host(), 96, 54, rfilter="box", level=3)
    rp = abi.render_params(max_depth=6, rr_depth=5, spp=8, block_size=32)
    gpu, gst, sc = _gpu_render(b2, sb, rp)
    cpu, cst = _oracle_render(oracle, sb, rp)
    assert gst["n_segments"] == cst["n_segments"] and gst["n_shadow_rays"] == cst["n_shadow_rays"]
    assert int((gpu.view(np.uint32) != cpu.view(np.uint32)).sum()) == 0


@pytest.mark.parametrize("scene", ["matball", "veach", "flattened_lattice"])
def test_yield_resume_equals_static_trace(b2, scene):
    """k_trace_resume (a warp suspends its unfinished rays below a lane threshold, the next launch continues them in full
    warps) must trace exactly the rays of the static kernel and find the same hits: films bit-identical, counters equal,
    for every threshold (1 = never suspends ... 28 = suspends at once), with one lane set and two, and with a wavefront
    narrow enough that lanes wait in the RESUME list across many launches."""
    if scene == "matball":
        sb = scenes.matball(b2.Host(), 320, 180, rfilter="box", level=5)
    elif scene == "veach":
        sb = scenes.veach_mis(b2.Host(), 320, 180)
    else:
        sb = scenes.instanced_grid(b2.Host(), 256, 256, grid=41, lattice=3, flatten=True)     # 86 k triangles, deep BVH, no instances
    base = None
    R = abi.FLAG_RESUME_TRACE
    for flags, slots in [(0, 0), (R, 0), (R | abi.flag_yield_threshold(1), 0), (R | abi.flag_yield_threshold(20), 0),
                         (R | abi.flag_yield_threshold(27), 4096), (R | abi.FLAG_SINGLE_SET, 0), (R | abi.FLAG_SINGLE_SET | abi.flag_yield_threshold(24), 8192),
                         (abi.flag_blocks(2, 2), 0), (R | abi.flag_blocks(3, 2), 0)]:
        rp = abi.render_params(max_depth=6, rr_depth=5, spp=8, block_size=32, flags=flags, lane_slots=slots)
        film, st, _ = _gpu_render(b2, sb, rp)
        key = (st["n_samples"], st["n_segments"], st["n_shadow_rays"], st["n_invalid"])
        if base is None:
            base = (film, key)
            assert st["n_samples"] == sb.film.crop_size[0] * sb.film.crop_size[1] * 8
        else:
            assert key == base[1], (flags, slots, key, base[1])
            # (bit for bit, up to the unordered pair of additions of a pixel-edge sample: see test_render_bitexact_single_pass)
            ga, ca = film.view(np.int32).astype(np.int64).ravel(), base[0].view(np.int32).astype(np.int64).ravel()
            bad = np.nonzero(ga != ca)[0]
            assert len(bad) <= 32 and (len(bad) == 0 or np.abs(ga[bad] - ca[bad]).max() <= 2), (flags, slots, len(bad))


@pytest.mark.parametrize("workload", ["cbox", "grid"])
def test_full_size_properties(b2, workload):
    """BASELINE configs[1] (cbox 1024^2 x 256 spp) and configs[4] (9802 triangles x 1000 instances, 2048^2 x 64 spp)
    at FULL size, through size-independent properties: every camera sample lands on the film exactly once (box
    filter: the W channel of a pixel is spp up to the filter's epsilon-wide overlap, whatever the pass / lane-slot schedule), alpha <= weight, all
    values finite and non-negative, no sample rejected, and the image agrees with a low-resolution render of the
    same scene (pixel-block means) to Monte Carlo accuracy."""
    if workload == "cbox":
        W = H = 1024; spp = 256; sb = scenes.cbox(b2.Host(), W, H); small = scenes.cbox(b2.Host(), 64, 64)
    else:
        W = H = 2048; spp = 64; sb = scenes.instanced_grid(b2.Host(), W, H); small = scenes.instanced_grid(b2.Host(), 64, 64)
    integ = b2.PathIntegrator(max_depth=6, rr_depth=5, samples_per_pass=16, block_size=32)
    film = integ.render(b2.Scene(sb), spp)
    st = integ.stats
    assert st["n_samples"] == W * H * spp and st["n_invalid"] == 0
    assert np.isfinite(film).all() and (film[..., :3] >= -1e-5).all()
    # W = number of samples that landed on the pixel: spp, except that the box filter's radius is 0.5 + RayEpsilon
    # (box.cpp:24-29), so a sample within 1e-4 of a pixel edge lands on the neighbour, or is dropped at a block edge
    # (imageblock.cpp:162-169) -- like in the reference
    assert np.abs(film[..., 4] - spp).max() <= 4
    assert abs(film[..., 4].sum(dtype=np.float64) / (float(W) * H * spp) - 1.0) < 1e-4
    assert (film[..., 3] <= film[..., 4]).all() and (film[..., 3] >= 0).all()
    ref = b2.PathIntegrator(max_depth=6, rr_depth=5, block_size=32).render(b2.Scene(small), 1024)
    a = (film[..., :3] / spp).reshape(64, H // 64, 64, W // 64, 3).mean(axis=(1, 3))
    b = ref[..., :3] / 1024
    assert abs(a.mean() - b.mean()) < 0.02 * b.mean()
    assert np.linalg.norm(a - b) / np.linalg.norm(b) < 0.1


def test_develop(b2, oracle):
    sb, rp, ref = G.load("cbox_64_spp16", b2.Host())
    gpu, gst, sc = _gpu_render(b2, sb, rp)
    assert np.array_equal(sc.develop(gpu).view(np.uint32), oracle.develop(gpu).view(np.uint32))


def test_c1_reference_parity(b2, oracle):
    """BASELINE configs[0] (cbox 256x256, 16 spp, max_depth 6) and its well-posed variant
    (coplanar ceiling patch removed): GPU vs the reference's image."""
    sb, rp, ref = G.load("cbox_nopatch_256_spp16", b2.Host())
    gpu, gst, sc = _gpu_render(b2, sb, rp)
    s = G.parity_stats(sc.develop(gpu), ref)
    assert s["rel_l2"] < 1e-4, s           # north_star tolerance
    sb, rp, ref = G.load("cbox_256_spp16", b2.Host())
    gpu, gst, sc = _gpu_render(b2, sb, rp)
    s = G.parity_stats(sc.develop(gpu), ref)
    assert s["rel_l2_unflipped"] < 1e-5 and s["flipped_pixels"] < 200, s


def test_lane_slots_and_sharding_invariance(b2, oracle):
    """Seeds are a pure function of (block, pixel, pass): the image must not depend on the
    wavefront width, and rank-sharded films must sum to the single-GPU film."""
    sb, rp, ref = G.load("cbox_64_spp16", b2.Host())
    full, _, sc = _gpu_render(b2, sb, rp)
    rp2 = abi.render_params(rp.max_depth, rp.rr_depth, rp.spp, rp.samples_per_pass, rp.block_size, rp.seed, 0, 1, 512)
    small, _, _ = _gpu_render(b2, sb, rp2)
    assert np.array_equal(full.view(np.uint32), small.view(np.uint32))
    acc = np.zeros_like(full)
    for r in range(3):
        rpr = abi.render_params(rp.max_depth, rp.rr_depth, rp.spp, rp.samples_per_pass, rp.block_size, rp.seed, r, 3)
        part, st, _ = _gpu_render(b2, sb, rpr)
        acc += part
    assert np.array_equal(full.view(np.uint32), acc.view(np.uint32))   # disjoint pixels: exact


def test_errors(b2):
    sb = scenes.cbox(b2.Host(), 32, 32)
    sc = b2.Scene(sb)
    integ = b2.PathIntegrator(max_depth=6)
    with pytest.raises(b2.B2Error):          # integrator.cpp:62-64
        integ.samples_per_pass = 3
        integ.render(sc, spp=16)


@pytest.mark.parametrize("name", ["cbox_64_spp16", "cbox_64_spp16_ggx", "cbox_64_spp16_dielectric_lift", "cbox_64_spp16_env_area",
                                  "cbox_64_spp16_instances", "cbox_64_spp16_roughglass_ggx"])
def test_stage_shade_bitexact(b2, oracle, name):
    """b2_stage_shade (ONE k_shade launch on caller queues) vs the oracle's restatement of the loop body (path.cpp:124-208),
    chained over three bounces: next ray, throughput, result, RNG state, depth / flags and the emitter-sample queue must
    match bit for bit — the per-stage parity hook of the fused shading kernel."""
    sb, rp, _ = G.load(name, b2.Host())
    sc = b2.Scene(sb)
    desc = sb.desc()
    W, H = sb.film.crop_size[0], sb.film.crop_size[1]
    rng = np.random.default_rng(5)
    n = 6000
    px = (rng.integers(0, W, n) + sb.film.crop_offset[0]).astype(np.uint32); py = (rng.integers(0, H, n) + sb.film.crop_offset[1]).astype(np.uint32)
    ro, rd, state, pos = sc.sample_camera_rays(px, py, rng.integers(0, 2**40, n).astype(np.uint64), base_seed=3)
    tp = np.ones((n, 4), np.float32); res = np.zeros((n, 4), np.float32); res[:, 3] = 1.0
    df = np.ones(n, np.uint32)
    rp.max_depth = 6; rp.rr_depth = 2        # (Russian roulette from the third segment on)
    instanced = len(sb.instances) > 0
    for bounce in range(3):
        hit = np.zeros((len(ro), 4), np.float32); inst = np.full(len(ro), -1, np.int32)
        assert b2.lib().b2_stage_intersect(sc._h, len(ro), ro.ctypes.data, rd.ctypes.data, hit.ctypes.data, inst.ctypes.data) == 0
        keep = np.isfinite(hit[:, 0]) & (rd[:, 3] >= 0)
        ro, rd, hit, inst, tp, res, state, df = [np.ascontiguousarray(a[keep]) for a in (ro, rd, hit, inst, tp, res, state, df)]
        m = len(ro)
        assert m > 500
        g = [a.copy() for a in (ro, rd, tp, res, state, df)]
        c = [a.copy() for a in (ro, rd, tp, res, state, df)]
        gs = sc.shade(rp, g[0], g[1], hit, inst if instanced else None, g[2], g[3], g[4], g[5])
        ns = C.c_uint32(); so = np.zeros((m, 4), np.float32); sd = np.zeros((m, 4), np.float32); scn = np.zeros((m, 4), np.float32)
        oracle.lib().b2o_stage_shade(C.byref(desc), C.byref(rp), m, c[0].ctypes.data, c[1].ctypes.data, hit.ctypes.data,
                                     inst.ctypes.data if instanced else None, c[2].ctypes.data, c[3].ctypes.data, c[4].ctypes.data, c[5].ctypes.data,
                                     C.byref(ns), so.ctypes.data, sd.ctypes.data, scn.ctypes.data)
        assert np.array_equal(g[4], c[4]) and np.array_equal(g[5], c[5])              # RNG state, depth | flags
        parked = (g[5] & 0x2000) != 0
        assert np.array_equal(g[1].view(np.uint32), c[1].view(np.uint32))
        assert np.array_equal(g[0][~parked].view(np.uint32), c[0][~parked].view(np.uint32))
        assert np.array_equal(g[2].view(np.uint32), c[2].view(np.uint32)) and np.array_equal(g[3].view(np.uint32), c[3].view(np.uint32))
        k = ns.value
        assert len(gs[0]) == k and k > 100
        for a, b in zip(gs, (so[:k], sd[:k], scn[:k])):
            assert np.array_equal(a.view(np.uint32), b.view(np.uint32))
        ro, rd, tp, res, state, df = g


@pytest.mark.parametrize("rfilter", ["box", "gaussian"])
def test_stage_film_put(b2, oracle, rfilter):
    """b2_stage_film_put (ImageBlock::put on the device) vs the oracle's block put + merge: box filter with at most one
    sample per pixel is exact; the gaussian taps are fp32 atomics (order differs: rel 1e-6)."""
    sb = scenes.cbox(b2.Host(), 80, 48, rfilter=rfilter)
    sc = b2.Scene(sb)
    rng = np.random.default_rng(9)
    n = 3000 if rfilter == "gaussian" else 80 * 48
    if rfilter == "box":
        idx = rng.permutation(80 * 48)[:n]; px = (idx % 80).astype(np.uint32); py = (idx // 80).astype(np.uint32)
    else:
        px = rng.integers(0, 80, n).astype(np.uint32); py = rng.integers(0, 48, n).astype(np.uint32)
    pos = np.stack([px + rng.random(n, dtype=np.float32), py + rng.random(n, dtype=np.float32)], 1).astype(np.float32)
    pos[:7] = np.stack([px[:7], py[:7]], 1)                  # samples exactly on pixel corners: the x-1 / y-1 / dropped cases
    rgb = rng.random((n, 3), dtype=np.float32) * 4
    rgb[5] = np.nan; rgb[6] = -1.0                           # rejected by the validity rule (imageblock.cpp:85-109)
    alpha = (rng.random(n) < 0.8).astype(np.uint8)
    film, bad = sc.film_put(32, px, py, pos, rgb, alpha)
    desc = sb.desc(); ref = np.zeros_like(film); nb = C.c_uint32()
    oracle.lib().b2o_stage_film_put(C.byref(desc), 32, n, px.ctypes.data, py.ctypes.data, pos.ctypes.data, rgb.ctypes.data, alpha.ctypes.data,
                                    ref.ctypes.data, C.byref(nb))
    assert bad == nb.value == 2
    if rfilter == "box":
        assert np.array_equal(film.view(np.uint32), ref.view(np.uint32))
    else:
        assert oracle.rel_l2(film, ref) < 1e-6 and np.abs(film - ref).max() < 1e-4


def test_progressive_snapshots_and_cancel(b2):
    """b2_set_progress_callback: partial films while the render runs (SIGHUP / progress semantics of the reference), the
    final film is unaffected; a callback that asks to stop ends the render with B2_ERR_CANCELLED and leaves the samples
    accumulated so far on the film."""
    sb = scenes.cbox(b2.Host(), 512, 512)
    sc = b2.Scene(sb)
    integ = b2.PathIntegrator(max_depth=6, rr_depth=5, block_size=32)      # (one pass: the box-filter film is then deterministic bit for bit)
    ref = integ.render(sc, 64).copy()
    seen = []

    def on_progress(snap, done, total):
        w = float(snap[..., 4].sum(dtype=np.float64))
        seen.append((done, total, w, bool(np.isfinite(snap).all())))
        return False
    sc.set_progress_callback(on_progress, 0)
    film = integ.render(sc, 64)

    def same(a, b):        # bit-identical up to the few pixel-edge samples that land on a neighbouring lane's pixel (see test_render_bitexact_single_pass)
        x, y = a.view(np.int32).astype(np.int64).ravel(), b.view(np.int32).astype(np.int64).ravel()
        bad = np.nonzero(x != y)[0]
        return len(bad) <= 64 and (len(bad) == 0 or np.abs(x[bad] - y[bad]).max() <= 2)
    assert same(film, ref)
    assert len(seen) >= 2 and all(s[3] for s in seen) and all(s[1] == 512 * 512 * 64 for s in seen)
    assert all(a[0] <= b[0] for a, b in zip(seen, seen[1:])) and 0 < seen[-1][0] <= 512 * 512 * 64
    assert all(abs(s[2] - s[0]) <= 0.02 * 512 * 512 * 64 for s in seen)       # W of the snapshot ~ samples put so far (copy is not a cut)
    # develop a partial snapshot: XYZ / W is a valid estimate of the final image
    calls = []

    def stop_early(snap, done, total):
        calls.append(done)
        return done > total // 4
    sc.set_progress_callback(stop_early, 0)
    with pytest.raises(b2.B2Error) as e:
        integ.render(sc, 64)
    assert e.value.code == 3
    sc.set_progress_callback(None)
    out = np.empty((512, 512, 5), np.float32); st = abi.B2Stats()
    rp = integ.params(64)
    sc.set_progress_callback(stop_early, 0)
    rc = b2.lib().b2_render(sc._h, C.byref(rp), out.ctypes.data, C.byref(st))
    assert rc == 3
    wsum = float(out[..., 4].sum(dtype=np.float64))
    assert 0.2 * 512 * 512 * 64 < wsum < 512 * 512 * 64            # a partial film came back
    # developing it (XYZ / W per pixel) gives a valid estimate wherever samples have landed
    done = out[..., 4] >= 16
    assert done.sum() > 0.15 * 512 * 512
    part = (out[..., :3] / np.maximum(out[..., 4:5], 1))[done].mean(axis=0); full = (ref[..., :3] / ref[..., 4:5])[done].mean(axis=0)
    assert np.allclose(part, full, rtol=0.1), (part, full)
    sc.set_progress_callback(None)
    assert same(integ.render(sc, 64), ref)


def test_non_blocking_stream(b2, oracle):
    """b2_set_stream with a NON-BLOCKING stream (what PyTorch hands out): the stage entry points order their host <-> device
    copies on that stream, and a render on it gives the same film as on the legacy stream."""
    import torch
    sb, rp, _ = G.load("cbox_64_spp16", b2.Host())
    sc = b2.Scene(sb)
    ro, rd = _random_rays(sb, 50000, 7)
    hit0 = sc.ray_intersect_preliminary(ro, rd)
    occ0 = sc.ray_test(ro, rd)
    film0 = np.empty((64, 64, 5), np.float32); st = abi.B2Stats()
    assert b2.lib().b2_render(sc._h, C.byref(rp), film0.ctypes.data, C.byref(st)) == 0
    stream = torch.cuda.Stream()
    assert b2.lib().b2_set_stream(sc._h, C.c_void_p(stream.cuda_stream)) == 0
    for _ in range(3):
        hit1 = sc.ray_intersect_preliminary(ro, rd)
        occ1 = sc.ray_test(ro, rd)
        assert np.array_equal(hit0.view(np.uint32), hit1.view(np.uint32)) and np.array_equal(occ0, occ1)
    film1 = np.empty_like(film0)
    assert b2.lib().b2_render(sc._h, C.byref(rp), film1.ctypes.data, C.byref(st)) == 0
    assert np.array_equal(film0.view(np.uint32), film1.view(np.uint32))
    assert np.array_equal(sc.develop(film1).view(np.uint32), oracle.develop(film1).view(np.uint32))
    assert b2.lib().b2_set_stream(sc._h, None) == 0


def test_rank_without_blocks(b2):
    """world_size larger than the number of spiral blocks: the ranks that own no block return an empty film without
    launching, and the rank films still sum to the single-GPU film."""
    sb, rp, _ = G.load("cbox_64_spp16", b2.Host())          # 64 x 64, block size 32: 4 blocks
    full, _, sc = _gpu_render(b2, sb, rp)
    acc = np.zeros_like(full)
    for r in range(6):
        rpr = abi.render_params(rp.max_depth, rp.rr_depth, rp.spp, rp.samples_per_pass, rp.block_size, rp.seed, r, 6)
        part, st, _ = _gpu_render(b2, sb, rpr)
        if r >= 4:
            assert st["n_samples"] == 0 and st["n_launches"] == 0 and not part.any()
        acc += part
    assert np.array_equal(full.view(np.uint32), acc.view(np.uint32))


def test_stage_compact(b2):
    """b2_stage_compact (K5: the warp-ballot compaction of the live lanes) vs numpy, ragged sizes included."""
    sc = b2.Scene(scenes.cbox(b2.Host(), 32, 32))
    rng = np.random.default_rng(3)
    for n in (1, 31, 32, 33, 1000, 70001):
        maxt = np.where(rng.random(n) < 0.9, np.inf, -1.0).astype(np.float32)
        t = np.where(rng.random(n) < 0.7, rng.random(n) * 100, np.inf).astype(np.float32)
        depth = (rng.integers(1, 8, n) | (rng.integers(0, 4, n) << 8)).astype(np.uint32)       # depth | flag bits
        c, x = sc.compact(6, maxt, t, depth)
        has = maxt >= 0
        cont = has & np.isfinite(t) & ((depth & 0xff) < 6)
        assert np.array_equal(c, np.nonzero(cont)[0]) and np.array_equal(x, np.nonzero(has & ~cont)[0])
    c, x = sc.compact(-1, np.full(5, np.inf, np.float32), np.array([1, np.inf, 2, 3, np.inf], np.float32), np.full(5, 200, np.uint32))
    assert list(c) == [0, 2, 3] and list(x) == [1, 4]          # max_depth = -1: only misses end a path
