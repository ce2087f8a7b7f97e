"""CPU-only known-answer / property tests of the oracle's building blocks (oracle/b2o.c), the way the reference tests
its own: src/textures/tests, src/samplers/tests/test_*.py (stratification), src/bsdfs/tests (sample == eval / pdf)."""
import ctypes as C

import numpy as np
import pytest

from mitsuba2_b200 import abi


# ---- textures: oracle vs an independent numpy restatement of bitmap.cpp:386-474 / checkerboard.cpp:46-63 -------------
def _np_wrap(v, res, mode):
    if mode == "clamp":
        return np.clip(v, 0, res - 1)
    div = np.trunc(v / res).astype(np.int64)            # truncating division, like enoki::divisor
    mod = v - div * res
    mod = np.where(mod < 0, mod + res, mod)
    if mode == "mirror":
        keep = ((div & 1) == 0) ^ (v < 0)
        mod = np.where(keep, mod, res - 1 - mod)
    return mod


def _np_bitmap(data, uv, filt, wrap):
    h, w = data.shape[:2]
    d = data if data.ndim == 3 else np.repeat(data[..., None], 3, axis=2)
    u, v = uv[:, 0].astype(np.float64), uv[:, 1].astype(np.float64)
    if filt == "nearest":
        x = _np_wrap(np.floor(u * w).astype(np.int64), w, wrap); y = _np_wrap(np.floor(v * h).astype(np.int64), h, wrap)
        return d[y, x]
    u = u * w - 0.5; v = v * h - 0.5
    ix, iy = np.floor(u).astype(np.int64), np.floor(v).astype(np.int64)
    w1x, w1y = (u - ix)[:, None], (v - iy)[:, None]
    x0, x1, y0, y1 = _np_wrap(ix, w, wrap), _np_wrap(ix + 1, w, wrap), _np_wrap(iy, h, wrap), _np_wrap(iy + 1, h, wrap)
    v0 = (1 - w1x) * d[y0, x0] + w1x * d[y0, x1]
    v1 = (1 - w1x) * d[y1, x0] + w1x * d[y1, x1]
    return (1 - w1y) * v0 + w1y * v1


@pytest.mark.parametrize("filt", ["nearest", "bilinear"])
@pytest.mark.parametrize("wrap", ["repeat", "mirror", "clamp"])
@pytest.mark.parametrize("channels", [1, 3])
def test_bitmap_lookup(oracle, filt, wrap, channels):
    rng = np.random.default_rng(7)
    data = rng.uniform(0, 1, (5, 7) if channels == 1 else (5, 7, 3)).astype(np.float32)
    sb = abi.SceneBuffers()
    sb.add_bitmap(data, filt, wrap)
    uv = rng.uniform(-3.2, 4.1, (4000, 2)).astype(np.float32)       # far outside [0,1]^2, both signs
    out = np.zeros((len(uv), 3), np.float32)
    L = oracle.lib()
    L.b2o_texture_eval.argtypes = [C.POINTER(abi.B2Texture), C.c_uint32, C.c_void_p, C.c_void_p]
    L.b2o_texture_eval(C.byref(sb.textures[0]), len(uv), uv.ctypes.data, out.ctypes.data)
    ref = _np_bitmap(data, uv, filt, wrap)
    if filt == "nearest":       # exact unless the scaled coordinate sits within rounding of a texel edge
        frac = np.minimum(np.abs(uv[:, 0] * 7 - np.round(uv[:, 0] * 7)), np.abs(uv[:, 1] * 5 - np.round(uv[:, 1] * 5)))
        ok = frac > 1e-4
        assert np.array_equal(out[ok], ref[ok].astype(np.float32))
    else:
        assert np.abs(out - ref).max() < 2e-5


def test_checkerboard_lookup(oracle):
    sb = abi.SceneBuffers()
    sb.add_checkerboard((0.9, 0.8, 0.7), (0.1, 0.2, 0.3), [[4, 0, 0.25], [0, 2, -0.5], [0, 0, 1]])
    rng = np.random.default_rng(3)
    uv = rng.uniform(-2, 2, (4000, 2)).astype(np.float32)
    out = np.zeros((len(uv), 3), np.float32)
    L = oracle.lib()
    L.b2o_texture_eval.argtypes = [C.POINTER(abi.B2Texture), C.c_uint32, C.c_void_p, C.c_void_p]
    L.b2o_texture_eval(C.byref(sb.textures[0]), len(uv), uv.ctypes.data, out.ctypes.data)
    u = uv[:, 0].astype(np.float64) * 4 + 0.25; v = uv[:, 1].astype(np.float64) * 2 - 0.5
    fu, fv = u - np.floor(u), v - np.floor(v)
    safe = (np.abs(fu - 0.5) > 1e-5) & (np.abs(fv - 0.5) > 1e-5) & (fu > 1e-5) & (fv > 1e-5)
    same = (fu > 0.5) == (fv > 0.5)
    ref = np.where(same[:, None], np.float32([0.9, 0.8, 0.7]), np.float32([0.1, 0.2, 0.3]))
    assert np.array_equal(out[safe], ref[safe])


# ---- samplers: the stratification properties the reference's sampler tests check ---------------------------------------
def _draws(oracle, sampler, spp, pattern, seed_offset=12345, jitter=True, seed=0):
    rp = abi.render_params(spp=spp, sampler=sampler, jitter=jitter, seed=seed)
    n = sum(2 if c == "2" else 1 for c in pattern)
    out = np.zeros((spp, n), np.float32)
    L = oracle.lib()
    L.b2o_sampler_draws.argtypes = [C.POINTER(abi.B2RenderParams), C.c_uint64, C.c_uint32, C.c_char_p, C.c_void_p]
    L.b2o_sampler_draws(C.byref(rp), seed_offset, spp, pattern.encode(), out.ctypes.data)
    return out


@pytest.mark.parametrize("sampler,spp", [("stratified", 16), ("stratified", 64), ("multijitter", 16), ("multijitter", 8),
                                         ("multijitter", 30), ("ldsampler", 16), ("ldsampler", 64)])
def test_sampler_stratification(oracle, sampler, spp):
    d = _draws(oracle, sampler, spp, "2121")
    assert (d >= 0).all() and (d <= 1).all()            # (sobol_2 may return exactly 1.0 in fp32, like the reference)
    # every 1-D dimension: exactly one sample per 1/spp stratum
    for col in (2, 5):
        assert sorted(np.minimum((d[:, col] * spp).astype(int), spp - 1)) == list(range(spp))
    # every 2-D dimension: exactly one sample per cell of the rx x ry grid
    if sampler == "multijitter":
        ry = int(np.sqrt(np.float32(spp))); rx = (spp + ry - 1) // ry
    else:
        rx = ry = int(round(np.sqrt(spp)))
    for c0 in (0, 3):
        cells = np.minimum((d[:, c0] * rx).astype(int), rx - 1) + rx * np.minimum((d[:, c0 + 1] * ry).astype(int), ry - 1)
        assert sorted(cells) == list(range(spp)), (sampler, spp)
    if sampler == "multijitter":                         # ... and the N-rooks property of correlated multi-jittering
        for c0 in (0, 3):
            assert sorted(np.minimum((d[:, c0] * spp).astype(int), spp - 1)) == list(range(spp))
            assert sorted(np.minimum((d[:, c0 + 1] * spp).astype(int), spp - 1)) == list(range(spp))


def test_sampler_sequences_depend_on_pixel_and_seed(oracle):
    a = _draws(oracle, "stratified", 16, "21", seed_offset=1)
    assert not np.array_equal(a, _draws(oracle, "stratified", 16, "21", seed_offset=2))
    assert not np.array_equal(a, _draws(oracle, "stratified", 16, "21", seed_offset=1, seed=5))
    assert np.array_equal(a, _draws(oracle, "stratified", 16, "21", seed_offset=1))
    nj = _draws(oracle, "stratified", 16, "21", jitter=False)
    assert np.allclose((nj[:, 2] * 16) % 1.0, 0.5)      # stratum centres without jitter


def test_independent_sampler_is_pcg32(oracle):
    """IndependentSampler == consecutive PCG32 floats (independent.cpp:73-82) — pinned by enoki's known first outputs."""
    d = _draws(oracle, "independent", 4, "21", seed_offset=0).ravel()
    L = oracle.lib()
    st = (C.c_uint64 * 2)()
    L.b2o_pcg32_seed(0, 0, st)
    ref = np.array([L.b2o_pcg32_next_float(st) for _ in range(12)], np.float32)
    assert np.array_equal(d, ref)


# ---- BSDFs: sample() weight == eval / pdf for the smooth lobes, delta lobes carry no density ---------------------------
def _bsdf_records():
    sb = abi.SceneBuffers()
    sb.add_diffuse((0.2, 0.5, 0.8))
    sb.add_roughconductor(alpha=0.3, distribution="ggx", eta=(0.2, 0.9, 1.1), k=(3.9, 2.4, 2.2))
    sb.add_roughconductor(alpha=0.2, distribution="beckmann")
    sb.add_plastic((0.3, 0.4, 0.5), 1.6, 1.0)
    sb.add_conductor((0.2, 0.9, 1.1), (3.9, 2.4, 2.2))
    sb.add_dielectric(1.5, 1.0)
    return sb.bsdfs


@pytest.mark.parametrize("idx", range(6))
def test_bsdf_sample_weight_is_eval_over_pdf(oracle, idx):
    L = oracle.lib()
    L.b2o_bsdf_eval.argtypes = [C.POINTER(abi.B2Bsdf), C.POINTER(C.c_float), C.POINTER(C.c_float), C.POINTER(C.c_float), C.POINTER(C.c_float)]
    b = _bsdf_records()[idx]
    rng = np.random.default_rng(idx)
    n_smooth = 0
    for _ in range(400):
        wi = rng.normal(size=3); wi[2] = abs(wi[2]) + 0.05; wi /= np.linalg.norm(wi)
        wi_c = (C.c_float * 3)(*wi)
        s2 = (C.c_float * 2)(*rng.uniform(0, 1, 2))
        wo = (C.c_float * 3)(); w = (C.c_float * 3)(); pdf = C.c_float(); eta = C.c_float(); delta = C.c_int()
        L.b2o_bsdf_sample(C.byref(b), wi_c, float(rng.uniform()), s2, wo, C.byref(pdf), C.byref(eta), C.byref(delta), w)
        ev = (C.c_float * 3)(); p2 = C.c_float()
        L.b2o_bsdf_eval(C.byref(b), wi_c, wo, ev, C.byref(p2))
        weight = np.array(w[:])
        assert np.isfinite(weight).all() and (weight >= 0).all()
        if b.type != abi.BSDF_PLASTIC:       # single-lobe BSDFs: the weight is an albedo (plastic divides by a lobe-selection probability)
            assert weight.max() <= 1.0 + 1e-4
        if delta.value:
            if b.type != abi.BSDF_PLASTIC:        # (plastic: eval / pdf describe its diffuse lobe in every direction)
                assert p2.value == 0.0 and not any(ev[:])        # a delta lobe has no density (conductor.cpp:242-250)
            continue
        if pdf.value <= 0 or not weight.any():
            continue
        n_smooth += 1
        if b.type == abi.BSDF_PLASTIC:        # sample(): pdf of the chosen (diffuse) lobe == pdf(); weight == eval / pdf
            assert abs(p2.value - pdf.value) <= 1e-5 * pdf.value
        else:
            assert abs(p2.value - pdf.value) <= 2e-4 * pdf.value
        assert np.allclose(np.array(ev[:]) / pdf.value, weight, rtol=5e-4, atol=1e-6)
    if b.type in (abi.BSDF_DIFFUSE, abi.BSDF_ROUGHCONDUCTOR, abi.BSDF_PLASTIC):
        assert n_smooth > 100
