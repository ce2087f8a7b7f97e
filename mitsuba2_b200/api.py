"""Python binding of the C-ABI (include/b2pt.h) — thin ctypes layer, no compute here.

Mirrors the reference's host-side interface for the path (same names / meaning):
  Host                    Transform::look_at, parse_fov, perspective_projection, rfilters
  Scene(desc)             Scene::Scene + accel_init      (src/librender/scene.cpp:22-103)
  PathIntegrator.render   SamplingIntegrator::render      (src/librender/integrator.cpp:51)
  Scene.ray_intersect_preliminary / ray_test             (include/mitsuba/render/scene.h:38-128)
The CUDA library has no CPU fallback: loading works without a GPU (host helpers, symbol
checks), every compute call raises B2Error(B2_ERR_NO_DEVICE) when no device is present.
"""
import ctypes as C
import os
import subprocess

import numpy as np

from . import abi

_DIR = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("B2PT_LIB", os.path.join(_DIR, "libb2pt.so"))     # B2PT_LIB: A/B builds of the same ABI
_LIB = None

EXPORTS = [
    "b2_scene_create", "b2_scene_destroy", "b2_scene_update", "b2_last_error", "b2_device_count", "b2_render", "b2_render_device",
    "b2_cancel", "b2_develop", "b2_set_stream", "b2_stage_raygen", "b2_stage_intersect", "b2_stage_occluded",
    "b2_stage_si_fill", "b2_host_look_at", "b2_host_perspective_sensor", "b2_host_parse_fov", "b2_host_film_box",
    "b2_host_film_gaussian", "b2_host_spiral", "b2_host_build_bvh4", "b2_render_reduce", "b2_develop_device",
    "b2_host_bsdf_plastic", "b2_host_shape_sphere", "b2_host_shape_rectangle", "b2_stage_shade", "b2_stage_film_put",
    "b2_set_progress_callback", "b2_stage_compact",
]

PROGRESS_FN = C.CFUNCTYPE(C.c_int, C.c_void_p, C.POINTER(C.c_float), C.c_uint64, C.c_uint64)     # b2_progress_fn
REDUCE_FN = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p)     # b2_reduce_fn


class B2Error(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"[b2pt error {code}] {msg}")
        self.code = code


def build(force=False):
    """Compile libb2pt.so in-tree for sm_100a (nvcc cross-compiles without a GPU)."""
    src = os.path.join(_DIR, "csrc")
    if force:
        subprocess.run(["make", "-C", src, "clean"], check=True, capture_output=True)
    r = subprocess.run(["make", "-j4", "-C", src], capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("libb2pt build failed:\n" + r.stdout + r.stderr)
    return LIB_PATH


def lib():
    global _LIB
    if _LIB is None:
        if not os.path.exists(LIB_PATH):
            raise B2Error(-1, f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                              "(there is no CPU fallback)")
        L = C.CDLL(LIB_PATH)
        L.b2_last_error.restype = C.c_char_p
        L.b2_host_parse_fov.restype = C.c_float
        L.b2_host_parse_fov.argtypes = [C.c_float, C.c_char_p, C.c_int, C.c_int]
        L.b2_host_perspective_sensor.argtypes = [C.POINTER(C.c_float), C.c_float, C.c_float, C.c_float,
                                                 C.POINTER(abi.B2Film), C.POINTER(abi.B2Sensor)]
        L.b2_host_film_gaussian.argtypes = [C.c_int, C.c_int, C.c_float, C.POINTER(abi.B2Film)]
        L.b2_host_film_box.argtypes = [C.c_int, C.c_int, C.POINTER(abi.B2Film)]
        L.b2_host_spiral.argtypes = [C.c_int, C.c_int, C.c_int, C.POINTER(C.c_uint32), C.c_void_p]
        L.b2_scene_create.argtypes = [C.POINTER(abi.B2SceneDesc), C.POINTER(C.c_void_p)]
        L.b2_scene_destroy.argtypes = [C.c_void_p]
        L.b2_scene_update.argtypes = [C.c_void_p, C.POINTER(abi.B2SceneDesc)]
        L.b2_render.argtypes = [C.c_void_p, C.POINTER(abi.B2RenderParams), C.c_void_p, C.POINTER(abi.B2Stats)]
        L.b2_render_device.argtypes = [C.c_void_p, C.POINTER(abi.B2RenderParams), C.POINTER(C.c_void_p),
                                       C.POINTER(abi.B2Stats)]
        if hasattr(L, "b2_render_reduce"):      # (absent from older builds of the ABI loaded through B2PT_LIB for A/B runs)
            L.b2_render_reduce.argtypes = [C.c_void_p, C.POINTER(abi.B2RenderParams), REDUCE_FN, C.c_void_p, C.c_void_p,
                                           C.POINTER(abi.B2Stats)]
            L.b2_develop_device.argtypes = [C.c_void_p, C.c_void_p]
            L.b2_set_progress_callback.argtypes = [C.c_void_p, PROGRESS_FN, C.c_void_p, C.c_uint32]
            L.b2_stage_shade.argtypes = [C.c_void_p, C.POINTER(abi.B2RenderParams), C.c_uint32] + [C.c_void_p] * 8 + [C.POINTER(C.c_uint32)] + [C.c_void_p] * 3
            L.b2_stage_compact.argtypes = [C.c_void_p, C.c_int32, C.c_uint32, C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(C.c_uint32), C.c_void_p,
                                           C.POINTER(C.c_uint32), C.c_void_p]
            L.b2_stage_film_put.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32] + [C.c_void_p] * 6 + [C.POINTER(C.c_uint32)]
        L.b2_cancel.argtypes = [C.c_void_p]
        L.b2_develop.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        L.b2_set_stream.argtypes = [C.c_void_p, C.c_void_p]
        L.b2_stage_raygen.argtypes = [C.c_void_p, C.c_uint32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64,
                                      C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        L.b2_stage_intersect.argtypes = [C.c_void_p, C.c_uint32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        L.b2_stage_occluded.argtypes = [C.c_void_p, C.c_uint32, C.c_void_p, C.c_void_p, C.c_void_p]
        L.b2_stage_si_fill.argtypes = [C.c_void_p, C.c_uint32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        _LIB = L
    return _LIB


def _check(rc):
    if rc != 0:
        raise B2Error(rc, lib().b2_last_error().decode())


def device_count():
    return lib().b2_device_count()


class Host:
    """b2_host_*: C++ mirrors of the reference's host-side sensor / film / spiral setup."""

    def look_at(self, origin, target, up):
        out = (C.c_float * 16)()
        _check(lib().b2_host_look_at((C.c_float * 3)(*origin), (C.c_float * 3)(*target), (C.c_float * 3)(*up), out))
        return out

    def parse_fov(self, fov, axis, w, h):
        return lib().b2_host_parse_fov(fov, axis.encode(), w, h)

    def film_box(self, w, h):
        f = abi.B2Film(); _check(lib().b2_host_film_box(w, h, C.byref(f))); return f

    def film_gaussian(self, w, h, stddev=0.5):
        f = abi.B2Film(); _check(lib().b2_host_film_gaussian(w, h, stddev, C.byref(f))); return f

    def perspective_sensor(self, to_world, fov_x, near, far, film):
        s = abi.B2Sensor()
        _check(lib().b2_host_perspective_sensor(to_world, fov_x, near, far, C.byref(film), C.byref(s)))
        return s

    def spiral(self, w, h, bs):
        n = C.c_uint32()
        _check(lib().b2_host_spiral(w, h, bs, C.byref(n), None))
        order = np.zeros((n.value, 2), np.int32)
        _check(lib().b2_host_spiral(w, h, bs, C.byref(n), order.ctypes.data))
        return order


class Scene:
    """Device-resident scene (BVH + SoA triangles + materials). Wraps b2_scene_create."""

    def __init__(self, scene_buffers):
        self.sb = scene_buffers
        self._desc = scene_buffers.desc()
        self._h = C.c_void_p()
        _check(lib().b2_scene_create(C.byref(self._desc), C.byref(self._h)))
        self.height, self.width = scene_buffers.film.crop_size[1], scene_buffers.film.crop_size[0]

    def update(self, scene_buffers):
        """Scene::parameters_changed (scene.cpp:242-260): new geometry / parameters behind the same handle (b2_scene_update)."""
        desc = scene_buffers.desc()
        _check(lib().b2_scene_update(self._h, C.byref(desc)))
        self.sb, self._desc = scene_buffers, desc
        self.height, self.width = scene_buffers.film.crop_size[1], scene_buffers.film.crop_size[0]

    def close(self):
        if self._h:
            lib().b2_scene_destroy(self._h); self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # --- Scene::ray_intersect_preliminary / ray_test (stage API, host arrays) -----------------
    def ray_intersect_preliminary(self, ray_o, ray_d):
        ray_o = np.ascontiguousarray(ray_o, np.float32); ray_d = np.ascontiguousarray(ray_d, np.float32)
        n = len(ray_o)
        hit = np.zeros((n, 4), np.float32)
        _check(lib().b2_stage_intersect(self._h, n, ray_o.ctypes.data, ray_d.ctypes.data, hit.ctypes.data, None))
        return hit

    def ray_test(self, ray_o, ray_d):
        ray_o = np.ascontiguousarray(ray_o, np.float32); ray_d = np.ascontiguousarray(ray_d, np.float32)
        n = len(ray_o)
        occ = np.zeros(n, np.uint8)
        _check(lib().b2_stage_occluded(self._h, n, ray_o.ctypes.data, ray_d.ctypes.data, occ.ctypes.data))
        return occ.astype(bool)

    def compute_surface_interaction(self, ray_o, ray_d, hit):
        ray_o = np.ascontiguousarray(ray_o, np.float32); ray_d = np.ascontiguousarray(ray_d, np.float32)
        hit = np.ascontiguousarray(hit, np.float32)
        n = len(ray_o)
        si = np.zeros((n, 16), np.float32)
        _check(lib().b2_stage_si_fill(self._h, n, ray_o.ctypes.data, ray_d.ctypes.data, hit.ctypes.data, None, si.ctypes.data))
        return si

    def sample_camera_rays(self, px, py, seed_offset, base_seed=0):
        px = np.ascontiguousarray(px, np.uint32); py = np.ascontiguousarray(py, np.uint32)
        so = np.ascontiguousarray(seed_offset, np.uint64)
        n = len(px)
        ro = np.zeros((n, 4), np.float32); rd = np.zeros((n, 4), np.float32)
        rng = np.zeros(n, np.uint64); pos = np.zeros((n, 2), np.float32)
        _check(lib().b2_stage_raygen(self._h, n, px.ctypes.data, py.ctypes.data, so.ctypes.data, base_seed,
                                     ro.ctypes.data, rd.ctypes.data, rng.ctypes.data, pos.ctypes.data))
        return ro, rd, rng, pos

    def set_progress_callback(self, fn, min_interval_ms=0):
        """fn(xyzaw_snapshot [H, W, 5] (a view, valid during the call), samples_done, samples_total) -> truthy to cancel the
        render; None switches the snapshots off. Called from the rendering thread while the GPU keeps rendering."""
        if fn is None:
            self._progress = None
            _check(lib().b2_set_progress_callback(self._h, PROGRESS_FN(), None, 0))
            return

        def _cb(user, ptr, done, total):
            snap = np.ctypeslib.as_array(ptr, shape=(self.height, self.width, 5))
            return 1 if fn(snap, done, total) else 0
        self._progress = PROGRESS_FN(_cb)          # (keeps the thunk alive)
        _check(lib().b2_set_progress_callback(self._h, self._progress, None, int(min_interval_ms)))

    def shade(self, rp, ray_o, ray_d, hit, inst, throughput_eta, result_pdf, rng_state, depth_flags):
        """b2_stage_shade: one loop-body iteration of PathIntegrator::sample for n CONTINUE lanes. The state arrays are updated
        in place; returns the emitter-sample queue (sh_o, sh_d, sh_c) sorted by lane index."""
        n = len(ray_o)
        sh_o = np.zeros((n, 4), np.float32); sh_d = np.zeros((n, 4), np.float32); sh_c = np.zeros((n, 4), np.float32)
        ns = C.c_uint32()
        _check(lib().b2_stage_shade(self._h, C.byref(rp), n, ray_o.ctypes.data, ray_d.ctypes.data, hit.ctypes.data,
                                    inst.ctypes.data if inst is not None else None, throughput_eta.ctypes.data, result_pdf.ctypes.data,
                                    rng_state.ctypes.data, depth_flags.ctypes.data, C.byref(ns), sh_o.ctypes.data, sh_d.ctypes.data, sh_c.ctypes.data))
        k = ns.value
        order = np.argsort(sh_c[:k, 3].view(np.uint32), kind="stable")
        return sh_o[:k][order], sh_d[:k][order], sh_c[:k][order]

    def compact(self, max_depth, ray_maxt, hit_t, depth):
        """b2_stage_compact: -> (continue lanes, terminate lanes), each sorted ascending"""
        ray_maxt = np.ascontiguousarray(ray_maxt, np.float32); hit_t = np.ascontiguousarray(hit_t, np.float32); depth = np.ascontiguousarray(depth, np.uint32)
        n = len(hit_t)
        a = np.zeros(n, np.uint32); b = np.zeros(n, np.uint32); na = C.c_uint32(); nb = C.c_uint32()
        _check(lib().b2_stage_compact(self._h, max_depth, n, ray_maxt.ctypes.data, hit_t.ctypes.data, depth.ctypes.data, C.byref(na), a.ctypes.data,
                                      C.byref(nb), b.ctypes.data))
        return np.sort(a[:na.value]), np.sort(b[:nb.value])

    def film_put(self, block_size, px, py, pos, rgb, alpha):
        """b2_stage_film_put: ImageBlock::put of n samples -> (XYZAW film [H, W, 5], number of rejected samples)"""
        px = np.ascontiguousarray(px, np.uint32); py = np.ascontiguousarray(py, np.uint32)
        pos = np.ascontiguousarray(pos, np.float32); rgb = np.ascontiguousarray(rgb, np.float32); alpha = np.ascontiguousarray(alpha, np.uint8)
        film = np.zeros((self.height, self.width, 5), np.float32); bad = C.c_uint32()
        _check(lib().b2_stage_film_put(self._h, block_size, len(px), px.ctypes.data, py.ctypes.data, pos.ctypes.data, rgb.ctypes.data,
                                       alpha.ctypes.data, film.ctypes.data, C.byref(bad)))
        return film, bad.value

    def develop(self, xyzaw):
        x = np.ascontiguousarray(xyzaw, np.float32)
        rgb = np.zeros((self.height, self.width, 3), np.float32)
        _check(lib().b2_develop(self._h, x.ctypes.data, rgb.ctypes.data))
        return rgb


class PathIntegrator:
    """PathIntegrator (src/integrators/path.cpp) + SamplingIntegrator::render driver."""

    def __init__(self, max_depth=-1, rr_depth=5, samples_per_pass=0, block_size=32, sampler="independent", jitter=True):
        """sampler: "independent" (src/samplers/independent.cpp), "stratified" (src/samplers/stratified.cpp; spp must be a
        perfect square) "multijitter" (src/samplers/multijitter.cpp; spp = rx * ry with ry = floor(sqrt(spp))) or "ldsampler"
        (src/samplers/ldsampler.cpp; spp = 4^k)."""
        self.max_depth, self.rr_depth = max_depth, rr_depth
        self.samples_per_pass, self.block_size = samples_per_pass, block_size
        self.sampler, self.jitter = sampler, jitter
        self.stats = None

    def params(self, spp, seed=0, rank=0, world_size=1, lane_slots=0, flags=0):
        return abi.render_params(self.max_depth, self.rr_depth, spp, self.samples_per_pass, self.block_size, seed,
                                 rank, world_size, lane_slots, flags, self.sampler, self.jitter)

    def render(self, scene, spp, seed=0, rank=0, world_size=1, lane_slots=0, out=None, flags=0):
        """-> XYZAW film as float32 [H, W, 5] in HOST memory (copies inside the call)."""
        rp = self.params(spp, seed, rank, world_size, lane_slots, flags)
        if out is None:
            out = np.empty((scene.height, scene.width, 5), np.float32)
        st = abi.B2Stats()
        _check(lib().b2_render(scene._h, C.byref(rp), out.ctypes.data, C.byref(st)))
        self.stats = st.as_dict()
        return out

    def render_reduce(self, scene, spp, reduce, seed=0, rank=0, world_size=1, lane_slots=0, out=None, flags=0):
        """Multi-GPU render through the API: `reduce(ptr, n_floats, stream)` sums this rank's DEVICE film into rank 0's
        (e.g. torch.distributed.reduce on a tensor wrapping `ptr`); the film is copied to the host only on rank 0
        (-> [H, W, 5] array there, None elsewhere)."""
        rp = self.params(spp, seed, rank, world_size, lane_slots, flags)
        if out is None and rank == 0:
            out = np.empty((scene.height, scene.width, 5), np.float32)

        def _cb(user, ptr, n, stream):
            try:
                reduce(ptr, n, stream)
                return 0
            except Exception:      # noqa: BLE001 — reported through the status code
                import traceback
                traceback.print_exc()
                return 1
        cb = REDUCE_FN(_cb)
        st = abi.B2Stats()
        _check(lib().b2_render_reduce(scene._h, C.byref(rp), cb, None, out.ctypes.data if rank == 0 else None, C.byref(st)))
        self.stats = st.as_dict()
        return out if rank == 0 else None

    def render_device(self, scene, spp, seed=0, rank=0, world_size=1, lane_slots=0, flags=0):
        """-> device pointer (int) of the XYZAW film; film stays in HBM."""
        rp = self.params(spp, seed, rank, world_size, lane_slots, flags)
        st = abi.B2Stats(); ptr = C.c_void_p()
        _check(lib().b2_render_device(scene._h, C.byref(rp), C.byref(ptr), C.byref(st)))
        self.stats = st.as_dict()
        return ptr.value


class DirectIntegrator(PathIntegrator):
    """DirectIntegrator (src/integrators/direct.cpp:78-196): emitter_samples + bsdf_samples at the first hit, combined by
    multiple importance sampling. With the default shading_samples = 1 it consumes the sampler exactly like PathIntegrator with
    max_depth = 2 and uses the same MIS weights (pinned by tests/golden/ref_cbox_64_spp16_direct.npz), so that case runs as
    such; any other combination runs the direct mode of the wavefront (B2RenderParams.direct_*)."""

    def __init__(self, shading_samples=1, emitter_samples=None, bsdf_samples=None, samples_per_pass=0, block_size=32, sampler="independent", jitter=True):
        if shading_samples != 1 and (emitter_samples is not None or bsdf_samples is not None):
            raise ValueError("Cannot specify both 'shading_samples' and ('emitter_samples' and/or 'bsdf_samples').")      # direct.cpp:79-84
        self.emitter_samples = shading_samples if emitter_samples is None else emitter_samples
        self.bsdf_samples = shading_samples if bsdf_samples is None else bsdf_samples
        if self.emitter_samples + self.bsdf_samples == 0:
            raise ValueError("Must have at least 1 BSDF or emitter sample!")                                               # direct.cpp:93-94
        super().__init__(max_depth=2, rr_depth=5, samples_per_pass=samples_per_pass, block_size=block_size, sampler=sampler, jitter=jitter)

    def params(self, *args, **kw):
        rp = super().params(*args, **kw)
        if (self.emitter_samples, self.bsdf_samples) != (1, 1):
            rp.direct_emitter_samples, rp.direct_bsdf_samples = self.emitter_samples, self.bsdf_samples
        return rp
