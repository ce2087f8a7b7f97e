"""ctypes loader for the CPU oracle (oracle/libb2oracle.so). TEST INFRASTRUCTURE: imported
only from tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / reference legs."""
import ctypes as C
import os
import subprocess

import numpy as np

from mitsuba2_b200 import abi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_LIB = None


def build():
    subprocess.run(["make", "-C", os.path.join(ROOT, "oracle")], check=True, capture_output=True)


def lib():
    global _LIB
    if _LIB is None:
        path = os.path.join(ROOT, "oracle", "libb2oracle.so")
        src = os.path.join(ROOT, "oracle", "b2o.c")
        if not os.path.exists(path) or os.path.getmtime(path) < os.path.getmtime(src):
            build()
        path = os.environ.get("B2O_LIB", path)      # (scripts/refrcp_analysis.py: the analysis build of the oracle)
        L = C.CDLL(path)
        L.b2o_parse_fov.restype = C.c_float
        L.b2o_parse_fov.argtypes = [C.c_float, C.c_char_p, C.c_int, C.c_int]
        L.b2o_perspective_sensor.argtypes = [C.POINTER(C.c_float), C.c_float, C.c_float, C.c_float,
                                             C.POINTER(abi.B2Film), C.POINTER(abi.B2Sensor)]
        L.b2o_film_gaussian.argtypes = [C.c_int, C.c_int, C.c_float, C.POINTER(abi.B2Film)]
        L.b2o_pcg32_next_float.restype = C.c_float
        L.b2o_pcg32_next_uint32.restype = C.c_uint32
        L.b2o_pcg32_seed.argtypes = [C.c_uint64, C.c_uint64, C.POINTER(C.c_uint64)]
        L.b2o_fresnel_conductor.restype = C.c_float
        L.b2o_fresnel_conductor.argtypes = [C.c_float, C.c_float, C.c_float]
        L.b2o_fresnel.argtypes = [C.c_float, C.c_float, C.POINTER(C.c_float)]
        L.b2o_math_probe.argtypes = [C.c_float, C.POINTER(C.c_float)]
        L.b2o_warp_cosine_hemisphere.argtypes = [C.c_float, C.c_float, C.POINTER(C.c_float)]
        L.b2o_bsdf_sample.argtypes = [C.POINTER(abi.B2Bsdf), C.POINTER(C.c_float), C.c_float, C.POINTER(C.c_float),
                                      C.POINTER(C.c_float), C.POINTER(C.c_float), C.POINTER(C.c_float),
                                      C.POINTER(C.c_int), C.POINTER(C.c_float)]
        L.b2o_stage_raygen.argtypes = [C.POINTER(abi.B2SceneDesc), C.c_uint32, C.c_void_p, C.c_void_p, C.c_void_p,
                                       C.c_uint64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        L.b2o_stage_intersect.argtypes = [C.POINTER(abi.B2SceneDesc), C.c_uint32, C.c_void_p, C.c_void_p, C.c_void_p,
                                          C.c_void_p, C.c_int]
        L.b2o_stage_occluded.argtypes = [C.POINTER(abi.B2SceneDesc), C.c_uint32, C.c_void_p, C.c_void_p, C.c_void_p,
                                         C.c_int]
        L.b2o_stage_si_fill.argtypes = [C.POINTER(abi.B2SceneDesc), C.c_uint32, C.c_void_p, C.c_void_p, C.c_void_p,
                                        C.c_void_p, C.c_void_p]
        L.b2o_stage_shade.argtypes = [C.POINTER(abi.B2SceneDesc), C.POINTER(abi.B2RenderParams), C.c_uint32] + [C.c_void_p] * 8 + \
                                     [C.POINTER(C.c_uint32)] + [C.c_void_p] * 3
        L.b2o_stage_film_put.argtypes = [C.POINTER(abi.B2SceneDesc), C.c_uint32, C.c_uint32] + [C.c_void_p] * 6 + [C.POINTER(C.c_uint32)]
        _LIB = L
    return _LIB


class OracleHost:
    """The b2o_* restatements of the reference's host-side setup code."""

    def look_at(self, origin, target, up):
        out = (C.c_float * 16)()
        lib().b2o_look_at((C.c_float * 3)(*origin), (C.c_float * 3)(*target), (C.c_float * 3)(*up), out)
        return out

    def parse_fov(self, fov, axis, w, h):
        return lib().b2o_parse_fov(fov, axis.encode(), w, h)

    def film_box(self, w, h):
        f = abi.B2Film(); lib().b2o_film_box(w, h, C.byref(f)); return f

    def film_gaussian(self, w, h, stddev=0.5):
        f = abi.B2Film(); lib().b2o_film_gaussian(w, h, stddev, C.byref(f)); return f

    def perspective_sensor(self, to_world, fov_x, near, far, film):
        s = abi.B2Sensor()
        lib().b2o_perspective_sensor(to_world, fov_x, near, far, C.byref(film), C.byref(s))
        return s


def render(sb, params, threads=None, brute=False):
    """Oracle render -> (xyzaw float32 [H,W,5], stats dict)."""
    d = sb.desc()
    H, W = sb.film.crop_size[1], sb.film.crop_size[0]
    out = np.zeros((H, W, 5), np.float32)
    st = abi.B2Stats()
    threads = threads or os.cpu_count() or 1
    rc = lib().b2o_render(C.byref(d), C.byref(params), out.ctypes.data_as(C.POINTER(C.c_float)), C.byref(st),
                          int(threads), int(brute))
    if rc != 0:
        raise RuntimeError(f"b2o_render failed: {rc}")
    return out, st.as_dict()


def develop(xyzaw):
    H, W = xyzaw.shape[:2]
    rgb = np.zeros((H, W, 3), np.float32)
    x = np.ascontiguousarray(xyzaw, np.float32)
    lib().b2o_develop(x.ctypes.data_as(C.POINTER(C.c_float)), H * W, rgb.ctypes.data_as(C.POINTER(C.c_float)))
    return rgb


def rel_l2(a, b):
    a = np.asarray(a, np.float64); b = np.asarray(b, np.float64)
    return float(np.linalg.norm(a - b) / np.linalg.norm(b))
