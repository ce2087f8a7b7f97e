"""ctypes mirror of include/b2pt.h (the C-ABI structs).

Python is only the binding layer here: the structs below are byte-for-byte the C structs,
and the same definitions are used to feed the CUDA library (mitsuba2_b200/libb2pt.so) and,
from tests/ only, the CPU oracle (oracle/libb2oracle.so) with identical inputs.
"""
import ctypes as C

import numpy as np

B2_OK, B2_ERR_INVALID, B2_ERR_CUDA, B2_ERR_CANCELLED, B2_ERR_NO_DEVICE = range(5)
BSDF_DIFFUSE, BSDF_DIELECTRIC, BSDF_ROUGHCONDUCTOR, BSDF_NULL = range(4)
MF_BECKMANN, MF_GGX = 0, 1
EMITTER_AREA, EMITTER_ENVMAP, EMITTER_CONSTANT = range(3)
FILTER_RESOLUTION = 31
FLAG_STAGE_TIMING = 1


class B2Bsdf(C.Structure):
    _fields_ = [
        ("type", C.c_int32), ("distribution", C.c_int32), ("sample_visible", C.c_int32), ("_pad", C.c_int32),
        ("reflectance", C.c_float * 3), ("eta", C.c_float),
        ("specular_reflectance", C.c_float * 3), ("alpha_u", C.c_float),
        ("specular_transmittance", C.c_float * 3), ("alpha_v", C.c_float),
        ("cond_eta", C.c_float * 3), ("_pad1", C.c_float),
        ("cond_k", C.c_float * 3), ("_pad2", C.c_float),
    ]


class B2Emitter(C.Structure):
    _fields_ = [
        ("type", C.c_int32), ("mesh", C.c_int32),
        ("radiance", C.c_float * 3), ("scale", C.c_float),
        ("envmap", C.POINTER(C.c_float)),
        ("env_width", C.c_int32), ("env_height", C.c_int32),
        ("env_to_world", C.c_float * 16),
    ]


class B2Mesh(C.Structure):
    _fields_ = [
        ("n_vertices", C.c_uint32), ("n_faces", C.c_uint32),
        ("positions", C.POINTER(C.c_float)), ("normals", C.POINTER(C.c_float)),
        ("texcoords", C.POINTER(C.c_float)), ("faces", C.POINTER(C.c_uint32)),
        ("bsdf", C.c_int32), ("emitter", C.c_int32),
    ]


class B2Instance(C.Structure):
    _fields_ = [("group", C.c_int32), ("to_world", C.c_float * 16), ("to_object", C.c_float * 16)]


class B2Sensor(C.Structure):
    _fields_ = [
        ("sample_to_camera", C.c_float * 16), ("to_world", C.c_float * 16),
        ("near_clip", C.c_float), ("far_clip", C.c_float),
        ("principal_point_offset", C.c_float * 2),
    ]


class B2Film(C.Structure):
    _fields_ = [
        ("width", C.c_int32), ("height", C.c_int32),
        ("crop_offset", C.c_int32 * 2), ("crop_size", C.c_int32 * 2),
        ("filter_radius", C.c_float), ("filter_scale", C.c_float),
        ("filter_lut", C.c_float * (FILTER_RESOLUTION + 1)),
        ("border", C.c_int32),
    ]


class B2SceneDesc(C.Structure):
    _fields_ = [
        ("n_meshes", C.c_uint32), ("meshes", C.POINTER(B2Mesh)),
        ("n_bsdfs", C.c_uint32), ("bsdfs", C.POINTER(B2Bsdf)),
        ("n_emitters", C.c_uint32), ("emitters", C.POINTER(B2Emitter)),
        ("n_groups", C.c_uint32), ("group_first", C.POINTER(C.c_uint32)),
        ("n_instances", C.c_uint32), ("instances", C.POINTER(B2Instance)),
        ("n_top_meshes", C.c_uint32),
        ("sensor", B2Sensor), ("film", B2Film),
    ]


class B2RenderParams(C.Structure):
    _fields_ = [
        ("max_depth", C.c_int32), ("rr_depth", C.c_int32),
        ("spp", C.c_uint32), ("samples_per_pass", C.c_uint32),
        ("block_size", C.c_uint32), ("lane_slots", C.c_uint32),
        ("seed", C.c_uint64),
        ("rank", C.c_int32), ("world_size", C.c_int32),
        ("flags", C.c_uint32), ("_pad", C.c_uint32),
    ]


class B2Stats(C.Structure):
    _fields_ = [
        ("n_samples", C.c_uint64), ("n_segments", C.c_uint64), ("n_shadow_rays", C.c_uint64),
        ("n_invalid", C.c_uint64), ("n_iterations", C.c_uint64), ("n_launches", C.c_uint64),
        ("ms_total", C.c_float),
        ("ms_regen", C.c_float), ("ms_intersect", C.c_float), ("ms_shade", C.c_float), ("ms_shadow", C.c_float),
        ("bvh_nodes", C.c_uint64), ("bvh_tris", C.c_uint64),
        ("n_put_shadow", C.c_uint64),
    ]

    def as_dict(self):
        return {n: getattr(self, n) for n, _ in self._fields_}


def fptr(a):
    return a.ctypes.data_as(C.POINTER(C.c_float))


def uptr(a):
    return a.ctypes.data_as(C.POINTER(C.c_uint32))


def render_params(max_depth=6, rr_depth=5, spp=16, samples_per_pass=0, block_size=32, seed=0,
                  rank=0, world_size=1, lane_slots=0, flags=0):
    return B2RenderParams(max_depth, rr_depth, spp, samples_per_pass, block_size, lane_slots, seed, rank, world_size,
                          flags, 0)


class SceneBuffers:
    """Owns the numpy arrays behind a B2SceneDesc (keeps them alive) — the host-side
    flattening of Mitsuba's Scene object graph (scene.cpp:23-59, mesh.h:344-361)."""

    def __init__(self):
        self.meshes = []      # dicts: positions (n,3) f32, normals or None, faces (m,3) u32, bsdf, emitter
        self.bsdfs = []       # B2Bsdf
        self.emitters = []    # B2Emitter
        self.sensor = B2Sensor()
        self.film = B2Film()
        self._keep = []

    # -- BSDF / emitter constructors with the reference plugins' defaults --------------------
    def add_diffuse(self, reflectance=(0.5, 0.5, 0.5)):
        b = B2Bsdf(); b.type = BSDF_DIFFUSE
        b.reflectance[:] = reflectance
        b.specular_reflectance[:] = (1, 1, 1); b.specular_transmittance[:] = (1, 1, 1)
        self.bsdfs.append(b); return len(self.bsdfs) - 1

    def add_dielectric(self, int_ior=1.5046, ext_ior=1.000277, specular_reflectance=(1, 1, 1),
                       specular_transmittance=(1, 1, 1)):
        b = B2Bsdf(); b.type = BSDF_DIELECTRIC
        b.eta = float(np.float32(int_ior) / np.float32(ext_ior))     # dielectric.cpp:186
        b.specular_reflectance[:] = specular_reflectance
        b.specular_transmittance[:] = specular_transmittance
        self.bsdfs.append(b); return len(self.bsdfs) - 1

    def add_roughconductor(self, alpha=0.1, distribution="beckmann", eta=(0, 0, 0), k=(1, 1, 1),
                           specular_reflectance=(1, 1, 1), sample_visible=True, alpha_v=None):
        b = B2Bsdf(); b.type = BSDF_ROUGHCONDUCTOR
        b.distribution = MF_GGX if distribution == "ggx" else MF_BECKMANN
        b.sample_visible = 1 if sample_visible else 0
        b.alpha_u = alpha; b.alpha_v = alpha if alpha_v is None else alpha_v
        b.cond_eta[:] = eta; b.cond_k[:] = k
        b.specular_reflectance[:] = specular_reflectance; b.specular_transmittance[:] = (1, 1, 1)
        self.bsdfs.append(b); return len(self.bsdfs) - 1

    def add_area_emitter(self, mesh_index, radiance):
        e = B2Emitter(); e.type = EMITTER_AREA; e.mesh = mesh_index
        e.radiance[:] = radiance; e.scale = 1.0
        self.emitters.append(e)
        self.meshes[mesh_index]["emitter"] = len(self.emitters) - 1
        return len(self.emitters) - 1

    def add_mesh(self, positions, faces, normals=None, bsdf=0):
        m = dict(positions=np.ascontiguousarray(positions, np.float32).reshape(-1, 3),
                 faces=np.ascontiguousarray(faces, np.uint32).reshape(-1, 3),
                 normals=None if normals is None else np.ascontiguousarray(normals, np.float32).reshape(-1, 3),
                 bsdf=bsdf, emitter=-1)
        self.meshes.append(m); return len(self.meshes) - 1

    @property
    def n_triangles(self):
        return int(sum(len(m["faces"]) for m in self.meshes))

    def desc(self):
        d = B2SceneDesc()
        ms = (B2Mesh * max(1, len(self.meshes)))()
        for i, m in enumerate(self.meshes):
            ms[i].n_vertices = len(m["positions"]); ms[i].n_faces = len(m["faces"])
            ms[i].positions = fptr(m["positions"])
            ms[i].normals = fptr(m["normals"]) if m["normals"] is not None else None
            ms[i].texcoords = None
            ms[i].faces = uptr(m["faces"])
            ms[i].bsdf = m["bsdf"]; ms[i].emitter = m["emitter"]
        bs = (B2Bsdf * max(1, len(self.bsdfs)))(*self.bsdfs)
        es = (B2Emitter * max(1, len(self.emitters)))(*self.emitters)
        d.n_meshes = len(self.meshes); d.meshes = ms
        d.n_bsdfs = len(self.bsdfs); d.bsdfs = bs
        d.n_emitters = len(self.emitters); d.emitters = es
        d.n_groups = 0; d.group_first = None; d.n_instances = 0; d.instances = None
        d.n_top_meshes = len(self.meshes)
        d.sensor = self.sensor; d.film = self.film
        self._keep = [ms, bs, es]
        return d
