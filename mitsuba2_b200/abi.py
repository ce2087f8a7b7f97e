"""ctypes mirror of include/b2pt.h (the C-ABI structs).

Python is only the binding layer here: the structs below are byte-for-byte the C structs,
and the same definitions are used to feed the CUDA library (mitsuba2_b200/libb2pt.so) and,
from tests/ only, the CPU oracle (oracle/libb2oracle.so) with identical inputs.
"""
import ctypes as C

import numpy as np

B2_OK, B2_ERR_INVALID, B2_ERR_CUDA, B2_ERR_CANCELLED, B2_ERR_NO_DEVICE = range(5)
BSDF_DIFFUSE, BSDF_DIELECTRIC, BSDF_ROUGHCONDUCTOR, BSDF_NULL, BSDF_CONDUCTOR, BSDF_PLASTIC, BSDF_ROUGHDIELECTRIC = range(7)
MF_BECKMANN, MF_GGX = 0, 1
EMITTER_AREA, EMITTER_ENVMAP, EMITTER_CONSTANT, EMITTER_POINT, EMITTER_SPOT, EMITTER_DIRECTIONAL = range(6)
FILTER_RESOLUTION = 31
FLAG_STAGE_TIMING = 1
FLAG_SINGLE_SET = 2
FLAG_COUNT_BVH = 4
FLAG_PERSISTENT_TRACE = 8
FLAG_RESUME_TRACE = 16


def flag_yield_threshold(n):
    """B2_FLAG_YIELD_THRESHOLD(n)"""
    return (n & 0x1f) << 8


def flag_blocks(trace_sel=0, shade_sel=0):
    """B2_FLAG_TRACE_BLOCK | B2_FLAG_SHADE_BLOCK: CTA size of the two stage kernels: 0 = auto, 1 = 256, 2 = 128, 3 = 64 threads (tuning)"""
    return ((trace_sel & 3) << 16) | ((shade_sel & 3) << 18)

BUILD_HOST_SAH, BUILD_DEVICE_LBVH = 0, 1


class B2Bsdf(C.Structure):
    _fields_ = [
        ("type", C.c_int32), ("distribution", C.c_int32), ("sample_visible", C.c_int32), ("reflectance_texture", C.c_int32),
        ("reflectance", C.c_float * 3), ("eta", C.c_float),
        ("specular_reflectance", C.c_float * 3), ("alpha_u", C.c_float),
        ("specular_transmittance", C.c_float * 3), ("alpha_v", C.c_float),
        ("cond_eta", C.c_float * 3), ("twosided_back", C.c_int32),
        ("cond_k", C.c_float * 3), ("_pad2", C.c_float),
        ("fdr_int", C.c_float), ("specular_sampling_weight", C.c_float), ("inv_eta_2", C.c_float), ("nonlinear", C.c_int32),
    ]


TEX_CHECKERBOARD, TEX_BITMAP = range(2)
TEXFILTER_NEAREST, TEXFILTER_BILINEAR = range(2)
WRAP_REPEAT, WRAP_MIRROR, WRAP_CLAMP = range(3)


class B2Texture(C.Structure):
    _fields_ = [
        ("type", C.c_int32), ("filter", C.c_int32), ("wrap", C.c_int32),
        ("width", C.c_int32), ("height", C.c_int32), ("channels", C.c_int32),
        ("to_uv", C.c_float * 9), ("color0", C.c_float * 3), ("color1", C.c_float * 3),
        ("data", C.POINTER(C.c_float)),
    ]


class B2Emitter(C.Structure):
    _fields_ = [
        ("type", C.c_int32), ("mesh", C.c_int32),
        ("radiance", C.c_float * 3), ("scale", C.c_float),
        ("envmap", C.POINTER(C.c_float)),
        ("env_width", C.c_int32), ("env_height", C.c_int32),
        ("env_to_world", C.c_float * 16), ("env_to_object", C.c_float * 16),
        ("shape", C.c_int32), ("_pad", C.c_int32),
        ("cutoff_angle", C.c_float), ("beam_width", C.c_float),
    ]


class B2Mesh(C.Structure):
    _fields_ = [
        ("n_vertices", C.c_uint32), ("n_faces", C.c_uint32),
        ("positions", C.POINTER(C.c_float)), ("normals", C.POINTER(C.c_float)),
        ("texcoords", C.POINTER(C.c_float)), ("faces", C.POINTER(C.c_uint32)),
        ("bsdf", C.c_int32), ("emitter", C.c_int32),
    ]


SHAPE_RECTANGLE, SHAPE_SPHERE = range(2)


class B2Shape(C.Structure):
    _fields_ = [
        ("type", C.c_int32), ("bsdf", C.c_int32), ("emitter", C.c_int32), ("flip_normals", C.c_int32),
        ("to_world", C.c_float * 16), ("to_object", C.c_float * 16),
        ("center", C.c_float * 3), ("radius", C.c_float),
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
        ("n_textures", C.c_uint32), ("textures", C.POINTER(B2Texture)),
        ("n_shapes", C.c_uint32), ("shapes", C.POINTER(B2Shape)),
        ("build_flags", C.c_uint32),
    ]


class B2RenderParams(C.Structure):
    _fields_ = [
        ("max_depth", C.c_int32), ("rr_depth", C.c_int32),
        ("spp", C.c_uint32), ("samples_per_pass", C.c_uint32),
        ("block_size", C.c_uint32), ("lane_slots", C.c_uint32),
        ("seed", C.c_uint64),
        ("rank", C.c_int32), ("world_size", C.c_int32),
        ("flags", C.c_uint32), ("sampler", C.c_uint32),
        ("direct_emitter_samples", C.c_uint32), ("direct_bsdf_samples", C.c_uint32),
    ]


class B2Stats(C.Structure):
    _fields_ = [
        ("n_samples", C.c_uint64), ("n_segments", C.c_uint64), ("n_shadow_rays", C.c_uint64),
        ("n_invalid", C.c_uint64), ("n_iterations", C.c_uint64), ("n_launches", C.c_uint64),
        ("ms_total", C.c_float),
        ("ms_regen", C.c_float), ("ms_intersect", C.c_float), ("ms_shade", C.c_float), ("ms_shadow", C.c_float),
        ("bvh_nodes", C.c_uint64), ("bvh_tris", C.c_uint64),
        ("n_nodes_visited", C.c_uint64), ("n_tris_tested", C.c_uint64),
        ("bvh_depth", C.c_uint32), ("_pad", C.c_uint32),
        ("ms_build", C.c_float), ("ms_upload", C.c_float),
    ]

    def as_dict(self):
        return {n: getattr(self, n) for n, _ in self._fields_}


def fptr(a):
    return a.ctypes.data_as(C.POINTER(C.c_float))


def uptr(a):
    return a.ctypes.data_as(C.POINTER(C.c_uint32))


def render_params(max_depth=6, rr_depth=5, spp=16, samples_per_pass=0, block_size=32, seed=0,
                  rank=0, world_size=1, lane_slots=0, flags=0, sampler="independent", jitter=True):
    smp = {"independent": SAMPLER_INDEPENDENT, "stratified": SAMPLER_STRATIFIED, "multijitter": SAMPLER_MULTIJITTER, "ldsampler": SAMPLER_LOW_DISCREPANCY, "orthogonal": SAMPLER_ORTHOGONAL}[sampler] | (0 if jitter else SAMPLER_NO_JITTER)
    return B2RenderParams(max_depth, rr_depth, spp, samples_per_pass, block_size, lane_slots, seed, rank, world_size,
                          flags, smp)


SAMPLER_INDEPENDENT, SAMPLER_STRATIFIED, SAMPLER_MULTIJITTER, SAMPLER_LOW_DISCREPANCY, SAMPLER_ORTHOGONAL, SAMPLER_NO_JITTER = 0, 1, 2, 3, 4, 0x100


def fresnel_diffuse_reflectance(eta):
    """include/mitsuba/render/fresnel.h:327-361 in fp32 (left-to-right evaluation, no contraction)."""
    f = np.float32
    eta = f(eta)
    if eta < f(1):
        return f(f(f(f(-1.4399) * f(eta * eta)) + f(f(0.7099) * eta)) + f(0.6681)) + f(f(0.0636) / eta)
    i1 = f(1) / eta; i2 = f(i1 * i1); i3 = f(i2 * i1); i4 = f(i3 * i1); i5 = f(i4 * i1)
    r = f(f(0.919317) - f(f(3.4793) * i1))
    r = f(r + f(f(6.75335) * i2)); r = f(r - f(f(7.80989) * i3)); r = f(r + f(f(4.98554) * i4)); r = f(r - f(f(1.36881) * i5))
    return r


def _rgb_mean(c):
    """Texture::mean() of a constant rgb / uniform texture (srgb.cpp:52-57: hsum * (1/3); uniform: the value)."""
    f = np.float32
    if np.isscalar(c):
        return f(c)
    c = [f(x) for x in c]
    if c[0] == c[1] == c[2]:
        pass
    return f(f(f(c[0] + c[1]) + c[2]) * f(f(1) / f(3)))


class SceneBuffers:
    """Owns the numpy arrays behind a B2SceneDesc (keeps them alive) — the host-side
    flattening of Mitsuba's Scene object graph (scene.cpp:23-59, mesh.h:344-361)."""

    def __init__(self):
        self.meshes = []      # dicts: positions (n,3) f32, normals or None, faces (m,3) u32, bsdf, emitter
        self.bsdfs = []       # B2Bsdf
        self.emitters = []    # B2Emitter
        self.sensor = B2Sensor()
        self.film = B2Film()
        self.groups = []      # list of lists of mesh indices (shapegroups, src/librender/shapegroup.cpp)
        self.instances = []   # (group index, to_world 4x4, to_object 4x4)   (src/shapes/instance.cpp)
        self.shapes = []      # B2Shape: analytic rectangles / spheres (src/shapes/{rectangle,sphere}.cpp)
        self.textures = []    # B2Texture (src/textures/{checkerboard,bitmap}.cpp)
        self._tex_mean = []; self._tex_data = []
        self._keep = []
        self.build_flags = BUILD_HOST_SAH      # B2SceneDesc.build_flags (BUILD_DEVICE_LBVH: build the BVHs on the GPU)

    # -- BSDF / emitter constructors with the reference plugins' defaults --------------------
    def add_diffuse(self, reflectance=(0.5, 0.5, 0.5), texture=0):
        b = B2Bsdf(); b.type = BSDF_DIFFUSE
        b.reflectance[:] = reflectance; b.reflectance_texture = texture
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

    def add_roughdielectric(self, alpha=0.1, distribution="beckmann", int_ior=1.5046, ext_ior=1.000277,
                            specular_reflectance=(1, 1, 1), specular_transmittance=(1, 1, 1)):
        """RoughDielectric (src/bsdfs/roughdielectric.cpp). Oracle only in round 1: b2_scene_create rejects it."""
        b = B2Bsdf(); b.type = BSDF_ROUGHDIELECTRIC
        b.distribution = MF_GGX if distribution == "ggx" else MF_BECKMANN
        b.sample_visible = 1
        b.alpha_u = b.alpha_v = alpha
        b.eta = float(np.float32(int_ior) / np.float32(ext_ior))     # roughdielectric.cpp:171
        b.specular_reflectance[:] = specular_reflectance; b.specular_transmittance[:] = specular_transmittance
        self.bsdfs.append(b); return len(self.bsdfs) - 1

    def add_conductor(self, eta=(0, 0, 0), k=(1, 1, 1), specular_reflectance=(1, 1, 1)):
        """SmoothConductor (conductor.cpp:160-176); material 'none' = eta 0, k 1 (a perfect mirror)."""
        b = B2Bsdf(); b.type = BSDF_CONDUCTOR
        b.cond_eta[:] = eta; b.cond_k[:] = k
        b.specular_reflectance[:] = specular_reflectance; b.specular_transmittance[:] = (1, 1, 1)
        self.bsdfs.append(b); return len(self.bsdfs) - 1

    def add_plastic(self, diffuse_reflectance=(0.5, 0.5, 0.5), int_ior=1.49, ext_ior=1.000277, specular_reflectance=None,
                    nonlinear=False, texture=0):
        """SmoothPlastic (plastic.cpp:136-180), with the host-side values of parameters_changed()
        (:162-176) computed in fp32 like the reference: 1/eta^2, the diffuse Fresnel reflectance fit
        (fresnel.h:327-361) and the specular sampling weight from the textures' means
        (srgb.cpp:52-57: hmean of the rgb triple; uniform.cpp: the value)."""
        f = np.float32
        b = B2Bsdf(); b.type = BSDF_PLASTIC
        eta = f(int_ior) / f(ext_ior)
        b.eta = float(eta)
        b.inv_eta_2 = float(f(1) / (eta * eta))
        b.fdr_int = float(fresnel_diffuse_reflectance(f(1) / eta))
        b.reflectance[:] = diffuse_reflectance
        b.specular_reflectance[:] = (1, 1, 1) if specular_reflectance is None else specular_reflectance
        b.specular_transmittance[:] = (1, 1, 1)
        d_mean = self._tex_mean[texture - 1] if texture else _rgb_mean(diffuse_reflectance)
        b.reflectance_texture = texture
        s_mean = f(1) if specular_reflectance is None else _rgb_mean(specular_reflectance)
        b.specular_sampling_weight = float(s_mean / (d_mean + s_mean))
        b.nonlinear = 1 if nonlinear else 0
        self.bsdfs.append(b); return len(self.bsdfs) - 1

    def make_twosided(self, front, back=None):
        """TwoSidedBRDF (src/bsdfs/twosided.cpp): bsdfs[front] becomes two-sided with bsdfs[back] on the back side
        (back=None: the same BRDF on both sides). Returns `front`. Transmissive BSDFs cannot be nested (:90-91)."""
        back = front if back is None else back
        for i in (front, back):
            if self.bsdfs[i].type == BSDF_DIELECTRIC:
                raise ValueError("Only materials without a transmission component can be nested!")
        self.bsdfs[front].twosided_back = back + 1
        return front

    def add_area_emitter(self, mesh_index, radiance):
        e = B2Emitter(); e.type = EMITTER_AREA; e.mesh = mesh_index
        e.radiance[:] = radiance; e.scale = 1.0
        self.emitters.append(e)
        self.meshes[mesh_index]["emitter"] = len(self.emitters) - 1
        return len(self.emitters) - 1

    def add_constant_emitter(self, radiance=(1.0, 1.0, 1.0)):
        """ConstantBackgroundEmitter (src/emitters/constant.cpp): uniform environment radiance."""
        e = B2Emitter(); e.type = EMITTER_CONSTANT; e.mesh = -1; e.scale = 1.0
        e.radiance[:] = radiance
        self.emitters.append(e)
        return len(self.emitters) - 1

    def add_point_light(self, position, intensity=(1.0, 1.0, 1.0)):
        """PointLight (src/emitters/point.cpp): isotropic point emitter at `position` (to_world = translate(position))."""
        e = B2Emitter(); e.type = EMITTER_POINT; e.mesh = -1; e.scale = 1.0
        e.radiance[:] = intensity
        m = np.eye(4, dtype=np.float32); m[:3, 3] = position
        e.env_to_world[:] = [float(x) for x in m.reshape(-1)]
        inv = np.eye(4, dtype=np.float32); inv[:3, 3] = -np.asarray(position, np.float32)
        e.env_to_object[:] = [float(x) for x in inv.reshape(-1)]
        self.emitters.append(e)
        return len(self.emitters) - 1

    def add_spot_light(self, to_world, intensity=(1.0, 1.0, 1.0), cutoff_angle=20.0, beam_width=None):
        """SpotLight (src/emitters/spot.cpp), untextured; angles in degrees like the plugin. Oracle only in round 1."""
        e = B2Emitter(); e.type = EMITTER_SPOT; e.mesh = -1; e.scale = 1.0
        e.radiance[:] = intensity
        m = np.asarray(to_world, np.float32).reshape(4, 4)
        e.env_to_world[:] = [float(x) for x in m.reshape(-1)]
        e.env_to_object[:] = [float(x) for x in np.linalg.inv(m.astype(np.float64)).astype(np.float32).reshape(-1)]
        f = np.float32
        bw = f(cutoff_angle) * f(3.0) / f(4.0) if beam_width is None else f(beam_width)      # spot.cpp:88-92
        d2r = f(np.pi / 180.0)
        e.cutoff_angle = float(f(cutoff_angle) * d2r); e.beam_width = float(f(bw) * d2r)
        self.emitters.append(e); return len(self.emitters) - 1

    def add_directional_light(self, to_world, irradiance=(1.0, 1.0, 1.0)):
        """DirectionalEmitter (src/emitters/directional.cpp): emits along to_world * (0, 0, 1). Oracle only in round 1."""
        e = B2Emitter(); e.type = EMITTER_DIRECTIONAL; e.mesh = -1; e.scale = 1.0
        e.radiance[:] = irradiance
        m = np.asarray(to_world, np.float32).reshape(4, 4)
        e.env_to_world[:] = [float(x) for x in m.reshape(-1)]
        e.env_to_object[:] = [float(x) for x in np.linalg.inv(m.astype(np.float64)).astype(np.float32).reshape(-1)]
        self.emitters.append(e); return len(self.emitters) - 1

    def add_envmap(self, rgb, scale=1.0, to_world=None):
        """EnvironmentMapEmitter (src/emitters/envmap.cpp): rgb = float32 [H, W, 3] lat-long image."""
        rgb = np.asarray(rgb, np.float32)
        H, W = rgb.shape[:2]
        rgba = np.ones((H, W, 4), np.float32); rgba[..., :3] = rgb[..., :3]
        self._env_data = np.ascontiguousarray(rgba)
        e = B2Emitter(); e.type = EMITTER_ENVMAP; e.mesh = -1; e.scale = scale
        e.envmap = fptr(self._env_data); e.env_width = W; e.env_height = H
        m = np.eye(4, dtype=np.float32) if to_world is None else np.asarray(to_world, np.float32).reshape(4, 4)
        inv = np.linalg.inv(m.astype(np.float64)).astype(np.float32)
        e.env_to_world[:] = [float(x) for x in m.reshape(-1)]
        e.env_to_object[:] = [float(x) for x in inv.reshape(-1)]
        self.emitters.append(e)
        return len(self.emitters) - 1

    # -- textures (1-based handles; pass as texture= to add_diffuse / add_plastic) ---------------
    def _add_texture(self, t, to_uv, mean):
        m = np.eye(3, dtype=np.float32) if to_uv is None else np.asarray(to_uv, np.float32).reshape(3, 3)
        t.to_uv[:] = [float(x) for x in m.reshape(-1)]
        self.textures.append(t); self._tex_mean.append(np.float32(mean))
        return len(self.textures)

    def add_checkerboard(self, color0=(0.4, 0.4, 0.4), color1=(0.2, 0.2, 0.2), to_uv=None):
        """Checkerboard (src/textures/checkerboard.cpp:40-44), constant colours; mean() :82-84."""
        t = B2Texture(); t.type = TEX_CHECKERBOARD
        t.color0[:] = color0; t.color1[:] = color1
        return self._add_texture(t, to_uv, np.float32(0.5) * (_rgb_mean(color0) + _rgb_mean(color1)))

    def add_bitmap(self, data, filter_type="bilinear", wrap_mode="repeat", to_uv=None):
        """BitmapTexture (src/textures/bitmap.cpp): data = float32 [H, W] or [H, W, 3], linear, as the
        reference holds it after Bitmap::convert; mean() is the mean luminance accumulated in double (:159-205)."""
        d = np.ascontiguousarray(data, np.float32)
        if d.ndim == 2:
            d = d[..., None]
        t = B2Texture(); t.type = TEX_BITMAP
        t.filter = TEXFILTER_NEAREST if filter_type == "nearest" else TEXFILTER_BILINEAR
        t.wrap = {"repeat": WRAP_REPEAT, "mirror": WRAP_MIRROR, "clamp": WRAP_CLAMP}[wrap_mode]
        t.height, t.width, t.channels = d.shape
        t.data = fptr(d)
        self._tex_data.append(d)
        lum = d[..., 0] if d.shape[2] == 1 else (np.float32(0.212671) * d[..., 0] + np.float32(0.715160) * d[..., 1]
                                                  + np.float32(0.072169) * d[..., 2])
        return self._add_texture(t, to_uv, np.float32(lum.astype(np.float64).mean()))

    def add_mesh(self, positions, faces, normals=None, bsdf=0, texcoords=None):
        m = dict(positions=np.ascontiguousarray(positions, np.float32).reshape(-1, 3),
                 faces=np.ascontiguousarray(faces, np.uint32).reshape(-1, 3),
                 normals=None if normals is None else np.ascontiguousarray(normals, np.float32).reshape(-1, 3),
                 bsdf=bsdf, emitter=-1,
                 texcoords=None if texcoords is None else np.ascontiguousarray(texcoords, np.float32).reshape(-1, 2))
        self.meshes.append(m); return len(self.meshes) - 1

    # -- analytic shapes (top level; handles are 0-based indices into B2SceneDesc.shapes) ----------
    def _add_shape(self, sh, to_world, bsdf):
        m = np.asarray(to_world, np.float32).reshape(4, 4)
        inv = np.linalg.inv(m.astype(np.float64)).astype(np.float32)
        sh.to_world[:] = [float(x) for x in m.reshape(-1)]
        sh.to_object[:] = [float(x) for x in inv.reshape(-1)]
        sh.bsdf = bsdf; sh.emitter = -1
        self.shapes.append(sh)
        return len(self.shapes) - 1

    def add_rectangle(self, to_world, bsdf=0, flip_normals=False):
        """Rectangle (src/shapes/rectangle.cpp): the [-1,1]^2 square of the xy plane under to_world (4x4);
        flip_normals is folded into the transform like the constructor does (:75-77)."""
        m = np.asarray(to_world, np.float32).reshape(4, 4)
        if flip_normals:
            m = (m.astype(np.float64) @ np.diag([1.0, 1.0, -1.0, 1.0])).astype(np.float32)
        sh = B2Shape(); sh.type = SHAPE_RECTANGLE
        return self._add_shape(sh, m, bsdf)

    def add_sphere(self, center=(0.0, 0.0, 0.0), radius=1.0, bsdf=0, flip_normals=False):
        """Sphere (src/shapes/sphere.cpp) given by 'center' / 'radius': to_world = translate(center) * scale(radius)
        (:100-101); update() recovers exactly these for a shear-free, uniformly scaled transform (:107-130)."""
        sh = B2Shape(); sh.type = SHAPE_SPHERE
        sh.center[:] = center; sh.radius = radius; sh.flip_normals = 1 if flip_normals else 0
        m = np.eye(4, dtype=np.float32)
        m[0, 0] = m[1, 1] = m[2, 2] = radius; m[:3, 3] = center
        return self._add_shape(sh, m, bsdf)

    def add_shape_emitter(self, shape_index, radiance):
        e = B2Emitter(); e.type = EMITTER_AREA; e.mesh = -1; e.shape = shape_index + 1
        e.radiance[:] = radiance; e.scale = 1.0
        self.emitters.append(e)
        self.shapes[shape_index].emitter = len(self.emitters) - 1
        return len(self.emitters) - 1

    def add_shapegroup(self, mesh_indices):
        """Move the given meshes (added with add_mesh) into a new shapegroup; returns its index."""
        self.groups.append(list(mesh_indices)); return len(self.groups) - 1

    def add_instance(self, group, to_world):
        """to_world: 4x4 (row-major). The inverse is computed in fp64 and rounded to fp32 — exact for
        the rigid / axis-permuting / power-of-two transforms used by the parity scenes."""
        m = np.asarray(to_world, np.float32).reshape(4, 4)
        inv = np.linalg.inv(m.astype(np.float64)).astype(np.float32)
        self.instances.append((group, m, inv)); return len(self.instances) - 1

    @property
    def n_triangles(self):
        return int(sum(len(m["faces"]) for m in self.meshes))

    def desc(self):
        d = B2SceneDesc()
        # top-level meshes first, then the shapegroups' meshes (B2SceneDesc.group_first)
        grouped = [m for g in self.groups for m in g]
        order = [i for i in range(len(self.meshes)) if i not in grouped] + grouped
        remap = {old: new for new, old in enumerate(order)}
        n_top = len(self.meshes) - len(grouped)
        ms = (B2Mesh * max(1, len(self.meshes)))()
        for new, old in enumerate(order):
            m = self.meshes[old]
            ms[new].n_vertices = len(m["positions"]); ms[new].n_faces = len(m["faces"])
            ms[new].positions = fptr(m["positions"])
            ms[new].normals = fptr(m["normals"]) if m["normals"] is not None else None
            ms[new].texcoords = fptr(m["texcoords"]) if m.get("texcoords") is not None else None
            ms[new].faces = uptr(m["faces"])
            ms[new].bsdf = m["bsdf"]; ms[new].emitter = m["emitter"]
        bs = (B2Bsdf * max(1, len(self.bsdfs)))(*self.bsdfs)
        es = (B2Emitter * max(1, len(self.emitters)))()
        for i, e in enumerate(self.emitters):
            C.memmove(C.byref(es[i]), C.byref(e), C.sizeof(B2Emitter))
            es[i].mesh = remap[e.mesh] if e.mesh >= 0 else -1
        d.n_meshes = len(self.meshes); d.meshes = ms
        d.n_bsdfs = len(self.bsdfs); d.bsdfs = bs
        d.n_emitters = len(self.emitters); d.emitters = es
        gf = (C.c_uint32 * (len(self.groups) + 1))()
        pos = n_top
        for gi, g in enumerate(self.groups):
            gf[gi] = pos; pos += len(g)
        gf[len(self.groups)] = pos
        ins = (B2Instance * max(1, len(self.instances)))()
        for i, (g, m, inv) in enumerate(self.instances):
            ins[i].group = g
            ins[i].to_world[:] = [float(x) for x in m.reshape(-1)]
            ins[i].to_object[:] = [float(x) for x in inv.reshape(-1)]
        d.n_groups = len(self.groups); d.group_first = gf if self.groups else None
        d.n_instances = len(self.instances); d.instances = ins if self.instances else None
        d.n_top_meshes = n_top
        d.sensor = self.sensor; d.film = self.film
        texs = self.textures
        ts = (B2Texture * max(1, len(texs)))(*texs)
        d.n_textures = len(texs); d.textures = ts if texs else None
        shs = self.shapes
        ss = (B2Shape * max(1, len(shs)))(*shs)
        d.n_shapes = len(shs); d.shapes = ss if shs else None
        self._keep = [ms, bs, es, gf, ins, ts, ss]
        d.build_flags = self.build_flags
        return d
