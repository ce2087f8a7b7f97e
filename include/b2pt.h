/*
 * b2pt.h — C-ABI of libb2pt: the B200-native wavefront path tracer that drops in under
 * Mitsuba 2's scalar_rgb PathIntegrator / Scene hot path.
 *
 * Plain C: pointers + sizes only, no C++/torch types, no exceptions. Every entry point
 * returns an int status (0 = ok); b2_last_error() gives the message for the calling thread.
 * All file:line citations are relative to the reference tree (mitsuba2 @ 62863cb3).
 *
 * What each group of entry points replaces on the reference side:
 *
 *   b2_scene_create / b2_scene_destroy
 *        Scene::accel_init_gpu / accel_release_gpu     include/mitsuba/render/scene.h:184-190
 *        (+ the flattening of Mesh / BSDF / Emitter / Sensor / Film plugin state the
 *         reference reads through virtual calls: mesh.h:344-361, scene.cpp:23-59)
 *   b2_render / b2_render_device / b2_cancel
 *        Integrator::render(Scene*, Sensor*) / cancel() include/mitsuba/render/integrator.h:42-51
 *        = SamplingIntegrator::render + render_block + render_sample (src/librender/
 *          integrator.cpp:51-288) + PathIntegrator::sample (src/integrators/path.cpp:100-211)
 *   b2_stage_*           (device SoA queues; per-stage parity tests and ncu)
 *        raygen    : render_sample + PerspectiveCamera::sample_ray_differential
 *                    (integrator.cpp:233-262, src/sensors/perspective.cpp:182-216)
 *        intersect : Scene::ray_intersect_preliminary_gpu   scene.h:196, scene_optix.inl:426
 *        occluded  : Scene::ray_test_gpu                    scene.h:202, scene_optix.inl:554
 *        si_fill   : Mesh::compute_surface_interaction      src/librender/mesh.cpp:449-545
 *        (+ what widened from there, SURVEY 8(f): textures B2Texture, analytic shapes B2Shape, twosided / conductor /
 *         plastic BSDFs, constant and point emitters, the direct integrator's default, stratified / multijitter /
 *         low-discrepancy samplers B2RenderParams.sampler — each citing its reference file below)
 *   b2_develop
 *        HDRFilm::bitmap(): XYZAW -> linear sRGB, divide by W  src/films/hdrfilm.cpp:251-322
 *   b2_host_*            (host-side C++ helpers that mirror reference host code)
 *        perspective_projection / update_camera_transforms  include/mitsuba/render/sensor.h:196-232,
 *                                                            src/sensors/perspective.cpp:119-139
 *        Transform::look_at                                 include/mitsuba/core/transform.h:241-270
 *        ReconstructionFilter::init_discretization          src/libcore/rfilter.cpp:9-20
 *        Spiral::next_block order                           src/librender/spiral.cpp:27-72
 */
#ifndef B2PT_H
#define B2PT_H

#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define B2_API __attribute__((visibility("default")))

/* ---- status codes --------------------------------------------------------------------- */
enum {
    B2_OK = 0,
    B2_ERR_INVALID = 1,     /* bad argument / unsupported scene feature (reference: Throw())  */
    B2_ERR_CUDA = 2,        /* CUDA runtime failure                                           */
    B2_ERR_CANCELLED = 3,   /* b2_cancel() observed (reference: render() returns false)       */
    B2_ERR_NO_DEVICE = 4    /* no CUDA device: the product path has NO CPU fallback           */
};

/* ---- BSDF records (src/bsdfs/{diffuse,dielectric,roughconductor,conductor,plastic,twosided}.cpp) */
enum {
    B2_BSDF_DIFFUSE = 0,        /* SmoothDiffuse      diffuse.cpp:78-135                      */
    B2_BSDF_DIELECTRIC = 1,     /* SmoothDielectric   dielectric.cpp:201-320                  */
    B2_BSDF_ROUGHCONDUCTOR = 2, /* RoughConductor     roughconductor.cpp:196-382              */
    B2_BSDF_NULL = 3,           /* reflectance-0 diffuse given to emitters without a BSDF     */
    B2_BSDF_CONDUCTOR = 4,      /* SmoothConductor    conductor.cpp:197-258                   */
    B2_BSDF_PLASTIC = 5,        /* SmoothPlastic      plastic.cpp:181-314                     */
    B2_BSDF_ROUGHDIELECTRIC = 6 /* RoughDielectric    roughdielectric.cpp:203-452 — ORACLE ONLY in round 1: restated and
                                   pinned in oracle/b2o.c; b2_scene_create rejects it until the CUDA path has it */
};
enum { B2_MF_BECKMANN = 0, B2_MF_GGX = 1 };   /* MicrofacetType, microfacet.h:27-33          */

typedef struct B2Bsdf {
    int32_t type;               /* B2_BSDF_*                                                   */
    int32_t distribution;       /* B2_MF_* (roughconductor)                                    */
    int32_t sample_visible;     /* roughconductor 'sample_visible' (default 1)                 */
    int32_t reflectance_texture;/* 0: 'reflectance' is the constant below; k > 0: B2SceneDesc.textures[k-1]
                                   (diffuse 'reflectance', plastic 'diffuse_reflectance')        */
    float reflectance[3];       /* diffuse 'reflectance' (constant rgb texture)                */
    float eta;                  /* dielectric int_ior / ext_ior                                */
    float specular_reflectance[3];    /* dielectric / roughconductor, default 1               */
    float alpha_u;
    float specular_transmittance[3];  /* dielectric, default 1                                */
    float alpha_v;
    float cond_eta[3];          /* (rough)conductor complex IOR, real part (default 0)         */
    int32_t twosided_back;      /* TwoSidedBRDF (twosided.cpp:63-92): 0 = one-sided; k > 0 = this record is the front-side
                                   BRDF and bsdfs[k-1] the back-side one (k-1 = own index when a single BRDF is nested) */
    float cond_k[3];            /* imaginary part (default 1)                                  */
    float _pad2;
    /* plastic (plastic.cpp:136-180): reflectance = diffuse_reflectance, eta = int_ior / ext_ior, and the
       values parameters_changed() derives on the host                                           */
    float fdr_int;              /* fresnel_diffuse_reflectance(1 / eta), fresnel.h:327-361     */
    float specular_sampling_weight;   /* s_mean / (d_mean + s_mean), plastic.cpp:174           */
    float inv_eta_2;            /* 1 / eta^2                                                   */
    int32_t nonlinear;          /* 'nonlinear' (default 0)                                     */
} B2Bsdf;

/* ---- textures (src/textures/{checkerboard,bitmap}.cpp), evaluated at si.uv ----------------- */
enum { B2_TEX_CHECKERBOARD = 0, B2_TEX_BITMAP = 1 };
enum { B2_TEXFILTER_NEAREST = 0, B2_TEXFILTER_BILINEAR = 1 };         /* bitmap.cpp:77 */
enum { B2_WRAP_REPEAT = 0, B2_WRAP_MIRROR = 1, B2_WRAP_CLAMP = 2 };   /* bitmap.cpp:78 */

typedef struct B2Texture {
    int32_t type;               /* B2_TEX_*                                                    */
    int32_t filter, wrap;       /* bitmap 'filter_type' (default bilinear), 'wrap_mode' (repeat) */
    int32_t width, height;      /* bitmap resolution                                           */
    int32_t channels;           /* bitmap: 1 (Y) or 3 (linear RGB), bitmap.cpp:121-141         */
    float to_uv[9];             /* row-major 3x3, ScalarTransform3f m_transform ('to_uv')      */
    float color0[3], color1[3]; /* checkerboard 'color0' (default .4) / 'color1' (.2), constant */
    const float *data;          /* bitmap texels as the reference holds them after conversion:
                                   height x width x channels fp32, row-major                   */
} B2Texture;

/* ---- emitters (src/emitters/{area,envmap,constant,point}.cpp) --------------------------- */
enum { B2_EMITTER_AREA = 0, B2_EMITTER_ENVMAP = 1, B2_EMITTER_CONSTANT = 2, B2_EMITTER_POINT = 3 /* src/emitters/point.cpp */,
       /* ORACLE ONLY in round 1 (restated and pinned in oracle/b2o.c, rejected by b2_scene_create): */
       B2_EMITTER_SPOT = 4 /* src/emitters/spot.cpp, untextured */, B2_EMITTER_DIRECTIONAL = 5 /* src/emitters/directional.cpp */ };

typedef struct B2Emitter {
    int32_t type;               /* B2_EMITTER_*                                                */
    int32_t mesh;               /* area: index of the mesh it is attached to                   */
    float radiance[3];          /* area / constant: rgb radiance (srgb_d65 in rgb mode); point: 'intensity' */
    float scale;                /* envmap 'scale'                                              */
    const float *envmap;        /* envmap: RGBA fp32 texels, (height x width x 4), row-major   */
    int32_t env_width, env_height;
    float env_to_world[16];     /* envmap / point to_world, row-major 4x4 (point: its translation is the position) */
    float env_to_object[16];    /* its inverse, as the reference's Transform stores it         */
    int32_t shape;              /* area: 0 = attached to `mesh`; k > 0 = attached to B2SceneDesc.shapes[k-1] */
    int32_t _pad;
    float cutoff_angle, beam_width;   /* spot: radians, as SpotLight stores them (spot.cpp:88-92)       */
} B2Emitter;

/* ---- meshes (include/mitsuba/render/mesh.h:344-361) ------------------------------------- */
typedef struct B2Mesh {
    uint32_t n_vertices, n_faces;
    const float *positions;     /* 3 fp32 / vertex, world space (obj.cpp:172)                  */
    const float *normals;       /* 3 fp32 / vertex or NULL (face normals)                      */
    const float *texcoords;     /* 2 fp32 / vertex or NULL                                     */
    const uint32_t *faces;      /* 3 u32 / face                                                */
    int32_t bsdf;               /* index into B2SceneDesc.bsdfs                                */
    int32_t emitter;            /* index into B2SceneDesc.emitters or -1                       */
} B2Mesh;

/* ---- analytic shapes (src/shapes/rectangle.cpp, src/shapes/sphere.cpp), top level only ---- */
enum { B2_SHAPE_RECTANGLE = 0, B2_SHAPE_SPHERE = 1 };

typedef struct B2Shape {
    int32_t type;               /* B2_SHAPE_*                                                  */
    int32_t bsdf;               /* index into B2SceneDesc.bsdfs                                */
    int32_t emitter;            /* index into B2SceneDesc.emitters or -1                       */
    int32_t flip_normals;       /* sphere 'flip_normals' (rectangle folds it into to_world, rectangle.cpp:75-77) */
    float to_world[16];         /* row-major m_to_world (sphere: as recomposed by update(), sphere.cpp:127) */
    float to_object[16];        /* its inverse, as the reference's Transform stores it         */
    float center[3];            /* sphere m_center                                             */
    float radius;               /* sphere m_radius                                             */
} B2Shape;

/* one-level instancing (src/shapes/instance.cpp:107-164, src/librender/shapegroup.cpp) */
typedef struct B2Instance {
    int32_t group;              /* index of the shapegroup                                     */
    float to_world[16];         /* row-major 4x4                                               */
    float to_object[16];        /* its inverse, as the reference stores it                     */
} B2Instance;

/* ---- sensor + film (perspective.cpp, hdrfilm.cpp, rfilter.h) ---------------------------- */
typedef struct B2Sensor {
    float sample_to_camera[16]; /* row-major; m_sample_to_camera  perspective.cpp:124          */
    float to_world[16];         /* row-major; m_world_transform                                */
    float near_clip, far_clip;
    float principal_point_offset[2];
} B2Sensor;

#define B2_FILTER_RESOLUTION 31  /* MTS_FILTER_RESOLUTION, include/mitsuba/core/rfilter.h:10   */

typedef struct B2Film {
    int32_t width, height;              /* film->size()                                        */
    int32_t crop_offset[2], crop_size[2];
    float filter_radius;                /* ReconstructionFilter::radius()                      */
    float filter_scale;                 /* m_scale_factor = 31 / radius                        */
    float filter_lut[B2_FILTER_RESOLUTION + 1];
    int32_t border;                     /* border_size()                                       */
} B2Film;

typedef struct B2SceneDesc {
    uint32_t n_meshes;   const B2Mesh *meshes;
    uint32_t n_bsdfs;    const B2Bsdf *bsdfs;
    uint32_t n_emitters; const B2Emitter *emitters;   /* scene.m_emitters order (scene.cpp:23-59) */
    /* shapegroups: meshes [group_first[g], group_first[g+1]) belong to group g and are NOT
       part of the top-level scene; n_groups == 0 for flat scenes. */
    uint32_t n_groups;    const uint32_t *group_first;
    uint32_t n_instances; const B2Instance *instances;
    uint32_t n_top_meshes;              /* meshes [0, n_top_meshes) are top-level shapes       */
    B2Sensor sensor;
    B2Film film;
    uint32_t n_textures; const B2Texture *textures;
    /* analytic shapes; in the closest-hit tie rule they follow all mesh triangles (primitive index = total
       triangle count + shape index, like a ShapeKDTree whose meshes come first) */
    uint32_t n_shapes; const B2Shape *shapes;
    uint32_t build_flags;               /* B2_BUILD_*: which BVH builder b2_scene_create uses */
} B2SceneDesc;

/* BVH builders (both replace ShapeKDTree::build, kdtree.h:957-1900; the rendered film does not depend on the choice:
   closest hits are defined independently of the tree) */
#define B2_BUILD_HOST_SAH 0u      /* binned-SAH BVH2 on the host, collapsed to BVH4 (best traversal speed)                 */
#define B2_BUILD_DEVICE_LBVH 1u   /* LBVH on the device (Morton sort + Karras hierarchy + fit + collapse): 10 M triangles in
                                     tens of ms, for large or per-frame-changing geometry; triangle meshes and shapegroups
                                     (the BVH over analytic shapes / instances stays on the host: it is tiny)            */

/* ---- render parameters (integrator.cpp:23-40, sampler.cpp:20-30) ------------------------ */
typedef struct B2RenderParams {
    int32_t max_depth;          /* MonteCarloIntegrator 'max_depth' (-1 = infinite)            */
    int32_t rr_depth;           /* 'rr_depth' (default 5)                                      */
    uint32_t spp;               /* sampler 'sample_count'                                      */
    uint32_t samples_per_pass;  /* integrator 'samples_per_pass' (0 = spp)                     */
    uint32_t block_size;        /* integrator 'block_size' (power of two; MUST be pinned, it
                                   feeds the per-pixel seed, integrator.cpp:198)               */
    uint32_t lane_slots;        /* wavefront width (0 = auto)                                  */
    uint64_t seed;              /* sampler 'seed' (m_base_seed)                                */
    int32_t rank, world_size;   /* this process renders spiral blocks b with b % world == rank */
    uint32_t flags;             /* B2_FLAG_* */
    uint32_t sampler;           /* B2_SAMPLER_* (| B2_SAMPLER_NO_JITTER for stratified 'jitter' = false) */
    /* DirectIntegrator (src/integrators/direct.cpp:78-196) instead of PathIntegrator: 'emitter_samples' / 'bsdf_samples'
       (both 0 = path integrator; max_depth / rr_depth are ignored when either is set). */
    uint32_t direct_emitter_samples, direct_bsdf_samples;
} B2RenderParams;

/* samplers: IndependentSampler (src/samplers/independent.cpp) and StratifiedSampler (src/samplers/stratified.cpp;
   spp must be a perfect square, :60-75; multijitter: spp = rx * ry with ry = floor(sqrt(spp)), multijitter.cpp:68-81) */
enum { B2_SAMPLER_INDEPENDENT = 0, B2_SAMPLER_STRATIFIED = 1, B2_SAMPLER_MULTIJITTER = 2 /* src/samplers/multijitter.cpp */,
       B2_SAMPLER_LOW_DISCREPANCY = 3 /* src/samplers/ldsampler.cpp: (0,2)-sequence; spp = 4^k */,
       B2_SAMPLER_ORTHOGONAL = 4 /* src/samplers/orthogonal.cpp, default strength 2 (Bose construction, :188-218): spp = p^2, p prime */,
       B2_SAMPLER_NO_JITTER = 0x100 };

#define B2_FLAG_SINGLE_SET 2u   /* one lane set on one stream (default: two sets on two streams) */
#define B2_FLAG_STAGE_TIMING 1u /* bracket every stage launch with CUDA events -> B2Stats.ms_<stage> */
#define B2_FLAG_PERSISTENT_TRACE 8u /* traverse with the persistent-warp / dynamic-ray-fetch kernel (k_trace_inst) also on scenes
                                   without instances (always used with instances; ignored for tiny scenes on the flat leaf scan) */
#define B2_FLAG_RESUME_TRACE 16u /* BVH4 scenes without instances: trace with the yield / resume kernel (k_trace_resume: a warp
                                   suspends its unfinished rays when few of its lanes are still traversing; the next launch
                                   continues them in full warps) instead of the static ray -> lane kernel. Same film either way */
#define B2_FLAG_YIELD_THRESHOLD(n) (((n) & 0x1fu) << 8)   /* k_trace_resume: suspend below n traversing lanes (0 = default 12) */
#define B2_FLAG_TRACE_BLOCK(sel) (((sel) & 3u) << 16)   /* tuning: threads per CTA of the traversal kernel: 0 = auto (128 for scenes that
                                                          stage <= 8 KB per CTA, else 256), 1 = 256, 2 = 128, 3 = 64 */
#define B2_FLAG_SHADE_BLOCK(sel) (((sel) & 3u) << 18)   /* ... of the shading kernel */
#define B2_FLAG_COUNT_BVH 4u    /* run the counting instantiation of the traversal kernel -> B2Stats.n_nodes_visited /
                                   n_tris_tested (BVH bytes = nodes * 128 B + tris * 48 B); slower, for measurement */

typedef struct B2Stats {
    uint64_t n_samples;         /* camera samples put on the film                              */
    uint64_t n_segments;        /* closest-hit rays traced (camera + bounce rays)              */
    uint64_t n_shadow_rays;     /* any-hit rays traced                                         */
    uint64_t n_invalid;         /* samples dropped by the ImageBlock::put validity rule        */
    uint64_t n_iterations;      /* wavefront iterations                                        */
    uint64_t n_launches;        /* kernels launched                                            */
    float ms_total;             /* device time of the whole render (CUDA events)               */
    float ms_regen, ms_intersect, ms_shade, ms_shadow;   /* per-stage device time: ms_regen = k_init,
                                                            ms_intersect = k_trace (closest + any hit), ms_shadow unused */
    uint64_t bvh_nodes, bvh_tris;
    uint64_t n_nodes_visited;   /* B2_FLAG_COUNT_BVH: BVH4 nodes fetched by all rays (closest + any hit) */
    uint64_t n_tris_tested;     /* B2_FLAG_COUNT_BVH: primitives tested                          */
    uint32_t bvh_depth;         /* levels of the deepest BVH4 (+ the instance BVH's)             */
    uint32_t _pad;
    float ms_build;             /* b2_scene_create: BVH build (host or device)                  */
    float ms_upload;            /* b2_scene_create: flattening + upload                          */
} B2Stats;

typedef struct B2Scene B2Scene;     /* opaque */

/* ---- lifecycle --------------------------------------------------------------------------- */
B2_API int b2_scene_create(const B2SceneDesc *desc, B2Scene **out);
B2_API int b2_scene_destroy(B2Scene *scene);
/* Scene::parameters_changed -> accel_parameters_changed (scene.cpp:242-260, scene_optix.inl:236): replace the geometry /
 * materials / emitters / sensor behind an existing handle by a new description (moved vertices, edited parameters). The BVHs
 * are rebuilt as in b2_scene_create (B2_BUILD_DEVICE_LBVH: on the GPU, milliseconds); render state (queues, streams, pinned
 * buffers, work-item tables, callbacks) is kept. On failure the handle keeps its old contents. Not concurrent with a render
 * on the same handle. */
B2_API int b2_scene_update(B2Scene *scene, const B2SceneDesc *desc);
B2_API const char *b2_last_error(void);
B2_API int b2_device_count(void);

/* ---- whole-path entry points ------------------------------------------------------------- */
/* Renders into a HOST buffer xyzaw[crop_h * crop_w * 5] (channel order X,Y,Z,A,W,
   integrator.cpp:71-72). Host<->device copies are inside the call. */
B2_API int b2_render(B2Scene *scene, const B2RenderParams *params, float *xyzaw, B2Stats *stats);
/* Same, but leaves the film in device memory (pointer returned, owned by the scene, valid
   until the next render / destroy). For multi-GPU the caller reduces it (ncclReduce). */
B2_API int b2_render_device(B2Scene *scene, const B2RenderParams *params, float **d_xyzaw,
                            B2Stats *stats);
/* Multi-GPU through the API: every rank (B2RenderParams.rank / world_size) renders its spiral blocks; `reduce` is then
   called ONCE with this rank's DEVICE film (n floats, enqueued work must go to `cuda_stream`) and must sum the rank
   films into rank 0's — e.g. ncclReduce(film, film, n, ncclFloat, ncclSum, 0, comm, stream), or torch.distributed.reduce
   on a tensor wrapping the pointer. Only after it the film is copied to `xyzaw` (may be NULL on ranks != 0), so the only
   device->host copy of the job is rank 0's. Returns non-zero from `reduce` to signal failure. */
typedef int (*b2_reduce_fn)(void *user, float *d_film, size_t n_floats, void *cuda_stream);
B2_API int b2_render_reduce(B2Scene *scene, const B2RenderParams *params, b2_reduce_fn reduce, void *user,
                            float *xyzaw, B2Stats *stats);
/* Progressive output (the reference develops the partially rendered film on SIGHUP, src/mitsuba/mitsuba.cpp:95-121, and shows
   progress per block, integrator.cpp:100-135): while a render runs, at most every `min_interval_ms` the film accumulated so far
   is copied to a pinned host buffer on a SIDE stream — the copy engine works while the wavefront kernels keep the SMs busy —
   and handed to `fn` from the rendering thread. Every pixel carries its own sample count in W, so developing the snapshot
   (XYZ / W) gives a valid partial image at any time. `fn` returns non-zero to cancel the render. fn = NULL switches it off. */
typedef int (*b2_progress_fn)(void *user, const float *xyzaw_snapshot, uint64_t samples_done, uint64_t samples_total);
B2_API int b2_set_progress_callback(B2Scene *scene, b2_progress_fn fn, void *user, uint32_t min_interval_ms);
/* b2_cancel(): async-safe, sets a flag; the running render returns B2_ERR_CANCELLED within one batch of iterations and the
   film (host buffer of b2_render / device film of b2_render_device) holds the samples accumulated so far, with their count
   in W — develop it like the reference develops a partial film on SIGHUP (src/mitsuba/mitsuba.cpp:112-121). */
B2_API int b2_cancel(B2Scene *scene);
/* HDRFilm::bitmap(): rgb[h*w*3] = M_xyz2rgb * xyz / W  (host in, host out, done on device) */
B2_API int b2_develop(B2Scene *scene, const float *xyzaw, float *rgb);
/* Same on the DEVICE film of the last b2_render_device / b2_render (no host round trip of the XYZAW film); rgb is a host buffer */
B2_API int b2_develop_device(B2Scene *scene, float *rgb);
/* run stream for b2_render_device (cudaStream_t as void*); default = legacy stream */
B2_API int b2_set_stream(B2Scene *scene, void *cuda_stream);

/* ---- stage-level entry points (host arrays in/out; the kernels run on device queues) ----- */
/* n lanes; (px,py) = integer pixel of each lane, seed_offset = value passed to Sampler::seed.
   Outputs: ray_o[n*4] = (o.xyz, mint), ray_d[n*4] = (d.xyz, maxt), rng_state[n] after the
   3 draws of render_sample (integrator.cpp:242,252), pos[n*2] = position_sample. */
B2_API int b2_stage_raygen(B2Scene *scene, uint32_t n, const uint32_t *px, const uint32_t *py,
                           const uint64_t *seed_offset, uint64_t base_seed,
                           float *ray_o, float *ray_d, uint64_t *rng_state, float *pos);
/* hit[n*4] = (t, u, v, prim-as-bits); t = +inf on miss; prim = global triangle index
   (mesh order, then face order); inst[n] = instance index or -1 (may be NULL). */
B2_API int b2_stage_intersect(B2Scene *scene, uint32_t n, const float *ray_o, const float *ray_d,
                              float *hit, int32_t *inst);
B2_API int b2_stage_occluded(B2Scene *scene, uint32_t n, const float *ray_o, const float *ray_d,
                             uint8_t *occluded);
/* si[n*16] = p3, n3, sh_n3, sh_s3, wi3, (bsdf index as bits) */
B2_API int b2_stage_si_fill(B2Scene *scene, uint32_t n, const float *ray_o, const float *ray_d,
                            const float *hit, const int32_t *inst, float *si);

/* shade: ONE iteration of PathIntegrator::sample's loop body (path.cpp:124-208) — SI fill, emitter hit + MIS weight, Russian
   roulette, emitter sampling, BSDF sampling, spawn_ray — i.e. one k_shade launch over n caller-provided lanes. Every lane
   must have hit something (hit t < inf, as b2_stage_intersect returns it) with depth < max_depth: the lanes the traversal
   stage puts on its CONTINUE list. Independent sampler only.
     in/out  ray_o, ray_d [n*4]       the segment's ray -> the next ray (maxt < 0: the path ended inside the loop body)
     in      hit [n*4], inst [n]      (t, u, v, prim bits); instance index or NULL for flat scenes
     in/out  throughput_eta [n*4]     (throughput.rgb, eta)
     in/out  result_pdf [n*4]         (result.rgb, pdf of the last BSDF sample — feeds the MIS weight of the next emitter hit)
     in/out  rng_state [n]            PCG32 state (default stream), as b2_stage_raygen returns it
     in/out  depth_flags [n]          depth (1 = camera segment) | B2_LANE_*
     out     n_shadow, sh_o/sh_d/sh_c [n*4]   the dense emitter-sample queue: ray + (contribution.rgb, lane index bits),
                                       in the (arbitrary) order the warps appended them
   Uses params->max_depth / rr_depth only. */
#define B2_LANE_ALPHA  0x100u    /* out: the camera ray hit something (valid_ray, path.cpp:121)          */
#define B2_LANE_DELTA  0x200u    /* in/out: the last BSDF sample was a delta lobe                          */
#define B2_LANE_PARKED 0x2000u   /* out: Russian roulette or a zero throughput ended the path in this call */
B2_API int b2_stage_shade(B2Scene *scene, const B2RenderParams *params, uint32_t n, float *ray_o, float *ray_d,
                          const float *hit, const int32_t *inst, float *throughput_eta, float *result_pdf,
                          uint64_t *rng_state, uint32_t *depth_flags,
                          uint32_t *n_shadow, float *sh_o, float *sh_d, float *sh_c);
/* compact: K5 alone — the warp-ballot compaction the traversal stage ends with. Lane i takes part if it has a ray
   (ray_maxt[i] >= 0); it goes on the CONTINUE list if it hit something (hit_t[i] < inf) with depth < max_depth (path.cpp:131,147-149:
   the path goes on), otherwise on the TERMINATE list. The lists come back in the order the warps appended them (each warp's
   32 lanes ascending; terminate_lanes as stored from the back of the list array). */
B2_API int b2_stage_compact(B2Scene *scene, int32_t max_depth, uint32_t n, const float *ray_maxt, const float *hit_t, const uint32_t *depth,
                            uint32_t *n_continue, uint32_t *continue_lanes, uint32_t *n_terminate, uint32_t *terminate_lanes);
/* film_put: ImageBlock::put (imageblock.cpp:79-172) of n samples: sample i belongs to the image block of pixel (px, py)
   (crop-relative) and lands at pos[2i..] = position_sample (film coordinates); rgb is converted to XYZ like render_sample
   does (integrator.cpp:273). xyzaw[crop_h * crop_w * 5] receives the accumulated film (box: 5 reductions per sample,
   wider filters: the LUT taps), n_invalid the samples rejected by the validity rule (imageblock.cpp:85-109). */
B2_API int b2_stage_film_put(B2Scene *scene, uint32_t block_size, uint32_t n, const uint32_t *px, const uint32_t *py,
                             const float *pos, const float *rgb, const uint8_t *alpha, float *xyzaw, uint32_t *n_invalid);

/* ---- host-side helpers (C++ mirrors of reference host code; no GPU needed) --------------- */
B2_API int b2_host_look_at(const float origin[3], const float target[3], const float up[3],
                           float to_world[16]);
B2_API int b2_host_perspective_sensor(const float to_world[16], float fov_x_deg, float near_clip,
                                      float far_clip, const B2Film *film, B2Sensor *out);
/* fov given on another axis: 'fov_axis' = "x","y","smaller","larger","diagonal" (sensor.cpp parse_fov) */
B2_API float b2_host_parse_fov(float fov_deg, const char *fov_axis, int width, int height);
B2_API int b2_host_film_box(int width, int height, B2Film *out);                 /* box.cpp      */
B2_API int b2_host_film_gaussian(int width, int height, float stddev, B2Film *out); /* gaussian.cpp */
/* order[i] = (bx, by) of the i-th block handed out by Spiral::next_block */
/* BSDF / shape records from plugin parameters, for hosts that do not link the reference: the values
   SmoothPlastic::parameters_changed() derives (plastic.cpp:162-176: 1/eta^2, fresnel_diffuse_reflectance(1/eta) of
   fresnel.h:327-361, s_mean / (d_mean + s_mean)); Sphere 'center' / 'radius' (sphere.cpp:95-130); Rectangle 'to_world'
   with 'flip_normals' folded in (rectangle.cpp:75-92). to_world: row-major 4x4, affine. */
B2_API int b2_host_bsdf_plastic(float int_ior, float ext_ior, const float diffuse_reflectance[3], float diffuse_mean,
                                const float specular_reflectance[3] /* or NULL */, float specular_mean, int nonlinear, B2Bsdf *out);
B2_API int b2_host_shape_sphere(const float center[3], float radius, int flip_normals, int bsdf, B2Shape *out);
B2_API int b2_host_shape_rectangle(const float to_world[16], int flip_normals, int bsdf, B2Shape *out);
B2_API int b2_host_spiral(int width, int height, int block_size, uint32_t *n_blocks,
                          int32_t *order_xy /* 2 * n_blocks or NULL to query n_blocks */);

/* Debug / inspection: build the BVH4 the library would use for a triangle soup (host only).
   tri_vertices: n * 9 floats. nodes_out: capacity node_cap * 32 floats (layout in
   mitsuba2_b200/csrc/b2_device.cuh); order_out[n]: slot -> triangle. */
B2_API int b2_host_build_bvh4(const float *tri_vertices, uint32_t n, float *nodes_out, uint32_t node_cap,
                              uint32_t *n_nodes, uint32_t *order_out);

#ifdef __cplusplus
}
#endif
#endif /* B2PT_H */
