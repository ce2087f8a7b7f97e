// b2_api.cu — C-ABI implementation (include/b2pt.h): scene flattening + upload, the wavefront
// driver loop, and the stage-level entry points. There is NO CPU fallback: without a CUDA
// device every compute entry point fails with B2_ERR_NO_DEVICE.
#include "../../include/b2pt.h"
#include "b2_bvh.h"
#include "b2_kernels.cuh"

#include <algorithm>
#include <atomic>
#include <chrono>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <string>
#include <utility>
#include <vector>

namespace {

thread_local std::string g_error;

int fail(int code, const char *fmt, ...) {
    char buf[1024];
    va_list ap; va_start(ap, fmt); vsnprintf(buf, sizeof buf, fmt, ap); va_end(ap);
    g_error = buf;
    return code;
}

#define CU(call)                                                                                   \
    do {                                                                                           \
        cudaError_t e_ = (call);                                                                   \
        if (e_ != cudaSuccess)                                                                     \
            return fail(e_ == cudaErrorNoDevice || e_ == cudaErrorInsufficientDriver ? B2_ERR_NO_DEVICE : B2_ERR_CUDA, \
                        "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
    } while (0)

template <typename T> struct DBuf {
    T *p = nullptr; size_t n = 0;
    ~DBuf() { release(); }
    void release() { if (p) cudaFree(p); p = nullptr; n = 0; }
    cudaError_t alloc(size_t count) {
        if (count <= n && p) return cudaSuccess;
        release();
        cudaError_t e = cudaMalloc(&p, std::max<size_t>(count, 1) * sizeof(T));
        if (e == cudaSuccess) n = count;
        return e;
    }
    cudaError_t upload(const std::vector<T> &h) {
        cudaError_t e = alloc(h.size());
        if (e != cudaSuccess || h.empty()) return e;
        return cudaMemcpy(p, h.data(), h.size() * sizeof(T), cudaMemcpyHostToDevice);
    }
};

}  // namespace

// ---- kernel variants by scene features (b2_device.cuh F_*) -----------------------------------------
using ShadeFn = void (*)(DScene, Queues, int);
using TraceFn = void (*)(DScene, Queues, int);
template <uint32_t... I> static void fill_shade(ShadeFn *t, std::integer_sequence<uint32_t, I...>) { ((t[I] = k_shade<I>), ...); }
static ShadeFn shade_variant(uint32_t features) {
    static ShadeFn table[F_COUNT];
    static bool init = false;
    if (!init) { fill_shade(table, std::make_integer_sequence<uint32_t, F_COUNT>{}); init = true; }
    return table[features & (F_COUNT - 1)];
}
static TraceFn trace_variant(bool inst, bool shp, bool count, bool resume = false) {
    if (resume && !inst) {
        if (count) return shp ? k_trace_resume<true, true> : k_trace_resume<false, true>;
        return shp ? k_trace_resume<true, false> : k_trace_resume<false, false>;
    }
    if (count) {
        if (shp) return inst ? k_trace_inst<true, true> : k_trace<false, true, true>;
        return inst ? k_trace_inst<false, true> : k_trace<false, false, true>;
    }
    if (shp) return inst ? k_trace_inst<true, false> : k_trace<false, true, false>;
    return inst ? k_trace_inst<false, false> : k_trace<false, false, false>;
}

// CUDA events with scope lifetime (every early return of the driver releases them)
struct EventPool {
    std::vector<cudaEvent_t> ev;
    ~EventPool() { for (cudaEvent_t e : ev) cudaEventDestroy(e); }
    cudaEvent_t make(bool timing) {
        cudaEvent_t e = nullptr;
        if (cudaEventCreateWithFlags(&e, timing ? cudaEventDefault : cudaEventDisableTiming) != cudaSuccess) return nullptr;
        ev.push_back(e);
        return e;
    }
};

struct B2Scene {
    DScene d{};
    DBuf<float4> nodes, tri_isect, flat_leaves;
    DBuf<DInstance> instances;
    DBuf<float4> env_data;
    DBuf<float> env_levels;
    DBuf<uint32_t> inst_order;
    DBuf<int32_t> hit_inst;
    DBuf<unsigned char> blob;       // tri_pos | tri_nrm | prim_to_slot | cdf | bsdfs | emitters | tri_uv | textures
    DBuf<float> tex_data;           // bitmap texels
    int shade_smem_bytes = 0;
    uint32_t features = 0;          // F_* of this scene
    bool has_shapes = false;        // analytic shapes present (selects the SHP traversal variants)
    uint32_t n_tris = 0, n_nodes = 0, bvh_depth = 0;
    float ms_build = 0.f, ms_upload = 0.f;
    int smem_bytes = 0;
    int sm_count = 148;
    cudaStream_t stream = nullptr;
    cudaStream_t stream2 = nullptr;     // second lane set
    std::atomic<int> cancel{0};
    // render state (allocated lazily, reused across renders)
    DBuf<float4> ray_o, ray_d, hit, tp, res, sh_o, sh_d, sh_c;
    DBuf<uint4> st;
    DBuf<float2> pos;
    DBuf<uint32_t> pixel, list, cnt, item_pixel, item_seed;
    DBuf<uint8_t> cls;
    DBuf<uint32_t> sh_cnt;                                                // counters of the shadow sub-queues (sh_counter())
    DBuf<uint32_t> resume; DBuf<int4> tsave;                              // k_trace_resume: suspended lanes + their traversal state
    DBuf<float4> prim_o, prim_d, prim_hit; DBuf<int32_t> prim_inst;      // DirectIntegrator: the camera segment per lane
    DBuf<unsigned long long> stats;
    DBuf<float> film;
    DBuf<float> film_wide;          // wide filters: padded accumulation film, 8 floats / pixel
    DBuf<uint2> smp;                // per-lane sampler state of the non-independent samplers
    uint32_t *h_cnt = nullptr;      // pinned readback
    // progressive snapshots (b2_set_progress_callback)
    b2_progress_fn progress_fn = nullptr; void *progress_user = nullptr; uint32_t progress_ms = 0;
    cudaStream_t stream_snap = nullptr; float *h_snap = nullptr; size_t h_snap_floats = 0; unsigned long long *h_snap_samples = nullptr;
    DBuf<float> snap_film;          // wide filters: packed copy of the padded film
    // the work-item tables of the last render (spiral order x Morton order) are kept: they depend only on this key
    struct { int W = 0, H = 0; uint32_t bs = 0; int rank = -1, world = 0; uint32_t n_items = 0, n_blocks = 0; } items_key;
    ~B2Scene() {
        if (h_cnt) cudaFreeHost(h_cnt);
        if (h_snap) cudaFreeHost(h_snap);
        if (h_snap_samples) cudaFreeHost(h_snap_samples);
        if (stream2) cudaStreamDestroy(stream2);
        if (stream_snap) cudaStreamDestroy(stream_snap);
    }
};

extern "C" {

const char *b2_last_error(void) { return g_error.c_str(); }

int b2_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

// --------------------------------------------------------------------------------------------
// Scene flattening: what Scene::Scene / accel_init do on the reference side (scene.cpp:22-103)
// --------------------------------------------------------------------------------------------
int b2_scene_create(const B2SceneDesc *desc, B2Scene **out) {
    if (!desc || !out) return fail(B2_ERR_INVALID, "b2_scene_create: null argument");
    *out = nullptr;
    // (the description is validated first, so that malformed input is reported the same way with or without a device;
    //  a VALID scene then fails loudly without a GPU: there is no CPU fallback)
    uint32_t n_top_meshes = desc->n_groups ? desc->n_top_meshes : desc->n_meshes;
    if (n_top_meshes > desc->n_meshes) return fail(B2_ERR_INVALID, "n_top_meshes out of range");
    for (uint32_t g = 0; g < desc->n_groups; ++g)
        if (desc->group_first[g] < n_top_meshes || desc->group_first[g] > desc->group_first[g + 1] || desc->group_first[g + 1] > desc->n_meshes)
            return fail(B2_ERR_INVALID, "shapegroup %u: bad mesh range", g);
    // the groups must cover [n_top_meshes, n_meshes) exactly: every mesh is either a top-level shape or a member of one group
    if (desc->n_groups && (desc->group_first[0] != n_top_meshes || desc->group_first[desc->n_groups] != desc->n_meshes))
        return fail(B2_ERR_INVALID, "shapegroups must cover the meshes [n_top_meshes, n_meshes) exactly");
    for (uint32_t i = 0; i < desc->n_instances; ++i)
        if (desc->instances[i].group < 0 || (uint32_t) desc->instances[i].group >= desc->n_groups)
            return fail(B2_ERR_INVALID, "instance %u: bad shapegroup index", i);
    for (uint32_t m = n_top_meshes; m < desc->n_meshes; ++m)
        if (desc->meshes[m].emitter >= 0) return fail(B2_ERR_INVALID, "Instancing of emitters is not supported");   // shapegroup.cpp:24-25
    for (uint32_t e = 0; e < desc->n_emitters; ++e)
        if (desc->emitters[e].type < B2_EMITTER_AREA || desc->emitters[e].type > B2_EMITTER_DIRECTIONAL)
            return fail(B2_ERR_INVALID, "emitter %u: only area, envmap, constant, point, spot and directional emitters are supported", e);
    {
        int n_env = 0;
        for (uint32_t e = 0; e < desc->n_emitters; ++e) {
            const B2Emitter &em = desc->emitters[e];
            if (em.type != B2_EMITTER_ENVMAP && em.type != B2_EMITTER_CONSTANT) continue;
            if (++n_env > 1) return fail(B2_ERR_INVALID, "Only one environment emitter can be specified per scene.");   // scene.cpp:48-49
            if (em.type == B2_EMITTER_ENVMAP && (!em.envmap || em.env_width < 2 || em.env_height < 2)) return fail(B2_ERR_INVALID, "Distribution2D(): input array resolution must be >= 2!");
        }
    }
    for (uint32_t b = 0; b < desc->n_bsdfs; ++b) {
        const int32_t t = desc->bsdfs[b].type;
        if (t < B2_BSDF_DIFFUSE || t > B2_BSDF_ROUGHDIELECTRIC) return fail(B2_ERR_INVALID, "bsdf %u: unsupported type %d", b, t);
        const bool rough = t == B2_BSDF_ROUGHCONDUCTOR || t == B2_BSDF_ROUGHDIELECTRIC;
        if (rough && desc->bsdfs[b].distribution != B2_MF_BECKMANN && desc->bsdfs[b].distribution != B2_MF_GGX)
            return fail(B2_ERR_INVALID, "bsdf %u: Specified an invalid distribution (must be \"beckmann\" or \"ggx\")", b);   // microfacet.h
        if (rough && !desc->bsdfs[b].sample_visible)
            return fail(B2_ERR_INVALID, "bsdf %u: rough BSDF with sample_visible = false is not supported", b);
    }
    for (uint32_t m = 0; m < desc->n_meshes; ++m) {
        const B2Mesh &me = desc->meshes[m];
        if (me.bsdf < 0 || (uint32_t) me.bsdf >= desc->n_bsdfs) return fail(B2_ERR_INVALID, "mesh %u: bad bsdf index", m);
        if (me.emitter >= (int32_t) desc->n_emitters) return fail(B2_ERR_INVALID, "mesh %u: bad emitter index", m);
        for (size_t f = 0; f < 3 * (size_t) me.n_faces; ++f)
            if (me.faces[f] >= me.n_vertices) return fail(B2_ERR_INVALID, "mesh %u: face index out of range", m);
    }
    const uint32_t ns = desc->shapes ? desc->n_shapes : 0;
    for (uint32_t k = 0; k < ns; ++k) {
        const B2Shape &sh = desc->shapes[k];
        if (sh.type != B2_SHAPE_RECTANGLE && sh.type != B2_SHAPE_SPHERE) return fail(B2_ERR_INVALID, "shape %u: unsupported type %d", k, sh.type);
        if (sh.bsdf < 0 || (uint32_t) sh.bsdf >= desc->n_bsdfs) return fail(B2_ERR_INVALID, "shape %u: bad bsdf index", k);
        if (sh.emitter >= (int32_t) desc->n_emitters) return fail(B2_ERR_INVALID, "shape %u: bad emitter index", k);
        if (sh.type == B2_SHAPE_SPHERE && !(sh.radius > 0.f)) return fail(B2_ERR_INVALID, "shape %u: sphere radius must be positive", k);
    }
    if (desc->film.crop_size[0] <= 0 || desc->film.crop_size[1] <= 0 || desc->film.crop_size[0] > 65535 || desc->film.crop_size[1] > 65535)
        return fail(B2_ERR_INVALID, "film crop size must be in [1, 65535]");

    if (b2_device_count() == 0) return fail(B2_ERR_NO_DEVICE, "b2_scene_create: no CUDA device (libb2pt has no CPU fallback)");
    B2Scene *s = new B2Scene();
    std::vector<uint32_t> mesh_first(desc->n_meshes + 1, 0);
    uint32_t nt = 0;
    for (uint32_t m = 0; m < desc->n_meshes; ++m) { mesh_first[m] = nt; nt += desc->meshes[m].n_faces; }
    mesh_first[desc->n_meshes] = nt;
    if (nt >= (1u << 27)) { delete s; return fail(B2_ERR_INVALID, "too many triangles (%u)", nt); }

    std::vector<B2BvhTri> tris(nt);
    for (uint32_t m = 0; m < desc->n_meshes; ++m) {
        const B2Mesh &me = desc->meshes[m];
        for (uint32_t f = 0; f < me.n_faces; ++f)
            for (int v = 0; v < 3; ++v)
                std::memcpy(tris[mesh_first[m] + f].p[v], me.positions + 3 * (size_t) me.faces[3 * f + v], 12);
    }
    // one BVH over the top-level triangles, one per shapegroup, one over the instances' world bounds;
    // all in one node array (top-level first: it is the part staged into shared memory)
    uint32_t n_top_tris = mesh_first[n_top_meshes];
    struct { std::vector<float> nodes; std::vector<uint32_t> order; uint32_t n_nodes = 0, depth = 0; } bvh;
    uint32_t top_nodes = 0; bool top_bfs = true;
    // builder: host binned SAH, or (B2_BUILD_DEVICE_LBVH) the device LBVH with the host builder as fallback
    auto t_build0 = std::chrono::steady_clock::now();
    float ms_device_build = 0.f; bool used_device = false;
    auto build_tris = [&](const std::vector<B2BvhTri> &sub, B2Bvh4 &b) {
        if ((desc->build_flags & B2_BUILD_DEVICE_LBVH) && sub.size() >= 2) {
            float ms = 0.f; const char *why = nullptr;
            if (b2_build_bvh4_device(&sub[0].p[0][0], (uint32_t) sub.size(), false, 2, b, &ms, &why)) { ms_device_build += ms; used_device = true; return; }
        }
        b2_build_bvh4(sub, b);
    };
    if (ns == 0) {
        std::vector<B2BvhTri> sub(tris.begin(), tris.begin() + n_top_tris);
        B2Bvh4 b; build_tris(sub, b);
        b2_bvh4_append(bvh.nodes, b, 0);
        bvh.order = b.order; bvh.depth = b.depth; top_nodes = b.n_nodes; top_bfs = b.breadth_first;
    } else {
        // analytic shapes share the top-level BVH: primitive boxes = the top-level triangles' boxes, then one box per shape
        // (Rectangle::bbox rectangle.cpp:94-101, Sphere::bbox sphere.cpp:131-136); shape k is primitive nt + k
        std::vector<B2BvhBox> boxes(n_top_tris + ns);
        for (uint32_t i = 0; i < n_top_tris; ++i) {
            B2BvhBox &bx = boxes[i];
            for (int k = 0; k < 3; ++k) { bx.lo[k] = INFINITY; bx.hi[k] = -INFINITY; }
            for (int v = 0; v < 3; ++v) for (int k = 0; k < 3; ++k) { bx.lo[k] = std::min(bx.lo[k], tris[i].p[v][k]); bx.hi[k] = std::max(bx.hi[k], tris[i].p[v][k]); }
        }
        for (uint32_t k = 0; k < ns; ++k) {
            const B2Shape &sh = desc->shapes[k];
            B2BvhBox &bx = boxes[n_top_tris + k];
            for (int q = 0; q < 3; ++q) { bx.lo[q] = INFINITY; bx.hi[q] = -INFINITY; }
            if (sh.type == B2_SHAPE_RECTANGLE) {
                for (int c = 0; c < 4; ++c) for (int r = 0; r < 3; ++r) {
                    const float *M = sh.to_world + 4 * r;
                    double w = (double) M[3] + (double) M[0] * (c & 1 ? 1.0 : -1.0) + (double) M[1] * (c & 2 ? 1.0 : -1.0);
                    bx.lo[r] = std::min(bx.lo[r], std::nextafter((float) w, -INFINITY)); bx.hi[r] = std::max(bx.hi[r], std::nextafter((float) w, INFINITY));
                }
            } else {
                for (int r = 0; r < 3; ++r) {
                    bx.lo[r] = std::nextafter(sh.center[r] - sh.radius, -INFINITY); bx.hi[r] = std::nextafter(sh.center[r] + sh.radius, INFINITY);
                }
            }
            // the analytic tests work on rounded object-space / double quantities: pad by 1e-4 of the box size
            for (int q = 0; q < 3; ++q) { float e = 1e-4f * (std::fabs(bx.lo[q]) + std::fabs(bx.hi[q]) + 1e-3f); bx.lo[q] -= e; bx.hi[q] += e; }
        }
        B2Bvh4 b; b2_build_bvh4_boxes(boxes, 2, b);
        b2_bvh4_append(bvh.nodes, b, 0);
        bvh.order = b.order; bvh.depth = b.depth; top_nodes = b.n_nodes;
        for (uint32_t &o : bvh.order) if (o >= n_top_tris) o = nt + (o - n_top_tris);
    }
    const uint32_t n_slots = nt + ns, n_top_prims = n_top_tris + ns;
    std::vector<uint32_t> group_root(desc->n_groups, 0);
    std::vector<B2BvhBox> group_box(desc->n_groups);
    for (uint32_t g = 0; g < desc->n_groups; ++g) {
        uint32_t t0 = mesh_first[desc->group_first[g]], t1 = mesh_first[desc->group_first[g + 1]];
        std::vector<B2BvhTri> sub(tris.begin() + t0, tris.begin() + t1);
        B2Bvh4 b; build_tris(sub, b);
        group_root[g] = b2_bvh4_append(bvh.nodes, b, (uint32_t) bvh.order.size());
        for (uint32_t o : b.order) bvh.order.push_back(t0 + o);
        bvh.depth = std::max(bvh.depth, b.depth);
        B2BvhBox bx; for (int k = 0; k < 3; ++k) { bx.lo[k] = INFINITY; bx.hi[k] = -INFINITY; }
        for (const B2BvhTri &t : sub) for (int v = 0; v < 3; ++v) for (int k = 0; k < 3; ++k) {
            bx.lo[k] = std::min(bx.lo[k], t.p[v][k]); bx.hi[k] = std::max(bx.hi[k], t.p[v][k]);
        }
        group_box[g] = bx;
    }
    std::vector<DInstance> dinst(desc->n_instances);
    std::vector<uint32_t> inst_order;
    uint32_t tlas_root = 0, tlas_levels = 0;
    if (desc->n_instances) {
        std::vector<B2BvhBox> ib(desc->n_instances);
        for (uint32_t i = 0; i < desc->n_instances; ++i) {
            const B2Instance &in = desc->instances[i];
            std::memcpy(dinst[i].to_world, in.to_world, 48); std::memcpy(dinst[i].to_object, in.to_object, 48);
            dinst[i].root = (int32_t) group_root[in.group];
            const B2BvhBox &gb = group_box[in.group];
            B2BvhBox w; for (int k = 0; k < 3; ++k) { w.lo[k] = INFINITY; w.hi[k] = -INFINITY; }
            for (int c = 0; c < 8; ++c) {                           // Instance::bbox(), instance.cpp:83-94
                double p[3] = { c & 1 ? gb.hi[0] : gb.lo[0], c & 2 ? gb.hi[1] : gb.lo[1], c & 4 ? gb.hi[2] : gb.lo[2] };
                for (int r = 0; r < 3; ++r) {
                    const float *M = in.to_world + 4 * r;
                    double v = (double) M[3] + (double) M[0] * p[0] + (double) M[1] * p[1] + (double) M[2] * p[2];
                    float lo = (float) v, hi = (float) v;
                    lo = std::nextafter(lo, -INFINITY); hi = std::nextafter(hi, INFINITY);
                    w.lo[r] = std::min(w.lo[r], lo); w.hi[r] = std::max(w.hi[r], hi);
                }
            }
            // the object-space triangle test sees rounded coordinates: pad by 1e-4 of the box size
            for (int k = 0; k < 3; ++k) { float e = 1e-4f * (std::fabs(w.lo[k]) + std::fabs(w.hi[k]) + 1e-3f); w.lo[k] -= e; w.hi[k] += e; }
            ib[i] = w;
        }
        B2Bvh4 t; b2_build_bvh4_boxes(ib, 1, t);
        tlas_root = b2_bvh4_append(bvh.nodes, t, 0);
        inst_order = t.order;
        bvh.depth += t.depth; tlas_levels = t.depth;
    }
    bvh.n_nodes = (uint32_t) (bvh.nodes.size() / 32);
    s->ms_build = used_device ? ms_device_build : std::chrono::duration<float, std::milli>(std::chrono::steady_clock::now() - t_build0).count();
    auto t_upload0 = std::chrono::steady_clock::now();
    {   // the traversal stacks are fixed-size local arrays (B2_STACK entries per BVH, B2_INST_STACK for k_trace_inst's
        // single two-level stack); a BVH4 level pushes at most 3 entries. The builders bound the depth (b2_bvh.cpp:
        // median splits once the level budget runs short), so this is a consistency check, not a limit on the input.
        const uint32_t tlas_depth = desc->n_instances ? tlas_levels : 0, blas_depth = bvh.depth - tlas_depth;
        if (3 * blas_depth + 2 > B2_STACK || (desc->n_instances && (3 * tlas_depth + 2 > B2_STACK || 3 * (tlas_depth + blas_depth) + 4 > B2_INST_STACK))) {
            delete s;
            return fail(B2_ERR_INVALID, "BVH too deep for the traversal stack (%u + %u levels)", blas_depth, tlas_depth);
        }
    }

    if (bvh.order.size() != n_slots) { delete s; return fail(B2_ERR_INVALID, "internal: BVH leaf order covers %zu of %u primitives", bvh.order.size(), n_slots); }
    std::vector<float4> tri_pos(3 * (size_t) n_slots), tri_nrm(3 * (size_t) n_slots), tri_isect(3 * (size_t) n_slots);
    std::vector<uint32_t> prim_to_slot(n_slots);
    bool any_uv = false;
    for (uint32_t m = 0; m < desc->n_meshes; ++m) any_uv = any_uv || desc->meshes[m].texcoords != nullptr;
    std::vector<float4> tri_uv(any_uv ? 2 * (size_t) n_slots : 0);
    auto u2f = [](uint32_t u) { float f; std::memcpy(&f, &u, 4); return f; };
    for (uint32_t slot = 0; slot < n_slots; ++slot) {
        uint32_t prim = bvh.order[slot];
        prim_to_slot[prim] = slot;
        if (prim >= nt) {        // analytic shape: no vertices, (kind, shape index) words instead
            const uint32_t k = prim - nt;
            const B2Shape &sh = desc->shapes[k];
            const uint32_t kind = 1u + (uint32_t) sh.type;
            tri_isect[3 * (size_t) slot] = make_float4(0.f, 0.f, 0.f, u2f(prim));
            tri_isect[3 * (size_t) slot + 1] = make_float4(0.f, 0.f, 0.f, u2f(kind));
            tri_isect[3 * (size_t) slot + 2] = make_float4(0.f, 0.f, 0.f, u2f(k));
            tri_pos[3 * (size_t) slot] = make_float4(0.f, 0.f, 0.f, u2f(prim));
            tri_pos[3 * (size_t) slot + 1] = make_float4(0.f, 0.f, 0.f, u2f(0xffffffffu));
            tri_pos[3 * (size_t) slot + 2] = make_float4(0.f, 0.f, 0.f, u2f(k));
            tri_nrm[3 * (size_t) slot] = make_float4(0.f, 0.f, 0.f, u2f((uint32_t) sh.bsdf));
            tri_nrm[3 * (size_t) slot + 1] = make_float4(0.f, 0.f, 0.f, u2f((uint32_t) sh.emitter));
            tri_nrm[3 * (size_t) slot + 2] = make_float4(0.f, 0.f, 0.f, u2f(1u + kind));      // 2 + type
            if (any_uv) { tri_uv[2 * (size_t) slot] = make_float4(0.f, 0.f, 0.f, 0.f); tri_uv[2 * (size_t) slot + 1] = make_float4(0.f, 0.f, 0.f, 0.f); }
            continue;
        }
        uint32_t m = (uint32_t) (std::upper_bound(mesh_first.begin(), mesh_first.end(), prim) - mesh_first.begin()) - 1;
        const B2Mesh &me = desc->meshes[m];
        uint32_t f = prim - mesh_first[m];
        const float wpos[3] = { u2f(prim), u2f(m), 0.f };
        const float wnrm[3] = { u2f((uint32_t) me.bsdf), u2f((uint32_t) me.emitter), u2f(me.normals ? 1u : 0u) };
        {
            const float *q0 = me.positions + 3 * (size_t) me.faces[3 * f], *q1 = me.positions + 3 * (size_t) me.faces[3 * f + 1],
                        *q2 = me.positions + 3 * (size_t) me.faces[3 * f + 2];
            tri_isect[3 * (size_t) slot] = make_float4(q0[0], q0[1], q0[2], wpos[0]);
            tri_isect[3 * (size_t) slot + 1] = make_float4(q1[0] - q0[0], q1[1] - q0[1], q1[2] - q0[2], 0.f);
            tri_isect[3 * (size_t) slot + 2] = make_float4(q2[0] - q0[0], q2[1] - q0[1], q2[2] - q0[2], 0.f);
        }
        if (any_uv) {        // vertex texture coordinates (mesh.cpp:492-497): (uv0, uv1), (uv2, has_uv, -)
            if (me.texcoords) {
                const float *t0 = me.texcoords + 2 * (size_t) me.faces[3 * f], *t1 = me.texcoords + 2 * (size_t) me.faces[3 * f + 1],
                            *t2 = me.texcoords + 2 * (size_t) me.faces[3 * f + 2];
                tri_uv[2 * (size_t) slot] = make_float4(t0[0], t0[1], t1[0], t1[1]);
                tri_uv[2 * (size_t) slot + 1] = make_float4(t2[0], t2[1], u2f(1u), 0.f);
            } else { tri_uv[2 * (size_t) slot] = make_float4(0.f, 0.f, 0.f, 0.f); tri_uv[2 * (size_t) slot + 1] = make_float4(0.f, 0.f, 0.f, 0.f); }
        }
        for (int v = 0; v < 3; ++v) {
            const float *p = me.positions + 3 * (size_t) me.faces[3 * f + v];
            tri_pos[3 * (size_t) slot + v] = make_float4(p[0], p[1], p[2], wpos[v]);
            if (me.normals) {
                const float *n = me.normals + 3 * (size_t) me.faces[3 * f + v];
                tri_nrm[3 * (size_t) slot + v] = make_float4(n[0], n[1], n[2], wnrm[v]);
            } else
                tri_nrm[3 * (size_t) slot + v] = make_float4(0.f, 0.f, 0.f, wnrm[v]);
        }
    }
    // emitters: Mesh::build_pmf (mesh.cpp:285-312) + DiscreteDistribution::update (distr_1d.h:49-90)
    std::vector<DEmitter> ems(desc->n_emitters);
    std::vector<float> cdf;
    for (uint32_t e = 0; e < desc->n_emitters; ++e) {
        const B2Emitter &src = desc->emitters[e];
        if (src.type != B2_EMITTER_AREA) {
            std::memset(&ems[e], 0, sizeof(DEmitter)); ems[e].type = src.type; ems[e].mesh = -1; std::memcpy(ems[e].radiance, src.radiance, 12);
            if (src.type == B2_EMITTER_POINT || src.type == B2_EMITTER_SPOT) {
                ems[e].position[0] = src.env_to_world[3]; ems[e].position[1] = src.env_to_world[7]; ems[e].position[2] = src.env_to_world[11]; s->features |= F_TEX;
            }
            if (src.type == B2_EMITTER_SPOT) {                            // SpotLight::SpotLight, spot.cpp:88-95
                std::memcpy(ems[e].to_object, src.env_to_object, 48);
                ems[e].cutoff_angle = src.cutoff_angle; ems[e].beam_width = src.beam_width;
                ems[e].cos_cutoff = std::cos(src.cutoff_angle); ems[e].cos_beam = std::cos(src.beam_width);
                ems[e].inv_transition = 1.0f / (src.cutoff_angle - src.beam_width);
            }
            if (src.type == B2_EMITTER_DIRECTIONAL) {                     // d = to_world * (0,0,1), directional.cpp:113
                const float *M = src.env_to_world;
                for (int r = 0; r < 3; ++r) ems[e].position[r] = std::fma(M[4 * r + 2], 1.f, std::fma(M[4 * r + 1], 0.f, M[4 * r] * 0.f));
                s->features |= F_TEX;
            }
            continue;
        }
        if (src.shape > 0) {          // area emitter on an analytic shape
            if ((uint32_t) src.shape > ns) { delete s; return fail(B2_ERR_INVALID, "emitter %u: bad shape index", e); }
            std::memset(&ems[e], 0, sizeof(DEmitter));
            ems[e].type = src.type; ems[e].mesh = -1; ems[e].shape = src.shape;
            std::memcpy(ems[e].radiance, src.radiance, 12);
            continue;
        }
        if (src.mesh < 0 || (uint32_t) src.mesh >= desc->n_meshes) { delete s; return fail(B2_ERR_INVALID, "emitter %u: bad mesh index", e); }
        const B2Mesh &me = desc->meshes[src.mesh];
        if (me.n_faces == 0) { delete s; return fail(B2_ERR_INVALID, "Cannot create sampling table for an empty mesh"); }
        DEmitter &d = ems[e];
        d.type = src.type; d.mesh = src.mesh;
        std::memcpy(d.radiance, src.radiance, 12);
        d.first_prim = mesh_first[src.mesh]; d.n_faces = me.n_faces; d.cdf_offset = (uint32_t) cdf.size();
        double sum = 0.0; d.valid_first = 0xffffffffu; d.valid_last = 0;
        for (uint32_t f = 0; f < me.n_faces; ++f) {
            const B2BvhTri &t = tris[d.first_prim + f];
            float e0[3], e1[3];
            for (int k = 0; k < 3; ++k) { e0[k] = t.p[1][k] - t.p[0][k]; e1[k] = t.p[2][k] - t.p[0][k]; }
            float c[3] = { std::fma(e0[1], e1[2], -(e0[2] * e1[1])), std::fma(e0[2], e1[0], -(e0[0] * e1[2])),
                           std::fma(e0[0], e1[1], -(e0[1] * e1[0])) };
            float area = 0.5f * std::sqrt((c[0] * c[0] + c[1] * c[1]) + c[2] * c[2]);   // mesh.h:108-116
            sum += (double) area;
            cdf.push_back((float) sum);
            if (area > 0.f) { if (d.valid_first == 0xffffffffu) d.valid_first = f; d.valid_last = f; }
        }
        if (d.valid_first == 0xffffffffu) { delete s; return fail(B2_ERR_INVALID, "DiscreteDistribution: no probability mass found!"); }
        d.cdf_sum = (float) sum; d.cdf_norm = (float) (1.0 / sum);
    }
    // analytic shapes: the values Rectangle::update() / Sphere::update() derive (rectangle.cpp:82-92, sphere.cpp:107-130)
    std::vector<DShape> dshapes(ns);
    for (uint32_t k = 0; k < ns; ++k) {
        const B2Shape &sh = desc->shapes[k];
        DShape &ds = dshapes[k];
        std::memset(&ds, 0, sizeof ds);
        ds.type = sh.type; ds.flip_normals = sh.flip_normals; ds.bsdf = sh.bsdf; ds.emitter = sh.emitter;
        std::memcpy(ds.to_world, sh.to_world, 48); std::memcpy(ds.to_object, sh.to_object, 48);
        std::memcpy(ds.center, sh.center, 12); ds.radius = sh.radius;
        if (sh.type == B2_SHAPE_RECTANGLE) {
            const float *M = sh.to_world, *I = sh.to_object;
            // the same fmadd chains as Transform::operator*(Vector / Normal) (transform.h:88-151), so that even the signs of
            // zero components match: to_world * (2,0,0), to_world * (0,2,0), inverse transpose * (0,0,1)
            auto xfv = [&](int r, float x, float y, float z) { return std::fma(M[4 * r + 2], z, std::fma(M[4 * r + 1], y, M[4 * r] * x)); };
            auto xfn = [&](int c, float x, float y, float z) { return std::fma(I[8 + c], z, std::fma(I[4 + c], y, I[c] * x)); };
            float fs[3] = { xfv(0, 2.f, 0.f, 0.f), xfv(1, 2.f, 0.f, 0.f), xfv(2, 2.f, 0.f, 0.f) },
                  ft[3] = { xfv(0, 0.f, 2.f, 0.f), xfv(1, 0.f, 2.f, 0.f), xfv(2, 0.f, 2.f, 0.f) };
            float n[3] = { xfn(0, 0.f, 0.f, 1.f), xfn(1, 0.f, 0.f, 1.f), xfn(2, 0.f, 0.f, 1.f) };
            float inv_len = 1.f / std::sqrt((n[0] * n[0] + n[1] * n[1]) + n[2] * n[2]);
            for (int q = 0; q < 3; ++q) { ds.n[q] = n[q] * inv_len; ds.dp_du[q] = fs[q]; }
            float c[3] = { std::fma(fs[1], ft[2], -(fs[2] * ft[1])), std::fma(fs[2], ft[0], -(fs[0] * ft[2])), std::fma(fs[0], ft[1], -(fs[1] * ft[0])) };
            ds.inv_area = 1.f / std::sqrt((c[0] * c[0] + c[1] * c[1]) + c[2] * c[2]);
        } else {
            ds.inv_area = 1.f / (4.f * 3.14159265358979323846f * sh.radius * sh.radius);
        }
    }
    std::vector<B2Bsdf> bsdfs(desc->bsdfs, desc->bsdfs + desc->n_bsdfs);
    // textures (checkerboard.cpp, bitmap.cpp): texel data of all bitmaps in one device buffer
    std::vector<B2Texture> texs(desc->textures, desc->textures + (desc->textures ? desc->n_textures : 0));
    std::vector<float> tex_data; std::vector<size_t> tex_off(texs.size(), 0);
    for (size_t t = 0; t < texs.size(); ++t) {
        const B2Texture &tx = texs[t];
        if (tx.type == B2_TEX_BITMAP) {
            if (!tx.data || tx.width < 1 || tx.height < 1 || (tx.channels != 1 && tx.channels != 3)) {
                delete s; return fail(B2_ERR_INVALID, "texture %zu: Unsupported channel count: %d (expected 1 or 3)", t, tx.channels); }
            tex_off[t] = tex_data.size();
            tex_data.insert(tex_data.end(), tx.data, tx.data + (size_t) tx.width * tx.height * tx.channels);
            while (tex_data.size() % 4) tex_data.push_back(0.f);
        } else if (tx.type != B2_TEX_CHECKERBOARD) { delete s; return fail(B2_ERR_INVALID, "texture %zu: unsupported type %d", t, tx.type); }
    }
    for (const B2Bsdf &b : bsdfs)
        if (b.reflectance_texture < 0 || (uint32_t) b.reflectance_texture > texs.size()) { delete s; return fail(B2_ERR_INVALID, "BSDF references texture %d of %zu", b.reflectance_texture, texs.size()); }
    if (!tex_data.empty()) {
        cudaError_t te = s->tex_data.upload(tex_data);
        if (te != cudaSuccess) { delete s; return fail(B2_ERR_CUDA, "texture upload failed: %s", cudaGetErrorString(te)); }
        for (size_t t = 0; t < texs.size(); ++t) texs[t].data = texs[t].type == B2_TEX_BITMAP ? s->tex_data.p + tex_off[t] : nullptr;
    }
    std::vector<float4> nodes(bvh.nodes.size() / 4);
    std::memcpy(nodes.data(), bvh.nodes.data(), bvh.nodes.size() * sizeof(float));

    // one blob for the shading tables (so that small scenes can be staged with one bulk copy)
    auto align16 = [](size_t x) { return (x + 15) & ~(size_t) 15; };
    size_t off_pos = 0, off_nrm = off_pos + align16(tri_pos.size() * 16), off_p2s = off_nrm + align16(tri_nrm.size() * 16),
           off_cdf = off_p2s + align16(prim_to_slot.size() * 4), off_bsdf = off_cdf + align16(cdf.size() * 4),
           off_em = off_bsdf + align16(bsdfs.size() * sizeof(B2Bsdf)), off_uv = off_em + align16(ems.size() * sizeof(DEmitter)),
           off_tex = off_uv + align16(tri_uv.size() * 16), off_shp = off_tex + align16(texs.size() * sizeof(B2Texture)),
           blob_bytes = off_shp + align16(dshapes.size() * sizeof(DShape));
    std::vector<unsigned char> blob(std::max<size_t>(blob_bytes, 16), 0);
    if (!tri_pos.empty()) std::memcpy(blob.data() + off_pos, tri_pos.data(), tri_pos.size() * 16);
    if (!tri_nrm.empty()) std::memcpy(blob.data() + off_nrm, tri_nrm.data(), tri_nrm.size() * 16);
    if (!prim_to_slot.empty()) std::memcpy(blob.data() + off_p2s, prim_to_slot.data(), prim_to_slot.size() * 4);
    if (!cdf.empty()) std::memcpy(blob.data() + off_cdf, cdf.data(), cdf.size() * 4);
    if (!bsdfs.empty()) std::memcpy(blob.data() + off_bsdf, bsdfs.data(), bsdfs.size() * sizeof(B2Bsdf));
    if (!ems.empty()) std::memcpy(blob.data() + off_em, ems.data(), ems.size() * sizeof(DEmitter));
    if (!tri_uv.empty()) std::memcpy(blob.data() + off_uv, tri_uv.data(), tri_uv.size() * 16);
    if (!texs.empty()) std::memcpy(blob.data() + off_tex, texs.data(), texs.size() * sizeof(B2Texture));
    if (!dshapes.empty()) std::memcpy(blob.data() + off_shp, dshapes.data(), dshapes.size() * sizeof(DShape));
    cudaError_t e = cudaSuccess;
    if ((e = s->nodes.upload(nodes)) != cudaSuccess || (e = s->tri_isect.upload(tri_isect)) != cudaSuccess ||
        (e = s->blob.upload(blob)) != cudaSuccess || (e = s->instances.upload(dinst)) != cudaSuccess ||
        (e = s->inst_order.upload(inst_order)) != cudaSuccess) {
        delete s;
        return fail(B2_ERR_CUDA, "scene upload failed: %s", cudaGetErrorString(e));
    }
    s->n_tris = nt; s->n_nodes = bvh.n_nodes; s->bvh_depth = bvh.depth;
    DScene &d = s->d;
    d.nodes = s->nodes.p; d.tri_isect = s->tri_isect.p;
    d.blob = s->blob.p; d.blob_bytes = (uint32_t) blob_bytes;
    d.tri_pos = (const float4 *) (s->blob.p + off_pos); d.tri_nrm = (const float4 *) (s->blob.p + off_nrm);
    d.prim_to_slot = (const uint32_t *) (s->blob.p + off_p2s); d.cdf = (const float *) (s->blob.p + off_cdf);
    d.bsdfs = (const B2Bsdf *) (s->blob.p + off_bsdf); d.emitters = (const DEmitter *) (s->blob.p + off_em);
    d.tri_uv = tri_uv.empty() ? nullptr : (const float4 *) (s->blob.p + off_uv);
    d.textures = texs.empty() ? nullptr : (const B2Texture *) (s->blob.p + off_tex);
    d.shapes = dshapes.empty() ? nullptr : (const DShape *) (s->blob.p + off_shp);
    if (any_uv || !texs.empty() || ns) s->features |= F_TEX;
    d.shade_smem = blob_bytes > 0 && blob_bytes <= 48 * 1024 ? 1 : 0;
    s->shade_smem_bytes = d.shade_smem ? (int) blob_bytes + 16 : 0;
    for (const B2Bsdf &b : bsdfs) {
        if (b.twosided_back != 0) {                                            // twosided.cpp:63-92
            if (b.twosided_back < 0 || (uint32_t) b.twosided_back > bsdfs.size()) { delete s; return fail(B2_ERR_INVALID, "twosided BSDF: bad back-side index"); }
            if (b.type == B2_BSDF_DIELECTRIC || bsdfs[b.twosided_back - 1].type == B2_BSDF_DIELECTRIC) {
                delete s; return fail(B2_ERR_INVALID, "Only materials without a transmission component can be nested!"); }
            s->features |= F_TEX;
        }
        if (b.type == B2_BSDF_ROUGHCONDUCTOR || b.type == B2_BSDF_ROUGHDIELECTRIC) s->features |= F_ROUGH;
        if (b.type == B2_BSDF_DIELECTRIC || b.type == B2_BSDF_CONDUCTOR || b.type == B2_BSDF_PLASTIC) s->features |= F_DIELECTRIC;
    }
    if (desc->film.filter_radius > 0.5f + 1500.0f * 5.9604644775390625e-08f) s->features |= F_WIDE;
    d.instances = desc->n_instances ? s->instances.p : nullptr; d.inst_order = s->inst_order.p;
    d.n_instances = desc->n_instances; d.tlas_root = tlas_root; d.n_top_tris = n_top_prims;
    auto scene_bsphere_radius = [&]() -> float {
        // scene bbox = union of the top-level shapes' boxes (scene.cpp:35-39); instances: to_world * corner (instance.cpp:83-94)
        float lo[3] = { INFINITY, INFINITY, INFINITY }, hi[3] = { -INFINITY, -INFINITY, -INFINITY };
        for (uint32_t m = 0; m < n_top_meshes; ++m)
            for (uint32_t vtx = 0; vtx < desc->meshes[m].n_vertices; ++vtx)
                for (int k = 0; k < 3; ++k) { float c = desc->meshes[m].positions[3 * (size_t) vtx + k]; lo[k] = std::min(lo[k], c); hi[k] = std::max(hi[k], c); }
        for (uint32_t i = 0; i < desc->n_instances; ++i) {
            const B2Instance &in2 = desc->instances[i];
            const B2BvhBox &gb = group_box[in2.group];
            if (!(gb.lo[0] <= gb.hi[0])) continue;
            for (int c = 0; c < 8; ++c) {
                float p[3] = { c & 1 ? gb.hi[0] : gb.lo[0], c & 2 ? gb.hi[1] : gb.lo[1], c & 4 ? gb.hi[2] : gb.lo[2] };
                for (int r = 0; r < 3; ++r) {
                    const float *M = in2.to_world + 4 * r;
                    float w = std::fma(M[2], p[2], std::fma(M[1], p[1], std::fma(M[0], p[0], M[3])));
                    lo[r] = std::min(lo[r], w); hi[r] = std::max(hi[r], w);
                }
            }
        }
        for (uint32_t k = 0; k < ns; ++k) {                            // Rectangle::bbox rectangle.cpp:94-101, Sphere::bbox sphere.cpp:131-136
            const B2Shape &sh = desc->shapes[k];
            for (int cn = 0; cn < 4; ++cn)
                for (int r = 0; r < 3; ++r) {
                    float w;
                    if (sh.type == B2_SHAPE_RECTANGLE) {
                        const float *M = sh.to_world + 4 * r;
                        w = std::fma(M[2], 0.f, std::fma(M[1], cn & 2 ? 1.f : -1.f, std::fma(M[0], cn & 1 ? 1.f : -1.f, M[3])));
                    } else w = cn & 1 ? sh.center[r] - sh.radius : sh.center[r] + sh.radius;
                    lo[r] = std::min(lo[r], w); hi[r] = std::max(hi[r], w);
                }
        }
        float c[3] = { (lo[0] + hi[0]) * .5f, (lo[1] + hi[1]) * .5f, (lo[2] + hi[2]) * .5f };
        float dd[3] = { c[0] - hi[0], c[1] - hi[1], c[2] - hi[2] };
        float radius = std::sqrt((dd[0] * dd[0] + dd[1] * dd[1]) + dd[2] * dd[2]);     // bbox.h:329-332
        const float ray_eps = 1500.0f * 5.9604644775390625e-08f;
        return std::max(ray_eps, radius * (1.f + ray_eps));
    };
    // environment map: Hierarchical2D levels (distr_2d.h:372-471) + scene bounding sphere (envmap.cpp:127-131)
    d.env.index = -1;
    for (uint32_t e = 0; e < desc->n_emitters; ++e) if (desc->emitters[e].type == B2_EMITTER_ENVMAP || desc->emitters[e].type == B2_EMITTER_CONSTANT) s->features |= F_ENV;
    for (uint32_t e = 0; e < desc->n_emitters; ++e) {
        const B2Emitter &em = desc->emitters[e];
        if (em.type != B2_EMITTER_ENVMAP && em.type != B2_EMITTER_CONSTANT) continue;
        DEnv &v = d.env;
        v.index = (int32_t) e; v.type = em.type; std::memcpy(v.radiance, em.radiance, 12);
        if (em.type == B2_EMITTER_ENVMAP) {
        uint32_t W = (uint32_t) em.env_width, H = (uint32_t) em.env_height;
        std::vector<float> lum((size_t) W * H);
        for (uint32_t y = 0; y < H; ++y) {                            // envmap.cpp:86-117
            float sin_theta = std::sin((float) y / (float) (H - 1) * 3.14159265358979323846f);
            for (uint32_t x = 0; x < W; ++x) {
                const float *p = em.envmap + 4 * ((size_t) y * W + x);
                float l = p[0] * 0.212671f + p[1] * 0.715160f + p[2] * 0.072169f;
                lum[(size_t) y * W + x] = l * sin_theta;
            }
        }
        auto lvl_index = [](uint32_t x, uint32_t y, uint32_t width) { return ((x & 1u) | (((x & ~1u) | (y & 1u)) << 1)) + ((y & ~1u) * width); };
        uint32_t npx = W - 1, npy = H - 1, mx = std::max(npx, npy), max_level = 0;
        while ((1u << max_level) < mx) ++max_level;
        int nl = (int) max_level + 2;
        if (nl > 16) { delete s; return fail(B2_ERR_INVALID, "environment map too large"); }
        std::vector<uint32_t> lw(nl), lh(nl), loff(nl);
        lw[0] = W; lh[0] = H;
        { uint32_t lx = npx, ly = npy; for (int level = (int) max_level, k = 1; level >= 0; --level, ++k) { lx += lx & 1u; ly += ly & 1u; lw[k] = lx; lh[k] = ly; lx >>= 1; ly >>= 1; } }
        size_t total = 0;
        for (int k = 0; k < nl; ++k) { loff[k] = (uint32_t) total; total += ((size_t) lw[k] * lh[k] + 3) & ~(size_t) 3; }   // keep 16-byte alignment
        std::vector<float> lv(total, 0.f);
        double sum = 0.0;
        const float *in = lum.data();
        for (uint32_t y = 0; y < npy; ++y) {
            for (uint32_t x = 0; x < npx; ++x) {
                float avg = (in[0] + in[1] + in[W] + in[W + 1]) * .25f;
                sum += (double) avg;
                lv[loff[1] + lvl_index(x, y, lw[1])] = avg;
                ++in;
            }
            ++in;
        }
        float scale = (float) ((double) (npx * npy) / sum);
        for (size_t i = 0; i < (size_t) W * H; ++i) lv[loff[0] + i] = lum[i] * scale;
        for (size_t i = 0; i < (size_t) lw[1] * lh[1]; ++i) lv[loff[1] + i] *= scale;
        { uint32_t lx = npx, ly = npy;
          for (int level = 2; level <= (int) max_level + 1; ++level) {
              lx = (lx + 1u) >> 1; ly = (ly + 1u) >> 1;
              for (uint32_t y = 0; y < ly; ++y) for (uint32_t x = 0; x < lx; ++x) {
                  const float *d0 = &lv[loff[level - 1] + lvl_index(x * 2, y * 2, lw[level - 1])];
                  lv[loff[level] + lvl_index(x, y, lw[level])] = d0[0] + d0[1] + d0[2] + d0[3];
              }
          } }
        std::vector<float4> texels((size_t) W * H);
        std::memcpy(texels.data(), em.envmap, (size_t) W * H * 16);
        if (s->env_data.upload(texels) != cudaSuccess || s->env_levels.upload(lv) != cudaSuccess) { delete s; return fail(B2_ERR_CUDA, "envmap upload failed"); }
        v.data = s->env_data.p; v.levels = s->env_levels.p; v.width = (int32_t) W; v.height = (int32_t) H; v.n_levels = nl;
        for (int k = 0; k < nl; ++k) { v.level_off[k] = loff[k]; v.level_w[k] = lw[k]; }
        v.scale = em.scale;
        std::memcpy(v.to_world, em.env_to_world, 48); std::memcpy(v.to_object, em.env_to_object, 48);
        }   // (constant background emitter, constant.cpp: only the bounding sphere below)
        v.radius = scene_bsphere_radius();
    }
    // DirectionalEmitter::set_scene (directional.cpp:64-69) needs the bounding sphere too
    if (d.env.index < 0)
        for (uint32_t e = 0; e < desc->n_emitters; ++e)
            if (desc->emitters[e].type == B2_EMITTER_DIRECTIONAL) { d.env.radius = scene_bsphere_radius(); break; }
    d.n_nodes = bvh.n_nodes; d.n_tris = n_slots; d.n_emitters = desc->n_emitters; d.n_bsdfs = desc->n_bsdfs;
    d.sensor = desc->sensor; d.film = desc->film;
    // shared-memory staging plan: whole scene if it fits in 48 KB, else the top of the tree (32 KB)
    const uint32_t budget_all = 48 * 1024, budget_top = 32 * 1024;
    if (bvh.n_nodes * 128u + n_slots * 48u <= budget_all) { d.smem_nodes = bvh.n_nodes; d.smem_tris = 1; }
    else { d.smem_nodes = top_bfs ? std::min<uint32_t>(top_nodes, budget_top / 128u) : 0u; d.smem_tris = 0; }   // (a prefix of a breadth-first numbering)
    // tiny flat scenes: list the BVH leaves (box + code) for the exhaustive leaf scan (traverse_flat)
    d.n_flat_leaves = 0; d.flat_leaves = nullptr;
    if (d.smem_tris && desc->n_instances == 0 && nt > 0 && ns == 0) {
        std::vector<float4> flat;
        for (uint32_t n = 0; n < bvh.n_nodes; ++n)
            for (int k = 0; k < 4; ++k) {
                const float *N = &bvh.nodes[32 * (size_t) n];
                int32_t code; std::memcpy(&code, &N[24 + k], 4);
                if (code < 0) {
                    float cf; std::memcpy(&cf, &code, 4);
                    flat.push_back(make_float4(N[0 + k], N[4 + k], N[8 + k], cf));
                    flat.push_back(make_float4(N[12 + k], N[16 + k], N[20 + k], 0.f));
                }
            }
        if (flat.size() / 2 <= 32 && s->flat_leaves.upload(flat) == cudaSuccess) { d.n_flat_leaves = (uint32_t) (flat.size() / 2); d.flat_leaves = s->flat_leaves.p; }
    }
    s->smem_bytes = 16 + (int) (d.smem_nodes * 128u + (d.smem_tris ? n_slots * 48u : 0u) + d.n_flat_leaves * 32u);
    int dev = 0; cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&s->sm_count, cudaDevAttrMultiProcessorCount, dev);
    s->has_shapes = ns != 0;
    s->ms_upload = std::chrono::duration<float, std::milli>(std::chrono::steady_clock::now() - t_upload0).count();
    // (cudaFuncAttributeMaxDynamicSharedMemorySize is a per-function, process-wide attribute: it is set right before the
    //  launches of b2_render_device / b2_stage_*, not here, so that several live scenes cannot clobber each other's limit)
    *out = s;
    return B2_OK;
}

int b2_scene_destroy(B2Scene *scene) {
    if (!scene) return B2_OK;
    cudaDeviceSynchronize();
    delete scene;
    return B2_OK;
}

// Scene::parameters_changed -> accel_parameters_changed (scene.cpp:242-260; the OptiX backend rebuilds its acceleration structure,
// scene_optix.inl:236): the new description is flattened and built like in b2_scene_create (with B2_BUILD_DEVICE_LBVH the
// rebuild is a few ms for millions of triangles) and swapped in under the same handle; the wavefront queues, the streams, the
// pinned buffers, the work-item tables and the callbacks are kept. On failure the handle keeps its old contents.
int b2_scene_update(B2Scene *scene, const B2SceneDesc *desc) {
    if (!scene) return fail(B2_ERR_INVALID, "null scene");
    B2Scene *fresh = nullptr;
    int rc = b2_scene_create(desc, &fresh);
    if (rc != B2_OK) return rc;
    cudaStreamSynchronize(scene->stream);           // no render of the old geometry may still be in flight
    if (scene->stream2) cudaStreamSynchronize(scene->stream2);
    B2Scene &a = *scene, &b = *fresh;
    std::swap(a.d, b.d);
    auto sw = [](auto &x, auto &y) { std::swap(x.p, y.p); std::swap(x.n, y.n); };
    sw(a.nodes, b.nodes); sw(a.tri_isect, b.tri_isect); sw(a.flat_leaves, b.flat_leaves); sw(a.instances, b.instances);
    sw(a.env_data, b.env_data); sw(a.env_levels, b.env_levels); sw(a.inst_order, b.inst_order); sw(a.blob, b.blob); sw(a.tex_data, b.tex_data);
    std::swap(a.shade_smem_bytes, b.shade_smem_bytes); std::swap(a.features, b.features); std::swap(a.has_shapes, b.has_shapes);
    std::swap(a.n_tris, b.n_tris); std::swap(a.n_nodes, b.n_nodes); std::swap(a.bvh_depth, b.bvh_depth);
    std::swap(a.ms_build, b.ms_build); std::swap(a.ms_upload, b.ms_upload); std::swap(a.smem_bytes, b.smem_bytes);
    delete fresh;                                   // releases the old geometry
    return B2_OK;
}

int b2_set_stream(B2Scene *scene, void *stream) {
    if (!scene) return fail(B2_ERR_INVALID, "null scene");
    scene->stream = (cudaStream_t) stream;
    return B2_OK;
}

int b2_set_progress_callback(B2Scene *scene, b2_progress_fn fn, void *user, uint32_t min_interval_ms) {
    if (!scene) return fail(B2_ERR_INVALID, "null scene");
    scene->progress_fn = fn; scene->progress_user = user; scene->progress_ms = min_interval_ms;
    return B2_OK;
}

int b2_cancel(B2Scene *scene) {
    if (scene) scene->cancel.store(1);
    return B2_OK;
}

// --------------------------------------------------------------------------------------------
// SamplingIntegrator::render (integrator.cpp:51-179) as a wavefront loop
// --------------------------------------------------------------------------------------------
// Work-item table of one pass: this rank's blocks in spiral order (spiral.cpp:27-72), Morton order inside a
// block (integrator.cpp:196-209). Depends only on (crop size, block size, rank, world): kept across renders.
static int prepare_work_items(B2Scene *s, int W, int H, uint32_t bs, int rank, int world) {
    auto &k = s->items_key;
    if (k.W == W && k.H == H && k.bs == bs && k.rank == rank && k.world == world) return B2_OK;
    uint32_t n_blocks = 0;
    b2_host_spiral(W, H, (int) bs, &n_blocks, nullptr);
    std::vector<int32_t> order(2 * (size_t) n_blocks);
    b2_host_spiral(W, H, (int) bs, &n_blocks, order.data());
    // Morton decode table of one block: index -> (x, y)
    std::vector<uint16_t> mxy(2 * (size_t) bs * bs);
    for (uint32_t i = 0; i < bs * bs; ++i) {
        uint32_t mx = 0, my = 0;
        for (uint32_t bit = 0; bit < 16; ++bit) { mx |= ((i >> (2 * bit)) & 1u) << bit; my |= ((i >> (2 * bit + 1)) & 1u) << bit; }
        mxy[2 * i] = (uint16_t) mx; mxy[2 * i + 1] = (uint16_t) my;
    }
    std::vector<uint32_t> item_pixel, item_seed;
    item_pixel.reserve((size_t) W * H / (size_t) world + bs * bs); item_seed.reserve(item_pixel.capacity());
    for (uint32_t b = 0; b < n_blocks; ++b) {
        if ((int) (b % (uint32_t) world) != rank) continue;
        int ox = order[2 * b] * (int) bs, oy = order[2 * b + 1] * (int) bs;
        int bw = std::min((int) bs, W - ox), bh = std::min((int) bs, H - oy);
        // Morton indices in increasing order, skipping pixels outside (integrator.cpp:200-202)
        for (uint32_t i = 0; i < bs * bs; ++i) {
            int mx = mxy[2 * i], my = mxy[2 * i + 1];
            if (mx >= bw || my >= bh) continue;
            item_pixel.push_back((uint32_t) (ox + mx) | ((uint32_t) (oy + my) << 16));
            item_seed.push_back(b * bs * bs + i);
        }
    }
    CU(s->item_pixel.upload(item_pixel)); CU(s->item_seed.upload(item_seed));
    k.W = W; k.H = H; k.bs = bs; k.rank = rank; k.world = world; k.n_items = (uint32_t) item_pixel.size(); k.n_blocks = n_blocks;
    return B2_OK;
}

int b2_render_device(B2Scene *s, const B2RenderParams *rp, float **d_xyzaw, B2Stats *stats) {
    if (!s || !rp) return fail(B2_ERR_INVALID, "b2_render: null argument");
    const B2Film &film = s->d.film;
    uint32_t bs = rp->block_size;
    if (bs == 0 || (bs & (bs - 1)) || bs > 256) return fail(B2_ERR_INVALID, "block_size must be a power of two in [1,256] and pinned by the caller");
    if (rp->spp == 0) return fail(B2_ERR_INVALID, "sample_count must be > 0");
    uint32_t spp_pass = rp->samples_per_pass ? std::min(rp->samples_per_pass, rp->spp) : rp->spp;
    if (rp->spp % spp_pass) return fail(B2_ERR_INVALID, "sample_count (%u) must be a multiple of samples_per_pass (%u).", rp->spp, spp_pass);   // integrator.cpp:62-64
    if (spp_pass > 65535) return fail(B2_ERR_INVALID, "samples_per_pass must be <= 65535");
    uint32_t n_passes = rp->spp / spp_pass;
    int W = film.crop_size[0], H = film.crop_size[1];
    int world = rp->world_size > 0 ? rp->world_size : 1, rank = rp->rank;
    if (rank < 0 || rank >= world) return fail(B2_ERR_INVALID, "bad rank/world_size");
    s->cancel.store(0);
    cudaStream_t stream = s->stream;

    { int rc = prepare_work_items(s, W, H, bs, rank, world); if (rc != B2_OK) return rc; }
    const uint32_t n_items_pass = s->items_key.n_items, n_blocks = s->items_key.n_blocks;
    uint64_t total64 = (uint64_t) n_items_pass * n_passes;
    bool wide = film.filter_radius > 0.5f + 1500.0f * 5.9604644775390625e-08f;
    CU(s->film.alloc((size_t) W * H * 5));
    CU(s->stats.alloc(B2_STAT_COPIES * B2_STAT_STRIDE));
    if (total64 == 0) {
        // a rank that owns no spiral block (world_size > number of blocks): empty film, nothing to launch
        CU(cudaMemsetAsync(s->film.p, 0, (size_t) W * H * 5 * sizeof(float), stream));
        CU(cudaStreamSynchronize(stream));
        if (stats) { std::memset(stats, 0, sizeof *stats); stats->bvh_nodes = s->n_nodes; stats->bvh_tris = s->n_tris; stats->bvh_depth = s->bvh_depth; }
        if (d_xyzaw) *d_xyzaw = s->film.p;
        return B2_OK;
    }
    // default wavefront width: a whole number of lanes per thread for every stage kernel
    // (sm_count * 3072 = 3 lanes/thread at 4 CTAs/SM, 4 lanes/thread at 3 CTAs/SM) — and small
    // enough (75 MB of queues on 148 SMs) to stay resident in the 126 MB L2 between the stages
    // (instanced scenes: ray cost varies by orders of magnitude, so the launches are made 16x longer to
    //  amortise their tails — measured 680 -> 890 Msamples/s on the BASELINE lattice)
    uint32_t slots = rp->lane_slots ? rp->lane_slots : (uint32_t) s->sm_count * 3072u * (s->d.n_instances ? 16u : 1u);
    if ((uint64_t) slots > total64) slots = (uint32_t) std::max<uint64_t>(total64, 1);
    slots = (slots + 31u) & ~31u;
    if (total64 + slots >= 0xffffffffull) return fail(B2_ERR_INVALID, "too many work items (%llu)", (unsigned long long) total64);

    const size_t cap = (size_t) slots + 128;      // two lane sets, each rounded up to a multiple of 32
    // DirectIntegrator (direct.cpp:78-100): E emitter samples + B BSDF samples at the camera hit
    const uint32_t dE = rp->direct_emitter_samples, dB = rp->direct_bsdf_samples;
    const bool direct = (dE | dB) != 0;
    if (direct && (dE > 64 || dB > 64)) return fail(B2_ERR_INVALID, "direct integrator: at most 64 emitter / BSDF samples");
    const size_t sh_cap = cap * (size_t) std::max<uint32_t>(dE, 1u) + 2u * B2_SH_MAXSUB * 32u;    // (+ the rounding of the sub-queues)
    CU(s->ray_o.alloc(cap)); CU(s->ray_d.alloc(cap)); CU(s->hit.alloc(cap)); CU(s->tp.alloc(cap)); CU(s->res.alloc(cap));
    CU(s->sh_o.alloc(sh_cap)); CU(s->sh_d.alloc(sh_cap)); CU(s->sh_c.alloc(sh_cap));
    if (direct) { CU(s->prim_o.alloc(cap)); CU(s->prim_d.alloc(cap)); CU(s->prim_hit.alloc(cap)); CU(s->prim_inst.alloc(cap)); }
    CU(s->st.alloc(cap)); CU(s->pixel.alloc(cap)); CU(s->list.alloc(2 * cap));
    if (wide) CU(s->pos.alloc(cap));
    // persistent-warp traversal (dynamic ray fetch): always for instanced scenes; on request for the others (the flat
    // leaf scan of tiny scenes has no use for it)
    const bool inst = s->d.n_instances != 0 || ((rp->flags & B2_FLAG_PERSISTENT_TRACE) && s->d.n_flat_leaves == 0);
    if (inst) { CU(s->hit_inst.alloc(cap)); CU(s->cls.alloc(cap)); CU(cudaMemsetAsync(s->cls.p, 0, cap, stream)); }
    // yield / resume traversal (k_trace_resume), on request: BVH4 scenes without instances whose rays need a real tree walk
    // (the flat leaf scan of tiny scenes has no tail to cut). Measured (profiles/README.md): 23 % shorter traversal launches,
    // but the suspended lanes miss their shading turn, so a render needs 31 % more iterations — no net gain.
    const bool resume = !inst && s->d.n_flat_leaves == 0 && (rp->flags & B2_FLAG_RESUME_TRACE);
    if (resume) { CU(s->resume.alloc(2 * cap)); CU(s->tsave.alloc(cap * B2_TSAVE_INT4)); }
    CU(s->cnt.alloc(B2_CNT_TOTAL));
    CU(s->sh_cnt.alloc(2 * 2 * B2_SH_MAXSUB * 32));
    if (wide) CU(s->film_wide.alloc((size_t) W * H * 8));
    // samplers (B2RenderParams.sampler): the stratified sampler needs a square sample count (stratified.cpp:63-72 rounds it
    // up and warns; here it is an error so that spp keeps its meaning)
    const uint32_t sampler_type = rp->sampler & 0xffu;
    uint32_t smp_res = 1, smp_res_y = 1;
    if (sampler_type != B2_SAMPLER_INDEPENDENT) {
        if (sampler_type == B2_SAMPLER_STRATIFIED) {
            while (smp_res * smp_res < rp->spp) ++smp_res;
            smp_res_y = smp_res;
            if (smp_res * smp_res != rp->spp) return fail(B2_ERR_INVALID, "Sample count should be square and power of two (got %u)", rp->spp);
        } else if (sampler_type == B2_SAMPLER_MULTIJITTER) {           // multijitter.cpp:68-81
            smp_res_y = (uint32_t) std::sqrt((float) rp->spp);
            smp_res = (rp->spp + smp_res_y - 1) / smp_res_y;
            if (smp_res * smp_res_y != rp->spp) return fail(B2_ERR_INVALID, "Sample count rounded up to %u (got %u)", smp_res * smp_res_y, rp->spp);
        } else if (sampler_type == B2_SAMPLER_LOW_DISCREPANCY) {       // ldsampler.cpp:68-78: spp = res^2, res a power of two
            uint32_t lg = 0; while ((1u << lg) < rp->spp) ++lg;
            if ((1u << lg) != rp->spp || (lg & 1u) || rp->spp < 4) return fail(B2_ERR_INVALID, "Sample count should be square and power of two (got %u)", rp->spp);
            smp_res = lg;                                                // log2 of the sample count (permute(), random.h:194-207)
        } else if (sampler_type == B2_SAMPLER_ORTHOGONAL) {            // orthogonal.cpp:76-90: spp = p^2 with p prime (strength 2)
            uint32_t r = 2;
            for (;; ++r) {
                bool prime = true;
                for (uint32_t i = 2; i <= r / 2; ++i) if (r % i == 0) { prime = false; break; }
                if (prime && r * r >= rp->spp) break;
            }
            if (r * r != rp->spp) return fail(B2_ERR_INVALID, "Sample count should be the square of a prime number, rounding to %u (got %u)", r * r, rp->spp);
            smp_res = smp_res_y = r;
        } else return fail(B2_ERR_INVALID, "unsupported sampler %u", sampler_type);
        CU(s->smp.alloc(cap));
    }
    if (!s->h_cnt) CU(cudaMallocHost(&s->h_cnt, 2 * B2_CNT_TOTAL * sizeof(uint32_t)));

    // The lane slots are split into two independent lane SETS, each with its own counters, lists and shadow
    // queue, driven on its own stream: while one set is in its (issue-bound) traversal launch the
    // other is in its (latency-bound) shading launch, so the two kinds of kernel share the SMs and the
    // launch tails of one set are filled by the other. Both sets pull work items from one counter.
    // (per-stage timing runs one set so that a stage's time is its own)
    const int n_sets = (slots >= 2 * 65536 && !(rp->flags & (B2_FLAG_SINGLE_SET | B2_FLAG_STAGE_TIMING))) ? 2 : 1;
    const uint32_t set_slots = ((slots / n_sets) + 31u) & ~31u;
    if (n_sets > 1) slots = set_slots * n_sets;     // (allocations above were sized with the rounded-up count)
    const int WORK_WORD = B2_CNT_WORK;
    Queues qs[2];
    for (int k = 0; k < n_sets; ++k) {
        Queues &q = qs[k];
        q = Queues{};
        size_t o = (size_t) k * set_slots;
        q.ray_o = s->ray_o.p + o; q.ray_d = s->ray_d.p + o; q.hit = s->hit.p + o; q.tp = s->tp.p + o; q.res = s->res.p + o; q.st = s->st.p + o;
        q.pixel = s->pixel.p + o; q.list[0] = s->list.p + 2 * o; q.list[1] = s->list.p + 2 * o + set_slots;
        q.pos = wide ? s->pos.p + o : nullptr;
        q.hit_inst = inst ? s->hit_inst.p + o : nullptr; q.cls = inst ? s->cls.p + o : nullptr;
        if (resume) {
            q.resume[0] = s->resume.p + 2 * o; q.resume[1] = s->resume.p + 2 * o + set_slots; q.tsave = s->tsave.p + o * B2_TSAVE_INT4;
            const uint32_t thr = (rp->flags >> 8) & 0x1fu;                // B2_FLAG_YIELD_THRESHOLD(n): tuning override
            q.yield_threshold = thr ? thr : 12u;
        }
        // the shadow queue of a lane set: 8 sub-queues with their own counters (one 128-byte line each) once the wavefront is
        // wide enough for the single counter to be a hot spot (b2_kernels.cuh: sh_counter()); the direct integrator's entries
        // are addressed by lane
        const uint32_t sh_sub = (!direct && set_slots >= 8192u) ? 8u : 0u;
        const size_t so = direct ? o * (size_t) dE : o + (size_t) k * B2_SH_MAXSUB * 32u;
        q.sh_o = s->sh_o.p + so; q.sh_d = s->sh_d.p + so; q.sh_c = s->sh_c.p + so;
        q.sh_sub = sh_sub; q.sh_cnt = s->sh_cnt.p + (size_t) k * 2 * B2_SH_MAXSUB * 32;
        q.sh_cap = sh_sub ? ((((set_slots + 31u) / 32u) + sh_sub - 1u) / sh_sub) * 32u : 0u;
        if (direct) {
            q.direct_e = dE; q.direct_b = dB;
            const float sum = (float) (dE + dB);                         // direct.cpp:96-100
            q.d_weight_bsdf = 1.f / (float) dB; q.d_weight_lum = 1.f / (float) dE; q.d_frac_bsdf = (float) dB / sum; q.d_frac_lum = (float) dE / sum;
            q.prim_o = s->prim_o.p + o; q.prim_d = s->prim_d.p + o; q.prim_hit = s->prim_hit.p + o; q.prim_inst = s->prim_inst.p + o;
        }
        q.cnt = s->cnt.p + B2_CNT_WORDS * k; q.work = s->cnt.p + WORK_WORD; q.stats = s->stats.p;
        if (sampler_type != B2_SAMPLER_INDEPENDENT) {
            q.smp = s->smp.p + o; q.smp_count = rp->spp; q.smp_res = smp_res; q.smp_jitter = (rp->sampler & B2_SAMPLER_NO_JITTER) ? 0u : 1u;
            q.smp_inv_count = 1.f / (float) rp->spp; q.smp_inv_res = 1.f / (float) smp_res;
            q.smp_type = sampler_type; q.smp_res_y = smp_res_y; q.smp_inv_res_y = 1.f / (float) smp_res_y;
        }
        q.film = wide ? s->film_wide.p : s->film.p; q.item_pixel = s->item_pixel.p; q.item_seed = s->item_seed.p;
        q.n_items_pass = n_items_pass; q.n_passes = n_passes; q.spp_pass = spp_pass; q.block_size = bs; q.n_blocks_total = n_blocks;
        q.total_items = (uint32_t) total64; q.n_slots = set_slots; q.base_seed = rp->seed; q.max_depth = rp->max_depth; q.rr_depth = rp->rr_depth;
        if (direct) { q.max_depth = 2; q.rr_depth = 0x7fffffff; }          // (one bounce, no Russian roulette)
    }

    EventPool pool;
    cudaEvent_t ev0 = pool.make(true), ev1 = pool.make(true), evb[2] = { pool.make(false), pool.make(false) }, evs = pool.make(false), evx = pool.make(false);
    if (!ev0 || !ev1 || !evb[0] || !evb[1] || !evs || !evx) return fail(B2_ERR_CUDA, "cudaEventCreate failed");
    if (n_sets > 1 && !s->stream2) CU(cudaStreamCreateWithFlags(&s->stream2, cudaStreamNonBlocking));
    cudaStream_t streams[2] = { stream, n_sets > 1 ? s->stream2 : stream };
    CU(cudaEventRecord(ev0, stream));
    CU(cudaMemsetAsync(s->film.p, 0, (size_t) W * H * 5 * sizeof(float), stream));
    if (wide) CU(cudaMemsetAsync(s->film_wide.p, 0, (size_t) W * H * 8 * sizeof(float), stream));
    CU(cudaMemsetAsync(s->cnt.p, 0, B2_CNT_TOTAL * sizeof(uint32_t), stream));
    CU(cudaMemsetAsync(s->sh_cnt.p, 0, 2 * 2 * B2_SH_MAXSUB * 32 * sizeof(uint32_t), stream));
    CU(cudaMemsetAsync(s->stats.p, 0, B2_STAT_COPIES * B2_STAT_STRIDE * sizeof(unsigned long long), stream));
    const int threads = 256;
    // CTA sizes of the two stage kernels (the kernels stride over the lane slots by warps, so any multiple of 32 is correct;
    // grids keep the same number of warps). Every CTA stages its tables into shared memory first: scenes that stage only a few
    // KB (the Cornell box) run 128-thread CTAs — finer-grained slots for the two lane sets' launches to interleave: +2.7 %,
    // measured — the others 256 (matball: 128-thread traversal CTAs re-stage the 48 KB BVH top twice as often: -10 %).
    // B2_FLAG_TRACE_BLOCK / B2_FLAG_SHADE_BLOCK override (tuning).
    const uint32_t tb_sel = (rp->flags >> 16) & 3u, sb_sel = (rp->flags >> 18) & 3u;
    const int threads_trace = tb_sel ? 512 >> tb_sel : (!inst && s->smem_bytes <= 8192 ? 128 : 256);
    const int threads_shade = sb_sel ? 512 >> sb_sel : (s->shade_smem_bytes > 0 && s->shade_smem_bytes <= 8192 ? 128 : 256);
    // non-independent samplers run in the extended (F_TEX) kernel variants
    const uint32_t run_features = s->features | ((sampler_type != B2_SAMPLER_INDEPENDENT || direct) ? (uint32_t) F_TEX : 0u);
    const ShadeFn shade_fn = shade_variant(run_features);
    const TraceFn trace_fn = trace_variant(inst, s->has_shapes, (rp->flags & B2_FLAG_COUNT_BVH) != 0, resume);
    // (the limit is a per-function, process-wide attribute: set it for THIS scene right before its launches)
    CU(cudaFuncSetAttribute(shade_fn, cudaFuncAttributeMaxDynamicSharedMemorySize, s->shade_smem_bytes));
    // flat scenes: + the per-warp candidate lists of the cooperative any-hit test (flat_any_hit_coop)
    const int trace_smem = s->d.n_flat_leaves && !inst ? ((s->smem_bytes + 15) & ~15) + (threads_trace / 32) * (int) B2_COOP_BYTES_PER_WARP : s->smem_bytes;
    CU(cudaFuncSetAttribute(trace_fn, cudaFuncAttributeMaxDynamicSharedMemorySize, trace_smem));
    uint32_t need_blocks = (set_slots + threads - 1) / threads;
    int grid = (int) std::max<uint32_t>(1, std::min<uint32_t>((uint32_t) s->sm_count * 8, need_blocks));
    // measured: 8 CTAs/SM queued beats exactly-resident grids
    int grid_int = (int) std::min<uint32_t>((uint32_t) s->sm_count * 8, need_blocks * 2) * (threads / threads_trace);
    int grid_sha = grid * (threads / threads_shade);
    if (inst) grid_int = s->sm_count * B2_INST_MINBLOCKS * (threads / threads_trace);           // k_trace_inst: persistent warps
    uint64_t iterations = 0, launches = 0;
    const int batch = 8;
    const bool timing = (rp->flags & B2_FLAG_STAGE_TIMING) != 0;
    std::vector<cudaEvent_t> tev;     // 3 events per iteration when stage timing is on
    auto tick = [&](cudaStream_t st) { if (timing) { cudaEvent_t e = pool.make(true); if (e) { cudaEventRecord(e, st); tev.push_back(e); } } };
    cudaEvent_t evi0 = nullptr, evi1 = nullptr;
    if (timing) { evi0 = pool.make(true); evi1 = pool.make(true); if (evi0) cudaEventRecord(evi0, stream); }
    for (int k = 0; k < n_sets; ++k) {
        if (sampler_type != B2_SAMPLER_INDEPENDENT || direct) k_init<true><<<grid, threads, 0, stream>>>(s->d, qs[k]);
        else k_init<false><<<grid, threads, 0, stream>>>(s->d, qs[k]);
        ++launches;
    }
    if (timing && evi1) cudaEventRecord(evi1, stream);
    if (n_sets > 1) { CU(cudaEventRecord(evs, stream)); CU(cudaStreamWaitEvent(streams[1], evs, 0)); }
    // progressive snapshots: film -> pinned host buffer on a side stream (copy engine), handed to the callback one batch later
    const size_t film_floats = (size_t) W * H * 5;
    cudaEvent_t ev_film = nullptr, ev_snap = nullptr;
    bool snap_in_flight = false;
    auto snap_last = std::chrono::steady_clock::now();
    if (s->progress_fn) {
        if (!s->stream_snap) CU(cudaStreamCreateWithFlags(&s->stream_snap, cudaStreamNonBlocking));
        if (s->h_snap_floats < film_floats) {
            if (s->h_snap) cudaFreeHost(s->h_snap);
            s->h_snap = nullptr; s->h_snap_floats = 0;
            CU(cudaMallocHost(&s->h_snap, film_floats * sizeof(float))); s->h_snap_floats = film_floats;
        }
        if (!s->h_snap_samples) CU(cudaMallocHost(&s->h_snap_samples, B2_STAT_COPIES * B2_STAT_STRIDE * sizeof(unsigned long long)));
        if (wide) CU(s->snap_film.alloc(film_floats));
        ev_film = pool.make(false); ev_snap = pool.make(false);
        if (!ev_film || !ev_snap) return fail(B2_ERR_CUDA, "cudaEventCreate failed");
    }
    int pending[2] = { 0, 0 };
    bool done = false;
    int status = B2_OK;
    int parity = 0;
    for (int b = 0; !done; ++b) {
        int slot = b & 1;
        for (int it = 0; it < batch; ++it) {
            // iteration: trace (closest hit + the previous iteration's emitter samples + compaction) -> shade
            for (int k = 0; k < n_sets; ++k) {
                cudaStream_t st = streams[k];
                tick(st);
                trace_fn<<<grid_int, threads_trace, trace_smem, st>>>(s->d, qs[k], parity);
                if (inst) { k_classify<<<grid, threads, 0, st>>>(qs[k], parity); ++launches; }
                tick(st);
                shade_fn<<<grid_sha, threads_shade, s->shade_smem_bytes, st>>>(s->d, qs[k], parity);
                tick(st);
                launches += 2;
            }
            parity ^= 1;
            ++iterations;
        }
        if (n_sets > 1) { CU(cudaEventRecord(evx, streams[1])); CU(cudaStreamWaitEvent(stream, evx, 0)); }
        CU(cudaMemcpyAsync(s->h_cnt + B2_CNT_TOTAL * slot, s->cnt.p, B2_CNT_TOTAL * sizeof(uint32_t), cudaMemcpyDeviceToHost, stream));
        CU(cudaEventRecord(evb[slot], stream));
        if (n_sets > 1) CU(cudaStreamWaitEvent(streams[1], evb[slot], 0));     // keep the streams within one batch of each other
        pending[slot] = 1;
        // wait for the PREVIOUS batch's readback so the GPU always has one batch queued
        int prev = slot ^ 1;
        if (pending[prev]) {
            CU(cudaEventSynchronize(evb[prev]));
            pending[prev] = 0;
            const uint32_t *c = s->h_cnt + B2_CNT_TOTAL * prev;
            // finished when every lane retired (parked lanes count as alive) and no emitter sample is unresolved
            bool all = true;
            for (int k = 0; k < n_sets; ++k)
                all = all && c[B2_CNT_WORDS * k + CNT_ALIVE] == 0 && c[B2_CNT_WORDS * k + CNT_SHADOW] == 0 && c[B2_CNT_WORDS * k + CNT_SHADOW + 1] == 0;
            if (all) done = true;
        }
        if (s->progress_fn) {
            if (snap_in_flight && cudaEventQuery(ev_snap) == cudaSuccess) {
                snap_in_flight = false;
                unsigned long long n_snap = 0;                              // (the counters are kept in B2_STAT_COPIES copies)
                for (int c = 0; c < B2_STAT_COPIES; ++c) n_snap += s->h_snap_samples[c * B2_STAT_STRIDE + ST_SAMPLES];
                if (s->progress_fn(s->progress_user, s->h_snap, n_snap, (uint64_t) W * H * rp->spp / (uint64_t) world)) s->cancel.store(1);
            }
            auto now = std::chrono::steady_clock::now();
            if (!snap_in_flight && !done && std::chrono::duration<double, std::milli>(now - snap_last).count() >= (double) s->progress_ms) {
                snap_last = now;
                CU(cudaEventRecord(ev_film, stream));                      // (both lane sets were joined into `stream` above)
                CU(cudaStreamWaitEvent(s->stream_snap, ev_film, 0));
                const float *src = s->film.p;
                if (wide) {
                    k_film_compact<<<(unsigned) (((size_t) W * H + 255) / 256), 256, 0, s->stream_snap>>>(s->film_wide.p, s->snap_film.p, (size_t) W * H);
                    src = s->snap_film.p;
                }
                CU(cudaMemcpyAsync(s->h_snap, src, film_floats * sizeof(float), cudaMemcpyDeviceToHost, s->stream_snap));
                CU(cudaMemcpyAsync(s->h_snap_samples, s->stats.p, B2_STAT_COPIES * B2_STAT_STRIDE * sizeof(unsigned long long), cudaMemcpyDeviceToHost, s->stream_snap));
                CU(cudaEventRecord(ev_snap, s->stream_snap));
                snap_in_flight = true;
            }
        }
        if (s->cancel.load()) { status = B2_ERR_CANCELLED; done = true; }
        if (cudaPeekAtLastError() != cudaSuccess) break;
    }
    if (n_sets > 1) { CU(cudaEventRecord(evx, streams[1])); CU(cudaStreamWaitEvent(stream, evx, 0)); }
    if (snap_in_flight) CU(cudaStreamSynchronize(s->stream_snap));       // (the snapshot buffers must not outlive this call in flight)
    CU(cudaGetLastError());
    if (wide) { size_t np = (size_t) W * H; k_film_compact<<<(unsigned) ((np + 255) / 256), 256, 0, stream>>>(s->film_wide.p, s->film.p, np); ++launches; }
    CU(cudaEventRecord(ev1, stream));
    CU(cudaEventSynchronize(ev1));
    float ms = 0.f; CU(cudaEventElapsedTime(&ms, ev0, ev1));
    if (stats) {
        unsigned long long hc[B2_STAT_COPIES * B2_STAT_STRIDE], h[8] = { 0, 0, 0, 0, 0, 0, 0, 0 };
        CU(cudaMemcpyAsync(hc, s->stats.p, sizeof hc, cudaMemcpyDeviceToHost, stream));
        CU(cudaStreamSynchronize(stream));
        for (int c = 0; c < B2_STAT_COPIES; ++c) for (int k = 0; k < 8; ++k) h[k] += hc[c * B2_STAT_STRIDE + k];
        std::memset(stats, 0, sizeof *stats);
        stats->n_samples = h[ST_SAMPLES]; stats->n_segments = h[ST_SEGMENTS]; stats->n_shadow_rays = h[ST_SHADOW];
        stats->n_invalid = h[ST_INVALID]; stats->n_iterations = iterations; stats->n_launches = launches;
        stats->n_nodes_visited = h[ST_NODES]; stats->n_tris_tested = h[ST_TRIS];
        stats->ms_total = ms; stats->bvh_nodes = s->n_nodes; stats->bvh_tris = s->n_tris; stats->bvh_depth = s->bvh_depth;
        stats->ms_build = s->ms_build; stats->ms_upload = s->ms_upload;
        for (size_t i = 0; i + 2 < tev.size(); i += 3) {
            float t[2] = { 0.f, 0.f };
            for (int k = 0; k < 2; ++k) cudaEventElapsedTime(&t[k], tev[i + k], tev[i + k + 1]);
            stats->ms_intersect += t[0]; stats->ms_shade += t[1];     // ms_intersect = k_trace (closest + any hit + compaction)
        }
        if (timing && evi0 && evi1) cudaEventElapsedTime(&stats->ms_regen, evi0, evi1);      // k_init
    }
    if (d_xyzaw) *d_xyzaw = s->film.p;
    // (a cancelled render leaves the samples accumulated so far on the film: develop it like the reference does on
    //  SIGHUP, mitsuba.cpp:112-121 / integrator.cpp:58-65 — W carries the per-pixel sample count)
    if (status == B2_ERR_CANCELLED) return fail(status, "render cancelled");
    return status;
}

int b2_render(B2Scene *s, const B2RenderParams *rp, float *xyzaw, B2Stats *stats) {
    return b2_render_reduce(s, rp, nullptr, nullptr, xyzaw, stats);
}

// Multi-GPU: every rank renders its blocks (b % world == rank); `reduce` is called once with this rank's DEVICE film
// (n floats, on `stream`) and sums it into rank 0's — e.g. ncclReduce(film, film, n, ncclFloat, ncclSum, 0, comm, stream)
// — and only then the film is copied to the host buffer (rank 0; other ranks may pass xyzaw = NULL).
int b2_render_reduce(B2Scene *s, const B2RenderParams *rp, b2_reduce_fn reduce, void *user, float *xyzaw, B2Stats *stats) {
    if (!s || !rp) return fail(B2_ERR_INVALID, "b2_render: null argument");
    if (!xyzaw && !reduce) return fail(B2_ERR_INVALID, "b2_render: null output");
    float *d = nullptr;
    int rc = b2_render_device(s, rp, &d, stats);
    if (rc != B2_OK && rc != B2_ERR_CANCELLED) return rc;
    const std::string msg = g_error;
    size_t n = (size_t) s->d.film.crop_size[0] * s->d.film.crop_size[1] * 5;
    if (reduce && rc == B2_OK) {
        int rr = reduce(user, d, n, (void *) s->stream);
        if (rr != 0) return fail(B2_ERR_CUDA, "b2_render_reduce: the reduce callback failed (%d)", rr);
    }
    if (xyzaw) {
        CU(cudaMemcpyAsync(xyzaw, d, n * sizeof(float), cudaMemcpyDeviceToHost, s->stream));
        CU(cudaStreamSynchronize(s->stream));
    }
    if (rc != B2_OK) g_error = msg;
    return rc;
}

// host <-> device copies of the stage entry points are ordered on the scene's stream (b2_set_stream may install a
// non-blocking stream, on which the legacy-stream cudaMemcpy would not be ordered with the kernel)
#define H2D(dst, src, bytes) CU(cudaMemcpyAsync((dst), (src), (bytes), cudaMemcpyHostToDevice, s->stream))
#define D2H(dst, src, bytes) CU(cudaMemcpyAsync((dst), (src), (bytes), cudaMemcpyDeviceToHost, s->stream))
#define SYNC() CU(cudaStreamSynchronize(s->stream))

int b2_develop(B2Scene *s, const float *xyzaw, float *rgb) {
    if (!s || !xyzaw || !rgb) return fail(B2_ERR_INVALID, "b2_develop: null argument");
    const size_t n = (size_t) s->d.film.crop_size[0] * (size_t) s->d.film.crop_size[1];
    DBuf<float> in, out;
    CU(in.alloc(n * 5)); CU(out.alloc(n * 3));
    H2D(in.p, xyzaw, n * 5 * sizeof(float));
    k_develop<<<(unsigned) ((n + 255) / 256), 256, 0, s->stream>>>(in.p, out.p, (int) n);
    CU(cudaGetLastError());
    D2H(rgb, out.p, n * 3 * sizeof(float));
    SYNC();
    return B2_OK;
}

// HDRFilm::bitmap() on the device film of the last b2_render_device (no host round trip of the XYZAW film)
int b2_develop_device(B2Scene *s, float *rgb) {
    if (!s || !rgb) return fail(B2_ERR_INVALID, "b2_develop_device: null argument");
    const size_t n = (size_t) s->d.film.crop_size[0] * (size_t) s->d.film.crop_size[1];
    if (!s->film.p || s->film.n < n * 5) return fail(B2_ERR_INVALID, "b2_develop_device: no film (call b2_render_device first)");
    DBuf<float> out;
    CU(out.alloc(n * 3));
    k_develop<<<(unsigned) ((n + 255) / 256), 256, 0, s->stream>>>(s->film.p, out.p, (int) n);
    CU(cudaGetLastError());
    D2H(rgb, out.p, n * 3 * sizeof(float));
    SYNC();
    return B2_OK;
}

// --------------------------------------------------------------------------------------------
// stage-level entry points
// --------------------------------------------------------------------------------------------
int b2_stage_raygen(B2Scene *s, uint32_t n, const uint32_t *px, const uint32_t *py, const uint64_t *seed_offset,
                    uint64_t base_seed, float *ray_o, float *ray_d, uint64_t *rng_state, float *pos) {
    if (!s || (n && (!px || !py || !seed_offset || !ray_o || !ray_d || !rng_state || !pos))) return fail(B2_ERR_INVALID, "b2_stage_raygen: null argument");
    if (n == 0) return B2_OK;
    const size_t N = n;
    DBuf<uint32_t> dx, dy; DBuf<uint64_t> ds, dr; DBuf<float4> o, d; DBuf<float2> p;
    CU(dx.alloc(N)); CU(dy.alloc(N)); CU(ds.alloc(N)); CU(dr.alloc(N)); CU(o.alloc(N)); CU(d.alloc(N)); CU(p.alloc(N));
    H2D(dx.p, px, N * 4); H2D(dy.p, py, N * 4); H2D(ds.p, seed_offset, N * 8);
    k_stage_raygen<<<(unsigned) ((N + 255) / 256), 256, 0, s->stream>>>(s->d, n, dx.p, dy.p, ds.p, base_seed, o.p, d.p, dr.p, p.p);
    CU(cudaGetLastError());
    D2H(ray_o, o.p, N * 16); D2H(ray_d, d.p, N * 16); D2H(rng_state, dr.p, N * 8); D2H(pos, p.p, N * 8);
    SYNC();
    return B2_OK;
}

static int stage_grid(const B2Scene *s, uint32_t n) {
    uint32_t need = (n + 255) / 256;
    return (int) std::max<uint32_t>(1, std::min<uint32_t>(need, (uint32_t) s->sm_count * 8));
}

int b2_stage_intersect(B2Scene *s, uint32_t n, const float *ray_o, const float *ray_d, float *hit, int32_t *inst) {
    if (!s || (n && (!ray_o || !ray_d || !hit))) return fail(B2_ERR_INVALID, "b2_stage_intersect: null argument");
    if (n == 0) return B2_OK;
    const size_t N = n;
    DBuf<float4> o, d, h; DBuf<int32_t> di;
    CU(o.alloc(N)); CU(d.alloc(N)); CU(h.alloc(N)); CU(di.alloc(N));
    H2D(o.p, ray_o, N * 16); H2D(d.p, ray_d, N * 16);
    auto fn = s->d.n_instances ? k_stage_intersect<true> : k_stage_intersect<false>;
    CU(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, s->smem_bytes));
    fn<<<stage_grid(s, n), 256, s->smem_bytes, s->stream>>>(s->d, n, o.p, d.p, h.p, di.p);
    CU(cudaGetLastError());
    D2H(hit, h.p, N * 16);
    if (inst) D2H(inst, di.p, N * 4);
    SYNC();
    return B2_OK;
}

int b2_stage_occluded(B2Scene *s, uint32_t n, const float *ray_o, const float *ray_d, uint8_t *occluded) {
    if (!s || (n && (!ray_o || !ray_d || !occluded))) return fail(B2_ERR_INVALID, "b2_stage_occluded: null argument");
    if (n == 0) return B2_OK;
    const size_t N = n;
    DBuf<float4> o, d; DBuf<uint8_t> r;
    CU(o.alloc(N)); CU(d.alloc(N)); CU(r.alloc(N));
    H2D(o.p, ray_o, N * 16); H2D(d.p, ray_d, N * 16);
    auto fn = s->d.n_instances ? k_stage_occluded<true> : k_stage_occluded<false>;
    CU(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, s->smem_bytes));
    fn<<<stage_grid(s, n), 256, s->smem_bytes, s->stream>>>(s->d, n, o.p, d.p, r.p);
    CU(cudaGetLastError());
    D2H(occluded, r.p, N);
    SYNC();
    return B2_OK;
}

int b2_stage_si_fill(B2Scene *s, uint32_t n, const float *ray_o, const float *ray_d, const float *hit, const int32_t *inst, float *si) {
    if (!s || (n && (!ray_o || !ray_d || !hit || !si))) return fail(B2_ERR_INVALID, "b2_stage_si_fill: null argument");
    if (n == 0) return B2_OK;
    const size_t N = n;
    DBuf<float4> o, d, h; DBuf<float> out; DBuf<int32_t> di;
    CU(o.alloc(N)); CU(d.alloc(N)); CU(h.alloc(N)); CU(out.alloc(N * 16));
    H2D(o.p, ray_o, N * 16); H2D(d.p, ray_d, N * 16); H2D(h.p, hit, N * 16);
    if (inst) { CU(di.alloc(N)); H2D(di.p, inst, N * 4); }
    k_stage_si_fill<<<(unsigned) ((N + 255) / 256), 256, 0, s->stream>>>(s->d, n, o.p, d.p, h.p, inst ? di.p : nullptr, out.p);
    CU(cudaGetLastError());
    D2H(si, out.p, N * 64);
    SYNC();
    return B2_OK;
}

// one k_shade launch over caller-provided lanes, all of them CONTINUE lanes (the stage as the wavefront runs it)
int b2_stage_shade(B2Scene *s, const B2RenderParams *rp, uint32_t n, float *ray_o, float *ray_d, const float *hit, const int32_t *inst,
                   float *throughput_eta, float *result_pdf, uint64_t *rng_state, uint32_t *depth_flags,
                   uint32_t *n_shadow, float *sh_o, float *sh_d, float *sh_c) {
    if (!s || !rp || (n && (!ray_o || !ray_d || !hit || !throughput_eta || !result_pdf || !rng_state || !depth_flags || !n_shadow || !sh_o || !sh_d || !sh_c)))
        return fail(B2_ERR_INVALID, "b2_stage_shade: null argument");
    if ((rp->sampler & 0xffu) != B2_SAMPLER_INDEPENDENT) return fail(B2_ERR_INVALID, "b2_stage_shade: independent sampler only");
    if (s->d.n_instances && !inst) return fail(B2_ERR_INVALID, "b2_stage_shade: instanced scene: inst[] is required");
    *n_shadow = 0;
    if (n == 0) return B2_OK;
    const size_t N = n, cap = (N + 31) & ~(size_t) 31;
    DBuf<float4> o, d, h, tp, res, so, sd, sc4; DBuf<uint4> st; DBuf<uint32_t> list, cnt, df, pix; DBuf<int32_t> di; DBuf<uint64_t> rng;
    DBuf<unsigned long long> stats; DBuf<float2> pos;
    CU(o.alloc(cap)); CU(d.alloc(cap)); CU(h.alloc(cap)); CU(tp.alloc(cap)); CU(res.alloc(cap)); CU(so.alloc(cap)); CU(sd.alloc(cap)); CU(sc4.alloc(cap));
    CU(st.alloc(cap)); CU(list.alloc(2 * cap)); CU(cnt.alloc(B2_CNT_TOTAL)); CU(df.alloc(cap)); CU(pix.alloc(cap)); CU(rng.alloc(cap)); CU(stats.alloc(B2_STAT_COPIES * B2_STAT_STRIDE)); CU(pos.alloc(cap));
    H2D(o.p, ray_o, N * 16); H2D(d.p, ray_d, N * 16); H2D(h.p, hit, N * 16); H2D(tp.p, throughput_eta, N * 16); H2D(res.p, result_pdf, N * 16);
    H2D(rng.p, rng_state, N * 8); H2D(df.p, depth_flags, N * 4);
    if (s->d.n_instances) { CU(di.alloc(cap)); H2D(di.p, inst, N * 4); }
    CU(cudaMemsetAsync(cnt.p, 0, B2_CNT_TOTAL * 4, s->stream)); CU(cudaMemsetAsync(stats.p, 0, B2_STAT_COPIES * B2_STAT_STRIDE * 8, s->stream));
    k_stage_pack<<<(unsigned) ((N + 255) / 256), 256, 0, s->stream>>>(s->d, n, h.p, rng.p, df.p, st.p, list.p);
    const unsigned long long n_list = (unsigned long long) n;        // n CONTINUE lanes, no TERMINATE lanes
    H2D(cnt.p + CNT_LIST, &n_list, 8);
    Queues q{};
    q.ray_o = o.p; q.ray_d = d.p; q.hit = h.p; q.hit_inst = s->d.n_instances ? di.p : nullptr; q.tp = tp.p; q.res = res.p; q.st = st.p; q.pixel = pix.p;
    q.pos = (s->features & F_WIDE) ? pos.p : nullptr;
    q.sh_o = so.p; q.sh_d = sd.p; q.sh_c = sc4.p; q.list[0] = list.p; q.list[1] = list.p + cap; q.cnt = cnt.p; q.work = cnt.p + B2_CNT_WORK; q.stats = stats.p;
    q.film = nullptr; q.n_slots = (uint32_t) cap; q.max_depth = rp->max_depth; q.rr_depth = rp->rr_depth; q.base_seed = rp->seed;
    q.n_items_pass = 1; q.n_passes = 1; q.spp_pass = 1; q.block_size = rp->block_size ? rp->block_size : 32; q.n_blocks_total = 1; q.total_items = 0;
    const ShadeFn fn = shade_variant(s->features);
    CU(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, s->shade_smem_bytes));
    fn<<<stage_grid(s, n), 256, s->shade_smem_bytes, s->stream>>>(s->d, q, 0);
    CU(cudaGetLastError());
    k_stage_unpack<<<(unsigned) ((N + 255) / 256), 256, 0, s->stream>>>(n, st.p, rng.p, df.p);
    uint32_t hc[8];
    D2H(hc, cnt.p, sizeof hc);
    D2H(ray_o, o.p, N * 16); D2H(ray_d, d.p, N * 16); D2H(throughput_eta, tp.p, N * 16); D2H(result_pdf, res.p, N * 16);
    D2H(rng_state, rng.p, N * 8); D2H(depth_flags, df.p, N * 4);
    SYNC();
    const uint32_t ns = hc[CNT_SHADOW];
    if (ns > n) return fail(B2_ERR_CUDA, "b2_stage_shade: shadow queue overflow");
    *n_shadow = ns;
    if (ns) { D2H(sh_o, so.p, (size_t) ns * 16); D2H(sh_d, sd.p, (size_t) ns * 16); D2H(sh_c, sc4.p, (size_t) ns * 16); SYNC(); }
    return B2_OK;
}

// K5 alone: the warp-ballot compaction of the live lanes into the CONTINUE / TERMINATE lists (what k_trace ends with)
int b2_stage_compact(B2Scene *s, int32_t max_depth, uint32_t n, const float *ray_maxt, const float *hit_t, const uint32_t *depth,
                     uint32_t *n_continue, uint32_t *continue_lanes, uint32_t *n_terminate, uint32_t *terminate_lanes) {
    if (!s || !n_continue || !n_terminate || (n && (!ray_maxt || !hit_t || !depth || !continue_lanes || !terminate_lanes)))
        return fail(B2_ERR_INVALID, "b2_stage_compact: null argument");
    *n_continue = *n_terminate = 0;
    if (n == 0) return B2_OK;
    const size_t N = n;
    DBuf<float> dm, dt; DBuf<uint32_t> dd, list, cnt;
    CU(dm.alloc(N)); CU(dt.alloc(N)); CU(dd.alloc(N)); CU(list.alloc(N)); CU(cnt.alloc(B2_CNT_WORDS));
    H2D(dm.p, ray_maxt, N * 4); H2D(dt.p, hit_t, N * 4); H2D(dd.p, depth, N * 4);
    CU(cudaMemsetAsync(cnt.p, 0, B2_CNT_WORDS * 4, s->stream));
    Queues q{};
    q.list[0] = list.p; q.list[1] = list.p; q.cnt = cnt.p; q.n_slots = n; q.max_depth = max_depth;
    k_stage_compact<<<stage_grid(s, n), 256, 0, s->stream>>>(q, n, dm.p, dt.p, dd.p);
    CU(cudaGetLastError());
    unsigned long long counts = 0;
    D2H(&counts, cnt.p + CNT_LIST, 8);
    SYNC();
    const uint32_t nc = (uint32_t) counts, nx = (uint32_t) (counts >> 32);
    if ((size_t) nc + nx > N) return fail(B2_ERR_CUDA, "b2_stage_compact: list overflow");
    *n_continue = nc; *n_terminate = nx;
    if (nc) D2H(continue_lanes, list.p, (size_t) nc * 4);
    if (nx) D2H(terminate_lanes, list.p + (N - nx), (size_t) nx * 4);        // (appended from the back: descending positions)
    SYNC();
    return B2_OK;
}

int b2_stage_film_put(B2Scene *s, uint32_t block_size, uint32_t n, const uint32_t *px, const uint32_t *py, const float *pos, const float *rgb,
                      const uint8_t *alpha, float *xyzaw, uint32_t *n_invalid) {
    if (!s || !xyzaw || (n && (!px || !py || !pos || !rgb || !alpha))) return fail(B2_ERR_INVALID, "b2_stage_film_put: null argument");
    if (block_size == 0 || (block_size & (block_size - 1))) return fail(B2_ERR_INVALID, "block_size must be a power of two");
    const size_t N = n, np = (size_t) s->d.film.crop_size[0] * (size_t) s->d.film.crop_size[1];
    const bool wide = (s->features & F_WIDE) != 0;
    DBuf<uint32_t> dx, dy, bad; DBuf<float2> dp; DBuf<float> drgb, film, film8; DBuf<uint8_t> da;
    CU(dx.alloc(N)); CU(dy.alloc(N)); CU(dp.alloc(N)); CU(drgb.alloc(N * 3)); CU(da.alloc(N)); CU(film.alloc(np * 5)); CU(bad.alloc(1));
    if (wide) CU(film8.alloc(np * 8));
    for (uint32_t i = 0; i < n; ++i)
        if (px[i] >= (uint32_t) s->d.film.crop_size[0] || py[i] >= (uint32_t) s->d.film.crop_size[1]) return fail(B2_ERR_INVALID, "b2_stage_film_put: pixel %u out of range", i);
    if (n) { H2D(dx.p, px, N * 4); H2D(dy.p, py, N * 4); H2D(dp.p, pos, N * 8); H2D(drgb.p, rgb, N * 12); H2D(da.p, alpha, N); }
    CU(cudaMemsetAsync(film.p, 0, np * 20, s->stream)); CU(cudaMemsetAsync(bad.p, 0, 4, s->stream));
    if (wide) CU(cudaMemsetAsync(film8.p, 0, np * 32, s->stream));
    if (n) {
        if (wide) k_stage_film_put<true><<<(unsigned) ((N + 255) / 256), 256, 0, s->stream>>>(s->d, n, block_size, dx.p, dy.p, dp.p, drgb.p, da.p, film8.p, bad.p);
        else k_stage_film_put<false><<<(unsigned) ((N + 255) / 256), 256, 0, s->stream>>>(s->d, n, block_size, dx.p, dy.p, dp.p, drgb.p, da.p, film.p, bad.p);
    }
    if (wide) k_film_compact<<<(unsigned) ((np + 255) / 256), 256, 0, s->stream>>>(film8.p, film.p, np);
    CU(cudaGetLastError());
    D2H(xyzaw, film.p, np * 20);
    uint32_t hb = 0; D2H(&hb, bad.p, 4);
    SYNC();
    if (n_invalid) *n_invalid = hb;
    return B2_OK;
}

}  // extern "C"
