/*
 * b2o.c — TEST INFRASTRUCTURE: CPU oracle for the b2pt hot path.
 *
 * A plain-C restatement of the reference's scalar_rgb path tracer (mitsuba2 @ 62863cb3),
 * written in the reference's own (depth-first, one path at a time) shape — NOT in the
 * wavefront shape of the CUDA product — so that a bit-for-bit match of the two is evidence
 * that the wavefront stages compute what the reference computes. Only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load this
 * library; the product (mitsuba2_b200/) never does.
 *
 * Pinned against the reference by tests/golden/ (images rendered with the unmodified
 * reference CLI built by oracle/build_ref.sh, identical sampler seeds) and by the
 * known-answer values transcribed from the reference's own tests (see tests/test_oracle_*.py).
 *
 * Every function cites the reference file:line it follows (paths relative to the reference).
 * Float operation order is fixed (see b2o_math.h); compile with -ffp-contract=off.
 */
#include "../include/b2pt.h"
#include "b2o_math.h"

#include <stdio.h>
#include <stdlib.h>
#include <pthread.h>

#define B2O_API __attribute__((visibility("default")))

/* ============================================================================================
 * PCG32  —  ext/enoki/include/enoki/random.h:62-119, src/librender/sampler.cpp:83-96
 * ============================================================================================ */
#define PCG32_DEFAULT_STREAM 0xda3e39cb94b95bdbULL
#define PCG32_MULT           0x5851f42d4c957f2dULL

typedef struct { uint64_t state, inc; } pcg32;

static inline uint32_t pcg32_next_uint32(pcg32 *r) {            /* random.h:71-77 */
    uint64_t oldstate = r->state;
    r->state = oldstate * PCG32_MULT + r->inc;
    uint32_t xorshifted = (uint32_t) (((oldstate >> 18u) ^ oldstate) >> 27u);
    uint32_t rot = (uint32_t) (oldstate >> 59u);
    return (xorshifted >> rot) | (xorshifted << ((-rot) & 31));
}
static inline void pcg32_seed(pcg32 *r, uint64_t initstate, uint64_t initseq) { /* random.h:62-68 */
    r->state = 0u;
    r->inc = (initseq << 1u) | 1u;
    pcg32_next_uint32(r);
    r->state += initstate;
    pcg32_next_uint32(r);
}
static inline float pcg32_next_float(pcg32 *r) {                 /* random.h:117-119 */
    return b2o_u2f((pcg32_next_uint32(r) >> 9) | 0x3f800000u) - 1.0f;
}
/* PCG32Sampler::seed(seed_offset): m_rng.seed(m_base_seed + seed_offset, PCG32_DEFAULT_STREAM)
   (sampler.cpp:83-96, scalar branch) */
static inline void sampler_seed(pcg32 *r, uint64_t base_seed, uint64_t seed_offset) {
    pcg32_seed(r, base_seed + seed_offset, PCG32_DEFAULT_STREAM);
}

B2O_API void b2o_pcg32_seed(uint64_t base_seed, uint64_t seed_offset, uint64_t *state_inc) {
    pcg32 r; sampler_seed(&r, base_seed, seed_offset);
    state_inc[0] = r.state; state_inc[1] = r.inc;
}
B2O_API float b2o_pcg32_next_float(uint64_t *state_inc) {
    pcg32 r = { state_inc[0], state_inc[1] };
    float f = pcg32_next_float(&r);
    state_inc[0] = r.state;
    return f;
}
B2O_API uint32_t b2o_pcg32_next_uint32(uint64_t *state_inc) {
    pcg32 r = { state_inc[0], state_inc[1] };
    uint32_t u = pcg32_next_uint32(&r);
    state_inc[0] = r.state;
    return u;
}

/* ============================================================================================
 * Host-side scene setup mirrors
 * ============================================================================================ */
/* 4x4 matrices are row-major float[16] at the ABI; Enoki's are column-major. m(r,c) = a[4*r+c] */
typedef struct { float a[16]; } m44;
static m44 m44_identity(void) { m44 m; memset(&m, 0, sizeof m); m.a[0] = m.a[5] = m.a[10] = m.a[15] = 1.f; return m; }
/* ext/enoki/include/enoki/matrix.h:154-167: column j of the result = m0.col(0)*m1(0,j), then
   fmadd over i = 1..3 */
static m44 m44_mul(const m44 *A, const m44 *B) {
    m44 R;
    for (int j = 0; j < 4; ++j)
        for (int r = 0; r < 4; ++r) {
            float sum = A->a[4 * r + 0] * B->a[4 * 0 + j];
            for (int i = 1; i < 4; ++i)
                sum = b2o_fma(A->a[4 * r + i], B->a[4 * i + j], sum);
            R.a[4 * r + j] = sum;
        }
    return R;
}
static m44 m44_transpose(const m44 *A) {
    m44 R;
    for (int r = 0; r < 4; ++r) for (int c = 0; c < 4; ++c) R.a[4 * r + c] = A->a[4 * c + r];
    return R;
}
typedef struct { m44 m, inv_t; } xform;   /* Transform: matrix + inverse_transpose (transform.h:29-33) */
static xform xf_mul(const xform *a, const xform *b) {   /* transform.h:46-49 */
    xform r; r.m = m44_mul(&a->m, &b->m); r.inv_t = m44_mul(&a->inv_t, &b->inv_t); return r;
}
static xform xf_scale(float x, float y, float z) {      /* transform.h:167-171 */
    xform r; r.m = m44_identity(); r.inv_t = m44_identity();
    r.m.a[0] = x; r.m.a[5] = y; r.m.a[10] = z;
    r.inv_t.a[0] = b2o_rcp(x); r.inv_t.a[5] = b2o_rcp(y); r.inv_t.a[10] = b2o_rcp(z);
    return r;
}
static xform xf_translate(float x, float y, float z) {  /* transform.h:161-164 */
    xform r; r.m = m44_identity();
    r.m.a[3] = x; r.m.a[7] = y; r.m.a[11] = z;
    m44 inv = m44_identity(); inv.a[3] = -x; inv.a[7] = -y; inv.a[11] = -z;
    r.inv_t = m44_transpose(&inv);
    return r;
}
static xform xf_perspective(float fov, float near_, float far_) {  /* transform.h:203-220 */
    float recip = 1.f / (far_ - near_);
    /* deg_to_rad(fov * .5f) = value * (Pi / 180) (include/mitsuba/core/math.h) ; enoki::tan -> tanf */
    float tn = tanf((fov * .5f) * (B2O_PI / 180.f));
    float cot = 1.f / tn;
    xform r; memset(&r, 0, sizeof r);
    r.m.a[0] = cot; r.m.a[5] = cot; r.m.a[10] = far_ * recip; r.m.a[15] = 0.f;
    r.m.a[4 * 2 + 3] = -near_ * far_ * recip;
    r.m.a[4 * 3 + 2] = 1.f;
    m44 inv; memset(&inv, 0, sizeof inv);
    inv.a[0] = tn; inv.a[5] = tn; inv.a[10] = 0.f; inv.a[15] = b2o_rcp(near_);
    inv.a[4 * 2 + 3] = 1.f;
    inv.a[4 * 3 + 2] = (near_ - far_) / (far_ * near_);
    r.inv_t = m44_transpose(&inv);
    return r;
}

/* Transform::look_at — include/mitsuba/core/transform.h:241-270 */
B2O_API int b2o_look_at(const float o[3], const float tg[3], const float up_[3], float out[16]) {
    v3 origin = v3_make(o[0], o[1], o[2]), target = v3_make(tg[0], tg[1], tg[2]), up = v3_make(up_[0], up_[1], up_[2]);
    v3 dir = v3_normalize(v3_sub(target, origin));
    dir = v3_normalize(dir);
    v3 left = v3_normalize(v3_cross(up, dir));
    v3 new_up = v3_cross(dir, left);
    float m[16] = { left.x, new_up.x, dir.x, origin.x,
                    left.y, new_up.y, dir.y, origin.y,
                    left.z, new_up.z, dir.z, origin.z,
                    0.f, 0.f, 0.f, 1.f };
    memcpy(out, m, sizeof m);
    return 0;
}

/* parse_fov — src/librender/sensor.cpp (fov_axis handling); float arithmetic as there */
B2O_API float b2o_parse_fov(float fov, const char *axis, int width, int height) {
    float aspect = (float) width / (float) height;
    int mode; /* 0 = x, 1 = y, 2 = diagonal */
    if (!strcmp(axis, "x")) mode = 0;
    else if (!strcmp(axis, "y")) mode = 1;
    else if (!strcmp(axis, "diagonal")) mode = 2;
    else if (!strcmp(axis, "smaller")) mode = aspect > 1.f ? 1 : 0;
    else if (!strcmp(axis, "larger")) mode = aspect > 1.f ? 0 : 1;
    else return -1.f;
    if (mode == 0) return fov;
    if (mode == 1)
        return (180.f / B2O_PI) * (2.f * atanf(tanf(.5f * (fov * (B2O_PI / 180.f))) * aspect));
    float diagonal = 2.f * tanf(.5f * (fov * (B2O_PI / 180.f)));
    float w = diagonal / sqrtf(1.f + 1.f / (aspect * aspect));
    return (180.f / B2O_PI) * (2.f * atanf(w * .5f));
}

/* perspective_projection + update_camera_transforms —
   include/mitsuba/render/sensor.h:196-232, src/sensors/perspective.cpp:119-139 */
B2O_API int b2o_perspective_sensor(const float to_world[16], float fov_x, float near_clip, float far_clip,
                                   const B2Film *film, B2Sensor *out) {
    float fw = (float) film->width, fh = (float) film->height;
    float rel_size_x = (float) film->crop_size[0] / fw, rel_size_y = (float) film->crop_size[1] / fh;
    float rel_off_x = (float) film->crop_offset[0] / fw, rel_off_y = (float) film->crop_offset[1] / fh;
    float aspect = fw / fh;
    xform s1 = xf_scale(1.f / rel_size_x, 1.f / rel_size_y, 1.f);
    xform t1 = xf_translate(-rel_off_x, -rel_off_y, 0.f);
    xform s2 = xf_scale(-0.5f, -0.5f * aspect, 1.f);
    xform t2 = xf_translate(-1.f, -1.f / aspect, 0.f);
    xform p = xf_perspective(fov_x, near_clip, far_clip);
    xform c2s = xf_mul(&s1, &t1);
    c2s = xf_mul(&c2s, &s2);
    c2s = xf_mul(&c2s, &t2);
    c2s = xf_mul(&c2s, &p);
    /* inverse(): matrix = transpose(inverse_transpose) (transform.h:59-61) */
    m44 s2c = m44_transpose(&c2s.inv_t);
    memcpy(out->sample_to_camera, s2c.a, sizeof s2c.a);
    memcpy(out->to_world, to_world, 16 * sizeof(float));
    out->near_clip = near_clip; out->far_clip = far_clip;
    out->principal_point_offset[0] = 0.f; out->principal_point_offset[1] = 0.f;
    return 0;
}

/* ReconstructionFilter::init_discretization — src/libcore/rfilter.cpp:9-20 */
static void film_init_filter(B2Film *f, float radius, float (*eval)(float, const float *), const float *p) {
    f->filter_radius = radius;
    for (int i = 0; i < B2_FILTER_RESOLUTION; ++i)
        f->filter_lut[i] = eval((radius * (float) i) / (float) B2_FILTER_RESOLUTION, p);
    f->filter_lut[B2_FILTER_RESOLUTION] = 0.f;
    f->filter_scale = (float) B2_FILTER_RESOLUTION / radius;
    f->border = (int) ceilf(radius - .5f - 2.f * B2O_RAY_EPS);
}
static float box_eval(float x, const float *p) { return fabsf(x) <= p[0] ? 1.f : 0.f; }        /* box.cpp:31-33 */
static float gauss_eval(float x, const float *p) {                                               /* gaussian.cpp:45-47 */
    return b2o_max(0.f, expf(p[0] * (x * x)) - p[1]);
}
static void film_dims(B2Film *f, int w, int h) {
    memset(f, 0, sizeof *f);
    f->width = w; f->height = h; f->crop_offset[0] = f->crop_offset[1] = 0; f->crop_size[0] = w; f->crop_size[1] = h;
}
B2O_API int b2o_film_box(int w, int h, B2Film *f) {            /* src/rfilters/box.cpp:24-29 */
    film_dims(f, w, h);
    float radius = .5f + B2O_RAY_EPS;
    film_init_filter(f, radius, box_eval, &radius);
    return 0;
}
B2O_API int b2o_film_gaussian(int w, int h, float stddev, B2Film *f) {   /* src/rfilters/gaussian.cpp:32-43 */
    film_dims(f, w, h);
    float radius = 4 * stddev;
    float p[2];
    p[0] = -1.f / (2.f * stddev * stddev);
    p[1] = expf(p[0] * (radius * radius));
    film_init_filter(f, radius, gauss_eval, p);
    return 0;
}

/* Spiral — src/librender/spiral.cpp:8-72. order_xy[2*i] = block position of the i-th block. */
B2O_API int b2o_spiral(int width, int height, int bs, uint32_t *n_blocks, int32_t *order_xy) {
    int bx = (int) ceilf((float) width / (float) bs), by = (int) ceilf((float) height / (float) bs);
    uint32_t count = (uint32_t) (bx * by);
    *n_blocks = count;
    if (!order_xy) return 0;
    int px = bx / 2, py = by / 2, dir = 0 /* Right, Down, Left, Up */, steps_left = 1, steps = 1;
    for (uint32_t counter = 0; counter < count;) {
        order_xy[2 * counter] = px; order_xy[2 * counter + 1] = py;
        ++counter;
        if (counter != count) {
            do {
                switch (dir) {
                    case 0: ++px; break;
                    case 1: ++py; break;
                    case 2: --px; break;
                    case 3: --py; break;
                }
                if (--steps_left == 0) {
                    dir = (dir + 1) % 4;
                    if (dir == 2 || dir == 0) ++steps;
                    steps_left = steps;
                }
            } while (px < 0 || py < 0 || px >= bx || py >= by);
        }
    }
    return 0;
}

/* ============================================================================================
 * Scene: flattened triangles + a plain BVH2 (median split) used ONLY for speed. Results are
 * independent of the acceleration structure: closest hit = lexicographic min of (t, prim)
 * over all triangles whose Moeller-Trumbore test passes (see DESIGN.md "tie rule").
 * ============================================================================================ */
typedef struct {
    v3 p0, p1, p2;
    v3 n0, n1, n2;
    int has_normals;
    int mesh;
    int has_uv; float uv0[2], uv1[2], uv2[2];
} otri;

typedef struct { float lo[3], hi[3]; int left, right, first, count; } onode;

/* a BVH over the contiguous triangle range [first, first+count) of oscene.tris */
typedef struct { onode *nodes; int n_nodes; uint32_t *order; uint32_t first, count; } obvh;

typedef struct {
    const B2SceneDesc *desc;
    uint32_t n_tris;
    otri *tris;
    obvh top;             /* top-level meshes [0, n_top_meshes) */
    obvh *groups;         /* one per shapegroup (shapegroup.cpp:7-47) */
    float (*inst_lo)[3], (*inst_hi)[3];   /* world bounds of the instances (instance.cpp:83-94) */
    /* per-mesh area distribution (mesh.cpp:285-312, distr_1d.h:49-90) */
    float **cdf; float *cdf_sum, *cdf_norm; uint32_t *mesh_first;
    /* environment map (envmap.cpp): Hierarchical2D<Float,0> levels + scene bounding sphere */
    int env;                       /* index of the envmap emitter or -1 */
    int env_nlevels; float **env_level; uint32_t *env_level_w;
    float env_radius;
} oscene;

static void tri_bounds(const otri *t, float lo[3], float hi[3]) {
    const v3 *p[3] = { &t->p0, &t->p1, &t->p2 };
    lo[0] = lo[1] = lo[2] = INFINITY; hi[0] = hi[1] = hi[2] = -INFINITY;
    for (int i = 0; i < 3; ++i) {
        float c[3] = { p[i]->x, p[i]->y, p[i]->z };
        for (int k = 0; k < 3; ++k) { if (c[k] < lo[k]) lo[k] = c[k]; if (c[k] > hi[k]) hi[k] = c[k]; }
    }
}
static int g_axis; static const otri *g_tris;
static int cmp_centroid(const void *a, const void *b) {
    const otri *ta = &g_tris[*(const uint32_t *) a], *tb = &g_tris[*(const uint32_t *) b];
    float la[3], ha[3], lb[3], hb[3];
    tri_bounds(ta, la, ha); tri_bounds(tb, lb, hb);
    float ca = la[g_axis] + ha[g_axis], cb = lb[g_axis] + hb[g_axis];
    return (ca > cb) - (ca < cb);
}
static int build_node(const otri *tris, obvh *s, int first, int count) {
    int id = s->n_nodes++;
    onode *n = &s->nodes[id];
    for (int k = 0; k < 3; ++k) { n->lo[k] = INFINITY; n->hi[k] = -INFINITY; }
    for (int i = first; i < first + count; ++i) {
        float lo[3], hi[3]; tri_bounds(&tris[s->order[i]], lo, hi);
        for (int k = 0; k < 3; ++k) { if (lo[k] < n->lo[k]) n->lo[k] = lo[k]; if (hi[k] > n->hi[k]) n->hi[k] = hi[k]; }
    }
    n->first = first; n->count = count; n->left = n->right = -1;
    if (count > 4) {
        int ax = 0; float ext = n->hi[0] - n->lo[0];
        for (int k = 1; k < 3; ++k) if (n->hi[k] - n->lo[k] > ext) { ext = n->hi[k] - n->lo[k]; ax = k; }
        g_axis = ax; g_tris = tris;
        qsort(s->order + first, (size_t) count, sizeof(uint32_t), cmp_centroid);
        int half = count / 2;
        int l = build_node(tris, s, first, half);
        int r = build_node(tris, s, first + half, count - half);
        n = &s->nodes[id];
        n->left = l; n->right = r; n->count = 0;
    }
    return id;
}

static void obvh_build(const otri *tris, obvh *b, uint32_t first, uint32_t count) {
    b->first = first; b->count = count; b->n_nodes = 0;
    b->order = (uint32_t *) calloc(count ? count : 1, sizeof(uint32_t));
    b->nodes = (onode *) calloc(2 * (count ? count : 1) + 1, sizeof(onode));
    for (uint32_t i = 0; i < count; ++i) b->order[i] = first + i;
    if (count) build_node(tris, b, 0, (int) count);
}
static void obvh_free(obvh *b) { free(b->order); free(b->nodes); }

static inline v3 xf_point(const float *M, v3 p);
static void env_build(oscene *s, const B2Emitter *em);

static void mesh_vertex(const B2Mesh *m, const float *buf, uint32_t i, v3 *out) {
    out->x = buf[3 * i]; out->y = buf[3 * i + 1]; out->z = buf[3 * i + 2]; (void) m;
}

static oscene *oscene_create(const B2SceneDesc *d) {
    oscene *s = (oscene *) calloc(1, sizeof *s);
    s->desc = d;
    uint32_t nt = 0;
    s->mesh_first = (uint32_t *) calloc(d->n_meshes + 1, sizeof(uint32_t));
    for (uint32_t m = 0; m < d->n_meshes; ++m) { s->mesh_first[m] = nt; nt += d->meshes[m].n_faces; }
    s->mesh_first[d->n_meshes] = nt;
    s->n_tris = nt;
    s->tris = (otri *) calloc(nt ? nt : 1, sizeof(otri));
    s->cdf = (float **) calloc(d->n_meshes, sizeof(float *));
    s->cdf_sum = (float *) calloc(d->n_meshes, sizeof(float));
    s->cdf_norm = (float *) calloc(d->n_meshes, sizeof(float));
    for (uint32_t m = 0; m < d->n_meshes; ++m) {
        const B2Mesh *me = &d->meshes[m];
        for (uint32_t f = 0; f < me->n_faces; ++f) {
            otri *t = &s->tris[s->mesh_first[m] + f];
            uint32_t i0 = me->faces[3 * f], i1 = me->faces[3 * f + 1], i2 = me->faces[3 * f + 2];
            mesh_vertex(me, me->positions, i0, &t->p0);
            mesh_vertex(me, me->positions, i1, &t->p1);
            mesh_vertex(me, me->positions, i2, &t->p2);
            t->has_normals = me->normals != NULL;
            if (me->normals) {
                mesh_vertex(me, me->normals, i0, &t->n0);
                mesh_vertex(me, me->normals, i1, &t->n1);
                mesh_vertex(me, me->normals, i2, &t->n2);
            }
            t->mesh = (int) m;
            t->has_uv = me->texcoords != NULL;
            if (me->texcoords)
                for (int k = 0; k < 2; ++k) {
                    t->uv0[k] = me->texcoords[2 * i0 + k]; t->uv1[k] = me->texcoords[2 * i1 + k]; t->uv2[k] = me->texcoords[2 * i2 + k];
                }
        }
        /* area pmf + cdf only needed for emitter meshes: Mesh::build_pmf mesh.cpp:285-312,
           face_area mesh.h:108-116, DiscreteDistribution::update distr_1d.h:49-90 */
        if (me->emitter >= 0) {
            s->cdf[m] = (float *) calloc(me->n_faces, sizeof(float));
            double sum = 0.0;
            for (uint32_t f = 0; f < me->n_faces; ++f) {
                const otri *t = &s->tris[s->mesh_first[m] + f];
                v3 c = v3_cross(v3_sub(t->p1, t->p0), v3_sub(t->p2, t->p0));
                float area = 0.5f * sqrtf(v3_dot(c, c));
                sum += (double) area;
                s->cdf[m][f] = (float) sum;
            }
            s->cdf_sum[m] = (float) sum;
            s->cdf_norm[m] = (float) (1.0 / sum);
        }
    }
    /* top-level BVH over the meshes [0, n_top_meshes); one BVH per shapegroup; instance bounds */
    uint32_t n_top = d->n_groups ? d->n_top_meshes : d->n_meshes;
    obvh_build(s->tris, &s->top, 0, s->mesh_first[n_top]);
    s->groups = (obvh *) calloc(d->n_groups ? d->n_groups : 1, sizeof(obvh));
    for (uint32_t g = 0; g < d->n_groups; ++g) {
        uint32_t m0 = d->group_first[g], m1 = d->group_first[g + 1];
        obvh_build(s->tris, &s->groups[g], s->mesh_first[m0], s->mesh_first[m1] - s->mesh_first[m0]);
    }
    s->inst_lo = calloc(d->n_instances ? d->n_instances : 1, sizeof(float[3]));
    s->inst_hi = calloc(d->n_instances ? d->n_instances : 1, sizeof(float[3]));
    for (uint32_t i = 0; i < d->n_instances; ++i) {
        const B2Instance *in = &d->instances[i];
        const obvh *g = &s->groups[in->group];
        for (int k = 0; k < 3; ++k) { s->inst_lo[i][k] = INFINITY; s->inst_hi[i][k] = -INFINITY; }
        if (!g->count) continue;
        const onode *root = &g->nodes[0];
        for (int c = 0; c < 8; ++c) {
            float p[3] = { c & 1 ? root->hi[0] : root->lo[0], c & 2 ? root->hi[1] : root->lo[1], c & 4 ? root->hi[2] : root->lo[2] };
            for (int r = 0; r < 3; ++r) {
                const float *M = in->to_world + 4 * r;
                float w = M[3] + M[0] * p[0] + M[1] * p[1] + M[2] * p[2];
                if (w < s->inst_lo[i][r]) s->inst_lo[i][r] = w;
                if (w > s->inst_hi[i][r]) s->inst_hi[i][r] = w;
            }
        }
    }
    /* environment: Scene::Scene expands m_bbox by every top-level shape's bbox (scene.cpp:35-39);
       EnvironmentMapEmitter::set_scene (envmap.cpp:127-131) takes its bounding sphere */
    s->env = -1;
    for (uint32_t e = 0; e < d->n_emitters; ++e) if (d->emitters[e].type == B2_EMITTER_ENVMAP || d->emitters[e].type == B2_EMITTER_CONSTANT) s->env = (int) e;   /* is_environment() */
    int need_bsphere = s->env >= 0;                 /* also DirectionalEmitter::set_scene, directional.cpp:64-69 */
    for (uint32_t e = 0; e < d->n_emitters; ++e) if (d->emitters[e].type == B2_EMITTER_DIRECTIONAL) need_bsphere = 1;
    if (need_bsphere) {
        float lo[3] = { INFINITY, INFINITY, INFINITY }, hi[3] = { -INFINITY, -INFINITY, -INFINITY };
        for (uint32_t m = 0; m < n_top; ++m)
            for (uint32_t v = 0; v < d->meshes[m].n_vertices; ++v)
                for (int k = 0; k < 3; ++k) {
                    float c = d->meshes[m].positions[3 * v + k];
                    if (c < lo[k]) lo[k] = c;
                    if (c > hi[k]) hi[k] = c;
                }
        for (uint32_t i = 0; i < d->n_instances; ++i) {               /* Instance::bbox(): m_to_world * corner (instance.cpp:83-94) */
            const B2Instance *in = &d->instances[i];
            const obvh *g = &s->groups[in->group];
            if (!g->count) continue;
            float glo[3] = { INFINITY, INFINITY, INFINITY }, ghi[3] = { -INFINITY, -INFINITY, -INFINITY };
            for (uint32_t t = g->first; t < g->first + g->count; ++t) { float a[3], b[3]; tri_bounds(&s->tris[t], a, b); for (int k = 0; k < 3; ++k) { if (a[k] < glo[k]) glo[k] = a[k]; if (b[k] > ghi[k]) ghi[k] = b[k]; } }
            for (int c = 0; c < 8; ++c) {
                v3 p = v3_make(c & 1 ? ghi[0] : glo[0], c & 2 ? ghi[1] : glo[1], c & 4 ? ghi[2] : glo[2]);
                v3 w = xf_point(in->to_world, p);
                float wc[3] = { w.x, w.y, w.z };
                for (int k = 0; k < 3; ++k) { if (wc[k] < lo[k]) lo[k] = wc[k]; if (wc[k] > hi[k]) hi[k] = wc[k]; }
            }
        }
        for (uint32_t k = 0; k < d->n_shapes; ++k) {                   /* Rectangle::bbox rectangle.cpp:94-101, Sphere::bbox sphere.cpp:131-136 */
            const B2Shape *sh = &d->shapes[k];
            for (int cnr = 0; cnr < 4; ++cnr) {
                v3 w;
                if (sh->type == B2_SHAPE_RECTANGLE) w = xf_point(sh->to_world, v3_make(cnr & 1 ? 1.f : -1.f, cnr & 2 ? 1.f : -1.f, 0.f));
                else w = cnr & 1 ? v3_make(sh->center[0] - sh->radius, sh->center[1] - sh->radius, sh->center[2] - sh->radius)
                                 : v3_make(sh->center[0] + sh->radius, sh->center[1] + sh->radius, sh->center[2] + sh->radius);
                float wc[3] = { w.x, w.y, w.z };
                for (int q = 0; q < 3; ++q) { if (wc[q] < lo[q]) lo[q] = wc[q]; if (wc[q] > hi[q]) hi[q] = wc[q]; }
            }
        }
        v3 c = v3_make((lo[0] + hi[0]) * .5f, (lo[1] + hi[1]) * .5f, (lo[2] + hi[2]) * .5f);   /* bbox.h center() */
        v3 dmax = v3_sub(c, v3_make(hi[0], hi[1], hi[2]));
        float radius = sqrtf(v3_dot(dmax, dmax));                      /* bbox.h:329-332 */
        s->env_radius = b2o_max(B2O_RAY_EPS, radius * (1.f + B2O_RAY_EPS));
        if (s->env >= 0 && d->emitters[s->env].type == B2_EMITTER_ENVMAP) env_build(s, &d->emitters[s->env]);
    }
    return s;
}
static void oscene_destroy(oscene *s) {
    if (!s) return;
    for (uint32_t m = 0; m < s->desc->n_meshes; ++m) free(s->cdf[m]);
    free(s->cdf); free(s->cdf_sum); free(s->cdf_norm); free(s->mesh_first);
    obvh_free(&s->top);
    for (uint32_t g = 0; g < s->desc->n_groups; ++g) obvh_free(&s->groups[g]);
    free(s->groups); free(s->inst_lo); free(s->inst_hi);
    for (int l = 0; l < s->env_nlevels; ++l) free(s->env_level[l]);
    free(s->env_level); free(s->env_level_w);
    free(s->tris); free(s);
}

typedef struct { v3 o, d; float mint, maxt; } oray;
typedef struct { float t, u, v; uint32_t prim; int inst; } ohit;   /* t = +inf: miss; inst = -1: top-level shape */

/* Mesh::ray_intersect_triangle — include/mitsuba/render/mesh.h:194-226 */
static inline int tri_intersect(const otri *tr, const oray *ray, float *t_out, float *u_out, float *v_out) {
    v3 e1 = v3_sub(tr->p1, tr->p0), e2 = v3_sub(tr->p2, tr->p0);
    v3 pvec = v3_cross(ray->d, e2);
    float inv_det = b2o_rcp(v3_dot(e1, pvec));
    v3 tvec = v3_sub(ray->o, tr->p0);
    float u = v3_dot(tvec, pvec) * inv_det;
    int active = u >= 0.f && u <= 1.f;
    v3 qvec = v3_cross(tvec, e1);
    float v = v3_dot(ray->d, qvec) * inv_det;
    active &= v >= 0.f && u + v <= 1.f;
    float t = v3_dot(e2, qvec) * inv_det;
    active &= t >= ray->mint && t <= ray->maxt;
    *t_out = t; *u_out = u; *v_out = v;
    return active;
}

static inline int box_hit(const onode *n, const oray *r, float tmax) {
    /* conservative slab test in double (never culls a box the fp32 triangle test could hit) */
    double t0 = (double) r->mint, t1 = (double) tmax;
    const float o[3] = { r->o.x, r->o.y, r->o.z }, d[3] = { r->d.x, r->d.y, r->d.z };
    for (int k = 0; k < 3; ++k) {
        if (d[k] == 0.f) { if (o[k] < n->lo[k] - 1e-3f || o[k] > n->hi[k] + 1e-3f) return 0; continue; }
        double inv = 1.0 / (double) d[k];
        double a = ((double) n->lo[k] - o[k]) * inv, b = ((double) n->hi[k] - o[k]) * inv;
        if (a > b) { double tmp = a; a = b; b = tmp; }
        a -= 1e-4 * (fabs(a) + 1.0); b += 1e-4 * (fabs(b) + 1.0);
        if (a > t0) t0 = a;
        if (b < t1) t1 = b;
        if (t0 > t1) return 0;
    }
    return 1;
}

/* tie rule: smaller t wins; on exactly equal t the larger (instance + 1, prim) key wins */
static inline int hit_better(float t, int inst, uint32_t prim, const ohit *best) {
    if (t < best->t) return 1;
    if (t > best->t || best->prim == 0xffffffffu) return t == best->t;
    uint64_t k = ((uint64_t) (inst + 1) << 32) | prim, kb = ((uint64_t) (best->inst + 1) << 32) | best->prim;
    return k > kb;
}

/* ---- analytic shapes: Rectangle::ray_intersect_preliminary (rectangle.cpp:139-158) and
 *      Sphere::ray_intersect_preliminary (sphere.cpp:273-305, in double; math.h:371-413) ---- */
static inline v3 xf_point(const float *M, v3 p);
static inline v3 xf_vector(const float *M, v3 d);
static int shape_intersect(const B2Shape *sh, const oray *ray, float *t_out, float *u_out, float *v_out) {
    *u_out = 0.f; *v_out = 0.f;
    if (sh->type == B2_SHAPE_RECTANGLE) {
        v3 o = xf_point(sh->to_object, ray->o), d = xf_vector(sh->to_object, ray->d);
        float t = -o.z * b2o_rcp(d.z);
        float lx = b2o_fma(d.x, t, o.x), ly = b2o_fma(d.y, t, o.y);
        *t_out = t; *u_out = lx; *v_out = ly;
        return t >= ray->mint && t <= ray->maxt && fabsf(lx) <= 1.f && fabsf(ly) <= 1.f;
    }
    double mint = (double) ray->mint, maxt = (double) ray->maxt;
    double ox = (double) ray->o.x - (double) sh->center[0], oy = (double) ray->o.y - (double) sh->center[1],
           oz = (double) ray->o.z - (double) sh->center[2];
    double dx = ray->d.x, dy = ray->d.y, dz = ray->d.z;
    double A = (dx * dx + dy * dy) + dz * dz;
    double B = 2.0 * ((ox * dx + oy * dy) + oz * dz);
    double C = ((ox * ox + oy * oy) + oz * oz) - (double) sh->radius * (double) sh->radius;
    /* solve_quadratic */
    int linear_case = A == 0.0, valid_linear = linear_case && B != 0.0;
    double x0 = -C / B, x1 = x0;
    double discrim = __builtin_fma(B, B, -(4.0 * A * C));
    int valid_quadratic = !linear_case && discrim >= 0.0;
    if (valid_quadratic) {
        double sq = sqrt(discrim);
        double temp = -0.5 * (B + copysign(sq, B));
        double x0p = temp / A, x1p = C / temp;
        double x0m = x0p < x1p ? x0p : x1p, x1m = x0p > x1p ? x0p : x1p;     /* min/max (NaN-free here) */
        x0 = x0m; x1 = x1m;
    }
    int found = valid_linear || valid_quadratic;
    int out_bounds = !(x0 <= maxt && x1 >= mint), in_bounds = x0 < mint && x1 > maxt;
    *t_out = x0 < mint ? (float) x1 : (float) x0;
    return found && !out_bounds && !in_bounds;
}

static void bvh_intersect(const oscene *s, const obvh *b, const oray *ray, int inst, int brute, ohit *best) {
    if (!b->count) return;
    float t, u, v;
    if (brute) {
        for (uint32_t i = b->first; i < b->first + b->count; ++i)
            if (tri_intersect(&s->tris[i], ray, &t, &u, &v) && hit_better(t, inst, i, best)) {
                best->t = t; best->u = u; best->v = v; best->prim = i; best->inst = inst;
            }
        return;
    }
    int stack[128], sp = 0; stack[sp++] = 0;
    while (sp) {
        const onode *n = &b->nodes[stack[--sp]];
        if (!box_hit(n, ray, best->t < ray->maxt ? best->t : ray->maxt)) continue;
        if (n->left < 0) {
            for (int i = n->first; i < n->first + n->count; ++i) {
                uint32_t prim = b->order[i];
                if (tri_intersect(&s->tris[prim], ray, &t, &u, &v) && hit_better(t, inst, prim, best)) {
                    best->t = t; best->u = u; best->v = v; best->prim = prim; best->inst = inst;
                }
            }
        } else { stack[sp++] = n->left; stack[sp++] = n->right; }
    }
}
static int bvh_occluded(const oscene *s, const obvh *b, const oray *ray, int brute) {
    if (!b->count) return 0;
    float t, u, v;
    if (brute) {
        for (uint32_t i = b->first; i < b->first + b->count; ++i) if (tri_intersect(&s->tris[i], ray, &t, &u, &v)) return 1;
        return 0;
    }
    int stack[128], sp = 0; stack[sp++] = 0;
    while (sp) {
        const onode *n = &b->nodes[stack[--sp]];
        if (!box_hit(n, ray, ray->maxt)) continue;
        if (n->left < 0) {
            for (int i = n->first; i < n->first + n->count; ++i)
                if (tri_intersect(&s->tris[b->order[i]], ray, &t, &u, &v)) return 1;
        } else { stack[sp++] = n->left; stack[sp++] = n->right; }
    }
    return 0;
}

/* Transform::transform_affine for points / vectors / normals (transform.h:88-151): row-major M */
static inline v3 xf_point(const float *M, v3 p) {
    v3 r;
    r.x = b2o_fma(M[2], p.z, b2o_fma(M[1], p.y, b2o_fma(M[0], p.x, M[3])));
    r.y = b2o_fma(M[6], p.z, b2o_fma(M[5], p.y, b2o_fma(M[4], p.x, M[7])));
    r.z = b2o_fma(M[10], p.z, b2o_fma(M[9], p.y, b2o_fma(M[8], p.x, M[11])));
    return r;
}
static inline v3 xf_vector(const float *M, v3 d) {
    v3 r;
    r.x = b2o_fma(M[2], d.z, b2o_fma(M[1], d.y, M[0] * d.x));
    r.y = b2o_fma(M[6], d.z, b2o_fma(M[5], d.y, M[4] * d.x));
    r.z = b2o_fma(M[10], d.z, b2o_fma(M[9], d.y, M[8] * d.x));
    return r;
}
/* normals use the inverse transpose: columns of `inv` (the inverse matrix, row-major) */
static inline v3 xf_normal(const float *inv, v3 n) {
    v3 r;
    r.x = b2o_fma(inv[8], n.z, b2o_fma(inv[4], n.y, inv[0] * n.x));
    r.y = b2o_fma(inv[9], n.z, b2o_fma(inv[5], n.y, inv[1] * n.x));
    r.z = b2o_fma(inv[10], n.z, b2o_fma(inv[6], n.y, inv[2] * n.x));
    return r;
}
/* Instance::ray_intersect_preliminary: m_to_object.transform_affine(ray) — origin as a point,
   direction as an UN-normalised vector, mint/maxt unchanged (instance.cpp:107-117) */
static inline oray instance_ray(const B2Instance *in, const oray *ray) {
    oray r; r.o = xf_point(in->to_object, ray->o); r.d = xf_vector(in->to_object, ray->d);
    r.mint = ray->mint; r.maxt = ray->maxt; return r;
}
static inline int inst_box_hit(const oscene *s, uint32_t i, const oray *ray, float tmax) {
    onode n; for (int k = 0; k < 3; ++k) { n.lo[k] = s->inst_lo[i][k] - 1e-4f * (fabsf(s->inst_lo[i][k]) + 1.f); n.hi[k] = s->inst_hi[i][k] + 1e-4f * (fabsf(s->inst_hi[i][k]) + 1.f); }
    return box_hit(&n, ray, tmax);
}

/* Scene::ray_intersect_preliminary (scene.cpp:131-137) with the order-independent tie rule */
static ohit scene_intersect(const oscene *s, const oray *ray, int brute) {
    ohit best; best.t = INFINITY; best.u = best.v = 0.f; best.prim = 0xffffffffu; best.inst = -1;
    bvh_intersect(s, &s->top, ray, -1, brute, &best);
    for (uint32_t i = 0; i < s->desc->n_instances; ++i) {
        if (!brute && !inst_box_hit(s, i, ray, best.t < ray->maxt ? best.t : ray->maxt)) continue;
        oray ro = instance_ray(&s->desc->instances[i], ray);
        bvh_intersect(s, &s->groups[s->desc->instances[i].group], &ro, (int) i, brute, &best);
    }
    for (uint32_t k = 0; k < s->desc->n_shapes; ++k) {          /* primitive index = n_tris + k */
        float t, u, v;
        if (shape_intersect(&s->desc->shapes[k], ray, &t, &u, &v) && hit_better(t, -1, s->n_tris + k, &best)) {
            best.t = t; best.u = u; best.v = v; best.prim = s->n_tris + k; best.inst = -1;
        }
    }
    return best;
}
/* Scene::ray_test (scene.cpp:154-162) */
static int scene_occluded(const oscene *s, const oray *ray, int brute) {
    if (bvh_occluded(s, &s->top, ray, brute)) return 1;
    for (uint32_t i = 0; i < s->desc->n_instances; ++i) {
        if (!brute && !inst_box_hit(s, i, ray, ray->maxt)) continue;
        oray ro = instance_ray(&s->desc->instances[i], ray);
        if (bvh_occluded(s, &s->groups[s->desc->instances[i].group], &ro, brute)) return 1;
    }
    for (uint32_t k = 0; k < s->desc->n_shapes; ++k) {
        float t, u, v;
        if (shape_intersect(&s->desc->shapes[k], ray, &t, &u, &v)) return 1;
    }
    return 0;
}

/* ============================================================================================
 * SurfaceInteraction — mesh.cpp:449-545, interaction.h:153-156,571-596, vector.h:116-136
 * ============================================================================================ */
typedef struct {
    float t; v3 p, n, sh_n, sh_s, sh_t, wi;
    uint32_t prim; int mesh;      /* mesh = -1: miss; -2 - k: analytic shape k */
    float uv[2];
    int bsdf, emitter;            /* of the shape that was hit (-1: none) */
} osi;

static inline void coordinate_system_s(v3 n, v3 *s) {   /* vector.h:116-136 (first vector only) */
    float sign = b2o_sign(n.z);
    float a = -b2o_rcp(sign + n.z);
    float b = (n.x * n.y) * a;
    s->x = b2o_mulsign(b2o_sqr(n.x) * a, n.z) + 1.f;
    s->y = b2o_mulsign(b, n.z);
    s->z = b2o_mulsign_neg(n.x, n.z);
}
static inline void coordinate_system(v3 n, v3 *s, v3 *t) {
    float sign = b2o_sign(n.z);
    float a = -b2o_rcp(sign + n.z);
    float b = (n.x * n.y) * a;
    coordinate_system_s(n, s);
    t->x = b; t->y = sign + b2o_sqr(n.y) * a; t->z = -n.y;
}
static inline v3 to_local(const osi *si, v3 v) { return v3_make(v3_dot(v, si->sh_s), v3_dot(v, si->sh_t), v3_dot(v, si->sh_n)); }
static inline v3 to_world(const osi *si, v3 v) {   /* frame.h:35-37: s*v.x + t*v.y + n*v.z */
    return v3_add(v3_add(v3_scale(si->sh_s, v.x), v3_scale(si->sh_t, v.y)), v3_scale(si->sh_n, v.z));
}

static inline v3 xf_normal(const float *inv, v3 n);
/* Rectangle::compute_surface_interaction (rectangle.cpp:174-207, frame from update() :82-92) and
 * Sphere::compute_surface_interaction (sphere.cpp:333-395) */
static void si_fill_shape(const oscene *s, const oray *ray, const ohit *h, osi *si) {
    uint32_t k = h->prim - s->n_tris;
    const B2Shape *sh = &s->desc->shapes[k];
    si->mesh = -2 - (int) k; si->bsdf = sh->bsdf; si->emitter = sh->emitter;
    v3 dp_du;
    v3 hitp = v3_make(b2o_fma(ray->d.x, h->t, ray->o.x), b2o_fma(ray->d.y, h->t, ray->o.y), b2o_fma(ray->d.z, h->t, ray->o.z));
    if (sh->type == B2_SHAPE_RECTANGLE) {
        si->p = hitp;
        si->n = v3_normalize(xf_normal(sh->to_object, v3_make(0.f, 0.f, 1.f)));
        si->sh_n = si->n;
        dp_du = xf_vector(sh->to_world, v3_make(2.f, 0.f, 0.f));
        si->uv[0] = b2o_fma(h->u, .5f, .5f); si->uv[1] = b2o_fma(h->v, .5f, .5f);
    } else {
        v3 c = v3_make(sh->center[0], sh->center[1], sh->center[2]);
        v3 n = v3_normalize(v3_sub(hitp, c));
        si->p = v3_make(b2o_fma(n.x, sh->radius, c.x), b2o_fma(n.y, sh->radius, c.y), b2o_fma(n.z, sh->radius, c.z));
        v3 local = xf_point(sh->to_object, si->p);
        float theta = 2.f * b2o_asin(.5f * sqrtf(b2o_sqr(local.x) + b2o_sqr(local.y) + b2o_sqr(local.z - b2o_sign(local.z))));
        if (!(local.z >= 0.f)) theta = 3.14159265358979323846f - theta;                 /* unit_angle_z */
        float phi = b2o_atan2(local.y, local.x);
        if (phi < 0.f) phi += 2.f * B2O_PI;
        si->uv[0] = phi * (.5f * B2O_INV_PI); si->uv[1] = theta * B2O_INV_PI;
        dp_du = v3_scale(xf_vector(sh->to_world, v3_make(-local.y, local.x, 0.f)), 2.f * B2O_PI);
        if (sh->flip_normals) n = v3_neg(n);
        si->sh_n = n; si->n = n;
    }
    si->sh_s = v3_normalize(v3_fnmadd(si->sh_n, v3_dot(si->sh_n, dp_du), dp_du));    /* interaction.h:153-156 */
    si->sh_t = v3_cross(si->sh_n, si->sh_s);
    si->wi = to_local(si, v3_neg(ray->d));
}

static void si_fill(const oscene *s, const oray *ray, const ohit *h, osi *si) {
    memset(si, 0, sizeof *si);
    si->t = h->t; si->prim = h->prim; si->mesh = -1;
    si->bsdf = -1; si->emitter = -1;
    if (!(h->t < INFINITY)) { si->wi = v3_neg(ray->d); return; }   /* interaction.h:591 */
    if (h->prim >= s->n_tris) { si_fill_shape(s, ray, h, si); return; }
    const otri *tr = &s->tris[h->prim];
    si->mesh = tr->mesh;
    si->bsdf = s->desc->meshes[tr->mesh].bsdf; si->emitter = s->desc->meshes[tr->mesh].emitter;
    float b1 = h->u, b2 = h->v, b0 = 1.f - b1 - b2;
    v3 dp0 = v3_sub(tr->p1, tr->p0), dp1 = v3_sub(tr->p2, tr->p0);
    si->p = v3_add(v3_add(v3_scale(tr->p0, b0), v3_scale(tr->p1, b1)), v3_scale(tr->p2, b2));  /* mesh.cpp:484 */
    si->n = v3_normalize(v3_cross(dp0, dp1));                                                   /* mesh.cpp:487 */
    v3 dp_du; coordinate_system_s(si->n, &dp_du);                                               /* mesh.cpp:491 */
    si->uv[0] = b1; si->uv[1] = b2;                                                             /* mesh.cpp:490 */
    if (tr->has_uv) {                                                                            /* mesh.cpp:492-511 */
        for (int k = 0; k < 2; ++k) si->uv[k] = (tr->uv0[k] * b0 + tr->uv1[k] * b1) + tr->uv2[k] * b2;
        float duv0x = tr->uv1[0] - tr->uv0[0], duv0y = tr->uv1[1] - tr->uv0[1],
              duv1x = tr->uv2[0] - tr->uv0[0], duv1y = tr->uv2[1] - tr->uv0[1];
        float det = b2o_fma(duv0x, duv1y, -(duv0y * duv1x)), inv_det = b2o_rcp(det);
        if (det != 0.f) {
            /* fmsub(duv1.y, dp0, duv0.y * dp1) * inv_det */
            dp_du = v3_make(b2o_fma(duv1y, dp0.x, -(duv0y * dp1.x)) * inv_det, b2o_fma(duv1y, dp0.y, -(duv0y * dp1.y)) * inv_det,
                            b2o_fma(duv1y, dp0.z, -(duv0y * dp1.z)) * inv_det);
        }
    }
    if (tr->has_normals)                                                                         /* mesh.cpp:514-519 */
        si->sh_n = v3_normalize(v3_add(v3_add(v3_scale(tr->n0, b0), v3_scale(tr->n1, b1)), v3_scale(tr->n2, b2)));
    else
        si->sh_n = si->n;
    if (h->inst >= 0) {      /* Instance::compute_surface_interaction instance.cpp:124-142 */
        const B2Instance *in = &s->desc->instances[h->inst];
        si->p = xf_point(in->to_world, si->p);
        si->n = v3_normalize(xf_normal(in->to_object, si->n));
        si->sh_n = v3_normalize(xf_normal(in->to_object, si->sh_n));
        dp_du = xf_vector(in->to_world, dp_du);
    }
    /* initialize_sh_frame interaction.h:153-156 (for instances: the call in interaction.h:587-588,
       made after dp_du was mapped to world space) */
    si->sh_s = v3_normalize(v3_fnmadd(si->sh_n, v3_dot(si->sh_n, dp_du), dp_du));
    si->sh_t = v3_cross(si->sh_n, si->sh_s);
    si->wi = to_local(si, v3_neg(ray->d));                                                       /* interaction.h:591 */
}

/* ============================================================================================
 * Warps — include/mitsuba/core/warp.h
 * ============================================================================================ */
static inline void square_to_uniform_disk_concentric(float sx, float sy, float *ox, float *oy) { /* warp.h:53-90 */
    float x = b2o_fma(2.f, sx, -1.f), y = b2o_fma(2.f, sy, -1.f);
    int is_zero = x == 0.f && y == 0.f, q13 = fabsf(x) < fabsf(y);
    float r = q13 ? y : x, rp = q13 ? x : y;
    float phi = ((.25f * B2O_PI) * rp) / r;
    if (q13) phi = .5f * B2O_PI - phi;
    if (is_zero) phi = 0.f;
    float s, c; b2o_sincos(phi, &s, &c);
    *ox = r * c; *oy = r * s;
}
static inline v3 square_to_cosine_hemisphere(float sx, float sy) {   /* warp.h:325-333 */
    float px, py; square_to_uniform_disk_concentric(sx, sy, &px, &py);
    float z = b2o_safe_sqrt(1.f - (px * px + py * py));
    return v3_make(px, py, z);
}

/* ============================================================================================
 * Fresnel + microfacet — include/mitsuba/render/fresnel.h, microfacet.h
 * ============================================================================================ */
static inline void fresnel(float cos_theta_i, float eta, float *r_out, float *cos_theta_t_out,
                           float *eta_it_out, float *eta_ti_out) {       /* fresnel.h:34-70 */
    int outside = cos_theta_i >= 0.f;
    float rcp_eta = b2o_rcp(eta), eta_it = outside ? eta : rcp_eta, eta_ti = outside ? rcp_eta : eta;
    float cos_theta_t_sqr = b2o_fma(-b2o_fma(-cos_theta_i, cos_theta_i, 1.f), eta_ti * eta_ti, 1.f);
    float cos_theta_i_abs = fabsf(cos_theta_i);
    float cos_theta_t_abs = b2o_safe_sqrt(cos_theta_t_sqr);
    int index_matched = eta == 1.f, special_case = index_matched || cos_theta_i_abs == 0.f;
    float r_sc = index_matched ? 0.f : 1.f;
    float a_s = b2o_fma(-eta_it, cos_theta_t_abs, cos_theta_i_abs) / b2o_fma(eta_it, cos_theta_t_abs, cos_theta_i_abs);
    float a_p = b2o_fma(-eta_it, cos_theta_i_abs, cos_theta_t_abs) / b2o_fma(eta_it, cos_theta_i_abs, cos_theta_t_abs);
    float r = .5f * (b2o_sqr(a_s) + b2o_sqr(a_p));
    if (special_case) r = r_sc;
    *r_out = r; *cos_theta_t_out = b2o_mulsign_neg(cos_theta_t_abs, cos_theta_i);
    *eta_it_out = eta_it; *eta_ti_out = eta_ti;
}
static inline float fresnel_conductor(float cos_theta_i, float eta_r, float eta_i) {   /* fresnel.h:92-116 */
    float cos_theta_i_2 = cos_theta_i * cos_theta_i, sin_theta_i_2 = 1.f - cos_theta_i_2,
          sin_theta_i_4 = sin_theta_i_2 * sin_theta_i_2;
    float temp_1 = eta_r * eta_r - eta_i * eta_i - sin_theta_i_2,
          a_2_pb_2 = b2o_safe_sqrt(temp_1 * temp_1 + 4.f * eta_i * eta_i * eta_r * eta_r),
          a = b2o_safe_sqrt(.5f * (a_2_pb_2 + temp_1));
    float term_1 = a_2_pb_2 + cos_theta_i_2, term_2 = 2.f * cos_theta_i * a;
    float r_s = (term_1 - term_2) / (term_1 + term_2);
    float term_3 = a_2_pb_2 * cos_theta_i_2 + sin_theta_i_4, term_4 = term_2 * sin_theta_i_2;
    float r_p = r_s * (term_3 - term_4) / (term_3 + term_4);
    return .5f * (r_s + r_p);
}

typedef struct { int type; float au, av; int sample_visible; } omf;
static inline omf mf_make(const B2Bsdf *b) {     /* microfacet.h configure(): alpha = max(alpha, 1e-4) */
    omf m; m.type = b->distribution; m.au = b2o_max(b->alpha_u, 1e-4f); m.av = b2o_max(b->alpha_v, 1e-4f);
    m.sample_visible = b->sample_visible; return m;
}
static inline float mf_eval(const omf *d, v3 m) {   /* microfacet.h:184-203 */
    float alpha_uv = d->au * d->av, cos_theta = m.z, cos_theta_2 = b2o_sqr(cos_theta), result;
    if (d->type == B2_MF_BECKMANN)
        result = b2o_exp(-(b2o_sqr(m.x / d->au) + b2o_sqr(m.y / d->av)) / cos_theta_2) /
                 (B2O_PI * alpha_uv * b2o_sqr(cos_theta_2));
    else
        result = b2o_rcp(B2O_PI * alpha_uv * b2o_sqr(b2o_sqr(m.x / d->au) + b2o_sqr(m.y / d->av) + b2o_sqr(m.z)));
    return result * cos_theta > 1e-20f ? result : 0.f;
}
static inline float mf_smith_g1(const omf *d, v3 v, v3 m) {    /* microfacet.h:331-355 */
    float xy_alpha_2 = b2o_sqr(d->au * v.x) + b2o_sqr(d->av * v.y),
          tan_theta_alpha_2 = xy_alpha_2 / b2o_sqr(v.z), result;
    if (d->type == B2_MF_BECKMANN) {
        float a = b2o_rsqrt(tan_theta_alpha_2), a_sqr = b2o_sqr(a);
        result = a >= 1.6f ? 1.f : (3.535f * a + 2.181f * a_sqr) / (1.f + 2.276f * a + 2.577f * a_sqr);
    } else
        result = 2.f / (1.f + sqrtf(1.f + tan_theta_alpha_2));
    if (xy_alpha_2 == 0.f) result = 1.f;
    if (v3_dot(v, m) * v.z <= 0.f) result = 0.f;
    return result;
}
static inline void mf_sample_visible_11(const omf *d, float cos_theta_i, float sx, float sy, float *ox, float *oy) {
    if (d->type == B2_MF_BECKMANN) {                                    /* microfacet.h:359-394 */
        float tan_theta_i = b2o_safe_sqrt(b2o_fma(-cos_theta_i, cos_theta_i, 1.f)) / cos_theta_i;
        float cot_theta_i = b2o_rcp(tan_theta_i);
        float maxval = b2o_erf(cot_theta_i);
        sx = b2o_max(b2o_min(sx, 1.f - 1e-6f), 1e-6f);
        sy = b2o_max(b2o_min(sy, 1.f - 1e-6f), 1e-6f);
        float x = maxval - (maxval + 1.f) * b2o_erf(sqrtf(-b2o_log(sx)));
        sx *= 1.f + maxval + B2O_INV_SQRT_PI * tan_theta_i * b2o_exp(-b2o_sqr(cot_theta_i));
        for (int i = 0; i < 3; ++i) {
            float slope = b2o_erfinv(x),
                  value = 1.f + x + B2O_INV_SQRT_PI * tan_theta_i * b2o_exp(-b2o_sqr(slope)) - sx,
                  derivative = 1.f - slope * tan_theta_i;
            x -= value / derivative;
        }
        *ox = b2o_erfinv(x); *oy = b2o_erfinv(b2o_fma(2.f, sy, -1.f));
    } else {                                                             /* microfacet.h:395-410 */
        float px, py; square_to_uniform_disk_concentric(sx, sy, &px, &py);
        float s = .5f * (1.f + cos_theta_i);
        /* enoki::lerp(a, b, t) = fmadd(b, t, fnmadd(a, t, a)) (array_router.h) */
        float a0 = b2o_safe_sqrt(1.f - b2o_sqr(px));
        py = b2o_fma(py, s, b2o_fma(-a0, s, a0));
        float x = px, y = py, z = b2o_safe_sqrt(1.f - (px * px + py * py));
        float sin_theta_i = b2o_safe_sqrt(1.f - b2o_sqr(cos_theta_i));
        float norm = b2o_rcp(b2o_fma(sin_theta_i, y, cos_theta_i * z));
        *ox = b2o_fma(cos_theta_i, y, -(sin_theta_i * z)) * norm; *oy = x * norm;
    }
}
/* MicrofacetDistribution::sample, visible-normal branch — microfacet.h:283-316 */
static inline v3 mf_sample(const omf *d, v3 wi, float sx, float sy, float *pdf) {
    v3 wi_p = v3_normalize(v3_make(d->au * wi.x, d->av * wi.y, wi.z));
    /* Frame3f::sincos_phi (frame.h:105-117): sin_theta_2 = fnmadd(z,z,1); inv = rsqrt; result = (y,x)*inv,
       or (0,1) if |sin_theta_2| <= 4 eps */
    float sin_theta_2 = b2o_fma(-wi_p.z, wi_p.z, 1.f), inv_sin_theta = b2o_rsqrt(sin_theta_2);
    float sin_phi = wi_p.y * inv_sin_theta, cos_phi = wi_p.x * inv_sin_theta;
    if (fabsf(sin_theta_2) <= 4.f * 5.9604644775390625e-08f) { sin_phi = 0.f; cos_phi = 1.f; }
    float cos_theta = wi_p.z;
    float slx, sly; mf_sample_visible_11(d, cos_theta, sx, sy, &slx, &sly);
    float s0 = b2o_fma(cos_phi, slx, -(sin_phi * sly)) * d->au;
    float s1 = b2o_fma(sin_phi, slx, cos_phi * sly) * d->av;
    v3 m = v3_normalize(v3_make(-s0, -s1, 1.f));
    *pdf = mf_eval(d, m) * mf_smith_g1(d, wi, m) * fabsf(v3_dot(wi, m)) / wi.z;
    return m;
}

/* ============================================================================================
 * Textures — src/textures/checkerboard.cpp:46-63, src/textures/bitmap.cpp:386-474
 * ============================================================================================ */
static inline void tex_to_uv(const B2Texture *t, const float uv[2], float *u, float *v) {
    /* Transform3f::transform_affine (transform.h): result = col2; result = fmadd(col_i, uv_i, result) */
    const float *M = t->to_uv;
    *u = b2o_fma(M[1], uv[1], b2o_fma(M[0], uv[0], M[2]));
    *v = b2o_fma(M[4], uv[1], b2o_fma(M[3], uv[0], M[5]));
}
static inline int tex_wrap(const B2Texture *t, int value, int res) {            /* bitmap.cpp:386-401 */
    if (t->wrap == B2_WRAP_CLAMP) return value < 0 ? 0 : (value > res - 1 ? res - 1 : value);
    int div = value / res, mod = value - div * res;                             /* enoki::divisor: truncating */
    if (mod < 0) mod += res;
    if (t->wrap == B2_WRAP_MIRROR && !(((div & 1) == 0) ^ (value < 0))) mod = res - 1 - mod;
    return mod;
}
static inline v3 tex_fetch(const B2Texture *t, int x, int y) {
    const float *p = t->data + (size_t) (x + y * t->width) * (size_t) t->channels;
    return t->channels == 1 ? v3_make(p[0], p[0], p[0]) : v3_make(p[0], p[1], p[2]);
}
static v3 texture_eval(const B2Texture *t, const float si_uv[2]) {
    float u, v; tex_to_uv(t, si_uv, &u, &v);
    if (t->type == B2_TEX_CHECKERBOARD) {                                       /* checkerboard.cpp:46-63 */
        int mx = u - floorf(u) > .5f, my = v - floorf(v) > .5f;
        const float *c = mx == my ? t->color0 : t->color1;
        return v3_make(c[0], c[1], c[2]);
    }
    if (t->filter == B2_TEXFILTER_BILINEAR) {                                   /* bitmap.cpp:412-457 */
        u = b2o_fma(u, (float) t->width, -.5f); v = b2o_fma(v, (float) t->height, -.5f);
        int ix = (int) floorf(u), iy = (int) floorf(v);
        float w1x = u - (float) ix, w1y = v - (float) iy, w0x = 1.f - w1x, w0y = 1.f - w1y;
        int x0 = tex_wrap(t, ix, t->width), x1 = tex_wrap(t, ix + 1, t->width),
            y0 = tex_wrap(t, iy, t->height), y1 = tex_wrap(t, iy + 1, t->height);
        v3 v00 = tex_fetch(t, x0, y0), v10 = tex_fetch(t, x1, y0), v01 = tex_fetch(t, x0, y1), v11 = tex_fetch(t, x1, y1);
        v3 v0 = v3_make(b2o_fma(w0x, v00.x, w1x * v10.x), b2o_fma(w0x, v00.y, w1x * v10.y), b2o_fma(w0x, v00.z, w1x * v10.z));
        v3 v1 = v3_make(b2o_fma(w0x, v01.x, w1x * v11.x), b2o_fma(w0x, v01.y, w1x * v11.y), b2o_fma(w0x, v01.z, w1x * v11.z));
        return v3_make(b2o_fma(w0y, v0.x, w1y * v1.x), b2o_fma(w0y, v0.y, w1y * v1.y), b2o_fma(w0y, v0.z, w1y * v1.z));
    }
    u *= (float) t->width; v *= (float) t->height;                              /* bitmap.cpp:458-473 */
    return tex_fetch(t, tex_wrap(t, (int) floorf(u), t->width), tex_wrap(t, (int) floorf(v), t->height));
}
/* the BSDF record with its textured parameters evaluated at si (Texture::eval(si) inside BSDF::eval/sample).
 * TwoSidedBRDF (twosided.cpp:94-186): the front record for cos_theta(wi) > 0, the back record (with wi.z and the
 * resulting wo.z mirrored: *flip = 1) for cos_theta(wi) < 0, nothing at exactly 0 (*side = 0). */
static inline B2Bsdf bsdf_at(const B2SceneDesc *d, const B2Bsdf *b, const osi *si, int *flip, int *smooth) {
    *flip = 0;
    *smooth = b->type != B2_BSDF_DIELECTRIC && b->type != B2_BSDF_CONDUCTOR;       /* has_flag(flags, Smooth), path.cpp:155 */
    if (b->twosided_back > 0) {
        const B2Bsdf *back = &d->bsdfs[b->twosided_back - 1];
        *smooth = *smooth || (back->type != B2_BSDF_DIELECTRIC && back->type != B2_BSDF_CONDUCTOR);   /* union of the nested flags */
        if (si->wi.z < 0.f) { b = back; *flip = 1; }
        else if (!(si->wi.z > 0.f)) { B2Bsdf z; memset(&z, 0, sizeof z); z.type = B2_BSDF_NULL; *flip = 0; return z; }   /* both masks false */
    }
    B2Bsdf r = *b;
    r.twosided_back = 0;
    if (b->reflectance_texture > 0 && (uint32_t) b->reflectance_texture <= d->n_textures) {
        v3 c = texture_eval(&d->textures[b->reflectance_texture - 1], si->uv);
        r.reflectance[0] = c.x; r.reflectance[1] = c.y; r.reflectance[2] = c.z;
    }
    return r;
}

/* ============================================================================================
 * BSDFs — src/bsdfs/{diffuse,dielectric,roughconductor}.cpp
 * ============================================================================================ */
typedef struct { v3 wo; float pdf, eta; int delta; } obs;

static inline int bsdf_is_smooth(const B2Bsdf *b) {   /* has_flag(flags, BSDFFlags::Smooth), path.cpp:155 */
    return b->type != B2_BSDF_DIELECTRIC && b->type != B2_BSDF_CONDUCTOR;    /* delta-only BSDFs */
}

/* SmoothPlastic: diffuse_reflectance / (1 - fdr_int) with the optional nonlinear variant (plastic.cpp:231-232, 262-263).
 * Spectrum /= scalar multiplies by the reciprocal; Spectrum / Spectrum divides. */
static inline v3 plastic_diffuse(const B2Bsdf *b) {
    v3 d = v3_make(b->reflectance[0], b->reflectance[1], b->reflectance[2]);
    if (b->nonlinear)
        return v3_make(d.x / (1.f - d.x * b->fdr_int), d.y / (1.f - d.y * b->fdr_int), d.z / (1.f - d.z * b->fdr_int));
    return v3_scale(d, 1.f / (1.f - b->fdr_int));
}
static inline float fresnel_r(float cos_theta_i, float eta) {
    float r, ct, eit, eti; fresnel(cos_theta_i, eta, &r, &ct, &eit, &eti); return r;
}

static v3 bsdf_eval(const B2Bsdf *b, const osi *si, v3 wo) {
    float cti = si->wi.z, cto = wo.z;
    v3 zero = v3_make(0.f, 0.f, 0.f);
    switch (b->type) {
        case B2_BSDF_DIFFUSE: case B2_BSDF_NULL: {                        /* diffuse.cpp:104-119 */
            if (!(cti > 0.f && cto > 0.f)) return zero;
            const float *r = b->reflectance;
            return v3_make((r[0] * B2O_INV_PI) * cto, (r[1] * B2O_INV_PI) * cto, (r[2] * B2O_INV_PI) * cto);
        }
        case B2_BSDF_ROUGHCONDUCTOR: {                                     /* roughconductor.cpp:277-343 */
            if (!(cti > 0.f && cto > 0.f)) return zero;
            v3 H = v3_normalize(v3_add(wo, si->wi));
            omf d = mf_make(b);
            float D = mf_eval(&d, H);
            if (D == 0.f) return zero;
            float G = mf_smith_g1(&d, si->wi, H) * mf_smith_g1(&d, wo, H);
            float result = D * G / (4.f * cti);
            float c = v3_dot(si->wi, H);
            v3 F = v3_make(fresnel_conductor(c, b->cond_eta[0], b->cond_k[0]),
                           fresnel_conductor(c, b->cond_eta[1], b->cond_k[1]),
                           fresnel_conductor(c, b->cond_eta[2], b->cond_k[2]));
            v3 res = v3_make(result * b->specular_reflectance[0], result * b->specular_reflectance[1],
                             result * b->specular_reflectance[2]);
            return v3_mul(F, res);
        }
        case B2_BSDF_ROUGHDIELECTRIC: {                                    /* roughdielectric.cpp:312-384 */
            if (cti == 0.f) return zero;
            int reflect = cti * cto > 0.f;
            float eta = cti > 0.f ? b->eta : b2o_rcp(b->eta), inv_eta = cti > 0.f ? b2o_rcp(b->eta) : b->eta;   /* m_inv_eta = 1 / m_eta (:172) */
            v3 m = v3_normalize(v3_add(si->wi, v3_scale(wo, reflect ? 1.f : eta)));
            if (m.z < 0.f || (m.z == 0.f && signbit(m.z))) m = v3_neg(m);                                        /* mulsign(m, cos_theta(m)) */
            omf d = mf_make(b);
            float D = mf_eval(&d, m);
            float F = fresnel_r(v3_dot(si->wi, m), b->eta);
            float G = mf_smith_g1(&d, si->wi, m) * mf_smith_g1(&d, wo, m);
            if (reflect) {
                float value = F * D * G / (4.f * fabsf(cti));
                return v3_make(value * b->specular_reflectance[0], value * b->specular_reflectance[1], value * b->specular_reflectance[2]);
            }
            float scale = b2o_sqr(inv_eta);
            float dwm = v3_dot(si->wi, m), dom = v3_dot(wo, m);
            float value = fabsf((scale * (1.f - F) * D * G * eta * eta * dwm * dom) / (cti * b2o_sqr(dwm + eta * dom)));
            return v3_make(value * b->specular_transmittance[0], value * b->specular_transmittance[1], value * b->specular_transmittance[2]);
        }
        case B2_BSDF_PLASTIC: {                                            /* plastic.cpp:242-266 */
            if (!(cti > 0.f && cto > 0.f)) return zero;
            float f_i = fresnel_r(cti, b->eta), f_o = fresnel_r(cto, b->eta);
            v3 diff = plastic_diffuse(b);
            float s = (B2O_INV_PI * cto) * b->inv_eta_2 * (1.f - f_i) * (1.f - f_o);
            return v3_scale(diff, s);
        }
        default: return zero;                                              /* dielectric.cpp:312-315, conductor.cpp:242-245 */
    }
}
static float bsdf_pdf(const B2Bsdf *b, const osi *si, v3 wo) {
    float cti = si->wi.z, cto = wo.z;
    switch (b->type) {
        case B2_BSDF_DIFFUSE: case B2_BSDF_NULL:                           /* diffuse.cpp:121-135 */
            return (cti > 0.f && cto > 0.f) ? B2O_INV_PI * cto : 0.f;
        case B2_BSDF_ROUGHCONDUCTOR: {                                     /* roughconductor.cpp:345-382 */
            v3 m = v3_normalize(v3_add(wo, si->wi));
            if (!(cti > 0.f && cto > 0.f && v3_dot(si->wi, m) > 0.f && v3_dot(wo, m) > 0.f)) return 0.f;
            omf d = mf_make(b);
            return mf_eval(&d, m) * mf_smith_g1(&d, si->wi, m) / (4.f * cti);
        }
        case B2_BSDF_ROUGHDIELECTRIC: {                                    /* roughdielectric.cpp:386-452 */
            if (cti == 0.f) return 0.f;
            int reflect = cti * cto > 0.f;
            float eta = cti > 0.f ? b->eta : b2o_rcp(b->eta);
            v3 m = v3_normalize(v3_add(si->wi, v3_scale(wo, reflect ? 1.f : eta)));
            if (m.z < 0.f || (m.z == 0.f && signbit(m.z))) m = v3_neg(m);
            if (!(v3_dot(si->wi, m) * cti > 0.f && v3_dot(wo, m) * cto > 0.f)) return 0.f;
            float dwh_dwo = reflect ? b2o_rcp(4.f * v3_dot(wo, m))
                                    : (eta * eta * v3_dot(wo, m)) / b2o_sqr(v3_dot(si->wi, m) + eta * v3_dot(wo, m));
            omf d = mf_make(b);
            v3 wi_s = cti < 0.f ? v3_neg(si->wi) : si->wi;                                                        /* mulsign(wi, cos_theta(wi)) */
            float prob = mf_eval(&d, m) * mf_smith_g1(&d, wi_s, m) * fabsf(v3_dot(wi_s, m)) / wi_s.z;            /* distr.pdf, microfacet.h:214-223 */
            float F = fresnel_r(v3_dot(si->wi, m), b->eta);
            prob *= reflect ? F : 1.f - F;
            return prob * fabsf(dwh_dwo);
        }
        case B2_BSDF_PLASTIC: {                                            /* plastic.cpp:268-293 */
            if (!(cti > 0.f && cto > 0.f)) return 0.f;
            float f_i = fresnel_r(cti, b->eta);
            float prob_specular = f_i * b->specular_sampling_weight;
            float prob_diffuse = (1.f - f_i) * (1.f - b->specular_sampling_weight);
            prob_diffuse = prob_diffuse / (prob_specular + prob_diffuse);
            return (B2O_INV_PI * cto) * prob_diffuse;
        }
        default: return 0.f;
    }
}
/* returns the weight f/pdf; sample1 / (s2x,s2y) as in BSDF::sample(ctx, si, sample1, sample2) */
static v3 bsdf_sample(const B2Bsdf *b, const osi *si, float sample1, float s2x, float s2y, obs *bs) {
    v3 zero = v3_make(0.f, 0.f, 0.f);
    memset(bs, 0, sizeof *bs);
    float cti = si->wi.z;
    switch (b->type) {
        case B2_BSDF_DIFFUSE: case B2_BSDF_NULL: {                         /* diffuse.cpp:78-102 */
            if (!(cti > 0.f)) return zero;
            bs->wo = square_to_cosine_hemisphere(s2x, s2y);
            bs->pdf = B2O_INV_PI * bs->wo.z;
            bs->eta = 1.f; bs->delta = 0;
            if (!(bs->pdf > 0.f)) return zero;
            return v3_make(b->reflectance[0], b->reflectance[1], b->reflectance[2]);
        }
        case B2_BSDF_DIELECTRIC: {                                         /* dielectric.cpp:201-310 */
            float r_i, cos_theta_t, eta_it, eta_ti;
            fresnel(cti, b->eta, &r_i, &cos_theta_t, &eta_it, &eta_ti);
            float t_i = 1.f - r_i;
            int selected_r = sample1 <= r_i;
            bs->pdf = selected_r ? r_i : t_i;
            bs->delta = 1;
            if (selected_r) { bs->wo = v3_make(-si->wi.x, -si->wi.y, si->wi.z); bs->eta = 1.f; }
            else { bs->wo = v3_make(-eta_ti * si->wi.x, -eta_ti * si->wi.y, cos_theta_t); bs->eta = eta_it; }
            v3 w = v3_make(1.f, 1.f, 1.f);
            if (selected_r) w = v3_mul(w, v3_make(b->specular_reflectance[0], b->specular_reflectance[1], b->specular_reflectance[2]));
            else {
                w = v3_mul(w, v3_make(b->specular_transmittance[0], b->specular_transmittance[1], b->specular_transmittance[2]));
                w = v3_scale(w, b2o_sqr(eta_ti));                            /* dielectric.cpp:303-307 (Radiance mode) */
            }
            return w;
        }
        case B2_BSDF_ROUGHCONDUCTOR: {                                     /* roughconductor.cpp:196-275 */
            if (!(cti > 0.f)) return zero;
            omf d = mf_make(b);
            float pdf; v3 m = mf_sample(&d, si->wi, s2x, s2y, &pdf);
            /* reflect(wi, m) = fmsub(m, 2*dot(wi,m), wi)  fresnel.h:281-283 */
            float dwm = 2.f * v3_dot(si->wi, m);
            bs->wo = v3_make(b2o_fma(m.x, dwm, -si->wi.x), b2o_fma(m.y, dwm, -si->wi.y), b2o_fma(m.z, dwm, -si->wi.z));
            bs->eta = 1.f; bs->delta = 0;
            int active = pdf != 0.f && bs->wo.z > 0.f;
            float weight = mf_smith_g1(&d, bs->wo, m);
            bs->pdf = pdf / (4.f * v3_dot(bs->wo, m));
            float c = v3_dot(si->wi, m);
            v3 F = v3_make(fresnel_conductor(c, b->cond_eta[0], b->cond_k[0]),
                           fresnel_conductor(c, b->cond_eta[1], b->cond_k[1]),
                           fresnel_conductor(c, b->cond_eta[2], b->cond_k[2]));
            v3 wv = v3_make(weight * b->specular_reflectance[0], weight * b->specular_reflectance[1],
                            weight * b->specular_reflectance[2]);
            return active ? v3_mul(F, wv) : zero;
        }
        case B2_BSDF_ROUGHDIELECTRIC: {                                    /* roughdielectric.cpp:203-310 */
            int active = cti != 0.f;
            omf d = mf_make(b);
            v3 wi_s = cti < 0.f ? v3_neg(si->wi) : si->wi;
            float pdf; v3 m = mf_sample(&d, wi_s, s2x, s2y, &pdf);
            active = active && pdf != 0.f;
            float F, cos_theta_t, eta_it, eta_ti;
            fresnel(v3_dot(si->wi, m), b->eta, &F, &cos_theta_t, &eta_it, &eta_ti);
            int selected_r = sample1 <= F && active;
            v3 weight = v3_make(1.f, 1.f, 1.f);
            pdf *= selected_r ? F : 1.f - F;
            bs->eta = selected_r ? 1.f : eta_it; bs->delta = 0;
            float dwh_dwo = 0.f;
            float dwm = v3_dot(si->wi, m);
            if (selected_r) {
                float s = 2.f * dwm;                                       /* reflect(wi, m) = fmsub(m, 2 dot(wi, m), wi) */
                bs->wo = v3_make(b2o_fma(m.x, s, -si->wi.x), b2o_fma(m.y, s, -si->wi.y), b2o_fma(m.z, s, -si->wi.z));
                weight = v3_mul(weight, v3_make(b->specular_reflectance[0], b->specular_reflectance[1], b->specular_reflectance[2]));
                dwh_dwo = b2o_rcp(4.f * v3_dot(bs->wo, m));
            } else if (active) {
                float s = b2o_fma(dwm, eta_ti, cos_theta_t);               /* refract(wi, m, cos_theta_t, eta_ti), fresnel.h:310-313 */
                bs->wo = v3_make(b2o_fma(m.x, s, -(si->wi.x * eta_ti)), b2o_fma(m.y, s, -(si->wi.y * eta_ti)), b2o_fma(m.z, s, -(si->wi.z * eta_ti)));
                float fct = b2o_sqr(eta_ti);
                weight = v3_mul(weight, v3_make(fct * b->specular_transmittance[0], fct * b->specular_transmittance[1], fct * b->specular_transmittance[2]));
                dwh_dwo = (b2o_sqr(bs->eta) * v3_dot(bs->wo, m)) / b2o_sqr(dwm + bs->eta * v3_dot(bs->wo, m));
            }
            weight = v3_scale(weight, mf_smith_g1(&d, bs->wo, m));
            bs->pdf = pdf * fabsf(dwh_dwo);
            return active ? weight : zero;
        }
        case B2_BSDF_CONDUCTOR: {                                          /* conductor.cpp:197-240 */
            if (!(cti > 0.f)) return zero;
            bs->wo = v3_make(-si->wi.x, -si->wi.y, si->wi.z);              /* reflect(wi), fresnel.h:276-278 */
            bs->eta = 1.f; bs->pdf = 1.f; bs->delta = 1;
            v3 F = v3_make(fresnel_conductor(cti, b->cond_eta[0], b->cond_k[0]),
                           fresnel_conductor(cti, b->cond_eta[1], b->cond_k[1]),
                           fresnel_conductor(cti, b->cond_eta[2], b->cond_k[2]));
            return v3_mul(v3_make(b->specular_reflectance[0], b->specular_reflectance[1], b->specular_reflectance[2]), F);
        }
        case B2_BSDF_PLASTIC: {                                            /* plastic.cpp:181-240 */
            if (!(cti > 0.f)) return zero;
            float f_i = fresnel_r(cti, b->eta);
            float prob_specular = f_i * b->specular_sampling_weight;
            float prob_diffuse = (1.f - f_i) * (1.f - b->specular_sampling_weight);
            prob_specular = prob_specular / (prob_specular + prob_diffuse);
            prob_diffuse = 1.f - prob_specular;
            bs->eta = 1.f;
            if (sample1 < prob_specular) {
                bs->wo = v3_make(-si->wi.x, -si->wi.y, si->wi.z);
                bs->pdf = prob_specular; bs->delta = 1;
                float value = f_i / bs->pdf;
                return v3_make(value * b->specular_reflectance[0], value * b->specular_reflectance[1], value * b->specular_reflectance[2]);
            }
            bs->wo = square_to_cosine_hemisphere(s2x, s2y);
            bs->pdf = prob_diffuse * (B2O_INV_PI * bs->wo.z); bs->delta = 0;
            float f_o = fresnel_r(bs->wo.z, b->eta);
            v3 value = plastic_diffuse(b);
            float s = b->inv_eta_2 * (1.f - f_i) * (1.f - f_o) / prob_diffuse;
            return v3_scale(value, s);
        }
        default: return zero;
    }
}

/* ============================================================================================
 * Emitters — src/emitters/area.cpp, src/librender/shape.cpp:292-323, mesh.cpp:352-397,
 *            scene.cpp:164-231, distr_1d.h:144-204
 * ============================================================================================ */
typedef struct { v3 p, n, d; float pdf, dist; int emitter; int delta; } ods;

static uint32_t distr_sample(const float *cdf, uint32_t n, float sum, float value) {   /* distr_1d.h:144-155 */
    value *= sum;
    /* m_valid: first/last index with nonzero pmf; binary_search array_utils.h:153-170 */
    uint32_t first = 0, last = n - 1;
    while (first < n - 1 && !(cdf[first] > 0.f)) ++first;
    while (last > 0 && cdf[last] == cdf[last - 1]) --last;
    uint32_t start = first, end = last;
    uint32_t iterations = 0;
    if (start < end) { uint32_t x = end - start; while (x) { ++iterations; x >>= 1; } } /* log2i(x)+1 */
    for (uint32_t i = 0; i < iterations; ++i) {
        uint32_t middle = (start + end) >> 1;
        int cond = cdf[middle] < value;
        if (cond) start = (middle + 1 < end) ? middle + 1 : end; else end = middle;
    }
    return start;
}

/* Mesh::sample_position mesh.cpp:352-397 + Shape::sample_direction shape.cpp:292-310 +
   AreaLight::sample_direction area.cpp:121-134 ; returns spec = radiance / pdf (0 if inactive) */
/* Rectangle: Shape::sample_direction (shape.cpp:292-323) over Rectangle::sample_position (rectangle.cpp:113-128);
 * Sphere::sample_direction (sphere.cpp:174-253); then AreaLight::sample_direction (area.cpp:118-141) */
static inline float rect_inv_area(const B2Shape *sh) {        /* rectangle.cpp:82-92, 103-105 */
    v3 fs = xf_vector(sh->to_world, v3_make(2.f, 0.f, 0.f)), ft = xf_vector(sh->to_world, v3_make(0.f, 2.f, 0.f));
    v3 c = v3_cross(fs, ft);
    return b2o_rcp(sqrtf(v3_dot(c, c)));
}
static inline float sphere_inv_area(const B2Shape *sh) { return b2o_rcp(4.f * B2O_PI * sh->radius * sh->radius); }   /* sphere.cpp:138-140 */
static inline v3 square_to_uniform_sphere(float sx, float sy) {      /* warp.h:255-260 */
    float z = b2o_fma(-2.f, sy, 1.f), r = b2o_safe_sqrt(b2o_fma(-z, z, 1.f));
    float sn, cs; b2o_sincos(2.f * B2O_PI * sx, &sn, &cs);
    return v3_make(r * cs, r * sn, z);
}
static v3 shape_sample_direction(const oscene *s, int emitter, v3 ref_p, float sx, float sy, ods *ds) {
    const B2Emitter *em = &s->desc->emitters[emitter];
    const B2Shape *sh = &s->desc->shapes[em->shape - 1];
    v3 zero = v3_make(0.f, 0.f, 0.f);
    ds->emitter = emitter;
    if (sh->type == B2_SHAPE_RECTANGLE) {
        ds->p = xf_point(sh->to_world, v3_make(sx * 2.f - 1.f, sy * 2.f - 1.f, 0.f));
        ds->n = v3_normalize(xf_normal(sh->to_object, v3_make(0.f, 0.f, 1.f)));
        float pdf = rect_inv_area(sh);
        ds->d = v3_sub(ds->p, ref_p);
        float dist_squared = v3_dot(ds->d, ds->d);
        ds->dist = sqrtf(dist_squared);
        ds->d = v3_scale(ds->d, b2o_rcp(ds->dist));
        float dp = fabsf(v3_dot(ds->d, ds->n));
        pdf *= dp != 0.f ? dist_squared / dp : 0.f;
        ds->pdf = pdf;
    } else {
        v3 c = v3_make(sh->center[0], sh->center[1], sh->center[2]);
        v3 dc_v = v3_sub(c, ref_p);
        float dc_2 = v3_dot(dc_v, dc_v);
        float radius_adj = sh->radius * (sh->flip_normals ? (1.f + B2O_RAY_EPS) : (1.f - B2O_RAY_EPS));
        v3 d;
        if (dc_2 > b2o_sqr(radius_adj)) {                               /* reference point outside: cone sampling */
            float inv_dc = b2o_rsqrt(dc_2), sin_theta_max = sh->radius * inv_dc, sin_theta_max_2 = b2o_sqr(sin_theta_max),
                  inv_sin_theta_max = b2o_rcp(sin_theta_max), cos_theta_max = b2o_safe_sqrt(1.f - sin_theta_max_2);
            float sin_theta_2 = sin_theta_max_2 > 0.00068523f ? 1.f - b2o_sqr(b2o_fma(cos_theta_max - 1.f, sx, 1.f)) : sin_theta_max_2 * sx,
                  cos_theta = b2o_safe_sqrt(1.f - sin_theta_2);
            float cos_alpha = sin_theta_2 * inv_sin_theta_max + cos_theta * b2o_safe_sqrt(b2o_fma(-sin_theta_2, b2o_sqr(inv_sin_theta_max), 1.f)),
                  sin_alpha = b2o_safe_sqrt(b2o_fma(-cos_alpha, cos_alpha, 1.f));
            float sin_phi, cos_phi; b2o_sincos(sy * (2.f * B2O_PI), &sin_phi, &cos_phi);
            v3 fn = v3_scale(dc_v, -inv_dc), fs, ft;                      /* Frame3f(n): coordinate_system(n) */
            coordinate_system(fn, &fs, &ft);
            v3 l = v3_make(cos_phi * sin_alpha, sin_phi * sin_alpha, cos_alpha);
            d = v3_add(v3_add(v3_scale(fs, l.x), v3_scale(ft, l.y)), v3_scale(fn, l.z));
            ds->p = v3_make(b2o_fma(d.x, sh->radius, c.x), b2o_fma(d.y, sh->radius, c.y), b2o_fma(d.z, sh->radius, c.z));
            ds->n = d;
            ds->d = v3_sub(ds->p, ref_p);
            float dist2 = v3_dot(ds->d, ds->d);
            ds->dist = sqrtf(dist2);
            ds->d = v3_scale(ds->d, b2o_rcp(ds->dist));
            ds->pdf = (.5f * B2O_INV_PI) / (1.f - cos_theta_max);      /* square_to_uniform_cone_pdf, warp.h:481-490 */
            if (ds->dist == 0.f) ds->pdf = 0.f;
        } else {                                                        /* inside: uniform area sampling */
            d = square_to_uniform_sphere(sx, sy);
            ds->p = v3_make(b2o_fma(d.x, sh->radius, c.x), b2o_fma(d.y, sh->radius, c.y), b2o_fma(d.z, sh->radius, c.z));
            ds->n = d;
            ds->d = v3_sub(ds->p, ref_p);
            float dist2 = v3_dot(ds->d, ds->d);
            ds->dist = sqrtf(dist2);
            ds->d = v3_scale(ds->d, b2o_rcp(ds->dist));
            ds->pdf = sphere_inv_area(sh) * dist2 / fabsf(v3_dot(ds->d, ds->n));
        }
        if (sh->flip_normals) ds->n = v3_neg(ds->n);
    }
    int active = v3_dot(ds->d, ds->n) < 0.f && ds->pdf != 0.f;            /* area.cpp:131-134 */
    if (!active) return zero;
    float ip = b2o_rcp(ds->pdf);
    return v3_make(em->radiance[0] * ip, em->radiance[1] * ip, em->radiance[2] * ip);
}

static v3 area_sample_direction(const oscene *s, int emitter, v3 ref_p, float sx, float sy, ods *ds) {
    const B2Emitter *em = &s->desc->emitters[emitter];
    if (em->shape > 0) return shape_sample_direction(s, emitter, ref_p, sx, sy, ds);
    int m = em->mesh;
    const B2Mesh *me = &s->desc->meshes[m];
    const float *cdf = s->cdf[m];
    uint32_t face = distr_sample(cdf, me->n_faces, s->cdf_sum[m], sy);
    /* sample_reuse distr_1d.h:194-204: pmf(index) = pmf[index]*norm ; pmf[i] is the stored face area */
    const otri *tr = &s->tris[s->mesh_first[m] + face];
    v3 e0 = v3_sub(tr->p1, tr->p0), e1 = v3_sub(tr->p2, tr->p0);
    v3 cr = v3_cross(e0, e1);
    float area = 0.5f * sqrtf(v3_dot(cr, cr));
    float pmf = area * s->cdf_norm[m];
    float cdfv = face > 0 ? cdf[face - 1] * s->cdf_norm[m] : 0.f;
    sy = (sy - cdfv) / pmf;
    float t = b2o_safe_sqrt(1.f - sx);                      /* warp.h:152-156 */
    float bx = 1.f - t, by = t * sy;
    ds->p = v3_add(v3_add(tr->p0, v3_scale(e0, bx)), v3_scale(e1, by));
    float pdf = s->cdf_norm[m];
    if (tr->has_normals)
        ds->n = v3_normalize(v3_add(v3_add(v3_scale(tr->n0, 1.f - bx - by), v3_scale(tr->n1, bx)), v3_scale(tr->n2, by)));
    else
        ds->n = v3_normalize(cr);
    ds->d = v3_sub(ds->p, ref_p);
    float dist_squared = v3_dot(ds->d, ds->d);
    ds->dist = sqrtf(dist_squared);
    ds->d = v3_scale(ds->d, b2o_rcp(ds->dist));
    float dp = fabsf(v3_dot(ds->d, ds->n));
    pdf *= dp != 0.f ? dist_squared / dp : 0.f;
    ds->pdf = pdf; ds->emitter = emitter;
    int active = v3_dot(ds->d, ds->n) < 0.f && pdf != 0.f;
    if (!active) return v3_make(0.f, 0.f, 0.f);
    float ip = b2o_rcp(pdf);
    return v3_make(em->radiance[0] * ip, em->radiance[1] * ip, em->radiance[2] * ip);
}

static v3 env_sample_direction(const oscene *s, int emitter, v3 ref_p, float sx, float sy, ods *ds);
static float env_pdf_direction(const oscene *s, v3 d_world);
static v3 emitter_sample_direction(const oscene *s, int emitter, v3 ref_p, float sx, float sy, ods *ds) {
    ds->delta = 0;
    if (s->desc->emitters[emitter].type == B2_EMITTER_SPOT) {                  /* SpotLight::sample_direction + falloff_curve, spot.cpp:97-160 */
        const B2Emitter *em = &s->desc->emitters[emitter];
        ds->p = v3_make(em->env_to_world[3], em->env_to_world[7], em->env_to_world[11]);
        ds->n = v3_make(0.f, 0.f, 0.f);
        ds->pdf = 1.f; ds->delta = 1; ds->emitter = emitter;
        ds->d = v3_sub(ds->p, ref_p);
        ds->dist = sqrtf(v3_dot(ds->d, ds->d));
        float inv_dist = b2o_rcp(ds->dist);
        ds->d = v3_scale(ds->d, inv_dist);
        v3 local_dir = v3_normalize(xf_vector(em->env_to_object, v3_neg(ds->d)));
        float cos_theta = local_dir.z;
        float cos_cutoff = cosf(em->cutoff_angle), cos_beam = cosf(em->beam_width);         /* host values, spot.cpp:93-94 */
        float inv_transition = 1.0f / (em->cutoff_angle - em->beam_width);
        float f = cos_theta >= cos_beam ? 1.f : (em->cutoff_angle - b2o_safe_acos(cos_theta)) * inv_transition;
        if (cos_theta <= cos_cutoff) f = 0.f;
        float w = inv_dist * inv_dist;
        v3 res = v3_make(em->radiance[0], em->radiance[1], em->radiance[2]);
        if (!(cos_theta >= cos_beam)) res = v3_scale(res, f);
        if (cos_theta <= cos_cutoff) res = v3_make(0.f, 0.f, 0.f);
        return v3_scale(res, w);
    }
    if (s->desc->emitters[emitter].type == B2_EMITTER_DIRECTIONAL) {           /* DirectionalEmitter::sample_direction, directional.cpp:107-134 */
        const B2Emitter *em = &s->desc->emitters[emitter];
        v3 d = xf_vector(em->env_to_world, v3_make(0.f, 0.f, 1.f));
        float dist = 2.f * s->env_radius;
        ds->p = v3_sub(ref_p, v3_scale(d, dist));
        ds->n = d; ds->pdf = 1.f; ds->delta = 1; ds->emitter = emitter;
        ds->d = v3_neg(d); ds->dist = dist;
        return v3_make(em->radiance[0], em->radiance[1], em->radiance[2]);
    }
    if (s->desc->emitters[emitter].type == B2_EMITTER_POINT) {                 /* PointLight::sample_direction, point.cpp:86-113 */
        const B2Emitter *em = &s->desc->emitters[emitter];
        ds->p = v3_make(em->env_to_world[3], em->env_to_world[7], em->env_to_world[11]);
        ds->n = v3_make(0.f, 0.f, 0.f);
        ds->pdf = 1.f; ds->delta = 1; ds->emitter = emitter;
        ds->d = v3_sub(ds->p, ref_p);
        ds->dist = sqrtf(v3_dot(ds->d, ds->d));
        float inv_dist = b2o_rcp(ds->dist);
        ds->d = v3_scale(ds->d, inv_dist);
        float w = b2o_sqr(inv_dist);
        return v3_make(em->radiance[0] * w, em->radiance[1] * w, em->radiance[2] * w);
    }
    return (s->desc->emitters[emitter].type == B2_EMITTER_ENVMAP || s->desc->emitters[emitter].type == B2_EMITTER_CONSTANT) ? env_sample_direction(s, emitter, ref_p, sx, sy, ds)
                                                                 : area_sample_direction(s, emitter, ref_p, sx, sy, ds);
}
/* Scene::sample_emitter_direction scene.cpp:164-214 (visibility test done by the caller) */
static v3 scene_sample_emitter_direction(const oscene *s, v3 ref_p, float sx, float sy, ods *ds) {
    uint32_t ne = s->desc->n_emitters;
    if (ne == 0) { memset(ds, 0, sizeof *ds); return v3_make(0.f, 0.f, 0.f); }
    if (ne == 1) return emitter_sample_direction(s, 0, ref_p, sx, sy, ds);
    float emitter_pdf = 1.f / (float) ne;
    uint32_t index = (uint32_t) (sx * (float) ne);
    if (index > ne - 1) index = ne - 1;
    sx = (sx - (float) index * emitter_pdf) * (float) ne;
    v3 spec = emitter_sample_direction(s, (int) index, ref_p, sx, sy, ds);
    ds->pdf *= emitter_pdf;
    return v3_scale(spec, b2o_rcp(emitter_pdf));
}

/* Scene::pdf_emitter_direction scene.cpp:216-231 -> AreaLight::pdf_direction area.cpp:168-187
   -> Shape::pdf_direction shape.cpp:312-323 -> Mesh::pdf_position mesh.cpp:417-420 */
static float scene_pdf_emitter_direction(const oscene *s, int emitter, v3 ref_p, v3 d, v3 n, float dist) {
    const B2Emitter *em = &s->desc->emitters[emitter];
    if (em->type == B2_EMITTER_ENVMAP || em->type == B2_EMITTER_CONSTANT) {
        float pdf = env_pdf_direction(s, d);
        if (s->desc->n_emitters != 1) pdf = pdf * (1.f / (float) s->desc->n_emitters);
        return pdf;
    }
    float dp = v3_dot(d, n);
    if (!(dp < 0.f)) return 0.f;
    float pdf, adp = fabsf(dp);
    if (em->shape > 0 && s->desc->shapes[em->shape - 1].type == B2_SHAPE_SPHERE) {     /* Sphere::pdf_direction sphere.cpp:255-267 */
        const B2Shape *sh = &s->desc->shapes[em->shape - 1];
        v3 dc = v3_sub(v3_make(sh->center[0], sh->center[1], sh->center[2]), ref_p);
        float sin_alpha = sh->radius * b2o_rcp(sqrtf(v3_dot(dc, dc))), cos_alpha = b2o_safe_sqrt(1.f - sin_alpha * sin_alpha);
        pdf = sin_alpha < B2O_ONE_MINUS_EPS ? (.5f * B2O_INV_PI) / (1.f - cos_alpha) : sphere_inv_area(sh) * b2o_sqr(dist) / adp;
    } else {
        pdf = em->shape > 0 ? rect_inv_area(&s->desc->shapes[em->shape - 1]) : s->cdf_norm[em->mesh];
        pdf *= adp != 0.f ? (dist * dist) / adp : 0.f;                                   /* shape.cpp:325-336 */
    }
    if (s->desc->n_emitters != 1) pdf = pdf * (1.f / (float) s->desc->n_emitters);
    return pdf;
}

/* ============================================================================================
 * Environment map — src/emitters/envmap.cpp, include/mitsuba/core/distr_2d.h:336-700 (Hierarchical2D),
 * include/mitsuba/core/warp.h:359-434
 * ============================================================================================ */
static inline uint32_t lvl_index(uint32_t x, uint32_t y, uint32_t width) {     /* distr_2d.h Level::index */
    return ((x & 1u) | (((x & ~1u) | (y & 1u)) << 1)) + ((y & ~1u) * width);
}
static inline float b2o_lerp(float a, float b, float t) { return b2o_fma(b, t, b2o_fma(-a, t, a)); }  /* enoki::lerp */
static inline float interval_to_linear(float v0, float v1, float sample) {       /* warp.h:359-366 */
    if (fabsf(v0 - v1) > 1e-4f * (v0 + v1))
        return (v0 - b2o_safe_sqrt(b2o_lerp(b2o_sqr(v0), b2o_sqr(v1), sample))) / (v0 - v1);
    return sample;
}
static inline float clamp01(float x) { return b2o_min(b2o_max(x, 0.f), 1.f); }

/* Hierarchical2D constructor (distr_2d.h:372-471), normalize = true, enable_sampling = true */
static void env_build(oscene *s, const B2Emitter *em) {
    uint32_t W = (uint32_t) em->env_width, H = (uint32_t) em->env_height;
    float *lum = (float *) malloc(sizeof(float) * W * H);
    for (uint32_t y = 0; y < H; ++y) {                                /* envmap.cpp:86-117 */
        float sin_theta = sinf((float) y / (float) (H - 1) * B2O_PI);
        for (uint32_t x = 0; x < W; ++x) {
            const float *p = em->envmap + 4 * ((size_t) y * W + x);
            float l = p[0] * 0.212671f + p[1] * 0.715160f + p[2] * 0.072169f;     /* spectrum.h:246-248 */
            lum[y * W + x] = l * sin_theta;
        }
    }
    uint32_t npx = W - 1, npy = H - 1;
    uint32_t mx = npx > npy ? npx : npy, max_level = 0;
    while ((1u << max_level) < mx) ++max_level;                       /* log2i_ceil */
    int nl = (int) max_level + 2;
    s->env_nlevels = nl;
    s->env_level = (float **) calloc((size_t) nl, sizeof(float *));
    s->env_level_w = (uint32_t *) calloc((size_t) nl, sizeof(uint32_t));
    uint32_t *lh = (uint32_t *) calloc((size_t) nl, sizeof(uint32_t));
    s->env_level[0] = (float *) calloc((size_t) W * H, sizeof(float)); s->env_level_w[0] = W; lh[0] = H;
    uint32_t lx = npx, ly = npy;
    for (int level = (int) max_level, k = 1; level >= 0; --level, ++k) {
        lx += lx & 1u; ly += ly & 1u;                                 /* zero-pad */
        s->env_level[k] = (float *) calloc((size_t) lx * ly, sizeof(float)); s->env_level_w[k] = lx; lh[k] = ly;
        lx >>= 1; ly >>= 1;
    }
    double sum = 0.0;
    const float *in = lum;
    for (uint32_t y = 0; y < npy; ++y) {
        for (uint32_t x = 0; x < npx; ++x) {
            float avg = (in[0] + in[1] + in[W] + in[W + 1]) * .25f;
            sum += (double) avg;
            s->env_level[1][lvl_index(x, y, s->env_level_w[1])] = avg;
            ++in;
        }
        ++in;
    }
    float scale = (float) ((double) (npx * npy) / sum);
    for (uint32_t i = 0; i < W * H; ++i) s->env_level[0][i] = lum[i] * scale;
    for (uint32_t i = 0; i < s->env_level_w[1] * lh[1]; ++i) s->env_level[1][i] *= scale;
    lx = npx; ly = npy;
    for (int level = 2; level <= (int) max_level + 1; ++level) {
        lx = (lx + 1u) >> 1; ly = (ly + 1u) >> 1;
        for (uint32_t y = 0; y < ly; ++y)
            for (uint32_t x = 0; x < lx; ++x) {
                const float *d0 = s->env_level[level - 1] + lvl_index(x * 2, y * 2, s->env_level_w[level - 1]);
                s->env_level[level][lvl_index(x, y, s->env_level_w[level])] = d0[0] + d0[1] + d0[2] + d0[3];
            }
    }
    free(lum); free(lh);
}

/* Hierarchical2D::sample (distr_2d.h:473-560) */
static void env_warp_sample(const oscene *s, const B2Emitter *em, float sx, float sy, float *u, float *v, float *pdf) {
    sx = clamp01(sx); sy = clamp01(sy);
    uint32_t ox = 0, oy = 0;
    for (int l = s->env_nlevels - 2; l > 0; --l) {
        const float *lv = s->env_level[l];
        ox <<= 1; oy <<= 1;
        uint32_t oi = lvl_index(ox, oy, s->env_level_w[l]);
        float v00 = lv[oi], v10 = lv[oi + 1], v01 = lv[oi + 2], v11 = lv[oi + 3];
        sx = clamp01(sx); sy = clamp01(sy);
        float r0 = v00 + v10, r1 = v01 + v11;
        sy *= r0 + r1;
        int mask = sy > r0;
        if (mask) { oy += 1u; sy -= r0; }
        sy /= mask ? r1 : r0;
        float c0 = mask ? v01 : v00, c1 = mask ? v11 : v10;
        sx *= c0 + c1;
        mask = sx > c0;
        if (mask) sx -= c0;
        sx /= mask ? c1 : c0;
        if (mask) ox += 1u;
    }
    uint32_t W = (uint32_t) em->env_width;
    const float *l0 = s->env_level[0];
    uint32_t oi = ox + oy * W;
    float v00 = l0[oi], v10 = l0[oi + 1], v01 = l0[oi + W], v11 = l0[oi + W + 1];
    /* warp::square_to_bilinear (warp.h:393-407) */
    float r0 = v00 + v10, r1 = v01 + v11;
    sy = interval_to_linear(r0, r1, sy);
    float c0 = b2o_lerp(v00, v01, sy), c1 = b2o_lerp(v10, v11, sy);
    sx = interval_to_linear(c0, c1, sx);
    *pdf = b2o_lerp(c0, c1, sx);
    float psx = 1.f / (float) (em->env_width - 1), psy = 1.f / (float) (em->env_height - 1);   /* m_patch_size */
    *u = ((float) (int) ox + sx) * psx; *v = ((float) (int) oy + sy) * psy;
}
/* Hierarchical2D::eval (distr_2d.h:651-686) */
static float env_warp_eval(const oscene *s, const B2Emitter *em, float px, float py) {
    px = clamp01(px); py = clamp01(py);
    uint32_t W = (uint32_t) em->env_width, H = (uint32_t) em->env_height;
    px *= (float) (W - 1); py *= (float) (H - 1);                     /* m_inv_patch_size */
    uint32_t ox = (uint32_t) (int) px, oy = (uint32_t) (int) py;
    if (ox > W - 2) ox = W - 2;
    if (oy > H - 2) oy = H - 2;
    px -= (float) (int) ox; py -= (float) (int) oy;
    const float *l0 = s->env_level[0];
    uint32_t oi = ox + oy * W;
    float v00 = l0[oi], v10 = l0[oi + 1], v01 = l0[oi + W], v11 = l0[oi + W + 1];
    return b2o_lerp(b2o_lerp(v00, v10, px), b2o_lerp(v01, v11, px), py);     /* warp.h:428-434 */
}
/* EnvironmentMapEmitter::eval_spectrum (envmap.cpp:269-319), rgb branch */
static v3 env_eval_spectrum(const B2Emitter *em, float u, float v) {
    uint32_t W = (uint32_t) em->env_width, H = (uint32_t) em->env_height;
    u *= (float) (W - 1u); v *= (float) (H - 1u);
    uint32_t px = (uint32_t) u, py = (uint32_t) v;
    if (px > W - 2u) px = W - 2u;
    if (py > H - 2u) py = H - 2u;
    float w1x = u - (float) px, w1y = v - (float) py, w0x = 1.f - w1x, w0y = 1.f - w1y;
    const float *d = em->envmap;
    size_t idx = (size_t) px + (size_t) py * W;
    const float *v00 = d + 4 * idx, *v10 = d + 4 * (idx + 1), *v01 = d + 4 * (idx + W), *v11 = d + 4 * (idx + W + 1);
    float out[3];
    for (int k = 0; k < 3; ++k) {
        float a0 = b2o_fma(w0x, v00[k], w1x * v10[k]), a1 = b2o_fma(w0x, v01[k], w1x * v11[k]);
        out[k] = b2o_fma(w0y, a0, w1y * a1) * em->scale;
    }
    return v3_make(out[0], out[1], out[2]);
}
/* direction (world) -> uv (envmap.cpp:137-144, 198-203) */
static inline void env_dir_to_uv(const B2Emitter *em, v3 d_world, v3 *d_local, float *u, float *v) {
    v3 d = xf_vector(em->env_to_object, d_world);
    *d_local = d;
    float uu = b2o_atan2(d.x, -d.z) * (0.5f * B2O_INV_PI), vv = b2o_safe_acos(d.y) * B2O_INV_PI;
    uu -= floorf(uu); vv -= floorf(vv);
    *u = uu; *v = vv;
}
/* EnvironmentMapEmitter::eval (envmap.cpp:134-147): radiance arriving along ray direction d */
static v3 env_eval(const oscene *s, v3 ray_d);
static v3 env_eval(const oscene *s, v3 ray_d) {
    const B2Emitter *em = &s->desc->emitters[s->env];
    if (em->type == B2_EMITTER_CONSTANT) return v3_make(em->radiance[0], em->radiance[1], em->radiance[2]);   /* constant.cpp:54-57 */
    v3 dl; float u, v; env_dir_to_uv(em, ray_d, &dl, &u, &v);
    return env_eval_spectrum(em, u, v);
}
static inline float env_inv_sin_theta(v3 d) {      /* safe_rsqrt(max(d.x^2 + d.z^2, Epsilon^2)) */
    float x = b2o_max(b2o_sqr(d.x) + b2o_sqr(d.z), b2o_sqr(5.9604644775390625e-08f));
    return b2o_rsqrt(x);
}
/* EnvironmentMapEmitter::pdf_direction (envmap.cpp:192-208) */
static float env_pdf_direction(const oscene *s, v3 d_world) {
    const B2Emitter *em = &s->desc->emitters[s->env];
    if (em->type == B2_EMITTER_CONSTANT) return .25f * B2O_INV_PI;             /* square_to_uniform_sphere_pdf, constant.cpp:110-114 */
    v3 d; float u, v; env_dir_to_uv(em, d_world, &d, &u, &v);
    return env_warp_eval(s, em, u, v) * env_inv_sin_theta(d) * (1.f / (2.f * b2o_sqr(B2O_PI)));
}
/* EnvironmentMapEmitter::sample_direction (envmap.cpp:156-190) */
static v3 env_sample_direction(const oscene *s, int emitter, v3 ref_p, float sx, float sy, ods *ds) {
    const B2Emitter *em = &s->desc->emitters[emitter];
    if (em->type == B2_EMITTER_CONSTANT) {                                      /* ConstantBackgroundEmitter::sample_direction, constant.cpp:83-108 */
        v3 dd = square_to_uniform_sphere(sx, sy);
        float dist0 = 2.f * s->env_radius;
        ds->p = v3_add(ref_p, v3_scale(dd, dist0));
        ds->n = v3_neg(dd);
        ds->pdf = .25f * B2O_INV_PI;
        ds->d = dd; ds->dist = dist0; ds->emitter = emitter;
        float ip = b2o_rcp(ds->pdf);
        return v3_make(em->radiance[0] * ip, em->radiance[1] * ip, em->radiance[2] * ip);
    }
    float u, v, pdf; env_warp_sample(s, em, sx, sy, &u, &v, &pdf);
    float theta = v * B2O_PI, phi = u * (2.f * B2O_PI);
    float st, ct, sp, cp; b2o_sincos(theta, &st, &ct); b2o_sincos(phi, &sp, &cp);
    v3 d0 = v3_make(cp * st, sp * st, ct);                            /* math::sphdir */
    v3 d = v3_make(d0.y, d0.z, -d0.x);
    float dist = 2.f * s->env_radius;
    float inv_sin_theta = env_inv_sin_theta(d);
    d = xf_vector(em->env_to_world, d);
    ds->p = v3_add(ref_p, v3_scale(d, dist));
    ds->n = v3_neg(d);
    ds->pdf = pdf > 0.f ? pdf * inv_sin_theta * (1.f / (2.f * b2o_sqr(B2O_PI))) : 0.f;
    ds->d = d; ds->dist = dist; ds->emitter = emitter;
    v3 spec = env_eval_spectrum(em, u, v);
    return v3_scale(spec, b2o_rcp(ds->pdf));
}

/* ============================================================================================
 * PathIntegrator::sample — src/integrators/path.cpp:100-227
 * ============================================================================================ */
typedef struct { uint64_t segments, shadow_rays; } ocount;

/* ============================================================================================
 * Samplers — src/librender/sampler.cpp:22-71, src/samplers/independent.cpp:73-82, src/samplers/stratified.cpp:60-131,
 * include/mitsuba/core/random.h:75-85 (sample_tea_32), :229-274 (permute_kensler)
 * ============================================================================================ */
/* (x + (sy + jx) * inv_res.y) * inv_res.x, multijitter.cpp:142-143: a scalar expression the reference's compiler
 * (-ffp-contract=fast) fuses into one fmadd */
#define B2O_MJ_FMA(a, b, c) b2o_fma((a), (b), (c))
typedef struct {
    pcg32 rng;
    uint32_t type, jitter;
    uint32_t sample_index, dim, perm_seed;      /* m_sample_index, m_dimension_index, m_permutation_seed */
    uint32_t count, res; float inv_count, inv_res;
    uint32_t res_y; float inv_res_y;             /* multijitter: res = m_resolution.x, res_y = m_resolution.y */
} osampler;

static uint32_t sample_tea_32(uint32_t v0, uint32_t v1, int rounds) {
    uint32_t sum = 0;
    for (int i = 0; i < rounds; ++i) {
        sum += 0x9e3779b9u;
        v0 += ((v1 << 4) + 0xa341316cu) ^ (v1 + sum) ^ ((v1 >> 5) + 0xc8013ea4u);
        v1 += ((v0 << 4) + 0xad90777du) ^ (v0 + sum) ^ ((v0 >> 5) + 0x7e95761eu);
    }
    return v1;
}
static uint32_t permute_kensler(uint32_t index, uint32_t sample_count, uint32_t seed) {
    if (sample_count == 1) return 0;
    uint32_t w = sample_count - 1;
    w |= w >> 1; w |= w >> 2; w |= w >> 4; w |= w >> 8; w |= w >> 16;
    do {
        uint32_t tmp = index;
        tmp ^= seed; tmp *= 0xe170893du; tmp ^= seed >> 16;
        tmp ^= (tmp & w) >> 4; tmp ^= seed >> 8; tmp *= 0x0929eb3fu; tmp ^= seed >> 23;
        tmp ^= (tmp & w) >> 1; tmp *= 1u | seed >> 27; tmp *= 0x6935fa69u;
        tmp ^= (tmp & w) >> 11; tmp *= 0x74dcb303u;
        tmp ^= (tmp & w) >> 2; tmp *= 0x9e501cc3u;
        tmp ^= (tmp & w) >> 2; tmp *= 0xc860a3dfu;
        tmp &= w; tmp ^= tmp >> 5;
        index = tmp;
    } while (index >= sample_count);
    return (index + seed) % sample_count;
}
/* LowDiscrepancySampler (ldsampler.cpp:96-131): permute (random.h:194-207), radical_inverse_2 / sobol_2 (qmc.h:189-227) */
static uint32_t permute_tea(uint32_t index, uint32_t sample_count, uint32_t seed) {
    uint32_t n = 0; while ((1u << (n + 1)) <= sample_count) ++n;            /* log2i */
    for (uint32_t level = 0; level < n; ++level) {
        uint32_t bit = 1u << level;
        uint32_t rand = sample_tea_32(index | bit, seed, 2);
        if ((rand & bit) == bit) index ^= bit;
    }
    return index;
}
static float radical_inverse_2(uint32_t index, uint32_t scramble) {
    index = (index << 16) | (index >> 16);
    index = ((index & 0x00ff00ffu) << 8) | ((index & 0xff00ff00u) >> 8);
    index = ((index & 0x0f0f0f0fu) << 4) | ((index & 0xf0f0f0f0u) >> 4);
    index = ((index & 0x33333333u) << 2) | ((index & 0xccccccccu) >> 2);
    index = ((index & 0x55555555u) << 1) | ((index & 0xaaaaaaaau) >> 1);
    return b2o_u2f(((index ^ scramble) >> 9) | 0x3f800000u) - 1.f;
}
static float sobol_2(uint32_t index, uint32_t scramble) {
    for (uint32_t v = 1u << 31; index != 0; index >>= 1, v ^= v >> 1)
        if (index & 1u) scramble ^= v;
    return (float) scramble / 4294967296.f;
}
/* Sampler::seed + PCG32Sampler::seed (+ StratifiedSampler::seed); the constructor's derived values (stratified.cpp:60-75) */
static void smp_seed(osampler *s, const B2RenderParams *rp, uint64_t seed_offset) {
    sampler_seed(&s->rng, rp->seed, seed_offset);
    s->type = rp->sampler & 0xffu; s->jitter = !(rp->sampler & B2_SAMPLER_NO_JITTER);
    s->sample_index = 0; s->dim = 0;
    if (s->type == B2_SAMPLER_ORTHOGONAL) {                        /* orthogonal.cpp:76-90: the smallest prime p with p^2 >= spp */
        uint32_t r = 2;
        for (;; ++r) {
            int prime = 1;
            for (uint32_t i = 2; i <= r / 2; ++i) if (r % i == 0) { prime = 0; break; }
            if (prime && r * r >= rp->spp) break;
        }
        s->res = s->res_y = r; s->count = r * r;
    } else if (s->type == B2_SAMPLER_MULTIJITTER) {                /* multijitter.cpp:68-81 */
        s->res_y = (uint32_t) sqrtf((float) rp->spp);
        s->res = (rp->spp + s->res_y - 1) / s->res_y;
        s->count = s->res * s->res_y;
    } else {
        s->res = 1; while (s->res * s->res < rp->spp) s->res++;
        s->res_y = s->res;
        s->count = s->res * s->res;
    }
    s->inv_res_y = b2o_rcp((float) s->res_y);
    s->inv_count = b2o_rcp((float) s->count); s->inv_res = b2o_rcp((float) s->res);
    s->perm_seed = sample_tea_32((uint32_t) rp->seed, (uint32_t) seed_offset, 4);     /* compute_per_sequence_seed, sampler.cpp:55-60 */
}
static inline void smp_advance(osampler *s) { s->dim = 0; s->sample_index++; }
/* OrthogonalSampler::bose (orthogonal.cpp:188-218), strength 2: s->res = the prime resolution, s->count = res^2 */
static float smp_bose(osampler *s) {
    uint32_t j = s->dim++, p = s->perm_seed, res = s->res;
    uint32_t i = permute_kensler(s->sample_index % s->count, s->count, p);
    uint32_t a_i0 = i / res, a_i1 = i - a_i0 * res, a_ij, a_ik;
    if (j == 0) { a_ij = a_i0; a_ik = a_i1; }
    else if (j == 1) { a_ij = a_i1; a_ik = a_i0; }
    else {
        uint32_t k = (j % 2) ? j - 1 : j + 1;
        a_ij = (a_i0 + (j - 1) * a_i1) % res;
        a_ik = (a_i0 + (k - 1) * a_i1) % res;
    }
    uint32_t stratum = permute_kensler(a_ij, res, p * (j + 1) * 0x51633e2du);
    uint32_t sub_stratum = permute_kensler(a_ik, res, p * (j + 1) * 0x68bc21ebu);
    float jitter = s->jitter ? pcg32_next_float(&s->rng) : .5f;
    return ((float) stratum + ((float) sub_stratum + jitter) / (float) res) / (float) res;
}
static inline float smp_next_1d(osampler *s) {
    if (s->type == B2_SAMPLER_INDEPENDENT) return pcg32_next_float(&s->rng);
    if (s->type == B2_SAMPLER_ORTHOGONAL) return smp_bose(s);
    if (s->type == B2_SAMPLER_LOW_DISCREPANCY) {
        uint32_t i = permute_tea(s->sample_index, s->count, s->perm_seed + s->dim++);
        return radical_inverse_2(i, sample_tea_32(s->perm_seed, 0x48bc48ebu, 4));
    }
    uint32_t ps = s->perm_seed + s->dim++;
    if (s->type == B2_SAMPLER_MULTIJITTER) ps *= 0x45fbe943u;       /* multijitter.cpp:108 */
    uint32_t p = permute_kensler(s->sample_index, s->count, ps);
    float j = s->jitter ? pcg32_next_float(&s->rng) : .5f;
    return ((float) p + j) * s->inv_count;
}
static inline void smp_next_2d(osampler *s, float *x, float *y) {
    if (s->type == B2_SAMPLER_INDEPENDENT) { *x = pcg32_next_float(&s->rng); *y = pcg32_next_float(&s->rng); return; }
    if (s->type == B2_SAMPLER_ORTHOGONAL) { *x = smp_bose(s); *y = smp_bose(s); return; }     /* orthogonal.cpp:128-132 */
    if (s->type == B2_SAMPLER_LOW_DISCREPANCY) {
        uint32_t i = permute_tea(s->sample_index, s->count, s->perm_seed + s->dim++);
        *x = radical_inverse_2(i, sample_tea_32(s->perm_seed, 0x98bc51abu, 4));
        *y = sobol_2(i, sample_tea_32(s->perm_seed, 0x04223e2du, 4));
        return;
    }
    if (s->type == B2_SAMPLER_MULTIJITTER) {                        /* multijitter.cpp:116-144 */
        uint32_t ps = s->perm_seed + s->dim++;
        uint32_t sq = permute_kensler(s->sample_index, s->count, ps * 0x51633e2du);
        uint32_t cy = sq / s->res, cx = sq - cy * s->res;
        uint32_t sx = permute_kensler(cx, s->res, ps * 0x68bc21ebu), sy = permute_kensler(cy, s->res_y, ps * 0x02e5be93u);
        float jx = .5f, jy = .5f;
        if (s->jitter) { jx = pcg32_next_float(&s->rng); jy = pcg32_next_float(&s->rng); }
        *x = B2O_MJ_FMA(((float) sy + jx), s->inv_res_y, (float) cx) * s->inv_res;
        *y = B2O_MJ_FMA(((float) sx + jy), s->inv_res, (float) cy) * s->inv_res_y;
        return;
    }
    uint32_t p = permute_kensler(s->sample_index, s->count, s->perm_seed + s->dim++);
    uint32_t py = p / s->res, px = p - py * s->res;
    float jx = .5f, jy = .5f;
    if (s->jitter) { jx = pcg32_next_float(&s->rng); jy = pcg32_next_float(&s->rng); }
    *x = ((float) px + jx) * s->inv_res; *y = ((float) py + jy) * s->inv_res;
}

static inline float mis_weight(float a, float b) {   /* path.cpp:223-227 */
    a *= a; b *= b;
    return a > 0.f ? a / (a + b) : 0.f;
}

static v3 path_sample(const oscene *s, const B2RenderParams *rp, osampler *rng, oray ray, int *valid_ray,
                      int brute, ocount *cnt) {
    const B2SceneDesc *d = s->desc;
    float eta = 1.f, emission_weight = 1.f;
    v3 throughput = v3_make(1.f, 1.f, 1.f), result = v3_make(0.f, 0.f, 0.f);
    ohit h = scene_intersect(s, &ray, brute); cnt->segments++;
    osi si; si_fill(s, &ray, &h, &si);
    *valid_ray = h.t < INFINITY;
    int emitter = (si.mesh != -1) ? si.emitter : s->env;  /* scene.h:243-253 */
    int active = 1;
    for (int depth = 1;; ++depth) {
        if (emitter >= 0) {                                               /* path.cpp:128-129 */
            const B2Emitter *em = &d->emitters[emitter];
            v3 ev;
            if (si.mesh == -1) ev = env_eval(s, ray.d);                   /* envmap.cpp:134-147 (-si.wi = ray.d) */
            else ev = si.wi.z > 0.f ? v3_make(em->radiance[0], em->radiance[1], em->radiance[2]) : v3_make(0.f, 0.f, 0.f); /* area.cpp:63-71 */
            v3 wt = v3_scale(throughput, emission_weight);
            result = v3_add(result, v3_mul(wt, ev));
        }
        active &= si.t < INFINITY;                                        /* path.cpp:131 */
        if (depth > rp->rr_depth) {                                       /* path.cpp:137-141 */
            float q = b2o_min(v3_hmax(throughput) * b2o_sqr(eta), .95f);
            if (active) active &= smp_next_1d(rng) < q;
            throughput = v3_scale(throughput, b2o_rcp(q));
        }
        if ((uint32_t) depth >= (uint32_t) rp->max_depth || !active) break;   /* path.cpp:147-149 */

        int flip, smooth;
        const B2Bsdf bsdf_si = bsdf_at(d, &d->bsdfs[si.bsdf], &si, &flip, &smooth), *bsdf = &bsdf_si;
        osi sib = si;                                                      /* the interaction as the nested BRDF sees it */
        if (flip) sib.wi.z = -sib.wi.z;
        if (smooth) {                                                      /* path.cpp:153-172 */
            float sx, sy; smp_next_2d(rng, &sx, &sy);
            ods ds; v3 emitter_val = scene_sample_emitter_direction(s, si.p, sx, sy, &ds);
            if (ds.pdf != 0.f) {
                v3 wo = to_local(&si, ds.d);
                if (flip) wo.z = -wo.z;
                v3 bsdf_val = bsdf_eval(bsdf, &sib, wo);
                float bpdf = bsdf_pdf(bsdf, &sib, wo);
                float mis = ds.delta ? 1.f : mis_weight(ds.pdf, bpdf);             /* path.cpp:170 */
                v3 c = v3_mul(v3_mul(v3_scale(throughput, mis), bsdf_val), emitter_val);
                /* scene.cpp:203-207: spec[ray_test(ray)] = 0. The reference traces the shadow ray
                   whenever ds.pdf != 0; the outcome only matters when the contribution is not
                   already zero, so the oracle (and the CUDA path) trace it only then. */
                if (c.x != 0.f || c.y != 0.f || c.z != 0.f) {
                    oray sr; sr.o = si.p; sr.d = ds.d;
                    sr.mint = B2O_RAY_EPS * (1.f + v3_hmax_abs(si.p));
                    sr.maxt = ds.dist * (1.f - B2O_SHADOW_EPS);
                    cnt->shadow_rays++;
                    if (!scene_occluded(s, &sr, brute)) result = v3_add(result, c);
                }
            }
        }
        /* BSDF sampling: g++ evaluates the arguments right to left, so next_2d is drawn
           BEFORE next_1d (path.cpp:177-178) */
        float s2x, s2y; smp_next_2d(rng, &s2x, &s2y);
        float s1 = smp_next_1d(rng);
        obs bs; v3 bsdf_val = bsdf_sample(bsdf, &sib, s1, s2x, s2y, &bs);
        if (flip) bs.wo.z = -bs.wo.z;
        throughput = v3_mul(throughput, bsdf_val);
        active &= throughput.x != 0.f || throughput.y != 0.f || throughput.z != 0.f;
        if (!active) break;
        eta *= bs.eta;
        oray nr; nr.o = si.p; nr.d = to_world(&si, bs.wo);                 /* interaction.h:58-61 */
        nr.mint = (1.f + v3_hmax_abs(si.p)) * B2O_RAY_EPS; nr.maxt = INFINITY;
        ohit h2 = scene_intersect(s, &nr, brute); cnt->segments++;
        osi si2; si_fill(s, &nr, &h2, &si2);
        emitter = (si2.mesh != -1) ? si2.emitter : s->env;
        if (emitter >= 0) {                                               /* path.cpp:194-205, records.h:168-174 */
            v3 dd = v3_sub(si2.p, si.p);
            float dist = sqrtf(v3_dot(dd, dd));
            dd = v3_scale(dd, b2o_rcp(dist));
            if (si2.mesh == -1) dd = nr.d;                                /* d[!it.is_valid()] = -it.wi (records.h:173) */
            float emitter_pdf = !bs.delta ? scene_pdf_emitter_direction(s, emitter, si.p, dd, si2.sh_n, dist) : 0.f;
            emission_weight = mis_weight(bs.pdf, emitter_pdf);
        }
        ray = nr; si = si2;
    }
    return result;
}

/* DirectIntegrator::sample (direct.cpp:106-198) with emitter_samples = E, bsdf_samples = B */
static v3 direct_sample(const oscene *s, const B2RenderParams *rp, osampler *rng, oray ray, int *valid_ray, int brute, ocount *cnt) {
    const B2SceneDesc *d = s->desc;
    const uint32_t E = rp->direct_emitter_samples, B = rp->direct_bsdf_samples;
    const float sum = (float) (E + B);                                     /* direct.cpp:96-100 */
    const float weight_bsdf = 1.f / (float) B, weight_lum = 1.f / (float) E, frac_bsdf = (float) B / sum, frac_lum = (float) E / sum;
    v3 result = v3_make(0.f, 0.f, 0.f);
    ohit h = scene_intersect(s, &ray, brute); cnt->segments++;
    osi si; si_fill(s, &ray, &h, &si);
    *valid_ray = h.t < INFINITY;
    int emitter = (si.mesh != -1) ? si.emitter : s->env;                   /* visible emitters, direct.cpp:120-124 */
    if (emitter >= 0) {
        const B2Emitter *em = &d->emitters[emitter];
        v3 ev;
        if (si.mesh == -1) ev = env_eval(s, ray.d);
        else ev = si.wi.z > 0.f ? v3_make(em->radiance[0], em->radiance[1], em->radiance[2]) : v3_make(0.f, 0.f, 0.f);
        result = v3_add(result, ev);
    }
    if (!(si.t < INFINITY)) return result;
    int flip, smooth;
    const B2Bsdf bsdf_si = bsdf_at(d, &d->bsdfs[si.bsdf], &si, &flip, &smooth), *bsdf = &bsdf_si;
    osi sib = si;
    if (flip) sib.wi.z = -sib.wi.z;
    if (smooth)                                                            /* emitter sampling, direct.cpp:135-163 */
        for (uint32_t i = 0; i < E; ++i) {
            float sx, sy; smp_next_2d(rng, &sx, &sy);
            ods ds; v3 emitter_val = scene_sample_emitter_direction(s, si.p, sx, sy, &ds);
            if (ds.pdf == 0.f) continue;
            v3 wo = to_local(&si, ds.d);
            if (flip) wo.z = -wo.z;
            v3 bsdf_val = bsdf_eval(bsdf, &sib, wo);
            float bpdf = bsdf_pdf(bsdf, &sib, wo);
            float mis = ds.delta ? 1.f : mis_weight(ds.pdf * frac_lum, bpdf * frac_bsdf) * weight_lum;
            v3 c = v3_mul(v3_scale(bsdf_val, mis), emitter_val);
            if (c.x != 0.f || c.y != 0.f || c.z != 0.f) {                  /* scene.cpp:203-207 (see path_sample) */
                oray sr; sr.o = si.p; sr.d = ds.d;
                sr.mint = B2O_RAY_EPS * (1.f + v3_hmax_abs(si.p));
                sr.maxt = ds.dist * (1.f - B2O_SHADOW_EPS);
                cnt->shadow_rays++;
                if (!scene_occluded(s, &sr, brute)) result = v3_add(result, c);
            }
        }
    for (uint32_t i = 0; i < B; ++i) {                                      /* BSDF sampling, direct.cpp:167-195 */
        float s2x, s2y; smp_next_2d(rng, &s2x, &s2y);                      /* g++: next_2d before next_1d */
        float s1 = smp_next_1d(rng);
        obs bs; v3 bsdf_val = bsdf_sample(bsdf, &sib, s1, s2x, s2y, &bs);
        if (flip) bs.wo.z = -bs.wo.z;
        if (!(bsdf_val.x != 0.f || bsdf_val.y != 0.f || bsdf_val.z != 0.f)) continue;
        oray nr; nr.o = si.p; nr.d = to_world(&si, bs.wo);
        nr.mint = (1.f + v3_hmax_abs(si.p)) * B2O_RAY_EPS; nr.maxt = INFINITY;
        ohit h2 = scene_intersect(s, &nr, brute); cnt->segments++;
        osi si2; si_fill(s, &nr, &h2, &si2);
        int em2 = (si2.mesh != -1) ? si2.emitter : s->env;
        if (em2 < 0) continue;
        const B2Emitter *em = &d->emitters[em2];
        v3 ev;
        if (si2.mesh == -1) ev = env_eval(s, nr.d);
        else ev = si2.wi.z > 0.f ? v3_make(em->radiance[0], em->radiance[1], em->radiance[2]) : v3_make(0.f, 0.f, 0.f);
        v3 dd = v3_sub(si2.p, si.p);
        float dist = sqrtf(v3_dot(dd, dd));
        dd = v3_scale(dd, b2o_rcp(dist));
        if (si2.mesh == -1) dd = nr.d;
        float emitter_pdf = !bs.delta ? scene_pdf_emitter_direction(s, em2, si.p, dd, si2.sh_n, dist) : 0.f;
        float w = mis_weight(bs.pdf * frac_bsdf, emitter_pdf * frac_lum);
        result = v3_add(result, v3_scale(v3_scale(v3_mul(bsdf_val, ev), w), weight_bsdf));
    }
    return result;
}

/* ============================================================================================
 * Sensor + render driver — perspective.cpp:182-216, integrator.cpp:181-288, imageblock.cpp:79-172
 * ============================================================================================ */
static oray sensor_sample_ray(const B2Sensor *sn, float sx, float sy) {
    const float *M = sn->sample_to_camera;
    float px = sx + sn->principal_point_offset[0], py = sy + sn->principal_point_offset[1];
    /* Transform * Point (transform.h:103-111): result = col3; fmadd(col_i, arg_i, result); / w.
       arg.z = 0 contributes fmadd(col2, 0, r) = r. */
    float r[4];
    for (int k = 0; k < 4; ++k) {
        float acc = M[4 * k + 3];
        acc = b2o_fma(M[4 * k + 0], px, acc);
        acc = b2o_fma(M[4 * k + 1], py, acc);
        acc = b2o_fma(M[4 * k + 2], 0.f, acc);
        r[k] = acc;
    }
    float iw = b2o_rcp(r[3]);
    v3 near_p = v3_make(r[0] * iw, r[1] * iw, r[2] * iw);
    v3 d = v3_normalize(near_p);
    float inv_z = b2o_rcp(d.z);
    oray ray;
    ray.mint = sn->near_clip * inv_z; ray.maxt = sn->far_clip * inv_z;
    const float *T = sn->to_world;
    ray.o = v3_make(T[3], T[7], T[11]);                                    /* transform_affine(Point(0)) */
    /* Transform * Vector (transform.h:117-125) */
    ray.d.x = b2o_fma(T[2], d.z, b2o_fma(T[1], d.y, T[0] * d.x));
    ray.d.y = b2o_fma(T[6], d.z, b2o_fma(T[5], d.y, T[4] * d.x));
    ray.d.z = b2o_fma(T[10], d.z, b2o_fma(T[9], d.y, T[8] * d.x));
    return ray;
}

static inline uint32_t compact1by1(uint32_t x) {    /* morton_decode: even bits -> x, odd bits -> y */
    x &= 0x55555555u; x = (x ^ (x >> 1)) & 0x33333333u; x = (x ^ (x >> 2)) & 0x0f0f0f0fu;
    x = (x ^ (x >> 4)) & 0x00ff00ffu; x = (x ^ (x >> 8)) & 0x0000ffffu; return x;
}

typedef struct {
    int ox, oy, w, h, border;       /* block offset (film coords), size */
    float *data;                    /* (h+2b) x (w+2b) x 5 */
} oblock;

/* ImageBlock::put(pos, value) — imageblock.cpp:79-172 (m_normalize = false, warn flags on) */
static int block_put(oblock *blk, const B2Film *film, float posx, float posy, const float value[5]) {
    for (int k = 0; k < 5; ++k) {
        if (!(value[k] >= -1e-5f)) return 0;
        if (!(fabsf(value[k]) < INFINITY)) return 0;
    }
    float radius = film->filter_radius;
    int sx = blk->w + 2 * blk->border, sy = blk->h + 2 * blk->border;
    float px = posx - ((float) (blk->ox - blk->border) + .5f), py = posy - ((float) (blk->oy - blk->border) + .5f);
    if (radius > 0.5f + B2O_RAY_EPS) {
        int lox = (int) ceilf(px - radius), loy = (int) ceilf(py - radius);
        int hix = (int) floorf(px + radius), hiy = (int) floorf(py + radius);
        if (lox < 0) lox = 0;
        if (loy < 0) loy = 0;
        if (hix > sx - 1) hix = sx - 1;
        if (hiy > sy - 1) hiy = sy - 1;
        uint32_t n = (uint32_t) ceilf((radius - 2.f * B2O_RAY_EPS) * 2.f);
        float basex = (float) lox - px, basey = (float) loy - py;
        float wx[16], wy[16];
        for (uint32_t i = 0; i < n && i < 16; ++i) {
            float fx = basex + (float) i, fy = basey + (float) i;
            int ix = (int) fabsf(fx * film->filter_scale), iy = (int) fabsf(fy * film->filter_scale);
            if (ix > B2_FILTER_RESOLUTION) ix = B2_FILTER_RESOLUTION;
            if (iy > B2_FILTER_RESOLUTION) iy = B2_FILTER_RESOLUTION;
            wx[i] = film->filter_lut[ix]; wy[i] = film->filter_lut[iy];
        }
        for (uint32_t yr = 0; yr < n; ++yr) {
            int y = loy + (int) yr;
            int enabled = y <= hiy;
            for (uint32_t xr = 0; xr < n; ++xr) {
                int x = lox + (int) xr;
                float weight = wy[yr] * wx[xr];
                enabled &= x <= hix;          /* imageblock.cpp:158 */
                if (enabled) {
                    float *dst = blk->data + 5 * ((size_t) y * sx + x);
                    for (int k = 0; k < 5; ++k) dst[k] += value[k] * weight;
                }
            }
        }
    } else {
        int lox = (int) ceilf(px - .5f), loy = (int) ceilf(py - .5f);
        if (lox >= 0 && loy >= 0 && lox < sx && loy < sy) {
            float *dst = blk->data + 5 * ((size_t) loy * sx + lox);
            for (int k = 0; k < 5; ++k) dst[k] += value[k];
        }
    }
    return 1;
}

typedef struct {
    const oscene *s; const B2RenderParams *rp;
    uint32_t n_blocks; const int32_t *order; uint32_t n_passes, spp_pass;
    oblock *blocks;           /* one per (pass, block) work item */
    uint32_t next, total; pthread_mutex_t mu;
    int brute;
    ocount cnt; uint64_t invalid, samples;
} ojob;

/* render_block (integrator.cpp:181-209) + render_sample (:233-288) for work item `w` */
static void render_item(ojob *j, uint32_t w, ocount *cnt, uint64_t *invalid, uint64_t *samples) {
    const B2SceneDesc *d = j->s->desc; const B2Film *film = &d->film; const B2RenderParams *rp = j->rp;
    uint32_t pass = w / j->n_blocks, b = w % j->n_blocks;        /* pass 0 runs first: remaining_passes = n_passes */
    uint32_t bs = rp->block_size;
    uint64_t block_id = (uint64_t) b + (uint64_t) (j->n_passes - pass - 1) * j->n_blocks;    /* spiral.cpp:41 */
    oblock *blk = &j->blocks[w];
    blk->ox = j->order[2 * b] * (int) bs; blk->oy = j->order[2 * b + 1] * (int) bs;
    blk->w = (int) bs < film->crop_size[0] - blk->ox ? (int) bs : film->crop_size[0] - blk->ox;
    blk->h = (int) bs < film->crop_size[1] - blk->oy ? (int) bs : film->crop_size[1] - blk->oy;
    blk->ox += film->crop_offset[0]; blk->oy += film->crop_offset[1];
    blk->border = film->border;
    blk->data = (float *) calloc((size_t) 5 * (blk->w + 2 * blk->border) * (blk->h + 2 * blk->border), sizeof(float));
    uint32_t pixel_count = bs * bs;
    for (uint32_t i = 0; i < pixel_count; ++i) {
        osampler rng; smp_seed(&rng, rp, block_id * pixel_count + i);                       /* integrator.cpp:198 */
        uint32_t mx = compact1by1(i), my = compact1by1(i >> 1);
        if ((int) mx >= blk->w || (int) my >= blk->h) continue;
        float posx = (float) (mx + (uint32_t) blk->ox), posy = (float) (my + (uint32_t) blk->oy);
        for (uint32_t k = 0; k < j->spp_pass; ++k) {
            float jx, jy; smp_next_2d(&rng, &jx, &jy);                                       /* :242 */
            float psx = posx + jx, psy = posy + jy;
            (void) smp_next_1d(&rng);                                                        /* :252 wavelength sample */
            float ax = (psx - (float) film->crop_offset[0]) / (float) film->crop_size[0],
                  ay = (psy - (float) film->crop_offset[1]) / (float) film->crop_size[1];   /* :254-256 */
            oray ray = sensor_sample_ray(&d->sensor, ax, ay);
            int valid;
            v3 L = (rp->direct_emitter_samples | rp->direct_bsdf_samples) ? direct_sample(j->s, rp, &rng, ray, &valid, j->brute, cnt)
                                                                           : path_sample(j->s, rp, &rng, ray, &valid, j->brute, cnt);
            /* ray_weight = 1 in rgb mode ; srgb_to_xyz spectrum.h:221-228 (M * rgb, matrix.h:170-177) */
            float val[5];
            val[0] = b2o_fma(0.180423f, L.z, b2o_fma(0.357580f, L.y, 0.412453f * L.x));
            val[1] = b2o_fma(0.072169f, L.z, b2o_fma(0.715160f, L.y, 0.212671f * L.x));
            val[2] = b2o_fma(0.950227f, L.z, b2o_fma(0.119193f, L.y, 0.019334f * L.x));
            val[3] = valid ? 1.f : 0.f; val[4] = 1.f;
            if (!block_put(blk, film, psx, psy, val)) (*invalid)++;
            (*samples)++;
            smp_advance(&rng);                                                               /* :206 */
        }
    }
}

static void *worker(void *arg) {
    ojob *j = (ojob *) arg;
    ocount cnt = { 0, 0 }; uint64_t invalid = 0, samples = 0;
    for (;;) {
        pthread_mutex_lock(&j->mu);
        uint32_t w = j->next < j->total ? j->next++ : 0xffffffffu;
        pthread_mutex_unlock(&j->mu);
        if (w == 0xffffffffu) break;
        render_item(j, w, &cnt, &invalid, &samples);
    }
    pthread_mutex_lock(&j->mu);
    j->cnt.segments += cnt.segments; j->cnt.shadow_rays += cnt.shadow_rays; j->invalid += invalid; j->samples += samples;
    pthread_mutex_unlock(&j->mu);
    return NULL;
}

/* SamplingIntegrator::render (integrator.cpp:51-179) + HDRFilm::put (hdrfilm.cpp:207-211).
   Blocks are rendered by `threads` workers into private buffers and merged into the film in
   the reference's single-thread order (pass by pass, spiral order), so the result does not
   depend on the thread count. Honors rank/world_size (blocks b % world == rank). */
B2O_API int b2o_render(const B2SceneDesc *d, const B2RenderParams *rp, float *xyzaw, B2Stats *stats,
                       int threads, int brute) {
    oscene *s = oscene_create(d);
    const B2Film *film = &d->film;
    uint32_t nb = 0; b2o_spiral(film->crop_size[0], film->crop_size[1], (int) rp->block_size, &nb, NULL);
    int32_t *order = (int32_t *) malloc(sizeof(int32_t) * 2 * nb);
    b2o_spiral(film->crop_size[0], film->crop_size[1], (int) rp->block_size, &nb, order);
    uint32_t spp_pass = rp->samples_per_pass ? (rp->samples_per_pass < rp->spp ? rp->samples_per_pass : rp->spp) : rp->spp;
    if (rp->spp % spp_pass) { free(order); oscene_destroy(s); return B2_ERR_INVALID; }
    ojob j; memset(&j, 0, sizeof j);
    j.s = s; j.rp = rp; j.n_blocks = nb; j.order = order; j.n_passes = rp->spp / spp_pass; j.spp_pass = spp_pass;
    j.total = nb * j.n_passes; j.brute = brute;
    j.blocks = (oblock *) calloc(j.total, sizeof(oblock));
    pthread_mutex_init(&j.mu, NULL);
    int world = rp->world_size > 0 ? rp->world_size : 1;
    if (threads < 1) threads = 1;
    /* rank filter: render only owned blocks (others stay empty) */
    if (world > 1) {
        /* simple: single-threaded loop over owned items */
        ocount cnt = { 0, 0 }; uint64_t inv = 0, smp = 0;
        for (uint32_t w = 0; w < j.total; ++w)
            if ((int) ((w % nb) % (uint32_t) world) == rp->rank) render_item(&j, w, &cnt, &inv, &smp);
        j.cnt = cnt; j.invalid = inv; j.samples = smp;
    } else if (threads == 1) {
        j.next = 0; worker(&j);
    } else {
        pthread_t *th = (pthread_t *) malloc(sizeof(pthread_t) * (size_t) threads);
        for (int t = 0; t < threads; ++t) pthread_create(&th[t], NULL, worker, &j);
        for (int t = 0; t < threads; ++t) pthread_join(th[t], NULL);
        free(th);
    }
    /* merge: ImageBlock::put(block) imageblock.cpp:49-77 — film += block over the overlap */
    int W = film->crop_size[0], H = film->crop_size[1];
    memset(xyzaw, 0, sizeof(float) * 5 * (size_t) W * H);
    for (uint32_t w = 0; w < j.total; ++w) {
        oblock *blk = &j.blocks[w];
        if (!blk->data) continue;
        int sx = blk->w + 2 * blk->border, sy = blk->h + 2 * blk->border;
        for (int y = 0; y < sy; ++y) {
            int fy = blk->oy - film->crop_offset[1] - blk->border + y;
            if (fy < 0 || fy >= H) continue;
            for (int x = 0; x < sx; ++x) {
                int fx = blk->ox - film->crop_offset[0] - blk->border + x;
                if (fx < 0 || fx >= W) continue;
                const float *src = blk->data + 5 * ((size_t) y * sx + x);
                float *dst = xyzaw + 5 * ((size_t) fy * W + fx);
                for (int k = 0; k < 5; ++k) dst[k] += src[k];
            }
        }
        free(blk->data);
    }
    if (stats) {
        memset(stats, 0, sizeof *stats);
        stats->n_samples = j.samples; stats->n_segments = j.cnt.segments; stats->n_shadow_rays = j.cnt.shadow_rays;
        stats->n_invalid = j.invalid; stats->bvh_nodes = (uint64_t) s->top.n_nodes; stats->bvh_tris = s->n_tris;
    }
    free(j.blocks); free(order); pthread_mutex_destroy(&j.mu);
    oscene_destroy(s);
    return 0;
}

/* HDRFilm::bitmap(): XYZAW -> RGB — hdrfilm.cpp:251-322, bitmap.cpp:420-427, struct.cpp:1750-1812
   (value = blend of the channels in fp32 — first product rounded, the others fused —, then * (1/W)) */
B2O_API int b2o_develop(const float *xyzaw, int n_pixels, float *rgb) {
    for (int i = 0; i < n_pixels; ++i) {
        const float *p = xyzaw + 5 * (size_t) i;
        float iw = 1.f / p[4];
        /* the compiled conversion loop: first product rounded, the others fused — struct.cpp:1380-1386 (muls, then fmadd231) */
        float r = 3.240479f * p[0], g = -0.969256f * p[0], b = 0.055648f * p[0];
        r = b2o_fma(-1.537150f, p[1], r); r = b2o_fma(-0.498535f, p[2], r);
        g = b2o_fma(1.875991f, p[1], g); g = b2o_fma(0.041556f, p[2], g);
        b = b2o_fma(-0.204043f, p[1], b); b = b2o_fma(1.057311f, p[2], b);
        rgb[3 * i] = r * iw; rgb[3 * i + 1] = g * iw; rgb[3 * i + 2] = b * iw;
    }
    return 0;
}

/* ============================================================================================
 * Stage-level entry points (same signatures as include/b2pt.h b2_stage_*, scene passed as desc)
 * ============================================================================================ */
B2O_API int b2o_stage_raygen(const B2SceneDesc *d, uint32_t n, const uint32_t *px, const uint32_t *py,
                             const uint64_t *seed_offset, uint64_t base_seed,
                             float *ray_o, float *ray_d, uint64_t *rng_state, float *pos) {
    const B2Film *film = &d->film;
    for (uint32_t i = 0; i < n; ++i) {
        pcg32 rng; sampler_seed(&rng, base_seed, seed_offset[i]);
        float jx = pcg32_next_float(&rng), jy = pcg32_next_float(&rng);
        float psx = (float) px[i] + jx, psy = (float) py[i] + jy;
        (void) pcg32_next_float(&rng);
        float ax = (psx - (float) film->crop_offset[0]) / (float) film->crop_size[0],
              ay = (psy - (float) film->crop_offset[1]) / (float) film->crop_size[1];
        oray r = sensor_sample_ray(&d->sensor, ax, ay);
        ray_o[4 * i] = r.o.x; ray_o[4 * i + 1] = r.o.y; ray_o[4 * i + 2] = r.o.z; ray_o[4 * i + 3] = r.mint;
        ray_d[4 * i] = r.d.x; ray_d[4 * i + 1] = r.d.y; ray_d[4 * i + 2] = r.d.z; ray_d[4 * i + 3] = r.maxt;
        rng_state[i] = rng.state; pos[2 * i] = psx; pos[2 * i + 1] = psy;
    }
    return 0;
}
static oray load_ray(const float *ray_o, const float *ray_d, uint32_t i) {
    oray r; r.o = v3_make(ray_o[4 * i], ray_o[4 * i + 1], ray_o[4 * i + 2]); r.mint = ray_o[4 * i + 3];
    r.d = v3_make(ray_d[4 * i], ray_d[4 * i + 1], ray_d[4 * i + 2]); r.maxt = ray_d[4 * i + 3];
    return r;
}
B2O_API int b2o_stage_intersect(const B2SceneDesc *d, uint32_t n, const float *ray_o, const float *ray_d,
                                float *hit, int32_t *inst, int brute) {
    oscene *s = oscene_create(d);
    for (uint32_t i = 0; i < n; ++i) {
        oray r = load_ray(ray_o, ray_d, i);
        ohit h = scene_intersect(s, &r, brute);
        hit[4 * i] = h.t; hit[4 * i + 1] = h.u; hit[4 * i + 2] = h.v; hit[4 * i + 3] = b2o_u2f(h.prim);
        if (inst) inst[i] = h.inst;
    }
    oscene_destroy(s);
    return 0;
}
B2O_API int b2o_stage_occluded(const B2SceneDesc *d, uint32_t n, const float *ray_o, const float *ray_d,
                               uint8_t *occluded, int brute) {
    oscene *s = oscene_create(d);
    for (uint32_t i = 0; i < n; ++i) {
        oray r = load_ray(ray_o, ray_d, i);
        occluded[i] = (uint8_t) scene_occluded(s, &r, brute);
    }
    oscene_destroy(s);
    return 0;
}
B2O_API int b2o_stage_si_fill(const B2SceneDesc *d, uint32_t n, const float *ray_o, const float *ray_d,
                              const float *hit, const int32_t *inst, float *si_out) {
    oscene *s = oscene_create(d);
    for (uint32_t i = 0; i < n; ++i) {
        oray r = load_ray(ray_o, ray_d, i);
        ohit h; h.t = hit[4 * i]; h.u = hit[4 * i + 1]; h.v = hit[4 * i + 2]; h.prim = b2o_f2u(hit[4 * i + 3]); h.inst = inst ? inst[i] : -1;
        osi si; si_fill(s, &r, &h, &si);
        float *o = si_out + 16 * (size_t) i;
        o[0] = si.p.x; o[1] = si.p.y; o[2] = si.p.z;
        o[3] = si.n.x; o[4] = si.n.y; o[5] = si.n.z;
        o[6] = si.sh_n.x; o[7] = si.sh_n.y; o[8] = si.sh_n.z;
        o[9] = si.sh_s.x; o[10] = si.sh_s.y; o[11] = si.sh_s.z;
        o[12] = si.wi.x; o[13] = si.wi.y; o[14] = si.wi.z;
        o[15] = b2o_u2f(si.mesh != -1 ? (uint32_t) si.bsdf : 0xffffffffu);
    }
    oscene_destroy(s);
    return 0;
}

/* One iteration of PathIntegrator::sample's loop body (path.cpp:124-208) from an explicit per-lane state: the stage as
   the wavefront runs it (include/b2pt.h b2_stage_shade). Every lane must have hit something (hit t < inf) and have
   depth < max_depth (the lanes k_trace puts on its CONTINUE list). The emission weight of an emitter hit is formed on
   arrival from the stored pdf of the last BSDF sample (result_pdf.w), which is what path.cpp:194-205 computes one
   statement earlier. Independent sampler only (rng_state = PCG32 state, default stream).
   depth_flags: depth | B2_LANE_DELTA (0x200): the last BSDF sample was a delta lobe | B2_LANE_ALPHA (0x100, out) |
                B2_LANE_PARKED (0x2000, out): the path ended inside the loop body (ray_d.w = -1)
   Shadow entries are written in lane order: sh_c = (contribution.rgb, lane index as bits). */
B2O_API int b2o_stage_shade(const B2SceneDesc *d, const B2RenderParams *rp, uint32_t n, float *ray_o, float *ray_d,
                            const float *hit, const int32_t *inst, float *throughput_eta, float *result_pdf,
                            uint64_t *rng_state, uint32_t *depth_flags, uint32_t *n_shadow, float *sh_o, float *sh_d, float *sh_c) {
    oscene *s = oscene_create(d);
    uint32_t ns = 0;
    for (uint32_t i = 0; i < n; ++i) {
        oray ray = load_ray(ray_o, ray_d, i);
        ohit h; h.t = hit[4 * i]; h.u = hit[4 * i + 1]; h.v = hit[4 * i + 2]; h.prim = b2o_f2u(hit[4 * i + 3]); h.inst = inst ? inst[i] : -1;
        osi si; si_fill(s, &ray, &h, &si);
        v3 throughput = v3_make(throughput_eta[4 * i], throughput_eta[4 * i + 1], throughput_eta[4 * i + 2]);
        float eta = throughput_eta[4 * i + 3];
        v3 result = v3_make(result_pdf[4 * i], result_pdf[4 * i + 1], result_pdf[4 * i + 2]);
        float bs_pdf = result_pdf[4 * i + 3];
        osampler rng; memset(&rng, 0, sizeof rng);
        rng.type = B2_SAMPLER_INDEPENDENT; rng.rng.state = rng_state[i]; rng.rng.inc = (PCG32_DEFAULT_STREAM << 1u) | 1u;
        uint32_t fl = depth_flags[i];
        int depth = (int) (fl & 0xffu), parked = 0;
        if (depth == 1) fl |= 0x100u;                                     /* valid_ray, path.cpp:121 */
        int emitter = si.emitter;
        if (emitter >= 0) {                                               /* path.cpp:128-129 with the weight of :194-205 */
            float emission_weight = 1.f;
            if (depth > 1) {
                float emitter_pdf = 0.f;
                if (!(fl & 0x200u)) {
                    v3 dd = v3_sub(si.p, ray.o);
                    float dist = sqrtf(v3_dot(dd, dd));
                    dd = v3_scale(dd, b2o_rcp(dist));
                    emitter_pdf = scene_pdf_emitter_direction(s, emitter, ray.o, dd, si.sh_n, dist);
                }
                emission_weight = mis_weight(bs_pdf, emitter_pdf);
            }
            const B2Emitter *em = &d->emitters[emitter];
            v3 ev = si.wi.z > 0.f ? v3_make(em->radiance[0], em->radiance[1], em->radiance[2]) : v3_make(0.f, 0.f, 0.f);
            result = v3_add(result, v3_mul(v3_scale(throughput, emission_weight), ev));
        }
        int active = 1;
        if (depth > rp->rr_depth) {                                       /* path.cpp:137-141 */
            float q = b2o_min(v3_hmax(throughput) * b2o_sqr(eta), .95f);
            active = smp_next_1d(&rng) < q;
            throughput = v3_scale(throughput, b2o_rcp(q));
        }
        oray nr = ray;
        if (!active) parked = 1;
        else {
            int flip, smooth;
            const B2Bsdf bsdf_si = bsdf_at(d, &d->bsdfs[si.bsdf], &si, &flip, &smooth), *bsdf = &bsdf_si;
            osi sib = si;
            if (flip) sib.wi.z = -sib.wi.z;
            if (smooth) {                                                  /* path.cpp:153-172 */
                float sx, sy; smp_next_2d(&rng, &sx, &sy);
                ods ds; v3 emitter_val = scene_sample_emitter_direction(s, si.p, sx, sy, &ds);
                if (ds.pdf != 0.f) {
                    v3 wo = to_local(&si, ds.d);
                    if (flip) wo.z = -wo.z;
                    v3 bsdf_val = bsdf_eval(bsdf, &sib, wo);
                    float bpdf = bsdf_pdf(bsdf, &sib, wo);
                    float mis = ds.delta ? 1.f : mis_weight(ds.pdf, bpdf);
                    v3 c = v3_mul(v3_mul(v3_scale(throughput, mis), bsdf_val), emitter_val);
                    if (c.x != 0.f || c.y != 0.f || c.z != 0.f) {
                        float *o = sh_o + 4 * (size_t) ns, *dd = sh_d + 4 * (size_t) ns, *cc = sh_c + 4 * (size_t) ns;
                        o[0] = si.p.x; o[1] = si.p.y; o[2] = si.p.z; o[3] = B2O_RAY_EPS * (1.f + v3_hmax_abs(si.p));
                        dd[0] = ds.d.x; dd[1] = ds.d.y; dd[2] = ds.d.z; dd[3] = ds.dist * (1.f - B2O_SHADOW_EPS);
                        cc[0] = c.x; cc[1] = c.y; cc[2] = c.z; cc[3] = b2o_u2f(i);
                        ++ns;
                    }
                }
            }
            float s2x, s2y; smp_next_2d(&rng, &s2x, &s2y);               /* path.cpp:177-178 (g++: next_2d first) */
            float s1 = smp_next_1d(&rng);
            obs bs; v3 bsdf_val = bsdf_sample(bsdf, &sib, s1, s2x, s2y, &bs);
            if (flip) bs.wo.z = -bs.wo.z;
            throughput = v3_mul(throughput, bsdf_val);
            if (!(throughput.x != 0.f || throughput.y != 0.f || throughput.z != 0.f)) parked = 1;   /* path.cpp:182-184 */
            else {
                eta *= bs.eta;
                nr.o = si.p; nr.d = to_world(&si, bs.wo);
                nr.mint = (1.f + v3_hmax_abs(si.p)) * B2O_RAY_EPS; nr.maxt = INFINITY;
                bs_pdf = bs.pdf;
                fl = (fl & ~(0xffu | 0x200u)) | (uint32_t) (depth + 1) | (bs.delta ? 0x200u : 0u);
            }
        }
        if (parked) { fl |= 0x2000u; ray_d[4 * i] = 0.f; ray_d[4 * i + 1] = 0.f; ray_d[4 * i + 2] = 0.f; ray_d[4 * i + 3] = -1.f; }
        else {
            ray_o[4 * i] = nr.o.x; ray_o[4 * i + 1] = nr.o.y; ray_o[4 * i + 2] = nr.o.z; ray_o[4 * i + 3] = nr.mint;
            ray_d[4 * i] = nr.d.x; ray_d[4 * i + 1] = nr.d.y; ray_d[4 * i + 2] = nr.d.z; ray_d[4 * i + 3] = nr.maxt;
        }
        throughput_eta[4 * i] = throughput.x; throughput_eta[4 * i + 1] = throughput.y; throughput_eta[4 * i + 2] = throughput.z; throughput_eta[4 * i + 3] = eta;
        result_pdf[4 * i] = result.x; result_pdf[4 * i + 1] = result.y; result_pdf[4 * i + 2] = result.z; result_pdf[4 * i + 3] = bs_pdf;
        rng_state[i] = rng.rng.state; depth_flags[i] = fl;
    }
    *n_shadow = ns;
    oscene_destroy(s);
    return 0;
}

/* ImageBlock::put (imageblock.cpp:79-172) of n samples into a film of crop_h x crop_w x 5 (XYZAW): sample i belongs to the
   image block of pixel (px, py) (crop-relative) and lands at pos = position_sample (film coordinates); value rgb is
   converted to XYZ like render_sample does (integrator.cpp:273). Blocks are merged like HDRFilm::put. Returns the number
   of samples rejected by the validity rule in *n_invalid. */
B2O_API int b2o_stage_film_put(const B2SceneDesc *d, uint32_t block_size, uint32_t n, const uint32_t *px, const uint32_t *py,
                               const float *pos, const float *rgb, const uint8_t *alpha, float *xyzaw, uint32_t *n_invalid) {
    const B2Film *film = &d->film;
    int W = film->crop_size[0], H = film->crop_size[1];
    memset(xyzaw, 0, (size_t) W * H * 5 * sizeof(float));
    uint32_t bad = 0;
    for (uint32_t i = 0; i < n; ++i) {
        oblock blk;
        int bs = (int) block_size;
        blk.ox = ((int) px[i] / bs) * bs; blk.oy = ((int) py[i] / bs) * bs;
        blk.w = bs < W - blk.ox ? bs : W - blk.ox; blk.h = bs < H - blk.oy ? bs : H - blk.oy;
        blk.ox += film->crop_offset[0]; blk.oy += film->crop_offset[1];
        blk.border = film->border;
        int sx = blk.w + 2 * blk.border, sy = blk.h + 2 * blk.border;
        blk.data = (float *) calloc((size_t) 5 * sx * sy, sizeof(float));
        const float *L = rgb + 3 * (size_t) i;
        float val[5];
        val[0] = b2o_fma(0.180423f, L[2], b2o_fma(0.357580f, L[1], 0.412453f * L[0]));
        val[1] = b2o_fma(0.072169f, L[2], b2o_fma(0.715160f, L[1], 0.212671f * L[0]));
        val[2] = b2o_fma(0.950227f, L[2], b2o_fma(0.119193f, L[1], 0.019334f * L[0]));
        val[3] = alpha[i] ? 1.f : 0.f; val[4] = 1.f;
        if (!block_put(&blk, film, pos[2 * i], pos[2 * i + 1], val)) ++bad;
        for (int y = 0; y < sy; ++y)                                       /* HDRFilm::put -> ImageBlock::put(block), imageblock.cpp:49-77 */
            for (int x = 0; x < sx; ++x) {
                int fx = blk.ox - film->crop_offset[0] - blk.border + x, fy = blk.oy - film->crop_offset[1] - blk.border + y;
                if (fx < 0 || fy < 0 || fx >= W || fy >= H) continue;
                for (int k = 0; k < 5; ++k) xyzaw[5 * ((size_t) fy * W + fx) + k] += blk.data[5 * ((size_t) y * sx + x) + k];
            }
        free(blk.data);
    }
    *n_invalid = bad;
    return 0;
}

/* ============================================================================================
 * BSDF unit entry points (for the reference's closed-form tests: test_diffuse.py,
 * test_dielectric.py, test_microfacet.py, test_fresnel.py). wi/wo in local coordinates.
 * ============================================================================================ */
static osi local_si(const float wi[3]) {
    osi si; memset(&si, 0, sizeof si);
    si.t = 1.f; si.sh_n = v3_make(0, 0, 1); si.sh_s = v3_make(1, 0, 0); si.sh_t = v3_make(0, 1, 0);
    si.n = si.sh_n; si.wi = v3_make(wi[0], wi[1], wi[2]);
    return si;
}
B2O_API void b2o_bsdf_eval(const B2Bsdf *b, const float wi[3], const float wo[3], float out_rgb[3], float *out_pdf) {
    osi si = local_si(wi);
    v3 w = v3_make(wo[0], wo[1], wo[2]);
    v3 v = bsdf_eval(b, &si, w);
    out_rgb[0] = v.x; out_rgb[1] = v.y; out_rgb[2] = v.z;
    *out_pdf = bsdf_pdf(b, &si, w);
}
B2O_API void b2o_bsdf_sample(const B2Bsdf *b, const float wi[3], float sample1, const float sample2[2],
                             float out_wo[3], float *out_pdf, float *out_eta, int *out_delta, float out_weight[3]) {
    osi si = local_si(wi);
    obs bs; v3 w = bsdf_sample(b, &si, sample1, sample2[0], sample2[1], &bs);
    out_wo[0] = bs.wo.x; out_wo[1] = bs.wo.y; out_wo[2] = bs.wo.z;
    *out_pdf = bs.pdf; *out_eta = bs.eta; *out_delta = bs.delta;
    out_weight[0] = w.x; out_weight[1] = w.y; out_weight[2] = w.z;
}
/* unit-test entry points: texture lookups and sampler sequences */
B2O_API void b2o_texture_eval(const B2Texture *t, uint32_t n, const float *uv, float *out_rgb) {
    for (uint32_t i = 0; i < n; ++i) {
        v3 c = texture_eval(t, uv + 2 * (size_t) i);
        out_rgb[3 * (size_t) i] = c.x; out_rgb[3 * (size_t) i + 1] = c.y; out_rgb[3 * (size_t) i + 2] = c.z;
    }
}
/* the first `n_samples` samples of one pixel's sampler (seed_offset as in render_block): per sample the draws named by
   `pattern` ('1' = next_1d, '2' = next_2d), written consecutively */
B2O_API void b2o_sampler_draws(const B2RenderParams *rp, uint64_t seed_offset, uint32_t n_samples, const char *pattern, float *out) {
    osampler s; smp_seed(&s, rp, seed_offset);
    size_t k = 0;
    for (uint32_t i = 0; i < n_samples; ++i) {
        for (const char *p = pattern; *p; ++p) {
            if (*p == '2') { smp_next_2d(&s, &out[k], &out[k + 1]); k += 2; }
            else out[k++] = smp_next_1d(&s);
        }
        smp_advance(&s);
    }
}
B2O_API void b2o_fresnel(float cos_theta_i, float eta, float out[4]) {
    fresnel(cos_theta_i, eta, &out[0], &out[1], &out[2], &out[3]);
}
B2O_API float b2o_fresnel_conductor(float cos_theta_i, float eta_r, float eta_i) {
    return fresnel_conductor(cos_theta_i, eta_r, eta_i);
}
B2O_API void b2o_math_probe(float x, float out[6]) {
    b2o_sincos(x, &out[0], &out[1]);
    out[2] = b2o_exp(x); out[3] = b2o_log(x); out[4] = b2o_erf(x); out[5] = b2o_erfinv(x);
}
B2O_API void b2o_warp_cosine_hemisphere(float sx, float sy, float out[3]) {
    v3 v = square_to_cosine_hemisphere(sx, sy); out[0] = v.x; out[1] = v.y; out[2] = v.z;
}
