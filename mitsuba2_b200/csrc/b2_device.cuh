// b2_device.cuh — device-side scene layout and the per-lane shading functions of the
// wavefront stages (SurfaceInteraction fill, BSDFs, emitter sampling, MIS).
// Citations are reference file:line (mitsuba2 @ 62863cb3).
#pragma once
#include "../../include/b2pt.h"
#include "b2_math.cuh"

// Scene-table loads: the tables live in global memory or (small scenes) in a shared-memory copy,
// so they are read through generic pointers.
#define B2_LD(p) (*(p))

// ---- scene layout in HBM --------------------------------------------------------------------
// Triangles are stored in BVH leaf order ("slots"), two SoA streams of 3 x float4 per slot:
//   tri_pos[3*s + 0] = (p0.xyz, prim id bits)     prim = global triangle index (mesh order, face order)
//   tri_pos[3*s + 1] = (p1.xyz, mesh id bits)
//   tri_pos[3*s + 2] = (p2.xyz, 0)
//   tri_isect[3*s + 0..2] = (p0.xyz, prim id bits), (p1-p0, 0), (p2-p0, 0)   traversal copy
//   tri_nrm[3*s + 0] = (n0.xyz, bsdf index bits)  vertex normals (mesh.h:344-348)
//   tri_nrm[3*s + 1] = (n1.xyz, emitter index bits or -1)
//   tri_nrm[3*s + 2] = (n2.xyz, has_normals flag bits)
// BVH4 node = 8 x float4 = 128 B (one cache line):
//   [0] lo.x of children 0..3   [1] lo.y   [2] lo.z   [3] hi.x   [4] hi.y   [5] hi.z
//   [6] child[0..3] as int bits: 0 = empty, > 0 = inner node index, < 0 = leaf:
//       first slot = c & 0x07ffffff, count = (c >> 27) & 0xf
//   [7] unused
struct DEmitter {
    int32_t type, mesh;
    float radiance[3];
    uint32_t first_prim, n_faces, cdf_offset;
    float cdf_sum, cdf_norm;
    uint32_t valid_first, valid_last;
    int32_t shape;              // 0: attached to a mesh; k > 0: analytic shape k - 1
    float position[3];          // point / spot: trafo.translation() (point.cpp:92, spot.cpp:108); directional: to_world * (0,0,1)
    float to_object[12];        // spot: world -> local (3x4 row-major), spot.cpp:117
    float cutoff_angle, beam_width, cos_cutoff, cos_beam, inv_transition;    // spot.cpp:88-95 (host-derived)
    float _pad[3];
};

// analytic shapes (src/shapes/rectangle.cpp, sphere.cpp): 3x4 row-major transforms + the values update() derives
struct DShape {
    int32_t type, flip_normals, bsdf, emitter;
    float to_world[12], to_object[12];
    float center[3], radius;
    float n[3], inv_area;       // rectangle: m_frame.n; both: m_inv_surface_area
    float dp_du[3], _pad;       // rectangle: m_frame.s
};

// environment map (src/emitters/envmap.cpp): RGBA texels + the Hierarchical2D<Float,0> MIP levels
// (include/mitsuba/core/distr_2d.h:372-471) concatenated in one buffer
struct DEnv {
    const float4 *data;         // height x width RGBA
    const float *levels;        // level 0 (normalised luminance*sin(theta)), level 1 (patch averages), ... coarsest
    int32_t index;              // index of the envmap in the emitter list, -1: no environment
    int32_t width, height, n_levels;
    uint32_t level_off[16], level_w[16];
    float scale, radius;        // 'scale'; bounding-sphere radius of the scene (envmap.cpp:127-131)
    int32_t type;               // B2_EMITTER_ENVMAP or B2_EMITTER_CONSTANT (src/emitters/constant.cpp)
    float radiance[3];          // constant: the radiance
    float to_world[12], to_object[12];
};

// one-level instancing (src/shapes/instance.cpp): affine 3x4 transforms, row-major
struct DInstance {
    float to_world[12];
    float to_object[12];
    int32_t root;               // root node of the shapegroup's BVH in DScene.nodes
    int32_t _pad[3];
};

struct DScene {
    const float4 *nodes;
    const float4 *tri_pos;
    const float4 *tri_nrm;
    const float4 *tri_isect;    // traversal copy: (p0, prim), (e1, -), (e2, -)
    const B2Bsdf *bsdfs;
    const DEmitter *emitters;
    const float *cdf;
    const uint32_t *prim_to_slot;
    const float4 *tri_uv;       // NULL unless a mesh has texture coordinates: (uv0, uv1), (uv2, has_uv, -) per triangle
    const B2Texture *textures;  // `data` patched to device memory
    const DShape *shapes;       // analytic shapes; their slots carry (kind, shape index) instead of vertices
    uint32_t n_nodes, n_tris, n_emitters, n_bsdfs;
    const DInstance *instances; // NULL for flat scenes
    const uint32_t *inst_order; // TLAS leaf order -> instance index
    uint32_t n_instances, tlas_root, n_top_tris, _pad0;
    uint32_t smem_nodes;        // nodes [0, smem_nodes) are staged in shared memory
    const float4 *flat_leaves;  // tiny flat scenes: the BVH leaves as a list, 2 float4 each: (lo.xyz, leaf code), (hi.xyz, -)
    uint32_t n_flat_leaves;     // 0: use the BVH; 1..32: exhaustive leaf-box scan (no tree, no stack, no divergence)
    uint32_t _pad1;
    uint32_t smem_tris;         // 1: all triangle positions staged in shared memory too
    const unsigned char *blob;  // one allocation holding tri_pos | tri_nrm | prim_to_slot | cdf | bsdfs | emitters
    uint32_t blob_bytes;        // size of the blob (multiple of 16)
    uint32_t shade_smem;        // 1: k_shade stages the whole blob into shared memory
    B2Sensor sensor;
    B2Film film;
    DEnv env;
};

// The shading tables as seen by one kernel: pointers into global memory or into the CTA's
// shared-memory copy (kept in registers; the kernel parameter block itself is never written).
struct Tables {
    const float4 *tri_pos, *tri_nrm;
    const uint32_t *prim_to_slot;
    const float *cdf;
    const B2Bsdf *bsdfs;
    const DEmitter *emitters;
    const DInstance *instances;
    uint32_t n_emitters;
    const float4 *tri_uv;
    const B2Texture *textures;
    const DShape *shapes;
};
B2_D Tables make_tables(const DScene &sc) {
    Tables t; t.tri_pos = sc.tri_pos; t.tri_nrm = sc.tri_nrm; t.prim_to_slot = sc.prim_to_slot; t.cdf = sc.cdf;
    t.bsdfs = sc.bsdfs; t.emitters = sc.emitters; t.instances = sc.instances; t.n_emitters = sc.n_emitters; t.tri_uv = sc.tri_uv; t.textures = sc.textures; t.shapes = sc.shapes; return t;
}

// Compile-time scene features: the stage kernels are instantiated per feature set so that a scene
// pays (instruction cache, registers) only for the BSDFs / emitters / filter it contains.
// F_TEX ("extended scene"): a mesh has texture coordinates (uv-derived dp_du, mesh.cpp:492-511), a BSDF parameter is
// textured, or the scene contains analytic shapes
enum : uint32_t { F_ROUGH = 1u, F_DIELECTRIC = 2u, F_ENV = 4u, F_WIDE = 8u, F_TEX = 16u, F_COUNT = 32u };

struct Ray { float3 o, d; float mint, maxt; };

// Mesh::ray_intersect_triangle — include/mitsuba/render/mesh.h:194-226
B2_D bool tri_intersect(float3 p0, float3 p1, float3 p2, const Ray &ray, float &t, float &u, float &v) {
    float3 e1 = p1 - p0, e2 = p2 - p0;
    float3 pvec = cross(ray.d, e2);
    float inv_det = b2_rcp(dot(e1, pvec));
    float3 tvec = ray.o - p0;
    u = dot(tvec, pvec) * inv_det;
    bool active = u >= 0.f && u <= 1.f;
    float3 qvec = cross(tvec, e1);
    v = dot(ray.d, qvec) * inv_det;
    active &= v >= 0.f && u + v <= 1.f;
    t = dot(e2, qvec) * inv_det;
    active &= t >= ray.mint && t <= ray.maxt;
    return active;
}

// Rectangle::ray_intersect_preliminary (rectangle.cpp:139-158) / Sphere::ray_intersect_preliminary (sphere.cpp:273-305:
// the quadratic is solved in double, math.h:371-413). M = 3x4 row-major.
B2_D float3 xf_point(const float *M, float3 p);
B2_D float3 xf_vector(const float *M, float3 d);
B2_D bool shape_isect(const DShape &sh, const Ray &ray, float tmax, float &t, float &u, float &v) {
    u = 0.f; v = 0.f;
    if (sh.type == B2_SHAPE_RECTANGLE) {
        float3 o = xf_point(sh.to_object, ray.o), d = xf_vector(sh.to_object, ray.d);
        t = -o.z * b2_rcp(d.z);
        u = b2_fma(d.x, t, o.x); v = b2_fma(d.y, t, o.y);
        return t >= ray.mint && t <= tmax && fabsf(u) <= 1.f && fabsf(v) <= 1.f;
    }
    double mint = (double) ray.mint, maxt = (double) tmax;
    double ox = (double) ray.o.x - (double) sh.center[0], oy = (double) ray.o.y - (double) sh.center[1], oz = (double) ray.o.z - (double) sh.center[2];
    double dx = ray.d.x, dy = ray.d.y, dz = ray.d.z;
    double A = (dx * dx + dy * dy) + dz * dz;
    double B = 2.0 * ((ox * dx + oy * dy) + oz * dz);
    double C = ((ox * ox + oy * oy) + oz * oz) - (double) sh.radius * (double) sh.radius;
    bool linear_case = A == 0.0, valid_linear = linear_case && B != 0.0;
    double x0 = -C / B, x1 = x0;
    double discrim = fma(B, B, -(4.0 * A * C));
    bool valid_quadratic = !linear_case && discrim >= 0.0;
    if (valid_quadratic) {
        double sq = sqrt(discrim);
        double temp = -0.5 * (B + copysign(sq, B));
        double x0p = temp / A, x1p = C / temp;
        x0 = x0p < x1p ? x0p : x1p; x1 = x0p > x1p ? x0p : x1p;
    }
    bool found = valid_linear || valid_quadratic;
    bool out_bounds = !(x0 <= maxt && x1 >= mint), in_bounds = x0 < mint && x1 > maxt;
    t = x0 < mint ? (float) x1 : (float) x0;
    return found && !out_bounds && !in_bounds;
}

// ---- SurfaceInteraction (mesh.cpp:449-545, interaction.h:153-156,571-596) -------------------
struct SI {
    float3 p, n, sh_n, sh_s, sh_t, wi;
    int bsdf, emitter;
    float2 uv;                   // only maintained by si_fill<true>
};

B2_D float3 coordinate_system_s(float3 n) {        // include/mitsuba/core/vector.h:116-136 (first vector)
    float sign = b2_sign(n.z);
    float a = -b2_rcp(sign + n.z);
    float b = (n.x * n.y) * a;
    return v3(b2_mulsign(b2_sqr(n.x) * a, n.z) + 1.f, b2_mulsign(b, n.z), b2_mulsign_neg(n.x, n.z));
}
B2_D float3 to_local(const SI &si, float3 v) { return v3(dot(v, si.sh_s), dot(v, si.sh_t), dot(v, si.sh_n)); }
B2_D float3 to_world(const SI &si, float3 v) { return (si.sh_s * v.x + si.sh_t * v.y) + si.sh_n * v.z; }   // frame.h:35-37

// Transform::transform_affine for points / vectors / normals (include/mitsuba/core/transform.h:88-151);
// M = 3x4 row-major; normals use the inverse transpose = columns of the inverse matrix
B2_D float3 xf_point(const float *M, float3 p) {
    return v3(b2_fma(M[2], p.z, b2_fma(M[1], p.y, b2_fma(M[0], p.x, M[3]))),
              b2_fma(M[6], p.z, b2_fma(M[5], p.y, b2_fma(M[4], p.x, M[7]))),
              b2_fma(M[10], p.z, b2_fma(M[9], p.y, b2_fma(M[8], p.x, M[11]))));
}
B2_D float3 xf_vector(const float *M, float3 d) {
    return v3(b2_fma(M[2], d.z, b2_fma(M[1], d.y, M[0] * d.x)),
              b2_fma(M[6], d.z, b2_fma(M[5], d.y, M[4] * d.x)),
              b2_fma(M[10], d.z, b2_fma(M[9], d.y, M[8] * d.x)));
}
B2_D float3 xf_normal(const float *inv, float3 n) {
    return v3(b2_fma(inv[8], n.z, b2_fma(inv[4], n.y, inv[0] * n.x)),
              b2_fma(inv[9], n.z, b2_fma(inv[5], n.y, inv[1] * n.x)),
              b2_fma(inv[10], n.z, b2_fma(inv[6], n.y, inv[2] * n.x)));
}

template <bool TEX = false>
B2_D void si_fill(const Tables &sc, uint32_t slot, float t, float u, float v, float3 ray_o, float3 ray_d, int inst, SI &si) {
    float4 a0 = B2_LD(&sc.tri_pos[3 * slot]), a1 = B2_LD(&sc.tri_pos[3 * slot + 1]), a2 = B2_LD(&sc.tri_pos[3 * slot + 2]);
    float4 m0 = B2_LD(&sc.tri_nrm[3 * slot]), m1 = B2_LD(&sc.tri_nrm[3 * slot + 1]), m2 = B2_LD(&sc.tri_nrm[3 * slot + 2]);
    si.bsdf = (int) __float_as_uint(m0.w);
    si.emitter = (int) __float_as_uint(m1.w);
    if constexpr (TEX) if (__float_as_uint(m2.w) >= 2u) {           // analytic shape: tri_nrm[..+2].w = 2 + type, tri_pos[..+2].w = index
        // Rectangle::compute_surface_interaction (rectangle.cpp:174-207) / Sphere:: (sphere.cpp:333-395); (u, v) = prim_uv
        const DShape &sh = sc.shapes[__float_as_uint(a2.w)];
        float3 hitp = v3(b2_fma(ray_d.x, t, ray_o.x), b2_fma(ray_d.y, t, ray_o.y), b2_fma(ray_d.z, t, ray_o.z));
        float3 dp_du;
        if (sh.type == B2_SHAPE_RECTANGLE) {
            si.p = hitp;
            si.n = v3(sh.n[0], sh.n[1], sh.n[2]); si.sh_n = si.n;
            dp_du = v3(sh.dp_du[0], sh.dp_du[1], sh.dp_du[2]);
            si.uv = make_float2(b2_fma(u, .5f, .5f), b2_fma(v, .5f, .5f));
        } else {
            float3 c = v3(sh.center[0], sh.center[1], sh.center[2]);
            float3 n = normalize(hitp - c);
            si.p = v3(b2_fma(n.x, sh.radius, c.x), b2_fma(n.y, sh.radius, c.y), b2_fma(n.z, sh.radius, c.z));
            float3 local = xf_point(sh.to_object, si.p);
            float theta = 2.f * b2_asin(.5f * sqrtf(b2_sqr(local.x) + b2_sqr(local.y) + b2_sqr(local.z - b2_sign(local.z))));
            if (!(local.z >= 0.f)) theta = 3.14159265358979323846f - theta;          // unit_angle_z
            float phi = b2_atan2(local.y, local.x);
            if (phi < 0.f) phi += 2.f * B2_PI;
            si.uv = make_float2(phi * (.5f * B2_INV_PI), theta * B2_INV_PI);
            dp_du = xf_vector(sh.to_world, v3(-local.y, local.x, 0.f)) * (2.f * B2_PI);
            if (sh.flip_normals) n = -n;
            si.sh_n = n; si.n = n;
        }
        si.sh_s = normalize(fnmadd(si.sh_n, dot(si.sh_n, dp_du), dp_du));
        si.sh_t = cross(si.sh_n, si.sh_s);
        si.wi = to_local(si, -ray_d);
        return;
    }
    float3 p0 = v3(a0), p1 = v3(a1), p2 = v3(a2);
    float b1 = u, b2 = v, b0 = 1.f - b1 - b2;
    float3 dp0 = p1 - p0, dp1 = p2 - p0;
    si.p = (p0 * b0 + p1 * b1) + p2 * b2;                       // mesh.cpp:484
    si.n = normalize(cross(dp0, dp1));                           // mesh.cpp:487
    float3 dp_du = coordinate_system_s(si.n);                    // mesh.cpp:491
    if constexpr (TEX) {
        si.uv = make_float2(b1, b2);                             // mesh.cpp:490
        if (sc.tri_uv) {
            float4 t0 = B2_LD(&sc.tri_uv[2 * slot]), t1 = B2_LD(&sc.tri_uv[2 * slot + 1]);
            if (__float_as_uint(t1.z)) {                         // has_vertex_texcoords(), mesh.cpp:492-511
                si.uv = make_float2((t0.x * b0 + t0.z * b1) + t1.x * b2, (t0.y * b0 + t0.w * b1) + t1.y * b2);
                float duv0x = t0.z - t0.x, duv0y = t0.w - t0.y, duv1x = t1.x - t0.x, duv1y = t1.y - t0.y;
                float det = b2_fma(duv0x, duv1y, -(duv0y * duv1x)), inv_det = b2_rcp(det);
                if (det != 0.f)                                  // fmsub(duv1.y, dp0, duv0.y * dp1) * inv_det
                    dp_du = v3(b2_fma(duv1y, dp0.x, -(duv0y * dp1.x)) * inv_det, b2_fma(duv1y, dp0.y, -(duv0y * dp1.y)) * inv_det,
                               b2_fma(duv1y, dp0.z, -(duv0y * dp1.z)) * inv_det);
            }
        }
    }
    if (__float_as_uint(m2.w) == 1u)                             // mesh.cpp:514-519
        si.sh_n = normalize((v3(m0) * b0 + v3(m1) * b1) + v3(m2) * b2);
    else
        si.sh_n = si.n;
    if (inst >= 0) {                                             // Instance::compute_surface_interaction, instance.cpp:124-142
        const DInstance &in = sc.instances[inst];
        si.p = xf_point(in.to_world, si.p);
        si.n = normalize(xf_normal(in.to_object, si.n));
        si.sh_n = normalize(xf_normal(in.to_object, si.sh_n));
        dp_du = xf_vector(in.to_world, dp_du);
    }
    si.sh_s = normalize(fnmadd(si.sh_n, dot(si.sh_n, dp_du), dp_du));   // interaction.h:153-156
    si.sh_t = cross(si.sh_n, si.sh_s);
    si.wi = to_local(si, -ray_d);                                // interaction.h:591
}

// ---- warps (include/mitsuba/core/warp.h) -----------------------------------------------------
B2_D void square_to_uniform_disk_concentric(float sx, float sy, float &ox, float &oy) {   // warp.h:53-90
    float x = b2_fma(2.f, sx, -1.f), y = b2_fma(2.f, sy, -1.f);
    bool is_zero = x == 0.f && y == 0.f, q13 = fabsf(x) < fabsf(y);
    float r = q13 ? y : x, rp = q13 ? x : y;
    float phi = ((.25f * B2_PI) * rp) / r;
    if (q13) phi = .5f * B2_PI - phi;
    if (is_zero) phi = 0.f;
    float s, c; b2_sincos(phi, &s, &c);
    ox = r * c; oy = r * s;
}
B2_D float3 square_to_cosine_hemisphere(float sx, float sy) {   // warp.h:325-333
    float px, py; square_to_uniform_disk_concentric(sx, sy, px, py);
    float z = b2_safe_sqrt(1.f - (px * px + py * py));
    return v3(px, py, z);
}

// ---- Fresnel (include/mitsuba/render/fresnel.h) -----------------------------------------------
B2_D void fresnel(float cos_theta_i, float eta, float &r_out, float &cos_theta_t_out, float &eta_it_out, float &eta_ti_out) {
    bool outside = cos_theta_i >= 0.f;                           // fresnel.h:34-70
    float rcp_eta = b2_rcp(eta), eta_it = outside ? eta : rcp_eta, eta_ti = outside ? rcp_eta : eta;
    float cos_theta_t_sqr = b2_fma(-b2_fma(-cos_theta_i, cos_theta_i, 1.f), eta_ti * eta_ti, 1.f);
    float cos_theta_i_abs = fabsf(cos_theta_i);
    float cos_theta_t_abs = b2_safe_sqrt(cos_theta_t_sqr);
    bool index_matched = eta == 1.f, special_case = index_matched || cos_theta_i_abs == 0.f;
    float r_sc = index_matched ? 0.f : 1.f;
    float a_s = b2_fma(-eta_it, cos_theta_t_abs, cos_theta_i_abs) / b2_fma(eta_it, cos_theta_t_abs, cos_theta_i_abs);
    float a_p = b2_fma(-eta_it, cos_theta_i_abs, cos_theta_t_abs) / b2_fma(eta_it, cos_theta_i_abs, cos_theta_t_abs);
    float r = .5f * (b2_sqr(a_s) + b2_sqr(a_p));
    if (special_case) r = r_sc;
    r_out = r; cos_theta_t_out = b2_mulsign_neg(cos_theta_t_abs, cos_theta_i);
    eta_it_out = eta_it; eta_ti_out = eta_ti;
}
B2_D float fresnel_conductor(float cos_theta_i, float eta_r, float eta_i) {   // fresnel.h:92-116
    float cos_theta_i_2 = cos_theta_i * cos_theta_i, sin_theta_i_2 = 1.f - cos_theta_i_2,
          sin_theta_i_4 = sin_theta_i_2 * sin_theta_i_2;
    float temp_1 = eta_r * eta_r - eta_i * eta_i - sin_theta_i_2,
          a_2_pb_2 = b2_safe_sqrt(temp_1 * temp_1 + 4.f * eta_i * eta_i * eta_r * eta_r),
          a = b2_safe_sqrt(.5f * (a_2_pb_2 + temp_1));
    float term_1 = a_2_pb_2 + cos_theta_i_2, term_2 = 2.f * cos_theta_i * a;
    float r_s = (term_1 - term_2) / (term_1 + term_2);
    float term_3 = a_2_pb_2 * cos_theta_i_2 + sin_theta_i_4, term_4 = term_2 * sin_theta_i_2;
    float r_p = r_s * (term_3 - term_4) / (term_3 + term_4);
    return .5f * (r_s + r_p);
}

// ---- microfacet distribution (include/mitsuba/render/microfacet.h) ----------------------------
struct MF { int type; float au, av; };
B2_D MF mf_make(const B2Bsdf &b) {       // configure(): alpha = max(alpha, 1e-4)
    MF m; m.type = b.distribution; m.au = b2_max(b.alpha_u, 1e-4f); m.av = b2_max(b.alpha_v, 1e-4f); return m;
}
B2_D float mf_eval(const MF &d, float3 m) {     // microfacet.h:184-203
    float alpha_uv = d.au * d.av, cos_theta = m.z, cos_theta_2 = b2_sqr(cos_theta), result;
    if (d.type == B2_MF_BECKMANN)
        result = b2_exp(-(b2_sqr(m.x / d.au) + b2_sqr(m.y / d.av)) / cos_theta_2) / (B2_PI * alpha_uv * b2_sqr(cos_theta_2));
    else
        result = b2_rcp(B2_PI * alpha_uv * b2_sqr(b2_sqr(m.x / d.au) + b2_sqr(m.y / d.av) + b2_sqr(m.z)));
    return result * cos_theta > 1e-20f ? result : 0.f;
}
B2_D float mf_smith_g1(const MF &d, float3 v, float3 m) {   // microfacet.h:331-355
    float xy_alpha_2 = b2_sqr(d.au * v.x) + b2_sqr(d.av * v.y), tan_theta_alpha_2 = xy_alpha_2 / b2_sqr(v.z), result;
    if (d.type == B2_MF_BECKMANN) {
        float a = b2_rsqrt(tan_theta_alpha_2), a_sqr = b2_sqr(a);
        result = a >= 1.6f ? 1.f : (3.535f * a + 2.181f * a_sqr) / (1.f + 2.276f * a + 2.577f * a_sqr);
    } else
        result = 2.f / (1.f + sqrtf(1.f + tan_theta_alpha_2));
    if (xy_alpha_2 == 0.f) result = 1.f;
    if (dot(v, m) * v.z <= 0.f) result = 0.f;
    return result;
}
B2_D void mf_sample_visible_11(const MF &d, float cos_theta_i, float sx, float sy, float &ox, float &oy) {
    if (d.type == B2_MF_BECKMANN) {                               // microfacet.h:359-394
        float tan_theta_i = b2_safe_sqrt(b2_fma(-cos_theta_i, cos_theta_i, 1.f)) / cos_theta_i;
        float cot_theta_i = b2_rcp(tan_theta_i);
        float maxval = b2_erf(cot_theta_i);
        sx = b2_max(b2_min(sx, 1.f - 1e-6f), 1e-6f);
        sy = b2_max(b2_min(sy, 1.f - 1e-6f), 1e-6f);
        float x = maxval - (maxval + 1.f) * b2_erf(sqrtf(-b2_log(sx)));
        sx *= 1.f + maxval + B2_INV_SQRT_PI * tan_theta_i * b2_exp(-b2_sqr(cot_theta_i));
        #pragma unroll 1
        for (int i = 0; i < 3; ++i) {
            float slope = b2_erfinv(x),
                  value = 1.f + x + B2_INV_SQRT_PI * tan_theta_i * b2_exp(-b2_sqr(slope)) - sx,
                  derivative = 1.f - slope * tan_theta_i;
            x -= value / derivative;
        }
        ox = b2_erfinv(x); oy = b2_erfinv(b2_fma(2.f, sy, -1.f));
    } else {                                                       // microfacet.h:395-410
        float px, py; square_to_uniform_disk_concentric(sx, sy, px, py);
        float s = .5f * (1.f + cos_theta_i);
        float a0 = b2_safe_sqrt(1.f - b2_sqr(px));
        py = b2_fma(py, s, b2_fma(-a0, s, a0));                    // enoki::lerp
        float x = px, y = py, z = b2_safe_sqrt(1.f - (px * px + py * py));
        float sin_theta_i = b2_safe_sqrt(1.f - b2_sqr(cos_theta_i));
        float norm = b2_rcp(b2_fma(sin_theta_i, y, cos_theta_i * z));
        ox = b2_fma(cos_theta_i, y, -(sin_theta_i * z)) * norm; oy = x * norm;
    }
}
B2_D float3 mf_sample(const MF &d, float3 wi, float sx, float sy, float &pdf) {   // microfacet.h:283-316
    float3 wi_p = normalize(v3(d.au * wi.x, d.av * wi.y, wi.z));
    float sin_theta_2 = b2_fma(-wi_p.z, wi_p.z, 1.f), inv_sin_theta = b2_rsqrt(sin_theta_2);   // frame.h sincos_phi
    float sin_phi = wi_p.y * inv_sin_theta, cos_phi = wi_p.x * inv_sin_theta;
    if (fabsf(sin_theta_2) <= 4.f * B2_EPS) { sin_phi = 0.f; cos_phi = 1.f; }
    float slx, sly; mf_sample_visible_11(d, wi_p.z, sx, sy, slx, sly);
    float s0 = b2_fma(cos_phi, slx, -(sin_phi * sly)) * d.au;
    float s1 = b2_fma(sin_phi, slx, cos_phi * sly) * d.av;
    float3 m = normalize(v3(-s0, -s1, 1.f));
    pdf = mf_eval(d, m) * mf_smith_g1(d, wi, m) * fabsf(dot(wi, m)) / wi.z;
    return m;
}

B2_D float3 rgb3(const float *p) { return v3(p[0], p[1], p[2]); }

// ---- textures (src/textures/checkerboard.cpp:46-63, bitmap.cpp:386-474) ----------------------
B2_D int tex_wrap(const B2Texture &t, int value, int res) {       // bitmap.cpp:386-401
    if (t.wrap == B2_WRAP_CLAMP) return max(0, min(value, res - 1));
    int div = value / res, mod = value - div * res;               // enoki::divisor: truncating division
    if (mod < 0) mod += res;
    if (t.wrap == B2_WRAP_MIRROR && !(((div & 1) == 0) != (value < 0))) mod = res - 1 - mod;
    return mod;
}
B2_D float3 tex_fetch(const B2Texture &t, int x, int y) {
    const float *p = t.data + (size_t) (x + y * t.width) * (size_t) t.channels;
    float a = __ldg(p);
    return t.channels == 1 ? v3(a, a, a) : v3(a, __ldg(p + 1), __ldg(p + 2));
}
B2_D float3 texture_eval(const B2Texture &t, float2 si_uv) {
    // Transform3f::transform_affine: result = col2; result = fmadd(col_i, uv_i, result)
    float u = b2_fma(t.to_uv[1], si_uv.y, b2_fma(t.to_uv[0], si_uv.x, t.to_uv[2]));
    float v = b2_fma(t.to_uv[4], si_uv.y, b2_fma(t.to_uv[3], si_uv.x, t.to_uv[5]));
    if (t.type == B2_TEX_CHECKERBOARD) {                          // checkerboard.cpp:46-63
        bool mx = u - floorf(u) > .5f, my = v - floorf(v) > .5f;
        return mx == my ? rgb3(t.color0) : rgb3(t.color1);
    }
    if (t.filter == B2_TEXFILTER_BILINEAR) {                      // bitmap.cpp:412-457
        u = b2_fma(u, (float) t.width, -.5f); v = b2_fma(v, (float) t.height, -.5f);
        int ix = (int) floorf(u), iy = (int) floorf(v);
        float w1x = u - (float) ix, w1y = v - (float) iy, w0x = 1.f - w1x, w0y = 1.f - w1y;
        int x0 = tex_wrap(t, ix, t.width), x1 = tex_wrap(t, ix + 1, t.width), y0 = tex_wrap(t, iy, t.height), y1 = tex_wrap(t, iy + 1, t.height);
        float3 v00 = tex_fetch(t, x0, y0), v10 = tex_fetch(t, x1, y0), v01 = tex_fetch(t, x0, y1), v11 = tex_fetch(t, x1, y1);
        float3 v0 = v3(b2_fma(w0x, v00.x, w1x * v10.x), b2_fma(w0x, v00.y, w1x * v10.y), b2_fma(w0x, v00.z, w1x * v10.z));
        float3 v1 = v3(b2_fma(w0x, v01.x, w1x * v11.x), b2_fma(w0x, v01.y, w1x * v11.y), b2_fma(w0x, v01.z, w1x * v11.z));
        return v3(b2_fma(w0y, v0.x, w1y * v1.x), b2_fma(w0y, v0.y, w1y * v1.y), b2_fma(w0y, v0.z, w1y * v1.z));
    }
    u *= (float) t.width; v *= (float) t.height;                  // bitmap.cpp:458-473
    return tex_fetch(t, tex_wrap(t, (int) floorf(u), t.width), tex_wrap(t, (int) floorf(v), t.height));
}
// 'reflectance' of the BSDF at si: the constant, or its texture evaluated at si.uv (Texture::eval(si) inside BSDF::eval / sample)
template <uint32_t F> B2_D float3 bsdf_reflectance(const Tables &tb, const B2Bsdf &b, const SI &si) {
    if constexpr ((F & F_TEX) != 0)
        if (b.reflectance_texture > 0) return texture_eval(tb.textures[b.reflectance_texture - 1], si.uv);
    return v3(b.reflectance[0], b.reflectance[1], b.reflectance[2]);
}

// ---- BSDFs (src/bsdfs/{diffuse,dielectric,roughconductor,conductor,plastic}.cpp) --------------
// F_DIELECTRIC marks scenes with any of the BSDFs that have a delta lobe (dielectric, conductor, plastic)
template <uint32_t F> B2_D bool bsdf_is_smooth(const B2Bsdf &b) {     // has_flag(flags, Smooth), path.cpp:155
    return !(F & F_DIELECTRIC) || (b.type != B2_BSDF_DIELECTRIC && b.type != B2_BSDF_CONDUCTOR);
}
// The BSDF record that applies at si. TwoSidedBRDF (twosided.cpp:94-186): the front record for cos_theta(wi) > 0, the
// back record with wi.z mirrored (and the resulting wo.z mirrored by the caller: `flip`) for cos_theta(wi) < 0; at exactly
// 0 every nestable BRDF evaluates to zero by itself. `smooth` = has_flag(flags, Smooth) over the union of both sides.
// si.bsdf is updated to the side record's index, si.wi.z is mirrored in place (nothing after the BSDF phase reads wi).
template <uint32_t F> B2_D const B2Bsdf &bsdf_side(const Tables &tb, SI &si, bool &flip, bool &smooth) {
    const B2Bsdf *b = &tb.bsdfs[si.bsdf];
    smooth = bsdf_is_smooth<F>(*b);
    flip = false;
    if constexpr ((F & F_TEX) != 0) if (b->twosided_back > 0) {
        const int back = b->twosided_back - 1;
        smooth = smooth || bsdf_is_smooth<F>(tb.bsdfs[back]);
        if (si.wi.z < 0.f) { si.bsdf = back; b = &tb.bsdfs[back]; flip = true; si.wi.z = -si.wi.z; }
    }
    return *b;
}
// SmoothPlastic: diffuse_reflectance / (1 - fdr_int), optional nonlinear variant (plastic.cpp:231-232, 262-263)
B2_D float3 plastic_diffuse(const B2Bsdf &b, float3 d) {
    if (b.nonlinear)
        return v3(d.x / (1.f - d.x * b.fdr_int), d.y / (1.f - d.y * b.fdr_int), d.z / (1.f - d.z * b.fdr_int));
    return d * (1.f / (1.f - b.fdr_int));
}
B2_D float fresnel_r(float cos_theta_i, float eta) {
    float r, ct, eit, eti; fresnel(cos_theta_i, eta, r, ct, eit, eti); return r;
}

template <uint32_t F> B2_D float3 bsdf_eval(const B2Bsdf &b, const SI &si, float3 wo, float3 refl) {
    float cti = si.wi.z, cto = wo.z;
    float3 zero = v3(0.f, 0.f, 0.f);
    switch (b.type) {
        case B2_BSDF_DIFFUSE: case B2_BSDF_NULL: {                 // diffuse.cpp:104-119
            if (!(cti > 0.f && cto > 0.f)) return zero;
            return v3((refl.x * B2_INV_PI) * cto, (refl.y * B2_INV_PI) * cto, (refl.z * B2_INV_PI) * cto);
        }
        case B2_BSDF_ROUGHCONDUCTOR: if constexpr ((F & F_ROUGH) != 0) {   // roughconductor.cpp:277-343
            if (!(cti > 0.f && cto > 0.f)) return zero;
            float3 H = normalize(wo + si.wi);
            MF d = mf_make(b);
            float D = mf_eval(d, H);
            if (D == 0.f) return zero;
            float G = mf_smith_g1(d, si.wi, H) * mf_smith_g1(d, wo, H);
            float result = D * G / (4.f * cti);
            float c = dot(si.wi, H);
            float3 Fr = v3(fresnel_conductor(c, b.cond_eta[0], b.cond_k[0]), fresnel_conductor(c, b.cond_eta[1], b.cond_k[1]),
                           fresnel_conductor(c, b.cond_eta[2], b.cond_k[2]));
            float3 res = v3(result * b.specular_reflectance[0], result * b.specular_reflectance[1], result * b.specular_reflectance[2]);
            return Fr * res;
        } else return zero;
        case B2_BSDF_ROUGHDIELECTRIC: if constexpr ((F & F_ROUGH) != 0) {  // roughdielectric.cpp:312-384
            if (cti == 0.f) return zero;
            bool reflect = cti * cto > 0.f;
            float eta = cti > 0.f ? b.eta : b2_rcp(b.eta), inv_eta = cti > 0.f ? b2_rcp(b.eta) : b.eta;
            float3 m = normalize(si.wi + wo * (reflect ? 1.f : eta));
            if (__float_as_uint(m.z) & 0x80000000u) m = -m;             // mulsign(m, cos_theta(m))
            MF d = mf_make(b);
            float D = mf_eval(d, m);
            float Fr = fresnel_r(dot(si.wi, m), b.eta);
            float G = mf_smith_g1(d, si.wi, m) * mf_smith_g1(d, wo, m);
            if (reflect) {
                float value = Fr * D * G / (4.f * fabsf(cti));
                return v3(value * b.specular_reflectance[0], value * b.specular_reflectance[1], value * b.specular_reflectance[2]);
            }
            float scale = b2_sqr(inv_eta);
            float dwm = dot(si.wi, m), dom = dot(wo, m);
            float value = fabsf((scale * (1.f - Fr) * D * G * eta * eta * dwm * dom) / (cti * b2_sqr(dwm + eta * dom)));
            return v3(value * b.specular_transmittance[0], value * b.specular_transmittance[1], value * b.specular_transmittance[2]);
        } else return zero;
        case B2_BSDF_PLASTIC: if constexpr ((F & F_DIELECTRIC) != 0) {     // plastic.cpp:242-266
            if (!(cti > 0.f && cto > 0.f)) return zero;
            float f_i = fresnel_r(cti, b.eta), f_o = fresnel_r(cto, b.eta);
            float3 diff = plastic_diffuse(b, refl);
            float s = (B2_INV_PI * cto) * b.inv_eta_2 * (1.f - f_i) * (1.f - f_o);
            return diff * s;
        } else return zero;
        default: return zero;                                       // dielectric.cpp:312-315, conductor.cpp:242-245
    }
}
template <uint32_t F> B2_D float bsdf_pdf(const B2Bsdf &b, const SI &si, float3 wo) {
    float cti = si.wi.z, cto = wo.z;
    switch (b.type) {
        case B2_BSDF_DIFFUSE: case B2_BSDF_NULL:                    // diffuse.cpp:121-135
            return (cti > 0.f && cto > 0.f) ? B2_INV_PI * cto : 0.f;
        case B2_BSDF_ROUGHCONDUCTOR: if constexpr ((F & F_ROUGH) != 0) {   // roughconductor.cpp:345-382
            float3 m = normalize(wo + si.wi);
            if (!(cti > 0.f && cto > 0.f && dot(si.wi, m) > 0.f && dot(wo, m) > 0.f)) return 0.f;
            MF d = mf_make(b);
            return mf_eval(d, m) * mf_smith_g1(d, si.wi, m) / (4.f * cti);
        } else return 0.f;
        case B2_BSDF_ROUGHDIELECTRIC: if constexpr ((F & F_ROUGH) != 0) {  // roughdielectric.cpp:386-452
            if (cti == 0.f) return 0.f;
            bool reflect = cti * cto > 0.f;
            float eta = cti > 0.f ? b.eta : b2_rcp(b.eta);
            float3 m = normalize(si.wi + wo * (reflect ? 1.f : eta));
            if (__float_as_uint(m.z) & 0x80000000u) m = -m;
            if (!(dot(si.wi, m) * cti > 0.f && dot(wo, m) * cto > 0.f)) return 0.f;
            float dwh_dwo = reflect ? b2_rcp(4.f * dot(wo, m)) : (eta * eta * dot(wo, m)) / b2_sqr(dot(si.wi, m) + eta * dot(wo, m));
            MF d = mf_make(b);
            float3 wi_s = cti < 0.f ? -si.wi : si.wi;                   // mulsign(wi, cos_theta(wi))
            float prob = mf_eval(d, m) * mf_smith_g1(d, wi_s, m) * fabsf(dot(wi_s, m)) / wi_s.z;   // distr.pdf, microfacet.h:214-223
            float Fr = fresnel_r(dot(si.wi, m), b.eta);
            prob *= reflect ? Fr : 1.f - Fr;
            return prob * fabsf(dwh_dwo);
        } else return 0.f;
        case B2_BSDF_PLASTIC: if constexpr ((F & F_DIELECTRIC) != 0) {     // plastic.cpp:268-293
            if (!(cti > 0.f && cto > 0.f)) return 0.f;
            float f_i = fresnel_r(cti, b.eta);
            float prob_specular = f_i * b.specular_sampling_weight;
            float prob_diffuse = (1.f - f_i) * (1.f - b.specular_sampling_weight);
            prob_diffuse = prob_diffuse / (prob_specular + prob_diffuse);
            return (B2_INV_PI * cto) * prob_diffuse;
        } else return 0.f;
        default: return 0.f;
    }
}
struct BS { float3 wo; float pdf, eta; bool delta; };
template <uint32_t F> B2_D float3 bsdf_sample(const B2Bsdf &b, const SI &si, float sample1, float s2x, float s2y, BS &bs, float3 refl) {
    float3 zero = v3(0.f, 0.f, 0.f);
    bs.wo = zero; bs.pdf = 0.f; bs.eta = 0.f; bs.delta = false;
    float cti = si.wi.z;
    switch (b.type) {
        case B2_BSDF_DIFFUSE: case B2_BSDF_NULL: {                  // diffuse.cpp:78-102
            if (!(cti > 0.f)) return zero;
            bs.wo = square_to_cosine_hemisphere(s2x, s2y);
            bs.pdf = B2_INV_PI * bs.wo.z;
            bs.eta = 1.f;
            if (!(bs.pdf > 0.f)) return zero;
            return refl;
        }
        case B2_BSDF_DIELECTRIC: if constexpr ((F & F_DIELECTRIC) != 0) {   // dielectric.cpp:201-310
            float r_i, cos_theta_t, eta_it, eta_ti;
            fresnel(cti, b.eta, r_i, cos_theta_t, eta_it, eta_ti);
            float t_i = 1.f - r_i;
            bool selected_r = sample1 <= r_i;
            bs.pdf = selected_r ? r_i : t_i;
            bs.delta = true;
            float3 w = v3(1.f, 1.f, 1.f);
            if (selected_r) {
                bs.wo = v3(-si.wi.x, -si.wi.y, si.wi.z); bs.eta = 1.f;
                w = w * rgb3(b.specular_reflectance);
            } else {
                bs.wo = v3(-eta_ti * si.wi.x, -eta_ti * si.wi.y, cos_theta_t); bs.eta = eta_it;
                w = w * rgb3(b.specular_transmittance);
                w = w * b2_sqr(eta_ti);                              // dielectric.cpp:303-307
            }
            return w;
        } else return zero;
        case B2_BSDF_ROUGHCONDUCTOR: if constexpr ((F & F_ROUGH) != 0) {   // roughconductor.cpp:196-275
            if (!(cti > 0.f)) return zero;
            MF d = mf_make(b);
            float pdf; float3 m = mf_sample(d, si.wi, s2x, s2y, pdf);
            float dwm = 2.f * dot(si.wi, m);                         // reflect(wi, m) fresnel.h:281-283
            bs.wo = v3(b2_fma(m.x, dwm, -si.wi.x), b2_fma(m.y, dwm, -si.wi.y), b2_fma(m.z, dwm, -si.wi.z));
            bs.eta = 1.f;
            bool active = pdf != 0.f && bs.wo.z > 0.f;
            float weight = mf_smith_g1(d, bs.wo, m);
            bs.pdf = pdf / (4.f * dot(bs.wo, m));
            float c = dot(si.wi, m);
            float3 Fr = v3(fresnel_conductor(c, b.cond_eta[0], b.cond_k[0]), fresnel_conductor(c, b.cond_eta[1], b.cond_k[1]),
                           fresnel_conductor(c, b.cond_eta[2], b.cond_k[2]));
            float3 wv = v3(weight * b.specular_reflectance[0], weight * b.specular_reflectance[1], weight * b.specular_reflectance[2]);
            return active ? Fr * wv : zero;
        } else return zero;
        case B2_BSDF_ROUGHDIELECTRIC: if constexpr ((F & F_ROUGH) != 0) {  // roughdielectric.cpp:203-310
            bool active = cti != 0.f;
            MF d = mf_make(b);
            float3 wi_s = cti < 0.f ? -si.wi : si.wi;
            float pdf; float3 m = mf_sample(d, wi_s, s2x, s2y, pdf);
            active = active && pdf != 0.f;
            float Fr, cos_theta_t, eta_it, eta_ti;
            fresnel(dot(si.wi, m), b.eta, Fr, cos_theta_t, eta_it, eta_ti);
            bool selected_r = sample1 <= Fr && active;
            float3 weight = v3(1.f, 1.f, 1.f);
            pdf *= selected_r ? Fr : 1.f - Fr;
            bs.eta = selected_r ? 1.f : eta_it;
            float dwh_dwo = 0.f;
            float dwm = dot(si.wi, m);
            if (selected_r) {
                float s = 2.f * dwm;                                 // reflect(wi, m) = fmsub(m, 2 dot(wi, m), wi)
                bs.wo = v3(b2_fma(m.x, s, -si.wi.x), b2_fma(m.y, s, -si.wi.y), b2_fma(m.z, s, -si.wi.z));
                weight = weight * rgb3(b.specular_reflectance);
                dwh_dwo = b2_rcp(4.f * dot(bs.wo, m));
            } else if (active) {
                float s = b2_fma(dwm, eta_ti, cos_theta_t);          // refract(wi, m, cos_theta_t, eta_ti), fresnel.h:310-313
                bs.wo = v3(b2_fma(m.x, s, -(si.wi.x * eta_ti)), b2_fma(m.y, s, -(si.wi.y * eta_ti)), b2_fma(m.z, s, -(si.wi.z * eta_ti)));
                float fct = b2_sqr(eta_ti);
                weight = weight * v3(fct * b.specular_transmittance[0], fct * b.specular_transmittance[1], fct * b.specular_transmittance[2]);
                dwh_dwo = (b2_sqr(bs.eta) * dot(bs.wo, m)) / b2_sqr(dwm + bs.eta * dot(bs.wo, m));
            }
            weight = weight * mf_smith_g1(d, bs.wo, m);
            bs.pdf = pdf * fabsf(dwh_dwo);
            return active ? weight : zero;
        } else return zero;
        case B2_BSDF_CONDUCTOR: if constexpr ((F & F_DIELECTRIC) != 0) {   // conductor.cpp:197-240
            if (!(cti > 0.f)) return zero;
            bs.wo = v3(-si.wi.x, -si.wi.y, si.wi.z);                 // reflect(wi), fresnel.h:276-278
            bs.eta = 1.f; bs.pdf = 1.f; bs.delta = true;
            float3 Fr = v3(fresnel_conductor(cti, b.cond_eta[0], b.cond_k[0]), fresnel_conductor(cti, b.cond_eta[1], b.cond_k[1]),
                           fresnel_conductor(cti, b.cond_eta[2], b.cond_k[2]));
            return rgb3(b.specular_reflectance) * Fr;
        } else return zero;
        case B2_BSDF_PLASTIC: if constexpr ((F & F_DIELECTRIC) != 0) {     // plastic.cpp:181-240
            if (!(cti > 0.f)) return zero;
            float f_i = fresnel_r(cti, b.eta);
            float prob_specular = f_i * b.specular_sampling_weight;
            float prob_diffuse = (1.f - f_i) * (1.f - b.specular_sampling_weight);
            prob_specular = prob_specular / (prob_specular + prob_diffuse);
            prob_diffuse = 1.f - prob_specular;
            bs.eta = 1.f;
            if (sample1 < prob_specular) {
                bs.wo = v3(-si.wi.x, -si.wi.y, si.wi.z);
                bs.pdf = prob_specular; bs.delta = true;
                float value = f_i / bs.pdf;
                return v3(value * b.specular_reflectance[0], value * b.specular_reflectance[1], value * b.specular_reflectance[2]);
            }
            bs.wo = square_to_cosine_hemisphere(s2x, s2y);
            bs.pdf = prob_diffuse * (B2_INV_PI * bs.wo.z);
            float f_o = fresnel_r(bs.wo.z, b.eta);
            float3 value = plastic_diffuse(b, refl);
            float s = b.inv_eta_2 * (1.f - f_i) * (1.f - f_o) / prob_diffuse;
            return value * s;
        } else return zero;
        default: return zero;
    }
}

// ---- emitter sampling (area.cpp, shape.cpp:292-323, mesh.cpp:352-397, scene.cpp:164-231) ------
struct DS { float3 p, n, d; float pdf, dist; bool delta; };

B2_D uint32_t distr_sample(const float *cdf, uint32_t first, uint32_t last, float sum, float value) {   // distr_1d.h:144-155
    value *= sum;
    uint32_t start = first, end = last;
    uint32_t iterations = start < end ? (32 - __clz(end - start)) : 0;     // log2i(end-start)+1
    for (uint32_t i = 0; i < iterations; ++i) {                              // array_utils.h:153-170
        uint32_t middle = (start + end) >> 1;
        bool cond = B2_LD(&cdf[middle]) < value;
        if (cond) start = (middle + 1 < end) ? middle + 1 : end; else end = middle;
    }
    return start;
}

B2_D float3 area_sample_direction(const Tables &sc, const DEmitter &em, float3 ref_p, float sx, float sy, DS &ds) {
    const float *cdf = sc.cdf + em.cdf_offset;
    uint32_t face = distr_sample(cdf, em.valid_first, em.valid_last, em.cdf_sum, sy);
    uint32_t slot = B2_LD(&sc.prim_to_slot[em.first_prim + face]);
    float4 a0 = B2_LD(&sc.tri_pos[3 * slot]), a1 = B2_LD(&sc.tri_pos[3 * slot + 1]), a2 = B2_LD(&sc.tri_pos[3 * slot + 2]);
    float3 p0 = v3(a0), e0 = v3(a1) - p0, e1 = v3(a2) - p0;
    float3 cr = cross(e0, e1);
    float area = 0.5f * sqrtf(dot(cr, cr));                      // mesh.h:108-116
    float pmf = area * em.cdf_norm;                                // distr_1d.h:194-204
    float cdfv = face > 0 ? B2_LD(&cdf[face - 1]) * em.cdf_norm : 0.f;
    sy = (sy - cdfv) / pmf;
    float t = b2_safe_sqrt(1.f - sx);                             // warp.h:152-156
    float bx = 1.f - t, by = t * sy;
    ds.p = (p0 + e0 * bx) + e1 * by;                               // mesh.cpp:370
    float pdf = em.cdf_norm;
    float4 m2 = B2_LD(&sc.tri_nrm[3 * slot + 2]);
    if (__float_as_uint(m2.w)) {                                   // mesh.cpp:386-394
        float4 m0 = B2_LD(&sc.tri_nrm[3 * slot]), m1 = B2_LD(&sc.tri_nrm[3 * slot + 1]);
        ds.n = normalize((v3(m0) * (1.f - bx - by) + v3(m1) * bx) + v3(m2) * by);
    } else
        ds.n = normalize(cr);
    ds.d = ds.p - ref_p;                                           // shape.cpp:298-306
    float dist_squared = dot(ds.d, ds.d);
    ds.dist = sqrtf(dist_squared);
    ds.d = ds.d * b2_rcp(ds.dist);
    float dp = fabsf(dot(ds.d, ds.n));
    pdf *= dp != 0.f ? dist_squared / dp : 0.f;
    ds.pdf = pdf;
    bool active = dot(ds.d, ds.n) < 0.f && pdf != 0.f;             // area.cpp:132
    if (!active) return v3(0.f, 0.f, 0.f);
    float ip = b2_rcp(pdf);
    return v3(em.radiance[0] * ip, em.radiance[1] * ip, em.radiance[2] * ip);
}

// ---- environment map (src/emitters/envmap.cpp, distr_2d.h:473-686, warp.h:359-434) ----------------
B2_D uint32_t lvl_index(uint32_t x, uint32_t y, uint32_t width) {            // distr_2d.h Level::index
    return ((x & 1u) | (((x & ~1u) | (y & 1u)) << 1)) + ((y & ~1u) * width);
}
B2_D float b2_lerp(float a, float b, float t) { return b2_fma(b, t, b2_fma(-a, t, a)); }   // enoki::lerp
B2_D float interval_to_linear(float v0, float v1, float sample) {            // warp.h:359-366
    if (fabsf(v0 - v1) > 1e-4f * (v0 + v1))
        return (v0 - b2_safe_sqrt(b2_lerp(b2_sqr(v0), b2_sqr(v1), sample))) / (v0 - v1);
    return sample;
}
B2_D float clamp01(float x) { return b2_min(b2_max(x, 0.f), 1.f); }

B2_D void env_warp_sample(const DEnv &env, float sx, float sy, float &u, float &v, float &pdf) {   // distr_2d.h:473-560
    sx = clamp01(sx); sy = clamp01(sy);
    uint32_t ox = 0, oy = 0;
    for (int l = env.n_levels - 2; l > 0; --l) {
        const float *lv = env.levels + env.level_off[l];
        ox <<= 1; oy <<= 1;
        uint32_t oi = lvl_index(ox, oy, env.level_w[l]);
        float4 q = __ldg((const float4 *) (lv + oi));              // 2x2 patches are contiguous and 16-byte aligned
        float v00 = q.x, v10 = q.y, v01 = q.z, v11 = q.w;
        sx = clamp01(sx); sy = clamp01(sy);
        float r0 = v00 + v10, r1 = v01 + v11;
        sy *= r0 + r1;
        bool mask = sy > r0;
        if (mask) { oy += 1u; sy -= r0; }
        sy /= mask ? r1 : r0;
        float c0 = mask ? v01 : v00, c1 = mask ? v11 : v10;
        sx *= c0 + c1;
        mask = sx > c0;
        if (mask) sx -= c0;
        sx /= mask ? c1 : c0;
        if (mask) ox += 1u;
    }
    uint32_t W = (uint32_t) env.width;
    const float *l0 = env.levels;
    uint32_t oi = ox + oy * W;
    float v00 = __ldg(l0 + oi), v10 = __ldg(l0 + oi + 1), v01 = __ldg(l0 + oi + W), v11 = __ldg(l0 + oi + W + 1);
    float r0 = v00 + v10, r1 = v01 + v11;                           // warp::square_to_bilinear, warp.h:393-407
    sy = interval_to_linear(r0, r1, sy);
    float c0 = b2_lerp(v00, v01, sy), c1 = b2_lerp(v10, v11, sy);
    sx = interval_to_linear(c0, c1, sx);
    pdf = b2_lerp(c0, c1, sx);
    float psx = 1.f / (float) (env.width - 1), psy = 1.f / (float) (env.height - 1);
    u = ((float) (int) ox + sx) * psx; v = ((float) (int) oy + sy) * psy;
}
B2_D float env_warp_eval(const DEnv &env, float px, float py) {               // distr_2d.h:651-686
    px = clamp01(px); py = clamp01(py);
    uint32_t W = (uint32_t) env.width, H = (uint32_t) env.height;
    px *= (float) (W - 1); py *= (float) (H - 1);
    uint32_t ox = (uint32_t) (int) px, oy = (uint32_t) (int) py;
    if (ox > W - 2) ox = W - 2;
    if (oy > H - 2) oy = H - 2;
    px -= (float) (int) ox; py -= (float) (int) oy;
    const float *l0 = env.levels;
    uint32_t oi = ox + oy * W;
    float v00 = __ldg(l0 + oi), v10 = __ldg(l0 + oi + 1), v01 = __ldg(l0 + oi + W), v11 = __ldg(l0 + oi + W + 1);
    return b2_lerp(b2_lerp(v00, v10, px), b2_lerp(v01, v11, px), py);        // warp.h:428-434
}
B2_D float3 env_eval_spectrum(const DEnv &env, float u, float v) {            // envmap.cpp:269-319 (rgb branch)
    uint32_t W = (uint32_t) env.width, H = (uint32_t) env.height;
    u *= (float) (W - 1u); v *= (float) (H - 1u);
    uint32_t px = (uint32_t) u, py = (uint32_t) v;
    if (px > W - 2u) px = W - 2u;
    if (py > H - 2u) py = H - 2u;
    float w1x = u - (float) px, w1y = v - (float) py, w0x = 1.f - w1x, w0y = 1.f - w1y;
    size_t idx = (size_t) px + (size_t) py * W;
    float4 v00 = __ldg(env.data + idx), v10 = __ldg(env.data + idx + 1), v01 = __ldg(env.data + idx + W), v11 = __ldg(env.data + idx + W + 1);
    float3 a0 = v3(b2_fma(w0x, v00.x, w1x * v10.x), b2_fma(w0x, v00.y, w1x * v10.y), b2_fma(w0x, v00.z, w1x * v10.z));
    float3 a1 = v3(b2_fma(w0x, v01.x, w1x * v11.x), b2_fma(w0x, v01.y, w1x * v11.y), b2_fma(w0x, v01.z, w1x * v11.z));
    return v3(b2_fma(w0y, a0.x, w1y * a1.x) * env.scale, b2_fma(w0y, a0.y, w1y * a1.y) * env.scale, b2_fma(w0y, a0.z, w1y * a1.z) * env.scale);
}
B2_D void env_dir_to_uv(const DEnv &env, float3 d_world, float3 &d, float &u, float &v) {   // envmap.cpp:137-144
    d = xf_vector(env.to_object, d_world);
    u = b2_atan2(d.x, -d.z) * (0.5f * B2_INV_PI); v = b2_safe_acos(d.y) * B2_INV_PI;
    u -= floorf(u); v -= floorf(v);
}
B2_D float3 env_eval(const DEnv &env, float3 ray_d) {                         // envmap.cpp:134-147
    if (env.type == B2_EMITTER_CONSTANT) return v3(env.radiance[0], env.radiance[1], env.radiance[2]);   // constant.cpp:54-57
    float3 d; float u, v; env_dir_to_uv(env, ray_d, d, u, v);
    return env_eval_spectrum(env, u, v);
}
B2_D float env_inv_sin_theta(float3 d) { return b2_rsqrt(b2_max(b2_sqr(d.x) + b2_sqr(d.z), b2_sqr(B2_EPS))); }
B2_D float env_pdf_direction(const DEnv &env, float3 d_world) {               // envmap.cpp:192-208
    if (env.type == B2_EMITTER_CONSTANT) return .25f * B2_INV_PI;             // square_to_uniform_sphere_pdf, constant.cpp:110-114
    float3 d; float u, v; env_dir_to_uv(env, d_world, d, u, v);
    return env_warp_eval(env, u, v) * env_inv_sin_theta(d) * (1.f / (2.f * b2_sqr(B2_PI)));
}
B2_D float3 square_to_uniform_sphere(float sx, float sy);
B2_D float3 env_sample_direction(const DEnv &env, float3 ref_p, float sx, float sy, DS &ds) {   // envmap.cpp:156-190
    if (env.type == B2_EMITTER_CONSTANT) {                             // ConstantBackgroundEmitter::sample_direction, constant.cpp:83-108
        float3 dd = square_to_uniform_sphere(sx, sy);
        float dist0 = 2.f * env.radius;
        ds.p = ref_p + dd * dist0; ds.n = -dd;
        ds.pdf = .25f * B2_INV_PI;
        ds.d = dd; ds.dist = dist0;
        float ip = b2_rcp(ds.pdf);
        return v3(env.radiance[0] * ip, env.radiance[1] * ip, env.radiance[2] * ip);
    }
    float u, v, pdf; env_warp_sample(env, sx, sy, u, v, pdf);
    float theta = v * B2_PI, phi = u * (2.f * B2_PI);
    float st, ct, sp, cp; b2_sincos(theta, &st, &ct); b2_sincos(phi, &sp, &cp);
    float3 d0 = v3(cp * st, sp * st, ct);                            // math::sphdir
    float3 d = v3(d0.y, d0.z, -d0.x);
    float dist = 2.f * env.radius;
    float inv_sin_theta = env_inv_sin_theta(d);
    d = xf_vector(env.to_world, d);
    ds.p = ref_p + d * dist; ds.n = -d;
    ds.pdf = pdf > 0.f ? pdf * inv_sin_theta * (1.f / (2.f * b2_sqr(B2_PI))) : 0.f;
    ds.d = d; ds.dist = dist;
    return env_eval_spectrum(env, u, v) * b2_rcp(ds.pdf);
}

// Rectangle: Shape::sample_direction (shape.cpp:292-323) over Rectangle::sample_position (rectangle.cpp:113-128);
// Sphere::sample_direction (sphere.cpp:174-253); then AreaLight::sample_direction (area.cpp:118-141)
B2_D float3 coordinate_system_t(float3 n) {        // vector.h:116-136 (second vector)
    float sign = b2_sign(n.z);
    float a = -b2_rcp(sign + n.z);
    float b = (n.x * n.y) * a;
    return v3(b, sign + b2_sqr(n.y) * a, -n.y);
}
B2_D float3 square_to_uniform_sphere(float sx, float sy) {      // warp.h:255-260
    float z = b2_fma(-2.f, sy, 1.f), r = b2_safe_sqrt(b2_fma(-z, z, 1.f));
    float sn, cs; b2_sincos(2.f * B2_PI * sx, &sn, &cs);
    return v3(r * cs, r * sn, z);
}
B2_D float3 shape_sample_direction(const Tables &sc, const DEmitter &em, float3 ref_p, float sx, float sy, DS &ds) {
    const DShape &sh = sc.shapes[em.shape - 1];
    if (sh.type == B2_SHAPE_RECTANGLE) {
        ds.p = xf_point(sh.to_world, v3(sx * 2.f - 1.f, sy * 2.f - 1.f, 0.f));
        ds.n = v3(sh.n[0], sh.n[1], sh.n[2]);
        float pdf = sh.inv_area;
        ds.d = ds.p - ref_p;
        float dist_squared = dot(ds.d, ds.d);
        ds.dist = sqrtf(dist_squared);
        ds.d = ds.d * b2_rcp(ds.dist);
        float dp = fabsf(dot(ds.d, ds.n));
        pdf *= dp != 0.f ? dist_squared / dp : 0.f;
        ds.pdf = pdf;
    } else {
        float3 c = v3(sh.center[0], sh.center[1], sh.center[2]);
        float3 dc_v = c - ref_p;
        float dc_2 = dot(dc_v, dc_v);
        float radius_adj = sh.radius * (sh.flip_normals ? (1.f + B2_RAY_EPS) : (1.f - B2_RAY_EPS));
        float3 dd;
        if (dc_2 > b2_sqr(radius_adj)) {                              // reference point outside: cone sampling
            float inv_dc = b2_rsqrt(dc_2), sin_theta_max = sh.radius * inv_dc, sin_theta_max_2 = b2_sqr(sin_theta_max),
                  inv_sin_theta_max = b2_rcp(sin_theta_max), cos_theta_max = b2_safe_sqrt(1.f - sin_theta_max_2);
            float sin_theta_2 = sin_theta_max_2 > 0.00068523f ? 1.f - b2_sqr(b2_fma(cos_theta_max - 1.f, sx, 1.f)) : sin_theta_max_2 * sx,
                  cos_theta = b2_safe_sqrt(1.f - sin_theta_2);
            float cos_alpha = sin_theta_2 * inv_sin_theta_max + cos_theta * b2_safe_sqrt(b2_fma(-sin_theta_2, b2_sqr(inv_sin_theta_max), 1.f)),
                  sin_alpha = b2_safe_sqrt(b2_fma(-cos_alpha, cos_alpha, 1.f));
            float sin_phi, cos_phi; b2_sincos(sy * (2.f * B2_PI), &sin_phi, &cos_phi);
            float3 fn = dc_v * (-inv_dc), fs = coordinate_system_s(fn), ft = coordinate_system_t(fn);   // Frame3f(n)
            float3 l = v3(cos_phi * sin_alpha, sin_phi * sin_alpha, cos_alpha);
            dd = (fs * l.x + ft * l.y) + fn * l.z;
            ds.p = v3(b2_fma(dd.x, sh.radius, c.x), b2_fma(dd.y, sh.radius, c.y), b2_fma(dd.z, sh.radius, c.z));
            ds.n = dd;
            ds.d = ds.p - ref_p;
            float dist2 = dot(ds.d, ds.d);
            ds.dist = sqrtf(dist2);
            ds.d = ds.d * b2_rcp(ds.dist);
            ds.pdf = (.5f * B2_INV_PI) / (1.f - cos_theta_max);       // square_to_uniform_cone_pdf, warp.h:481-490
            if (ds.dist == 0.f) ds.pdf = 0.f;
        } else {                                                       // inside: uniform area sampling
            dd = square_to_uniform_sphere(sx, sy);
            ds.p = v3(b2_fma(dd.x, sh.radius, c.x), b2_fma(dd.y, sh.radius, c.y), b2_fma(dd.z, sh.radius, c.z));
            ds.n = dd;
            ds.d = ds.p - ref_p;
            float dist2 = dot(ds.d, ds.d);
            ds.dist = sqrtf(dist2);
            ds.d = ds.d * b2_rcp(ds.dist);
            ds.pdf = sh.inv_area * dist2 / fabsf(dot(ds.d, ds.n));
        }
        if (sh.flip_normals) ds.n = -ds.n;
    }
    bool active = dot(ds.d, ds.n) < 0.f && ds.pdf != 0.f;             // area.cpp:131-134
    if (!active) return v3(0.f, 0.f, 0.f);
    float ip = b2_rcp(ds.pdf);
    return v3(em.radiance[0] * ip, em.radiance[1] * ip, em.radiance[2] * ip);
}

template <uint32_t F> B2_D float3 emitter_sample_direction(const Tables &sc, const DEnv &env, uint32_t e, float3 ref_p, float sx, float sy, DS &ds) {
    ds.delta = false;
    if constexpr ((F & F_ENV) != 0) if ((int) e == env.index) return env_sample_direction(env, ref_p, sx, sy, ds);
    if constexpr ((F & F_TEX) != 0) {
        const DEmitter &em = sc.emitters[e];
        if (em.type == B2_EMITTER_POINT) {                            // PointLight::sample_direction, point.cpp:86-113
            ds.p = v3(em.position[0], em.position[1], em.position[2]);
            ds.n = v3(0.f, 0.f, 0.f);
            ds.pdf = 1.f; ds.delta = true;
            ds.d = ds.p - ref_p;
            ds.dist = sqrtf(dot(ds.d, ds.d));
            float inv_dist = b2_rcp(ds.dist);
            ds.d = ds.d * inv_dist;
            float w = b2_sqr(inv_dist);
            return v3(em.radiance[0] * w, em.radiance[1] * w, em.radiance[2] * w);
        }
        if (em.type == B2_EMITTER_SPOT) {                             // SpotLight::sample_direction + falloff_curve, spot.cpp:97-160
            ds.p = v3(em.position[0], em.position[1], em.position[2]);
            ds.n = v3(0.f, 0.f, 0.f);
            ds.pdf = 1.f; ds.delta = true;
            ds.d = ds.p - ref_p;
            ds.dist = sqrtf(dot(ds.d, ds.d));
            float inv_dist = b2_rcp(ds.dist);
            ds.d = ds.d * inv_dist;
            float3 local_dir = normalize(xf_vector(em.to_object, -ds.d));
            float cos_theta = local_dir.z;
            float f = cos_theta >= em.cos_beam ? 1.f : (em.cutoff_angle - b2_safe_acos(cos_theta)) * em.inv_transition;
            float3 res = v3(em.radiance[0], em.radiance[1], em.radiance[2]);
            if (!(cos_theta >= em.cos_beam)) res = res * f;
            if (cos_theta <= em.cos_cutoff) res = v3(0.f, 0.f, 0.f);
            return res * (inv_dist * inv_dist);
        }
        if (em.type == B2_EMITTER_DIRECTIONAL) {                      // DirectionalEmitter::sample_direction, directional.cpp:107-134
            float3 d = v3(em.position[0], em.position[1], em.position[2]);
            float dist = 2.f * env.radius;
            ds.p = ref_p - d * dist;
            ds.n = d; ds.pdf = 1.f; ds.delta = true;
            ds.d = -d; ds.dist = dist;
            return v3(em.radiance[0], em.radiance[1], em.radiance[2]);
        }
        if (em.shape > 0) return shape_sample_direction(sc, em, ref_p, sx, sy, ds);
    }
    return area_sample_direction(sc, sc.emitters[e], ref_p, sx, sy, ds);
}

template <uint32_t F> B2_D float3 scene_sample_emitter_direction(const Tables &sc, const DEnv &env, float3 ref_p, float sx, float sy, DS &ds) {   // scene.cpp:164-214
    uint32_t ne = sc.n_emitters;
    if (ne == 0) { ds.pdf = 0.f; ds.dist = 0.f; ds.delta = false; ds.p = ds.n = ds.d = v3(0.f, 0.f, 0.f); return v3(0.f, 0.f, 0.f); }
    if (ne == 1) return emitter_sample_direction<F>(sc, env, 0, ref_p, sx, sy, ds);
    float emitter_pdf = 1.f / (float) ne;
    uint32_t index = (uint32_t) (sx * (float) ne);
    if (index > ne - 1) index = ne - 1;
    sx = (sx - (float) index * emitter_pdf) * (float) ne;
    float3 spec = emitter_sample_direction<F>(sc, env, index, ref_p, sx, sy, ds);
    ds.pdf *= emitter_pdf;
    return spec * b2_rcp(emitter_pdf);
}

template <uint32_t F> B2_D float scene_pdf_emitter_direction(const Tables &sc, const DEnv &env, int emitter, float3 ref_p, float3 d, float3 n, float dist) {   // scene.cpp:216-231
    if constexpr ((F & F_ENV) != 0) if (emitter == env.index) {
        float pdf = env_pdf_direction(env, d);
        if (sc.n_emitters != 1) pdf = pdf * (1.f / (float) sc.n_emitters);
        return pdf;
    }
    const DEmitter &em = sc.emitters[emitter];
    float dp = dot(d, n);
    if (!(dp < 0.f)) return 0.f;                                   // area.cpp:170-171
    float pdf = em.cdf_norm;                                       // mesh.cpp:417-420
    float adp = fabsf(dp);
    bool sphere = false;
    if constexpr ((F & F_TEX) != 0) if (em.shape > 0) {
        const DShape &sh = sc.shapes[em.shape - 1];
        pdf = sh.inv_area;
        if (sh.type == B2_SHAPE_SPHERE) {                          // Sphere::pdf_direction, sphere.cpp:255-267
            sphere = true;
            float3 dc = v3(sh.center[0], sh.center[1], sh.center[2]) - ref_p;
            float sin_alpha = sh.radius * b2_rcp(sqrtf(dot(dc, dc))), cos_alpha = b2_safe_sqrt(1.f - sin_alpha * sin_alpha);
            pdf = sin_alpha < B2_ONE_MINUS_EPS ? (.5f * B2_INV_PI) / (1.f - cos_alpha) : sh.inv_area * b2_sqr(dist) / adp;
        }
    }
    if (!sphere) pdf *= adp != 0.f ? (dist * dist) / adp : 0.f;    // shape.cpp:312-323
    if (sc.n_emitters != 1) pdf = pdf * (1.f / (float) sc.n_emitters);
    return pdf;
}

B2_D float mis_weight(float a, float b) {    // path.cpp:223-227
    a *= a; b *= b;
    return a > 0.f ? a / (a + b) : 0.f;
}

// ---- sensor (src/sensors/perspective.cpp:182-216) ---------------------------------------------
B2_D Ray sensor_sample_ray(const B2Sensor &sn, float sx, float sy) {
    const float *M = sn.sample_to_camera;
    float px = sx + sn.principal_point_offset[0], py = sy + sn.principal_point_offset[1];
    float r[4];
    #pragma unroll
    for (int k = 0; k < 4; ++k) {                                  // transform.h:103-111
        float acc = M[4 * k + 3];
        acc = b2_fma(M[4 * k + 0], px, acc);
        acc = b2_fma(M[4 * k + 1], py, acc);
        acc = b2_fma(M[4 * k + 2], 0.f, acc);
        r[k] = acc;
    }
    float iw = b2_rcp(r[3]);
    float3 d = normalize(v3(r[0] * iw, r[1] * iw, r[2] * iw));
    float inv_z = b2_rcp(d.z);
    Ray ray;
    ray.mint = sn.near_clip * inv_z; ray.maxt = sn.far_clip * inv_z;
    const float *T = sn.to_world;
    ray.o = v3(T[3], T[7], T[11]);
    ray.d.x = b2_fma(T[2], d.z, b2_fma(T[1], d.y, T[0] * d.x));    // transform.h:117-125
    ray.d.y = b2_fma(T[6], d.z, b2_fma(T[5], d.y, T[4] * d.x));
    ray.d.z = b2_fma(T[10], d.z, b2_fma(T[9], d.y, T[8] * d.x));
    return ray;
}
