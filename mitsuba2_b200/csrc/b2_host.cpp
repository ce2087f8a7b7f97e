// b2_host.cpp — host-side C++ mirrors of the reference's scene-setup code that feeds the hot
// path (the b2_host_* part of include/b2pt.h). No GPU involved.
#include "../../include/b2pt.h"

#include <cmath>
#include <cstring>
#include <vector>

namespace {

constexpr float kPi = 3.14159265358979323846f;
constexpr float kRayEps = 1500.0f * 5.9604644775390625e-08f;   // include/mitsuba/core/math.h:36

inline float fma_(float a, float b, float c) { return std::fma(a, b, c); }

// Row-major 4x4; Enoki's Matrix4f is column-major, m(r, c) == a[4 * r + c] here.
struct Mat4 {
    float a[16];
    static Mat4 identity() { Mat4 m{}; m.a[0] = m.a[5] = m.a[10] = m.a[15] = 1.f; return m; }
    float &operator()(int r, int c) { return a[4 * r + c]; }
    float operator()(int r, int c) const { return a[4 * r + c]; }
    Mat4 transpose() const { Mat4 t{}; for (int r = 0; r < 4; ++r) for (int c = 0; c < 4; ++c) t(r, c) = (*this)(c, r); return t; }
};
// ext/enoki/include/enoki/matrix.h:154-167 — per column: m0.col(0) * m1(0,j), then fmadd for i = 1..3
Mat4 operator*(const Mat4 &A, const Mat4 &B) {
    Mat4 R{};
    for (int j = 0; j < 4; ++j)
        for (int r = 0; r < 4; ++r) {
            float sum = A(r, 0) * B(0, j);
            for (int i = 1; i < 4; ++i) sum = fma_(A(r, i), B(i, j), sum);
            R(r, j) = sum;
        }
    return R;
}
// include/mitsuba/core/transform.h:29-61 — a Transform carries its matrix and inverse transpose
struct Transform {
    Mat4 m, inv_t;
    Transform operator*(const Transform &o) const { return { m * o.m, inv_t * o.inv_t }; }
    static Transform scale(float x, float y, float z) {                 // transform.h:167-171
        Transform t{ Mat4::identity(), Mat4::identity() };
        t.m(0, 0) = x; t.m(1, 1) = y; t.m(2, 2) = z;
        t.inv_t(0, 0) = 1.f / x; t.inv_t(1, 1) = 1.f / y; t.inv_t(2, 2) = 1.f / z;
        return t;
    }
    static Transform translate(float x, float y, float z) {             // transform.h:161-164
        Transform t{ Mat4::identity(), Mat4::identity() };
        t.m(0, 3) = x; t.m(1, 3) = y; t.m(2, 3) = z;
        Mat4 inv = Mat4::identity(); inv(0, 3) = -x; inv(1, 3) = -y; inv(2, 3) = -z;
        t.inv_t = inv.transpose();
        return t;
    }
    static Transform perspective(float fov, float near_, float far_) {  // transform.h:203-220
        float recip = 1.f / (far_ - near_);
        float tn = std::tan((fov * .5f) * (kPi / 180.f)), cot = 1.f / tn;
        Transform t{};
        t.m(0, 0) = cot; t.m(1, 1) = cot; t.m(2, 2) = far_ * recip; t.m(2, 3) = -near_ * far_ * recip; t.m(3, 2) = 1.f;
        Mat4 inv{};
        inv(0, 0) = tn; inv(1, 1) = tn; inv(3, 3) = 1.f / near_; inv(2, 3) = 1.f; inv(3, 2) = (near_ - far_) / (far_ * near_);
        t.inv_t = inv.transpose();
        return t;
    }
};

struct V3 { float x, y, z; };
V3 sub(V3 a, V3 b) { return { a.x - b.x, a.y - b.y, a.z - b.z }; }
float dot(V3 a, V3 b) { return (a.x * b.x + a.y * b.y) + a.z * b.z; }
V3 cross(V3 a, V3 b) { return { fma_(a.y, b.z, -(a.z * b.y)), fma_(a.z, b.x, -(a.x * b.z)), fma_(a.x, b.y, -(a.y * b.x)) }; }
V3 normalize(V3 a) { float s = 1.f / std::sqrt(dot(a, a)); return { a.x * s, a.y * s, a.z * s }; }

void init_filter(B2Film *f, float radius, float (*eval)(float, const float *), const float *p) {   // rfilter.cpp:9-20
    f->filter_radius = radius;
    for (int i = 0; i < B2_FILTER_RESOLUTION; ++i) f->filter_lut[i] = eval((radius * (float) i) / (float) B2_FILTER_RESOLUTION, p);
    f->filter_lut[B2_FILTER_RESOLUTION] = 0.f;
    f->filter_scale = (float) B2_FILTER_RESOLUTION / radius;
    f->border = (int) std::ceil(radius - .5f - 2.f * kRayEps);
}
void film_dims(B2Film *f, int w, int h) {
    std::memset(f, 0, sizeof *f);
    f->width = w; f->height = h; f->crop_size[0] = w; f->crop_size[1] = h;
}

}  // namespace

extern "C" {

int b2_host_look_at(const float o[3], const float tg[3], const float up_[3], float out[16]) {   // transform.h:241-270
    V3 origin{ o[0], o[1], o[2] }, target{ tg[0], tg[1], tg[2] }, up{ up_[0], up_[1], up_[2] };
    V3 dir = normalize(sub(target, origin));
    dir = normalize(dir);
    V3 left = normalize(cross(up, dir));
    V3 new_up = cross(dir, left);
    const float m[16] = { left.x, new_up.x, dir.x, origin.x, left.y, new_up.y, dir.y, origin.y,
                          left.z, new_up.z, dir.z, origin.z, 0.f, 0.f, 0.f, 1.f };
    std::memcpy(out, m, sizeof m);
    return B2_OK;
}

float b2_host_parse_fov(float fov, const char *axis, int width, int height) {   // src/librender/sensor.cpp:113-166
    float aspect = (float) width / (float) height;
    int mode;
    if (!std::strcmp(axis, "x")) mode = 0;
    else if (!std::strcmp(axis, "y")) mode = 1;
    else if (!std::strcmp(axis, "diagonal")) mode = 2;
    else if (!std::strcmp(axis, "smaller")) mode = aspect > 1.f ? 1 : 0;
    else if (!std::strcmp(axis, "larger")) mode = aspect > 1.f ? 0 : 1;
    else return -1.f;
    if (mode == 0) return fov;
    if (mode == 1) return (180.f / kPi) * (2.f * std::atan(std::tan(.5f * (fov * (kPi / 180.f))) * aspect));
    float diagonal = 2.f * std::tan(.5f * (fov * (kPi / 180.f)));
    float w = diagonal / std::sqrt(1.f + 1.f / (aspect * aspect));
    return (180.f / kPi) * (2.f * std::atan(w * .5f));
}

// perspective_projection (include/mitsuba/render/sensor.h:196-232) + update_camera_transforms
// (src/sensors/perspective.cpp:119-139)
int b2_host_perspective_sensor(const float to_world[16], float fov_x, float near_clip, float far_clip,
                               const B2Film *film, B2Sensor *out) {
    if (!to_world || !film || !out) return B2_ERR_INVALID;
    float fw = (float) film->width, fh = (float) film->height;
    float rsx = (float) film->crop_size[0] / fw, rsy = (float) film->crop_size[1] / fh;
    float rox = (float) film->crop_offset[0] / fw, roy = (float) film->crop_offset[1] / fh;
    float aspect = fw / fh;
    Transform c2s = Transform::scale(1.f / rsx, 1.f / rsy, 1.f) * Transform::translate(-rox, -roy, 0.f) *
                    Transform::scale(-0.5f, -0.5f * aspect, 1.f) * Transform::translate(-1.f, -1.f / aspect, 0.f) *
                    Transform::perspective(fov_x, near_clip, far_clip);
    Mat4 s2c = c2s.inv_t.transpose();                                   // Transform::inverse(), transform.h:59-61
    std::memcpy(out->sample_to_camera, s2c.a, sizeof s2c.a);
    std::memcpy(out->to_world, to_world, 16 * sizeof(float));
    out->near_clip = near_clip; out->far_clip = far_clip;
    out->principal_point_offset[0] = out->principal_point_offset[1] = 0.f;
    return B2_OK;
}

// SmoothPlastic (plastic.cpp:136-180) without the reference: parameters_changed()'s derived values in fp32
static float fresnel_diffuse_reflectance(float eta) {                    // fresnel.h:327-361
    if (eta < 1.f) return -1.4399f * (eta * eta) + 0.7099f * eta + 0.6681f + 0.0636f / eta;
    float i1 = 1.f / eta, i2 = i1 * i1, i3 = i2 * i1, i4 = i3 * i1, i5 = i4 * i1;
    return 0.919317f - 3.4793f * i1 + 6.75335f * i2 - 7.80989f * i3 + 4.98554f * i4 - 1.36881f * i5;
}
int b2_host_bsdf_plastic(float int_ior, float ext_ior, const float diffuse_reflectance[3], float diffuse_mean,
                         const float specular_reflectance[3], float specular_mean, int nonlinear, B2Bsdf *out) {
    if (!out || !diffuse_reflectance || !(int_ior > 0.f) || !(ext_ior > 0.f)) return B2_ERR_INVALID;   // plastic.cpp:143-145
    std::memset(out, 0, sizeof *out);
    out->type = B2_BSDF_PLASTIC;
    float eta = int_ior / ext_ior;
    out->eta = eta; out->inv_eta_2 = 1.f / (eta * eta); out->fdr_int = fresnel_diffuse_reflectance(1.f / eta);
    for (int k = 0; k < 3; ++k) {
        out->reflectance[k] = diffuse_reflectance[k];
        out->specular_reflectance[k] = specular_reflectance ? specular_reflectance[k] : 1.f;
        out->specular_transmittance[k] = 1.f;
    }
    float s_mean = specular_reflectance ? specular_mean : 1.f;
    out->specular_sampling_weight = s_mean / (diffuse_mean + s_mean);
    out->nonlinear = nonlinear ? 1 : 0;
    return B2_OK;
}

// general 4x4 affine inverse (in double, rounded once) for the shape records
static bool invert_affine(const float M[16], float inv[16]) {
    double a[3][3], t[3];
    for (int r = 0; r < 3; ++r) { for (int c = 0; c < 3; ++c) a[r][c] = M[4 * r + c]; t[r] = M[4 * r + 3]; }
    double det = a[0][0] * (a[1][1] * a[2][2] - a[1][2] * a[2][1]) - a[0][1] * (a[1][0] * a[2][2] - a[1][2] * a[2][0]) +
                 a[0][2] * (a[1][0] * a[2][1] - a[1][1] * a[2][0]);
    if (det == 0.0 || !std::isfinite(det)) return false;
    double id = 1.0 / det, b[3][3];
    b[0][0] = (a[1][1] * a[2][2] - a[1][2] * a[2][1]) * id; b[0][1] = (a[0][2] * a[2][1] - a[0][1] * a[2][2]) * id; b[0][2] = (a[0][1] * a[1][2] - a[0][2] * a[1][1]) * id;
    b[1][0] = (a[1][2] * a[2][0] - a[1][0] * a[2][2]) * id; b[1][1] = (a[0][0] * a[2][2] - a[0][2] * a[2][0]) * id; b[1][2] = (a[0][2] * a[1][0] - a[0][0] * a[1][2]) * id;
    b[2][0] = (a[1][0] * a[2][1] - a[1][1] * a[2][0]) * id; b[2][1] = (a[0][1] * a[2][0] - a[0][0] * a[2][1]) * id; b[2][2] = (a[0][0] * a[1][1] - a[0][1] * a[1][0]) * id;
    for (int r = 0; r < 3; ++r) {
        for (int c = 0; c < 3; ++c) inv[4 * r + c] = (float) b[r][c];
        inv[4 * r + 3] = (float) -(b[r][0] * t[0] + b[r][1] * t[1] + b[r][2] * t[2]);
    }
    inv[12] = inv[13] = inv[14] = 0.f; inv[15] = 1.f;
    return true;
}
int b2_host_shape_sphere(const float center[3], float radius, int flip_normals, int bsdf, B2Shape *out) {
    if (!out || !center || radius == 0.f) return B2_ERR_INVALID;
    std::memset(out, 0, sizeof *out);
    out->type = B2_SHAPE_SPHERE; out->bsdf = bsdf; out->emitter = -1;
    if (radius < 0.f) { radius = -radius; flip_normals = !flip_normals; }      // sphere.cpp:121-124
    out->flip_normals = flip_normals ? 1 : 0; out->radius = radius;
    for (int k = 0; k < 3; ++k) {
        out->center[k] = center[k];
        out->to_world[5 * k] = radius; out->to_world[4 * k + 3] = center[k];   // translate(center) * scale(radius), sphere.cpp:100-101
    }
    out->to_world[15] = 1.f;
    return invert_affine(out->to_world, out->to_object) ? B2_OK : B2_ERR_INVALID;
}
int b2_host_shape_rectangle(const float to_world[16], int flip_normals, int bsdf, B2Shape *out) {
    if (!out || !to_world) return B2_ERR_INVALID;
    std::memset(out, 0, sizeof *out);
    out->type = B2_SHAPE_RECTANGLE; out->bsdf = bsdf; out->emitter = -1;
    std::memcpy(out->to_world, to_world, 64);
    if (flip_normals) for (int r = 0; r < 4; ++r) out->to_world[4 * r + 2] = -out->to_world[4 * r + 2];   // * scale(1, 1, -1), rectangle.cpp:75-77
    return invert_affine(out->to_world, out->to_object) ? B2_OK : B2_ERR_INVALID;
}

int b2_host_film_box(int w, int h, B2Film *f) {                          // src/rfilters/box.cpp:24-33
    if (!f || w <= 0 || h <= 0) return B2_ERR_INVALID;
    film_dims(f, w, h);
    float radius = .5f + kRayEps;
    init_filter(f, radius, [](float x, const float *p) { return std::fabs(x) <= p[0] ? 1.f : 0.f; }, &radius);
    return B2_OK;
}

int b2_host_film_gaussian(int w, int h, float stddev, B2Film *f) {       // src/rfilters/gaussian.cpp:32-47
    if (!f || w <= 0 || h <= 0 || !(stddev > 0.f)) return B2_ERR_INVALID;
    film_dims(f, w, h);
    float radius = 4 * stddev;
    float p[2];
    p[0] = -1.f / (2.f * stddev * stddev);
    p[1] = std::exp(p[0] * (radius * radius));
    init_filter(f, radius, [](float x, const float *q) { float v = std::exp(q[0] * (x * x)) - q[1]; return v > 0.f ? v : 0.f; }, p);
    return B2_OK;
}

// Spiral::next_block order — src/librender/spiral.cpp:8-72
int b2_host_spiral(int width, int height, int bs, uint32_t *n_blocks, int32_t *order_xy) {
    if (!n_blocks || width <= 0 || height <= 0 || bs <= 0) return B2_ERR_INVALID;
    int bx = (int) std::ceil((float) width / (float) bs), by = (int) std::ceil((float) height / (float) bs);
    uint32_t count = (uint32_t) (bx * by);
    *n_blocks = count;
    if (!order_xy) return B2_OK;
    int px = bx / 2, py = by / 2, dir = 0 /* 0 right, 1 down, 2 left, 3 up */, steps_left = 1, steps = 1;
    for (uint32_t counter = 0; counter < count;) {
        order_xy[2 * counter] = px; order_xy[2 * counter + 1] = py;
        if (++counter == count) break;
        do {
            switch (dir) { case 0: ++px; break; case 1: ++py; break; case 2: --px; break; default: --py; break; }
            if (--steps_left == 0) {
                dir = (dir + 1) % 4;
                if (dir == 0 || dir == 2) ++steps;
                steps_left = steps;
            }
        } while (px < 0 || py < 0 || px >= bx || py >= by);
    }
    return B2_OK;
}

}  // extern "C"
