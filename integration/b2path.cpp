// b2path.cpp — the reference-side binding of libb2pt: a Mitsuba 2 integrator plugin.
//
//     <integrator type="b2path"> <integer name="max_depth" value="6"/> <integer name="block_size" value="32"/> </integrator>
//
// is a drop-in for <integrator type="path"> on the scalar_rgb variant: the scene is loaded by the
// reference's own XML / plugin / OBJ code, this adapter flattens the loaded object graph into a
// B2SceneDesc (include/b2pt.h), b2_render() runs the wavefront path tracer on the GPU, and the
// XYZAW film goes back through the reference's Film::put / develop.
//
// Built by integration/build_plugin.sh against the reference's headers (compiled with the same
// flags as src/integrators/path.cpp) and installed next to the reference's plugins in
// oracle/_ref/dist/plugins/. This file is written against the reference's public plugin API
// (include/mitsuba/render/*.h); nothing is copied from it.
#include <mitsuba/core/properties.h>
#include <mitsuba/core/thread.h>
#include <mitsuba/core/transform.h>
#include <mitsuba/render/bsdf.h>
#include <mitsuba/render/emitter.h>
#include <mitsuba/render/film.h>
#include <mitsuba/render/imageblock.h>
#include <mitsuba/render/integrator.h>
#include <mitsuba/render/mesh.h>
#include <mitsuba/render/sampler.h>
#include <mitsuba/render/scene.h>
#include <mitsuba/render/sensor.h>
#include <mitsuba/render/spiral.h>
#include <mitsuba/render/fresnel.h>
#include <mitsuba/render/texture.h>

#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstring>
#include <map>
#include <typeinfo>

#include <b2pt.h>

#include <mitsuba/render/kdtree.h>

NAMESPACE_BEGIN(mitsuba)

namespace {
// The unmodified reference has no accessor for a ShapeGroup's children or the Scene's shapegroup list (INTEGRATION.md lists
// the two one-line accessors a maintainer would add instead). An explicit template instantiation may name private members
// ([temp.explicit]), which is what this idiom uses — no header of the reference is edited.
template <typename Tag, typename Tag::type M> struct B2Rob { friend typename Tag::type b2_get(Tag) { return M; } };
using B2SG = ShapeGroup<float, Color<float, 3>>;
using B2SC = Scene<float, Color<float, 3>>;
using B2KD = ShapeKDTree<float, Color<float, 3>>;
struct B2TagKd { using type = ref<B2KD> B2SG::*; friend type b2_get(B2TagKd); };
template struct B2Rob<B2TagKd, &B2SG::m_kdtree>;
struct B2TagGroups { using type = std::vector<ref<B2SG>> B2SC::*; friend type b2_get(B2TagGroups); };
template struct B2Rob<B2TagGroups, &B2SC::m_shapegroups>;
// Shape::m_to_world / m_to_object are protected: reachable through a pointer to member formed in a derived class
template <typename Float, typename Spectrum> struct B2ShapePeek : Shape<Float, Spectrum> {
    using Base = Shape<Float, Spectrum>;
    static const typename Base::ScalarTransform4f &to_world(const Base *s) { return s->*(&B2ShapePeek::m_to_world); }
    static const typename Base::ScalarTransform4f &to_object(const Base *s) { return s->*(&B2ShapePeek::m_to_object); }
};

/// Collects what Object::traverse() exposes (diffuse.cpp:137-139, dielectric.cpp:322-328, ...)
struct ParamCollector : TraversalCallback {
    std::map<std::string, Object *> objects;
    std::map<std::string, std::pair<const std::type_info *, void *>> params;
    void put_object(const std::string &name, Object *obj) override { objects[name] = obj; }
    void put_parameter_impl(const std::string &name, const std::type_info &type, void *ptr) override {
        params[name] = { &type, ptr };
    }
    float get_float(const std::string &name, float def) const {
        auto it = params.find(name);
        if (it == params.end() || *it->second.first != typeid(float)) return def;
        return *(float *) it->second.second;
    }
};
}  // namespace

template <typename Float, typename Spectrum>
class B2PathIntegrator final : public MonteCarloIntegrator<Float, Spectrum> {
public:
    MTS_IMPORT_BASE(MonteCarloIntegrator, m_max_depth, m_rr_depth, m_block_size, m_samples_per_pass, m_stop)
    MTS_IMPORT_TYPES(Scene, Sensor, Film, Sampler, Mesh, BSDF, Emitter, Texture, ImageBlock, Shape, Medium)

    B2PathIntegrator(const Properties &props) : Base(props) {
        // the sampler keeps its base seed protected; mirror it here (default 0, like 'independent')
        m_seed = (uint64_t) props.size_("seed", 0);
        m_lbvh = props.bool_("device_bvh", false);       // build the BVHs on the GPU (LBVH) instead of the host SAH builder
#if defined(B2_DIRECT)
        // DirectIntegrator (direct.cpp:78-100, 106-196) with its default shading_samples = 1 draws the same random numbers
        // in the same order, with the same MIS weights, as PathIntegrator with max_depth = 2
        if (props.has_property("shading_samples") && (props.has_property("emitter_samples") || props.has_property("bsdf_samples")))
            Throw("Cannot specify both 'shading_samples' and ('emitter_samples' and/or 'bsdf_samples').");     // direct.cpp:79-84
        size_t shading = props.size_("shading_samples", 1);
        m_emitter_samples = (uint32_t) props.size_("emitter_samples", shading);
        m_bsdf_samples = (uint32_t) props.size_("bsdf_samples", shading);
        if (m_emitter_samples + m_bsdf_samples == 0)
            Throw("Must have at least 1 BSDF or emitter sample!");                                                // direct.cpp:93-94
        m_max_depth = 2; m_rr_depth = 5;
#endif
    }

    bool render(Scene *scene, Sensor *sensor) override {
        if constexpr (is_array_v<Float> || !is_rgb_v<Spectrum>) {
            NotImplementedError("b2path: scalar_rgb only");
        } else {
            m_stop = false;
            ref<Film> film = sensor->film();
            SurfaceInteraction3f si0 = zero<SurfaceInteraction3f>();

            // ---- flatten the loaded object graph (scene.cpp:23-59 order) -------------------
            std::vector<B2Mesh> meshes;
            std::vector<B2Bsdf> bsdfs;
            std::vector<B2Emitter> emitters;
            std::vector<B2Texture> textures;
            std::map<const BSDF *, int> bsdf_index;
            std::map<const Emitter *, int> emitter_of_mesh;
            std::vector<B2Shape> shapes;
            std::vector<B2Instance> instances;
            std::vector<uint32_t> group_first;
            std::map<const Emitter *, int> emitter_of_shape;
            auto bsdf_of = [&](const BSDF *bsdf) {
                auto it = bsdf_index.find(bsdf);
                if (it == bsdf_index.end()) it = bsdf_index.emplace(bsdf, add_bsdf(bsdf, si0, bsdfs, textures)).first;
                return it->second;
            };
            auto add_mesh = [&](const Mesh *mesh) {
                B2Mesh d;
                std::memset(&d, 0, sizeof d);
                d.n_vertices = (uint32_t) mesh->vertex_count();
                d.n_faces = (uint32_t) mesh->face_count();
                d.positions = mesh->vertex_positions_buffer().data();
                d.normals = mesh->has_vertex_normals() ? mesh->vertex_normals_buffer().data() : nullptr;
                d.texcoords = mesh->has_vertex_texcoords() ? mesh->vertex_texcoords_buffer().data() : nullptr;   // mesh.h:131-136
                d.faces = (const uint32_t *) mesh->faces_buffer().data();
                d.bsdf = bsdf_of(mesh->bsdf());
                d.emitter = -1;
                if (mesh->is_emitter())
                    emitter_of_mesh[mesh->emitter()] = (int) meshes.size();
                meshes.push_back(d);
            };
            // shapegroups of the scene (scene.cpp:36-37) and their children (shapegroup.cpp:14-43)
            const std::vector<ref<B2SG>> &groups = (*(const B2SC *) scene).*b2_get(B2TagGroups());
            std::vector<const Shape *> instance_shapes;
            // top-level shapes first (B2SceneDesc: meshes [0, n_top_meshes))
            for (auto &shape : scene->shapes()) {
                const std::string cls = shape->class_()->name();
                if (const Mesh *mesh = dynamic_cast<const Mesh *>(shape.get())) { add_mesh(mesh); continue; }
                if (cls == "Instance") { instance_shapes.push_back(shape.get()); continue; }
                if (cls == "Sphere" || cls == "Rectangle") {
                    B2Shape sh;
                    std::memset(&sh, 0, sizeof sh);
                    sh.type = cls == "Sphere" ? B2_SHAPE_SPHERE : B2_SHAPE_RECTANGLE;
                    sh.bsdf = bsdf_of(shape->bsdf()); sh.emitter = -1;
                    store_row_major(B2ShapePeek<Float, Spectrum>::to_world(shape.get()).matrix, sh.to_world);
                    store_row_major(B2ShapePeek<Float, Spectrum>::to_object(shape.get()).matrix, sh.to_object);
                    if (sh.type == B2_SHAPE_SPHERE) {
                        // m_center / m_radius / m_flip_normals are private (sphere.cpp:452-458): the centre is the translation of
                        // to_world (sphere.cpp:119,127); the radius is recovered from bbox() = centre -+ radius (sphere.cpp:131-138)
                        // and checked to reproduce it exactly; flip_normals from the normal of a sampled position (:159-160)
                        for (int k = 0; k < 3; ++k) sh.center[k] = sh.to_world[4 * k + 3];
                        auto bb = shape->bbox();
                        const float r0 = sh.center[0] - bb.min.x();
                        bool exact = false;
                        for (int step = 0; step < 9 && !exact; ++step) {       // r0, then its neighbours: +1, -1, +2, -2, ... ulp
                            float cand = r0;
                            for (int k = 0; k < (step + 1) / 2; ++k) cand = std::nextafter(cand, (step & 1) ? INFINITY : -INFINITY);
                            bool ok = true;
                            for (int k = 0; k < 3 && ok; ++k) ok = (sh.center[k] - cand == bb.min[k]) && (sh.center[k] + cand == bb.max[k]);
                            if (ok) { sh.radius = cand; exact = true; }
                        }
                        if (!exact) Throw("b2path: could not recover the radius of %s exactly", shape->to_string());
                        auto ps = shape->sample_position(0.f, Point2f(.3f, .6f));
                        Vector3f out(ps.p.x() - sh.center[0], ps.p.y() - sh.center[1], ps.p.z() - sh.center[2]);
                        sh.flip_normals = dot(out, Vector3f(ps.n)) < 0.f ? 1 : 0;
                    }
                    if (shape->is_emitter()) emitter_of_shape[shape->emitter()] = (int) shapes.size();
                    shapes.push_back(sh);
                    continue;
                }
                Throw("b2path: shape \"%s\" is not supported (triangle meshes, sphere, rectangle, instance are)", cls);
            }
            const uint32_t n_top_meshes = (uint32_t) meshes.size();
            for (auto &g : groups) {
                group_first.push_back((uint32_t) meshes.size());
                const B2KD *kd = ((*g).*b2_get(B2TagKd())).get();
                for (size_t i = 0; i < kd->shape_count(); ++i) {
                    const Mesh *mesh = dynamic_cast<const Mesh *>(kd->shape(i));
                    if (!mesh) Throw("b2path: only triangle meshes are supported inside a shapegroup (got %s)", kd->shape(i)->class_()->name());
                    add_mesh(mesh);
                }
            }
            group_first.push_back((uint32_t) meshes.size());
            for (const Shape *inst : instance_shapes) {
                // Instance::m_shapegroup is private to instance.cpp: it is the only data member after the Shape base, so it
                // sits at sizeof(Shape); the value is only used after it has been found in the scene's shapegroup list
                const B2SG *sg = *(const B2SG *const *) ((const char *) inst + sizeof(Shape));
                int gi = -1;
                for (size_t k = 0; k < groups.size(); ++k) if (groups[k].get() == sg) gi = (int) k;
                if (gi < 0) Throw("b2path: could not identify the shapegroup of %s", inst->to_string());
                B2Instance in;
                std::memset(&in, 0, sizeof in);
                in.group = gi;
                store_row_major(B2ShapePeek<Float, Spectrum>::to_world(inst).matrix, in.to_world);
                store_row_major(B2ShapePeek<Float, Spectrum>::to_object(inst).matrix, in.to_object);
                instances.push_back(in);
            }
            // emitters in scene->emitters() order (uniform selection index, scene.cpp:176-198)
            for (auto &em : scene->emitters()) {
                B2Emitter e;
                std::memset(&e, 0, sizeof e);
                ParamCollector pc;
                const_cast<Emitter *>(em.get())->traverse(&pc);
                if (em->is_environment() && std::string(em->class_()->name()) == "ConstantBackgroundEmitter") {   // constant.cpp:120-122
                    e.type = B2_EMITTER_CONSTANT; e.mesh = -1; e.scale = 1.f;
                    Color3f rad = ((Texture *) pc.objects["radiance"])->eval(si0);
                    for (int k = 0; k < 3; ++k) e.radiance[k] = rad[k];
                } else if (std::string(em->class_()->name()) == "PointLight") {                                     // point.cpp:86-113, 128-130
                    e.type = B2_EMITTER_POINT; e.mesh = -1; e.scale = 1.f;
                    Color3f in = ((Texture *) pc.objects["intensity"])->eval(si0);
                    for (int k = 0; k < 3; ++k) e.radiance[k] = in[k];
                    auto trafo = em->world_transform()->eval(0.f);
                    store_row_major(trafo.matrix, e.env_to_world);
                    store_row_major(trafo.inverse().matrix, e.env_to_object);
                } else if (std::string(em->class_()->name()) == "DirectionalEmitter") {                             // directional.cpp:107-134
                    e.type = B2_EMITTER_DIRECTIONAL; e.mesh = -1; e.scale = 1.f;
                    Color3f in = ((Texture *) pc.objects["irradiance"])->eval(si0);
                    for (int k = 0; k < 3; ++k) e.radiance[k] = in[k];
                    auto trafo = em->world_transform()->eval(0.f);
                    store_row_major(trafo.matrix, e.env_to_world);
                    store_row_major(trafo.inverse().matrix, e.env_to_object);
                } else if (std::string(em->class_()->name()) == "SpotLight") {                                      // spot.cpp:72-95
                    if (((Texture *) pc.objects["texture"])->is_spatially_varying()) Throw("b2path: textured spot lights are not supported");
                    e.type = B2_EMITTER_SPOT; e.mesh = -1; e.scale = 1.f;
                    Color3f in = ((Texture *) pc.objects["intensity"])->eval(si0);
                    for (int k = 0; k < 3; ++k) e.radiance[k] = in[k];
                    auto trafo = em->world_transform()->eval(0.f);
                    store_row_major(trafo.matrix, e.env_to_world);
                    store_row_major(trafo.inverse().matrix, e.env_to_object);
                    // the angles are only printed (to_string, 6 significant digits): take the printed radians back to the
                    // degrees the scene file gave (rounded to 1e-3 deg) and redo the constructor's arithmetic (spot.cpp:88-91)
                    std::string str = em->to_string();
                    auto field = [&](const char *name) {
                        size_t p = str.find(name);
                        if (p == std::string::npos) Throw("b2path: cannot read %s of %s", name, str);
                        return std::stod(str.substr(p + std::strlen(name)));
                    };
                    auto redo = [](double rad) { float deg = (float) (std::round(rad * 180.0 / 3.14159265358979323846 * 1000.0) / 1000.0); return deg_to_rad(deg); };
                    e.cutoff_angle = redo(field("cutoff_angle = ")); e.beam_width = redo(field("beam_width = "));
                } else if (em->is_environment()) {
                    if (pc.params.count("data") == 0 || pc.params.count("resolution") == 0)
                        Throw("b2path: unsupported environment emitter %s", em->class_()->name());
                    auto &data = *(DynamicBuffer<Float> *) pc.params["data"].second;
                    auto &res = *(ScalarVector2u *) pc.params["resolution"].second;
                    e.type = B2_EMITTER_ENVMAP; e.mesh = -1;
                    e.scale = pc.get_float("scale", 1.f);
                    e.envmap = data.data(); e.env_width = (int32_t) res.x(); e.env_height = (int32_t) res.y();
                    auto trafo = em->world_transform()->eval(0.f);
                    store_row_major(trafo.matrix, e.env_to_world);
                    store_row_major(trafo.inverse().matrix, e.env_to_object);
                } else {
                    auto it = emitter_of_mesh.find(em.get());
                    auto is = emitter_of_shape.find(em.get());
                    if ((it == emitter_of_mesh.end() && is == emitter_of_shape.end()) || pc.objects.count("radiance") == 0)
                        Throw("b2path: only area, envmap, constant, point, spot and directional emitters are supported (got %s)", em->class_()->name());
                    e.type = B2_EMITTER_AREA; e.scale = 1.f;
                    Color3f rad = ((Texture *) pc.objects["radiance"])->eval(si0);
                    for (int k = 0; k < 3; ++k) e.radiance[k] = rad[k];
                    if (it != emitter_of_mesh.end()) { e.mesh = it->second; meshes[it->second].emitter = (int32_t) emitters.size(); }
                    else { e.mesh = -1; e.shape = is->second + 1; shapes[is->second].emitter = (int32_t) emitters.size(); }
                }
                emitters.push_back(e);
            }

            B2SceneDesc desc;
            std::memset(&desc, 0, sizeof desc);
            desc.n_meshes = (uint32_t) meshes.size(); desc.meshes = meshes.data(); desc.n_top_meshes = n_top_meshes;
            if (!groups.empty()) { desc.n_groups = (uint32_t) groups.size(); desc.group_first = group_first.data(); }
            desc.n_instances = (uint32_t) instances.size(); desc.instances = instances.empty() ? nullptr : instances.data();
            desc.n_shapes = (uint32_t) shapes.size(); desc.shapes = shapes.empty() ? nullptr : shapes.data();
            desc.build_flags = m_lbvh ? B2_BUILD_DEVICE_LBVH : B2_BUILD_HOST_SAH;
            desc.n_bsdfs = (uint32_t) bsdfs.size(); desc.bsdfs = bsdfs.data();
            desc.n_emitters = (uint32_t) emitters.size(); desc.emitters = emitters.data();
            desc.n_textures = (uint32_t) textures.size(); desc.textures = textures.empty() ? nullptr : textures.data();

            // ---- film + reconstruction filter LUT (rfilter.h:62-65) -------------------------
            B2Film &f = desc.film;
            ScalarVector2i size = film->size(), crop_size = film->crop_size();
            ScalarPoint2i crop_offset = film->crop_offset();
            f.width = size.x(); f.height = size.y();
            f.crop_offset[0] = crop_offset.x(); f.crop_offset[1] = crop_offset.y();
            f.crop_size[0] = crop_size.x(); f.crop_size[1] = crop_size.y();
            const auto *rf = film->reconstruction_filter();
            f.filter_radius = rf->radius();
            f.filter_scale = (float) B2_FILTER_RESOLUTION / f.filter_radius;
            for (int i = 0; i <= B2_FILTER_RESOLUTION; ++i)
                f.filter_lut[i] = rf->eval_discretized(((float) i + .5f) / f.filter_scale);
            f.border = (int32_t) rf->border_size();

            // ---- sensor: same derivation as PerspectiveCamera::update_camera_transforms -----
            ParamCollector spc;
            sensor->traverse(&spc);
            if (spc.params.count("x_fov") == 0)
                Throw("b2path: only the perspective sensor is supported");
            const auto *cam = dynamic_cast<const ProjectiveCamera<Float, Spectrum> *>(sensor);
            float x_fov = spc.get_float("x_fov", 0.f);
            auto c2s = perspective_projection<ScalarFloat>(size, crop_size, ScalarVector2i(crop_offset), x_fov,
                                                           cam->near_clip(), cam->far_clip());
            store_row_major(c2s.inverse().matrix, desc.sensor.sample_to_camera);
            store_row_major(sensor->world_transform()->eval(0.f).matrix, desc.sensor.to_world);
            desc.sensor.near_clip = cam->near_clip(); desc.sensor.far_clip = cam->far_clip();

            B2Scene *handle = nullptr;
            if (b2_scene_create(&desc, &handle) != B2_OK)
                Throw("b2path: %s", b2_last_error());

            // ---- render ---------------------------------------------------------------------
            B2RenderParams rp;
            std::memset(&rp, 0, sizeof rp);
            rp.max_depth = m_max_depth; rp.rr_depth = m_rr_depth;
            rp.spp = (uint32_t) sensor->sampler()->sample_count();
            rp.samples_per_pass = m_samples_per_pass == (uint32_t) -1 ? 0 : m_samples_per_pass;
            // the block size feeds the per-pixel seed (integrator.cpp:198): when it is not given, derive it exactly
            // like SamplingIntegrator::render() does from the film size and the thread count (integrator.cpp:88-97)
            uint32_t block_size = m_block_size;
            if (block_size == 0) {
                block_size = MTS_BLOCK_SIZE;
                size_t n_threads = __global_thread_count;
                while (!(block_size == 1 || (size_t) hprod((crop_size + (int) block_size - 1) / (int) block_size) >= n_threads))
                    block_size /= 2;
            }
            rp.block_size = block_size;
            rp.seed = m_seed; rp.rank = 0; rp.world_size = 1;
            if (m_emitter_samples != 1 || m_bsdf_samples != 1) {       // (1, 1) is the path integrator with max_depth 2
                rp.direct_emitter_samples = m_emitter_samples; rp.direct_bsdf_samples = m_bsdf_samples;
            }
            {   // sampler plugin -> B2_SAMPLER_* ('jitter' is only printed, stratified.cpp:133-140)
                std::string scls = sensor->sampler()->class_()->name();
                if (scls == "StratifiedSampler")
                    rp.sampler = B2_SAMPLER_STRATIFIED | (sensor->sampler()->to_string().find("jitter = 0") != std::string::npos ? B2_SAMPLER_NO_JITTER : 0);
                else if (scls == "MultijitterSampler")
                    rp.sampler = B2_SAMPLER_MULTIJITTER | (sensor->sampler()->to_string().find("jitter = 0") != std::string::npos ? B2_SAMPLER_NO_JITTER : 0);
                else if (scls == "LowDiscrepancySampler")
                    rp.sampler = B2_SAMPLER_LOW_DISCREPANCY;
                else if (scls == "OrthogonalSampler")        // (its 'strength' is not exposed: the default 2 = Bose construction is assumed)
                    rp.sampler = B2_SAMPLER_ORTHOGONAL | (sensor->sampler()->to_string().find("jitter = 0") != std::string::npos ? B2_SAMPLER_NO_JITTER : 0);
                else if (scls != "IndependentSampler")
                    Throw("b2path: sampler \"%s\" is not supported (independent, stratified, multijitter, ldsampler, orthogonal are)", scls);
            }

            film->prepare({ "X", "Y", "Z", "A", "W" });                  // integrator.cpp:71-73
            ref<ImageBlock> block = new ImageBlock(crop_size, 5, film->reconstruction_filter(), false, false, false);
            block->set_offset(crop_offset);
            block->clear();
            m_handle = handle;
            B2Stats stats;
            int rc = b2_render(handle, &rp, block->data().data(), &stats);
            m_handle = nullptr;
            b2_scene_destroy(handle);
            if (rc == B2_ERR_CANCELLED)
                return false;                                            // integrator.cpp:174-178
            if (rc != B2_OK)
                Throw("b2path: %s", b2_last_error());
            film->put(block);                                            // hdrfilm.cpp:207-211
            Log(Info, "b2path: %i samples, %i segments, %i shadow rays, %.1f ms on the GPU (%.1f Msamples/s)",
                stats.n_samples, stats.n_segments, stats.n_shadow_rays, stats.ms_total,
                stats.n_samples / (stats.ms_total * 1e3f));
            return true;
        }
    }

    void cancel() override {
        B2Scene *h = m_handle.load();
        if (h) b2_cancel(h);
        Base::cancel();
    }

    std::pair<Spectrum, Mask> sample(const Scene *, Sampler *, const RayDifferential3f &, const Medium *,
                                     Float *, Mask) const override {
        NotImplementedError("sample");       // the whole render() runs on the GPU
    }

    std::string to_string() const override {
        return tfm::format("B2PathIntegrator[\n  max_depth = %i,\n  rr_depth = %i\n]", m_max_depth, m_rr_depth);
    }

    MTS_DECLARE_CLASS()
private:
    template <typename Matrix> static void store_row_major(const Matrix &m, float *out) {
        for (int r = 0; r < 4; ++r)
            for (int c = 0; c < 4; ++c)
                out[4 * r + c] = (float) m(r, c);
    }

    static void eval3(Object *tex, const SurfaceInteraction3f &si, float out[3], float def) {
        if (!tex) { out[0] = out[1] = out[2] = def; return; }
        Color3f v = ((Texture *) tex)->eval(si);
        for (int k = 0; k < 3; ++k) out[k] = v[k];
    }

    /// SmoothDiffuse / SmoothDielectric / RoughConductor / SmoothConductor / SmoothPlastic -> B2Bsdf, anything else: NotImplemented
    /// Bitmap lookup as libb2pt performs it (bitmap.cpp:386-474), used only to recover filter_type / wrap_mode,
    /// which BitmapTextureImpl exposes neither through traverse() nor to_string()
    static void probe_eval(const B2Texture &t, float su, float sv, float out[3]) {
        auto wrap = [&](int v, int res) {
            if (t.wrap == B2_WRAP_CLAMP) return std::max(0, std::min(v, res - 1));
            int div = v / res, mod = v - div * res;
            if (mod < 0) mod += res;
            if (t.wrap == B2_WRAP_MIRROR && !(((div & 1) == 0) != (v < 0))) mod = res - 1 - mod;
            return mod;
        };
        auto fetch = [&](int x, int y, int c) { return t.data[(size_t) (x + y * t.width) * t.channels + (t.channels == 1 ? 0 : c)]; };
        float u = std::fma(t.to_uv[1], sv, std::fma(t.to_uv[0], su, t.to_uv[2])), v = std::fma(t.to_uv[4], sv, std::fma(t.to_uv[3], su, t.to_uv[5]));
        if (t.filter == B2_TEXFILTER_BILINEAR) {
            u = std::fma(u, (float) t.width, -.5f); v = std::fma(v, (float) t.height, -.5f);
            int ix = (int) std::floor(u), iy = (int) std::floor(v);
            float w1x = u - (float) ix, w1y = v - (float) iy, w0x = 1.f - w1x, w0y = 1.f - w1y;
            int x0 = wrap(ix, t.width), x1 = wrap(ix + 1, t.width), y0 = wrap(iy, t.height), y1 = wrap(iy + 1, t.height);
            for (int c = 0; c < 3; ++c) {
                float v0 = std::fma(w0x, fetch(x0, y0, c), w1x * fetch(x1, y0, c)), v1 = std::fma(w0x, fetch(x0, y1, c), w1x * fetch(x1, y1, c));
                out[c] = std::fma(w0y, v0, w1y * v1);
            }
        } else {
            int x = wrap((int) std::floor(u * (float) t.width), t.width), y = wrap((int) std::floor(v * (float) t.height), t.height);
            for (int c = 0; c < 3; ++c) out[c] = fetch(x, y, c);
        }
    }

    /// Texture -> 0 (not spatially varying: the caller evaluates it once) or the 1-based index of a B2Texture
    int convert_texture(Object *obj, std::vector<B2Texture> &textures) const {
        Texture *tex = (Texture *) obj;
        if (!tex || !tex->is_spatially_varying()) return 0;
        SurfaceInteraction3f si = zero<SurfaceInteraction3f>();
        ParamCollector pc;
        tex->traverse(&pc);
        std::string cls = tex->class_()->name();
        B2Texture t;
        std::memset(&t, 0, sizeof t);
        auto tr = pc.params.find("transform");
        if (tr == pc.params.end()) Throw("b2path: texture \"%s\" is not supported (checkerboard, bitmap are)", cls);
        const auto &m = ((ScalarTransform3f *) tr->second.second)->matrix;
        for (int r = 0; r < 3; ++r) for (int c = 0; c < 3; ++c) t.to_uv[3 * r + c] = (float) m(r, c);
        if (cls == "Checkerboard") {                                  // checkerboard.cpp:40-44, 86-90
            t.type = B2_TEX_CHECKERBOARD;
            for (const char *n : { "color0", "color1" }) {
                Texture *c = (Texture *) pc.objects[n];
                if (!c || c->is_spatially_varying()) Throw("b2path: checkerboard with nested textures is not supported");
                Color3f v = c->eval(si);
                for (int k = 0; k < 3; ++k) (n[5] == '0' ? t.color0 : t.color1)[k] = v[k];
            }
        } else if (cls.rfind("BitmapTextureImpl", 0) == 0) {          // bitmap.cpp:561-565, 676-688
            t.type = B2_TEX_BITMAP;
            t.channels = cls.find("Impl_1") != std::string::npos ? 1 : 3;
            auto &data = *(DynamicBuffer<Float> *) pc.params["data"].second;
            auto &res = *(ScalarVector2i *) pc.params["resolution"].second;
            t.data = data.data(); t.width = res.x(); t.height = res.y();
            bool found = false;
            for (int f = 1; f >= 0 && !found; --f)
                for (int w = 0; w < 3 && !found; ++w) {
                    t.filter = f; t.wrap = w;
                    bool ok = true;
                    for (int i = 0; i < 11 && ok; ++i)
                        for (int j = 0; j < 11 && ok; ++j) {
                            si.uv = Point2f(-1.37f + .371f * i, -1.29f + .353f * j);
                            Color3f ref = tex->eval(si);
                            float mine[3]; probe_eval(t, si.uv.x(), si.uv.y(), mine);
                            for (int k = 0; k < 3; ++k) ok = ok && std::abs(mine[k] - ref[k]) <= 1e-5f * std::max(1.f, std::abs(ref[k]));
                        }
                    found = ok;
                }
            if (!found) Throw("b2path: could not determine filter_type / wrap_mode of %s", tex->to_string());
        } else {
            Throw("b2path: texture \"%s\" is not supported (checkerboard, bitmap are)", cls);
        }
        textures.push_back(t);
        return (int) textures.size();
    }

    /// Appends the record(s) of a BSDF and returns the index of its (front) record. TwoSidedBRDF (twosided.cpp:63-92,
    /// 188-191): the nested BRDFs become ordinary records, the front one pointing at the back one.
    int add_bsdf(const BSDF *bsdf, const SurfaceInteraction3f &si, std::vector<B2Bsdf> &bsdfs, std::vector<B2Texture> &textures) const {
        if (bsdf && std::string(bsdf->class_()->name()) == "TwoSidedBRDF") {
            ParamCollector pc;
            const_cast<BSDF *>(bsdf)->traverse(&pc);
            const BSDF *b0 = (const BSDF *) pc.objects["brdf_0"], *b1 = (const BSDF *) pc.objects["brdf_1"];
            int front = add_bsdf(b0, si, bsdfs, textures);
            int back = b1 == b0 ? front : add_bsdf(b1, si, bsdfs, textures);
            if (bsdfs[front].twosided_back || bsdfs[back].twosided_back) Throw("b2path: nested twosided BSDFs are not supported");
            bsdfs[front].twosided_back = back + 1;
            return front;
        }
        bsdfs.push_back(convert_bsdf(bsdf, si, textures));
        return (int) bsdfs.size() - 1;
    }

    B2Bsdf convert_bsdf(const BSDF *bsdf, const SurfaceInteraction3f &si, std::vector<B2Texture> &textures) const {
        B2Bsdf b;
        std::memset(&b, 0, sizeof b);
        for (int k = 0; k < 3; ++k) b.specular_reflectance[k] = b.specular_transmittance[k] = 1.f;
        if (!bsdf) { b.type = B2_BSDF_NULL; return b; }
        ParamCollector pc;
        const_cast<BSDF *>(bsdf)->traverse(&pc);
        std::string cls = bsdf->class_()->name();
        for (auto &kv : pc.objects) {       // only (diffuse_)reflectance may be spatially varying
            Texture *t = dynamic_cast<Texture *>(kv.second);
            if (t && t->is_spatially_varying() && kv.first != "reflectance" && kv.first != "diffuse_reflectance")
                Throw("b2path: textured BSDF parameter \"%s\" is not supported", kv.first);
        }
        auto obj = [&](const char *n) -> Object * { auto it = pc.objects.find(n); return it == pc.objects.end() ? nullptr : it->second; };
        if (cls == "SmoothDiffuse") {
            b.type = B2_BSDF_DIFFUSE;
            eval3(obj("reflectance"), si, b.reflectance, .5f);
            b.reflectance_texture = convert_texture(obj("reflectance"), textures);
        } else if (cls == "SmoothDielectric") {
            b.type = B2_BSDF_DIELECTRIC;
            b.eta = pc.get_float("eta", 1.5046f / 1.000277f);
            eval3(obj("specular_reflectance"), si, b.specular_reflectance, 1.f);
            eval3(obj("specular_transmittance"), si, b.specular_transmittance, 1.f);
        } else if (cls == "RoughConductor") {
            b.type = B2_BSDF_ROUGHCONDUCTOR;
            std::string s = bsdf->to_string();                       // distribution / sample_visible are only printed
            b.distribution = s.find("ggx") != std::string::npos ? B2_MF_GGX : B2_MF_BECKMANN;
            b.sample_visible = s.find("sample_visible = 0") != std::string::npos ? 0 : 1;
            Object *au = obj("alpha") ? obj("alpha") : obj("alpha_u"), *av = obj("alpha") ? obj("alpha") : obj("alpha_v");
            b.alpha_u = au ? ((Texture *) au)->eval_1(si) : .1f;
            b.alpha_v = av ? ((Texture *) av)->eval_1(si) : .1f;
            eval3(obj("eta"), si, b.cond_eta, 0.f);
            eval3(obj("k"), si, b.cond_k, 1.f);
            eval3(obj("specular_reflectance"), si, b.specular_reflectance, 1.f);
            if (!b.sample_visible) NotImplementedError("b2path: roughconductor with sample_visible=false");
        } else if (cls == "RoughDielectric") {                       // roughdielectric.cpp:140-201, 453-465
            b.type = B2_BSDF_ROUGHDIELECTRIC;
            std::string s = bsdf->to_string();                       // distribution / sample_visible are only printed
            b.distribution = s.find("ggx") != std::string::npos ? B2_MF_GGX : B2_MF_BECKMANN;
            b.sample_visible = s.find("sample_visible = 0") != std::string::npos ? 0 : 1;
            Object *au = obj("alpha") ? obj("alpha") : obj("alpha_u"), *av = obj("alpha") ? obj("alpha") : obj("alpha_v");
            b.alpha_u = au ? ((Texture *) au)->eval_1(si) : .1f;
            b.alpha_v = av ? ((Texture *) av)->eval_1(si) : .1f;
            b.eta = pc.get_float("eta", 1.5046f / 1.000277f);
            eval3(obj("specular_reflectance"), si, b.specular_reflectance, 1.f);
            eval3(obj("specular_transmittance"), si, b.specular_transmittance, 1.f);
            if (!b.sample_visible) NotImplementedError("b2path: roughdielectric with sample_visible=false");
        } else if (cls == "SmoothConductor") {                       // conductor.cpp:160-176, 247-251
            b.type = B2_BSDF_CONDUCTOR;
            eval3(obj("eta"), si, b.cond_eta, 0.f);
            eval3(obj("k"), si, b.cond_k, 1.f);
            eval3(obj("specular_reflectance"), si, b.specular_reflectance, 1.f);
        } else if (cls == "SmoothPlastic") {                         // plastic.cpp:136-180, 295-301
            b.type = B2_BSDF_PLASTIC;
            b.eta = pc.get_float("eta", 1.49f / 1.000277f);
            eval3(obj("diffuse_reflectance"), si, b.reflectance, .5f);
            b.reflectance_texture = convert_texture(obj("diffuse_reflectance"), textures);
            eval3(obj("specular_reflectance"), si, b.specular_reflectance, 1.f);
            // the values parameters_changed() derives (:162-176), with the reference's own functions
            b.inv_eta_2 = 1.f / (b.eta * b.eta);
            b.fdr_int = fresnel_diffuse_reflectance(1.f / b.eta);
            float d_mean = obj("diffuse_reflectance") ? ((Texture *) obj("diffuse_reflectance"))->mean() : .5f,
                  s_mean = obj("specular_reflectance") ? ((Texture *) obj("specular_reflectance"))->mean() : 1.f;
            b.specular_sampling_weight = s_mean / (d_mean + s_mean);
            b.nonlinear = bsdf->to_string().find("nonlinear = 1") != std::string::npos ? 1 : 0;   // only printed
        } else {
            Throw("b2path: BSDF \"%s\" is not supported (diffuse, dielectric, roughconductor, roughdielectric, conductor, plastic, twosided are)", cls);
        }
        return b;
    }

    uint64_t m_seed;
    uint32_t m_emitter_samples = 1, m_bsdf_samples = 1;       // b2direct only
    bool m_lbvh = false;
    std::atomic<B2Scene *> m_handle { nullptr };
};

MTS_IMPLEMENT_CLASS_VARIANT(B2PathIntegrator, MonteCarloIntegrator)
MTS_EXPORT_PLUGIN(B2PathIntegrator, "B200 wavefront path tracer (libb2pt)")
NAMESPACE_END(mitsuba)
