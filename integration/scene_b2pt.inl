// scene_b2pt.inl — Scene::accel_* backed by libb2pt: the reference keeps its own integrator (src/integrators/path.cpp on the
// CPU) and only its ray queries go to the GPU, through the stage entry points of include/b2pt.h.
//
// Reference-side patch (the way scene_embree.inl is selected, src/librender/scene.cpp:10-14):
//
//     #if defined(MTS_ENABLE_EMBREE)
//     #  include "scene_embree.inl"
//     #elif defined(MTS_ENABLE_B2PT)          // <- added
//     #  include "scene_b2pt.inl"             // <- added
//     #else
//     #  include "scene_native.inl"
//     #endif
//
// and `target_link_libraries(mitsuba-render PRIVATE b2pt)`. integration/build_accel_shim.sh applies exactly that to a scratch
// copy and installs the resulting libmitsuba-render.so under oracle/_ref/dist_b2accel/ (the unmodified library stays where
// it is). The hooks are the ones the OptiX back end implements (include/mitsuba/render/scene.h:184-205,
// src/librender/scene_optix.inl): accel_init -> b2_scene_create, ray_intersect_preliminary -> b2_stage_intersect,
// ray_test -> b2_stage_occluded, accel_release -> b2_scene_destroy, with m_accel holding the state.
//
// The scalar variants trace ONE ray per call, so every query is a kernel launch: this shim demonstrates the boundary and its
// parity (tests/test_plugin_dropin.py::test_accel_shim_keeps_path_on_the_cpu), it is not a fast path — the fast path is the
// b2path integrator, which keeps the whole loop on the device.
#include <b2pt.h>
#include <mitsuba/render/mesh.h>

NAMESPACE_BEGIN(mitsuba)

namespace {
struct B2Accel {
    B2Scene *scene = nullptr;
    std::vector<const void *> mesh_of;      // top-level meshes in scene order
    std::vector<uint32_t> mesh_first;       // global primitive index of each mesh's first face (+ total)
};
}

MTS_VARIANT void Scene<Float, Spectrum>::accel_init_cpu(const Properties &) {
    if constexpr (is_array_v<Float>) {
        Throw("scene_b2pt.inl: scalar variants only");
    } else {
        B2Accel *acc = new B2Accel();
        std::vector<B2Mesh> meshes;
        uint32_t nt = 0;
        for (Shape *shape : m_shapes) {
            const Mesh<Float, Spectrum> *mesh = dynamic_cast<const Mesh<Float, Spectrum> *>(shape);
            if (!mesh) Throw("scene_b2pt.inl: only triangle meshes are supported (got %s)", shape->class_()->name());
            B2Mesh d;
            std::memset(&d, 0, sizeof d);
            d.n_vertices = (uint32_t) mesh->vertex_count(); d.n_faces = (uint32_t) mesh->face_count();
            d.positions = mesh->vertex_positions_buffer().data();
            d.faces = (const uint32_t *) mesh->faces_buffer().data();
            d.bsdf = 0; d.emitter = -1;
            acc->mesh_of.push_back(mesh); acc->mesh_first.push_back(nt); nt += d.n_faces;
            meshes.push_back(d);
        }
        acc->mesh_first.push_back(nt);
        // the library wants a complete description: one placeholder BSDF, a 1 x 1 film and an identity sensor (unused by the stages)
        B2Bsdf bsdf; std::memset(&bsdf, 0, sizeof bsdf); bsdf.type = B2_BSDF_DIFFUSE;
        B2SceneDesc desc; std::memset(&desc, 0, sizeof desc);
        desc.n_meshes = desc.n_top_meshes = (uint32_t) meshes.size(); desc.meshes = meshes.data();
        desc.n_bsdfs = 1; desc.bsdfs = &bsdf;
        b2_host_film_box(1, 1, &desc.film);
        for (int i = 0; i < 4; ++i) desc.sensor.sample_to_camera[5 * i] = desc.sensor.to_world[5 * i] = 1.f;
        desc.sensor.near_clip = 1.f; desc.sensor.far_clip = 2.f;
        if (b2_scene_create(&desc, &acc->scene) != B2_OK) { std::string msg = b2_last_error(); delete acc; Throw("scene_b2pt.inl: %s", msg); }
        m_accel = acc;
    }
}

MTS_VARIANT void Scene<Float, Spectrum>::accel_release_cpu() {
    B2Accel *acc = (B2Accel *) m_accel;
    if (acc) { b2_scene_destroy(acc->scene); delete acc; }
    m_accel = nullptr;
}

MTS_VARIANT typename Scene<Float, Spectrum>::PreliminaryIntersection3f
Scene<Float, Spectrum>::ray_intersect_preliminary_cpu(const Ray3f &ray, Mask active) const {
    PreliminaryIntersection3f pi;
    if constexpr (!is_array_v<Float>) {
        if (!active) return pi;
        const B2Accel *acc = (const B2Accel *) m_accel;
        float ro[4] = { ray.o.x(), ray.o.y(), ray.o.z(), ray.mint }, rd[4] = { ray.d.x(), ray.d.y(), ray.d.z(), ray.maxt }, hit[4];
        int32_t inst = -1;
        if (b2_stage_intersect(acc->scene, 1, ro, rd, hit, &inst) != B2_OK) Throw("scene_b2pt.inl: %s", b2_last_error());
        if (hit[0] < math::Infinity<float>) {
            uint32_t prim; std::memcpy(&prim, &hit[3], 4);
            size_t m = std::upper_bound(acc->mesh_first.begin(), acc->mesh_first.end(), prim) - acc->mesh_first.begin() - 1;
            pi.t = hit[0]; pi.prim_uv = Point2f(hit[1], hit[2]);
            pi.prim_index = prim - acc->mesh_first[m]; pi.shape_index = (uint32_t) m;
            pi.shape = (const Shape *) acc->mesh_of[m];
        }
    }
    return pi;
}

MTS_VARIANT typename Scene<Float, Spectrum>::SurfaceInteraction3f
Scene<Float, Spectrum>::ray_intersect_cpu(const Ray3f &ray, HitComputeFlags flags, Mask active) const {
    PreliminaryIntersection3f pi = ray_intersect_preliminary_cpu(ray, active);
    active &= pi.is_valid();
    SurfaceInteraction3f si;
    if (likely(any(active)))
        si = pi.compute_surface_interaction(ray, flags, active);        // Mesh::compute_surface_interaction: the reference's own
    else {
        si.wavelengths = ray.wavelengths;
        si.wi = -ray.d;
        si.t = math::Infinity<Float>;
    }
    return si;
}

MTS_VARIANT typename Scene<Float, Spectrum>::SurfaceInteraction3f
Scene<Float, Spectrum>::ray_intersect_naive_cpu(const Ray3f &ray, Mask active) const {
    return ray_intersect_cpu(ray, HitComputeFlags::All, active);
}

MTS_VARIANT typename Scene<Float, Spectrum>::Mask
Scene<Float, Spectrum>::ray_test_cpu(const Ray3f &ray, Mask active) const {
    if constexpr (!is_array_v<Float>) {
        if (!active) return false;
        const B2Accel *acc = (const B2Accel *) m_accel;
        float ro[4] = { ray.o.x(), ray.o.y(), ray.o.z(), ray.mint }, rd[4] = { ray.d.x(), ray.d.y(), ray.d.z(), ray.maxt };
        uint8_t occ = 0;
        if (b2_stage_occluded(acc->scene, 1, ro, rd, &occ) != B2_OK) Throw("scene_b2pt.inl: %s", b2_last_error());
        return occ != 0;
    } else {
        return active;
    }
}

NAMESPACE_END(mitsuba)
