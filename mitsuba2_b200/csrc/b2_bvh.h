// b2_bvh.h — host-side BVH4 builder (binned SAH BVH2, collapsed to 4-wide nodes, emitted in
// breadth-first order so that the top of the tree is one contiguous prefix that the kernels
// stage into shared memory with a single bulk copy).
//
// Replaces the role of ShapeKDTree::build (include/mitsuba/render/kdtree.h:957-1900, built from
// src/librender/scene_native.inl:3-10) / ext/embree's BVH builder. Traversal results do not
// depend on the tree (see the tie rule in b2_kernels.cuh), so any correct builder is a drop-in.
#pragma once
#include <stdint.h>
#include <vector>

struct B2BvhTri { float p[3][3]; };
struct B2BvhBox { float lo[3], hi[3]; };

struct B2Bvh4 {
    std::vector<float> nodes;       // 32 floats (128 B) per node, layout in b2_device.cuh
    std::vector<uint32_t> order;    // slot -> input triangle index
    uint32_t n_nodes = 0;
    uint32_t depth = 0;
};

void b2_build_bvh4(const std::vector<B2BvhTri> &tris, B2Bvh4 &out);
// Same builder over arbitrary primitive boxes (used for the top-level BVH over instances).
// max_leaf = primitives per leaf (1..15); pad_scale multiplies the conservative inflation.
void b2_build_bvh4_boxes(const std::vector<B2BvhBox> &boxes, int max_leaf, B2Bvh4 &out);
// Append `src` to the global node array `dst` (32 floats per node): inner child indices are
// rebased by the node offset, leaf first-slot indices by `slot_base`. Returns the root index.
uint32_t b2_bvh4_append(std::vector<float> &dst, const B2Bvh4 &src, uint32_t slot_base);
