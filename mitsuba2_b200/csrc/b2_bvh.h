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

// Depth bound of every tree this builder emits (BVH2 levels; the collapsed BVH4 is never deeper than its BVH2). The
// traversal kernels size their stacks from it (b2_kernels.cuh: 3 * depth + 2 <= B2_STACK). Once the levels left
// are only just enough for a balanced tree over a node's primitives, the binned-SAH split gives way to object-median
// splits, so ANY input (long thin triangles, multi-scale soups, 10^8 primitives) builds within the bound.
#define B2_BVH_MAX_DEPTH 42

struct B2BvhTri { float p[3][3]; };
struct B2BvhBox { float lo[3], hi[3]; };

struct B2Bvh4 {
    std::vector<float> nodes;       // 32 floats (128 B) per node, layout in b2_device.cuh
    std::vector<uint32_t> order;    // slot -> input triangle index
    uint32_t n_nodes = 0;
    uint32_t depth = 0;
    bool breadth_first = true;      // node numbering: any prefix is a top of the tree (shared-memory staging)
};

void b2_build_bvh4(const std::vector<B2BvhTri> &tris, B2Bvh4 &out);
// Same builder over arbitrary primitive boxes (used for the top-level BVH over instances).
// max_leaf = primitives per leaf (1..15); pad_scale multiplies the conservative inflation.
void b2_build_bvh4_boxes(const std::vector<B2BvhBox> &boxes, int max_leaf, B2Bvh4 &out);
// Append `src` to the global node array `dst` (32 floats per node): inner child indices are
// rebased by the node offset, leaf first-slot indices by `slot_base`. Returns the root index.
uint32_t b2_bvh4_append(std::vector<float> &dst, const B2Bvh4 &src, uint32_t slot_base);

// Device builder (b2_lbvh.cu): LBVH over n primitives given in HOST memory as triangles (n * 9 floats) or boxes (n * 6
// floats: lo, hi), same output as the host builder (nodes / order copied back). ms_device = CUDA-event time of the build
// kernels (Morton codes, sort, hierarchy, fit, collapse), excluding the copies. Returns false when it cannot be used
// (no device, n < 2, tree deeper than B2_BVH_MAX_DEPTH): the caller then uses the host builder.
bool b2_build_bvh4_device(const float *prims, uint32_t n, bool is_boxes, int max_leaf, B2Bvh4 &out, float *ms_device, const char **error);
