// b2_bvh.cpp — see b2_bvh.h
#include <cstdlib>
#include <thread>
#include "b2_bvh.h"

#include <algorithm>
#include <cmath>
#include <cstring>
#include <limits>

namespace {

struct Box {
    float lo[3], hi[3];
    void reset() { for (int k = 0; k < 3; ++k) { lo[k] = std::numeric_limits<float>::infinity(); hi[k] = -lo[k]; } }
    void grow(const Box &b) { for (int k = 0; k < 3; ++k) { lo[k] = std::min(lo[k], b.lo[k]); hi[k] = std::max(hi[k], b.hi[k]); } }
    void grow(const float *p) { for (int k = 0; k < 3; ++k) { lo[k] = std::min(lo[k], p[k]); hi[k] = std::max(hi[k], p[k]); } }
    float area() const {
        float d[3] = { hi[0] - lo[0], hi[1] - lo[1], hi[2] - lo[2] };
        if (d[0] < 0) return 0.f;
        return 2.f * (d[0] * d[1] + d[1] * d[2] + d[2] * d[0]);
    }
};

struct Node2 { Box box; int left = -1, right = -1; uint32_t first = 0, count = 0; };

struct Builder {
    std::vector<Box> tbox;
    std::vector<float> cent;        // 3 per tri
    std::vector<uint32_t> idx;
    std::vector<Node2> nodes;
    static constexpr int kBins = 16;
    int kMaxLeaf = 2;

    // One node: bounds + binned-SAH split of idx[first, first+count). Returns false for a leaf.
    // levels a balanced tree over `count` primitives needs below this node
    int levels_needed(uint32_t count) const {
        uint32_t leaves = (count + (uint32_t) kMaxLeaf - 1) / (uint32_t) kMaxLeaf;
        int l = 0;
        while ((1u << l) < leaves) ++l;
        return l;
    }

    bool make_node(uint32_t first, uint32_t count, Node2 &node, uint32_t &mid, int depth) {
        Box box; box.reset();
        Box cbox; cbox.reset();
        for (uint32_t i = first; i < first + count; ++i) {
            box.grow(tbox[idx[i]]);
            cbox.grow(&cent[3 * idx[i]]);
        }
        node.box = box; node.first = first; node.count = count; node.left = node.right = -1;
        if (count <= (uint32_t) kMaxLeaf) return false;
        if (depth + levels_needed(count) >= B2_BVH_MAX_DEPTH - 1) {
            // level budget nearly spent: object-median split along the widest centroid axis (keeps the subtree balanced)
            int ax = 0;
            for (int k = 1; k < 3; ++k) if (cbox.hi[k] - cbox.lo[k] > cbox.hi[ax] - cbox.lo[ax]) ax = k;
            mid = first + count / 2;
            std::nth_element(idx.begin() + first, idx.begin() + mid, idx.begin() + first + count,
                             [&](uint32_t a, uint32_t b) { float ca = cent[3 * a + ax], cb = cent[3 * b + ax]; return ca < cb || (ca == cb && a < b); });
            node.count = 0;
            return true;
        }
        // binned SAH over the 3 axes
        float best_cost = std::numeric_limits<float>::infinity(); int best_axis = -1, best_bin = -1;
        for (int ax = 0; ax < 3; ++ax) {
            float lo = cbox.lo[ax], ext = cbox.hi[ax] - lo;
            if (!(ext > 0.f)) continue;
            Box bb[kBins]; uint32_t bc[kBins];
            for (int b = 0; b < kBins; ++b) { bb[b].reset(); bc[b] = 0; }
            float scale = kBins / ext;
            for (uint32_t i = first; i < first + count; ++i) {
                int b = std::min(kBins - 1, (int) ((cent[3 * idx[i] + ax] - lo) * scale));
                bb[b].grow(tbox[idx[i]]); bc[b]++;
            }
            float ra[kBins]; uint32_t rc[kBins];
            Box acc; acc.reset(); uint32_t c = 0;
            for (int b = kBins - 1; b > 0; --b) { acc.grow(bb[b]); c += bc[b]; ra[b] = acc.area(); rc[b] = c; }
            acc.reset(); c = 0;
            for (int b = 0; b < kBins - 1; ++b) {
                acc.grow(bb[b]); c += bc[b];
                if (c == 0 || rc[b + 1] == 0) continue;
                float cost = acc.area() * (float) c + ra[b + 1] * (float) rc[b + 1];
                if (cost < best_cost) { best_cost = cost; best_axis = ax; best_bin = b; }
            }
        }
        if (best_axis < 0) {
            mid = first + count / 2;            // all centroids coincide: split in the middle
        } else {
            float lo = cbox.lo[best_axis], scale = kBins / (cbox.hi[best_axis] - lo);
            auto it = std::partition(idx.begin() + first, idx.begin() + first + count, [&](uint32_t t) {
                int b = std::min(kBins - 1, (int) ((cent[3 * t + best_axis] - lo) * scale));
                return b <= best_bin;
            });
            mid = (uint32_t) (it - idx.begin());
            if (mid == first || mid == first + count) mid = first + count / 2;
        }
        node.count = 0;
        return true;
    }

    // serial build into `dst` (pre-order numbering, indices local to dst)
    static int build_serial(Builder &b, std::vector<Node2> &dst, uint32_t first, uint32_t count, int depth) {
        int id = (int) dst.size();
        dst.emplace_back();
        Node2 n; uint32_t mid = 0;
        bool inner = b.make_node(first, count, n, mid, depth);
        dst[id] = n;
        if (!inner) return id;
        int l = build_serial(b, dst, first, mid - first, depth + 1);
        int r = build_serial(b, dst, mid, first + count - mid, depth + 1);
        dst[id].left = l; dst[id].right = r;
        return id;
    }

    // Parallel build with the SAME result as the serial one: the two subtrees of the top `par_depth` levels are built
    // concurrently into their own arrays (they partition disjoint ranges of idx) and spliced in pre-order.
    static void build_parallel(Builder &b, std::vector<Node2> &dst, uint32_t first, uint32_t count, int par_depth, int depth) {
        if (par_depth <= 0 || count < 16384) { build_serial(b, dst, first, count, depth); return; }
        Node2 n; uint32_t mid = 0;
        if (!b.make_node(first, count, n, mid, depth)) { dst.push_back(n); return; }
        std::vector<Node2> L, R;
        std::thread th([&] { build_parallel(b, L, first, mid - first, par_depth - 1, depth + 1); });
        build_parallel(b, R, mid, first + count - mid, par_depth - 1, depth + 1);
        th.join();
        const int base = (int) dst.size(), lo = base + 1, ro = base + 1 + (int) L.size();
        n.left = lo; n.right = ro;
        dst.push_back(n);
        auto splice = [&](std::vector<Node2> &S, int off) {
            for (Node2 &c : S) { if (c.left >= 0) { c.left += off; c.right += off; } dst.push_back(c); }
        };
        splice(L, lo); splice(R, ro);
    }

    int build(uint32_t first, uint32_t count) {
        // B2_BVH_THREADS=1 forces the serial path (the result is identical; tests/test_abi_host.py checks it)
        int par_depth = 4;       // up to 16 concurrent subtrees
        if (const char *e = getenv("B2_BVH_THREADS")) { int t = atoi(e); par_depth = t <= 1 ? 0 : (t <= 2 ? 1 : (t <= 4 ? 2 : (t <= 8 ? 3 : 4))); }
        nodes.clear();
        build_parallel(*this, nodes, first, count, par_depth, 0);
        return 0;
    }
};

}  // namespace

void b2_build_bvh4(const std::vector<B2BvhTri> &tris, B2Bvh4 &out) {
    std::vector<B2BvhBox> boxes(tris.size());
    for (size_t i = 0; i < tris.size(); ++i) {
        Box bx; bx.reset();
        for (int v = 0; v < 3; ++v) bx.grow(tris[i].p[v]);
        for (int k = 0; k < 3; ++k) { boxes[i].lo[k] = bx.lo[k]; boxes[i].hi[k] = bx.hi[k]; }
    }
    b2_build_bvh4_boxes(boxes, 2, out);      // (3- and 4-triangle leaves were measured slower on the 9.8k-triangle lattice mesh)
}

uint32_t b2_bvh4_append(std::vector<float> &dst, const B2Bvh4 &src, uint32_t slot_base) {
    uint32_t node_base = (uint32_t) (dst.size() / 32);
    size_t off = dst.size();
    dst.insert(dst.end(), src.nodes.begin(), src.nodes.end());
    for (uint32_t n = 0; n < src.n_nodes; ++n)
        for (int k = 0; k < 4; ++k) {
            int32_t code; std::memcpy(&code, &dst[off + 32 * (size_t) n + 24 + k], 4);
            if (code > 0) code += (int32_t) node_base;
            else if (code < 0) {
                uint32_t u = (uint32_t) code, first = u & 0x07ffffffu;
                code = (int32_t) ((u & 0xf8000000u) | ((first + slot_base) & 0x07ffffffu));
            }
            std::memcpy(&dst[off + 32 * (size_t) n + 24 + k], &code, 4);
        }
    return node_base;
}

void b2_build_bvh4_boxes(const std::vector<B2BvhBox> &boxes, int max_leaf, B2Bvh4 &out) {
    out.nodes.clear(); out.order.clear(); out.n_nodes = 0; out.depth = 0;
    uint32_t n = (uint32_t) boxes.size();
    if (n == 0) {
        out.nodes.assign(32, 0.f); out.n_nodes = 1;
        for (int k = 0; k < 12; ++k) out.nodes[k] = std::numeric_limits<float>::infinity();
        for (int k = 12; k < 24; ++k) out.nodes[k] = -std::numeric_limits<float>::infinity();
        return;
    }
    Builder b;
    b.kMaxLeaf = max_leaf;
    b.tbox.resize(n); b.cent.resize(3 * (size_t) n); b.idx.resize(n);
    Box scene; scene.reset();
    for (uint32_t i = 0; i < n; ++i) {
        Box bx;
        for (int k = 0; k < 3; ++k) { bx.lo[k] = boxes[i].lo[k]; bx.hi[k] = boxes[i].hi[k]; }
        b.tbox[i] = bx; scene.grow(bx);
        for (int k = 0; k < 3; ++k) b.cent[3 * (size_t) i + k] = 0.5f * (bx.lo[k] + bx.hi[k]);
        b.idx[i] = i;
    }
    b.nodes.reserve(2 * (size_t) n);
    b.build(0, n);
    out.order = b.idx;

    // conservative inflation of every child box: the fp32 triangle test can accept hits a few
    // ulp outside the exact bounds; 1e-5 relative + 1e-7 of the scene size covers that.
    float diag = 0.f;
    for (int k = 0; k < 3; ++k) diag = std::max(diag, scene.hi[k] - scene.lo[k]);
    auto pad = [&](float v) { return 1e-5f * std::fabs(v) + 1e-7f * diag + 1e-30f; };

    // collapse BVH2 -> BVH4, breadth-first
    struct Item { int n2; };
    std::vector<int> queue; queue.push_back(0);
    std::vector<int> level_of; level_of.push_back(0);
    // each queue entry becomes one BVH4 node (root may be a leaf: wrap it)
    size_t head = 0;
    std::vector<float> &N = out.nodes;
    while (head < queue.size()) {
        int n2 = queue[head]; int level = level_of[head]; ++head;
        out.depth = std::max<uint32_t>(out.depth, (uint32_t) level + 1);
        int kids[4]; int nk = 0;
        if (b.nodes[n2].left < 0) kids[nk++] = n2;       // leaf root
        else { kids[nk++] = b.nodes[n2].left; kids[nk++] = b.nodes[n2].right; }
        while (nk < 4) {
            int best = -1; float ba = -1.f;
            for (int k = 0; k < nk; ++k)
                if (b.nodes[kids[k]].left >= 0 && b.nodes[kids[k]].box.area() > ba) { ba = b.nodes[kids[k]].box.area(); best = k; }
            if (best < 0) break;
            int c = kids[best];
            kids[best] = b.nodes[c].left; kids[nk++] = b.nodes[c].right;
        }
        size_t base = N.size(); N.resize(base + 32, 0.f);
        const float inf = std::numeric_limits<float>::infinity();
        for (int k = 0; k < 4; ++k) {
            int32_t code = 0;
            float lo[3] = { inf, inf, inf }, hi[3] = { -inf, -inf, -inf };
            if (k < nk) {
                const Node2 &c = b.nodes[kids[k]];
                for (int a = 0; a < 3; ++a) { lo[a] = c.box.lo[a] - pad(c.box.lo[a]); hi[a] = c.box.hi[a] + pad(c.box.hi[a]); }
                if (c.left < 0) code = (int32_t) (0x80000000u | (c.count << 27) | c.first);
                else { code = (int32_t) queue.size(); queue.push_back(kids[k]); level_of.push_back(level + 1); }
            }
            N[base + 0 + k] = lo[0]; N[base + 4 + k] = lo[1]; N[base + 8 + k] = lo[2];
            N[base + 12 + k] = hi[0]; N[base + 16 + k] = hi[1]; N[base + 20 + k] = hi[2];
            std::memcpy(&N[base + 24 + k], &code, 4);
        }
    }
    out.n_nodes = (uint32_t) (N.size() / 32);
}

#include "../../include/b2pt.h"
extern "C" int b2_host_build_bvh4(const float *tri_vertices, uint32_t n, float *nodes_out, uint32_t node_cap,
                                  uint32_t *n_nodes, uint32_t *order_out) {
    if ((n && !tri_vertices) || !n_nodes) return B2_ERR_INVALID;
    std::vector<B2BvhTri> tris(n);
    for (uint32_t i = 0; i < n; ++i) std::memcpy(tris[i].p, tri_vertices + 9 * (size_t) i, 36);
    B2Bvh4 b; b2_build_bvh4(tris, b);
    *n_nodes = b.n_nodes;
    if (nodes_out) {
        if (node_cap < b.n_nodes) return B2_ERR_INVALID;
        std::memcpy(nodes_out, b.nodes.data(), b.nodes.size() * sizeof(float));
    }
    if (order_out && n) std::memcpy(order_out, b.order.data(), n * sizeof(uint32_t));
    return B2_OK;
}
