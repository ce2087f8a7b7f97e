// b2_lbvh.cu — device BVH build: LBVH (Morton order + Karras 2012 hierarchy + bottom-up fit) collapsed to the 128-byte
// BVH4 node format of b2_device.cuh. Replaces, for large / dynamic scenes, the role of ShapeKDTree::build
// (include/mitsuba/render/kdtree.h:957-1900, built from src/librender/scene_native.inl:3-10): like any builder it only has
// to produce a VALID tree — closest hits are the lexicographic (t, -prim) minimum whatever the tree (b2_kernels.cuh), so
// films rendered with this builder and with the host SAH builder (b2_bvh.cpp) are identical bit for bit.
//
//   k_prim_bounds   primitive boxes + centroid bounds (warp reduce, ordered-int atomics)
//   k_morton        63-bit Morton code of the centroid (21 bits / axis)
//   cub::DeviceRadixSort::SortPairs (key = code, value = primitive) — library sort; not on the render path
//   k_hierarchy     Karras 2012: one thread per internal node finds its range and split (ties broken by index)
//   k_fit           leaves walk up; the second arrival at a node unions its children (one atomic per node)
//   k_select        a subtree with <= max_leaf primitives becomes a BVH4 leaf; the other internal nodes at EVEN depth
//                   become BVH4 nodes (each adopts its grandchildren), the odd ones dissolve
//   cub::DeviceScan (node numbering: deterministic, root = 0), k_emit (128-byte nodes, child boxes inflated like b2_bvh.cpp)
#include "b2_bvh.h"

#include <cub/cub.cuh>
#include <cuda_runtime.h>
#include <cstring>

namespace {

#define LB_CU(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { err = cudaGetErrorString(e_); goto fail; } } while (0)

__device__ __forceinline__ int f2ord(float f) { int i = __float_as_int(f); return i >= 0 ? i : i ^ 0x7fffffff; }
__device__ __forceinline__ float ord2f(int i) { return __int_as_float(i >= 0 ? i : i ^ 0x7fffffff); }

struct Bounds { int lo[3], hi[3]; };     // ordered-int encoded floats

__global__ void k_prim_bounds(const float *tri, uint32_t n, float4 *blo, float4 *bhi, Bounds *cb) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    float c[3] = { 0.f, 0.f, 0.f };
    bool valid = i < n;
    if (valid) {
        const float *p = tri + 9 * (size_t) i;
        float lo[3], hi[3];
        for (int k = 0; k < 3; ++k) { lo[k] = fminf(fminf(p[k], p[3 + k]), p[6 + k]); hi[k] = fmaxf(fmaxf(p[k], p[3 + k]), p[6 + k]); c[k] = 0.5f * (lo[k] + hi[k]); }
        blo[i] = make_float4(lo[0], lo[1], lo[2], 0.f); bhi[i] = make_float4(hi[0], hi[1], hi[2], 0.f);
    }
    for (int k = 0; k < 3; ++k) {
        int a = valid ? f2ord(c[k]) : 0x7fffffff, b = valid ? f2ord(c[k]) : (int) 0x80000000;
        a = __reduce_min_sync(0xffffffffu, a); b = __reduce_max_sync(0xffffffffu, b);
        if ((threadIdx.x & 31) == 0) { atomicMin(&cb->lo[k], a); atomicMax(&cb->hi[k], b); }
    }
}
// same for boxes given directly (lo.xyz, hi.xyz per primitive)
__global__ void k_box_bounds(const float *boxes, uint32_t n, float4 *blo, float4 *bhi, Bounds *cb) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    float c[3] = { 0.f, 0.f, 0.f };
    bool valid = i < n;
    if (valid) {
        const float *p = boxes + 6 * (size_t) i;
        for (int k = 0; k < 3; ++k) c[k] = 0.5f * (p[k] + p[3 + k]);
        blo[i] = make_float4(p[0], p[1], p[2], 0.f); bhi[i] = make_float4(p[3], p[4], p[5], 0.f);
    }
    for (int k = 0; k < 3; ++k) {
        int a = valid ? f2ord(c[k]) : 0x7fffffff, b = valid ? f2ord(c[k]) : (int) 0x80000000;
        a = __reduce_min_sync(0xffffffffu, a); b = __reduce_max_sync(0xffffffffu, b);
        if ((threadIdx.x & 31) == 0) { atomicMin(&cb->lo[k], a); atomicMax(&cb->hi[k], b); }
    }
}

__device__ __forceinline__ unsigned long long spread21(uint32_t v) {      // 21 bits -> every third bit
    unsigned long long x = v & 0x1fffffu;
    x = (x | x << 32) & 0x1f00000000ffffull;
    x = (x | x << 16) & 0x1f0000ff0000ffull;
    x = (x | x << 8) & 0x100f00f00f00f00full;
    x = (x | x << 4) & 0x10c30c30c30c30c3ull;
    x = (x | x << 2) & 0x1249249249249249ull;
    return x;
}

__global__ void k_morton(const float4 *blo, const float4 *bhi, uint32_t n, const Bounds *cb, unsigned long long *keys, uint32_t *vals) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float4 lo = blo[i], hi = bhi[i];
    float c[3] = { 0.5f * (lo.x + hi.x), 0.5f * (lo.y + hi.y), 0.5f * (lo.z + hi.z) };
    uint32_t q[3];
    for (int k = 0; k < 3; ++k) {
        float a = ord2f(cb->lo[k]), b = ord2f(cb->hi[k]);
        float e = b - a;
        float t = e > 0.f ? (c[k] - a) / e : 0.f;
        q[k] = (uint32_t) fminf(fmaxf(t * 2097152.f, 0.f), 2097151.f);
    }
    keys[i] = spread21(q[0]) << 2 | spread21(q[1]) << 1 | spread21(q[2]);
    vals[i] = i;
}

// delta(i, j): length of the common prefix of keys i and j (index as tie-break), -1 outside [0, n)
__device__ __forceinline__ int delta(const unsigned long long *k, uint32_t n, int i, int j) {
    if (j < 0 || j >= (int) n) return -1;
    unsigned long long a = k[i], b = k[j];
    if (a == b) return 64 + __clz((uint32_t) i ^ (uint32_t) j);
    return __clzll((long long) (a ^ b));
}

// Children are encoded as: internal node c -> c, leaf (sorted position p) -> ~p (negative)
struct Node2 { int left, right, parent; uint32_t first, last; };

__global__ void k_hierarchy(const unsigned long long *keys, uint32_t n, Node2 *nodes, int *leaf_parent) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (int) n - 1) return;
    int d = delta(keys, n, i, i + 1) - delta(keys, n, i, i - 1) >= 0 ? 1 : -1;
    int dmin = delta(keys, n, i, i - d);
    int lmax = 2;
    while (delta(keys, n, i, i + lmax * d) > dmin) lmax <<= 1;
    int l = 0;
    for (int t = lmax >> 1; t >= 1; t >>= 1)
        if (delta(keys, n, i, i + (l + t) * d) > dmin) l += t;
    int j = i + l * d;
    int dnode = delta(keys, n, i, j);
    int s = 0, t = l;
    do {
        t = (t + 1) >> 1;
        if (delta(keys, n, i, i + (s + t) * d) > dnode) s += t;
    } while (t > 1);
    int gamma = i + s * d + min(d, 0);
    int lo = min(i, j), hi = max(i, j);
    Node2 &nd = nodes[i];
    nd.first = (uint32_t) lo; nd.last = (uint32_t) hi;
    if (lo == gamma) { nd.left = ~gamma; leaf_parent[gamma] = i; } else { nd.left = gamma; nodes[gamma].parent = i; }
    if (hi == gamma + 1) { nd.right = ~(gamma + 1); leaf_parent[gamma + 1] = i; } else { nd.right = gamma + 1; nodes[gamma + 1].parent = i; }
    if (i == 0) nd.parent = -1;
}

__global__ void k_fit(const Node2 *nodes, const int *leaf_parent, const uint32_t *vals, const float4 *blo, const float4 *bhi, uint32_t n,
                      float4 *nlo, float4 *nhi, uint32_t *flags) {
    uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n) return;
    int node = leaf_parent[p];
    while (node >= 0) {
        __threadfence();
        if (atomicAdd(&flags[node], 1u) == 0u) return;          // the sibling subtree is not finished yet
        const Node2 &nd = nodes[node];
        float4 alo, ahi, clo, chi;
        if (nd.left < 0) { uint32_t v = vals[~nd.left]; alo = blo[v]; ahi = bhi[v]; } else { alo = __ldcg(&nlo[nd.left]); ahi = __ldcg(&nhi[nd.left]); }
        if (nd.right < 0) { uint32_t v = vals[~nd.right]; clo = blo[v]; chi = bhi[v]; } else { clo = __ldcg(&nlo[nd.right]); chi = __ldcg(&nhi[nd.right]); }
        __stcg(&nlo[node], make_float4(fminf(alo.x, clo.x), fminf(alo.y, clo.y), fminf(alo.z, clo.z), 0.f));
        __stcg(&nhi[node], make_float4(fmaxf(ahi.x, chi.x), fmaxf(ahi.y, chi.y), fmaxf(ahi.z, chi.z), 0.f));
        node = nd.parent;
    }
}

// selected[i] = 1: internal node i becomes a BVH4 node. depth = number of ancestors (all of them hold more primitives, so
// none of them is leaf-like): even depth -> BVH4 node, odd depth -> dissolved into its parent.
__global__ void k_select(const Node2 *nodes, uint32_t n, uint32_t max_leaf, uint32_t *selected, uint32_t *max_depth) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n - 1) return;
    const Node2 &nd = nodes[i];
    uint32_t sel = 0;
    if (nd.last - nd.first + 1u > max_leaf || i == 0) {
        uint32_t depth = 0;
        for (int p = nd.parent; p >= 0; p = nodes[p].parent) ++depth;
        if ((depth & 1u) == 0u) { sel = 1; atomicMax(max_depth, depth / 2u + 1u); }
    }
    selected[i] = sel;
}

struct Entry { float4 lo, hi; int code; };

__global__ void k_emit(const Node2 *nodes, uint32_t n, uint32_t max_leaf, const uint32_t *selected, const uint32_t *index, const uint32_t *vals,
                       const float4 *blo, const float4 *bhi, const float4 *nlo, const float4 *nhi, float pad_abs, float4 *out) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n - 1 || !selected[i]) return;
    Entry e[4]; int ne = 0;
    auto leaf_entry = [&](uint32_t first, uint32_t count, float4 lo, float4 hi) { e[ne].lo = lo; e[ne].hi = hi; e[ne].code = (int) (0x80000000u | (count << 27) | first); ++ne; };
    auto add = [&](int c, bool may_expand, auto &&self) -> void {
        if (c < 0) { uint32_t p = (uint32_t) ~c, v = vals[p]; leaf_entry(p, 1u, blo[v], bhi[v]); return; }
        const Node2 &cn = nodes[c];
        uint32_t cnt = cn.last - cn.first + 1u;
        if (cnt <= max_leaf) { leaf_entry(cn.first, cnt, nlo[c], nhi[c]); return; }
        if (may_expand) { self(cn.left, false, self); self(cn.right, false, self); return; }
        e[ne].lo = nlo[c]; e[ne].hi = nhi[c]; e[ne].code = (int) index[c]; ++ne;      // an even-depth internal node: selected
    };
    const Node2 &nd = nodes[i];
    if (nd.last - nd.first + 1u <= max_leaf) leaf_entry(nd.first, nd.last - nd.first + 1u, nlo[i], nhi[i]);      // (only the root can be leaf-like here)
    else { add(nd.left, true, add); add(nd.right, true, add); }
    float4 *N = out + 8 * (size_t) index[i];
    const float inf = __int_as_float(0x7f800000);
    float v[6][4]; int code[4];
    for (int k = 0; k < 4; ++k) {
        if (k < ne) {
            const float lo[3] = { e[k].lo.x, e[k].lo.y, e[k].lo.z }, hi[3] = { e[k].hi.x, e[k].hi.y, e[k].hi.z };
            // conservative inflation, as in b2_bvh.cpp: 1e-5 relative + 1e-7 of the scene size
            for (int a = 0; a < 3; ++a) { v[a][k] = lo[a] - (1e-5f * fabsf(lo[a]) + pad_abs); v[3 + a][k] = hi[a] + (1e-5f * fabsf(hi[a]) + pad_abs); }
            code[k] = e[k].code;
        } else { for (int a = 0; a < 3; ++a) { v[a][k] = inf; v[3 + a][k] = -inf; } code[k] = 0; }
    }
    for (int a = 0; a < 6; ++a) N[a] = make_float4(v[a][0], v[a][1], v[a][2], v[a][3]);
    N[6] = make_float4(__int_as_float(code[0]), __int_as_float(code[1]), __int_as_float(code[2]), __int_as_float(code[3]));
    N[7] = make_float4(0.f, 0.f, 0.f, 0.f);
}

template <typename T> struct Dev {
    T *p = nullptr;
    ~Dev() { if (p) cudaFree(p); }
    cudaError_t alloc(size_t n) { return cudaMalloc(&p, std::max<size_t>(n, 1) * sizeof(T)); }
};

}  // namespace

// prims: n * 9 floats (triangles, is_boxes = false) or n * 6 floats (lo, hi boxes) in HOST memory.
// Returns false (with *error set) when the device build cannot be used (no device, tree too deep, n < 2): the caller falls
// back to the host builder.
bool b2_build_bvh4_device(const float *prims, uint32_t n, bool is_boxes, int max_leaf, B2Bvh4 &out, float *ms_device, const char **error) {
    static const char *err = nullptr;
    err = nullptr;
    out.nodes.clear(); out.order.clear(); out.n_nodes = 0; out.depth = 0;
    if (ms_device) *ms_device = 0.f;
    if (n < 2) { if (error) *error = "fewer than 2 primitives"; return false; }
    const size_t in_floats = (size_t) n * (is_boxes ? 6 : 9);
    Dev<float> d_in; Dev<float4> blo, bhi, nlo, nhi, d_nodes; Dev<Bounds> cb; Dev<unsigned long long> k0, k1; Dev<uint32_t> v0, v1, flags, sel, idx, maxd;
    Dev<Node2> nodes; Dev<int> leaf_parent; Dev<unsigned char> tmp;
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    const unsigned T = 256, G = (n + T - 1) / T;
    size_t tmp_sort = 0, tmp_scan = 0;
    uint32_t n4 = 0, depth = 0;
    Bounds hb;
    float diag = 0.f;
    {
        LB_CU(d_in.alloc(in_floats)); LB_CU(blo.alloc(n)); LB_CU(bhi.alloc(n)); LB_CU(nlo.alloc(n)); LB_CU(nhi.alloc(n)); LB_CU(cb.alloc(1));
        LB_CU(k0.alloc(n)); LB_CU(k1.alloc(n)); LB_CU(v0.alloc(n)); LB_CU(v1.alloc(n)); LB_CU(flags.alloc(n)); LB_CU(sel.alloc(n)); LB_CU(idx.alloc(n)); LB_CU(maxd.alloc(1));
        LB_CU(nodes.alloc(n)); LB_CU(leaf_parent.alloc(n));
        LB_CU(cub::DeviceRadixSort::SortPairs(nullptr, tmp_sort, k0.p, k1.p, v0.p, v1.p, (int) n, 0, 63));
        LB_CU(cub::DeviceScan::ExclusiveSum(nullptr, tmp_scan, sel.p, idx.p, (int) (n - 1)));
        LB_CU(tmp.alloc(std::max(tmp_sort, tmp_scan)));
        LB_CU(cudaMemcpy(d_in.p, prims, in_floats * sizeof(float), cudaMemcpyHostToDevice));
        for (int k = 0; k < 3; ++k) { hb.lo[k] = 0x7fffffff; hb.hi[k] = (int) 0x80000000; }
        LB_CU(cudaMemcpy(cb.p, &hb, sizeof hb, cudaMemcpyHostToDevice));
        LB_CU(cudaEventCreate(&e0)); LB_CU(cudaEventCreate(&e1));
        LB_CU(cudaEventRecord(e0));
        LB_CU(cudaMemsetAsync(flags.p, 0, (size_t) n * 4)); LB_CU(cudaMemsetAsync(maxd.p, 0, 4));
        if (is_boxes) k_box_bounds<<<G, T>>>(d_in.p, n, blo.p, bhi.p, cb.p); else k_prim_bounds<<<G, T>>>(d_in.p, n, blo.p, bhi.p, cb.p);
        k_morton<<<G, T>>>(blo.p, bhi.p, n, cb.p, k0.p, v0.p);
        LB_CU(cub::DeviceRadixSort::SortPairs(tmp.p, tmp_sort, k0.p, k1.p, v0.p, v1.p, (int) n, 0, 63));
        k_hierarchy<<<G, T>>>(k1.p, n, nodes.p, leaf_parent.p);
        k_fit<<<G, T>>>(nodes.p, leaf_parent.p, v1.p, blo.p, bhi.p, n, nlo.p, nhi.p, flags.p);
        k_select<<<G, T>>>(nodes.p, n, (uint32_t) max_leaf, sel.p, maxd.p);
        LB_CU(cub::DeviceScan::ExclusiveSum(tmp.p, tmp_scan, sel.p, idx.p, (int) (n - 1)));
        {   // number of BVH4 nodes = index[n-2] + selected[n-2]
            uint32_t a = 0, b = 0;
            LB_CU(cudaMemcpy(&a, idx.p + (n - 2), 4, cudaMemcpyDeviceToHost)); LB_CU(cudaMemcpy(&b, sel.p + (n - 2), 4, cudaMemcpyDeviceToHost));
            n4 = a + b;
        }
        LB_CU(cudaMemcpy(&depth, maxd.p, 4, cudaMemcpyDeviceToHost));
        if (depth > B2_BVH_MAX_DEPTH) { err = "LBVH deeper than B2_BVH_MAX_DEPTH"; goto fail; }
        LB_CU(cudaMemcpy(&hb, cb.p, sizeof hb, cudaMemcpyDeviceToHost));
        // (the inflation constant uses the extent of the centroid bounds; the primitives' own extent is covered by the relative term)
        for (int k = 0; k < 3; ++k) {
            int lo = hb.lo[k], hi = hb.hi[k];
            float a, b; int la = lo >= 0 ? lo : lo ^ 0x7fffffff, lb = hi >= 0 ? hi : hi ^ 0x7fffffff;
            std::memcpy(&a, &la, 4); std::memcpy(&b, &lb, 4);
            diag = std::max(diag, b - a);
        }
        LB_CU(d_nodes.alloc((size_t) n4 * 8));
        k_emit<<<G, T>>>(nodes.p, n, (uint32_t) max_leaf, sel.p, idx.p, v1.p, blo.p, bhi.p, nlo.p, nhi.p, 1e-7f * diag + 1e-30f, d_nodes.p);
        LB_CU(cudaEventRecord(e1));
        LB_CU(cudaEventSynchronize(e1));
        LB_CU(cudaGetLastError());
        if (ms_device) LB_CU(cudaEventElapsedTime(ms_device, e0, e1));
        out.nodes.resize((size_t) n4 * 32); out.order.resize(n);
        LB_CU(cudaMemcpy(out.nodes.data(), d_nodes.p, (size_t) n4 * 128, cudaMemcpyDeviceToHost));
        LB_CU(cudaMemcpy(out.order.data(), v1.p, (size_t) n * 4, cudaMemcpyDeviceToHost));
        out.n_nodes = n4; out.depth = depth; out.breadth_first = false;     // (numbered by a scan over the binary nodes)
        cudaEventDestroy(e0); cudaEventDestroy(e1);
        return true;
    }
fail:
    if (e0) cudaEventDestroy(e0);
    if (e1) cudaEventDestroy(e1);
    cudaGetLastError();
    if (error) *error = err ? err : "device BVH build failed";
    out.nodes.clear(); out.order.clear(); out.n_nodes = 0; out.depth = 0;
    return false;
}
