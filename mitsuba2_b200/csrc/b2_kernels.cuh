// b2_kernels.cuh — the wavefront stage kernels (sm_100a).
//
// One wavefront iteration = TWO launches per lane set over the SoA lane queues. A lane slot keeps its
// index for its whole life (its RNG stream, work item and film pixel travel with it); between the stages
// the live lanes are COMPACTED by warp ballot into two dense index lists, so that every shading warp runs
// one kind of work on 32 lanes:
//
//   K2+K3+K5    k_trace / k_trace_inst
//                          closest hit of every lane that has a ray (BVH4 traversal or the flat leaf scan +
//                          Mesh::ray_intersect_triangle, mesh.h:194-226), any-hit visibility of the emitter
//                          samples queued by the previous k_shade (scene.cpp:203-207: unoccluded =>
//                          `result += contribution`), and the compaction: a lane whose path goes on (hit and
//                          depth < max_depth, path.cpp:131,147) is appended to the CONTINUE list, a lane whose
//                          path ends here (miss / depth limit) to the TERMINATE list — one 64-bit atomic per warp.
//                          k_trace_inst is the persistent-warp, dynamic-fetch variant for instanced scenes.
//   K4+K1+K6    k_shade    warps [0, nC/32): SI fill + one iteration of PathIntegrator::sample's loop body
//                          (path.cpp:124-208) for 32 CONTINUE lanes: emitter hit + MIS, Russian roulette,
//                          emitter sampling -> dense shadow queue (warp-ballot append), BSDF sampling -> next ray.
//                          warps [nC/32, ..): 32 TERMINATE lanes: last emitter / environment contribution,
//                          ImageBlock::put (imageblock.cpp:79-172), and regeneration of the lane with the next
//                          camera sample of its work item (render_sample, integrator.cpp:233-262;
//                          perspective.cpp:182-216) or the next work item from the global counter.
//                          A path that dies inside the loop body (roulette, zero throughput) is parked: its last
//                          emitter sample is resolved by the next k_trace, and the lane joins the next TERMINATE list.
//   (k_init = the regenerate step alone, run once to fill the lanes)
#pragma once
#include "b2_device.cuh"

struct Queues {
    float4 *ray_o, *ray_d;      // (o.xyz, mint), (d.xyz, maxt); maxt < 0: no ray to trace (retired or parked lane)   32 B
    float4 *hit;                // (t, u, v, slot bits); t = +inf: miss                           16 B
    int32_t *hit_inst;          // instance index of the hit (-1: top-level shape); NULL for flat scenes
    float4 *tp;                 // (throughput.rgb, eta)                                          16 B
    float4 *res;                // (result.rgb, pdf of the last BSDF sample)                      16 B
    uint4 *st;                  // (rng.lo, rng.hi, work item, depth|flags|samples_left)          16 B
    uint32_t *pixel;            // x | y << 16 of the lane's work item (crop-relative)             4 B
    float2 *pos;                // position_sample, only for wide (non-box) filters
    float4 *sh_o, *sh_d, *sh_c; // dense shadow queue: ray + (contrib.rgb, lane bits)
    uint8_t *cls;               // k_trace_inst -> k_classify: 0 no ray, 1 the path goes on, 2 the path ends with this segment   1 B
    uint32_t *list[2];          // per iteration parity: CONTINUE lanes from the front, TERMINATE lanes from the back   8 B
    uint32_t *cnt;              // per lane set: see CNT_*
    uint32_t *work;             // global work-item counter (shared by the lane sets)
    unsigned long long *stats;  // ST_*
    float *film;                // crop_h x crop_w x 5 (X,Y,Z,A,W)
    const uint32_t *item_pixel; // per work item of ONE pass: x | y << 16 (crop-relative)
    const uint32_t *item_seed;  // spiral_index * bs^2 + morton index
    uint32_t n_items_pass;      // work items per pass (this rank's pixels)
    uint32_t n_passes, spp_pass, block_size, n_blocks_total;
    uint32_t total_items;       // n_items_pass * n_passes
    uint32_t n_slots;
    uint64_t base_seed;
    int32_t max_depth, rr_depth;
    // non-independent samplers (stratified.cpp): per-lane (sample_index << 16 | dimension_index, permutation seed); NULL
    // for the independent sampler. The constructor's derived values (stratified.cpp:60-75) come from the host.
    uint2 *smp;
    uint32_t smp_count, smp_res, smp_jitter, smp_type, smp_res_y;      // smp_res = m_resolution.x, smp_res_y = .y (multijitter)
    float smp_inv_count, smp_inv_res, smp_inv_res_y;
    // DirectIntegrator (direct.cpp:78-196; extended kernel variants only): E emitter samples + B BSDF samples at the camera hit.
    // The emitter samples of lane L are the shadow entries [L * E, (L + 1) * E) (maxt < 0: no entry; k_trace leaves the
    // visibility in sh_c.w) and are summed in sample order; prim_* keep the camera segment, re-shaded for every BSDF sample.
    uint32_t direct_e, direct_b;
    float d_weight_bsdf, d_weight_lum, d_frac_bsdf, d_frac_lum;      // direct.cpp:96-100
    float4 *prim_o, *prim_d, *prim_hit; int32_t *prim_inst;
    // yield / resume traversal (k_trace_resume): per iteration parity the lanes whose closest-hit traversal was suspended,
    // and their saved traversal state (B2_TSAVE_WORDS words per lane); NULL when the static kernel runs
    uint32_t *resume[2];
    int4 *tsave;
    uint32_t yield_threshold;   // a warp suspends its unfinished rays when fewer lanes than this are still traversing
    // shadow queue split into sh_sub sub-queues (power of two <= B2_SH_MAXSUB; 0: one queue counted by cnt[CNT_SHADOW + parity]):
    // sub-queue k holds the entries [k * sh_cap, k * sh_cap + count_k), its counter of iteration parity p is
    // sh_cnt[(p * B2_SH_MAXSUB + k) * 32] — one 128-byte line per counter (see sh_counter())
    uint32_t *sh_cnt;
    uint32_t sh_sub, sh_cap;
};

// q.cnt words of one lane set (B2_CNT_WORDS each). Every counter that takes atomics from a whole launch has a 128-byte line of
// its own: the two lane sets' kernels run side by side, and atomics on neighbouring words serialise in one L2 sector.
enum { CNT_SHADOW = 0 /* and 1: per iteration parity — only when the shadow queue is not split (Queues::sh_sub == 0) */,
       CNT_ALIVE = 32, CNT_FETCH = 64 /* k_trace_inst's ray counter */,
       CNT_LIST = 96 /* 64-bit (n_continue | n_terminate << 32) per iteration parity: words 96 and 128 (CNT_LIST_P) */,
       CNT_RESUME = 160 /* and 161: suspended lanes per iteration parity (k_trace_resume) */ };
#define CNT_LIST_P(parity) (CNT_LIST + 32 * (parity))
#define B2_CNT_WORDS 192
#define B2_CNT_WORK (2 * B2_CNT_WORDS)      /* the work-item counter shared by the lane sets */
#define B2_CNT_TOTAL (B2_CNT_WORK + 32)
enum { ST_SAMPLES = 0, ST_SEGMENTS = 1, ST_SHADOW = 2, ST_INVALID = 3, ST_NODES = 4, ST_TRIS = 5 };
// q.stats = B2_STAT_COPIES copies of the ST_* block, one 128-byte line each; a CTA adds to copy blockIdx.x % B2_STAT_COPIES and the
// host sums the copies (every warp of every launch books its counts: on one line those reductions queue up in one L2 sector)
#define B2_STAT_COPIES 8
#define B2_STAT_STRIDE 16       /* unsigned long long words per copy */

// st.w layout
#define STW_DEPTH(w)     ((w) & 0xffu)
#define STW_ALPHA        0x100u      // camera ray hit something (valid_ray, path.cpp:121)
#define STW_DELTA        0x200u      // last BSDF sample was a delta lobe
#define STW_FILM_DX      0x400u      // box filter: sample lands on pixel x-1
#define STW_FILM_DY      0x800u      //             ... y-1
#define STW_FILM_DROP    0x1000u     // sample falls outside its block (imageblock.cpp:166)
#define STW_PARKED       0x2000u     // the path ended inside the loop body: the lane only waits for its film put
#define STW_DIRECT_EMIT  0x4000u     // DirectIntegrator: the lane's emitter samples have not been added to its result yet
#define STW_INFLIGHT     0x8000u     // k_trace_resume: the lane's closest-hit traversal is suspended (state in q.tsave)
#define STW_LEFT(w)      ((w) >> 16)

B2_D uint32_t lane_id() { return threadIdx.x & 31; }
B2_D void stat_add(const Queues &q, int which, unsigned long long v) {
    atomicAdd(q.stats + (blockIdx.x % B2_STAT_COPIES) * B2_STAT_STRIDE + which, v);
}

// warp-aggregated append: one atomic per warp, contiguous slots for the set lanes.
// Must be called by all 32 lanes of the warp.
B2_D uint32_t warp_append(uint32_t *counter, bool pred) {
    uint32_t mask = __ballot_sync(0xffffffffu, pred);
    uint32_t base = 0;
    if (lane_id() == 0 && mask) base = atomicAdd(counter, __popc(mask));
    base = __shfl_sync(0xffffffffu, base, 0);
    return base + __popc(mask & ((1u << lane_id()) - 1u));
}

// ---- the shadow queue's counters --------------------------------------------------------------------
// Every CONTINUE warp of k_shade reserves its emitter samples' slots with one atomic. On ONE counter those are ~11 000
// same-address atomics per launch, issued in bursts (the warps of a wave reach that point together): ncu shows 22 % of
// k_shade's stall samples on the first use of the atomic's result, ten times the wait k_trace sees on its (evenly spread) list
// atomics. The queue is therefore split into sh_sub sub-queues, chunk c appends to sub-queue c % sh_sub, and every counter
// sits in its own 128-byte line (adjacent words share an L2 sector and serialise just the same). Consumers see ONE dense
// index space: sh_total() entries, dense index j -> storage slot sh_index(j).
#define B2_SH_MAXSUB 16
B2_D uint32_t *sh_counter(const Queues &q, int parity, uint32_t k) {
    return q.sh_sub ? q.sh_cnt + ((uint32_t) parity * B2_SH_MAXSUB + k) * 32u : q.cnt + CNT_SHADOW + parity;
}
B2_D uint32_t sh_total(const Queues &q, int parity) {
    if (!q.sh_sub) return q.cnt[CNT_SHADOW + parity];
    uint32_t n = 0;
    for (uint32_t k = 0; k < q.sh_sub; ++k) n += *sh_counter(q, parity, k);
    return n;
}
B2_D uint32_t sh_index(const Queues &q, int parity, uint32_t j) {
    if (!q.sh_sub) return j;
    uint32_t k = 0;
    for (; k + 1u < q.sh_sub; ++k) { const uint32_t c = *sh_counter(q, parity, k); if (j < c) break; j -= c; }
    return k * q.sh_cap + j;
}
// k_trace*: block 0 resets the counters the next k_shade appends to and books the entries consumed by this launch
B2_D void sh_begin_trace(const Queues &q, int parity, uint32_t n_shadow) {
    if (blockIdx.x != 0) return;
    if (q.sh_sub) { if (threadIdx.x < q.sh_sub) *sh_counter(q, parity, threadIdx.x) = 0; }
    else if (threadIdx.x == 0) q.cnt[CNT_SHADOW + parity] = 0;
    if (threadIdx.x == 0) {
        *(unsigned long long *) (q.cnt + CNT_LIST_P(parity ^ 1)) = 0ull;   // consumed by the previous k_shade
        if (!q.direct_e) stat_add(q, ST_SHADOW, (unsigned long long) n_shadow);
    }
}

// ---- TMA (1-D bulk copy) staging of the BVH top + triangles into shared memory -----------------
B2_D uint32_t smem_u32(const void *p) { return (uint32_t) __cvta_generic_to_shared(p); }

B2_D void stage_bulk(void *smem_dst, const void *gmem_src, uint32_t bytes, uint64_t *mbar) {
    // called by ONE thread; bytes is a multiple of 16, both pointers 16-byte aligned
    uint32_t off = 0;
    while (off < bytes) {
        uint32_t chunk = bytes - off > 32768u ? 32768u : bytes - off;
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n"
                     :: "r"(smem_u32((char *) smem_dst + off)), "l"((const char *) gmem_src + off), "r"(chunk),
                        "r"(smem_u32(mbar)) : "memory");
        off += chunk;
    }
}
B2_D void mbar_init(uint64_t *mbar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" :: "r"(smem_u32(mbar)), "r"(count) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
}
B2_D void mbar_expect_tx(uint64_t *mbar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" :: "r"(smem_u32(mbar)), "r"(bytes) : "memory");
}
B2_D void mbar_wait(uint64_t *mbar, uint32_t parity) {
    asm volatile("{\n .reg .pred p;\n WAIT_%=:\n mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n @p bra DONE_%=;\n bra WAIT_%=;\n DONE_%=:\n}\n"
                 :: "r"(smem_u32(mbar)), "r"(parity) : "memory");
}

struct SmemScene { const float4 *nodes; const float4 *tris; const float4 *flat; };

// Stage sc.smem_nodes BVH nodes (+ all triangle positions if sc.smem_tris) with cp.async.bulk.
B2_D SmemScene stage_scene(const DScene &sc, unsigned char *smem_raw) {
    SmemScene s; s.nodes = nullptr; s.tris = nullptr; s.flat = nullptr;
    uint32_t node_bytes = sc.smem_nodes * 128u, tri_bytes = sc.smem_tris ? sc.n_tris * 48u : 0u, flat_bytes = sc.n_flat_leaves * 32u;
    if (node_bytes + tri_bytes == 0) return s;
    uint64_t *mbar = (uint64_t *) smem_raw;
    float4 *dst = (float4 *) (smem_raw + 16);
    if (threadIdx.x == 0) mbar_init(mbar, 1);
    __syncthreads();
    if (threadIdx.x == 0) {
        mbar_expect_tx(mbar, node_bytes + tri_bytes + flat_bytes);
        if (node_bytes) stage_bulk(dst, sc.nodes, node_bytes, mbar);
        if (tri_bytes) stage_bulk((char *) dst + node_bytes, sc.tri_isect, tri_bytes, mbar);
        if (flat_bytes) stage_bulk((char *) dst + node_bytes + tri_bytes, sc.flat_leaves, flat_bytes, mbar);
    }
    mbar_wait(mbar, 0);
    s.nodes = dst;
    if (tri_bytes) s.tris = (const float4 *) ((const char *) dst + node_bytes);
    if (flat_bytes) s.flat = (const float4 *) ((const char *) dst + node_bytes + tri_bytes);
    return s;
}

// ---- BVH4 traversal ------------------------------------------------------------------------------
// Closest hit = lexicographic min over (t, -prim) of all triangles passing the Moeller-Trumbore
// test, i.e. on exact t ties the HIGHER prim index wins (the reference's kd-tree keeps the last
// primitive of a leaf on `t <= maxt`, kdtree.h:2140-2154). The result does not depend on the
// traversal order, so the box tests below are free to be approximate as long as they are
// conservative (child boxes are inflated at build time, b2_bvh.cpp).
struct Hit { float t, u, v; uint32_t slot, prim; int inst; };

// Triangle test on the traversal copy of the geometry: tri_isect[3*s] = (p0, prim), (e1 = p1-p0, -),
// (e2 = p2-p0, -). e1/e2 are the same fp32 values Mesh::ray_intersect_triangle computes on the fly
// (mesh.h:200-203), so the result is bit-identical to tri_intersect().
B2_D bool tri_isect(float3 p0, float3 e1, float3 e2, const Ray &ray, float tmax, float &t, float &u, float &v) {
    float3 pvec = cross(ray.d, e2);
    float inv_det = b2_rcp(dot(e1, pvec));
    float3 tvec = ray.o - p0;
    u = dot(tvec, pvec) * inv_det;
    if (!(u >= 0.f && u <= 1.f)) return false;
    float3 qvec = cross(tvec, e1);
    v = dot(ray.d, qvec) * inv_det;
    if (!(v >= 0.f && u + v <= 1.f)) return false;
    t = dot(e2, qvec) * inv_det;
    return t >= ray.mint && t <= tmax;
}

// Traversal state of one ray. The loop is interruptible: run() returns early when too few lanes
// of the warp are still traversing and the caller can hand new rays to the idle lanes
// (persistent-warp dynamic ray fetch, Aila & Laine 2009).
#define B2_TRAV_DONE 0x7fffffff
// Traversal stacks are fixed-size local arrays (only the touched entries ever occupy cache lines). A BVH4 level
// pushes at most 3 entries; the host builder bounds the depth so that 3 * depth + 2 <= B2_STACK (b2_bvh.cpp).
#define B2_STACK 128
#define B2_MAX_BVH_DEPTH ((B2_STACK - 2) / 3)     /* 42 levels; the builders never exceed it (b2_bvh.h: B2_BVH_MAX_DEPTH) */
#ifndef B2_REFILL_THRESHOLD
#define B2_REFILL_THRESHOLD 22
#endif

struct Trav {
    Ray ray;
    float idx, idy, idz, oox, ooy, ooz;
    int nxi, fxi, nyi, fyi, nzi, fzi;
    int sp, cur;                // cur in [0, DONE): inner node index, < 0: leaf code, DONE: finished
    float tmax;
    Hit best;
    int inst;                   // instance being traversed (-1: top level) — part of the tie key
    bool any_hit;
    uint32_t n_nodes, n_tris;   // BVH4 nodes visited / primitives tested (only maintained by the COUNT instantiations)

    // (the traversal stack is a separate local array so that the scalar state stays in registers)
    B2_D void init(const DScene &sc, const Ray &r, int *stack) {
        best.t = B2_INF; best.u = 0.f; best.v = 0.f; best.slot = 0xffffffffu; best.prim = 0xffffffffu; best.inst = -1;
        any_hit = false; n_nodes = 0; n_tris = 0;
        restart(sc, r, stack, sc.n_top_tris ? 0 : B2_TRAV_DONE, -1);
    }
    // (re)start on another BVH (a shapegroup's, with the ray in object space) keeping the best hit so far
    B2_D void restart(const DScene &sc, const Ray &r, int *stack, int root, int instance) {
        set_ray(r); inst = instance;
        stack[0] = B2_TRAV_DONE; sp = 1;
        cur = root;
    }
    B2_D void set_ray(const Ray &r) {
        ray = r;
        tmax = fminf(r.maxt, best.t);
        // reciprocal direction for the slab tests only (approximate is fine: tests are conservative)
        float dx = fabsf(r.d.x) > 1e-20f ? r.d.x : (r.d.x < 0.f ? -1e-20f : 1e-20f);
        float dy = fabsf(r.d.y) > 1e-20f ? r.d.y : (r.d.y < 0.f ? -1e-20f : 1e-20f);
        float dz = fabsf(r.d.z) > 1e-20f ? r.d.z : (r.d.z < 0.f ? -1e-20f : 1e-20f);
        idx = __fdividef(1.f, dx); idy = __fdividef(1.f, dy); idz = __fdividef(1.f, dz);
        oox = -(r.o.x * idx); ooy = -(r.o.y * idy); ooz = -(r.o.z * idz);
        // near / far plane rows of the SoA node, selected once per ray by the direction signs
        nxi = dx < 0.f ? 3 : 0; fxi = dx < 0.f ? 0 : 3;
        nyi = dy < 0.f ? 4 : 1; fyi = dy < 0.f ? 1 : 4;
        nzi = dz < 0.f ? 5 : 2; fzi = dz < 0.f ? 2 : 5;
    }

    // while-while traversal (inner-node loop, then leaf loop) keeps the lanes of a warp in the same
    // kind of work. Returns when the ray is finished, or (can_yield) when the warp has thinned out.
    // SHP: the scene has analytic shapes (compile-time, so that plain triangle scenes do not carry the sphere / rectangle code)
    // one BVH4 node: slab tests of the 4 children, nearest hit child becomes `cur`, the others go on the stack
    template <bool COUNT>
    B2_D void node_step(const DScene &sc, const SmemScene &sm, int *stack) {
        if constexpr (COUNT) ++n_nodes;
        float4 nx, ny, nz, fx, fy, fz, chf;
        if ((uint32_t) cur < sc.smem_nodes) {
            const float4 *np = sm.nodes + 8 * cur;
            nx = np[nxi]; ny = np[nyi]; nz = np[nzi]; fx = np[fxi]; fy = np[fyi]; fz = np[fzi]; chf = np[6];
        } else {
            const float4 *gp = sc.nodes + 8 * (size_t) cur;
            nx = __ldg(gp + nxi); ny = __ldg(gp + nyi); nz = __ldg(gp + nzi);
            fx = __ldg(gp + fxi); fy = __ldg(gp + fyi); fz = __ldg(gp + fzi); chf = __ldg(gp + 6);
        }
        const float tlim = tmax * 1.0000004f;
        #define B2_SLAB(k) \
            float tn##k = fmaxf(fmaxf(b2_fma(nx.k, idx, oox), b2_fma(ny.k, idy, ooy)), fmaxf(b2_fma(nz.k, idz, ooz), ray.mint)); \
            float tf##k = fminf(fminf(b2_fma(fx.k, idx, oox), b2_fma(fy.k, idy, ooy)), fminf(b2_fma(fz.k, idz, ooz), tlim)); \
            int c##k = __float_as_int(chf.k); \
            bool h##k = c##k != 0 && tn##k <= tf##k;
        B2_SLAB(x) B2_SLAB(y) B2_SLAB(z) B2_SLAB(w)
        #undef B2_SLAB
        // nearest hit child first; the others go on the stack (unordered)
        float bd = B2_INF; int bc = 0;
        if (hx) { bd = tnx; bc = cx; }
        if (hy && tny < bd) { bd = tny; bc = cy; }
        if (hz && tnz < bd) { bd = tnz; bc = cz; }
        if (hw && tnw < bd) { bd = tnw; bc = cw; }
        if (hx && cx != bc) stack[sp++] = cx;
        if (hy && cy != bc) stack[sp++] = cy;
        if (hz && cz != bc) stack[sp++] = cz;
        if (hw && cw != bc) stack[sp++] = cw;
        cur = bc != 0 ? bc : stack[--sp];
    }
    // one leaf: its primitives against the ray; returns true when an any-hit ray is done
    template <bool ANY, bool SHP, bool COUNT>
    B2_D bool leaf_step(const DScene &sc, const SmemScene &sm, int leaf) {
        uint32_t first = (uint32_t) leaf & 0x07ffffffu, count = ((uint32_t) leaf >> 27) & 0xfu;
        if constexpr (COUNT) n_tris += count;
        for (uint32_t s = first; s < first + count; ++s) {
            float4 a0, a1, a2;
            if (sm.tris) { a0 = sm.tris[3 * s]; a1 = sm.tris[3 * s + 1]; a2 = sm.tris[3 * s + 2]; }
            else { a0 = __ldg(&sc.tri_isect[3 * (size_t) s]); a1 = __ldg(&sc.tri_isect[3 * (size_t) s + 1]); a2 = __ldg(&sc.tri_isect[3 * (size_t) s + 2]); }
            float t, u, v;
            // a slot with a non-zero kind word holds an analytic shape (rectangle / sphere) instead of a triangle
            bool hit;
            if (SHP && __float_as_uint(a1.w)) hit = shape_isect(sc.shapes[__float_as_uint(a2.w)], ray, ray.maxt, t, u, v);
            else hit = tri_isect(v3(a0), v3(a1), v3(a2), ray, tmax, t, u, v);
            if (hit) {
                if (ANY) { any_hit = true; return true; }
                uint32_t prim = __float_as_uint(a0.w);
                // tie rule: on equal t the larger (instance + 1, prim) key wins
                if (t < best.t || (t == best.t && (inst > best.inst || (inst == best.inst && prim > best.prim)))) {
                    best.t = t; best.u = u; best.v = v; best.slot = s; best.prim = prim; best.inst = inst;
                    tmax = t;
                }
            }
        }
        return false;
    }

    // while-while traversal (inner-node loop, then leaf loop) keeps the lanes of a warp in the same
    // kind of work. Returns when the ray is finished, or (can_yield) when the warp has thinned out.
    // SHP: the scene has analytic shapes (compile-time, so that plain triangle scenes do not carry the sphere / rectangle code)
    template <bool ANY, bool SHP, bool COUNT = false>
    B2_D void run(const DScene &sc, const SmemScene &sm, int *stack, bool can_yield) {
        while (cur != B2_TRAV_DONE) {
            while ((uint32_t) cur < (uint32_t) B2_TRAV_DONE) node_step<COUNT>(sc, sm, stack);
            while (cur < 0) {
                if (leaf_step<ANY, SHP, COUNT>(sc, sm, cur)) { cur = B2_TRAV_DONE; return; }
                cur = stack[--sp];
            }
            if (can_yield && __popc(__activemask()) < B2_REFILL_THRESHOLD) return;
        }
    }
};

// Tiny flat scenes (<= 32 leaves, everything in shared memory): no tree at all. Every lane scans ALL
// leaf boxes with the same slab test (uniform loop: no stack, no divergence), remembers the hit leaves
// in a bit mask, tests the nearest leaf first and then the remaining ones that are still closer than the
// best hit. The triangle tests and the tie rule are the same as in the BVH path, so the result is
// bit-identical; for the 38-triangle Cornell box this replaces ~3.2 node visits + a stack by 19 slab
// tests at full lane utilisation.
B2_D float rcp_approx(float x) { float r; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }

template <bool ANY, bool COUNT = false>
B2_D bool traverse_flat(const DScene &sc, const SmemScene &sm, const Ray &ray, Hit &best, uint32_t *counters = nullptr) {
    best.t = B2_INF; best.u = 0.f; best.v = 0.f; best.slot = 0xffffffffu; best.prim = 0xffffffffu; best.inst = -1;
    float tmax = ray.maxt;
    float dx = copysignf(fmaxf(fabsf(ray.d.x), 1e-20f), ray.d.x), dy = copysignf(fmaxf(fabsf(ray.d.y), 1e-20f), ray.d.y),
          dz = copysignf(fmaxf(fabsf(ray.d.z), 1e-20f), ray.d.z);
    const float idx = rcp_approx(dx), idy = rcp_approx(dy), idz = rcp_approx(dz);
    const float oox = -(ray.o.x * idx), ooy = -(ray.o.y * idy), ooz = -(ray.o.z * idz);
    const uint32_t n = sc.n_flat_leaves;
    if constexpr (COUNT) counters[0] += (n + 3u) / 4u;            // n leaf boxes = n / 4 BVH4 nodes' worth of box tests
    uint32_t mask = 0; float near_t = B2_INF; uint32_t near_l = 0;
    {
        const float tlim = tmax * 1.0000004f;
        #pragma unroll 4
        for (uint32_t l = 0; l < n; ++l) {
            float4 lo = sm.flat[2 * l], hi = sm.flat[2 * l + 1];
            float x0 = b2_fma(lo.x, idx, oox), x1 = b2_fma(hi.x, idx, oox), y0 = b2_fma(lo.y, idy, ooy), y1 = b2_fma(hi.y, idy, ooy),
                  z0 = b2_fma(lo.z, idz, ooz), z1 = b2_fma(hi.z, idz, ooz);
            float tn = fmaxf(fmaxf(fminf(x0, x1), fminf(y0, y1)), fmaxf(fminf(z0, z1), ray.mint));
            float tf = fminf(fminf(fmaxf(x0, x1), fmaxf(y0, y1)), fminf(fmaxf(z0, z1), tlim));
            bool h = tn <= tf;
            if (h) mask |= 1u << l;
            if (h && tn < near_t) { near_t = tn; near_l = l; }
        }
    }
    bool first = mask != 0;
    while (mask) {
        uint32_t l;
        if (first) { l = near_l; first = false; }
        else {
            l = (uint32_t) __ffs((int) mask) - 1u;
            if (!ANY && best.t < B2_INF) {      // still closer than the best hit?
                float4 lo = sm.flat[2 * l], hi = sm.flat[2 * l + 1];
                float x0 = b2_fma(lo.x, idx, oox), x1 = b2_fma(hi.x, idx, oox), y0 = b2_fma(lo.y, idy, ooy), y1 = b2_fma(hi.y, idy, ooy),
                      z0 = b2_fma(lo.z, idz, ooz), z1 = b2_fma(hi.z, idz, ooz);
                float tn = fmaxf(fmaxf(fminf(x0, x1), fminf(y0, y1)), fmaxf(fminf(z0, z1), ray.mint));
                if (tn > tmax * 1.0000004f) { mask &= ~(1u << l); continue; }
            }
        }
        mask &= ~(1u << l);
        uint32_t code = __float_as_uint(sm.flat[2 * l].w);
        uint32_t s0 = code & 0x07ffffffu, cnt = (code >> 27) & 0xfu;
        if constexpr (COUNT) counters[1] += cnt;
        for (uint32_t s = s0; s < s0 + cnt; ++s) {
            float4 a0 = sm.tris[3 * s], a1 = sm.tris[3 * s + 1], a2 = sm.tris[3 * s + 2];
            float t, u, v;
            if (tri_isect(v3(a0), v3(a1), v3(a2), ray, tmax, t, u, v)) {
                if (ANY) return true;
                uint32_t prim = __float_as_uint(a0.w);
                if (t < best.t || (t == best.t && prim > best.prim)) {
                    best.t = t; best.u = u; best.v = v; best.slot = s; best.prim = prim;
                    tmax = t;
                }
            }
        }
    }
    return best.t < B2_INF;
}

// Any-hit (Scene::ray_test) of 32 rays on the flat leaf list, WARP-COOPERATIVE. An unoccluded emitter sample has to test every
// leaf its segment's slab test lets through (no culling without a hit), and the number of such leaves differs from lane to lane:
// with one ray per lane a warp ran as many triangle-test rounds as its worst lane needs (ncu: 8 rounds at 10 / 32 lanes on the
// Cornell box). Here every lane scans the leaf boxes for its own ray (uniform, dense) and appends its (lane, leaf) candidates to
// a warp-private shared-memory list; the warp then tests the list 32 entries at a time — the ray of an entry comes from its
// owner's registers by shuffle — and an entry whose ray is already known to be occluded is skipped. Occluded <=> some candidate
// triangle passes the same Moeller-Trumbore test as in traverse_flat<true>: the result is identical.
// Must be called by all 32 lanes of a warp (`valid`: this lane holds a ray).
#define B2_COOP_ITEM_BYTES (32 * 32 * 2)                /* candidate list: 32 lanes x <= 32 leaves, 2 B each */
#define B2_COOP_BYTES_PER_WARP (B2_COOP_ITEM_BYTES + 16 + 32 * 8 + 3 * 32 * 4)   /* + occluded-lanes word, + per-lane (t, ~prim) keys, + (u, v, slot) of the winner */
template <bool COUNT>
B2_D bool flat_any_hit_coop(const DScene &sc, const SmemScene &sm, bool valid, const Ray &ray, unsigned char *coop_warp, uint32_t *counters) {
    uint16_t *items = (uint16_t *) coop_warp;
    volatile uint32_t *occ_word = (volatile uint32_t *) (coop_warp + B2_COOP_ITEM_BYTES);
    const uint32_t lane = lane_id(), n = sc.n_flat_leaves;
    uint32_t mask = 0;
    if (valid) {
        float dx = copysignf(fmaxf(fabsf(ray.d.x), 1e-20f), ray.d.x), dy = copysignf(fmaxf(fabsf(ray.d.y), 1e-20f), ray.d.y),
              dz = copysignf(fmaxf(fabsf(ray.d.z), 1e-20f), ray.d.z);
        const float idx = rcp_approx(dx), idy = rcp_approx(dy), idz = rcp_approx(dz);
        const float oox = -(ray.o.x * idx), ooy = -(ray.o.y * idy), ooz = -(ray.o.z * idz);
        const float tlim = ray.maxt * 1.0000004f;
        if constexpr (COUNT) counters[0] += (n + 3u) / 4u;
        #pragma unroll 4
        for (uint32_t l = 0; l < n; ++l) {
            float4 lo = sm.flat[2 * l], hi = sm.flat[2 * l + 1];
            float x0 = b2_fma(lo.x, idx, oox), x1 = b2_fma(hi.x, idx, oox), y0 = b2_fma(lo.y, idy, ooy), y1 = b2_fma(hi.y, idy, ooy),
                  z0 = b2_fma(lo.z, idz, ooz), z1 = b2_fma(hi.z, idz, ooz);
            float tn = fmaxf(fmaxf(fminf(x0, x1), fminf(y0, y1)), fmaxf(fminf(z0, z1), ray.mint));
            float tf = fminf(fminf(fmaxf(x0, x1), fmaxf(y0, y1)), fminf(fmaxf(z0, z1), tlim));
            if (tn <= tf) mask |= 1u << l;
        }
    }
    // the warp's candidate list: lane L's entries at [exclusive prefix sum of the counts, ...)
    const uint32_t cnt = __popc(mask);
    uint32_t incl = cnt;
    #pragma unroll
    for (int d = 1; d < 32; d <<= 1) { const uint32_t v = __shfl_up_sync(0xffffffffu, incl, d); if ((int) lane >= d) incl += v; }
    const uint32_t total = __shfl_sync(0xffffffffu, incl, 31);
    uint32_t pos = incl - cnt;
    for (uint32_t m = mask; m; m &= m - 1u) items[pos++] = (uint16_t) ((lane << 8) | ((uint32_t) __ffs((int) m) - 1u));
    if (lane == 0) *occ_word = 0u;
    __syncwarp();
    uint32_t done = 0;
    for (uint32_t base = 0; base < total; base += 32u) {
        const uint32_t k = base + lane;
        const bool have = k < total;
        const uint32_t it = have ? (uint32_t) items[k] : 0u, src = it >> 8, leaf = it & 0xffu;
        Ray r;
        r.o.x = __shfl_sync(0xffffffffu, ray.o.x, src); r.o.y = __shfl_sync(0xffffffffu, ray.o.y, src); r.o.z = __shfl_sync(0xffffffffu, ray.o.z, src);
        r.d.x = __shfl_sync(0xffffffffu, ray.d.x, src); r.d.y = __shfl_sync(0xffffffffu, ray.d.y, src); r.d.z = __shfl_sync(0xffffffffu, ray.d.z, src);
        r.mint = __shfl_sync(0xffffffffu, ray.mint, src); r.maxt = __shfl_sync(0xffffffffu, ray.maxt, src);
        bool hit = false;
        if (have && !((done >> src) & 1u)) {
            const uint32_t code = __float_as_uint(sm.flat[2 * leaf].w);
            const uint32_t s0 = code & 0x07ffffffu, c = (code >> 27) & 0xfu;
            if constexpr (COUNT) counters[1] += c;
            for (uint32_t t_ = s0; t_ < s0 + c && !hit; ++t_) {
                const float4 a0 = sm.tris[3 * t_], a1 = sm.tris[3 * t_ + 1], a2 = sm.tris[3 * t_ + 2];
                float t, u, v;
                hit = tri_isect(v3(a0), v3(a1), v3(a2), r, r.maxt, t, u, v);
            }
        }
        if (hit) atomicOr((uint32_t *) occ_word, 1u << src);
        __syncwarp();
        done = *occ_word;
    }
    return ((done >> lane) & 1u) != 0u;
}

// Closest hit (Scene::ray_intersect_preliminary) of 32 rays on the flat leaf list, WARP-COOPERATIVE after the first leaf. Every
// lane scans the leaf boxes and tests the leaf it enters first (dense: that leaf holds the hit for 93 % of the Cornell box's
// rays). The leaves a lane still has to look into after that — entered before its hit distance — are few (the per-lane loop
// tested 2.2 of them per WARP on average) but cost a whole triangle-test round each when every lane walks its own list (ncu: 3
// more rounds at 4 / 32 lanes, a quarter of the kernel's instructions). Here they go to the warp's shared-memory candidate list
// (≈ 13 entries per warp: the list is built against the hit distance after the first leaf, without the per-lane loop's
// progressive culling) and are tested 32 at a time, normally in ONE round;
// a hit competes for its ray through a 64-bit (t, ~prim) key with atomicMin in shared memory — the lexicographic (t, -prim)
// minimum of DESIGN.md 4, which does not depend on the order of the tests — and the winner leaves (u, v, slot) for the owner.
// Same triangle test, same tie rule: the result is identical to traverse_flat<false>. Called by all 32 lanes of a warp.
template <bool COUNT>
B2_D void flat_closest_hit_coop(const DScene &sc, const SmemScene &sm, bool valid, const Ray &ray, Hit &best, unsigned char *coop_warp, uint32_t *counters) {
    uint16_t *items = (uint16_t *) coop_warp;
    unsigned long long *keys = (unsigned long long *) (coop_warp + B2_COOP_ITEM_BYTES + 16);
    float *pay_u = (float *) (coop_warp + B2_COOP_ITEM_BYTES + 16 + 32 * 8), *pay_v = pay_u + 32;
    uint32_t *pay_s = (uint32_t *) (pay_v + 32);
    const uint32_t lane = lane_id(), n = sc.n_flat_leaves;
    best.t = B2_INF; best.u = 0.f; best.v = 0.f; best.slot = 0xffffffffu; best.prim = 0xffffffffu; best.inst = -1;
    float tmax = ray.maxt;
    uint32_t mask = 0, near_l = 0; float near_t = B2_INF;
    float idx = 0.f, idy = 0.f, idz = 0.f, oox = 0.f, ooy = 0.f, ooz = 0.f;
    if (valid) {
        float dx = copysignf(fmaxf(fabsf(ray.d.x), 1e-20f), ray.d.x), dy = copysignf(fmaxf(fabsf(ray.d.y), 1e-20f), ray.d.y),
              dz = copysignf(fmaxf(fabsf(ray.d.z), 1e-20f), ray.d.z);
        idx = rcp_approx(dx); idy = rcp_approx(dy); idz = rcp_approx(dz);
        oox = -(ray.o.x * idx); ooy = -(ray.o.y * idy); ooz = -(ray.o.z * idz);
        if constexpr (COUNT) counters[0] += (n + 3u) / 4u;
        const float tlim = tmax * 1.0000004f;
        #pragma unroll 4
        for (uint32_t l = 0; l < n; ++l) {
            float4 lo = sm.flat[2 * l], hi = sm.flat[2 * l + 1];
            float x0 = b2_fma(lo.x, idx, oox), x1 = b2_fma(hi.x, idx, oox), y0 = b2_fma(lo.y, idy, ooy), y1 = b2_fma(hi.y, idy, ooy),
                  z0 = b2_fma(lo.z, idz, ooz), z1 = b2_fma(hi.z, idz, ooz);
            float tn = fmaxf(fmaxf(fminf(x0, x1), fminf(y0, y1)), fmaxf(fminf(z0, z1), ray.mint));
            float tf = fminf(fminf(fmaxf(x0, x1), fmaxf(y0, y1)), fminf(fmaxf(z0, z1), tlim));
            bool h = tn <= tf;
            if (h) mask |= 1u << l;
            if (h && tn < near_t) { near_t = tn; near_l = l; }
        }
    }
    // the leaf entered first (dense)
    if (mask) {
        mask &= ~(1u << near_l);
        const uint32_t code = __float_as_uint(sm.flat[2 * near_l].w);
        const uint32_t s0 = code & 0x07ffffffu, c = (code >> 27) & 0xfu;
        if constexpr (COUNT) counters[1] += c;
        for (uint32_t s = s0; s < s0 + c; ++s) {
            const float4 a0 = sm.tris[3 * s], a1 = sm.tris[3 * s + 1], a2 = sm.tris[3 * s + 2];
            float t, u, v;
            if (tri_isect(v3(a0), v3(a1), v3(a2), ray, tmax, t, u, v)) {
                const uint32_t prim = __float_as_uint(a0.w);
                if (t < best.t || (t == best.t && prim > best.prim)) { best.t = t; best.u = u; best.v = v; best.slot = s; best.prim = prim; tmax = t; }
            }
        }
    }
    // the other leaves: those entered beyond the hit are dropped here (cheap, per lane), the rest go to the warp's list
    uint32_t keep = mask;
    if (best.t < B2_INF) {
        keep = 0;
        for (uint32_t m = mask; m; m &= m - 1u) {
            const uint32_t l = (uint32_t) __ffs((int) m) - 1u;
            float4 lo = sm.flat[2 * l], hi = sm.flat[2 * l + 1];
            float x0 = b2_fma(lo.x, idx, oox), x1 = b2_fma(hi.x, idx, oox), y0 = b2_fma(lo.y, idy, ooy), y1 = b2_fma(hi.y, idy, ooy),
                  z0 = b2_fma(lo.z, idz, ooz), z1 = b2_fma(hi.z, idz, ooz);
            float tn = fmaxf(fmaxf(fminf(x0, x1), fminf(y0, y1)), fmaxf(fminf(z0, z1), ray.mint));
            if (!(tn > tmax * 1.0000004f)) keep |= 1u << l;
        }
    }
    const uint32_t cnt = __popc(keep);
    uint32_t incl = cnt;
    #pragma unroll
    for (int d = 1; d < 32; d <<= 1) { const uint32_t v = __shfl_up_sync(0xffffffffu, incl, d); if ((int) lane >= d) incl += v; }
    const uint32_t total = __shfl_sync(0xffffffffu, incl, 31);
    if (total == 0u) return;                                   // (warp-uniform)
    uint32_t pos = incl - cnt;
    for (uint32_t m = keep; m; m &= m - 1u) items[pos++] = (uint16_t) ((lane << 8) | ((uint32_t) __ffs((int) m) - 1u));
    const unsigned long long my_key = ((unsigned long long) __float_as_uint(best.t) << 32) | (unsigned long long) (~best.prim);
    keys[lane] = my_key;                                       // (no hit yet: t = +inf)
    __syncwarp();
    for (uint32_t base = 0; base < total; base += 32u) {
        const uint32_t k = base + lane;
        const bool have = k < total;
        const uint32_t it = have ? (uint32_t) items[k] : 0u, src = it >> 8, leaf = it & 0xffu;
        Ray r;
        r.o.x = __shfl_sync(0xffffffffu, ray.o.x, src); r.o.y = __shfl_sync(0xffffffffu, ray.o.y, src); r.o.z = __shfl_sync(0xffffffffu, ray.o.z, src);
        r.d.x = __shfl_sync(0xffffffffu, ray.d.x, src); r.d.y = __shfl_sync(0xffffffffu, ray.d.y, src); r.d.z = __shfl_sync(0xffffffffu, ray.d.z, src);
        r.mint = __shfl_sync(0xffffffffu, ray.mint, src);
        const float tm = __shfl_sync(0xffffffffu, tmax, src);
        bool found = false; float bt = B2_INF, bu = 0.f, bv = 0.f; uint32_t bs = 0, bp = 0;
        if (have) {
            const uint32_t code = __float_as_uint(sm.flat[2 * leaf].w);
            const uint32_t s0 = code & 0x07ffffffu, c = (code >> 27) & 0xfu;
            if constexpr (COUNT) counters[1] += c;
            for (uint32_t s = s0; s < s0 + c; ++s) {
                const float4 a0 = sm.tris[3 * s], a1 = sm.tris[3 * s + 1], a2 = sm.tris[3 * s + 2];
                float t, u, v;
                if (tri_isect(v3(a0), v3(a1), v3(a2), r, tm, t, u, v)) {
                    const uint32_t prim = __float_as_uint(a0.w);
                    if (!found || t < bt || (t == bt && prim > bp)) { found = true; bt = t; bu = u; bv = v; bs = s; bp = prim; }
                }
            }
        }
        const unsigned long long key = ((unsigned long long) __float_as_uint(bt) << 32) | (unsigned long long) (~bp);
        if (found) atomicMin(&keys[src], key);
        __syncwarp();
        if (found && *(volatile unsigned long long *) &keys[src] == key) { pay_u[src] = bu; pay_v[src] = bv; pay_s[src] = bs; }
    }
    __syncwarp();
    const unsigned long long kf = *(volatile unsigned long long *) &keys[lane];
    if (kf != my_key) {
        best.t = __uint_as_float((uint32_t) (kf >> 32)); best.prim = ~(uint32_t) kf;
        best.u = *(volatile float *) &pay_u[lane]; best.v = *(volatile float *) &pay_v[lane]; best.slot = *(volatile uint32_t *) &pay_s[lane];
    }
    __syncwarp();                                              // (the list is reused by the warp's next chunk)
}

// Scene::ray_intersect_preliminary / ray_test: the top-level triangles, then (instanced scenes) a
// BVH4 over the instances' world bounds whose leaves restart the traversal on the shapegroup's BVH
// with the ray mapped to object space — origin as a point, direction as an UN-normalised vector,
// mint/maxt unchanged so t stays in world units (instance.cpp:107-117).
// INST is a compile-time switch so that flat scenes do not pay registers for the two-level code.
template <bool ANY, bool INST, bool SHP, bool COUNT = false>
B2_D bool traverse(const DScene &sc, const SmemScene &sm, const Ray &ray, Hit &best, uint32_t *counters = nullptr) {
    if (!INST && sc.n_flat_leaves) return traverse_flat<ANY, COUNT>(sc, sm, ray, best, counters);      // warp-uniform
    int stack[B2_STACK];
    Trav tr; tr.init(sc, ray, stack);
    tr.run<ANY, SHP, COUNT>(sc, sm, stack, false);
    if (INST && sc.n_instances && !(ANY && tr.any_hit)) {
        float dx = fabsf(ray.d.x) > 1e-20f ? ray.d.x : (ray.d.x < 0.f ? -1e-20f : 1e-20f);
        float dy = fabsf(ray.d.y) > 1e-20f ? ray.d.y : (ray.d.y < 0.f ? -1e-20f : 1e-20f);
        float dz = fabsf(ray.d.z) > 1e-20f ? ray.d.z : (ray.d.z < 0.f ? -1e-20f : 1e-20f);
        float idx = __fdividef(1.f, dx), idy = __fdividef(1.f, dy), idz = __fdividef(1.f, dz);
        float oox = -(ray.o.x * idx), ooy = -(ray.o.y * idy), ooz = -(ray.o.z * idz);
        // front-to-back over the instance boxes: nearest child first, the others are pushed with their
        // entry distance and culled on pop against the best hit found so far
        int tstack[B2_STACK]; float tdist[B2_STACK]; int tsp = 0;
        int cur = (int) sc.tlas_root;
        for (;;) {
            if (cur >= 0) {
                if constexpr (COUNT) ++tr.n_nodes;
                const float4 *gp = sc.nodes + 8 * (size_t) cur;
                float4 lox = __ldg(gp), loy = __ldg(gp + 1), loz = __ldg(gp + 2), hix = __ldg(gp + 3), hiy = __ldg(gp + 4),
                       hiz = __ldg(gp + 5), chf = __ldg(gp + 6);
                const float tlim = fminf(ray.maxt, tr.best.t) * 1.0000004f;
                float bd = B2_INF; int bc = 0;
                #define B2_TSLAB(k) { \
                    float a0 = b2_fma(lox.k, idx, oox), a1 = b2_fma(hix.k, idx, oox), b0 = b2_fma(loy.k, idy, ooy), b1 = b2_fma(hiy.k, idy, ooy), \
                          c0 = b2_fma(loz.k, idz, ooz), c1 = b2_fma(hiz.k, idz, ooz); \
                    float tn = fmaxf(fmaxf(fminf(a0, a1), fminf(b0, b1)), fmaxf(fminf(c0, c1), ray.mint)); \
                    float tf = fminf(fminf(fmaxf(a0, a1), fmaxf(b0, b1)), fminf(fmaxf(c0, c1), tlim)); \
                    int c = __float_as_int(chf.k); \
                    if (c != 0 && tn <= tf) { \
                        if (tn < bd) { if (bc != 0) { tstack[tsp] = bc; tdist[tsp] = bd; ++tsp; } bd = tn; bc = c; } \
                        else { tstack[tsp] = c; tdist[tsp] = tn; ++tsp; } } }
                B2_TSLAB(x) B2_TSLAB(y) B2_TSLAB(z) B2_TSLAB(w)
                #undef B2_TSLAB
                if (bc != 0) { cur = bc; continue; }
            } else {
                uint32_t first = (uint32_t) cur & 0x07ffffffu, count = ((uint32_t) cur >> 27) & 0xfu;
                for (uint32_t s = first; s < first + count; ++s) {
                    uint32_t ii = __ldg(&sc.inst_order[s]);
                    const DInstance &in = sc.instances[ii];
                    Ray ro; ro.o = xf_point(in.to_object, ray.o); ro.d = xf_vector(in.to_object, ray.d);
                    ro.mint = ray.mint; ro.maxt = ray.maxt;
                    tr.restart(sc, ro, stack, in.root, (int) ii);
                    tr.run<ANY, false, COUNT>(sc, sm, stack, false);       // (shapegroups hold meshes only)
                    if (ANY && tr.any_hit) { best = tr.best; if constexpr (COUNT) { counters[0] += tr.n_nodes; counters[1] += tr.n_tris; } return true; }
                }
            }
            // pop the next entry that can still contain a closer hit
            bool found = false;
            while (tsp > 0) {
                --tsp;
                if (tdist[tsp] <= fminf(ray.maxt, tr.best.t) * 1.0000004f) { cur = tstack[tsp]; found = true; break; }
            }
            if (!found) break;
        }
    }
    best = tr.best;
    if constexpr (COUNT) { counters[0] += tr.n_nodes; counters[1] += tr.n_tris; }
    return ANY ? tr.any_hit : best.t < B2_INF;
}

// ---- K6: ImageBlock::put (imageblock.cpp:79-172) -----------------------------------------------
// film word: x | y << 16 of the work item's pixel in bits 0..27 is not enough for 65535^2, so the
// word is the pixel's linear index (crop-relative) and the box-filter flags travel in `flags`.
template <bool WIDE>
B2_D bool film_put(const DScene &sc, float *film, uint32_t px, uint32_t py, uint32_t block_size,
                   uint32_t flags, float2 pos, float3 L, bool alpha) {
    // srgb_to_xyz (spectrum.h:221-228): M * rgb with fmadd accumulation (matrix.h:170-177)
    float val[5];
    val[0] = b2_fma(0.180423f, L.z, b2_fma(0.357580f, L.y, 0.412453f * L.x));
    val[1] = b2_fma(0.072169f, L.z, b2_fma(0.715160f, L.y, 0.212671f * L.x));
    val[2] = b2_fma(0.950227f, L.z, b2_fma(0.119193f, L.y, 0.019334f * L.x));
    val[3] = alpha ? 1.f : 0.f; val[4] = 1.f;
    #pragma unroll
    for (int k = 0; k < 5; ++k)                                    // imageblock.cpp:85-109
        if (!(val[k] >= -1e-5f) || !(fabsf(val[k]) < B2_INF)) return false;
    const B2Film &f = sc.film;
    int W = f.crop_size[0], H = f.crop_size[1];
    if constexpr (WIDE) {                                          // imageblock.cpp:117-161 (radius > 0.5 + RayEpsilon)
        int border = f.border;
        int ox = (int) (px / block_size) * (int) block_size, oy = (int) (py / block_size) * (int) block_size;
        int bw = min((int) block_size, W - ox), bh = min((int) block_size, H - oy);
        int sx = bw + 2 * border, sy = bh + 2 * border;
        float bx = pos.x - ((float) (ox + f.crop_offset[0] - border) + .5f),
              by = pos.y - ((float) (oy + f.crop_offset[1] - border) + .5f);
        float radius = f.filter_radius;
        int lox = max((int) ceilf(bx - radius), 0), loy = max((int) ceilf(by - radius), 0);
        int hix = min((int) floorf(bx + radius), sx - 1), hiy = min((int) floorf(by + radius), sy - 1);
        uint32_t n = (uint32_t) ceilf((radius - 2.f * B2_RAY_EPS) * 2.f);
        float basex = (float) lox - bx, basey = (float) loy - by;
        for (uint32_t yr = 0; yr < n; ++yr) {
            int y = loy + (int) yr;
            bool enabled = y <= hiy;
            int iy = min((int) fabsf((basey + (float) yr) * f.filter_scale), B2_FILTER_RESOLUTION);
            float wy = f.filter_lut[iy];
            for (uint32_t xr = 0; xr < n; ++xr) {
                int x = lox + (int) xr;
                int ix = min((int) fabsf((basex + (float) xr) * f.filter_scale), B2_FILTER_RESOLUTION);
                float weight = wy * f.filter_lut[ix];
                enabled &= x <= hix;
                int fx = ox - border + x, fy = oy - border + y;   // film (crop-relative) coordinates
                if (enabled && fx >= 0 && fy >= 0 && fx < W && fy < H) {
                    // wide filters accumulate into a padded film (8 floats / pixel: X,Y,Z,A | W,-,-,-) so that a tap is
                    // one 16-byte vector reduction + one scalar one instead of five (k_film_compact packs it at the end)
                    float *dst = film + 8 * ((size_t) fy * W + fx);
                    atomicAdd((float4 *) dst, make_float4(val[0] * weight, val[1] * weight, val[2] * weight, val[3] * weight));
                    atomicAdd(dst + 4, val[4] * weight);
                }
            }
        }
    } else {                                                       // imageblock.cpp:162-169
        if (flags & STW_FILM_DROP) return true;
        int fx = (int) px - ((flags & STW_FILM_DX) ? 1 : 0), fy = (int) py - ((flags & STW_FILM_DY) ? 1 : 0);
        float *dst = film + 5 * ((size_t) fy * W + fx);
        #pragma unroll
        for (int k = 0; k < 5; ++k) atomicAdd(dst + k, val[k]);
    }
    return true;
}

// ---- K2 + K3 + K5: one traversal launch per iteration ----------------------------------------------
// Warps stride over the lane slots: closest hit (BVH4 + Mesh::ray_intersect_triangle, mesh.h:194-226) of every
// lane that has a ray, then the COMPACTION of the live lanes by warp ballot: a lane whose path goes on (it hit
// something and depth < max_depth: path.cpp:131,147-149) is appended to the front of q.list[parity], a lane
// whose path ends with this segment to its back (one 64-bit atomic per warp reserves both ranges).
// Items [n_slots, n_slots + n_shadow) are the emitter samples queued by the previous k_shade (any hit,
// scene.cpp:203-207): unoccluded => `result += contribution` (path.cpp:171). Both kinds of ray are independent
// (a shadow entry only touches `res`, which the closest-hit part does not read), so they share one launch:
// one tail and one shared-memory staging instead of two.
B2_D void list_append(const Queues &q, int parity, uint32_t lane, bool cont, bool term, uint32_t mask) {
    // `mask`: the lanes of the warp executing this call together
    const uint32_t mc = __ballot_sync(mask, cont), mx = __ballot_sync(mask, term);
    if ((mc | mx) == 0) return;
    const int leader = __ffs(mask) - 1;
    unsigned long long base = 0;
    if ((int) lane_id() == leader)
        base = atomicAdd((unsigned long long *) (q.cnt + CNT_LIST_P(parity)), (unsigned long long) __popc(mc) | ((unsigned long long) __popc(mx) << 32));
    base = __shfl_sync(mask, base, leader);
    const uint32_t below = (1u << lane_id()) - 1u;
    if (cont) q.list[parity][(uint32_t) base + __popc(mc & below)] = lane;
    if (term) q.list[parity][q.n_slots - 1u - ((uint32_t) (base >> 32) + __popc(mx & below))] = lane;
}

#ifndef B2_TRACE_MINBLOCKS
#define B2_TRACE_MINBLOCKS 4
#endif
template <bool INST, bool SHP, bool COUNT>
__global__ void __launch_bounds__(256, INST ? 3 : B2_TRACE_MINBLOCKS) k_trace(DScene sc, Queues q, int parity) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    SmemScene sm = stage_scene(sc, smem_raw);
    const uint32_t n_shadow = sh_total(q, parity ^ 1);                    // filled by the previous k_shade
    sh_begin_trace(q, parity, n_shadow);
    uint32_t n_seg = 0;
    uint32_t counters[2] = { 0u, 0u };
    const uint32_t warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, n_warps = (gridDim.x * blockDim.x) >> 5;
    // the warp's candidate list / keys of the cooperative tests on the flat leaf list (behind the staged scene)
    unsigned char *coop = smem_raw + ((16u + sc.smem_nodes * 128u + (sc.smem_tris ? sc.n_tris * 48u : 0u) + sc.n_flat_leaves * 32u + 15u) & ~15u)
                          + (threadIdx.x >> 5) * B2_COOP_BYTES_PER_WARP;
    if (!INST && sc.n_flat_leaves) {
        for (uint32_t base = warp * 32; base < q.n_slots; base += n_warps * 32) {
            const uint32_t i = base + lane_id();
            bool has = false, cont = false, term = false;
            float4 rd = make_float4(0.f, 0.f, 0.f, -1.f), ro = rd;
            uint32_t stw = 0;
            // (all three loads are issued together: one L2 round trip per chunk instead of two)
            if (i < q.n_slots) { rd = q.ray_d[i]; ro = q.ray_o[i]; stw = q.st[i].w; has = rd.w >= 0.f; }      // (retired / parked lanes carry maxt < 0)
            Ray ray; ray.o = v3(ro); ray.d = v3(rd); ray.mint = ro.w; ray.maxt = rd.w;
            Hit h; flat_closest_hit_coop<COUNT>(sc, sm, has, ray, h, coop, counters);
            if (has) {
                q.hit[i] = make_float4(h.t, h.u, h.v, __uint_as_float(h.slot));
                ++n_seg;
                cont = h.t < B2_INF && STW_DEPTH(stw) < (uint32_t) q.max_depth;
                term = !cont;
            }
            list_append(q, parity, i, cont, term, 0xffffffffu);
        }
    } else
    for (uint32_t base = warp * 32; base < q.n_slots; base += n_warps * 32) {
        const uint32_t i = base + lane_id();
        bool cont = false, term = false;
        if (i < q.n_slots) {
            float4 rd = q.ray_d[i];
            if (rd.w >= 0.f) {                                             // (retired / parked lanes carry maxt < 0)
                float4 ro = q.ray_o[i];
                const uint32_t stw = q.st[i].w;
                Ray ray; ray.o = v3(ro); ray.d = v3(rd); ray.mint = ro.w; ray.maxt = rd.w;
                Hit h; traverse<false, INST, SHP, COUNT>(sc, sm, ray, h, counters);
                q.hit[i] = make_float4(h.t, h.u, h.v, __uint_as_float(h.slot));
                if (INST) q.hit_inst[i] = h.inst;
                ++n_seg;
                cont = h.t < B2_INF && STW_DEPTH(stw) < (uint32_t) q.max_depth;
                term = !cont;
            }
        }
        list_append(q, parity, i, cont, term, 0xffffffffu);
    }
    uint32_t n_dsh = 0;
    const uint32_t n_entries = q.direct_e ? q.n_slots * q.direct_e : n_shadow;      // (direct integrator: E entry slots per lane)
    if (!INST && sc.n_flat_leaves) {
        // flat scenes: 32 entries per warp, tested cooperatively (flat_any_hit_coop); same entry -> thread mapping as below
        for (uint32_t base = warp * 32; base < n_entries; base += n_warps * 32) {
            uint32_t j = base + lane_id();
            bool valid = j < n_entries;
            if (valid && !q.direct_e) j = sh_index(q, parity ^ 1, j);
            // the entry (ray + contribution) and the result word it may add to are fetched up front: their two dependent L2
            // round trips overlap the leaf scan instead of ending the chunk
            float4 rd = make_float4(0.f, 0.f, 0.f, -1.f), ro = rd, c = rd, r = rd;
            if (valid) { rd = q.sh_d[j]; ro = q.sh_o[j]; if (!q.direct_e) c = q.sh_c[j]; valid = !(q.direct_e && !(rd.w >= 0.f)); }
            const uint32_t lw = __float_as_uint(c.w);
            if (valid && !q.direct_e) r = q.res[lw];
            Ray ray; ray.o = v3(ro); ray.d = v3(rd); ray.mint = ro.w; ray.maxt = rd.w;
            const bool occluded = flat_any_hit_coop<COUNT>(sc, sm, valid, ray, coop, counters);
            if (!valid) continue;
            if (q.direct_e) { q.sh_c[j].w = occluded ? 0.f : 1.f; ++n_dsh; }
            else if (!occluded) {
                r.x += c.x; r.y += c.y; r.z += c.z;
                q.res[lw] = r;
            }
        }
    } else
    for (uint32_t jd = blockIdx.x * blockDim.x + threadIdx.x; jd < n_entries; jd += gridDim.x * blockDim.x) {
        const uint32_t j = q.direct_e ? jd : sh_index(q, parity ^ 1, jd);
        float4 rd = q.sh_d[j];
        if (q.direct_e && !(rd.w >= 0.f)) continue;
        float4 ro = q.sh_o[j];
        Ray ray; ray.o = v3(ro); ray.d = v3(rd); ray.mint = ro.w; ray.maxt = rd.w;
        Hit h;
        const bool occluded = traverse<true, INST, SHP, COUNT>(sc, sm, ray, h, counters);
        if (q.direct_e) { q.sh_c[j].w = occluded ? 0.f : 1.f; ++n_dsh; }
        else if (!occluded) {
            float4 c = q.sh_c[j];
            const uint32_t lw = __float_as_uint(c.w);
            float4 r = q.res[lw];
            r.x += c.x; r.y += c.y; r.z += c.z;
            q.res[lw] = r;
        }
    }
    n_seg = __reduce_add_sync(0xffffffffu, n_seg);
    if (lane_id() == 0 && n_seg) stat_add(q, ST_SEGMENTS, (unsigned long long) n_seg);
    if (q.direct_e) { n_dsh = __reduce_add_sync(0xffffffffu, n_dsh); if (lane_id() == 0 && n_dsh) stat_add(q, ST_SHADOW, (unsigned long long) n_dsh); }
    if constexpr (COUNT) {
        uint32_t a = __reduce_add_sync(0xffffffffu, counters[0]), b = __reduce_add_sync(0xffffffffu, counters[1]);
        if (lane_id() == 0) { stat_add(q, ST_NODES, (unsigned long long) a); stat_add(q, ST_TRIS, (unsigned long long) b); }
    }
}

// ---- general BVH4 scenes: yield / resume traversal across launches ------------------------------------
// The number of BVH steps of a ray is roughly exponentially distributed (ncu, matball: 7.9 node visits per ray on
// average, but 33 node rounds per warp: the longest of 32 rays), so with a static ray -> lane mapping a warp spends most
// of its rounds on a handful of lanes. Here a warp SUSPENDS its unfinished rays as soon as fewer than
// q.yield_threshold of its lanes are still traversing: their state (current node, best hit, the live part of the stack)
// goes to q.tsave, the lane is flagged STW_INFLIGHT and appended to a dense RESUME list, and the next launch of the same
// lane set continues those rays in full warps of 32 list entries. (The flag stays set until k_shade hands the lane its next
// ray: the launch that finishes a suspended ray also scans the lane slots for new rays, and must not take that one again.) A suspended lane simply takes part in a later
// iteration of the wavefront (it is on neither list until its ray is finished), so nothing else changes: its RNG stream,
// work item and film pixel stay with the lane, the hit is the same lexicographic (t, -prim) minimum, films stay bit-identical.
// Unlike the persistent-warp kernel this keeps ordinary short launches, so the other lane set's k_shade still shares the SMs.
// Any-hit rays (emitter samples) run to completion: their contribution must reach `res` before the lane's next k_shade.
#define B2_SAVE_MAX 28                          /* stack entries a suspended ray may hold (deeper: the warp keeps going) */
#define B2_TSAVE_INT4 (2 + B2_SAVE_MAX / 4)     /* int4 words of q.tsave per lane: (cur, sp, slot, prim) (t, u, v, -) + the stack */

template <bool SHP, bool COUNT>
__global__ void __launch_bounds__(256, B2_TRACE_MINBLOCKS) k_trace_resume(DScene sc, Queues q, int parity) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    SmemScene sm = stage_scene(sc, smem_raw);
    const uint32_t n_shadow = sh_total(q, parity ^ 1);                    // filled by the previous k_shade
    const uint32_t n_res = q.cnt[CNT_RESUME + (parity ^ 1)];            // suspended by the previous launch (zeroed by the next k_shade)
    sh_begin_trace(q, parity, n_shadow);
    uint32_t n_seg = 0;
    uint32_t counters[2] = { 0u, 0u };
    const uint32_t warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, n_warps = (gridDim.x * blockDim.x) >> 5;
    const uint32_t chunks_res = (n_res + 31u) >> 5, chunks_new = (q.n_slots + 31u) >> 5;
    const uint32_t thr = q.yield_threshold;
    int stack[B2_STACK];
    // the suspended rays first (they are the long ones), then the lanes that got a new ray from k_shade
    for (uint32_t c = warp; c < chunks_res + chunks_new; c += n_warps) {
        const bool resumed = c < chunks_res;
        uint32_t i = 0, stw = 0;
        bool have = false;
        if (resumed) { const uint32_t k = c * 32u + lane_id(); if (k < n_res) { i = q.resume[parity ^ 1][k]; have = true; } }
        else { i = (c - chunks_res) * 32u + lane_id(); have = i < q.n_slots; }
        Trav tr; tr.cur = B2_TRAV_DONE; tr.sp = 0; tr.n_nodes = 0; tr.n_tris = 0;
        if (have) {
            const float4 rd = q.ray_d[i];
            stw = q.st[i].w;
            have = resumed || (rd.w >= 0.f && !(stw & STW_INFLIGHT));       // (retired / parked lanes carry maxt < 0)
            if (have) {
                const float4 ro = q.ray_o[i];
                Ray ray; ray.o = v3(ro); ray.d = v3(rd); ray.mint = ro.w; ray.maxt = rd.w;
                if (!resumed) tr.init(sc, ray, stack);
                else {
                    const int4 *sv = q.tsave + (size_t) i * B2_TSAVE_INT4;
                    const int4 h0 = sv[0], h1 = sv[1];
                    tr.best.t = __int_as_float(h1.x); tr.best.u = __int_as_float(h1.y); tr.best.v = __int_as_float(h1.z);
                    tr.best.slot = (uint32_t) h0.z; tr.best.prim = (uint32_t) h0.w; tr.best.inst = -1;
                    tr.any_hit = false; tr.inst = -1;
                    tr.set_ray(ray);                                          // (tmax = min(maxt, best.t))
                    tr.cur = h0.x; tr.sp = h0.y;
                    for (int e = 0; e < tr.sp; e += 4) {
                        const int4 v = sv[2 + (e >> 2)];
                        stack[e] = v.x; stack[e + 1] = v.y; stack[e + 2] = v.z; stack[e + 3] = v.w;
                    }
                }
            }
        }
        // a chunk may suspend rays only if it started well above the threshold (the drain at the end of a render and the
        // last, partial chunk of a list run to completion)
        const bool can_yield = (uint32_t) __popc(__ballot_sync(0xffffffffu, have && tr.cur != B2_TRAV_DONE)) >= thr + 4u;
        for (;;) {
            const bool active = have && tr.cur != B2_TRAV_DONE;
            const uint32_t act = __ballot_sync(0xffffffffu, active);
            if (act == 0) break;
            if (can_yield && (uint32_t) __popc(act) < thr && !__any_sync(0xffffffffu, active && tr.sp > B2_SAVE_MAX)) break;
            if (active) {       // one while-while round: inner nodes down to a leaf, then the leaves until the next inner node
                while ((uint32_t) tr.cur < (uint32_t) B2_TRAV_DONE) tr.node_step<COUNT>(sc, sm, stack);
                while (tr.cur < 0) { tr.leaf_step<false, SHP, COUNT>(sc, sm, tr.cur); tr.cur = stack[--tr.sp]; }
            }
        }
        const bool suspended = have && tr.cur != B2_TRAV_DONE;
        bool cont = false, term = false;
        if (have && !suspended) {
            const Hit &h = tr.best;
            q.hit[i] = make_float4(h.t, h.u, h.v, __uint_as_float(h.slot));
            ++n_seg;
            cont = h.t < B2_INF && STW_DEPTH(stw) < (uint32_t) q.max_depth;
            term = !cont;
        }
        if (suspended) {
            int4 *sv = q.tsave + (size_t) i * B2_TSAVE_INT4;
            sv[0] = make_int4(tr.cur, tr.sp, (int) tr.best.slot, (int) tr.best.prim);
            sv[1] = make_int4(__float_as_int(tr.best.t), __float_as_int(tr.best.u), __float_as_int(tr.best.v), 0);
            for (int e = 0; e < tr.sp; e += 4) sv[2 + (e >> 2)] = make_int4(stack[e], stack[e + 1], stack[e + 2], stack[e + 3]);
            if (!resumed) q.st[i].w = stw | STW_INFLIGHT;
        }
        const uint32_t r = warp_append(&q.cnt[CNT_RESUME + parity], suspended);
        if (suspended) q.resume[parity][r] = i;
        list_append(q, parity, i, cont, term, 0xffffffffu);
        if constexpr (COUNT) { counters[0] += tr.n_nodes; counters[1] += tr.n_tris; }
    }
    uint32_t n_dsh = 0;
    const uint32_t n_entries = q.direct_e ? q.n_slots * q.direct_e : n_shadow;      // (direct integrator: E entry slots per lane)
    for (uint32_t jd = blockIdx.x * blockDim.x + threadIdx.x; jd < n_entries; jd += gridDim.x * blockDim.x) {
        const uint32_t j = q.direct_e ? jd : sh_index(q, parity ^ 1, jd);
        float4 rd = q.sh_d[j];
        if (q.direct_e && !(rd.w >= 0.f)) continue;
        float4 ro = q.sh_o[j];
        Ray ray; ray.o = v3(ro); ray.d = v3(rd); ray.mint = ro.w; ray.maxt = rd.w;
        Trav tr; tr.init(sc, ray, stack);
        tr.run<true, SHP, COUNT>(sc, sm, stack, false);
        const bool occluded = tr.any_hit;
        if constexpr (COUNT) { counters[0] += tr.n_nodes; counters[1] += tr.n_tris; }
        // (direct integrator: the entry stays until its lane terminates, which may be launches away when the lane's BSDF-sample
        //  ray is suspended: maxt = -2 marks it as resolved so that it is traced, and counted, once)
        if (q.direct_e) { q.sh_c[j].w = occluded ? 0.f : 1.f; q.sh_d[j].w = -2.f; ++n_dsh; }
        else if (!occluded) {
            float4 c = q.sh_c[j];
            const uint32_t lw = __float_as_uint(c.w);
            float4 r = q.res[lw];
            r.x += c.x; r.y += c.y; r.z += c.z;
            q.res[lw] = r;
        }
    }
    n_seg = __reduce_add_sync(0xffffffffu, n_seg);
    if (lane_id() == 0 && n_seg) stat_add(q, ST_SEGMENTS, (unsigned long long) n_seg);
    if (q.direct_e) { n_dsh = __reduce_add_sync(0xffffffffu, n_dsh); if (lane_id() == 0 && n_dsh) stat_add(q, ST_SHADOW, (unsigned long long) n_dsh); }
    if constexpr (COUNT) {
        uint32_t a = __reduce_add_sync(0xffffffffu, counters[0]), b = __reduce_add_sync(0xffffffffu, counters[1]);
        if (lane_id() == 0) { stat_add(q, ST_NODES, (unsigned long long) a); stat_add(q, ST_TRIS, (unsigned long long) b); }
    }
}

// ---- instanced scenes: persistent warps, ONE traversal loop over both levels, dynamic ray fetch ----
// On an instanced scene the cost of a ray varies by orders of magnitude (most camera rays of the
// BASELINE lattice miss everything, a few cross twenty instances), so a static ray -> lane mapping
// leaves ~7 of 32 lanes busy (ncu, profiles/r01). Here every lane owns a resumable traversal state:
// one stack holds the top-level triangles' BVH, the instance BVH (TLAS) and the shapegroup BVHs
// (same node format); descending into an instance pushes a POP marker and swaps the ray for its
// object-space image (instance.cpp:107-117: un-normalised direction, so t keeps its meaning), popping
// the marker swaps the world ray back (re-read from the queue). Stack entries carry their entry
// distance and are culled on pop. Whenever fewer than B2_REFILL_THRESHOLD lanes of a warp are still
// traversing, the idle lanes fetch new rays from a global counter (Aila & Laine 2009). The hit is the
// lexicographic (t, key) minimum, so the result is the same as traverse<ANY, true>() bit for bit.
#define B2_TRAV_POP 0x7ffffffe
#define B2_TRAV_NEEDPOP 0x7ffffffd
#define B2_INST_STACK (2 * B2_STACK)      /* TLAS + BLAS levels on ONE stack (+ POP markers) */
#ifndef B2_INST_MINBLOCKS
#define B2_INST_MINBLOCKS 4
#endif

template <bool SHP, bool COUNT>
__global__ void __launch_bounds__(256, B2_INST_MINBLOCKS) k_trace_inst(DScene sc, Queues q, int parity) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    SmemScene sm = stage_scene(sc, smem_raw);
    const uint32_t n_shadow = sh_total(q, parity ^ 1);
    sh_begin_trace(q, parity, n_shadow);
    const uint32_t total = q.n_slots + (q.direct_e ? q.n_slots * q.direct_e : n_shadow);
    // storage slot of the shadow entry behind work item `it` (>= n_slots)
    auto sh_slot = [&](uint32_t it) -> uint32_t { return q.direct_e ? it - q.n_slots : sh_index(q, parity ^ 1, it - q.n_slots); };
    uint32_t n_seg = 0, n_dsh = 0;
    int2 stk[B2_INST_STACK];          // (node / leaf code, entry distance bits): one 8-byte local-memory access per push / pop
    Trav tr; tr.cur = B2_TRAV_DONE; tr.sp = 0; tr.inst = -1; tr.any_hit = false; tr.tmax = 0.f; tr.n_nodes = 0; tr.n_tris = 0;
    uint32_t item = 0;
    bool busy = false, exhausted = false;
    for (;;) {
        uint32_t act = __ballot_sync(0xffffffffu, busy);
        if (!exhausted && __popc(act) < B2_REFILL_THRESHOLD) {
            const uint32_t idle = ~act, n_idle = __popc(idle);
            uint32_t base = 0;
            if (lane_id() == 0) base = atomicAdd(&q.cnt[CNT_FETCH], n_idle);
            base = __shfl_sync(0xffffffffu, base, 0);
            if (!busy) {
                uint32_t i = base + __popc(idle & ((1u << lane_id()) - 1u));
                if (i < total) {
                    float4 ro, rd;
                    if (i < q.n_slots) { rd = q.ray_d[i]; ro = q.ray_o[i]; }
                    else { const uint32_t sj = sh_slot(i); rd = q.sh_d[sj]; ro = q.sh_o[sj]; }
                    if (rd.w >= 0.f) {                                   // (retired lanes carry maxt < 0)
                        item = i; busy = true;
                        Ray ray; ray.o = v3(ro); ray.d = v3(rd); ray.mint = ro.w; ray.maxt = rd.w;
                        tr.best.t = B2_INF; tr.best.u = 0.f; tr.best.v = 0.f; tr.best.slot = 0xffffffffu; tr.best.prim = 0xffffffffu; tr.best.inst = -1;
                        tr.any_hit = false;
                        tr.set_ray(ray);
                        stk[0] = make_int2(B2_TRAV_DONE, __float_as_int(-B2_INF)); tr.sp = 1;
                        if (sc.n_top_tris) {                             // phase -2: the top-level triangles, then the TLAS
                            tr.inst = -2; tr.cur = 0;
                            if (sc.n_instances) { stk[1] = make_int2(B2_TRAV_POP, __float_as_int(-B2_INF)); tr.sp = 2; }
                        } else { tr.inst = -1; tr.cur = (int) sc.tlas_root; }
                    }
                }
            }
            if (base + n_idle >= total) exhausted = true;
            act = __ballot_sync(0xffffffffu, busy);
        }
        if (act == 0) { if (exhausted) break; continue; }
        if (busy) {
            const bool any_ray = item >= q.n_slots;
            // ---- inner nodes (either level) ----
            while ((uint32_t) tr.cur < (uint32_t) B2_TRAV_NEEDPOP) {
                if constexpr (COUNT) ++tr.n_nodes;
                float4 nx, ny, nz, fx, fy, fz, chf;
                if ((uint32_t) tr.cur < sc.smem_nodes) {
                    const float4 *np = sm.nodes + 8 * tr.cur;
                    nx = np[tr.nxi]; ny = np[tr.nyi]; nz = np[tr.nzi]; fx = np[tr.fxi]; fy = np[tr.fyi]; fz = np[tr.fzi]; chf = np[6];
                } else {
                    const float4 *gp = sc.nodes + 8 * (size_t) tr.cur;
                    nx = __ldg(gp + tr.nxi); ny = __ldg(gp + tr.nyi); nz = __ldg(gp + tr.nzi);
                    fx = __ldg(gp + tr.fxi); fy = __ldg(gp + tr.fyi); fz = __ldg(gp + tr.fzi); chf = __ldg(gp + 6);
                }
                const float tlim = tr.tmax * 1.0000004f;
                #define B2_SLAB(k) \
                    float tn##k = fmaxf(fmaxf(b2_fma(nx.k, tr.idx, tr.oox), b2_fma(ny.k, tr.idy, tr.ooy)), fmaxf(b2_fma(nz.k, tr.idz, tr.ooz), tr.ray.mint)); \
                    float tf##k = fminf(fminf(b2_fma(fx.k, tr.idx, tr.oox), b2_fma(fy.k, tr.idy, tr.ooy)), fminf(b2_fma(fz.k, tr.idz, tr.ooz), tlim)); \
                    int c##k = __float_as_int(chf.k); \
                    bool h##k = c##k != 0 && tn##k <= tf##k;
                B2_SLAB(x) B2_SLAB(y) B2_SLAB(z) B2_SLAB(w)
                #undef B2_SLAB
                float bd = B2_INF; int bc = 0;
                if (hx) { bd = tnx; bc = cx; }
                if (hy && tny < bd) { bd = tny; bc = cy; }
                if (hz && tnz < bd) { bd = tnz; bc = cz; }
                if (hw && tnw < bd) { bd = tnw; bc = cw; }
                if (hx && cx != bc) { stk[tr.sp++] = make_int2(cx, __float_as_int(tnx)); }
                if (hy && cy != bc) { stk[tr.sp++] = make_int2(cy, __float_as_int(tny)); }
                if (hz && cz != bc) { stk[tr.sp++] = make_int2(cz, __float_as_int(tnz)); }
                if (hw && cw != bc) { stk[tr.sp++] = make_int2(cw, __float_as_int(tnw)); }
                if (bc != 0) { tr.cur = bc; continue; }
                tr.cur = B2_TRAV_NEEDPOP;
            }
            // ---- one leaf (looping over consecutive leaves here was measured slower) ----
            if (tr.cur < 0) {
                const uint32_t first = (uint32_t) tr.cur & 0x07ffffffu, count = ((uint32_t) tr.cur >> 27) & 0xfu;
                if (tr.inst == -1) {
                    // TLAS leaf: enter the (first) instance, keep the others for later
                    for (uint32_t s = first + 1; s < first + count; ++s) {
                        stk[tr.sp++] = make_int2((int) (0x80000000u | (1u << 27) | s), __float_as_int(-B2_INF));
                    }
                    const uint32_t ii = __ldg(&sc.inst_order[first]);
                    const DInstance &in = sc.instances[ii];
                    Ray ro; ro.o = xf_point(in.to_object, tr.ray.o); ro.d = xf_vector(in.to_object, tr.ray.d);
                    ro.mint = tr.ray.mint; ro.maxt = tr.ray.maxt;
                    tr.set_ray(ro);
                    tr.inst = (int) ii;
                    stk[tr.sp++] = make_int2(B2_TRAV_POP, __float_as_int(-B2_INF));
                    tr.cur = in.root;
                } else {
                    const int key_inst = tr.inst < 0 ? -1 : tr.inst;
                    if constexpr (COUNT) tr.n_tris += count;
                    for (uint32_t s = first; s < first + count; ++s) {
                        float4 a0 = __ldg(&sc.tri_isect[3 * (size_t) s]), a1 = __ldg(&sc.tri_isect[3 * (size_t) s + 1]), a2 = __ldg(&sc.tri_isect[3 * (size_t) s + 2]);
                        float t, u, v;
                        bool hit;
                        if (SHP && __float_as_uint(a1.w)) hit = shape_isect(sc.shapes[__float_as_uint(a2.w)], tr.ray, tr.ray.maxt, t, u, v);
                        else hit = tri_isect(v3(a0), v3(a1), v3(a2), tr.ray, tr.tmax, t, u, v);
                        if (hit) {
                            if (any_ray) { tr.any_hit = true; break; }
                            uint32_t prim = __float_as_uint(a0.w);
                            if (t < tr.best.t || (t == tr.best.t && (key_inst > tr.best.inst || (key_inst == tr.best.inst && prim > tr.best.prim)))) {
                                tr.best.t = t; tr.best.u = u; tr.best.v = v; tr.best.slot = s; tr.best.prim = prim; tr.best.inst = key_inst;
                                tr.tmax = t;
                            }
                        }
                    }
                    tr.cur = tr.any_hit ? B2_TRAV_DONE : B2_TRAV_NEEDPOP;
                }
            }
            // ---- pop the next entry that can still hold a closer hit ----
            if (tr.cur == B2_TRAV_NEEDPOP) {
                for (;;) {
                    --tr.sp;
                    const int2 e = stk[tr.sp];
                    const int c = e.x;
                    if (c == B2_TRAV_DONE) { tr.cur = c; break; }
                    if (c == B2_TRAV_POP) {
                        if (tr.inst == -2) { tr.inst = -1; tr.cur = (int) sc.tlas_root; break; }
                        // back to world space: the ray is re-read from the queue
                        float4 ro, rd;
                        if (!any_ray) { rd = q.ray_d[item]; ro = q.ray_o[item]; }
                        else { const uint32_t sj = sh_slot(item); rd = q.sh_d[sj]; ro = q.sh_o[sj]; }
                        Ray ray; ray.o = v3(ro); ray.d = v3(rd); ray.mint = ro.w; ray.maxt = rd.w;
                        tr.set_ray(ray);
                        tr.inst = -1;
                        continue;
                    }
                    if (__int_as_float(e.y) <= tr.tmax * 1.0000004f) { tr.cur = c; break; }
                }
            }
            // ---- finished: write the result, the lane becomes idle (the compaction is k_classify's: rays finish here in
            // arbitrary order, and lists in lane order keep the gathers of k_shade on consecutive addresses) ----
            if (tr.cur == B2_TRAV_DONE) {
                busy = false;
                if (!any_ray) {
                    const Hit &h = tr.best;
                    q.hit[item] = make_float4(h.t, h.u, h.v, __uint_as_float(h.slot));
                    q.hit_inst[item] = h.inst;
                    q.cls[item] = (h.t < B2_INF && STW_DEPTH(q.st[item].w) < (uint32_t) q.max_depth) ? 1 : 2;
                    ++n_seg;
                } else if (q.direct_e) {
                    q.sh_c[sh_slot(item)].w = tr.any_hit ? 0.f : 1.f; ++n_dsh;
                } else if (!tr.any_hit) {
                    const uint32_t j = sh_slot(item);
                    float4 c = q.sh_c[j];
                    const uint32_t lw = __float_as_uint(c.w);
                    float4 r = q.res[lw];
                    r.x += c.x; r.y += c.y; r.z += c.z;
                    q.res[lw] = r;
                }
            }
        }
    }
    n_seg = __reduce_add_sync(0xffffffffu, n_seg);
    if (lane_id() == 0 && n_seg) stat_add(q, ST_SEGMENTS, (unsigned long long) n_seg);
    if (q.direct_e) { n_dsh = __reduce_add_sync(0xffffffffu, n_dsh); if (lane_id() == 0 && n_dsh) stat_add(q, ST_SHADOW, (unsigned long long) n_dsh); }
    if constexpr (COUNT) {
        uint32_t a = __reduce_add_sync(0xffffffffu, tr.n_nodes), b = __reduce_add_sync(0xffffffffu, tr.n_tris);
        if (lane_id() == 0) { stat_add(q, ST_NODES, (unsigned long long) a); stat_add(q, ST_TRIS, (unsigned long long) b); }
    }
}

// K5 for k_trace_inst: the CONTINUE / TERMINATE lists of k_trace, built in lane order from the class byte it left per lane
__global__ void __launch_bounds__(256) k_classify(Queues q, int parity) {
    const uint32_t warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, n_warps = (gridDim.x * blockDim.x) >> 5;
    for (uint32_t base = warp * 32; base < q.n_slots; base += n_warps * 32) {
        const uint32_t i = base + lane_id();
        uint32_t c = 0;
        if (i < q.n_slots) { c = q.cls[i]; if (c) q.cls[i] = 0; }
        list_append(q, parity, i, c == 1u, c == 2u, 0xffffffffu);
    }
}

// ---- samplers: IndependentSampler::next_1d / next_2d (independent.cpp:73-82) and, in the extended (EXT) kernel variants
// when q.smp is set, StratifiedSampler (stratified.cpp:92-131): permute_kensler / sample_tea_32 (random.h:75-85, 229-274)
B2_D uint32_t sample_tea_32(uint32_t v0, uint32_t v1) {
    uint32_t sum = 0;
    #pragma unroll
    for (int i = 0; i < 4; ++i) {
        sum += 0x9e3779b9u;
        v0 += ((v1 << 4) + 0xa341316cu) ^ (v1 + sum) ^ ((v1 >> 5) + 0xc8013ea4u);
        v1 += ((v0 << 4) + 0xad90777du) ^ (v0 + sum) ^ ((v0 >> 5) + 0x7e95761eu);
    }
    return v1;
}
B2_D uint32_t permute_kensler(uint32_t index, uint32_t sample_count, uint32_t seed) {
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
// LowDiscrepancySampler (ldsampler.cpp:96-131): permute (random.h:194-207; q.smp_res = log2 of the sample count),
// radical_inverse_2 / sobol_2 (qmc.h:189-227)
B2_D uint32_t sample_tea_32_r2(uint32_t v0, uint32_t v1) {
    uint32_t sum = 0;
    #pragma unroll
    for (int i = 0; i < 2; ++i) {
        sum += 0x9e3779b9u;
        v0 += ((v1 << 4) + 0xa341316cu) ^ (v1 + sum) ^ ((v1 >> 5) + 0xc8013ea4u);
        v1 += ((v0 << 4) + 0xad90777du) ^ (v0 + sum) ^ ((v0 >> 5) + 0x7e95761eu);
    }
    return v1;
}
B2_D uint32_t permute_tea(uint32_t index, uint32_t log2_count, uint32_t seed) {
    for (uint32_t level = 0; level < log2_count; ++level) {
        uint32_t bit = 1u << level;
        if ((sample_tea_32_r2(index | bit, seed) & bit) == bit) index ^= bit;
    }
    return index;
}
B2_D float radical_inverse_2(uint32_t index, uint32_t scramble) {
    index = __brev(index);
    return __uint_as_float(((index ^ scramble) >> 9) | 0x3f800000u) - 1.f;
}
B2_D float sobol_2(uint32_t index, uint32_t scramble) {
    for (uint32_t v = 1u << 31; index != 0; index >>= 1, v ^= v >> 1)
        if (index & 1u) scramble ^= v;
    return (float) scramble / 4294967296.f;
}
// OrthogonalSampler::bose (orthogonal.cpp:188-218), strength 2: q.smp_res = the prime resolution, q.smp_count = its square
B2_D float smp_bose(const Queues &q, uint64_t &rng, uint2 &sm) {
    const uint32_t j = sm.x & 0xffffu, p = sm.y, res = q.smp_res; ++sm.x;
    const uint32_t i = permute_kensler((sm.x >> 16) % q.smp_count, q.smp_count, p);
    const uint32_t a_i0 = i / res, a_i1 = i - a_i0 * res;
    uint32_t a_ij, a_ik;
    if (j == 0) { a_ij = a_i0; a_ik = a_i1; }
    else if (j == 1) { a_ij = a_i1; a_ik = a_i0; }
    else {
        const uint32_t k = (j & 1u) ? j - 1 : j + 1;
        a_ij = (a_i0 + (j - 1) * a_i1) % res;
        a_ik = (a_i0 + (k - 1) * a_i1) % res;
    }
    const uint32_t stratum = permute_kensler(a_ij, res, p * (j + 1) * 0x51633e2du);
    const uint32_t sub_stratum = permute_kensler(a_ik, res, p * (j + 1) * 0x68bc21ebu);
    const float jitter = q.smp_jitter ? pcg32_next_float(rng) : .5f;
    return ((float) stratum + ((float) sub_stratum + jitter) / (float) res) / (float) res;
}
template <bool EXT> B2_D float smp_next_1d(const Queues &q, uint64_t &rng, uint2 &sm) {
    if constexpr (EXT) if (q.smp) {
        if (q.smp_type == B2_SAMPLER_ORTHOGONAL) return smp_bose(q, rng, sm);
        if (q.smp_type == B2_SAMPLER_LOW_DISCREPANCY) {
            uint32_t i = permute_tea(sm.x >> 16, q.smp_res, sm.y + (sm.x & 0xffffu)); ++sm.x;
            return radical_inverse_2(i, sample_tea_32(sm.y, 0x48bc48ebu));
        }
        uint32_t ps = sm.y + (sm.x & 0xffffu);
        if (q.smp_type == B2_SAMPLER_MULTIJITTER) ps *= 0x45fbe943u;                                  // multijitter.cpp:108
        uint32_t p = permute_kensler(sm.x >> 16, q.smp_count, ps); ++sm.x;                            // m_dimension_index++
        float j = q.smp_jitter ? pcg32_next_float(rng) : .5f;
        return ((float) p + j) * q.smp_inv_count;
    }
    return pcg32_next_float(rng);
}
template <bool EXT> B2_D float2 smp_next_2d(const Queues &q, uint64_t &rng, uint2 &sm) {
    if constexpr (EXT) if (q.smp) {
        if (q.smp_type == B2_SAMPLER_ORTHOGONAL) { float a = smp_bose(q, rng, sm), b = smp_bose(q, rng, sm); return make_float2(a, b); }   // orthogonal.cpp:128-132
        if (q.smp_type == B2_SAMPLER_LOW_DISCREPANCY) {
            uint32_t i = permute_tea(sm.x >> 16, q.smp_res, sm.y + (sm.x & 0xffffu)); ++sm.x;
            return make_float2(radical_inverse_2(i, sample_tea_32(sm.y, 0x98bc51abu)), sobol_2(i, sample_tea_32(sm.y, 0x04223e2du)));
        }
        if (q.smp_type == B2_SAMPLER_MULTIJITTER) {                                                   // multijitter.cpp:116-144
            uint32_t ps = sm.y + (sm.x & 0xffffu);
            uint32_t sq = permute_kensler(sm.x >> 16, q.smp_count, ps * 0x51633e2du); ++sm.x;
            uint32_t cy = sq / q.smp_res, cx = sq - cy * q.smp_res;
            uint32_t sx = permute_kensler(cx, q.smp_res, ps * 0x68bc21ebu), sy = permute_kensler(cy, q.smp_res_y, ps * 0x02e5be93u);
            float jx = .5f, jy = .5f;
            if (q.smp_jitter) { jx = pcg32_next_float(rng); jy = pcg32_next_float(rng); }
            // (x + (sy + jx) * inv_res.y) * inv_res.x: the reference's compiler fuses the inner multiply-add
            return make_float2(b2_fma((float) sy + jx, q.smp_inv_res_y, (float) cx) * q.smp_inv_res,
                               b2_fma((float) sx + jy, q.smp_inv_res, (float) cy) * q.smp_inv_res_y);
        }
        uint32_t p = permute_kensler(sm.x >> 16, q.smp_count, sm.y + (sm.x & 0xffffu)); ++sm.x;
        uint32_t y = p / q.smp_res, x = p - y * q.smp_res;
        float jx = .5f, jy = .5f;
        if (q.smp_jitter) { jx = pcg32_next_float(rng); jy = pcg32_next_float(rng); }
        return make_float2(((float) x + jx) * q.smp_inv_res, ((float) y + jy) * q.smp_inv_res);
    }
    float a = pcg32_next_float(rng), b = pcg32_next_float(rng);
    return make_float2(a, b);
}

// ---- K1: regenerate — next camera sample of the lane's work item (integrator.cpp:233-262) ------
// `item`/`left`/`rng` are the lane's work-item state (left == 0: `w` is the freshly fetched work
// item). Returns false if the work is exhausted (lane retires).
template <bool EXT>
B2_D bool regen_sample(const DScene &sc, const Queues &q, uint32_t lane, uint32_t w, uint32_t &pix, uint32_t &item, uint32_t &left,
                       uint64_t &rng, uint2 &sm, Ray &ray, uint32_t &stw) {
    const B2Film &film = sc.film;
    if (left == 0) {
        if (w >= q.total_items) return false;
        item = w; left = q.spp_pass;
        uint32_t pass = item / q.n_items_pass, j = item % q.n_items_pass;
        // block_id = spiral index + (remaining_passes - 1) * block_count  (spiral.cpp:41);
        // seed = block_id * bs^2 + morton index                           (integrator.cpp:198)
        uint64_t seed_off = (uint64_t) __ldg(&q.item_seed[j]) +
                            (uint64_t) (q.n_passes - 1 - pass) * q.n_blocks_total * q.block_size * q.block_size;
        rng = pcg32_seed(q.base_seed + seed_off);
        if constexpr (EXT) sm.y = sample_tea_32((uint32_t) q.base_seed, (uint32_t) seed_off);   // compute_per_sequence_seed, sampler.cpp:55-60
        pix = __ldg(&q.item_pixel[j]);
    }
    if constexpr (EXT) sm.x = (q.spp_pass - left) << 16;                       // Sampler::advance(): m_sample_index, dimension 0
    uint32_t px = pix & 0xffffu, py = pix >> 16;
    float2 jit = smp_next_2d<EXT>(q, rng, sm);                                 // integrator.cpp:242
    float jx = jit.x, jy = jit.y;
    float psx = (float) (px + (uint32_t) film.crop_offset[0]) + jx, psy = (float) (py + (uint32_t) film.crop_offset[1]) + jy;
    (void) smp_next_1d<EXT>(q, rng, sm);                                       // :252 wavelength sample
    float ax = (psx - (float) film.crop_offset[0]) / (float) film.crop_size[0],
          ay = (psy - (float) film.crop_offset[1]) / (float) film.crop_size[1];   // :254-256
    ray = sensor_sample_ray(sc.sensor, ax, ay);
    stw = 1u | ((left - 1) << 16);
    if (!(film.filter_radius > 0.5f + B2_RAY_EPS)) {
        // box filter target pixel, in block coordinates (imageblock.cpp:112,163-166)
        uint32_t bs = q.block_size;
        int ox = (int) (px / bs) * (int) bs, oy = (int) (py / bs) * (int) bs;
        int bw = min((int) bs, film.crop_size[0] - ox), bh = min((int) bs, film.crop_size[1] - oy);
        float bx = psx - ((float) (ox + film.crop_offset[0]) + .5f), by = psy - ((float) (oy + film.crop_offset[1]) + .5f);
        int lox = (int) ceilf(bx - .5f), loy = (int) ceilf(by - .5f);
        if (lox < 0 || loy < 0 || lox >= bw || loy >= bh) stw |= STW_FILM_DROP;
        if (lox + ox != (int) px) stw |= STW_FILM_DX;
        if (loy + oy != (int) py) stw |= STW_FILM_DY;
        // (lo can only be px or px-1: psx = px + j with j in [0,1) rounds into [px, px+1])
        if (lox + ox != (int) px && lox + ox != (int) px - 1) stw |= STW_FILM_DROP;
        if (loy + oy != (int) py && loy + oy != (int) py - 1) stw |= STW_FILM_DROP;
    } else if (q.pos)
        q.pos[lane] = make_float2(psx, psy);
    return true;
}

// fill every lane with its first camera sample
template <bool EXT>
__global__ void __launch_bounds__(256) k_init(DScene sc, Queues q) {
    uint32_t warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, n_warps = (gridDim.x * blockDim.x) >> 5;
    for (uint32_t base = warp * 32; base < q.n_slots; base += n_warps * 32) {
        uint32_t lane = base + lane_id();
        bool valid = lane < q.n_slots;
        uint32_t w = warp_append(q.work, valid);
        uint32_t item = 0, left = 0, stw = 0, pix = 0; uint64_t rng = 0; Ray ray; uint2 sm = make_uint2(0u, 0u);
        bool alive = valid && regen_sample<EXT>(sc, q, lane, w, pix, item, left, rng, sm, ray, stw);
        if (alive) {
            q.ray_o[lane] = make_float4(ray.o.x, ray.o.y, ray.o.z, ray.mint);
            q.ray_d[lane] = make_float4(ray.d.x, ray.d.y, ray.d.z, ray.maxt);
            q.tp[lane] = make_float4(1.f, 1.f, 1.f, 1.f);
            q.res[lane] = make_float4(0.f, 0.f, 0.f, 1.f);
            q.st[lane] = make_uint4((uint32_t) rng, (uint32_t) (rng >> 32), item, stw);
            q.pixel[lane] = pix;
            if constexpr (EXT) if (q.smp) q.smp[lane] = sm;
        } else if (valid)
            q.ray_d[lane] = make_float4(0.f, 0.f, 0.f, -1.f);
        if constexpr (EXT) if (valid)
            for (uint32_t i = 0; i < q.direct_e; ++i) q.sh_d[lane * q.direct_e + i] = make_float4(0.f, 0.f, 0.f, -1.f);
        uint32_t n_alive = __popc(__ballot_sync(0xffffffffu, alive));
        if (lane_id() == 0 && n_alive) atomicAdd(&q.cnt[CNT_ALIVE], n_alive);
    }
}

// Small scenes: stage the shading tables (triangles, normals, CDFs, BSDF + emitter records) into
// shared memory with one bulk copy and rebase the table pointers onto the copy.
B2_D Tables stage_shade_tables(const DScene &sc, unsigned char *smem_raw) {
    Tables tb = make_tables(sc);
    if (!sc.shade_smem) return tb;
    uint64_t *mbar = (uint64_t *) smem_raw;
    unsigned char *dst = smem_raw + 16;
    if (threadIdx.x == 0) mbar_init(mbar, 1);
    __syncthreads();
    if (threadIdx.x == 0) { mbar_expect_tx(mbar, sc.blob_bytes); stage_bulk(dst, sc.blob, sc.blob_bytes, mbar); }
    mbar_wait(mbar, 0);
    const unsigned char *g = sc.blob;
    tb.tri_pos = (const float4 *) (dst + ((const unsigned char *) sc.tri_pos - g));
    tb.tri_nrm = (const float4 *) (dst + ((const unsigned char *) sc.tri_nrm - g));
    tb.prim_to_slot = (const uint32_t *) (dst + ((const unsigned char *) sc.prim_to_slot - g));
    tb.cdf = (const float *) (dst + ((const unsigned char *) sc.cdf - g));
    tb.bsdfs = (const B2Bsdf *) (dst + ((const unsigned char *) sc.bsdfs - g));
    tb.emitters = (const DEmitter *) (dst + ((const unsigned char *) sc.emitters - g));
    if (sc.tri_uv) tb.tri_uv = (const float4 *) (dst + ((const unsigned char *) sc.tri_uv - g));
    if (sc.textures) tb.textures = (const B2Texture *) (dst + ((const unsigned char *) sc.textures - g));
    return tb;
}

// ---- K4 + K1 + K6: shade ---------------------------------------------------------------------------
#ifndef B2_SHADE_MINBLOCKS
#define B2_SHADE_MINBLOCKS 3
#endif
#ifndef B2_SHADE_THREADS
#define B2_SHADE_THREADS 256        /* largest CTA k_shade may be launched with (b2_api.cu picks 128 or 256 per scene) */
#endif

// DirectIntegrator (direct.cpp:167-176): BSDF samples b, b + 1, ... of the camera hit until one has a non-zero value (its ray is
// traced next) or the B samples are spent. `si` / `bsdf` / `flip` as bsdf_side() left them.
template <uint32_t F>
B2_D bool direct_bsdf_samples(const Tables &tb, const Queues &q, const SI &si, const B2Bsdf &bsdf, bool flip, uint32_t &b,
                              uint64_t &rng, uint2 &sm, float3 &bsdf_val, BS &bs) {
    for (; b < q.direct_b; ++b) {
        float2 s2 = smp_next_2d<true>(q, rng, sm);                             // (g++: next_2d before next_1d, direct.cpp:168-169)
        float s1 = smp_next_1d<true>(q, rng, sm);
        bsdf_val = bsdf_sample<F>(bsdf, si, s1, s2.x, s2.y, bs, bsdf_reflectance<F>(tb, bsdf, si));
        if (flip) bs.wo.z = -bs.wo.z;
        if (any_nonzero(bsdf_val)) return true;
    }
    return false;
}
// the lane's E emitter samples, in sample order, once k_trace has resolved their visibility (direct.cpp:161)
B2_D float3 direct_add_emitter_samples(const Queues &q, uint32_t lane, float3 result) {
    for (uint32_t i = 0; i < q.direct_e; ++i) {
        const uint32_t j = lane * q.direct_e + i;
        const float4 d4 = q.sh_d[j];
        if (d4.w >= 0.f || d4.w == -2.f) {                                  // (-2: resolved by k_trace_resume)
            const float4 c = q.sh_c[j];
            if (c.w != 0.f) result = result + v3(c);
            q.sh_d[j] = make_float4(0.f, 0.f, 0.f, -1.f);
        }
    }
    return result;
}

// 32 CONTINUE lanes: one iteration of PathIntegrator::sample's loop body (path.cpp:124-208). Called by whole warps.
template <uint32_t F>
B2_D void shade_continue(const DScene &sc, const Tables &tb, const Queues &q, int parity, bool valid, uint32_t lane, uint32_t chunk) {
    constexpr bool EXT = (F & F_TEX) != 0;
    // all queue loads are issued up front (one round trip)
    const uint32_t ld = lane;        // (callers pass slot 0 for the padding lanes of the last warp)
    float4 rd = q.ray_d[ld], ro = q.ray_o[ld], h = q.hit[ld], tpe = q.tp[ld], rs = q.res[ld];
    uint4 st = q.st[ld];
    bool park = false, shadow = false, flip = false, direct_mode = false;
    float4 sho, shd, shc;
    float4 nro, nrd;                 // next ray
    uint64_t rng = (uint64_t) st.x | ((uint64_t) st.y << 32);
    uint2 sm = make_uint2(0u, 0u);
    if constexpr (EXT) if (q.smp) sm = q.smp[ld];
    const int depth = (int) STW_DEPTH(st.w);
    float3 throughput = v3(tpe), result = v3(rs);
    float eta = tpe.w, bs_pdf = rs.w;
    SI si; si.emitter = -1; si.bsdf = 0; si.p = si.sh_n = v3(0.f, 0.f, 0.f);
    // ---- phase 1: SI fill, emitter hit, Russian roulette, emitter sampling (the lane hit something and depth < max_depth) ----
    if (valid) {
        si_fill<EXT>(tb, __float_as_uint(h.w), h.x, h.y, h.z, v3(ro), v3(rd), q.hit_inst ? q.hit_inst[lane] : -1, si);
        if (depth == 1) st.w |= STW_ALPHA;                                       // path.cpp:121
        // ---- intersection with emitters (path.cpp:128-129) ----
        if (si.emitter >= 0) {
            float emission_weight = 1.f;
            if (depth > 1) {                                                    // path.cpp:194-205, records.h:168-174
                float emitter_pdf = 0.f;
                if (!(st.w & STW_DELTA)) {
                    float3 dd = si.p - v3(ro);
                    float dist = sqrtf(dot(dd, dd));
                    dd = dd * b2_rcp(dist);
                    emitter_pdf = scene_pdf_emitter_direction<F>(tb, sc.env, si.emitter, v3(ro), dd, si.sh_n, dist);
                }
                emission_weight = mis_weight(bs_pdf, emitter_pdf);
            }
            const DEmitter &em = tb.emitters[si.emitter];
            float3 ev = si.wi.z > 0.f ? v3(em.radiance[0], em.radiance[1], em.radiance[2]) : v3(0.f, 0.f, 0.f);   // area.cpp:63-71
            result = result + (throughput * emission_weight) * ev;
        }
        bool active = true;                                                     // path.cpp:131
        if (depth > q.rr_depth) {                                               // path.cpp:137-141
            float qq = b2_min(hmax(throughput) * b2_sqr(eta), .95f);
            active = smp_next_1d<EXT>(q, rng, sm) < qq;
            throughput = throughput * b2_rcp(qq);
        }
        bool direct = false;
        if constexpr (EXT) direct = (q.direct_e | q.direct_b) != 0;
        direct_mode = direct;
        if (direct) {
            if constexpr (EXT) {        // DirectIntegrator::sample at the camera hit (direct.cpp:128-176)
                bool smooth;
                const B2Bsdf &bsdf = bsdf_side<F>(tb, si, flip, smooth);
                for (uint32_t i = 0; i < q.direct_e; ++i) {                     // emitter sampling, direct.cpp:135-163
                    const uint32_t j = lane * q.direct_e + i;
                    float4 e_d = make_float4(0.f, 0.f, 0.f, -1.f);
                    if (smooth) {
                        float2 se = smp_next_2d<true>(q, rng, sm);
                        DS ds; float3 emitter_val = scene_sample_emitter_direction<F>(tb, sc.env, si.p, se.x, se.y, ds);
                        if (ds.pdf != 0.f) {
                            float3 wo = to_local(si, ds.d);
                            if (flip) wo.z = -wo.z;
                            float3 bsdf_val = bsdf_eval<F>(bsdf, si, wo, bsdf_reflectance<F>(tb, bsdf, si));
                            float bpdf = bsdf_pdf<F>(bsdf, si, wo);
                            float mis = ds.delta ? 1.f : mis_weight(ds.pdf * q.d_frac_lum, bpdf * q.d_frac_bsdf) * q.d_weight_lum;
                            float3 c = (bsdf_val * mis) * emitter_val;
                            if (any_nonzero(c)) {
                                q.sh_o[j] = make_float4(si.p.x, si.p.y, si.p.z, B2_RAY_EPS * (1.f + hmax_abs(si.p)));
                                q.sh_c[j] = make_float4(c.x, c.y, c.z, 0.f);
                                e_d = make_float4(ds.d.x, ds.d.y, ds.d.z, ds.dist * (1.f - B2_SHADOW_EPS));
                            }
                        }
                    }
                    q.sh_d[j] = e_d;
                }
                if (q.direct_e) st.w |= STW_DIRECT_EMIT;
                q.prim_o[lane] = ro; q.prim_d[lane] = rd; q.prim_hit[lane] = h;
                if (q.hit_inst) q.prim_inst[lane] = q.hit_inst[lane];
                uint32_t b = 0; float3 bsdf_val; BS bs;                          // BSDF sampling, direct.cpp:167-176
                if (!direct_bsdf_samples<F>(tb, q, si, tb.bsdfs[si.bsdf], flip, b, rng, sm, bsdf_val, bs)) park = true;
                else {
                    throughput = bsdf_val; eta = __uint_as_float(b);               // (tp.w carries the index of the sample in flight)
                    float3 nd = to_world(si, bs.wo);
                    nro = make_float4(si.p.x, si.p.y, si.p.z, (1.f + hmax_abs(si.p)) * B2_RAY_EPS);
                    nrd = make_float4(nd.x, nd.y, nd.z, B2_INF);
                    bs_pdf = bs.pdf;
                    st.w = (st.w & ~(0xffu | STW_DELTA)) | 2u | (bs.delta ? STW_DELTA : 0u);
                }
            }
        } else if (!active) park = true;                                        // path.cpp:147-149
        else {
            bool smooth;
            const B2Bsdf &bsdf = bsdf_side<F>(tb, si, flip, smooth);            // (twosided: side record, wi.z mirrored)
            if (smooth) {                                                          // path.cpp:153-172
                float2 se = smp_next_2d<EXT>(q, rng, sm); float sx = se.x, sy = se.y;
                DS ds; float3 emitter_val = scene_sample_emitter_direction<F>(tb, sc.env, si.p, sx, sy, ds);
                if (ds.pdf != 0.f) {
                    float3 wo = to_local(si, ds.d);
                    if constexpr (EXT) if (flip) wo.z = -wo.z;
                    float3 bsdf_val = bsdf_eval<F>(bsdf, si, wo, bsdf_reflectance<F>(tb, bsdf, si));
                    float bpdf = bsdf_pdf<F>(bsdf, si, wo);
                    float mis = mis_weight(ds.pdf, bpdf);
                    if constexpr (EXT) if (ds.delta) mis = 1.f;                // path.cpp:170
                    float3 c = ((throughput * mis) * bsdf_val) * emitter_val;
                    if (any_nonzero(c)) {                                       // visibility: scene.cpp:203-207
                        shadow = true;
                        sho = make_float4(si.p.x, si.p.y, si.p.z, B2_RAY_EPS * (1.f + hmax_abs(si.p)));
                        shd = make_float4(ds.d.x, ds.d.y, ds.d.z, ds.dist * (1.f - B2_SHADOW_EPS));
                        shc = make_float4(c.x, c.y, c.z, __uint_as_float(lane));
                    }
                }
            }
        }
    }
    __syncwarp();
    // ---- K5: warp-ballot compaction of the emitter samples into the dense shadow queue. The atomic is issued
    // here and its result is only consumed after phase 2 (latency hidden). ----
    const uint32_t sh_mask = __ballot_sync(0xffffffffu, shadow);
    uint32_t sq_base = 0;
    const uint32_t sh_k = q.sh_sub ? chunk & (q.sh_sub - 1u) : 0u;           // this chunk's sub-queue (see sh_counter())
    if (lane_id() == 0 && sh_mask) sq_base = atomicAdd(sh_counter(q, parity, sh_k), __popc(sh_mask));
    // ---- phase 2: BSDF sampling (the reference (g++) draws next_2d BEFORE next_1d, path.cpp:177-178) ----
    if (valid && !park && !direct_mode) {
        const B2Bsdf &bsdf = tb.bsdfs[si.bsdf];
        float2 s2 = smp_next_2d<EXT>(q, rng, sm); float s2x = s2.x, s2y = s2.y;
        float s1 = smp_next_1d<EXT>(q, rng, sm);
        BS bs; float3 bsdf_val = bsdf_sample<F>(bsdf, si, s1, s2x, s2y, bs, bsdf_reflectance<F>(tb, bsdf, si));
        if constexpr (EXT) if (flip) bs.wo.z = -bs.wo.z;                       // twosided.cpp:122
        throughput = throughput * bsdf_val;
        if (!any_nonzero(throughput)) park = true;                             // path.cpp:182-184
        else {
            eta *= bs.eta;
            float3 nd = to_world(si, bs.wo);                                    // interaction.h:58-61
            nro = make_float4(si.p.x, si.p.y, si.p.z, (1.f + hmax_abs(si.p)) * B2_RAY_EPS);
            nrd = make_float4(nd.x, nd.y, nd.z, B2_INF);
            bs_pdf = bs.pdf;
            st.w = (st.w & ~(0xffu | STW_DELTA)) | (uint32_t) (depth + 1) | (bs.delta ? STW_DELTA : 0u);
        }
    }
    __syncwarp();
    // ---- a path that ended here is parked: no ray, the next k_trace resolves its last emitter sample, and the lane
    // joins the TERMINATE list of the next iteration (appended behind the entries the next k_trace adds) ----
    const uint32_t pk_mask = __ballot_sync(0xffffffffu, park);
    if (pk_mask) {
        unsigned long long pb = 0;
        if (lane_id() == 0) pb = atomicAdd((unsigned long long *) (q.cnt + CNT_LIST_P(parity ^ 1)), (unsigned long long) __popc(pk_mask) << 32);
        pb = __shfl_sync(0xffffffffu, pb, 0);
        if (park) q.list[parity ^ 1][q.n_slots - 1u - ((uint32_t) (pb >> 32) + __popc(pk_mask & ((1u << lane_id()) - 1u)))] = lane;
    }
    if (sh_mask) {
        const uint32_t sq = sh_k * q.sh_cap + __shfl_sync(0xffffffffu, sq_base, 0) + __popc(sh_mask & ((1u << lane_id()) - 1u));
        if (shadow) { q.sh_o[sq] = sho; q.sh_d[sq] = shd; q.sh_c[sq] = shc; }
    }
    if (valid) {
        if (park) { nrd = make_float4(0.f, 0.f, 0.f, -1.f); st.w |= STW_PARKED; }
        else q.ray_o[lane] = nro;
        q.ray_d[lane] = nrd;
        q.tp[lane] = make_float4(throughput.x, throughput.y, throughput.z, eta);
        q.res[lane] = make_float4(result.x, result.y, result.z, bs_pdf);
        q.st[lane] = make_uint4((uint32_t) rng, (uint32_t) (rng >> 32), st.z, st.w & ~STW_INFLIGHT);   // (see k_trace_resume)
        if constexpr (EXT) if (q.smp) q.smp[lane] = sm;
    }
}

// 32 TERMINATE lanes: the path ends with this segment (miss, or a hit at the depth limit: path.cpp:128-149 up to the
// break), or it was parked by shade_continue. Film put (ImageBlock::put) and regeneration. Called by whole warps.
// DirectIntegrator (extended variants): a lane whose BSDF-sample ray came back adds that sample's contribution and, while
// BSDF samples are left, re-shades the camera hit and leaves with the next sample's ray instead of finishing.
template <uint32_t F>
B2_D void shade_terminate(const DScene &sc, const Tables &tb, const Queues &q, bool valid, uint32_t lane, uint32_t &n_put, uint32_t &n_bad) {
    constexpr bool EXT = (F & F_TEX) != 0;
    const uint32_t ld = lane;        // (callers pass slot 0 for the padding lanes of the last warp)
    uint4 st = q.st[ld];
    float4 h = q.hit[ld], tpe = q.tp[ld], rs = q.res[ld];
    uint32_t pix = q.pixel[ld];
    uint32_t item = st.z, left = STW_LEFT(st.w);
    bool direct = false;
    if constexpr (EXT) direct = (q.direct_e | q.direct_b) != 0;
    // the work-item fetch is issued first: its latency hides behind the emitter evaluation and the film put
    // (direct integrator: only once it is known that the lane finishes its sample)
    uint32_t w = 0;
    if (!direct) w = warp_append(q.work, valid && left == 0);
    uint64_t rng = (uint64_t) st.x | ((uint64_t) st.y << 32);
    uint2 sm = make_uint2(0u, 0u);
    if constexpr (EXT) if (q.smp) sm = q.smp[ld];
    float3 result = v3(rs);
    bool finishing = valid;
    if constexpr (EXT) if (direct && valid && (st.w & STW_DIRECT_EMIT)) {
        result = direct_add_emitter_samples(q, lane, result);
        st.w &= ~STW_DIRECT_EMIT;
    }
    if (valid && !(st.w & STW_PARKED)) {
        const int depth = (int) STW_DEPTH(st.w);
        const bool hit_valid = h.x < B2_INF;
        const float3 throughput = v3(tpe);
        const float bs_pdf = rs.w;
        int emitter = -1;
        if (hit_valid) emitter = (int) __float_as_uint(B2_LD(&tb.tri_nrm[3 * __float_as_uint(h.w) + 1]).w);
        else if constexpr ((F & F_ENV) != 0) emitter = sc.env.index;              // scene->environment(), scene.h:243-253
        if (depth == 1 && hit_valid) st.w |= STW_ALPHA;                           // path.cpp:121
        if (emitter >= 0) {                                                        // path.cpp:128-129
            const float4 rd = q.ray_d[lane], ro = q.ray_o[lane];
            SI si; si.emitter = emitter; si.p = si.sh_n = v3(0.f, 0.f, 0.f);
            if (hit_valid) si_fill<EXT>(tb, __float_as_uint(h.w), h.x, h.y, h.z, v3(ro), v3(rd), q.hit_inst ? q.hit_inst[lane] : -1, si);
            float emitter_pdf = 0.f;
            if (depth > 1 && !(st.w & STW_DELTA)) {                               // path.cpp:194-205, records.h:168-174
                float3 dd = si.p - v3(ro);
                float dist = sqrtf(dot(dd, dd));
                dd = dd * b2_rcp(dist);
                if (!hit_valid) dd = v3(rd);                                      // d[!it.is_valid()] = -it.wi
                emitter_pdf = scene_pdf_emitter_direction<F>(tb, sc.env, si.emitter, v3(ro), dd, si.sh_n, dist);
            }
            float3 ev;
            bool miss = false;
            if constexpr ((F & F_ENV) != 0) miss = !hit_valid;
            if (miss) ev = env_eval(sc.env, v3(rd));                              // envmap.cpp:134-147
            else {
                const DEmitter &em = tb.emitters[si.emitter];
                ev = si.wi.z > 0.f ? v3(em.radiance[0], em.radiance[1], em.radiance[2]) : v3(0.f, 0.f, 0.f);   // area.cpp:63-71
            }
            if (direct && depth > 1)                                              // direct.cpp:180-194
                result = result + ((throughput * ev) * mis_weight(bs_pdf * q.d_frac_bsdf, emitter_pdf * q.d_frac_lum)) * q.d_weight_bsdf;
            else {
                const float emission_weight = depth > 1 ? mis_weight(bs_pdf, emitter_pdf) : 1.f;
                result = result + (throughput * emission_weight) * ev;
            }
        }
        // Russian roulette still draws its sample before the depth test ends the path (path.cpp:137-149): the sampler
        // stream of the work item continues into the next camera sample
        if (depth > q.rr_depth && hit_valid) (void) smp_next_1d<EXT>(q, rng, sm);
        if constexpr (EXT) if (direct && depth > 1) {
            // the next BSDF sample of the camera hit (direct.cpp:167-176), re-shaded from the stored camera segment
            uint32_t b = __float_as_uint(tpe.w) + 1u;
            if (b < q.direct_b) {
                const float4 po = q.prim_o[lane], pd = q.prim_d[lane], ph = q.prim_hit[lane];
                SI si; si.emitter = -1; si.bsdf = 0; si.p = si.sh_n = v3(0.f, 0.f, 0.f);
                si_fill<EXT>(tb, __float_as_uint(ph.w), ph.x, ph.y, ph.z, v3(po), v3(pd), q.hit_inst ? q.prim_inst[lane] : -1, si);
                bool flip, smooth;
                (void) bsdf_side<F>(tb, si, flip, smooth);
                float3 bsdf_val; BS bs;
                if (direct_bsdf_samples<F>(tb, q, si, tb.bsdfs[si.bsdf], flip, b, rng, sm, bsdf_val, bs)) {
                    finishing = false;
                    float3 nd = to_world(si, bs.wo);
                    q.ray_o[lane] = make_float4(si.p.x, si.p.y, si.p.z, (1.f + hmax_abs(si.p)) * B2_RAY_EPS);
                    q.ray_d[lane] = make_float4(nd.x, nd.y, nd.z, B2_INF);
                    q.tp[lane] = make_float4(bsdf_val.x, bsdf_val.y, bsdf_val.z, __uint_as_float(b));
                    q.res[lane] = make_float4(result.x, result.y, result.z, bs.pdf);
                    st.w = (st.w & ~STW_DELTA) | (bs.delta ? STW_DELTA : 0u);
                    q.st[lane] = make_uint4((uint32_t) rng, (uint32_t) (rng >> 32), st.z, st.w & ~STW_INFLIGHT);
                    if (q.smp) q.smp[lane] = sm;
                }
            }
        }
    }
    if (direct) w = warp_append(q.work, finishing && left == 0);
    // ---- K6: ImageBlock::put ----
    if (finishing) {
        float2 pos = q.pos ? q.pos[lane] : make_float2(0.f, 0.f);
        bool ok = film_put<(F & F_WIDE) != 0>(sc, q.film, pix & 0xffffu, pix >> 16, q.block_size, st.w, pos, result, (st.w & STW_ALPHA) != 0);
        ++n_put; n_bad += ok ? 0 : 1;
    }
    // ---- K1: regenerate ----
    bool retire = false;
    if (finishing) {
        Ray ray; uint32_t stw;
        if (regen_sample<EXT>(sc, q, lane, w, pix, item, left, rng, sm, ray, stw)) {
            q.ray_o[lane] = make_float4(ray.o.x, ray.o.y, ray.o.z, ray.mint);
            q.ray_d[lane] = make_float4(ray.d.x, ray.d.y, ray.d.z, ray.maxt);
            q.tp[lane] = make_float4(1.f, 1.f, 1.f, 1.f);
            q.res[lane] = make_float4(0.f, 0.f, 0.f, 1.f);
            q.st[lane] = make_uint4((uint32_t) rng, (uint32_t) (rng >> 32), item, stw);
            q.pixel[lane] = pix;
            if constexpr (EXT) if (q.smp) q.smp[lane] = sm;
        } else { retire = true; q.ray_d[lane] = make_float4(0.f, 0.f, 0.f, -1.f); }
    }
    const uint32_t n_ret = __popc(__ballot_sync(0xffffffffu, retire));
    if (lane_id() == 0 && n_ret) atomicSub(&q.cnt[CNT_ALIVE], n_ret);
}

template <uint32_t F>
__global__ void __launch_bounds__(B2_SHADE_THREADS, B2_SHADE_MINBLOCKS) k_shade(DScene sc, Queues q, int parity) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const uint32_t warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, n_warps = (gridDim.x * blockDim.x) >> 5;
    const Tables tb = stage_shade_tables(sc, smem_raw);
    // the two dense lists written by k_trace (+ the lanes parked by the previous k_shade)
    const unsigned long long n_list = *(const unsigned long long *) (q.cnt + CNT_LIST_P(parity));
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        q.cnt[CNT_FETCH] = 0;                                             // ray counter of the next k_trace_inst
        if (q.tsave) q.cnt[CNT_RESUME + (parity ^ 1)] = 0;                // consumed by the k_trace_resume before this launch
    }
    const uint32_t n_cont = (uint32_t) n_list, n_term = (uint32_t) (n_list >> 32);
    const uint32_t wc = (n_cont + 31u) >> 5, wt = (n_term + 31u) >> 5;
    uint32_t n_put = 0, n_bad = 0;
    // lane slot of this thread in chunk w: the w-th group of 32 entries of the CONTINUE list, then of the TERMINATE list
    auto list_entry = [&](uint32_t w, bool &valid) -> uint32_t {
        if (w < wc) { const uint32_t k = w * 32u + lane_id(); valid = k < n_cont; return valid ? q.list[parity][k] : 0u; }
        const uint32_t k = (w - wc) * 32u + lane_id(); valid = k < n_term;
        return valid ? q.list[parity][q.n_slots - 1u - k] : 0u;
    };
    for (uint32_t wi = warp; wi < wc + wt; wi += n_warps) {
        bool valid;
        const uint32_t slot = list_entry(wi, valid);
        if (wi < wc) shade_continue<F>(sc, tb, q, parity, valid, slot, wi);
        else shade_terminate<F>(sc, tb, q, valid, slot, n_put, n_bad);
    }
    n_put = __reduce_add_sync(0xffffffffu, n_put); n_bad = __reduce_add_sync(0xffffffffu, n_bad);
    if (lane_id() == 0) {
        if (n_put) stat_add(q, ST_SAMPLES, (unsigned long long) n_put);
        if (n_bad) stat_add(q, ST_INVALID, (unsigned long long) n_bad);
    }
}

// padded accumulation film of the wide-filter path -> the XYZAW film of the C-ABI
__global__ void k_film_compact(const float *film8, float *film5, size_t n) {
    size_t i = (size_t) blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float4 a = *(const float4 *) (film8 + 8 * i); float w = film8[8 * i + 4];
    float *d = film5 + 5 * i;
    d[0] = a.x; d[1] = a.y; d[2] = a.z; d[3] = a.w; d[4] = w;
}

// ---- HDRFilm::bitmap(): XYZAW -> RGB (hdrfilm.cpp:251-322, struct.cpp:1750-1812) ----------------
__global__ void k_develop(const float *xyzaw, float *rgb, int n) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float *p = xyzaw + 5 * (size_t) i;
    float iw = 1.f / p[4];
    // the blend of StructConverter's compiled loop: first product rounded, the others fused (struct.cpp:1380-1386: muls, then
    // fmadd231 per further channel), then the multiplication by 1 / weight (struct.cpp:1398-1403)
    float r = 3.240479f * p[0], g = -0.969256f * p[0], b = 0.055648f * p[0];
    r = b2_fma(-1.537150f, p[1], r); r = b2_fma(-0.498535f, p[2], r);
    g = b2_fma(1.875991f, p[1], g); g = b2_fma(0.041556f, p[2], g);
    b = b2_fma(-0.204043f, p[1], b); b = b2_fma(1.057311f, p[2], b);
    rgb[3 * i] = r * iw; rgb[3 * i + 1] = g * iw; rgb[3 * i + 2] = b * iw;
}

// ---- stage-level kernels for b2_stage_* (flat arrays, identity lane mapping) --------------------
__global__ void k_stage_raygen(DScene sc, uint32_t n, const uint32_t *px, const uint32_t *py, const uint64_t *seed_off,
                               uint64_t base_seed, float4 *ray_o, float4 *ray_d, uint64_t *rng_out, float2 *pos) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const B2Film &film = sc.film;
    uint64_t rng = pcg32_seed(base_seed + seed_off[i]);
    float jx = pcg32_next_float(rng), jy = pcg32_next_float(rng);
    float psx = (float) px[i] + jx, psy = (float) py[i] + jy;
    (void) pcg32_next_float(rng);
    float ax = (psx - (float) film.crop_offset[0]) / (float) film.crop_size[0],
          ay = (psy - (float) film.crop_offset[1]) / (float) film.crop_size[1];
    Ray ray = sensor_sample_ray(sc.sensor, ax, ay);
    ray_o[i] = make_float4(ray.o.x, ray.o.y, ray.o.z, ray.mint);
    ray_d[i] = make_float4(ray.d.x, ray.d.y, ray.d.z, ray.maxt);
    rng_out[i] = rng; pos[i] = make_float2(psx, psy);
}
template <bool INST>
__global__ void __launch_bounds__(256) k_stage_intersect(DScene sc, uint32_t n, const float4 *ray_o, const float4 *ray_d, float4 *hit, int32_t *inst) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    SmemScene sm = stage_scene(sc, smem_raw);
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        float4 ro = ray_o[i], rd = ray_d[i];
        Ray ray; ray.o = v3(ro); ray.d = v3(rd); ray.mint = ro.w; ray.maxt = rd.w;
        Hit h; traverse<false, INST, true>(sc, sm, ray, h);
        hit[i] = make_float4(h.t, h.u, h.v, __uint_as_float(h.prim));     // prim id (not slot) at the ABI
        inst[i] = h.inst;
    }
}
template <bool INST>
__global__ void __launch_bounds__(256) k_stage_occluded(DScene sc, uint32_t n, const float4 *ray_o, const float4 *ray_d, uint8_t *occ) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    SmemScene sm = stage_scene(sc, smem_raw);
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        float4 ro = ray_o[i], rd = ray_d[i];
        Ray ray; ray.o = v3(ro); ray.d = v3(rd); ray.mint = ro.w; ray.maxt = rd.w;
        Hit h; occ[i] = traverse<true, INST, true>(sc, sm, ray, h) ? 1 : 0;
    }
}
// b2_stage_compact: K5 alone — the classification k_trace ends with, on caller data (ray present / hit distance / depth)
__global__ void __launch_bounds__(256) k_stage_compact(Queues q, uint32_t n, const float *ray_maxt, const float *hit_t, const uint32_t *depth) {
    const uint32_t warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, n_warps = (gridDim.x * blockDim.x) >> 5;
    for (uint32_t base = warp * 32; base < n; base += n_warps * 32) {
        const uint32_t i = base + lane_id();
        bool cont = false, term = false;
        if (i < n && ray_maxt[i] >= 0.f) {
            cont = hit_t[i] < B2_INF && STW_DEPTH(depth[i]) < (uint32_t) q.max_depth;
            term = !cont;
        }
        list_append(q, 0, i, cont, term, 0xffffffffu);
    }
}
// b2_stage_shade: the C-ABI's (t, u, v, prim) hit -> the queues' (t, u, v, slot); lane state words -> st
__global__ void k_stage_pack(DScene sc, uint32_t n, float4 *hit, const uint64_t *rng, const uint32_t *depth_flags, uint4 *st, uint32_t *list) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float4 h = hit[i];
    if (h.x < B2_INF) h.w = __uint_as_float(sc.prim_to_slot[__float_as_uint(h.w)]);
    hit[i] = h;
    st[i] = make_uint4((uint32_t) rng[i], (uint32_t) (rng[i] >> 32), 0u, depth_flags[i] & 0xffffu);
    list[i] = i;
}
__global__ void k_stage_unpack(uint32_t n, const uint4 *st, uint64_t *rng, uint32_t *depth_flags) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint4 s = st[i];
    rng[i] = (uint64_t) s.x | ((uint64_t) s.y << 32); depth_flags[i] = s.w & 0xffffu;
}
// b2_stage_film_put: ImageBlock::put of n samples; the box-filter target is derived like regen_sample does
template <bool WIDE>
__global__ void k_stage_film_put(DScene sc, uint32_t n, uint32_t block_size, const uint32_t *px, const uint32_t *py, const float2 *pos,
                                 const float *rgb, const uint8_t *alpha, float *film, uint32_t *n_bad) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const B2Film &f = sc.film;
    uint32_t flags = 0;
    float2 p = pos[i];
    if (!WIDE) {
        int ox = (int) (px[i] / block_size) * (int) block_size, oy = (int) (py[i] / block_size) * (int) block_size;
        int bw = min((int) block_size, f.crop_size[0] - ox), bh = min((int) block_size, f.crop_size[1] - oy);
        float bx = p.x - ((float) (ox + f.crop_offset[0]) + .5f), by = p.y - ((float) (oy + f.crop_offset[1]) + .5f);
        int lox = (int) ceilf(bx - .5f), loy = (int) ceilf(by - .5f);
        if (lox < 0 || loy < 0 || lox >= bw || loy >= bh) flags |= STW_FILM_DROP;
        if (lox + ox != (int) px[i]) flags |= STW_FILM_DX;
        if (loy + oy != (int) py[i]) flags |= STW_FILM_DY;
        if (lox + ox != (int) px[i] && lox + ox != (int) px[i] - 1) flags |= STW_FILM_DROP;
        if (loy + oy != (int) py[i] && loy + oy != (int) py[i] - 1) flags |= STW_FILM_DROP;
    }
    bool ok = film_put<WIDE>(sc, film, px[i], py[i], block_size, flags, p, v3(rgb[3 * i], rgb[3 * i + 1], rgb[3 * i + 2]), alpha[i] != 0);
    if (!ok) atomicAdd(n_bad, 1u);
}
__global__ void k_stage_si_fill(DScene sc, uint32_t n, const float4 *ray_o, const float4 *ray_d, const float4 *hit, const int32_t *inst, float *out) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float4 h = hit[i];
    float *o = out + 16 * (size_t) i;
    if (!(h.x < B2_INF)) {
        for (int k = 0; k < 12; ++k) o[k] = 0.f;
        float4 rd = ray_d[i];
        o[12] = -rd.x; o[13] = -rd.y; o[14] = -rd.z; o[15] = __uint_as_float(0xffffffffu);
        return;
    }
    uint32_t slot = sc.prim_to_slot[__float_as_uint(h.w)];
    SI si; si_fill<true>(make_tables(sc), slot, h.x, h.y, h.z, v3(ray_o[i]), v3(ray_d[i]), inst ? inst[i] : -1, si);
    o[0] = si.p.x; o[1] = si.p.y; o[2] = si.p.z; o[3] = si.n.x; o[4] = si.n.y; o[5] = si.n.z;
    o[6] = si.sh_n.x; o[7] = si.sh_n.y; o[8] = si.sh_n.z; o[9] = si.sh_s.x; o[10] = si.sh_s.y; o[11] = si.sh_s.z;
    o[12] = si.wi.x; o[13] = si.wi.y; o[14] = si.wi.z; o[15] = __uint_as_float((uint32_t) si.bsdf);
}
