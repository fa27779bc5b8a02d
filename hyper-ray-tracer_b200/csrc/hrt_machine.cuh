// hrt_machine.cuh — warp-level op-class scheduler over the op stream (render_kernel, trace_hits_sched_kernel), and the
// step_* class bodies every render kernel shares.
//
// Why: a straightforward per-lane interpreter (`traverse<>` in hrt_device.cuh) lets the 32 lanes of a warp sit at
// different record kinds, so every loop iteration pays the SUM of the class bodies present: box (~40 instructions) +
// sphere + cuboid (~100) + medium (~300) ..., each for a handful of lanes.  ncu on the first version: 5.4 of 32 lanes
// active per issued instruction (profiles/r01_render_kernel_summary.md).
//
// Here every lane keeps its traversal state in registers and PARKS at its current record; each round the warp votes and
// runs ONE class with a warp-uniform branch:
//   * box records (~85 % of all steps) whenever at least kBoxQuorum lanes are at a box (one ballot), or when the box
//     population is larger than every parked class;
//   * otherwise the non-box class with the most parked lanes: sphere, rect / cuboid, misc (ray-space change, medium),
//     done ("traversal finished -> shade") or new ("no path -> draw a camera sample").
// A lane therefore never waits for the longest traversal or the longest path of its warp, only for its class's turn.
// Traversal order per ray is unchanged (the reference's fixed depth-first order), so results are bit-identical to
// `traverse<>` (tests/test_gpu_parity.py::test_hit_records_through_warp_scheduler).
#pragma once
#include "hrt_device.cuh"

namespace HRT_NS {

enum LaneClass : int { CLS_BOX = 0, CLS_SPHERE = 1, CLS_RECT = 2, CLS_MISC = 3, CLS_DONE = 4, CLS_NEW = 5, CLS_IDLE = 6 };

struct Lane {
    float4 A, B;  // the record at pc, fetched as soon as pc is known (overlaps the vote and other classes' work)
    int pc;
    Ray cur;      // ray in the current context
    RayK k;
    int ctx;
    float closest;
    int best_pc, best_face, best_ctx;
};

__device__ __forceinline__ int record_class(uint32_t opc) {
    return (int)(opc >> 4) - 1;  // hrt_types.h: the opcode's high nibble is the class + 1
}
__device__ __forceinline__ void lane_fetch(const DeviceScene& S, Lane& L) { load_op(S, L.pc, L.A, L.B); }
__device__ __forceinline__ int lane_class(const Lane& L) { return record_class(__float_as_uint(L.B.w) & 0xffu); }

__device__ __forceinline__ void lane_start(const DeviceScene& S, Lane& L, const Ray& world, float tmax) {
    L.cur = world;
    L.k = make_rayk(world);
    L.ctx = 0;
    L.pc = 0;
    L.closest = tmax;
    L.best_pc = -1;
    L.best_face = 0;
    L.best_ctx = 0;
    lane_fetch(S, L);
}

__device__ __forceinline__ void lane_accept(Lane& L, float t, int face) {
    L.closest = t;
    L.best_pc = L.pc;
    L.best_face = face;
    L.best_ctx = L.ctx;
}

// ---- class bodies: each advances the lane by exactly one record and prefetches the next ----
__device__ __forceinline__ void step_box(const DeviceScene& S, Lane& L, float tmin, bool reference_boxes) {
    const uint32_t w7 = __float_as_uint(L.B.w);
    const bool loose = ((w7 & 0xffu) == OP_BOX_LOOSE) || reference_boxes;
    const bool hit = loose ? box_hit_reference(L.A, L.B, L.cur, L.k, tmin, L.closest)
                           : box_hit_tight(L.A, L.B, L.cur, L.k, tmin, L.closest);
    L.pc = hit ? L.pc + 1 : (int)(w7 >> 8);
    lane_fetch(S, L);
}
__device__ __forceinline__ void step_sphere(const DeviceScene& S, Lane& L, float tmin) {
    const bool moving = (__float_as_uint(L.B.w) & 0xffu) == OP_MSPHERE;
    V3 ctr = v3(L.A.x, L.A.y, L.A.z);
    if (moving) {
        float4 C, D;
        load_op(S, L.pc + 1, C, D);
        ctr = msphere_center(ctr, v3(C.x, C.y, C.z), C.w, D.x, L.cur.time);
    }
    float t;
    if (sphere_test(ctr, L.A.w, L.cur, L.k, tmin, L.closest, t)) lane_accept(L, t, 0);
    L.pc += moving ? 2 : 1;
    lane_fetch(S, L);
}
__device__ __forceinline__ void step_rect(const DeviceScene& S, Lane& L, float tmin) {
    const uint32_t opc = __float_as_uint(L.B.w) & 0xffu;
    const float4 A = L.A, B = L.B;
    const Ray& c = L.cur;
    float t;
    if (opc == OP_CUBOID) {
        int face = 0;
        if (cuboid_test(v3(A.x, A.y, A.z), v3(B.x, B.y, B.z), c, L.k, tmin, L.closest, t, face)) lane_accept(L, t, face);
    } else {
        bool h;
        if (opc == OP_RECT_XY) h = rect_test(c.o.z, c.d.z, L.k.inv.z, c.o.x, c.d.x, c.o.y, c.d.y, A.x, A.y, A.z, A.w, B.x, tmin, L.closest, t);
        else if (opc == OP_RECT_YZ) h = rect_test(c.o.x, c.d.x, L.k.inv.x, c.o.y, c.d.y, c.o.z, c.d.z, A.x, A.y, A.z, A.w, B.x, tmin, L.closest, t);
        else h = rect_test(c.o.y, c.d.y, L.k.inv.y, c.o.z, c.d.z, c.o.x, c.d.x, A.x, A.y, A.z, A.w, B.x, tmin, L.closest, t);
        if (h) lane_accept(L, t, 0);
    }
    L.pc += 1;
    lane_fetch(S, L);
}

// The tail of ConstantMedium::hit once both boundary hits are known (constant_medium.rs:40-75).
__device__ __forceinline__ void medium_finish(const DeviceScene& S, Lane& L, float t1, float t2, float tmin, const MediumXi& xi) {
    if (t1 < tmin) t1 = tmin;
    if (t2 > L.closest) t2 = L.closest;
    if (!(t1 >= t2)) {
        if (t1 < 0.0f) t1 = 0.0f;
        const float ray_length = sqrtf(L.k.dd);
        const float dist_inside = (t2 - t1) * ray_length;
        const float u = xi.draw(__float_as_int(L.A.z));
#if HRT_EXACT
        const float hit_distance = L.A.x * (logf(u) / S.ln_e);
#else
        const float hit_distance = L.A.x * logf(u);
#endif
        if (!(hit_distance > dist_inside)) lane_accept(L, t1 + hit_distance / ray_length, 0);
    }
}

// TRANSLATE / ROTATE / POP / MEDIUM / MEDIUM_SPHERE / BVH
__device__ __forceinline__ void step_misc(const DeviceScene& S, Lane& L, const Ray& world, float tmin, bool reference_boxes,
                                          const MediumXi& xi) {
    const uint32_t w7 = __float_as_uint(L.B.w);
    const uint32_t opc = w7 & 0xffu;
    if (opc < OP_MEDIUM) {
        // TRANSLATE / ROTATE / POP: enter or leave a ray space.  All three map the WORLD ray through the target
        // context's push records (the same operations in the same order as applying them incrementally, so the result is
        // bit-identical) — one out-of-line code path instead of three inlined ones.
        L.ctx = __float_as_int(L.A.w);
        L.cur = ray_in_ctx(S, world, L.ctx);
        L.k = make_rayk(L.cur);
        const int run = (int)(w7 >> 8);
        L.pc += run > 0 ? run : 1;
    } else if (opc == OP_BVH) {  // a sound BvhNode as a two-child tree: per-ray stack walk (hrt_device.cuh bvh2_walk)
        const TreeHit th = bvh2_walk(S, __float_as_int(L.A.x), L.cur, tmin, L.closest, L.best_pc, L.B.x, L.B.y);
        if (th.pc != L.best_pc) {
            L.closest = th.t; L.best_pc = th.pc; L.best_face = th.face; L.best_ctx = L.ctx;
        }
        L.pc = (int)(w7 >> 8);
    } else {  // constant_medium.rs:34-76
        const int end = (int)(w7 >> 8);
        bool done = false;
        if (opc == OP_MEDIUM_SPHERE) {
            // Boundary = one plain sphere: both boundary queries in closed form, with sphere_test's own arithmetic.
            //   query 1, range (-inf, +inf): the near root is always accepted;
            //   query 2, range (t1 + 1e-4, +inf): the near root (= t1) is below the range, so it is the far root or nothing.
            float4 C, D;
            load_op(S, L.pc + 1, C, D);
            const Ray& r = L.cur;
            const V3 oc = v3(__fsub_rn(r.o.x, C.x), __fsub_rn(r.o.y, C.y), __fsub_rn(r.o.z, C.z));
            const float a = L.k.dd;
            const float half_b = dot_rn(oc, r.d);
            const float c = __fsub_rn(dot_rn(oc, oc), __fmul_rn(C.w, C.w));
            const float disc = __fsub_rn(__fmul_rn(half_b, half_b), __fmul_rn(a, c));
            if (disc < 0.0f) {
                done = true;  // boundary missed -> None
            } else {
                const float sqrtd = sqrtf(disc);
                const float t1 = __fdiv_rn(-half_b - sqrtd, a);
                const float t2 = __fdiv_rn(-half_b + sqrtd, a);
                const float lo = t1 + 0.0001f;
                if (t1 == t1 && t2 == t2) {  // NaN roots take the generic path
                    done = true;
                    // sphere.rs:52-57 with range (lo, +inf): near root unless it is below lo, then the far root.  For
                    // |t1| >= 2048 the f32 sum t1 + 1e-4 equals t1, the near root is accepted AGAIN, t2 == t1 and the
                    // medium never scatters — reference behaviour, reproduced.
                    if (!(t1 < lo)) medium_finish(S, L, t1, t1, tmin, xi);
                    else if (!(t2 < lo)) medium_finish(S, L, t1, t2, tmin, xi);
                }
            }
        }
        if (!done) {
            const float t1 = boundary_hit(S, L.pc + 1, end, world, L.cur, L.ctx, -CUDART_INF_F, reference_boxes);
            if (t1 == t1) {
                const float t2 = boundary_hit(S, L.pc + 1, end, world, L.cur, L.ctx, t1 + 0.0001f, reference_boxes);
                if (t2 == t2) medium_finish(S, L, t1, t2, tmin, xi);
            }
        }
        L.pc = end;
    }
    lane_fetch(S, L);
}

// ---- scheduling policy ----
// Measured (profiles/r01_render_kernel_summary.md): the box test itself is ~3 % of the kernel's stall samples; the time
// is in the leaf / medium / shade bodies, which are long and were running with 3-5 lanes under a "service every parked
// class at once" policy.  So: boxes run whenever a small quorum is at a box (they are cheap and feed the other classes),
// otherwise the single non-box class with the MOST parked lanes runs.  Alternatives measured and dropped: a MATCH.ANY +
// REDUX.MAX vote (-9 %), prefetching both successors of a box (+-0), batching thresholds for the parked classes (+-0).
#ifndef HRT_BOX_QUORUM
#define HRT_BOX_QUORUM 8
#endif
constexpr int kBoxQuorum = HRT_BOX_QUORUM;

struct Tier {
    bool box, leaf, done, fill, any;  // what to run this round (warp-uniform; at most one of box/leaf/done/fill)
    int leaf_cls;                     // which leaf class when `leaf`
};
// Number of lanes at a box record (the fast path of the vote: one ballot).
__device__ __forceinline__ int warp_box_count(int cls) { return __popc(__ballot_sync(0xffffffffu, cls == CLS_BOX)); }
// Slow path, taken when fewer than kBoxQuorum lanes are at a box: populations of the five non-box classes in ONE
// REDUX.ADD over 6-bit packed counters (sphere, rect, misc, done, new), then the class with the most parked lanes.
__device__ __forceinline__ Tier warp_plan_slow(int cls, int nb) {
    const unsigned full = 0xffffffffu;
    const unsigned one = (cls >= CLS_SPHERE && cls <= CLS_NEW) ? (1u << (6 * (cls - 1))) : 0u;
    const unsigned packed = __reduce_add_sync(full, one);
    int best = -1, best_n = 0;
#pragma unroll
    for (int c = CLS_SPHERE; c <= CLS_NEW; ++c) {
        const int n = (int)((packed >> (6 * (c - 1))) & 63u);
        if (n > best_n) { best = c; best_n = n; }
    }
    Tier t;
    t.leaf = t.done = t.fill = false;
    t.leaf_cls = CLS_SPHERE;
    t.any = (nb > 0) || (best >= 0);
    t.box = nb > 0 && nb > best_n;  // a box population larger than every parked class still goes first
    if (t.box || best < 0) return t;
    t.leaf = best <= CLS_MISC;
    t.leaf_cls = best;
    t.done = best == CLS_DONE;
    t.fill = best == CLS_NEW;
    return t;
}
__device__ __forceinline__ Tier warp_plan(int cls) {
    const int nb = warp_box_count(cls);
    if (nb >= kBoxQuorum) {
        Tier t;
        t.box = t.any = true;
        t.leaf = t.done = t.fill = false;
        t.leaf_cls = CLS_SPHERE;
        return t;
    }
    return warp_plan_slow(cls, nb);
}

}  // namespace HRT_NS
