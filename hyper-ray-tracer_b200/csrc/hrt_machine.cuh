// hrt_machine.cuh — warp-level op-class scheduler over the op stream.
//
// Why: a straightforward per-lane interpreter (`traverse<>` in hrt_device.cuh) lets the 32 lanes of a warp sit at
// different record kinds, so every iteration executes the box code, the sphere code, the cuboid code, the medium
// code ... one after the other with a handful of lanes each.  ncu on the first version: 5.4 of 32 lanes active per
// issued instruction (profiles/r01_v1_render_kernel.txt).
//
// Here every lane keeps its traversal state in registers and PARKS at its current record.  Each round the warp
// votes (ballot + popc) and executes ONE record class — the one most lanes are waiting for — with a warp-uniform
// branch.  "Traversal finished -> shade" and "no path -> draw a new camera sample" are classes as well, so a lane
// never waits for the longest traversal or the longest path in its warp, only for its class to win a vote.
// Traversal order per ray is unchanged (the reference's fixed depth-first order), so results are identical to
// `traverse<>`.
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
    // OP_END 0 | BOX 1,2 | SPHERE 3,4 | (AUX 5) | RECT 6,7,8, CUBOID 9 | TRANSLATE 10, ROTATE 11, POP 12, MEDIUM 13
    return opc == 0u ? CLS_DONE : (opc <= 2u ? CLS_BOX : (opc <= 5u ? CLS_SPHERE : (opc <= 9u ? CLS_RECT : CLS_MISC)));
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
// TRANSLATE / ROTATE / POP / MEDIUM
__device__ __forceinline__ void step_misc(const DeviceScene& S, Lane& L, const Ray& world, float tmin, bool reference_boxes,
                                          const MediumXi& xi) {
    const uint32_t w7 = __float_as_uint(L.B.w);
    const uint32_t opc = w7 & 0xffu;
    if (opc != OP_MEDIUM) {
        // TRANSLATE / ROTATE / POP: enter or leave a ray space.  All three map the WORLD ray through the target
        // context's push records (the same operations in the same order as applying them incrementally, so the result is
        // bit-identical) — one out-of-line code path instead of three inlined ones.
        L.ctx = __float_as_int(L.A.w);
        L.cur = ray_in_ctx(S, world, L.ctx);
        L.k = make_rayk(L.cur);
        L.pc += 1;
    } else {  // OP_MEDIUM — constant_medium.rs:34-76
        const int end = (int)(w7 >> 8);
        float t1 = boundary_hit(S, L.pc + 1, end, world, L.cur, L.ctx, -CUDART_INF_F, reference_boxes);
        if (t1 == t1) {
            float t2 = boundary_hit(S, L.pc + 1, end, world, L.cur, L.ctx, t1 + 0.0001f, reference_boxes);
            if (t2 == t2) {
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
        }
        L.pc = end;
    }
    lane_fetch(S, L);
}

// Majority vote.  Boxes are ~85 % of all records, so their ballot is taken first and wins outright above a quorum;
// otherwise the class with the most parked lanes runs.  `cls` may be CLS_IDLE (never chosen).  Returns CLS_IDLE when
// no lane has anything to do.
constexpr int kBoxQuorum = 12;
__device__ __forceinline__ int warp_vote(int cls) {
    const unsigned full = 0xffffffffu;
    const int nb = __popc(__ballot_sync(full, cls == CLS_BOX));
    if (nb >= kBoxQuorum) return CLS_BOX;
    int best = CLS_IDLE, best_n = 0;
    if (nb > 0) { best = CLS_BOX; best_n = nb; }
    const int ns = __popc(__ballot_sync(full, cls == CLS_SPHERE));
    if (ns > best_n) { best = CLS_SPHERE; best_n = ns; }
    const int nr = __popc(__ballot_sync(full, cls == CLS_RECT));
    if (nr > best_n) { best = CLS_RECT; best_n = nr; }
    const int nm = __popc(__ballot_sync(full, cls == CLS_MISC));
    if (nm > best_n) { best = CLS_MISC; best_n = nm; }
    const int nd = __popc(__ballot_sync(full, cls == CLS_DONE));
    if (nd > best_n) { best = CLS_DONE; best_n = nd; }
    const int nn = __popc(__ballot_sync(full, cls == CLS_NEW));
    if (nn > best_n) { best = CLS_NEW; best_n = nn; }
    return best;
}

}  // namespace HRT_NS
