// hrt_scene.cpp — C-ABI scene builder and the flattener ("host-side Rust flattener" of the north star,
// delivered in C++ behind the C ABI because there is no Rust toolchain in this image).
//
// Builder calls mirror the reference constructors 1:1 (citations in include/hrt.h).  hrt_scene_commit()
// reproduces, in f32 and in the reference's operation order, everything the reference computes at
// construction time — every `bounding_box` (sphere.rs:77-83, moving_sphere.rs:98-110, rect.rs:88-103,
// cuboid.rs:104-106, list.rs:33-44, translation.rs:39-48, rotation.rs:43-89, constant_medium.rs:78-80),
// `Aabb::surrounding_box` (aabb.rs:49-63) and `BvhNode::new` (bvh_node.rs:27-100) — and then lowers the
// tree to the stackless op stream described in hrt_types.h.
//
// Build with -ffp-contract=off (no FMA contraction): box coordinates must equal the reference's bit for bit.
#include "hrt_scene.hpp"

#include <algorithm>
#include <array>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <limits>

#include "../../include/hrt.h"

namespace hrt {

static thread_local std::string g_last_error;
void set_error(const std::string& msg) { g_last_error = msg; }
int32_t fail(int32_t code, const std::string& msg) {
    g_last_error = msg;
    return code;
}

static const float PI_F = 3.14159265358979323846f;
static const float FMAX = std::numeric_limits<float>::max();

static Box3 surrounding_box(const Box3& a, const Box3& b) {  // aabb.rs:49-63
    Box3 o;
    for (int i = 0; i < 3; ++i) {
        o.mn[i] = std::fmin(a.mn[i], b.mn[i]);
        o.mx[i] = std::fmax(a.mx[i], b.mx[i]);
    }
    return o;
}
static bool contains(const Box3& outer, const Box3& inner) {
    for (int i = 0; i < 3; ++i)
        if (!(outer.mn[i] <= inner.mn[i] && outer.mx[i] >= inner.mx[i])) return false;
    return true;
}

static void rotation_axes(int axis, int& r, int& a, int& b) {  // rotation.rs:20-26
    switch (axis) {
        case 0: r = 0; a = 1; b = 2; break;
        case 1: r = 1; a = 2; b = 0; break;
        default: r = 2; a = 0; b = 1; break;
    }
}

// rotation.rs:43-89 applied to an arbitrary input box
static Box3 rotate_box(const Box3& bb, int axis, float sin_theta, float cos_theta) {
    int r_axis, a_axis, b_axis;
    rotation_axes(axis, r_axis, a_axis, b_axis);
    Box3 o;
    for (int i = 0; i < 3; ++i) { o.mn[i] = FMAX; o.mx[i] = -FMAX; }
    for (int i = 0; i < 2; ++i)
        for (int j = 0; j < 2; ++j)
            for (int k = 0; k < 2; ++k) {
                float r = (float)k * bb.mx[r_axis] + (float)(1 - k) * bb.mn[r_axis];
                float a = (float)i * bb.mx[a_axis] + (float)(1 - i) * bb.mn[a_axis];
                float b = (float)j * bb.mx[b_axis] + (float)(1 - j) * bb.mn[b_axis];
                float new_a = cos_theta * a - sin_theta * b;
                float new_b = sin_theta * a + cos_theta * b;
                if (new_a < o.mn[a_axis]) o.mn[a_axis] = new_a;
                if (new_b < o.mn[b_axis]) o.mn[b_axis] = new_b;
                if (r < o.mn[r_axis]) o.mn[r_axis] = r;
                if (new_a > o.mx[a_axis]) o.mx[a_axis] = new_a;
                if (new_b > o.mx[b_axis]) o.mx[b_axis] = new_b;
                if (r > o.mx[r_axis]) o.mx[r_axis] = r;
            }
    return o;
}

static void msphere_center(const Obj& o, float time, float out[3]) {  // moving_sphere.rs:53-57
    float f = (time - o.t0) / (o.t1 - o.t0);
    for (int i = 0; i < 3; ++i) out[i] = o.c0[i] + f * (o.c1[i] - o.c0[i]);
}

// `bounding_box(time_start, time_end)` of the reference, bit for bit.  `truth` = true returns instead the
// box that actually contains the geometry hit() accepts (differs for ZX rects, Q2).
static bool ref_box(const hrt_scene& s, int id, float ts, float te, bool truth, Box3& out) {
    const Obj& o = s.objects[id];
    switch (o.kind) {
        case OBJ_SPHERE:
            for (int i = 0; i < 3; ++i) { out.mn[i] = o.c0[i] - o.r; out.mx[i] = o.c0[i] + o.r; }
            return true;
        case OBJ_MSPHERE: {
            float ca[3], cb[3];
            msphere_center(o, ts, ca);
            msphere_center(o, te, cb);
            Box3 b0, b1;
            for (int i = 0; i < 3; ++i) {
                b0.mn[i] = ca[i] - o.r; b0.mx[i] = ca[i] + o.r;
                b1.mn[i] = cb[i] - o.r; b1.mx[i] = cb[i] + o.r;
            }
            out = surrounding_box(b0, b1);
            return true;
        }
        case OBJ_RECT: {
            const float lo = o.k - 0.0001f, hi = o.k + 0.0001f;
            switch (o.plane_or_axis) {
                case HRT_PLANE_XY: out = Box3{{o.a0, o.b0, lo}, {o.a1, o.b1, hi}}; break;
                case HRT_PLANE_YZ: out = Box3{{lo, o.a0, o.b0}, {hi, o.a1, o.b1}}; break;
                default:
                    if (truth) out = Box3{{o.b0, lo, o.a0}, {o.b1, hi, o.a1}};  // hit(): a is z, b is x (rect.rs:57)
                    else       out = Box3{{o.a0, lo, o.b0}, {o.a1, hi, o.b1}};  // rect.rs:98-101 (axis-swapped)
                    break;
            }
            return true;
        }
        case OBJ_CUBOID:
            for (int i = 0; i < 3; ++i) { out.mn[i] = o.c0[i]; out.mx[i] = o.c1[i]; }
            return true;
        case OBJ_TRANSLATE: {
            Box3 b;
            if (!ref_box(s, o.child, ts, te, truth, b)) return false;
            for (int i = 0; i < 3; ++i) { out.mn[i] = b.mn[i] + o.c0[i]; out.mx[i] = b.mx[i] + o.c0[i]; }
            return true;
        }
        case OBJ_ROTATE: {
            if (!truth) {
                if (!o.has_rot_box) return false;
                out = o.rot_box;
                return true;
            }
            Box3 b;
            if (!ref_box(s, o.child, 0.0f, 1.0f, true, b)) return false;
            out = rotate_box(b, o.plane_or_axis, o.sin_theta, o.cos_theta);
            return true;
        }
        case OBJ_MEDIUM:
            return ref_box(s, o.child, ts, te, truth, out);
        case OBJ_LIST: {
            if (o.children.empty()) return false;
            Box3 acc;
            if (!ref_box(s, o.children[0], ts, te, truth, acc)) return false;
            for (size_t i = 1; i < o.children.size(); ++i) {
                Box3 b;
                if (!ref_box(s, o.children[i], ts, te, truth, b)) return false;
                acc = surrounding_box(acc, b);
            }
            out = acc;
            return true;
        }
        case OBJ_BVH: {
            if (!truth) {
                out = o.bvh.nodes[o.bvh.root].box;
                return true;
            }
            Box3 acc;
            bool first = true;
            for (const BvhTreeNode& n : o.bvh.nodes) {
                if (n.leaf_obj < 0) continue;
                Box3 b;
                if (!ref_box(s, n.leaf_obj, o.t0, o.t1, true, b)) return false;
                acc = first ? b : surrounding_box(acc, b);
                first = false;
            }
            out = acc;
            return !first;
        }
    }
    return false;
}

// BvhNode::new (bvh_node.rs:27-63).  sort_unstable_by is an insertion sort for n <= 20 in rustc's
// implementation (stable order); for larger n the order of EQUAL keys is toolchain-defined.  A stable sort is used
// throughout, so the tree is the reference's up to the order of equal keys; ties cannot change hit results on sound
// boxes.
static int32_t build_bvh(const hrt_scene& s, BvhTree& tree, std::vector<int32_t> objs, float ts, float te) {
    std::pair<int, float> ranges[3];
    for (int axis = 0; axis < 3; ++axis) {  // axis_range, bvh_node.rs:83-100
        float mn = FMAX, mx = -FMAX;
        for (int32_t id : objs) {
            Box3 b;
            if (!ref_box(s, id, ts, te, false, b)) continue;
            mn = std::fmin(mn, b.mn[axis]);
            mx = std::fmax(mx, b.mx[axis]);
        }
        ranges[axis] = {axis, mx - mn};
    }
    std::stable_sort(ranges, ranges + 3,
                     [](const std::pair<int, float>& a, const std::pair<int, float>& b) { return a.second > b.second; });
    const int axis = ranges[0].first;
    std::vector<std::pair<float, int32_t>> keyed;
    keyed.reserve(objs.size());
    for (int32_t id : objs) {  // box_compare key, bvh_node.rs:70-76
        Box3 b;
        ref_box(s, id, ts, te, false, b);
        keyed.push_back({b.mn[axis] + b.mx[axis], id});
    }
    std::stable_sort(keyed.begin(), keyed.end(),
                     [](const std::pair<float, int32_t>& a, const std::pair<float, int32_t>& b) { return a.first < b.first; });
    const size_t len = keyed.size();
    BvhTreeNode node;
    if (len == 1) {
        node.leaf_obj = keyed[0].second;
        ref_box(s, node.leaf_obj, ts, te, false, node.box);
        tree.nodes.push_back(node);
        return (int32_t)tree.nodes.size() - 1;
    }
    std::vector<int32_t> l, r;
    for (size_t i = 0; i < len / 2; ++i) l.push_back(keyed[i].second);
    for (size_t i = len / 2; i < len; ++i) r.push_back(keyed[i].second);
    int32_t ri = build_bvh(s, tree, std::move(r), ts, te);
    int32_t li = build_bvh(s, tree, std::move(l), ts, te);
    node.left = li;
    node.right = ri;
    node.box = surrounding_box(tree.nodes[li].box, tree.nodes[ri].box);
    tree.nodes.push_back(node);
    return (int32_t)tree.nodes.size() - 1;
}

// Topology of the OP_BVH trees (hrt_types.h): a full-sweep surface-area heuristic over the three centroid orders down to
// kBvh2SahDepth, balanced object-median splits below that (bounds the tree depth, hence the per-ray stack).  The node
// boxes are unions of the leaves' reference boxes; only the topology differs from BvhNode::new's.
static float half_area(const Box3& b) {
    const float dx = b.mx[0] - b.mn[0], dy = b.mx[1] - b.mn[1], dz = b.mx[2] - b.mn[2];
    return dx * dy + dy * dz + dz * dx;
}
static int32_t build_bvh_sah(const hrt_scene& s, BvhTree& tree, std::vector<int32_t> objs, float ts, float te, int depth) {
    const size_t n = objs.size();
    BvhTreeNode node;
    if (n == 1) {
        node.leaf_obj = objs[0];
        ref_box(s, node.leaf_obj, ts, te, false, node.box);
        tree.nodes.push_back(node);
        return (int32_t)tree.nodes.size() - 1;
    }
    struct Item { float key; int32_t id; Box3 box; };
    std::vector<Item> items(n);
    std::vector<float> right_area(n + 1);
    float best_cost = std::numeric_limits<float>::infinity();
    int best_axis = 0;
    size_t best_split = n / 2;
    auto sort_axis = [&](int axis) {
        for (size_t i = 0; i < n; ++i) {
            items[i].id = objs[i];
            ref_box(s, objs[i], ts, te, false, items[i].box);
            items[i].key = items[i].box.mn[axis] + items[i].box.mx[axis];
        }
        std::stable_sort(items.begin(), items.end(), [](const Item& a, const Item& b) { return a.key < b.key; });
    };
    if (depth < kBvh2SahDepth) {
        for (int axis = 0; axis < 3; ++axis) {
            sort_axis(axis);
            Box3 acc = items[n - 1].box;
            right_area[n - 1] = half_area(acc);
            for (size_t i = n - 1; i-- > 0;) { acc = surrounding_box(acc, items[i].box); right_area[i] = half_area(acc); }
            acc = items[0].box;
            for (size_t i = 1; i < n; ++i) {  // left = [0, i), right = [i, n)
                const float cost = half_area(acc) * (float)i + right_area[i] * (float)(n - i);
                if (cost < best_cost) { best_cost = cost; best_axis = axis; best_split = i; }
                acc = surrounding_box(acc, items[i].box);
            }
        }
    } else {  // longest axis of the centroid range, object median
        float ext[3];
        for (int axis = 0; axis < 3; ++axis) {
            float mn = FMAX, mx = -FMAX;
            for (int32_t id : objs) {
                Box3 b;
                ref_box(s, id, ts, te, false, b);
                mn = std::fmin(mn, b.mn[axis] + b.mx[axis]);
                mx = std::fmax(mx, b.mn[axis] + b.mx[axis]);
            }
            ext[axis] = mx - mn;
        }
        best_axis = ext[0] >= ext[1] ? (ext[0] >= ext[2] ? 0 : 2) : (ext[1] >= ext[2] ? 1 : 2);
    }
    sort_axis(best_axis);
    std::vector<int32_t> l, r;
    for (size_t i = 0; i < best_split; ++i) l.push_back(items[i].id);
    for (size_t i = best_split; i < n; ++i) r.push_back(items[i].id);
    items.clear();
    const int32_t ri = build_bvh_sah(s, tree, std::move(r), ts, te, depth + 1);
    const int32_t li = build_bvh_sah(s, tree, std::move(l), ts, te, depth + 1);
    node.left = li;
    node.right = ri;
    node.box = surrounding_box(tree.nodes[li].box, tree.nodes[ri].box);
    tree.nodes.push_back(node);
    return (int32_t)tree.nodes.size() - 1;
}

static uint32_t count_of(const hrt_scene& s, int id) {
    const Obj& o = s.objects[id];
    switch (o.kind) {
        case OBJ_SPHERE: case OBJ_MSPHERE: case OBJ_RECT: return 1;
        case OBJ_CUBOID: return 6;                     // cuboid.rs:108-110 -> List::count over 6 rects
        case OBJ_TRANSLATE: return count_of(s, o.child);
        case OBJ_ROTATE: return 1;                     // rotation.rs:140-142
        case OBJ_MEDIUM: return count_of(s, o.child);
        case OBJ_LIST: { uint32_t n = 0; for (int c : o.children) n += count_of(s, c); return n; }
        case OBJ_BVH: { uint32_t n = 0; for (auto& nd : o.bvh.nodes) if (nd.leaf_obj >= 0) n += count_of(s, nd.leaf_obj); return n; }
    }
    return 0;
}

// ------------------------------------------------------------------------------------------------
// Flattener
// ------------------------------------------------------------------------------------------------
static uint16_t float_to_half_directed(float x, bool up);

struct Flattener {
    const hrt_scene& s;
    FlatScene& f;
    const bool trees;  // sound BVHs of plain primitives become OP_BVH trees
    // inner boxes over at most this many leaves are left out of the fast form (HRT_ELIDE_BOXES=0: diagnostic, keep all)
    static constexpr int32_t kElideInnerMaxLeaves = 16;
    const bool elide_boxes = [] { const char* e = std::getenv("HRT_ELIDE_BOXES"); return !(e && e[0] == '0'); }();
    std::string error;
    bool in_medium = false;
    Flattener(const hrt_scene& sc, FlatScene& out, bool bvh_trees) : s(sc), f(out), trees(bvh_trees) {}

    int32_t pc() const { return (int32_t)f.ops.size(); }
    Op& push(uint32_t opcode) {
        Op op;
        std::memset(&op, 0, sizeof(op));
        op.u[7] = opcode;
        f.ops.push_back(op);
        return f.ops.back();
    }
    int32_t new_ctx(int32_t parent, int32_t op_pc) {
        Ctx c = f.ctxs[parent];
        if (c.depth >= kMaxCtxDepth) { error = "Translation/Rotation nesting deeper than 6"; return -1; }
        c.op_pc[c.depth] = op_pc;
        c.depth += 1;
        c.parent = parent;
        f.ctxs.push_back(c);
        f.max_ctx_depth = std::max(f.max_ctx_depth, c.depth);
        return (int32_t)f.ctxs.size() - 1;
    }

    bool emit(int32_t id, int32_t ctx) {
        const Obj& o = s.objects[id];
        switch (o.kind) {
            case OBJ_SPHERE: {
                Op& op = push(OP_SPHERE);
                op.f[0] = o.c0[0]; op.f[1] = o.c0[1]; op.f[2] = o.c0[2]; op.f[3] = o.r;
                op.i[4] = o.mat; op.i[5] = id;
                f.n_prim_ops++;
                return true;
            }
            case OBJ_MSPHERE: {
                Op& op = push(OP_MSPHERE);
                op.f[0] = o.c0[0]; op.f[1] = o.c0[1]; op.f[2] = o.c0[2]; op.f[3] = o.r;
                op.i[4] = o.mat; op.i[5] = id;
                Op& aux = push(OP_MSPHERE_AUX);
                aux.f[0] = o.c1[0]; aux.f[1] = o.c1[1]; aux.f[2] = o.c1[2]; aux.f[3] = o.t0; aux.f[4] = o.t1;
                f.n_prim_ops++;
                return true;
            }
            case OBJ_RECT: {
                uint32_t opc = o.plane_or_axis == HRT_PLANE_XY ? OP_RECT_XY : (o.plane_or_axis == HRT_PLANE_YZ ? OP_RECT_YZ : OP_RECT_ZX);
                Op& op = push(opc);
                op.f[0] = o.a0; op.f[1] = o.a1; op.f[2] = o.b0; op.f[3] = o.b1; op.f[4] = o.k;
                op.i[5] = o.mat; op.i[6] = id;
                f.n_prim_ops++;
                return true;
            }
            case OBJ_CUBOID: {
                if (id >= (1 << 24)) { error = "more than 2^24 objects"; return false; }
                Op& op = push(OP_CUBOID);
                op.f[0] = o.c0[0]; op.f[1] = o.c0[1]; op.f[2] = o.c0[2]; op.i[3] = o.mat;
                op.f[4] = o.c1[0]; op.f[5] = o.c1[1]; op.f[6] = o.c1[2];
                op.u[7] = OP_CUBOID | ((uint32_t)id << 8);
                f.n_prim_ops++;
                return true;
            }
            case OBJ_TRANSLATE: {
                int32_t at = pc();
                int32_t nctx = new_ctx(ctx, at);
                if (nctx < 0) return false;
                Op& op = push(OP_TRANSLATE);
                op.f[0] = o.c0[0]; op.f[1] = o.c0[1]; op.f[2] = o.c0[2]; op.i[3] = nctx;
                if (!emit(o.child, nctx)) return false;
                Op& pop = push(OP_POP);
                pop.i[3] = ctx;
                return true;
            }
            case OBJ_ROTATE: {
                int32_t at = pc();
                int32_t nctx = new_ctx(ctx, at);
                if (nctx < 0) return false;
                Op& op = push(OP_ROTATE);
                op.f[0] = o.sin_theta; op.f[1] = o.cos_theta; op.i[2] = o.plane_or_axis; op.i[3] = nctx;
                if (!emit(o.child, nctx)) return false;
                Op& pop = push(OP_POP);
                pop.i[3] = ctx;
                return true;
            }
            case OBJ_MEDIUM: {
                if (in_medium) { error = "ConstantMedium nested inside a ConstantMedium boundary is not supported"; return false; }
                int32_t at = pc();
                {
                    Op& op = push(OP_MEDIUM);
                    op.f[0] = o.neg_inv_density; op.i[1] = o.mat; op.i[2] = f.n_media; op.i[3] = id;
                }
                f.n_media++;
                in_medium = true;
                bool ok = emit(o.child, ctx);
                in_medium = false;
                if (!ok) return false;
                if (pc() >= (1 << 24)) { error = "op stream longer than 2^24 records"; return false; }
                {
                    const bool single_sphere = (pc() == at + 2) && ((f.ops[at + 1].u[7] & 0xffu) == OP_SPHERE);
                    // [push x r] CUBOID [pop x r]: one cuboid in its own ray space (r = 0: bare)
                    bool boxed = false;
                    {
                        auto opc = [&](int32_t i) { return f.ops[i].u[7] & 0xffu; };
                        int32_t r = 0;
                        while (at + 1 + r < pc() && (opc(at + 1 + r) == OP_TRANSLATE || opc(at + 1 + r) == OP_ROTATE)) ++r;
                        if (pc() == at + 2 + 2 * r && opc(at + 1 + r) == OP_CUBOID) {
                            boxed = true;
                            for (int32_t i = 0; i < r; ++i) boxed = boxed && opc(at + 2 + r + i) == OP_POP;
                        }
                    }
                    f.ops[at].u[7] = (single_sphere ? OP_MEDIUM_SPHERE : (boxed ? OP_MEDIUM_CUBOID : OP_MEDIUM)) | ((uint32_t)pc() << 8);
                }
                return true;
            }
            case OBJ_LIST: {
                // List::hit (list.rs:20-31): sequential closest-hit with narrowing, no box tests — exactly
                // what falling through consecutive records does.
                for (int32_t c : o.children)
                    if (!emit(c, ctx)) return false;
                return true;
            }
            case OBJ_BVH: {
                if (trees && tree_eligible(o)) return emit_tree(o, ctx);
                return emit_bvh(o.bvh, o.t0, o.t1, o.bvh.root, ctx);
            }
        }
        return false;
    }

    // OP_BVH trees: every leaf is a plain primitive whose reference box contains it (so every union of leaf boxes is
    // sound in ANY topology, and the closest hit does not depend on the visit order), and there are enough of them for a
    // per-ray stack walk to pay.
    bool tree_eligible(const Obj& bvh) {
        if (bvh.children.size() < 4) return false;
        {
            std::vector<int32_t> ids = bvh.children;
            std::sort(ids.begin(), ids.end());
            if (std::adjacent_find(ids.begin(), ids.end()) != ids.end()) return false;  // the same object twice
        }
        for (const BvhTreeNode& n : bvh.bvh.nodes) {
            if (n.leaf_obj < 0) continue;
            const ObjKind k = s.objects[n.leaf_obj].kind;
            if (k != OBJ_SPHERE && k != OBJ_MSPHERE && k != OBJ_RECT && k != OBJ_CUBOID) return false;
            Box3 truth;
            if (!ref_box(s, n.leaf_obj, bvh.t0, bvh.t1, true, truth) || !contains(n.box, truth)) return false;
            for (int a = 0; a < 3; ++a)
                if (std::isnan(n.box.mn[a]) || std::isnan(n.box.mx[a])) return false;
        }
        return true;
    }

    static void box_to_half(const Box3& b, uint16_t h[6]) {
        for (int a = 0; a < 3; ++a) {
            h[a] = float_to_half_directed(b.mn[a], false);
            h[3 + a] = float_to_half_directed(b.mx[a], true);
        }
    }
    // Lay the inner nodes of `t` out depth-first from `node` (an inner node) and return its index relative to `base`.
    int32_t emit_tree_nodes(const BvhTree& t, int32_t node, const std::vector<int32_t>& leaf_pc, int32_t base, int depth, int& max_depth) {
        const BvhTreeNode& n = t.nodes[node];
        const int32_t at = (int32_t)f.nodes.size();
        f.nodes.push_back(Bvh2Node{});
        max_depth = std::max(max_depth, depth);
        const BvhTreeNode& l = t.nodes[n.left];
        const BvhTreeNode& r = t.nodes[n.right];
        const int32_t li = l.leaf_obj >= 0 ? ~leaf_pc[l.leaf_obj] : emit_tree_nodes(t, n.left, leaf_pc, base, depth + 1, max_depth);
        const int32_t ri = r.leaf_obj >= 0 ? ~leaf_pc[r.leaf_obj] : emit_tree_nodes(t, n.right, leaf_pc, base, depth + 1, max_depth);
        Bvh2Node& out = f.nodes[at];
        box_to_half(l.box, out.lbox);
        box_to_half(r.box, out.rbox);
        out.left = li;
        out.right = ri;
        return at - base;
    }
    // [BOX root, skip = end] [BVH] [leaf primitive records in the REFERENCE's depth-first order] end:
    bool emit_tree(const Obj& o, int32_t ctx) {
        const BvhTreeNode& root = o.bvh.nodes[o.bvh.root];
        const int32_t box_at = pc();
        {
            Op& op = push(OP_BOX);
            op.f[0] = root.box.mn[0]; op.f[1] = root.box.mn[1]; op.f[2] = root.box.mn[2];
            op.f[4] = root.box.mx[0]; op.f[5] = root.box.mx[1]; op.f[6] = root.box.mx[2];
        }
        f.n_box_ops++;
        const int32_t at = pc();
        push(OP_BVH);
        // leaves in the reference's own visit order: a larger pc is "later in the reference's order" (the tie rule)
        std::vector<int32_t> order;
        leaf_order_of(o.bvh, o.bvh.root, order);
        std::vector<int32_t> leaf_pc(s.objects.size(), -1);
        for (int32_t id : order) {
            leaf_pc[id] = pc();
            if (!emit(id, ctx)) return false;
        }
        if (pc() >= (1 << 24)) { error = "op stream longer than 2^24 records"; return false; }
        BvhTree alt;
        alt.root = build_bvh_sah(s, alt, o.children, o.t0, o.t1, 0);
        const int32_t base = (int32_t)f.nodes.size();
        int max_depth = 0;
        emit_tree_nodes(alt, alt.root, leaf_pc, base, 1, max_depth);
        if (max_depth > kBvh2Stack) { error = "BVH tree deeper than the traversal stack"; return false; }
        Op& op = f.ops[at];
        op.i[0] = base;
        op.i[1] = (int32_t)f.nodes.size() - base;
        op.i[2] = (int32_t)order.size();
        op.i[3] = max_depth;
        op.f[4] = o.t0; op.f[5] = o.t1;  // the BvhNode's time interval: moving-sphere leaf boxes depend on it
        op.i[6] = in_medium ? -1 : (int32_t)f.trees.size();
        if (!in_medium) {
            PreTree t;
            t.pc = at; t.ctx = ctx; t.base = base; t.ts = o.t0; t.te = o.t1;
            for (int a = 0; a < 3; ++a) { t.mn[a] = root.box.mn[a]; t.mx[a] = root.box.mx[a]; }
            t.last_pc = leaf_pc[order.back()];
            t.last_face = s.objects[order.back()].kind == OBJ_CUBOID ? 5 : 0;
            f.trees.push_back(t);
        }
        op.u[7] = OP_BVH | ((uint32_t)pc() << 8);
        f.ops[box_at].u[7] = OP_BOX | ((uint32_t)pc() << 8);
        f.n_bvh_trees++;
        f.max_tree_depth = std::max(f.max_tree_depth, max_depth);
        return true;
    }
    static void leaf_order_of(const BvhTree& t, int32_t node, std::vector<int32_t>& out) {
        const BvhTreeNode& n = t.nodes[node];
        if (n.leaf_obj >= 0) out.push_back(n.leaf_obj);
        else { leaf_order_of(t, n.left, out); leaf_order_of(t, n.right, out); }
    }

    bool emit_bvh(const BvhTree& tree, float t0, float t1, int32_t node_index, int32_t ctx) {
        const BvhTreeNode& n = tree.nodes[node_index];
        // Soundness: does the reference box contain everything hit() can accept beneath this node?
        Box3 truth;
        bool sound = true_extent(tree, t0, t1, node_index, truth) && contains(n.box, truth);
        // The fast form leaves out sound boxes that rarely spare a warp anything.  A sound box only prunes — whatever it
        // rejects cannot be hit beneath it, and a box beneath it that is kept rejects at least as much on every axis
        // (nested intervals, monotone rounding), exact ties included — so the hits do not change.  The warp-uniform walk
        // executes every record ANY of its rays reaches: an inner box of a small BVH (the 11-object top level of `final`)
        // is reached by some ray of nearly every warp, so it is a step for the warp that spares only single lanes the leaf
        // boxes beneath; without the inner boxes all lanes meet the same leaf boxes together.  Leaf boxes stay (they do
        // spare whole warps the primitive), except above an OP_BVH tree, whose own root box is the same box.
        bool elide = false;
        if (trees && sound && elide_boxes) {
            if (n.leaf_obj < 0) elide = leaves_of(tree, node_index) <= kElideInnerMaxLeaves;
            else elide = s.objects[n.leaf_obj].kind == OBJ_BVH && tree_eligible(s.objects[n.leaf_obj]);
        }
        int32_t at = pc();
        if (!elide) {
            Op& op = push(sound ? OP_BOX : OP_BOX_LOOSE);
            op.f[0] = n.box.mn[0]; op.f[1] = n.box.mn[1]; op.f[2] = n.box.mn[2];
            op.f[4] = n.box.mx[0]; op.f[5] = n.box.mx[1]; op.f[6] = n.box.mx[2];
            f.n_box_ops++;
            if (!sound) f.n_loose_boxes++;
        }
        if (n.leaf_obj >= 0) {
            if (!emit(n.leaf_obj, ctx)) return false;
        } else {
            if (!emit_bvh(tree, t0, t1, n.left, ctx)) return false;   // left first (bvh_node.rs:111)
            if (!emit_bvh(tree, t0, t1, n.right, ctx)) return false;
        }
        if (pc() >= (1 << 24)) { error = "op stream longer than 2^24 records"; return false; }
        if (!elide) f.ops[at].u[7] = (f.ops[at].u[7] & 0xffu) | ((uint32_t)pc() << 8);
        return true;
    }

    static int32_t leaves_of(const BvhTree& tree, int32_t node_index) {
        const BvhTreeNode& n = tree.nodes[node_index];
        return n.leaf_obj >= 0 ? 1 : leaves_of(tree, n.left) + leaves_of(tree, n.right);
    }

    bool true_extent(const BvhTree& tree, float t0, float t1, int32_t node_index, Box3& out) {
        const BvhTreeNode& n = tree.nodes[node_index];
        if (n.leaf_obj >= 0) return ref_box(s, n.leaf_obj, t0, t1, true, out);
        Box3 a, b;
        if (!true_extent(tree, t0, t1, n.left, a) || !true_extent(tree, t0, t1, n.right, b)) return false;
        out = surrounding_box(a, b);
        return true;
    }
};

// ---- fp16 with directed rounding (Bvh2Node boxes, hrt_types.h) ----
// Largest fp16 <= x (up == false) or smallest fp16 >= x (up == true); +-inf when x is outside the finite fp16 range on
// the side that keeps the inequality.  x must not be NaN.
static uint16_t float_to_half_directed(float x, bool up) {
    if (x == 0.0f) return 0;
    const bool neg = x < 0.0f;
    const float a = std::fabs(x);
    // the magnitude rounds towards zero when the direction points at zero, away from it otherwise
    const bool away = neg ? !up : up;
    uint16_t mag;
    if (a > 65504.0f) {
        mag = (away || std::isinf(a)) ? 0x7c00u : 0x7bffu;  // inf, or the largest finite
    } else {
        int e;
        std::frexp(a, &e);    // a = f * 2^e, f in [0.5, 1)
        int he = e - 1 + 15;  // biased fp16 exponent of a normal
        float scaled;
        if (he <= 0) { he = 0; scaled = std::ldexp(a, 24); }  // subnormal: units of 2^-24
        else scaled = std::ldexp(a, 10 - (e - 1));            // in [1024, 2048)
        const float fl = std::floor(scaled);
        uint32_t q = (uint32_t)fl;
        if (away && fl != scaled) q += 1;
        // q may reach 2048 (or 1024 from the subnormal side): the bit-pattern sum carries into the exponent
        mag = he == 0 ? (uint16_t)q : (uint16_t)(((uint32_t)he << 10) + (q - 1024u));
    }
    return (uint16_t)(mag | (neg ? 0x8000u : 0u));
}

}  // namespace hrt

using namespace hrt;

hrt_scene::~hrt_scene() {
    for (DeviceState* d : devices)
        if (d) release_device_state(d);
}

static bool tex_ok(const hrt_scene* s, int32_t i) { return i >= 0 && (size_t)i < s->textures.size(); }
static bool mat_ok(const hrt_scene* s, int32_t i) { return i >= 0 && (size_t)i < s->materials.size(); }
static bool obj_ok(const hrt_scene* s, int32_t i) { return i >= 0 && (size_t)i < s->objects.size(); }
static bool tex_needs_uv(const hrt_scene* s, int32_t t) {
    const Texture& x = s->textures[t];
    if (x.kind == TEX_IMAGE) return true;
    if (x.kind == TEX_CHECKER) return tex_needs_uv(s, x.i0) || tex_needs_uv(s, x.i1);
    return false;
}
#define HRT_CHECK_SCENE(s)                                                        \
    if (!(s)) return fail(HRT_ERR_INVALID, "null scene");                          \
    if ((s)->committed) return fail(HRT_ERR_STATE, "scene is already committed (immutable)")

static int32_t add_obj(hrt_scene* s, Obj&& o) {
    s->objects.push_back(std::move(o));
    return (int32_t)s->objects.size() - 1;
}
static int32_t add_mat(hrt_scene* s, const Material& m) {
    s->materials.push_back(m);
    return (int32_t)s->materials.size() - 1;
}

namespace hrt {
int32_t add_medium_with_material(hrt_scene* s, int32_t boundary, float density, int32_t mat) {
    if (!s || s->committed || !obj_ok(s, boundary) || !mat_ok(s, mat)) return fail(HRT_ERR_INVALID, "constant_medium: bad argument");
    Obj o;
    o.kind = OBJ_MEDIUM;
    o.child = boundary;
    o.density = density;
    o.neg_inv_density = -1.0f / density;
    o.mat = mat;
    return add_obj(s, std::move(o));
}
}  // namespace hrt

extern "C" {

const char* hrt_last_error(void) { return g_last_error.c_str(); }
int32_t hrt_abi_version(void) { return HRT_ABI_VERSION; }

int32_t hrt_scene_create(hrt_scene** out) {
    if (!out) return fail(HRT_ERR_INVALID, "null out pointer");
    *out = new hrt_scene();
    if (const char* env = std::getenv("HRT_BVH_BUILDER"))  // diagnostic default for scenes the caller does not configure
        (*out)->bvh_builder = std::strcmp(env, "reference") == 0 ? HRT_BVH_REFERENCE : HRT_BVH_TREES;
    return HRT_OK;
}
void hrt_scene_destroy(hrt_scene* scene) { delete scene; }

int32_t hrt_tex_solid(hrt_scene* s, const float rgb[3]) {
    HRT_CHECK_SCENE(s);
    if (!rgb) return fail(HRT_ERR_INVALID, "null rgb");
    Texture t;
    std::memset(&t, 0, sizeof(t));
    t.kind = TEX_SOLID;
    t.v[0] = rgb[0]; t.v[1] = rgb[1]; t.v[2] = rgb[2];
    s->textures.push_back(t);
    return (int32_t)s->textures.size() - 1;
}
int32_t hrt_tex_checker(hrt_scene* s, int32_t odd, int32_t even) {
    HRT_CHECK_SCENE(s);
    if (!tex_ok(s, odd) || !tex_ok(s, even)) return fail(HRT_ERR_INVALID, "checker: unknown texture id");
    Texture t;
    std::memset(&t, 0, sizeof(t));
    t.kind = TEX_CHECKER;
    t.i0 = odd; t.i1 = even;
    s->textures.push_back(t);
    return (int32_t)s->textures.size() - 1;
}
int32_t hrt_tex_noise(hrt_scene* s, float scale, const float* ranvec, const uint32_t* px, const uint32_t* py,
                      const uint32_t* pz) {
    HRT_CHECK_SCENE(s);
    if (!ranvec || !px || !py || !pz) return fail(HRT_ERR_INVALID, "noise: null table");
    NoiseTable nt;
    std::memset(&nt, 0, sizeof(nt));
    for (int i = 0; i < 256; ++i) {
        nt.ranvec[i][0] = ranvec[3 * i]; nt.ranvec[i][1] = ranvec[3 * i + 1]; nt.ranvec[i][2] = ranvec[3 * i + 2];
        if (px[i] > 255 || py[i] > 255 || pz[i] > 255) return fail(HRT_ERR_INVALID, "noise: permutation entry > 255");
        nt.perm[0][i] = (uint8_t)px[i]; nt.perm[1][i] = (uint8_t)py[i]; nt.perm[2][i] = (uint8_t)pz[i];
    }
    s->noise_tables.push_back(nt);
    Texture t;
    std::memset(&t, 0, sizeof(t));
    t.kind = TEX_NOISE;
    t.v[0] = scale;
    t.i0 = (int32_t)s->noise_tables.size() - 1;
    s->textures.push_back(t);
    return (int32_t)s->textures.size() - 1;
}
int32_t hrt_tex_image(hrt_scene* s, const uint8_t* data, uint32_t width, uint32_t height, uint32_t components) {
    HRT_CHECK_SCENE(s);
    Texture t;
    std::memset(&t, 0, sizeof(t));
    t.kind = TEX_IMAGE;
    t.i0 = -1;
    if (data && width && height) {
        if (components != 3 && components != 4)
            return fail(HRT_ERR_UNSUPPORTED, "image: components must be 3 or 4 (the reference reads 3 consecutive bytes per texel)");
        if ((int)s->images.size() >= kMaxImages) return fail(HRT_ERR_UNSUPPORTED, "more than 8 image textures");
        if (width > 32768 || height > 32768) return fail(HRT_ERR_UNSUPPORTED, "image larger than 32768 in a dimension");
        ImageData img;
        img.width = width; img.height = height;
        img.rgba.resize((size_t)width * height * 4);
        for (size_t p = 0; p < (size_t)width * height; ++p) {
            img.rgba[4 * p + 0] = data[components * p + 0];
            img.rgba[4 * p + 1] = data[components * p + 1];
            img.rgba[4 * p + 2] = data[components * p + 2];
            img.rgba[4 * p + 3] = 255;
        }
        s->images.push_back(std::move(img));
        t.i0 = (int32_t)s->images.size() - 1;
        t.i1 = (int32_t)width; t.i2 = (int32_t)height;
    }
    s->textures.push_back(t);
    return (int32_t)s->textures.size() - 1;
}

int32_t hrt_mat_lambertian(hrt_scene* s, int32_t tex) {
    HRT_CHECK_SCENE(s);
    if (!tex_ok(s, tex)) return fail(HRT_ERR_INVALID, "lambertian: unknown texture id");
    Material m;
    std::memset(&m, 0, sizeof(m));
    m.kind = MAT_LAMBERTIAN; m.tex = tex;
    m.flags = tex_needs_uv(s, tex) ? MATF_NEEDS_UV : 0;
    return add_mat(s, m);
}
int32_t hrt_mat_metal(hrt_scene* s, const float albedo[3], float fuzz) {
    HRT_CHECK_SCENE(s);
    if (!albedo) return fail(HRT_ERR_INVALID, "null albedo");
    Material m;
    std::memset(&m, 0, sizeof(m));
    m.kind = MAT_METAL; m.tex = -1;
    m.albedo[0] = albedo[0]; m.albedo[1] = albedo[1]; m.albedo[2] = albedo[2];
    m.param = fuzz;
    return add_mat(s, m);
}
int32_t hrt_mat_dielectric(hrt_scene* s, float ior) {
    HRT_CHECK_SCENE(s);
    Material m;
    std::memset(&m, 0, sizeof(m));
    m.kind = MAT_DIELECTRIC; m.tex = -1; m.param = ior;
    return add_mat(s, m);
}
int32_t hrt_mat_diffuse_light(hrt_scene* s, int32_t tex) {
    HRT_CHECK_SCENE(s);
    if (!tex_ok(s, tex)) return fail(HRT_ERR_INVALID, "diffuse_light: unknown texture id");
    Material m;
    std::memset(&m, 0, sizeof(m));
    m.kind = MAT_DIFFUSE_LIGHT; m.tex = tex;
    m.flags = tex_needs_uv(s, tex) ? MATF_NEEDS_UV : 0;
    return add_mat(s, m);
}

int32_t hrt_sphere(hrt_scene* s, const float c[3], float radius, int32_t mat) {
    HRT_CHECK_SCENE(s);
    if (!c || !mat_ok(s, mat)) return fail(HRT_ERR_INVALID, "sphere: bad argument");
    Obj o;
    o.kind = OBJ_SPHERE;
    std::memcpy(o.c0, c, 12);
    o.r = radius; o.mat = mat;
    return add_obj(s, std::move(o));
}
int32_t hrt_moving_sphere(hrt_scene* s, const float c0[3], const float c1[3], float t0, float t1, float radius,
                          int32_t mat) {
    HRT_CHECK_SCENE(s);
    if (!c0 || !c1 || !mat_ok(s, mat)) return fail(HRT_ERR_INVALID, "moving_sphere: bad argument");
    Obj o;
    o.kind = OBJ_MSPHERE;
    std::memcpy(o.c0, c0, 12);
    std::memcpy(o.c1, c1, 12);
    o.t0 = t0; o.t1 = t1; o.r = radius; o.mat = mat;
    return add_obj(s, std::move(o));
}
int32_t hrt_rect(hrt_scene* s, int32_t plane, float a0, float a1, float b0, float b1, float k, int32_t mat) {
    HRT_CHECK_SCENE(s);
    if (plane < 0 || plane > 2 || !mat_ok(s, mat)) return fail(HRT_ERR_INVALID, "rect: bad argument");
    Obj o;
    o.kind = OBJ_RECT;
    o.plane_or_axis = plane; o.a0 = a0; o.a1 = a1; o.b0 = b0; o.b1 = b1; o.k = k; o.mat = mat;
    return add_obj(s, std::move(o));
}
int32_t hrt_cuboid(hrt_scene* s, const float mn[3], const float mx[3], int32_t mat) {
    HRT_CHECK_SCENE(s);
    if (!mn || !mx || !mat_ok(s, mat)) return fail(HRT_ERR_INVALID, "cuboid: bad argument");
    Obj o;
    o.kind = OBJ_CUBOID;
    std::memcpy(o.c0, mn, 12);
    std::memcpy(o.c1, mx, 12);
    o.mat = mat;
    return add_obj(s, std::move(o));
}
int32_t hrt_translate(hrt_scene* s, int32_t child, const float d[3]) {
    HRT_CHECK_SCENE(s);
    if (!d || !obj_ok(s, child)) return fail(HRT_ERR_INVALID, "translate: bad argument");
    Obj o;
    o.kind = OBJ_TRANSLATE;
    o.child = child;
    std::memcpy(o.c0, d, 12);
    return add_obj(s, std::move(o));
}
int32_t hrt_rotate(hrt_scene* s, int32_t axis, int32_t child, float degrees) {
    HRT_CHECK_SCENE(s);
    if (axis < 0 || axis > 2 || !obj_ok(s, child)) return fail(HRT_ERR_INVALID, "rotate: bad argument");
    Obj o;
    o.kind = OBJ_ROTATE;
    o.child = child;
    o.plane_or_axis = axis;
    o.angle_degrees = degrees;
    float radians = (PI_F / 180.0f) * degrees;  // rotation.rs:40-42
    o.sin_theta = sinf(radians);
    o.cos_theta = cosf(radians);
    Box3 b;
    o.has_rot_box = ref_box(*s, child, 0.0f, 1.0f, false, b);
    if (o.has_rot_box) o.rot_box = rotate_box(b, axis, o.sin_theta, o.cos_theta);
    return add_obj(s, std::move(o));
}
int32_t hrt_constant_medium(hrt_scene* s, int32_t boundary, float density, int32_t tex) {
    HRT_CHECK_SCENE(s);
    if (!obj_ok(s, boundary) || !tex_ok(s, tex)) return fail(HRT_ERR_INVALID, "constant_medium: bad argument");
    Material m;
    std::memset(&m, 0, sizeof(m));
    m.kind = MAT_ISOTROPIC; m.tex = tex;
    m.flags = tex_needs_uv(s, tex) ? MATF_NEEDS_UV : 0;
    Obj o;
    o.kind = OBJ_MEDIUM;
    o.child = boundary;
    o.density = density;
    o.neg_inv_density = -1.0f / density;  // constant_medium.rs:27
    o.mat = add_mat(s, m);
    return add_obj(s, std::move(o));
}
int32_t hrt_list(hrt_scene* s, const int32_t* children, int32_t n) {
    HRT_CHECK_SCENE(s);
    if (n < 0 || (n > 0 && !children)) return fail(HRT_ERR_INVALID, "list: bad argument");
    Obj o;
    o.kind = OBJ_LIST;
    for (int i = 0; i < n; ++i) {
        if (!obj_ok(s, children[i])) return fail(HRT_ERR_INVALID, "list: unknown object id");
        o.children.push_back(children[i]);
    }
    return add_obj(s, std::move(o));
}
int32_t hrt_bvh(hrt_scene* s, const int32_t* children, int32_t n, float ts, float te) {
    HRT_CHECK_SCENE(s);
    if (n <= 0 || !children) return fail(HRT_ERR_INVALID, "bvh: no elements in scene");  // bvh_node.rs:38 panics
    Obj o;
    o.kind = OBJ_BVH;
    o.t0 = ts; o.t1 = te;
    for (int i = 0; i < n; ++i) {
        if (!obj_ok(s, children[i])) return fail(HRT_ERR_INVALID, "bvh: unknown object id");
        Box3 b;
        if (!ref_box(*s, children[i], ts, te, false, b))
            return fail(HRT_ERR_INVALID, "bvh: object without a bounding box");  // bvh_node.rs:41-43,77-79 panic
        for (int a = 0; a < 3; ++a)  // box_compare's partial_cmp().unwrap() panics on NaN keys (bvh_node.rs:70-76)
            if (std::isnan(b.mn[a]) || std::isnan(b.mx[a])) return fail(HRT_ERR_INVALID, "bvh: NaN bounding box");
        o.children.push_back(children[i]);
    }
    o.bvh.root = build_bvh(*s, o.bvh, o.children, ts, te);
    return add_obj(s, std::move(o));
}

int32_t hrt_scene_set_bvh_builder(hrt_scene* s, int32_t builder) {
    HRT_CHECK_SCENE(s);
    if (builder != HRT_BVH_REFERENCE && builder != HRT_BVH_TREES) return fail(HRT_ERR_INVALID, "set_bvh_builder: unknown builder");
    s->bvh_builder = builder;
    return HRT_OK;
}

// Flatten the tree under `root` into `f` (hrt_types.h).
static int32_t flatten(const hrt_scene* s, int32_t root, FlatScene& f, bool trees) {
    f = FlatScene{};
    Ctx world;
    std::memset(&world, 0, sizeof(world));
    world.parent = -1;
    f.ctxs.push_back(world);
    Flattener fl(*s, f, trees);
    if (!fl.emit(root, 0)) {
        f = FlatScene{};
        return fail(HRT_ERR_UNSUPPORTED, "commit: " + fl.error);
    }
    fl.push(OP_END);
    // Consecutive ray-space pushes (Translation(Rotation(x))) and consecutive pops are entered / left in ONE step: the
    // first record of a run carries the run's final context and its length; the others stay in the stream as data for
    // the context replay (ray_in_ctx) but are never executed.
    auto is_push = [](uint32_t o) { return o == OP_TRANSLATE || o == OP_ROTATE; };
    const size_t n = f.ops.size();
    for (size_t i = 0; i < n;) {
        const uint32_t o = f.ops[i].u[7] & 0xffu;
        if (is_push(o) || o == OP_POP) {
            size_t j = i;
            while (j + 1 < n) {
                const uint32_t o2 = f.ops[j + 1].u[7] & 0xffu;
                if (is_push(o) ? is_push(o2) : (o2 == OP_POP)) ++j; else break;
            }
            f.ops[i].i[3] = f.ops[j].i[3];
            f.ops[i].u[7] = o | ((uint32_t)(j - i + 1) << 8);
            i = j + 1;
        } else {
            ++i;
        }
    }
    // The records around each OP_BVH tree that exist only for it (PreTree::from_pc / to_pc): enclosing sound boxes that
    // cover nothing else (they only prune, and the tree's own closest hit decides what they would have pruned), and whole
    // runs of ray-space pushes in front matched by as many pops behind.
    for (PreTree& t : f.trees) {
        auto opc = [&](int32_t i) { return f.ops[i].u[7] & 0xffu; };
        auto payload = [&](int32_t i) { return (int32_t)(f.ops[i].u[7] >> 8); };
        int32_t from = t.pc - 1, to = payload(t.pc);
        for (;;) {
            if (from >= 1 && opc(from - 1) == OP_BOX && payload(from - 1) == to) { --from; continue; }
            int32_t pushes = 0, pops = 0;
            while (from - pushes >= 1 && is_push(opc(from - pushes - 1))) ++pushes;
            while (to + pops < (int32_t)n && opc(to + pops) == OP_POP) ++pops;
            // only a WHOLE run of pushes (its first record is where a walk arrives and carries the run length)
            if (pushes > 0 && pushes <= pops && payload(from - pushes) == pushes) { from -= pushes; to += pushes; continue; }
            break;
        }
        // a span begins at a box (the first record of a push run stays in place: ray_in_ctx reads it as data)
        while (is_push(opc(from))) { const int32_t run = payload(from); from += run; to -= run; }
        t.from_pc = from;
        t.to_pc = to;
    }
    // The WAVE form: the stream walk of the wavefront render holds the answers of the first kMaxPreTrees trees before it
    // starts, takes each where its span begins and goes on behind it (OP_BVH_PRE, hrt_types.h).
    f.wave_ops.clear();
    if (!f.trees.empty()) {
        f.wave_ops = f.ops;
        for (size_t i = 0; i < f.trees.size() && i < (size_t)kMaxPreTrees; ++i) {
            const PreTree& t = f.trees[i];
            Op op;
            std::memset(&op, 0, sizeof(op));
            op.i[0] = (int32_t)i; op.i[1] = t.ctx; op.f[4] = t.ts; op.f[5] = t.te;
            op.u[7] = OP_BVH_PRE | ((uint32_t)t.to_pc << 8);
            f.wave_ops[(size_t)t.from_pc] = op;
        }
    }
    return HRT_OK;
}

int32_t hrt_scene_commit(hrt_scene* s, int32_t root) {
    HRT_CHECK_SCENE(s);
    if (!obj_ok(s, root)) return fail(HRT_ERR_INVALID, "commit: unknown root id");
    s->any_bvh = false;
    s->time_min = -FMAX;
    s->time_max = FMAX;
    for (const Obj& o : s->objects)
        if (o.kind == OBJ_BVH) {  // harmless over-approximation: every BVH ever built bounds the valid shutter
            s->any_bvh = true;
            s->time_min = std::fmax(s->time_min, std::fmin(o.t0, o.t1));
            s->time_max = std::fmin(s->time_max, std::fmax(o.t0, o.t1));
        }
    int32_t rc = flatten(s, root, s->ref, false);
    if (rc == HRT_OK) rc = flatten(s, root, s->fast, s->bvh_builder == HRT_BVH_TREES);
    if (rc != HRT_OK) return rc;
    if (s->ref.n_media != s->fast.n_media || s->ref.n_prim_ops != s->fast.n_prim_ops)
        return fail(HRT_ERR_UNSUPPORTED, "commit: the two flattened forms disagree (internal error)");
    s->root = root;
    s->committed = true;
    return HRT_OK;
}

int32_t hrt_scene_count(const hrt_scene* s) {
    if (!s || !s->committed) return fail(HRT_ERR_STATE, "scene not committed");
    return (int32_t)count_of(*s, s->root);
}

int32_t hrt_scene_get_info(const hrt_scene* s, hrt_scene_info* out) {
    if (!s || !out) return fail(HRT_ERR_INVALID, "null argument");
    if (!s->committed) return fail(HRT_ERR_STATE, "scene not committed");
    out->n_ops = (int32_t)s->ref.ops.size();
    out->n_box_ops = s->ref.n_box_ops;
    out->n_loose_boxes = s->ref.n_loose_boxes;
    out->n_prim_ops = s->ref.n_prim_ops;
    out->n_materials = (int32_t)s->materials.size();
    out->n_textures = (int32_t)s->textures.size();
    out->n_noise_tables = (int32_t)s->noise_tables.size();
    out->n_images = (int32_t)s->images.size();
    out->n_media = s->ref.n_media;
    out->n_contexts = (int32_t)s->ref.ctxs.size();
    out->max_context_depth = s->ref.max_ctx_depth;
    out->time_min = s->time_min;
    out->time_max = s->time_max;
    out->n_fast_ops = (int32_t)s->fast.ops.size();
    out->n_fast_box_ops = s->fast.n_box_ops;
    out->n_bvh_trees = s->fast.n_bvh_trees;
    out->n_tree_nodes = (int32_t)s->fast.nodes.size();
    out->max_tree_depth = s->fast.max_tree_depth;
    return HRT_OK;
}
static const FlatScene* pick_flat(const hrt_scene* s, int32_t which) { return which == HRT_STREAM_FAST ? &s->fast : &s->ref; }
int32_t hrt_scene_get_ops(const hrt_scene* s, int32_t which, void* out, int32_t cap_ops) {
    if (!s) return fail(HRT_ERR_INVALID, "null scene");
    if (!s->committed) return fail(HRT_ERR_STATE, "scene not committed");
    if (which != HRT_STREAM_REFERENCE && which != HRT_STREAM_FAST && which != HRT_STREAM_WAVE)
        return fail(HRT_ERR_INVALID, "get_ops: unknown form of the stream");
    const FlatScene* f = pick_flat(s, which == HRT_STREAM_WAVE ? HRT_STREAM_FAST : which);
    const std::vector<Op>& ops = (which == HRT_STREAM_WAVE && !f->wave_ops.empty()) ? f->wave_ops : f->ops;
    int32_t n = (int32_t)ops.size();
    if (out && cap_ops > 0) std::memcpy(out, ops.data(), sizeof(Op) * (size_t)std::min(n, cap_ops));
    return n;
}
int32_t hrt_scene_get_tree_nodes(const hrt_scene* s, void* out, int32_t cap_nodes) {
    if (!s) return fail(HRT_ERR_INVALID, "null scene");
    if (!s->committed) return fail(HRT_ERR_STATE, "scene not committed");
    int32_t n = (int32_t)s->fast.nodes.size();
    if (out && cap_nodes > 0) std::memcpy(out, s->fast.nodes.data(), sizeof(Bvh2Node) * (size_t)std::min(n, cap_nodes));
    return n;
}

int32_t hrt_scene_get_tree_spans(const hrt_scene* s, int32_t* out, int32_t cap_trees) {
    if (!s) return fail(HRT_ERR_INVALID, "null scene");
    if (!s->committed) return fail(HRT_ERR_STATE, "scene not committed");
    const int32_t n = (int32_t)s->fast.trees.size();
    for (int32_t i = 0; out && i < std::min(n, cap_trees); ++i) {
        const PreTree& t = s->fast.trees[(size_t)i];
        out[4 * i + 0] = t.pc; out[4 * i + 1] = t.ctx; out[4 * i + 2] = t.from_pc; out[4 * i + 3] = t.to_pc;
    }
    return n;
}

static void leaf_order(const BvhTree& t, int32_t node, std::vector<int32_t>& out) {
    const BvhTreeNode& n = t.nodes[node];
    if (n.leaf_obj >= 0) out.push_back(n.leaf_obj);
    else { leaf_order(t, n.left, out); leaf_order(t, n.right, out); }
}
int32_t hrt_bvh_leaf_order(const hrt_scene* s, int32_t bvh, int32_t* out, int32_t cap) {
    if (!s || !obj_ok(s, bvh) || s->objects[bvh].kind != OBJ_BVH) return fail(HRT_ERR_INVALID, "not a bvh object");
    std::vector<int32_t> v;
    leaf_order(s->objects[bvh].bvh, s->objects[bvh].bvh.root, v);
    for (size_t i = 0; i < v.size() && (int32_t)i < cap; ++i) out[i] = v[i];
    return (int32_t)v.size();
}
int32_t hrt_bounding_box(const hrt_scene* s, int32_t obj, float out6[6]) {
    if (!s || !obj_ok(s, obj) || !out6) return fail(HRT_ERR_INVALID, "bounding_box: bad argument");
    Box3 b;
    if (!ref_box(*s, obj, 0.0f, 1.0f, false, b)) return fail(HRT_ERR_INVALID, "object has no bounding box");
    for (int i = 0; i < 3; ++i) { out6[i] = b.mn[i]; out6[3 + i] = b.mx[i]; }
    return HRT_OK;
}

// Camera::new + Camera::resize (src/camera.rs:34-83), f32, reference operation order.
int32_t hrt_camera_init(const hrt_camera_desc* d, hrt_camera_state* out) {
    if (!d || !out) return fail(HRT_ERR_INVALID, "null argument");
    if (d->width <= 0 || d->height <= 0) return fail(HRT_ERR_INVALID, "camera: non-positive size");
    auto dot3 = [](const float* a, const float* b) { return (a[0] * b[0] + a[1] * b[1]) + a[2] * b[2]; };
    auto normalize3 = [&](const float* a, float* o) {
        float inv = 1.0f / std::sqrt(dot3(a, a));
        for (int i = 0; i < 3; ++i) o[i] = a[i] * inv;
    };
    auto cross3 = [](const float* a, const float* b, float* o) {
        o[0] = a[1] * b[2] - a[2] * b[1];
        o[1] = a[2] * b[0] - a[0] * b[2];
        o[2] = a[0] * b[1] - a[1] * b[0];
    };
    float aspect_ratio = (float)d->width / (float)d->height;
    float theta = d->vfov * (PI_F / 180.0f);
    float h = tanf(theta / 2.0f);
    float viewport_height = 2.0f * h;
    float viewport_width = aspect_ratio * viewport_height;
    float diff[3] = {d->look_from[0] - d->look_at[0], d->look_from[1] - d->look_at[1], d->look_from[2] - d->look_at[2]};
    float w[3], u[3], v[3], c[3];
    normalize3(diff, w);
    const float vup[3] = {0.0f, 1.0f, 0.0f};
    cross3(vup, w, c);
    normalize3(c, u);
    cross3(w, u, v);
    const float fw = d->focus_dist * viewport_width, fh = d->focus_dist * viewport_height;
    for (int i = 0; i < 3; ++i) {
        out->origin[i] = d->look_from[i];
        out->horizontal[i] = fw * u[i];
        out->vertical[i] = fh * v[i];
        out->u[i] = u[i]; out->v[i] = v[i]; out->w[i] = w[i];
    }
    for (int i = 0; i < 3; ++i)
        out->lower_left_corner[i] =
            ((out->origin[i] - out->horizontal[i] / 2.0f) - out->vertical[i] / 2.0f) - d->focus_dist * w[i];
    out->lens_radius = d->aperture / 2.0f;
    out->time0 = d->time0;
    out->time1 = d->time1;
    return HRT_OK;
}

}  // extern "C"
