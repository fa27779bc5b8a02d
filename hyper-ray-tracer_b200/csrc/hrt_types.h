// hrt_types.h — flat scene tables shared by the host flattener (hrt_scene.cpp) and the CUDA kernels.
//
// The reference's world is a tree of `Box<dyn Hittable>` (src/hittable/mod.rs:19-25) walked by virtual
// recursion.  Here the whole tree — every BvhNode of every (nested) BVH, every primitive, every
// Translation/Rotation and every ConstantMedium — is flattened into ONE linear "op stream" of 32-byte
// records laid out in the reference's own depth-first visit order (left before right,
// src/hittable/bvh_node.rs:110-124).  Because that order is fixed (it does not depend on the ray), the
// traversal needs no stack: a box record that is missed jumps to its `skip` index, anything else falls
// through to pc+1.  Ray-space changes (Translation / Rotation) are push/pop records around the child's
// sub-stream; a ConstantMedium record owns the sub-stream [pc+1, end) of its boundary.
#pragma once
#include <stdint.h>

namespace hrt {

// The high nibble of an opcode is its record class (box 1, sphere 2, rect 3, ray space / medium / tree 4, end 5).
enum Opcode : uint32_t {
    OP_BOX = 0x10,        // sound box: intersected ("tight") slab test is result-identical to the reference test
    OP_BOX_LOOSE = 0x11,  // unsound box (Q2): MUST use the reference's per-axis test (src/aabb.rs:20-47)
    OP_SPHERE = 0x20,
    OP_MSPHERE = 0x21,    // followed by one OP_MSPHERE_AUX record
    OP_MSPHERE_AUX = 0x2f,
    OP_RECT_XY = 0x30,
    OP_RECT_YZ = 0x31,
    OP_RECT_ZX = 0x32,
    OP_CUBOID = 0x33,
    OP_TRANSLATE = 0x40,  // enter child ray space
    OP_ROTATE = 0x41,
    OP_POP = 0x42,        // leave child ray space (restore context in w3)
    OP_MEDIUM = 0x43,
    OP_MEDIUM_SPHERE = 0x44,  // OP_MEDIUM whose boundary sub-stream is exactly one OP_SPHERE record (closed-form path)
    OP_MEDIUM_CUBOID = 0x46,  // OP_MEDIUM whose boundary sub-stream is one OP_CUBOID, bare or inside one run of ray-space pushes
                              // (Translation(Rotation(Cuboid)): the smoke blocks of Cornell-smoke) — both queries in one step
    OP_BVH = 0x45,            // a whole sound BvhNode as a binary tree of Bvh2Node records walked with a per-ray stack,
                              // nearer child first; its leaves are the primitive records [pc+1, end)
    OP_BVH_PRE = 0x47,        // WAVE form only: stands where the records that exist only for one pre-walked tree begin
                              // (PreTree::from_pc) — the tree's answer is merged there and the walk goes on at to_pc
    OP_END = 0x50,
};

// 32-byte record = two float4.  w7 (the .w of the second float4) = opcode | (payload << 8).
//   BOX*        w0-2 min            w4-6 max                  w7 = op | skip_pc<<8
//   SPHERE      w0-2 centre, w3 r   w4 mat, w5 prim_id        w7 = op
//   MSPHERE     w0-2 c0, w3 r       w4 mat, w5 prim_id        w7 = op       (+AUX: w0-2 c1, w3 t0, w4 t1)
//   RECT_*      w0-3 a0,a1,b0,b1    w4 k, w5 mat, w6 prim_id  w7 = op
//   CUBOID      w0-2 min, w3 mat    w4-6 max                  w7 = op | prim_id<<8
//   TRANSLATE   w0-2 d, w3 ctx                                 w7 = op | run<<8   } run = number of consecutive push (or
//   ROTATE      w0 sin, w1 cos, w2 axis, w3 ctx                w7 = op | run<<8   } pop) records starting here: executing the
//   POP         w3 ctx to restore                              w7 = op | run<<8   } first one enters/leaves the whole chain
//   MEDIUM      w0 -1/density, w1 mat, w2 medium idx, w3 prim  w7 = op | end_pc<<8
//   BVH         w0 first node (index into the node table), w1 node count, w2 leaf count, w3 tree depth
//               w4 time_start, w5 time_end of the BvhNode, w6 index of the tree among those outside medium
//               boundaries (stream order; -1 inside one)       w7 = op | end_pc<<8
//   BVH_PRE     w0 index of the tree among the pre-walked ones, w1 its ray-space context
//               w4 time_start, w5 time_end                     w7 = op | to_pc<<8
struct alignas(16) Op {
    union {
        float f[8];
        uint32_t u[8];
        int32_t i[8];
    };
};
static_assert(sizeof(Op) == 32, "op record must be 32 bytes");

// Node of an OP_BVH tree: the boxes of BOTH children rounded OUTWARD to fp16 (min down, max up — any superset of a sound
// box is sound, so results do not change; +-inf beyond the fp16 range) and the two
// child links — >= 0: another node (index relative to the tree's first node), < 0: ~pc of the leaf's primitive record in
// the op stream.  The reference visits a BvhNode's children left first with a running t_max and lets the later child win
// an exact tie if it is still reached (bvh_node.rs:110-124); a sound tree can be walked in ANY order with the same
// closest hit as long as exact ties are settled the reference's way.  The leaf records of a tree sit in the stream in
// the reference's own depth-first order, so "later in the reference's order" is "larger pc" (hrt_device.cuh
// tie_goes_to_later).
struct alignas(16) Bvh2Node {
    uint16_t lbox[6];  // left child: min xyz rounded down, max xyz rounded up
    int32_t left;
    uint16_t rbox[6];
    int32_t right;
};
static_assert(sizeof(Bvh2Node) == 32, "bvh2 node must be 32 bytes");
constexpr int kBvh2SahDepth = 24;  // SAH splits down to this depth, then balanced median splits: depth <= 24 + log2(n)
constexpr int kBvh2Stack = 48;     // per-ray stack entries (one per tree level at most)
constexpr int kMaxPreTrees = 2;    // OP_BVH trees the wavefront render walks in its own, compacted stage AHEAD of the stream
                                   // walk (the first ones of the world ray's stream; later ones are walked inline)
// Where such a tree sits (render kernel parameters): its record, the ray space it lives in and its root box.
struct PreTree {
    int32_t pc;       // the OP_BVH record (its OP_BOX root record is pc - 1)
    int32_t ctx;      // ray-space context of the record
    int32_t base;     // first node
    float ts, te;     // the BvhNode's time interval
    float mn[3], mx[3];
    int32_t last_pc;  // the tree's last leaf record in the reference's order and, when it is a cuboid, its last side (5):
    int32_t last_face;  // what a ray whose origin or direction is all NaN "hits" (every test accepts it, the last one wins)
    // [from_pc, to_pc): the records that exist ONLY for this tree — its root box and OP_BVH record, the leaves, and around
    // them every sound OP_BOX whose skip is to_pc and every run of ray-space pushes matched by the pops behind the leaves.
    // A walk that already holds the tree's answer takes it at from_pc and goes on at to_pc (traverse_uniform).
    int32_t from_pc, to_pc;
};

constexpr int kMaxCtxDepth = 6;
// A ray-space context = the chain of TRANSLATE/ROTATE records (outermost first) that maps the world ray
// into it.  ctx 0 is world space.
struct alignas(16) Ctx {
    int32_t depth;
    int32_t parent;
    int32_t op_pc[kMaxCtxDepth];
};
static_assert(sizeof(Ctx) == 32, "ctx record must be 32 bytes");

enum MaterialKind : int32_t { MAT_LAMBERTIAN = 0, MAT_METAL = 1, MAT_DIELECTRIC = 2, MAT_DIFFUSE_LIGHT = 3, MAT_ISOTROPIC = 4 };
enum MaterialFlags : int32_t { MATF_NEEDS_UV = 1 };
struct alignas(16) Material {
    float albedo[3];  // metal
    float param;      // metal fuzz | dielectric ior
    int32_t kind;
    int32_t tex;      // lambertian / isotropic albedo, diffuse-light emit
    int32_t flags;
    int32_t pad;
};
static_assert(sizeof(Material) == 32, "material record must be 32 bytes");

enum TextureKind : int32_t { TEX_SOLID = 0, TEX_CHECKER = 1, TEX_NOISE = 2, TEX_IMAGE = 3 };
struct alignas(16) Texture {
    float v[4];   // solid: rgb | noise: scale
    int32_t kind;
    int32_t i0;   // checker: odd  | noise: table | image: image index (-1 = empty data)
    int32_t i1;   // checker: even | image: width
    int32_t i2;   //                 image: height
};
static_assert(sizeof(Texture) == 32, "texture record must be 32 bytes");

// src/perlin_noise.rs:13-18: 256 gradient vectors + three 256-entry permutations (values < 256 -> u8).
struct alignas(16) NoiseTable {
    float ranvec[256][4];
    uint8_t perm[3][256];
};
static_assert(sizeof(NoiseTable) == 4096 + 768, "noise table size");

constexpr int kMaxImages = 8;
constexpr int kMaxNoiseTablesShared = 2;  // staged in shared memory by the render kernel

}  // namespace hrt
