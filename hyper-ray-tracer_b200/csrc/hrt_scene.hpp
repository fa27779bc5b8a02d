// hrt_scene.hpp — host-side scene description + flattener state behind the opaque `hrt_scene` handle.
#pragma once
#include <cstdint>
#include <memory>
#include <string>
#include <vector>

#include "hrt_types.h"

namespace hrt {

struct Box3 {
    float mn[3], mx[3];
};

enum ObjKind : int32_t {
    OBJ_SPHERE, OBJ_MSPHERE, OBJ_RECT, OBJ_CUBOID, OBJ_TRANSLATE, OBJ_ROTATE, OBJ_MEDIUM, OBJ_LIST, OBJ_BVH
};

struct BvhTreeNode {
    int32_t left = -1, right = -1;  // indices into BvhTree::nodes (Branch)
    int32_t leaf_obj = -1;          // object id (Leaf)
    Box3 box;                       // reference box (bvh_node.rs:41-59)
};
struct BvhTree {
    std::vector<BvhTreeNode> nodes;
    int32_t root = -1;
};

struct Obj {
    ObjKind kind;
    float c0[3] = {0, 0, 0}, c1[3] = {0, 0, 0};  // centre(s) | box min/max | displacement
    float r = 0, t0 = 0, t1 = 1;                 // radius | time interval
    float a0 = 0, a1 = 0, b0 = 0, b1 = 0, k = 0; // rect
    float sin_theta = 0, cos_theta = 1;          // rotate
    float neg_inv_density = 0;                   // medium
    float density = 0, angle_degrees = 0;        // as passed to the builder (scene-instance files)
    int32_t plane_or_axis = 0;
    int32_t mat = -1;
    int32_t child = -1;
    std::vector<int32_t> children;               // list / bvh input order
    bool has_rot_box = false;
    Box3 rot_box;                                // Rotation caches its box at construction (rotation.rs:43-89)
    BvhTree bvh;
};

struct ImageData {
    std::vector<uint8_t> rgba;  // expanded to RGBA8
    uint32_t width = 0, height = 0;
};

// One flattened form of the scene (hrt_types.h): the op stream, the ray-space contexts whose records it refers to, and
// the node table of its OP_BVH trees.
struct FlatScene {
    std::vector<Op> ops;
    std::vector<Ctx> ctxs;
    std::vector<Bvh2Node> nodes;
    std::vector<PreTree> trees;  // the OP_BVH trees outside medium boundaries, in stream order
    std::vector<Op> wave_ops;    // the WAVE form: `ops` with an OP_BVH_PRE record at the from_pc of each of the first
                                 // kMaxPreTrees trees (what the wavefront render's stream walk reads; empty without trees)
    int32_t n_box_ops = 0, n_loose_boxes = 0, n_prim_ops = 0, n_media = 0, max_ctx_depth = 0;
    int32_t n_bvh_trees = 0, max_tree_depth = 0;
};

struct DeviceState;  // defined in hrt_api.cu

}  // namespace hrt

struct hrt_scene {
    std::vector<hrt::Texture> textures;
    std::vector<hrt::Material> materials;
    std::vector<hrt::Obj> objects;
    std::vector<hrt::NoiseTable> noise_tables;
    std::vector<hrt::ImageData> images;

    int32_t bvh_builder = 1;  // hrt_scene_set_bvh_builder: 0 reference trees only, 1 OP_BVH trees for sound BVHs

    // committed (flattened) forms: `ref` keeps every BvhNode as the reference built it (left-first box records); `fast`
    // is what renders by default — sound BVHs of plain primitives become OP_BVH trees.  Same primitives, materials,
    // medium indices and hit results; `ref` is what HRT_FLAG_REFERENCE_TRAVERSAL and out-of-interval shutters use.
    bool committed = false;
    int32_t root = -1;
    hrt::FlatScene ref, fast;
    float time_min = -3.402823466e38f, time_max = 3.402823466e38f;  // intersection of BVH build intervals
    bool any_bvh = false;

    std::vector<hrt::DeviceState*> devices;  // per CUDA device, filled by hrt_scene_upload

    ~hrt_scene();
};

namespace hrt {
void set_error(const std::string& msg);
int32_t fail(int32_t code, const std::string& msg);
void release_device_state(DeviceState*);
// hrt_constant_medium with its Isotropic material already in the material table (hrt_scene_load)
int32_t add_medium_with_material(hrt_scene*, int32_t boundary, float density, int32_t mat);
}  // namespace hrt
