/*
 * hrt.h — C ABI of libhrt.so: the B200-native (sm_100a CUDA) replacement for hyper-ray-tracer's
 * per-pixel path-tracing hot path.
 *
 * The reference (SkillerRaptor/hyper-ray-tracer, Rust) has no FFI or plugin interface; the seam this ABI
 * replaces is internal: `Application::render` + `Application::ray_color` (src/application.rs:393-495) and
 * every `Hittable::hit` / `Material::scatter` / `Texture::value` beneath them.  `dyn Hittable` is opaque, so
 * a flattener cannot introspect an existing tree: the front end *describes* the scene through builder calls
 * that mirror the reference constructors 1:1, then commits it.  Each entry point below cites the reference
 * interface it replaces.  INTEGRATION.md shows the Rust `extern "C"` block a maintainer would add.
 *
 * Conventions: plain C types only; ids are non-negative int32 handles into per-scene typed tables
 * (textures, materials, hittables — three separate id spaces, allocated in call order); a negative return
 * is an error code (hrt_status) and hrt_last_error() returns a thread-local message.  The library copies
 * every input array; the caller keeps ownership.  A scene is immutable after hrt_scene_commit (the
 * reference's world is an immutable Arc<Box<dyn Hittable>>, src/application.rs:68,228).  Calls on one
 * scene must be serialised by the caller; different scenes are independent.
 *
 * There is NO CPU fallback: every compute entry point fails with HRT_ERR_CUDA when no sm_100 device or no
 * CUDA runtime is available.
 */
#ifndef HRT_H
#define HRT_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define HRT_ABI_VERSION 3 /* 3: two flattened forms (reference / fast), OP_BVH trees, scene library, progressive render;
                             hrt_scene_info / hrt_scene_get_ops changed, hrt_scene_get_box16 and three render flags gone */

typedef enum hrt_status {
    HRT_OK = 0,
    HRT_ERR_INVALID = -1,     /* bad argument / unknown id / wrong call order                         */
    HRT_ERR_UNSUPPORTED = -2, /* scene shape outside what the flattener supports (message says which)  */
    HRT_ERR_CUDA = -3,        /* CUDA runtime / device failure (message carries cudaGetErrorString)    */
    HRT_ERR_STATE = -4        /* scene not committed / not uploaded                                    */
} hrt_status;

typedef struct hrt_scene hrt_scene; /* opaque */

/* rect.rs:13-17  `enum Plane { XY, YZ, ZX }`; rotation.rs:13-17 `enum Axis { X, Y, Z }` */
enum { HRT_PLANE_XY = 0, HRT_PLANE_YZ = 1, HRT_PLANE_ZX = 2 };
enum { HRT_AXIS_X = 0, HRT_AXIS_Y = 1, HRT_AXIS_Z = 2 };

/* Material kinds as reported in hit records / table dumps. */
enum { HRT_MAT_LAMBERTIAN = 0, HRT_MAT_METAL = 1, HRT_MAT_DIELECTRIC = 2, HRT_MAT_DIFFUSE_LIGHT = 3,
       HRT_MAT_ISOTROPIC = 4 };

const char* hrt_last_error(void);
int32_t hrt_abi_version(void);
/* Number of CUDA devices usable by the library (0 when there is no driver/GPU; never fails). */
int32_t hrt_device_count(void);

/* ---- scene lifetime ------------------------------------------------------------------------------- */
int32_t hrt_scene_create(hrt_scene** out);
void hrt_scene_destroy(hrt_scene* scene);

/* ---- textures (src/textures/ *.rs) ----------------------------------------------------------------- */
/* SolidColor::new(color)                                   src/textures/solid_color.rs:15 */
int32_t hrt_tex_solid(hrt_scene*, const float rgb[3]);
/* CheckerTexture::new(odd, even)                           src/textures/checker_texture.rs:16 */
int32_t hrt_tex_checker(hrt_scene*, int32_t odd_tex, int32_t even_tex);
/* NoiseTexture::new(scale) + PerlinNoise::new()            src/textures/noise_texture.rs:16, src/perlin_noise.rs:23-64.
 * The reference draws the tables from its own thread_rng; the front end passes them in:
 * ranvec = 256 x 3 f32 (unit vectors), perm_* = 256 x u32 permutations of 0..255. */
int32_t hrt_tex_noise(hrt_scene*, float scale, const float* ranvec, const uint32_t* perm_x, const uint32_t* perm_y,
                      const uint32_t* perm_z);
/* ImageTexture::new(path) — already-decoded bytes (the front end's `image` crate stays the decoder)
 *                                                          src/textures/image_texture.rs:19-32.
 * components must be 3 or 4; data may be NULL/empty (the reference then returns (1,0,1), :37-39). */
int32_t hrt_tex_image(hrt_scene*, const uint8_t* data, uint32_t width, uint32_t height, uint32_t components);

/* ---- materials (src/materials/ *.rs) --------------------------------------------------------------- */
int32_t hrt_mat_lambertian(hrt_scene*, int32_t albedo_tex);          /* lambertian.rs:21 */
int32_t hrt_mat_metal(hrt_scene*, const float albedo[3], float fuzz); /* metal.rs:23 (fuzz NOT clamped) */
int32_t hrt_mat_dielectric(hrt_scene*, float index_of_refraction);    /* dielectric.rs:23 */
int32_t hrt_mat_diffuse_light(hrt_scene*, int32_t emit_tex);          /* diffuse_light.rs:15 */

/* ---- hittables (src/hittable/ *.rs) ---------------------------------------------------------------- */
int32_t hrt_sphere(hrt_scene*, const float center[3], float radius, int32_t mat);                 /* sphere.rs:23 */
int32_t hrt_moving_sphere(hrt_scene*, const float center0[3], const float center1[3], float time0, float time1,
                          float radius, int32_t mat);                                             /* moving_sphere.rs:26 */
int32_t hrt_rect(hrt_scene*, int32_t plane, float a0, float a1, float b0, float b1, float k, int32_t mat); /* rect.rs:31 */
int32_t hrt_cuboid(hrt_scene*, const float box_min[3], const float box_max[3], int32_t mat);      /* cuboid.rs:30 */
int32_t hrt_translate(hrt_scene*, int32_t child, const float displacement[3]);                    /* translation.rs:15 */
int32_t hrt_rotate(hrt_scene*, int32_t axis, int32_t child, float angle_degrees);                 /* rotation.rs:38 */
/* ConstantMedium::new(boundary, density, texture): allocates the next material id for its Isotropic phase
 * function, exactly as the reference constructs one internally (constant_medium.rs:24-30). */
int32_t hrt_constant_medium(hrt_scene*, int32_t boundary, float density, int32_t albedo_tex);
int32_t hrt_list(hrt_scene*, const int32_t* children, int32_t n);                                 /* list.rs:14 */
/* BvhNode::new(objects, time_start, time_end)  — reproduces the reference's axis choice, centroid sort,
 * len/2 split and node boxes (bvh_node.rs:27-100); the tree is the reference's up to the order of objects with EQUAL
 * sort keys (sort_unstable_by leaves it toolchain-defined; it cannot change a hit on sound boxes).  A NaN bounding box is
 * refused (the reference panics in partial_cmp().unwrap()). */
int32_t hrt_bvh(hrt_scene*, const int32_t* children, int32_t n, float time_start, float time_end);

/* ---- scene library (src/application.rs:497-935, :132-197) ------------------------------------------- */
/* What Application::new picks per scene next to the world: Camera::new's arguments and the background colour
 * (src/application.rs:132-211; focus_dist 10, shutter [0, 1) for every scene). */
typedef struct hrt_scene_view {
    float look_from[3], look_at[3];
    float vfov, aperture, focus_dist, time0, time1;
    float background[3];
} hrt_scene_view;
/* The reference's scene generators — `--scene` of src/arguments.rs:10-19: "random" (generate_random_scene, application.rs:497),
 * "two-spheres" (:567), "two-perlin-spheres" (:589), "earth" (:604), "simple-light" (:614), "cornell" (:639),
 * "cornell-smoke" (:723), "final" (:817) — issued as builder calls on `scene` (not committed; *root_out is the world).
 * The reference draws sphere positions, box heights and perlin tables from an unseeded thread_rng, so no two of its runs
 * render the same world; here `seed` fixes the instance (the draws are those of numpy's PCG64 seeded with `seed`, 24 / 23
 * bit f32 uniforms as rand 0.8.5's gen / gen_range — hrt_rng.hpp — so the instances equal the ones the round-1 Python
 * harness generated and the committed goldens were rendered from).  `image`: the decoded texels of assets/earthmap.jpg for
 * "earth" and "final" (the front end's `image` crate stays the decoder); NULL gives the reference's empty-image colour. */
int32_t hrt_make_scene(hrt_scene*, const char* name, uint64_t seed, const uint8_t* image, uint32_t image_width, uint32_t image_height,
                       uint32_t image_components, int32_t* root_out, hrt_scene_view* view_out);
/* A scene INSTANCE on disk: every builder call made on `scene` so far (textures, materials, perlin tables, image texels,
 * hittables in creation order), the root and the view, as one flat little-endian file.  hrt_scene_load re-issues the
 * hittable calls in the same order on a fresh scene — same ids, same bounding boxes, same BvhNode trees, hence the same
 * flattened streams — and leaves it uncommitted.  (The reference never stores a world: Application::new generates it
 * and drops the generator's state, application.rs:132-199.) */
int32_t hrt_scene_save(const hrt_scene*, int32_t root, const hrt_scene_view* view, const char* path);
int32_t hrt_scene_load(const char* path, hrt_scene** out, int32_t* root_out, hrt_scene_view* view_out);

/* Flatten the tree under `root` into the device op stream + material/texture tables (host only). */
int32_t hrt_scene_commit(hrt_scene*, int32_t root);

/* `world.count()` as the reference logs it (src/application.rs:277; Rotation::count() == 1, rotation.rs:140). */
int32_t hrt_scene_count(const hrt_scene*);

/* ---- flattened-table introspection (tests, tools) -------------------------------------------------- */
typedef struct hrt_scene_info {
    /* the REFERENCE form of the op stream (every BvhNode as the reference built it) */
    int32_t n_ops;          /* 32-byte records in the op stream                              */
    int32_t n_box_ops;      /* BVH node boxes                                                 */
    int32_t n_loose_boxes;  /* boxes that must use the reference's per-axis test (unsound)    */
    int32_t n_prim_ops;     /* sphere / moving sphere / rect / cuboid records                 */
    int32_t n_materials, n_textures, n_noise_tables, n_images, n_media, n_contexts;
    int32_t max_context_depth;
    float time_min, time_max; /* BVH build interval (intersection over all hrt_bvh calls)      */
    /* the FAST form (what renders by default) */
    int32_t n_fast_ops, n_fast_box_ops;
    int32_t n_bvh_trees;      /* hrt_bvh objects flattened as OP_BVH trees                      */
    int32_t n_tree_nodes;     /* 32-byte two-child nodes of those trees                         */
    int32_t max_tree_depth;
} hrt_scene_info;
int32_t hrt_scene_get_info(const hrt_scene*, hrt_scene_info* out);
/* hrt_scene_commit flattens the scene twice.  The REFERENCE form keeps every BvhNode exactly as BvhNode::new built it
 * (longest axis, object median, bvh_node.rs:27-63) as box records visited left first — what HRT_FLAG_REFERENCE_TRAVERSAL
 * renders, and what a shutter outside the BVH build interval falls back to.  The FAST form (default for rendering) turns
 * every hrt_bvh that is SOUND — at least four children, all plain spheres / moving spheres / rects / cuboids whose
 * reference boxes contain them, i.e. no axis-swapped ZX rect — into a surface-area-heuristic binary tree walked with a
 * per-ray stack, nearer child first.  Box tests only prune, so the closest hit is the same; on an EXACT tie between two
 * surfaces the reference keeps the leaf that comes later in ITS depth-first order (bvh_node.rs:110-124: `t <= t_max`), and
 * so does the tree walk: the leaf records stay in the stream in the reference's order and an equal-t hit only replaces
 * an earlier record's.  BVHs that are not sound (the Cornell / final top levels, above the ceiling light) stay box records
 * in both forms; of those the fast form leaves out the SOUND inner boxes of small BVHs (at most 16 leaves beneath) and the
 * leaf box above a tree: a sound box only prunes, a kept box beneath it rejects at least as much on every axis, and the
 * warp-uniform walk pays for every record any of its 32 rays reaches.  Every loose box and every leaf box stays.
 *   HRT_BVH_TREES (default)  as above
 *   HRT_BVH_REFERENCE        the fast form is the reference form
 * Call before hrt_scene_commit.  hrt_bvh_leaf_order / hrt_bounding_box always describe the reference trees. */
enum { HRT_BVH_REFERENCE = 0, HRT_BVH_TREES = 1 };
int32_t hrt_scene_set_bvh_builder(hrt_scene*, int32_t builder);
enum { HRT_STREAM_REFERENCE = 0, HRT_STREAM_FAST = 1,
       HRT_STREAM_WAVE = 2 /* the fast form as the wavefront render's stream walk reads it: an OP_BVH_PRE record where the
                              span of each tree it walks ahead begins (hrt_scene_get_tree_spans); same length */ };
/* Copies up to cap_ops 32-byte records of the chosen form; returns its record count. */
int32_t hrt_scene_get_ops(const hrt_scene*, int32_t which, void* out, int32_t cap_ops);
/* Copies up to cap_nodes 32-byte tree nodes of the fast form (hrt_types.h Bvh2Node); returns the node count. */
int32_t hrt_scene_get_tree_nodes(const hrt_scene*, void* out, int32_t cap_nodes);
/* The OP_BVH trees of the fast form outside medium boundaries, in stream order: up to cap_trees rows of four int32
 * {pc of the OP_BVH record, ray-space context, from_pc, to_pc} where [from_pc, to_pc) are the records that exist only
 * for that tree (hrt_types.h PreTree: what the wavefront render's stream walk steps over once the tree has been
 * walked ahead); returns the tree count. */
int32_t hrt_scene_get_tree_spans(const hrt_scene*, int32_t* out, int32_t cap_trees);
/* DFS left->right leaf object ids of a hrt_bvh object; returns leaf count. */
int32_t hrt_bvh_leaf_order(const hrt_scene*, int32_t bvh, int32_t* out, int32_t cap);
/* Reference bounding box of any hittable over time [0,1] (what `bounding_box(0.0, 1.0)` returns). */
int32_t hrt_bounding_box(const hrt_scene*, int32_t obj, float out_min_max[6]);

/* ---- camera (src/camera.rs) ------------------------------------------------------------------------ */
typedef struct hrt_camera_desc { /* Camera::new arguments, camera.rs:34-44 */
    float look_from[3], look_at[3];
    float vfov, aperture, focus_dist, time0, time1;
    int32_t width, height;
} hrt_camera_desc;
typedef struct hrt_camera_state { /* what Camera::resize derives, camera.rs:67-83 */
    float origin[3], lower_left_corner[3], horizontal[3], vertical[3], u[3], v[3], w[3];
    float lens_radius, time0, time1;
} hrt_camera_state;
int32_t hrt_camera_init(const hrt_camera_desc*, hrt_camera_state* out);

/* ---- render (src/application.rs:393-495) ----------------------------------------------------------- */
enum {
    HRT_FLAG_REFERENCE_TRAVERSAL = 1, /* per-axis (loose) box test on every node, as aabb.rs:20-47        */
    HRT_FLAG_EXACT_MATH = 2,          /* no FMA contraction, IEEE div/sqrt, accurate libm (parity build) */
    HRT_FLAG_INTERPRETER = 8,         /* render: the persistent kernel with every lane interpreting its own ray's records
                                         (the plain form, kept as the baseline of the warp-uniform walk)        */
    HRT_FLAG_UNIFORM = 64,            /* render / hrt_trace_hits: the warp walks the op stream together (one record per
                                         step for the lanes that are at it; every branch warp-uniform).  Render: the
                                         persistent kernel (default for small jobs and scenes without OP_BVH trees) */
    HRT_FLAG_WAVEFRONT = 128          /* render: the wavefront render (default for big jobs on scenes with OP_BVH trees):
                                         path slots in device memory, a few small kernels per ray segment, tree walks
                                         compacted over the whole wave                                           */
};
typedef struct hrt_render_desc {
    int32_t width, height;
    int32_t samples;     /* total samples per pixel of the whole job (--samples)                  */
    int32_t depth;       /* --depth; 0..65535 (0 renders black, as ray_color(depth == 0) does)     */
    float background[3];
    uint64_t seed;       /* Philox key                                                            */
    int32_t sample_begin, sample_count; /* slice [begin, begin+count) rendered by THIS call (multi-GPU
                                           sharding); count <= 0 means the whole range            */
    uint32_t flags;
} hrt_render_desc;
typedef struct hrt_stats {
    uint64_t paths;       /* camera samples traced by this call                                   */
    uint64_t rays;        /* world.hit() calls issued from the bounce loop                        */
    float kernel_ms;      /* path-trace kernel, CUDA events on the launch stream                   */
    float resolve_ms;
    float h2d_ms, d2h_ms;
    int32_t launches;     /* kernels launched by this call                                         */
    int32_t grid, block;
} hrt_stats;

/* Upload the committed scene to CUDA device `device` (idempotent per device). */
int32_t hrt_scene_upload(hrt_scene*, int32_t device);

/* Whole render on one device with HOST output: out_rgba = width*height*4 f32, rows bottom-up (row 0 is the
 * bottom row, as the reference's GL upload expects), gamma-resolved sqrt(sum/samples), alpha 1.0 —
 * byte-compatible with concatenated reference `Tile.pixels` (src/application.rs:46-52,451-456). Blocking. */
int32_t hrt_render(hrt_scene*, int32_t device, const hrt_camera_desc*, const hrt_render_desc*, float* out_rgba,
                   hrt_stats* stats);
/* As hrt_render but returns the un-resolved linear sum (width*height*4 f32: r,g,b sums, w = samples). */
int32_t hrt_render_accum(hrt_scene*, int32_t device, const hrt_camera_desc*, const hrt_render_desc*, float* out_sum,
                         hrt_stats* stats);

/* Progressive delivery and early exit.  The reference hands finished tiles to the display loop one by one
 * (src/application.rs:284-306) and abandons a frame when the window is resized (:357-391).  Here the frame is rendered in
 * batches of `batch_samples` samples — disjoint slices of ONE render's sample set, accumulated on the device — and after
 * every batch the frame of the samples so far (gamma-resolved with the count so far, same layout as hrt_render) is
 * copied to `out_rgba` and `on_frame(user, samples_done, samples_total, out_rgba)` is called on the calling thread; a
 * non-zero return cancels the render (HRT_CANCELLED; out_rgba keeps the last delivered frame).  on_frame may be NULL.
 * The last frame equals hrt_render's up to f32 summation order. */
#define HRT_CANCELLED 1
typedef int32_t (*hrt_progress_fn)(void* user, int32_t samples_done, int32_t samples_total, const float* rgba);
int32_t hrt_render_progressive(hrt_scene*, int32_t device, const hrt_camera_desc*, const hrt_render_desc*, int32_t batch_samples,
                               hrt_progress_fn on_frame, void* user, float* out_rgba, hrt_stats* stats);

/* Single-process multi-GPU render on `n_devices` (1..8) CUDA devices of one node — the shape the reference needs, being one
 * process (src/main.rs:24-32): device k renders the k-th disjoint sample slice into its own accumulator, all devices run
 * concurrently, and devices[0] sums the peers' accumulators over NVLink peer memory INSIDE the gamma-resolve kernel (no
 * staging copy, no separate collective).  Output layout and blocking semantics as hrt_render / hrt_render_accum.
 * stats: paths / rays / launches summed, kernel_ms = slowest device. */
int32_t hrt_render_multi(hrt_scene*, const int32_t* devices, int32_t n_devices, const hrt_camera_desc*, const hrt_render_desc*,
                         float* out_rgba, hrt_stats* stats);
int32_t hrt_render_accum_multi(hrt_scene*, const int32_t* devices, int32_t n_devices, const hrt_camera_desc*,
                               const hrt_render_desc*, float* out_sum, hrt_stats* stats);

/* Device-resident variants for multi-GPU sample sharding: `d_accum` is a device pointer on `device` to
 * width*height*4 f32 that the call ADDS into (zero it first); `stream` is a cudaStream_t (0 = default): the work is
 * ordered after what `stream` holds at the call and before anything enqueued on it afterwards.  A big render on a scene
 * with OP_BVH trees is driven from the host in iterations (the wavefront render) and the call then returns when it is
 * complete; otherwise it only enqueues one kernel and returns (it blocks for the stats read-back when stats != NULL).
 * Renders of one scene on one device may overlap on different streams: every launch owns its counters. */
int32_t hrt_render_accum_device(hrt_scene*, int32_t device, const hrt_camera_desc*, const hrt_render_desc*,
                                void* d_accum, void* stream, hrt_stats* stats);
/* Gamma resolve (src/application.rs:451-456): d_out_rgba[i] = (sqrt(sum.rgb * (1/samples)), 1.0). */
int32_t hrt_resolve_device(int32_t device, const void* d_accum, int32_t width, int32_t height, int32_t samples,
                           void* d_out_rgba, void* stream);

/* Drop the device copy of the scene on `device` (next compute call uploads again). */
int32_t hrt_scene_evict(hrt_scene*, int32_t device);
/* Re-copy every scene table and image host->device into the EXISTING device allocations (uploads first if the scene
 * is not resident).  bench.py's end-to-end leg calls it every step so that the H2D input copy sits inside the timed region
 * without the allocator noise of evict + upload. */
int32_t hrt_scene_refresh(hrt_scene*, int32_t device);
/* Bytes hrt_scene_upload copies host->device (op stream + tables + image texels). */
int64_t hrt_scene_device_bytes(const hrt_scene*);

/* Roofline denominators measured on `device` with CUDA events (not in the reference): dependent-free FFMA
 * chains on every SM (FP32 TFLOP/s, FMA = 2 flops) and a repeated coalesced read of an L2-resident buffer. */
typedef struct hrt_peaks {
    float fp32_tflops;   /* measured FFMA throughput                                   */
    float l2_read_gbs;   /* measured L2-resident read bandwidth (32 MiB buffer)         */
    float fma_ms, l2_ms; /* durations of the two microbenchmarks                        */
    int32_t sm_count;
    int32_t clock_khz;   /* cudaDevAttrClockRate (max SM clock)                         */
} hrt_peaks;
int32_t hrt_measure_peaks(int32_t device, hrt_peaks* out);

/* ---- parity entry points --------------------------------------------------------------------------- */
typedef struct hrt_ray { float o[3], d[3], time, tmin, tmax; } hrt_ray;
typedef struct hrt_hit {
    int32_t hit;
    float t, p[3], n[3], u, v;
    int32_t front_face, material_id, prim_id, face;
} hrt_hit;
/* `world.hit(ray, tmin, tmax)` on explicit rays (host buffers).  xi[i] is the uniform returned by every
 * ConstantMedium draw on ray i (constant_medium.rs:59); may be NULL (0.5). */
int32_t hrt_trace_hits(hrt_scene*, int32_t device, const hrt_ray* rays, int32_t n, const float* xi, hrt_hit* out,
                       uint32_t flags);
/* `Texture::value(u, v, p)` on n tuples: uvp = n x 5 f32, out = n x 3 f32. */
int32_t hrt_tex_value(hrt_scene*, int32_t device, int32_t tex, const float* uvp, int32_t n, float* out, uint32_t flags);
/* `Material::scatter` + `emitted` under injected uniforms u4 (n x 4): the library's fixed-draw samplers. */
typedef struct hrt_scatter_out {
    int32_t scattered;
    float attenuation[3], o[3], d[3], time, emitted[3];
} hrt_scatter_out;
int32_t hrt_scatter(hrt_scene*, int32_t device, const hrt_ray* rays, const hrt_hit* hits, const float* u4, int32_t n,
                    hrt_scatter_out* out, uint32_t flags);
/* `Camera::get_ray(s,t)` under injected uniforms: stuuu = n x 5 (s, t, lens u1, lens u2, time u). */
int32_t hrt_camera_rays(int32_t device, const hrt_camera_desc*, const float* stuuu, int32_t n, hrt_ray* out,
                        uint32_t flags);
/* The first `n` uniforms the render kernel draws for (pixel, sample, bounce, block): Philox4x32-10 KAT. */
int32_t hrt_philox_uniforms(uint64_t seed, uint32_t pixel, uint32_t sample, uint32_t bounce, uint32_t block,
                            float out4[4]);

#ifdef __cplusplus
}
#endif
#endif /* HRT_H */
