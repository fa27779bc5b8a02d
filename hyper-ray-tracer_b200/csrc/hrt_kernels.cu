// hrt_kernels.cu — sm_100a kernels of the path-tracing hot path + their launchers.
// Compiled twice (see hrt_device.cuh): -DHRT_EXACT=1 --fmad=false and -DHRT_EXACT=0.
//
// Application::render's sample loop + ray_color (src/application.rs:393-495) has two implementations that trace the same
// paths (same Philox streams, same per-ray record order, same arithmetic):
//   render_interp_kernel<true>   persistent: warps pull (8x4-pixel tile, sample-chunk) work items from a global cursor,
//                                every lane owns a path, world.hit is the warp-uniform walk of the op stream
//                                (traverse_uniform, hrt_device.cuh).  Small jobs and scenes without OP_BVH trees.
//   render_interp_kernel<false>  the same with every lane interpreting its own ray's records (traverse<>): the plain form,
//                                kept as the baseline the uniform walk is measured against.
//   wave_logic / wave_noise / wave_tree / wave_trace / wave_finish
//                                the wavefront render: path slots in device memory, a few small kernels per ray segment,
//                                tree walks compacted over the whole wave.  Big jobs on scenes with OP_BVH trees.
//   resolve_kernel         the gamma resolve sqrt(sum * 1/spp), alpha 1 (src/application.rs:451-456)
//   reduce_resolve_kernel  the same, summing the accumulators of several devices over peer memory (hrt_render_multi)
//   trace_hits_kernel / trace_hits_uniform_kernel   world.hit() on explicit rays   (parity entry)
//   tex_value_kernel   Texture::value                          (parity entry)
//   scatter_kernel     Material::scatter / emitted             (parity entry)
//   camera_rays_kernel Camera::get_ray                         (parity entry)
#include "hrt_device.cuh"
#include "hrt_launch.h"

namespace HRT_NS {

static_assert(sizeof(DeviceScene) == sizeof(hrt::DeviceSceneHost), "DeviceScene host/device mirror mismatch");

constexpr int kBlock = 256;
constexpr int kWarpsPerBlock = kBlock / 32;
constexpr unsigned kFull = 0xffffffffu;

constexpr int kMaxTailChunks = 12;

// Sample range of chunk `c` (relative to sample_begin): big chunks first, the tail in shrinking chunks, so the last work
// items handed out by the global cursor are short and the end-of-kernel tail stays small.
struct RenderParams;

struct RenderParams {
    DeviceScene S;
    CameraK cam;
    int width, height, depth;
    float bg[3];
    uint32_t k0, k1;
    int sample_begin, sample_count;
    int chunk, n_chunks, tiles_x, tiles_y, n_tiles, n_items;
    // guided schedule: n_big chunks of `chunk` samples, then up to kMaxTailChunks geometrically smaller ones
    int n_big, tail_begin[kMaxTailChunks], tail_size[kMaxTailChunks];
    int reference_boxes;

    unsigned long long* counters;
    float4* accum;
};

// Copies the first min(n_noise, cap) noise tables into the block's shared memory at `dst` (ends with a block barrier).
__device__ __forceinline__ void stage_noise(const DeviceScene& S, TexEnv& E, NoiseTable* dst_tables, int cap) {
    E.sh_noise = dst_tables;
    E.n_shared_noise = S.n_noise < cap ? S.n_noise : cap;
    const int words = E.n_shared_noise * (int)(sizeof(NoiseTable) / 16);
    const uint4* src = reinterpret_cast<const uint4*>(S.noise);
    uint4* dst = reinterpret_cast<uint4*>(dst_tables);
    for (int i = threadIdx.x; i < words; i += blockDim.x) dst[i] = __ldg(src + i);
    __syncthreads();
}

// Sample range of work item `item` (tile-major inside a chunk; big chunks first, then the shrinking tail chunks).
__device__ __forceinline__ void decode_item(const RenderParams& P, unsigned long long item, int& tx, int& ty, int& s0, int& s_n) {
    const int tile = (int)(item % (unsigned long long)P.n_tiles);
    const int chunk = (int)(item / (unsigned long long)P.n_tiles);
    tx = tile % P.tiles_x;
    ty = tile / P.tiles_x;
    const int tc = chunk - P.n_big;  // >= 0: one of the shrinking tail chunks
    s0 = P.sample_begin + (tc < 0 ? chunk * P.chunk : P.tail_begin[tc]);
    s_n = tc < 0 ? P.chunk : P.tail_size[tc];
}
// Adds the item's radiance sums (lane = pixel of the 8x4 tile) into the frame accumulator.
__device__ __forceinline__ void flush_item(const RenderParams& P, int tx, int ty, int s_n, float (*acc)[3], int lane) {
    const int px = tx * 8 + (lane & 7), py = ty * 4 + (lane >> 3);
    if (px < P.width && py < P.height) {
        float* dst = reinterpret_cast<float*>(P.accum + (size_t)py * P.width + px);
        atomicAdd(dst + 0, acc[lane][0]);
        atomicAdd(dst + 1, acc[lane][1]);
        atomicAdd(dst + 2, acc[lane][2]);
        atomicAdd(dst + 3, (float)s_n);
    }
}

// The persistent render kernel.  kUniform = false: every lane runs the per-lane interpreter (`traverse<>`) for one whole ray
// segment per iteration and the warp re-converges for shading (HRT_FLAG_INTERPRETER: the plain form).  kUniform = true:
// the warp walks the stream together instead (traverse_uniform, hrt_device.cuh) — HRT_FLAG_UNIFORM, the default for
// small jobs and for scenes without OP_BVH trees.
template <bool kUniform>
__global__ void __launch_bounds__(kBlock, 2) render_interp_kernel(const __grid_constant__ RenderParams P) {
    __shared__ float sh_acc[kWarpsPerBlock][32][3];
    __shared__ NoiseTable sh_noise[kMaxNoiseTablesShared];
    TexEnv E;
    stage_noise(P.S, E, sh_noise, kMaxNoiseTablesShared);

    const DeviceScene& S = P.S;
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const unsigned lt_mask = (1u << lane) - 1u;
    const bool ref_boxes = P.reference_boxes != 0;
    const V3 bg = v3(P.bg[0], P.bg[1], P.bg[2]);
    const float div_w = (float)P.width - 1.0f, div_h = (float)P.height - 1.0f;  // application.rs:444-445
    unsigned long long n_rays = 0, n_paths = 0;

    for (;;) {
        unsigned long long item = 0;
        if (lane == 0) item = atomicAdd(P.counters, 1ULL);
        item = __shfl_sync(kFull, item, 0);
        if (item >= (unsigned long long)P.n_items) break;
        const int tile = (int)(item % (unsigned long long)P.n_tiles);
        const int chunk = (int)(item / (unsigned long long)P.n_tiles);
        const int tx = tile % P.tiles_x, ty = tile / P.tiles_x;
        const int tc = chunk - P.n_big;  // >= 0: one of the shrinking tail chunks
        const int s0 = P.sample_begin + (tc < 0 ? chunk * P.chunk : P.tail_begin[tc]);
        const int s_n = tc < 0 ? P.chunk : P.tail_size[tc];
        const int pool_size = 32 * s_n;
        int pool_next = 0;

        sh_acc[warp][lane][0] = 0.0f;
        sh_acc[warp][lane][1] = 0.0f;
        sh_acc[warp][lane][2] = 0.0f;
        __syncwarp();

        bool active = false;
        Ray ray;
        ray.o = v3(0.0f, 0.0f, 0.0f); ray.d = v3(1.0f, 1.0f, 1.0f); ray.time = 0.0f;
        V3 T = v3(1.0f, 1.0f, 1.0f);
        uint32_t bounce = 0;
        RngKey key;
        key.k0 = P.k0; key.k1 = P.k1; key.pixel = 0; key.sample = 0;
        int my_pl = lane;

        for (;;) {
            // ---- re-fill dead lanes from the warp-local pool (ballot + popc compaction) ----
            const unsigned need = __ballot_sync(kFull, !active);
            if (need) {
                const int idx = pool_next + __popc(need & lt_mask);
                pool_next += __popc(need);
                if (!active && idx < pool_size) {
                    const int pl = idx & 31;
                    const int px = tx * 8 + (pl & 7), py = ty * 4 + (pl >> 3);
                    if (px < P.width && py < P.height) {
                        key.pixel = (uint32_t)(py * P.width + px);
                        key.sample = (uint32_t)(s0 + (idx >> 5));
                        float c4[4];
                        rng_block(key, 0, RNG_BLOCK_CAMERA, c4);  // jitter x, jitter y, shutter time, lens u1
                        float lens_u2 = 0.0f;
                        if (P.cam.lens_radius != 0.0f) {
                            float l4[4];
                            rng_block(key, 0, RNG_BLOCK_CAMERA - 1, l4);
                            lens_u2 = l4[0];
                        }
                        const float u = ((float)px + c4[0]) / div_w;
                        const float v = ((float)py + c4[1]) / div_h;
                        ray = camera_get_ray(P.cam, u, v, c4[3], lens_u2, c4[2]);
                        T = v3(1.0f, 1.0f, 1.0f);
                        bounce = 0;
                        my_pl = pl;
                        active = P.depth > 0;  // ray_color(depth == 0) is black (application.rs:478-480)
                        n_paths++;
                    }
                }
            }
            if (!__any_sync(kFull, active)) {
                if (pool_next >= pool_size) break;
                continue;
            }
            // ---- one ray segment per live lane: world.hit + emitted + scatter (application.rs:477-495) ----
            MediumXi xi;
            xi.key = key; xi.bounce = bounce; xi.injected = 0.0f; xi.inject = false;
            Best best;
            best.pc = -1; best.t = 0.0f; best.face = 0; best.ctx = 0;
            float closest = CUDART_INF_F;
            bool hit = false;
            if (kUniform) hit = traverse_uniform(S, 0, S.n_ops, active, ray, ray, 0, 0.001f, closest, best, ref_boxes, xi);
            else if (active) hit = traverse<false>(S, 0, S.n_ops, ray, ray, 0, 0.001f, closest, best, ref_boxes, xi);
            if (active) {
                n_rays++;
                V3 add = v3(0.0f, 0.0f, 0.0f);
                if (!hit) {
                    add = T * bg;
                    active = false;
                } else {
                    HitRec h;
                    make_hit_record(S, ray, best, false, h);
                    const Material m = S.mats[h.mat];
                    if (m.kind == MAT_DIFFUSE_LIGHT) {
                        add = T * material_emitted(S, E, m, h);
                        active = false;  // DiffuseLight::scatter -> None
                    } else {
                        float u4[4];
                        rng_block(key, bounce, RNG_BLOCK_SCATTER, u4);
                        V3 att;
                        Ray sc;
                        if (material_scatter(S, E, m, ray, h, u4, att, sc)) {
                            T = T * att;
                            ray = sc;
                            bounce++;
                            if (bounce >= (uint32_t)P.depth) active = false;  // depth == 0 -> black
                        } else {
                            active = false;
                        }
                    }
                }
                if (add.x != 0.0f) atomicAdd(&sh_acc[warp][my_pl][0], add.x);
                if (add.y != 0.0f) atomicAdd(&sh_acc[warp][my_pl][1], add.y);
                if (add.z != 0.0f) atomicAdd(&sh_acc[warp][my_pl][2], add.z);
            }
        }
        __syncwarp();
        {
            const int px = tx * 8 + (lane & 7), py = ty * 4 + (lane >> 3);
            if (px < P.width && py < P.height) {
                float* dst = reinterpret_cast<float*>(P.accum + (size_t)py * P.width + px);
                atomicAdd(dst + 0, sh_acc[warp][lane][0]);
                atomicAdd(dst + 1, sh_acc[warp][lane][1]);
                atomicAdd(dst + 2, sh_acc[warp][lane][2]);
                atomicAdd(dst + 3, (float)s_n);
            }
        }
        __syncwarp();
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        n_rays += __shfl_xor_sync(kFull, n_rays, o);
        n_paths += __shfl_xor_sync(kFull, n_paths, o);
    }
    if (lane == 0) {
        atomicAdd(P.counters + 1, n_rays);
        atomicAdd(P.counters + 2, n_paths);
    }
}

// ---- wavefront render: two small kernels per ray segment instead of one big persistent kernel ----
// Why: the persistent kernels above carry the whole path tracer (camera, stream walk, tree walk, media, hit record,
// materials, textures: 4 000 - 5 000 SASS instructions = 64 - 80 KB), every warp cycles through all of it once per ray
// segment on its own schedule, and instruction fetch is then THE top stall: 5 - 11 warps per issue slot waiting in
// `no_instruction`, growing with the kernel's size (profiles/r02_*).  Here N path slots live in global memory and one
// iteration is
//   wave_logic_kernel   shade the segment traced last (emitted + scatter, application.rs:486-494), accumulate finished
//                       paths, start the next camera sample in every free slot (application.rs:443-448), test the next
//                       ray against the root box of every tree that is walked ahead and queue it
//   wave_noise_kernel   the NoiseTexture albedos the logic pass deferred, one evaluation per thread
//   wave_tree_kernel    the queued tree walks of all pre-walked trees, compacted over the whole wave
//   wave_trace_kernel   world.hit for every slot's ray segment: the warp-uniform walk of the WAVE form of the stream,
//                       which takes the trees' answers at their OP_BVH_PRE records (hrt_types.h)
// each a fraction of the code, each run by the whole GPU at once, so the instruction caches hold what is running.
// Same Philox streams and arithmetic as the other kernels: the same paths, summed in a different order.
constexpr int kWaveBlock = 256;
#ifndef HRT_TRACE_BLOCKS
#define HRT_TRACE_BLOCKS 3
#endif
#ifndef HRT_LOGIC_BLOCKS
#define HRT_LOGIC_BLOCKS 4  // 64 registers, 32 warps per SM: the kernel waits on memory (C5 499 -> 521 Mpaths/s against 3 blocks)
#endif
#ifndef HRT_LOGIC_THREADS
#define HRT_LOGIC_THREADS 256  // threads per block of wave_logic_kernel (its barriers couple the warps of a block)
#endif
constexpr int kLogicBlock = HRT_LOGIC_THREADS;
static_assert(kWaveBlock % kLogicBlock == 0 && kLogicBlock % 32 == 0, "logic block size");
#ifndef HRT_LOGIC_PREFETCH
#define HRT_LOGIC_PREFETCH 1  // diagnostic builds: 0 = plain loads of the slot state (which ptxas sinks into their branches)
#endif
enum WaveField {
    WF_OX, WF_OY, WF_OZ, WF_DX, WF_DY, WF_DZ, WF_TIME,  // the ray segment to trace / traced last
    WF_TX, WF_TY, WF_TZ, WF_PIXEL, WF_SAMPLE,           // throughput, Philox key
    WF_BOUNCE,                                          // >= 0: segments behind this path; -1: free slot
    WF_HIT_T, WF_HIT_PC, WF_HIT_FC,                     // result of the trace (pc -1: miss; FC = side | ctx << 8)
    WF_WORDS
};
struct WaveParams {
    DeviceScene S;
    CameraK cam;
    int width, height, depth;
    float bg[3];
    uint32_t k0, k1;
    int sample_begin;
    int n_pixels, n_slots;
    unsigned long long total_paths;
    int reference_boxes;
    float* st;                     // [WF_WORDS][n_slots]
    unsigned long long* counters;  // [0] next path index, [1] rays, [2] paths
    // the tree stage: the first n_pre OP_BVH trees of the stream are walked for all slots together, compacted
    int n_pre;
    PreTree pre[kMaxPreTrees];
    float4* tq;                    // [kMaxPreTrees][n_slots][2]: {o.xyz, time}, {d.xyz, slot} in the tree's ray space
    float2* pre_res;               // [n_slots][kMaxPreTrees]: {t, code} (traverse_uniform)
    float4* xq;                    // [n_slots] deferred NoiseTexture evaluations: {p.xyz, slot | texture << 22}
    int* tq_count;                 // [kMaxPreTrees] tree walks queued, [kMaxPreTrees] taken, [1] noise evaluations queued
    int tree_refill;               // idle lanes at which a tree-walk warp pulls new entries
    int tree_inner_min;            // lanes at an inner node for which the walk stays in its inner-node loop
    double* acc64;                 // [n_pixels][4] radiance sums + sample counts of this render (added into `accum` at the end)
    float4* accum;
    int* live_out;                 // when not null: += number of slots that carry a path after this logic pass
};
// One reduction per finished path.  In f64: the persistent kernels add a work item's (<= 256 samples') partial sum, here
// every sample is added on its own, and 10^4 f32 additions of near-equal terms drift by ~1e-4 relative (measured on the
// constant background of `earth`), enough to show in the pooled z-scores.
__device__ __forceinline__ void wave_accumulate(const WaveParams& P, uint32_t pixel, V3 add) {
    double* a = P.acc64 + 4 * (size_t)pixel;
    if (add.x != 0.0f) atomicAdd(a + 0, (double)add.x);
    if (add.y != 0.0f) atomicAdd(a + 1, (double)add.y);
    if (add.z != 0.0f) atomicAdd(a + 2, (double)add.z);
    atomicAdd(a + 3, 1.0);
}
__global__ void __launch_bounds__(256) wave_finish_kernel(double* __restrict__ acc64, int n_pixels, float4* __restrict__ accum) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_pixels) return;
    double* a = acc64 + 4 * (size_t)i;
    float4 v = accum[i];
    v.x += (float)a[0]; v.y += (float)a[1]; v.z += (float)a[2]; v.w += (float)a[3];
    accum[i] = v;
}
static_assert(WF_WORDS == hrt::kWaveStateWords, "wave state layout");
__device__ __forceinline__ float ld_volatile(const float* p) {
    float v;
    asm volatile("ld.volatile.global.f32 %0, [%1];" : "=f"(v) : "l"(p));
    return v;
}
#define WST(f, slot) P.st[(size_t)(f) * P.n_slots + (slot)]

__global__ void __launch_bounds__(kWaveBlock, HRT_TRACE_BLOCKS) wave_trace_kernel(const __grid_constant__ WaveParams P) {
    const int slot = blockIdx.x * kWaveBlock + threadIdx.x;
    const bool in_range = slot < P.n_slots;
    const int bounce = in_range ? __float_as_int(WST(WF_BOUNCE, slot)) : -1;
    const bool active = bounce >= 0;
    Ray ray;
    ray.o = v3(0.0f, 0.0f, 0.0f); ray.d = v3(1.0f, 1.0f, 1.0f); ray.time = 0.0f;
    MediumXi xi;
    xi.key.k0 = P.k0; xi.key.k1 = P.k1; xi.key.pixel = 0; xi.key.sample = 0;
    xi.bounce = 0; xi.injected = 0.0f; xi.inject = false;
    if (active) {
        ray.o = v3(WST(WF_OX, slot), WST(WF_OY, slot), WST(WF_OZ, slot));
        ray.d = v3(WST(WF_DX, slot), WST(WF_DY, slot), WST(WF_DZ, slot));
        ray.time = WST(WF_TIME, slot);
        xi.key.pixel = __float_as_uint(WST(WF_PIXEL, slot));
        xi.key.sample = __float_as_uint(WST(WF_SAMPLE, slot));
        xi.bounce = (uint32_t)bounce;
    }
    Best best;
    best.pc = -1; best.t = 0.0f; best.face = 0; best.ctx = 0;
    float closest = CUDART_INF_F;
    traverse_uniform<false>(P.S, 0, P.S.n_ops, active, ray, ray, 0, 0.001f, closest, best, P.reference_boxes != 0, xi,
                            P.pre_res + (size_t)(in_range ? slot : 0) * kMaxPreTrees, P.n_pre);
    if (blockIdx.x == 0 && threadIdx.x < hrt::kWaveCounters) P.tq_count[threadIdx.x] = 0;  // the queues are consumed
    if (active) {
        WST(WF_HIT_T, slot) = best.t;
        WST(WF_HIT_PC, slot) = __int_as_float(best.pc);
        WST(WF_HIT_FC, slot) = __int_as_float(best.face | (best.ctx << 8));
    }
    const unsigned m = __ballot_sync(kFull, active);
    if ((threadIdx.x & 31) == 0 && m) atomicAdd(P.counters + 1, (unsigned long long)__popc(m));
}

// Deferred NoiseTexture albedos (material_scatter): throughput *= value, one queued evaluation per thread.
__global__ void __launch_bounds__(kWaveBlock) wave_noise_kernel(const __grid_constant__ WaveParams P) {
    __shared__ NoiseTable sh_noise[kMaxNoiseTablesShared];
    TexEnv E;
    stage_noise(P.S, E, sh_noise, kMaxNoiseTablesShared);
    const int n = P.tq_count[2 * kMaxPreTrees];
    for (int i = blockIdx.x * kWaveBlock + threadIdx.x; i < n; i += gridDim.x * kWaveBlock) {
        const float4 q = P.xq[i];
        const int code = __float_as_int(q.w);
        const int slot = code & ((1 << 22) - 1);
        const V3 c = texture_value(P.S, E, code >> 22, 0.0f, 0.0f, v3(q.x, q.y, q.z));
        WST(WF_TX, slot) *= c.x; WST(WF_TY, slot) *= c.y; WST(WF_TZ, slot) *= c.z;
    }
}

// Tree stage, part 2: persistent warps pull walks from the trees' queues — a lane that finishes its walk takes the next
// entry — so the lanes stay filled however much the walks differ in length.  (Inside wave_trace_kernel the same walks
// were 64 % of its issued instructions at 3 - 7 of 32 lanes: only the few rays of a warp that reach one tree walk
// together, profiles/r02_w1_*.)  tree >= 0: that tree's queue only; tree < 0: the queues of all pre-walked trees one
// behind the other in ONE launch (one ramp-up and one drain per iteration instead of one per tree; the walks of different
// trees run the same code, only the first node differs).
__global__ void __launch_bounds__(kWaveBlock, 3) wave_tree_kernel(const __grid_constant__ WaveParams P, const int tree) {
    const DeviceScene& S = P.S;
    const int lane = threadIdx.x & 31;
    const unsigned lt_mask = (1u << lane) - 1u;
    const int first = tree < 0 ? 0 : tree, last = tree < 0 ? P.n_pre : tree + 1;
    int n_before[kMaxPreTrees + 1];  // entries of the queues in front of each tree's
    n_before[0] = 0;
#pragma unroll
    for (int t = 0; t < kMaxPreTrees; ++t) n_before[t + 1] = n_before[t] + ((t >= first && t < last) ? P.tq_count[t] : 0);
    const int n_entries = n_before[kMaxPreTrees];
    int* const next = P.tq_count + kMaxPreTrees + first;
    int slot = -1;  // < 0: this lane is idle
    int my_tree = first, base = P.pre[first].base;
    float ts = P.pre[first].ts, te = P.pre[first].te;
    int ref = 0, sp = 0;
    Ray cur;
    cur.o = v3(0.0f, 0.0f, 0.0f); cur.d = v3(1.0f, 1.0f, 1.0f); cur.time = 0.0f;
    RayK k = make_rayk(cur);
    TreeHit h;
    h.t = CUDART_INF_F; h.pc = -1; h.face = 0;
    int stack_ref[kBvh2Stack];
    float stack_t[kBvh2Stack];
    bool drained = false;  // warp-uniform: the queue has no entry left
    for (;;) {
        const unsigned idle = __ballot_sync(kFull, slot < 0);
        // refill when enough lanes are idle: one atomic on the queue head per refill, shared by every warp of the GPU, so
        // refilling single lanes makes that counter the bottleneck (measured: 836 us per tree and iteration)
        if (!drained && __popc(idle) >= P.tree_refill) {
            int e0 = 0;
            if (lane == 0) e0 = atomicAdd(next, __popc(idle));
            e0 = __shfl_sync(kFull, e0, 0);
            const int e = e0 + __popc(idle & lt_mask);
            if (e0 + __popc(idle) >= n_entries) drained = true;
            if (slot < 0 && e < n_entries) {
                my_tree = first;
                int before = 0;
#pragma unroll
                for (int t = 1; t < kMaxPreTrees; ++t)
                    if (t > first && t < last && e >= n_before[t]) { my_tree = t; before = n_before[t]; }
                const float4* const q = P.tq + 2 * ((size_t)my_tree * P.n_slots + (size_t)(e - before));
                const float4 a = __ldg(q), b = __ldg(q + 1);
                slot = __float_as_int(b.w);
                cur.o = v3(a.x, a.y, a.z); cur.d = v3(b.x, b.y, b.z); cur.time = a.w;
                k = make_rayk(cur);
                h.t = CUDART_INF_F; h.pc = -1; h.face = 0;
                ref = 0; sp = 0;
                base = P.pre[my_tree].base; ts = P.pre[my_tree].ts; te = P.pre[my_tree].te;
            }
        }
        if (!__any_sync(kFull, slot >= 0)) break;
        // "while-while": inner nodes for as long as most busy lanes stand at one (a lane that has reached a leaf waits),
        // then the leaves — a step that mixed both would run each with only the lanes that are at it
        for (;;) {
            const bool at_inner = slot >= 0 && ref >= 0;
            const int n_inner = __popc(__ballot_sync(kFull, at_inner));
            if (n_inner < P.tree_inner_min) break;
            if (at_inner && bvh2_inner(S, base, cur, k, 0.001f, ref, sp, stack_ref, stack_t, h)) {
                P.pre_res[(size_t)slot * kMaxPreTrees + my_tree] =
                    make_float2(h.t, __int_as_float(h.pc < 0 ? kPreNone : (h.pc | (h.face << 24))));
                slot = -1;
            }
        }
        if (slot >= 0) {
            if (bvh2_step(S, base, cur, k, 0.001f, ts, te, ref, sp, stack_ref, stack_t, h)) {
                P.pre_res[(size_t)slot * kMaxPreTrees + my_tree] =
                    make_float2(h.t, __int_as_float(h.pc < 0 ? kPreNone : (h.pc | (h.face << 24))));
                slot = -1;
            }
        }
    }
}

__global__ void __launch_bounds__(kLogicBlock, HRT_LOGIC_BLOCKS * (kWaveBlock / kLogicBlock)) wave_logic_kernel(const __grid_constant__ WaveParams P) {
    // NoiseTexture albedos are deferred to wave_noise_kernel, so no perlin table is staged here (9.7 KB per block and a
    // barrier before the first state load); a checker over noise or an emitting noise texture reads the tables in place
    TexEnv E;
    E.sh_noise = nullptr; E.n_shared_noise = 0;
    const DeviceScene& S = P.S;
    const int lane = threadIdx.x & 31;
    const int slot = blockIdx.x * kLogicBlock + threadIdx.x;
    const bool in_range = slot < P.n_slots;
    // the whole slot is fetched at once, not behind the test of its bounce word (a second trip to HBM on the critical
    // path of every block); free slots exist only while a render ramps up and drains
    // (ld.volatile: ptxas sinks plain loads back into the branches that use their values)
    const int ld = in_range ? slot : 0;
#if HRT_LOGIC_PREFETCH
#define WLD(f) ld_volatile(&WST(f, ld))
#else
#define WLD(f) WST(f, ld)
#endif
    int bounce = in_range ? __float_as_int(WLD(WF_BOUNCE)) : -1;
    RngKey key;
    key.k0 = P.k0; key.k1 = P.k1;
    key.pixel = __float_as_uint(WLD(WF_PIXEL));
    key.sample = __float_as_uint(WLD(WF_SAMPLE));
    Ray ray;
    ray.o = v3(WLD(WF_OX), WLD(WF_OY), WLD(WF_OZ));
    ray.d = v3(WLD(WF_DX), WLD(WF_DY), WLD(WF_DZ));
    ray.time = WLD(WF_TIME);
    V3 T = v3(WLD(WF_TX), WLD(WF_TY), WLD(WF_TZ));
    const int hit_pc = __float_as_int(WLD(WF_HIT_PC));
    const int hit_fc = __float_as_int(WLD(WF_HIT_FC));
    const float hit_t = WLD(WF_HIT_T);
#undef WLD
    Ray nxt;  // the segment this slot traces next (when it still carries a path after this pass)
    nxt.o = v3(0.0f, 0.0f, 0.0f); nxt.d = v3(1.0f, 1.0f, 1.0f); nxt.time = 0.0f;
    if (bounce >= 0) {
        // ---- emitted + scatter for the segment traced last (application.rs:482-494) ----
        V3 add = v3(0.0f, 0.0f, 0.0f);
        bool alive = false;
        if (hit_pc < 0) {
            add = T * v3(P.bg[0], P.bg[1], P.bg[2]);
        } else {
            const int fc = hit_fc;
            Best best;
            best.t = hit_t; best.pc = hit_pc; best.face = fc & 0xff; best.ctx = fc >> 8;
            HitRec h;
            make_hit_record(S, ray, best, false, h);
            const Material m = S.mats[h.mat];
            if (m.kind == MAT_DIFFUSE_LIGHT) {
                add = T * material_emitted(S, E, m, h);  // DiffuseLight::scatter -> None
            } else {
                float u4[4];
                rng_block(key, (uint32_t)bounce, RNG_BLOCK_SCATTER, u4);
                V3 att;
                Ray sc;
                int noise_tex = -1;
                if (material_scatter(S, E, m, ray, h, u4, att, sc, &noise_tex)) {
                    T = T * att;
                    bounce++;
                    if (bounce < P.depth) {  // ray_color(depth == 0) is black
                        alive = true;
                        if (noise_tex >= 0)  // the albedo is applied by wave_noise_kernel
                            P.xq[atomicAdd(P.tq_count + 2 * kMaxPreTrees, 1)] =
                                make_float4(h.p.x, h.p.y, h.p.z, __int_as_float(slot | (noise_tex << 22)));
                        nxt = sc;
                        WST(WF_OX, slot) = sc.o.x; WST(WF_OY, slot) = sc.o.y; WST(WF_OZ, slot) = sc.o.z;
                        WST(WF_DX, slot) = sc.d.x; WST(WF_DY, slot) = sc.d.y; WST(WF_DZ, slot) = sc.d.z;
                        WST(WF_TIME, slot) = sc.time;
                        WST(WF_TX, slot) = T.x; WST(WF_TY, slot) = T.y; WST(WF_TZ, slot) = T.z;
                        WST(WF_BOUNCE, slot) = __int_as_float(bounce);
                    }
                }
            }
        }
        if (!alive) {
            // the path is over: its radiance and its sample count go to the pixel (one 16-byte reduction)
            wave_accumulate(P, key.pixel, add);
            bounce = -1;
        }
    }
    // ---- free slots start the next camera samples (application.rs:443-448), drawn from one global index ----
    // ONE atomic per block: every warp of the GPU draws from the same counter, and a returning atomic per warp on one
    // address was 38 % of this kernel's stall samples (profiles/r02_wave_logic_kernel_by_function.txt)
    __shared__ int sh_cnt[kLogicBlock / 32];
    __shared__ unsigned long long sh_base64;
    const int warp = threadIdx.x >> 5;
    const bool is_free = in_range && bounce < 0;
    const unsigned need = __ballot_sync(kFull, is_free);
    if (lane == 0) sh_cnt[warp] = __popc(need);
    __syncthreads();
    if (threadIdx.x == 0) {
        int tot = 0;
#pragma unroll
        for (int ww = 0; ww < kLogicBlock / 32; ++ww) tot += sh_cnt[ww];
        sh_base64 = tot ? atomicAdd(P.counters, (unsigned long long)tot) : 0ULL;
    }
    __syncthreads();
    bool started = false;
    if (need) {
        unsigned long long base = sh_base64;
        for (int ww = 0; ww < warp; ++ww) base += (unsigned long long)sh_cnt[ww];
        const unsigned long long idx = base + (unsigned long long)__popc(need & ((1u << lane) - 1u));
        if (is_free && idx < P.total_paths) {
            key.pixel = (uint32_t)(idx % (unsigned long long)P.n_pixels);
            key.sample = (uint32_t)P.sample_begin + (uint32_t)(idx / (unsigned long long)P.n_pixels);
            if (P.depth > 0) {
                const int px = (int)(key.pixel % (uint32_t)P.width), py = (int)(key.pixel / (uint32_t)P.width);
                float c4[4];
                rng_block(key, 0, RNG_BLOCK_CAMERA, c4);  // jitter x, jitter y, shutter time, lens u1
                float lens_u2 = 0.0f;
                if (P.cam.lens_radius != 0.0f) {
                    float l4[4];
                    rng_block(key, 0, RNG_BLOCK_CAMERA - 1, l4);
                    lens_u2 = l4[0];
                }
                const float u = ((float)px + c4[0]) / ((float)P.width - 1.0f);   // application.rs:444-445
                const float v = ((float)py + c4[1]) / ((float)P.height - 1.0f);
                nxt = camera_get_ray(P.cam, u, v, c4[3], lens_u2, c4[2]);
                WST(WF_OX, slot) = nxt.o.x; WST(WF_OY, slot) = nxt.o.y; WST(WF_OZ, slot) = nxt.o.z;
                WST(WF_DX, slot) = nxt.d.x; WST(WF_DY, slot) = nxt.d.y; WST(WF_DZ, slot) = nxt.d.z;
                WST(WF_TIME, slot) = nxt.time;
                WST(WF_TX, slot) = 1.0f; WST(WF_TY, slot) = 1.0f; WST(WF_TZ, slot) = 1.0f;
                WST(WF_PIXEL, slot) = __uint_as_float(key.pixel);
                WST(WF_SAMPLE, slot) = __uint_as_float(key.sample);
                WST(WF_BOUNCE, slot) = __int_as_float(0);
                bounce = 0;
            } else {
                wave_accumulate(P, key.pixel, v3(0.0f, 0.0f, 0.0f));  // depth 0: black, but counted
            }
            started = true;
        }
    }
    if (in_range && bounce < 0) WST(WF_BOUNCE, slot) = __int_as_float(-1);
    // ---- tree stage, part 1: the ray this slot traces next against the root box of every pre-walked tree ----
    // (all trees in ONE round of block-wide counting: one barrier pair and one atomic per tree and block)
    if (P.n_pre > 0) {
        __shared__ int sh_tcnt[kMaxPreTrees][kLogicBlock / 32];
        __shared__ int sh_tbase[kMaxPreTrees];
        const bool live = bounce >= 0;
        Ray rr[kMaxPreTrees];
        unsigned mm[kMaxPreTrees];
        bool hit[kMaxPreTrees], nan_ray[kMaxPreTrees];
#pragma unroll
        for (int tree = 0; tree < kMaxPreTrees; ++tree) {
            rr[tree] = nxt; mm[tree] = 0u; hit[tree] = false; nan_ray[tree] = false;
            if (tree < P.n_pre) {
                const Ray r = P.pre[tree].ctx != 0 ? ray_in_ctx(S, nxt, P.pre[tree].ctx) : nxt;
                const RayK k = make_rayk(r);
                const float4 mn = make_float4(P.pre[tree].mn[0], P.pre[tree].mn[1], P.pre[tree].mn[2], 0.0f);
                const float4 mx = make_float4(P.pre[tree].mx[0], P.pre[tree].mx[1], P.pre[tree].mx[2], 0.0f);
                // A ray whose origin (or direction) is all NaN — the continuation of a path that hit something at t = NaN,
                // Q15 — passes every box test and is accepted by every primitive test, each accepted hit replacing the one
                // before (the reference's `t <= t_max` is never false): it visits the whole tree (~2000 steps against ~14
                // for an ordinary ray, and a wave of a million rays nearly always holds one) to end at the last leaf of the
                // reference's order with t = NaN.  That answer is written directly.
                nan_ray[tree] = (r.o.x != r.o.x && r.o.y != r.o.y && r.o.z != r.o.z) || (r.d.x != r.d.x && r.d.y != r.d.y && r.d.z != r.d.z);
                hit[tree] = live && !nan_ray[tree] && box_hit_tight(mn, mx, r, k, 0.001f, CUDART_INF_F);
                mm[tree] = __ballot_sync(kFull, hit[tree]);
                if (lane == 0) sh_tcnt[tree][warp] = __popc(mm[tree]);
                rr[tree] = r;
            }
        }
        __syncthreads();
        if (threadIdx.x < P.n_pre) {
            int tot = 0;
#pragma unroll
            for (int ww = 0; ww < kLogicBlock / 32; ++ww) tot += sh_tcnt[threadIdx.x][ww];
            sh_tbase[threadIdx.x] = tot ? atomicAdd(P.tq_count + threadIdx.x, tot) : 0;
        }
        __syncthreads();
#pragma unroll
        for (int tree = 0; tree < kMaxPreTrees; ++tree) {
            if (tree < P.n_pre) {
                if (hit[tree]) {
                    int at = sh_tbase[tree] + __popc(mm[tree] & ((1u << lane) - 1u));
                    for (int ww = 0; ww < warp; ++ww) at += sh_tcnt[tree][ww];
                    float4* q = P.tq + 2 * ((size_t)tree * P.n_slots + at);
                    const Ray& r = rr[tree];
                    q[0] = make_float4(r.o.x, r.o.y, r.o.z, r.time);
                    q[1] = make_float4(r.d.x, r.d.y, r.d.z, __int_as_float(slot));
                }
                if (live)
                    P.pre_res[(size_t)slot * kMaxPreTrees + tree] =
                        nan_ray[tree] ? make_float2(CUDART_NAN_F, __int_as_float(P.pre[tree].last_pc | (P.pre[tree].last_face << 24)))
                                      : make_float2(0.0f, __int_as_float(kPreNone));
            }
        }
    }
    const unsigned ms = __ballot_sync(kFull, started);
    if (lane == 0 && ms) atomicAdd(P.counters + 2, (unsigned long long)__popc(ms));
    if (P.live_out) {
        // depth 0 never keeps a path, but unstarted samples mean the job is not over
        const unsigned ml = __ballot_sync(kFull, bounce >= 0 || (started && P.depth <= 0));
        if (lane == 0 && ml) atomicAdd(P.live_out, __popc(ml));
    }
}
#undef WST

__global__ void __launch_bounds__(256) resolve_kernel(const float4* __restrict__ accum, int n_pixels, float scale,
                                                      float4* __restrict__ out) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_pixels) return;
    float4 a = accum[i];
    out[i] = make_float4(sqrtf(a.x * scale), sqrtf(a.y * scale), sqrtf(a.z * scale), 1.0f);
}

__global__ void __launch_bounds__(128) trace_hits_kernel(const __grid_constant__ DeviceScene S, const hrt_ray* __restrict__ rays,
                                                         int n, const float* __restrict__ xi_in, hrt_hit* __restrict__ out,
                                                         int reference_boxes) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    hrt_ray r = rays[i];
    Ray ray;
    ray.o = v3(r.o[0], r.o[1], r.o[2]);
    ray.d = v3(r.d[0], r.d[1], r.d[2]);
    ray.time = r.time;
    MediumXi xi;
    xi.key.k0 = 0; xi.key.k1 = 0; xi.key.pixel = 0; xi.key.sample = 0;
    xi.bounce = 0;
    xi.inject = true;
    xi.injected = xi_in ? xi_in[i] : 0.5f;
    Best best;
    best.pc = -1; best.t = 0.0f; best.face = 0; best.ctx = 0;
    float closest = r.tmax;
    bool hit = traverse<false>(S, 0, S.n_ops, ray, ray, 0, r.tmin, closest, best, reference_boxes != 0, xi);
    hrt_hit o;
    o.hit = 0; o.t = 0.0f;
    o.p[0] = o.p[1] = o.p[2] = 0.0f;
    o.n[0] = o.n[1] = o.n[2] = 0.0f;
    o.u = 0.0f; o.v = 0.0f;
    o.front_face = 0; o.material_id = -1; o.prim_id = -1; o.face = 0;
    if (hit) {
        HitRec h;
        make_hit_record(S, ray, best, true, h);
        o.hit = 1;
        o.t = h.t;
        o.p[0] = h.p.x; o.p[1] = h.p.y; o.p[2] = h.p.z;
        o.n[0] = h.n.x; o.n[1] = h.n.y; o.n[2] = h.n.z;
        o.u = h.u; o.v = h.v;
        o.front_face = h.front_face ? 1 : 0;
        o.material_id = h.mat;
        o.prim_id = h.prim;
        o.face = h.face;
    }
    out[i] = o;
}

// world.hit() through the warp-uniform walk (traverse_uniform<>): one ray per lane, whole warps.
__global__ void __launch_bounds__(128) trace_hits_uniform_kernel(const __grid_constant__ DeviceScene S, const hrt_ray* __restrict__ rays,
                                                                 int n, const float* __restrict__ xi_in, hrt_hit* __restrict__ out,
                                                                 int reference_boxes) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const bool active = i < n;
    Ray ray;
    ray.o = v3(0.0f, 0.0f, 0.0f); ray.d = v3(1.0f, 1.0f, 1.0f); ray.time = 0.0f;
    float tmin = 0.0f, closest = 0.0f;
    MediumXi xi;
    xi.key.k0 = 0; xi.key.k1 = 0; xi.key.pixel = 0; xi.key.sample = 0;
    xi.bounce = 0;
    xi.inject = true;
    xi.injected = 0.5f;
    if (active) {
        const hrt_ray r = rays[i];
        ray.o = v3(r.o[0], r.o[1], r.o[2]);
        ray.d = v3(r.d[0], r.d[1], r.d[2]);
        ray.time = r.time;
        tmin = r.tmin;
        closest = r.tmax;
        if (xi_in) xi.injected = xi_in[i];
    }
    Best best;
    best.pc = -1; best.t = 0.0f; best.face = 0; best.ctx = 0;
    const bool hit = traverse_uniform(S, 0, S.n_ops, active, ray, ray, 0, tmin, closest, best, reference_boxes != 0, xi);
    if (!active) return;
    hrt_hit o;
    o.hit = 0; o.t = 0.0f;
    o.p[0] = o.p[1] = o.p[2] = 0.0f;
    o.n[0] = o.n[1] = o.n[2] = 0.0f;
    o.u = 0.0f; o.v = 0.0f;
    o.front_face = 0; o.material_id = -1; o.prim_id = -1; o.face = 0;
    if (hit) {
        HitRec h;
        make_hit_record(S, ray, best, true, h);
        o.hit = 1;
        o.t = h.t;
        o.p[0] = h.p.x; o.p[1] = h.p.y; o.p[2] = h.p.z;
        o.n[0] = h.n.x; o.n[1] = h.n.y; o.n[2] = h.n.z;
        o.u = h.u; o.v = h.v;
        o.front_face = h.front_face ? 1 : 0;
        o.material_id = h.mat;
        o.prim_id = h.prim;
        o.face = h.face;
    }
    out[i] = o;
}

__global__ void __launch_bounds__(kBlock) tex_value_kernel(const __grid_constant__ DeviceScene S, int tex,
                                                           const float* __restrict__ uvp, int n, float* __restrict__ out) {
    __shared__ NoiseTable sh_noise[kMaxNoiseTablesShared];
    TexEnv E;
    stage_noise(S, E, sh_noise, kMaxNoiseTablesShared);
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float* q = uvp + 5 * (size_t)i;
    V3 c = texture_value(S, E, tex, q[0], q[1], v3(q[2], q[3], q[4]));
    out[3 * (size_t)i + 0] = c.x;
    out[3 * (size_t)i + 1] = c.y;
    out[3 * (size_t)i + 2] = c.z;
}

__global__ void __launch_bounds__(kBlock) scatter_kernel(const __grid_constant__ DeviceScene S, const hrt_ray* __restrict__ rays,
                                                         const hrt_hit* __restrict__ hits, const float* __restrict__ u4, int n,
                                                         hrt_scatter_out* __restrict__ out) {
    __shared__ NoiseTable sh_noise[kMaxNoiseTablesShared];
    TexEnv E;
    stage_noise(S, E, sh_noise, kMaxNoiseTablesShared);
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    hrt_scatter_out o;
    o.scattered = 0;
    for (int c = 0; c < 3; ++c) { o.attenuation[c] = 0.0f; o.o[c] = 0.0f; o.d[c] = 0.0f; o.emitted[c] = 0.0f; }
    o.time = 0.0f;
    const hrt_hit hh = hits[i];
    if (hh.hit && hh.material_id >= 0) {
        const hrt_ray r = rays[i];
        Ray ray;
        ray.o = v3(r.o[0], r.o[1], r.o[2]);
        ray.d = v3(r.d[0], r.d[1], r.d[2]);
        ray.time = r.time;
        HitRec h;
        h.p = v3(hh.p[0], hh.p[1], hh.p[2]);
        h.n = v3(hh.n[0], hh.n[1], hh.n[2]);
        h.t = hh.t; h.u = hh.u; h.v = hh.v;
        h.front_face = hh.front_face != 0;
        h.mat = hh.material_id; h.prim = hh.prim_id; h.face = hh.face;
        const Material m = S.mats[h.mat];
        V3 e = material_emitted(S, E, m, h);
        o.emitted[0] = e.x; o.emitted[1] = e.y; o.emitted[2] = e.z;
        float u[4] = {u4[4 * (size_t)i], u4[4 * (size_t)i + 1], u4[4 * (size_t)i + 2], u4[4 * (size_t)i + 3]};
        V3 att;
        Ray sc;
        if (material_scatter(S, E, m, ray, h, u, att, sc)) {
            o.scattered = 1;
            o.attenuation[0] = att.x; o.attenuation[1] = att.y; o.attenuation[2] = att.z;
            o.o[0] = sc.o.x; o.o[1] = sc.o.y; o.o[2] = sc.o.z;
            o.d[0] = sc.d.x; o.d[1] = sc.d.y; o.d[2] = sc.d.z;
            o.time = sc.time;
        }
    }
    out[i] = o;
}

__global__ void __launch_bounds__(kBlock) camera_rays_kernel(const __grid_constant__ CameraK cam, const float* __restrict__ stuuu,
                                                             int n, hrt_ray* __restrict__ out) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float* q = stuuu + 5 * (size_t)i;
    Ray r = camera_get_ray(cam, q[0], q[1], q[2], q[3], q[4]);
    hrt_ray o;
    o.o[0] = r.o.x; o.o[1] = r.o.y; o.o[2] = r.o.z;
    o.d[0] = r.d.x; o.d[1] = r.d.y; o.d[2] = r.d.z;
    o.time = r.time;
    o.tmin = 0.001f;
    o.tmax = CUDART_INF_F;
    out[i] = o;
}

#if !HRT_EXACT
// Fused reduce + resolve for single-process multi-GPU renders (hrt_render_multi): every device rendered a disjoint sample
// slice into its own accumulator; this kernel runs on the first device, loads the OTHER devices' accumulators directly
// over NVLink peer mappings (plain ld.global on peer pointers), sums them in a fixed order and applies the reference's
// gamma resolve (application.rs:451-456) in the same pass — no staging copies, no separate all-reduce.
struct PeerAccums {
    const float4* p[8];
    int n;
};
__global__ void __launch_bounds__(256) reduce_resolve_kernel(const PeerAccums A, int n_pixels, float scale, float4* __restrict__ out,
                                                             float4* __restrict__ sum_out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_pixels) return;
    float4 a = A.p[0][i];
    for (int k = 1; k < A.n; ++k) {
        const float4 b = A.p[k][i];  // peer memory over NVLink
        a.x += b.x; a.y += b.y; a.z += b.z; a.w += b.w;
    }
    if (sum_out) sum_out[i] = a;
    if (out) out[i] = make_float4(sqrtf(a.x * scale), sqrtf(a.y * scale), sqrtf(a.z * scale), 1.0f);
}
cudaError_t launch_reduce_resolve(const float* const* d_accums, int n, int n_pixels, int samples, float* d_out, float* d_sum_out,
                                  cudaStream_t stream) {
    PeerAccums A;
    A.n = n;
    for (int k = 0; k < 8; ++k) A.p[k] = reinterpret_cast<const float4*>(k < n ? d_accums[k] : d_accums[0]);
    const float scale = 1.0f / (float)samples;
    reduce_resolve_kernel<<<(n_pixels + 255) / 256, 256, 0, stream>>>(A, n_pixels, scale, reinterpret_cast<float4*>(d_out),
                                                                      reinterpret_cast<float4*>(d_sum_out));
    return cudaGetLastError();
}

// Roofline microbenchmarks (hrt_measure_peaks): 8 independent FFMA chains per thread; L2-resident float4 reads.
__global__ void __launch_bounds__(256) fma_peak_kernel(float* __restrict__ sink, int iters) {
    float a0 = threadIdx.x * 1e-3f, a1 = a0 + 1.0f, a2 = a0 + 2.0f, a3 = a0 + 3.0f, a4 = a0 + 4.0f, a5 = a0 + 5.0f,
          a6 = a0 + 6.0f, a7 = a0 + 7.0f;
    const float m = 0.999f + blockIdx.x * 1e-9f, c = 1e-3f;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            a0 = fmaf(a0, m, c); a1 = fmaf(a1, m, c); a2 = fmaf(a2, m, c); a3 = fmaf(a3, m, c);
            a4 = fmaf(a4, m, c); a5 = fmaf(a5, m, c); a6 = fmaf(a6, m, c); a7 = fmaf(a7, m, c);
        }
    }
    float r = ((a0 + a1) + (a2 + a3)) + ((a4 + a5) + (a6 + a7));
    if (r == 123.456f) sink[0] = r;
}
__global__ void __launch_bounds__(256) l2_read_kernel(const float4* __restrict__ buf, size_t n_vec, int repeats,
                                                      float* __restrict__ sink) {
    float acc = 0.0f;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (int r = 0; r < repeats; ++r)
        for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_vec; i += stride) {
            float4 v = __ldcg(buf + i);
            acc += (v.x + v.y) + (v.z + v.w);
        }
    if (acc == 123.456f) sink[0] = acc;
}
cudaError_t launch_fma_peak(float* d_sink, int grid, int iters, cudaStream_t stream) {
    fma_peak_kernel<<<grid, 256, 0, stream>>>(d_sink, iters);
    return cudaGetLastError();
}
cudaError_t launch_l2_read(const float4* d_buf, size_t n_vec, int repeats, float* d_sink, int grid, cudaStream_t stream) {
    l2_read_kernel<<<grid, 256, 0, stream>>>(d_buf, n_vec, repeats, d_sink);
    return cudaGetLastError();
}
#endif

// ------------------------------------------------------------------------------------------------
// Launchers
// ------------------------------------------------------------------------------------------------
static DeviceScene to_device_scene(const hrt::DeviceSceneHost& h) {
    DeviceScene S;
    S.ops = reinterpret_cast<const float4*>(h.ops);
    S.nodes = reinterpret_cast<const uint4*>(h.nodes);
    S.ctxs = reinterpret_cast<const Ctx*>(h.ctxs);
    S.mats = reinterpret_cast<const Material*>(h.mats);
    S.texs = reinterpret_cast<const Texture*>(h.texs);
    S.noise = reinterpret_cast<const NoiseTable*>(h.noise);
    for (int i = 0; i < kMaxImages; ++i) S.images[i] = h.images[i];
    S.n_ops = h.n_ops; S.n_noise = h.n_noise; S.n_media = h.n_media;
    S.ln_e = h.ln_e;

    return S;
}
static CameraK to_camera(const hrt_camera_state& c) {
    CameraK k;
    k.origin = V3{c.origin[0], c.origin[1], c.origin[2]};
    k.lower_left_corner = V3{c.lower_left_corner[0], c.lower_left_corner[1], c.lower_left_corner[2]};
    k.horizontal = V3{c.horizontal[0], c.horizontal[1], c.horizontal[2]};
    k.vertical = V3{c.vertical[0], c.vertical[1], c.vertical[2]};
    k.u = V3{c.u[0], c.u[1], c.u[2]};
    k.v = V3{c.v[0], c.v[1], c.v[2]};
    k.lens_radius = c.lens_radius;
    k.time0 = c.time0;
    k.time1 = c.time1;
    return k;
}

// Host loop of the wavefront render.  The path slots are split into partitions (2; HRT_WAVE_PARTS), each iterating on its own
// stream: every iteration ends with the slowest walk of its wave (a ray with a NaN component passes every box test and
// visits a whole 2000-node tree: ~0.8 ms against ~0.1 ms for the rest of the wave, and a million rays nearly always hold
// one), and while one partition sits in such a tail the others keep the SMs busy.  All partitions draw camera samples
// from the same global index, so a partition is finished for good when it ends a batch of iterations with no path left:
// batches are enqueued ahead and the count comes back through pinned memory.  Blocks the calling thread until the
// render is complete; on return the work is ordered before anything enqueued on `stream` afterwards.
cudaError_t launch_render_wave(hrt::RenderLaunch& L, hrt::WaveBuffers& W, int num_sms, cudaStream_t stream) {
    WaveParams P;
    P.S = to_device_scene(L.scene);
    P.cam = to_camera(L.cam);
    P.width = L.width; P.height = L.height; P.depth = L.depth;
    P.bg[0] = L.background[0]; P.bg[1] = L.background[1]; P.bg[2] = L.background[2];
    P.k0 = L.key0; P.k1 = L.key1;
    P.sample_begin = L.sample_begin;
    P.n_pixels = L.width * L.height;
    P.total_paths = (unsigned long long)P.n_pixels * (unsigned long long)L.sample_count;
    P.reference_boxes = L.reference_boxes;
    P.counters = L.counters;
    P.accum = reinterpret_cast<float4*>(L.accum);
    P.acc64 = W.acc64;
    P.n_pre = L.n_pre;
    for (int i = 0; i < kMaxPreTrees; ++i) P.pre[i] = L.pre[i];
    const bool one_tree_launch = !getenv("HRT_TREE_SPLIT");  // diagnostic: one wave_tree_kernel launch per tree
    P.tree_refill = 12;
    if (const char* env = getenv("HRT_TREE_REFILL")) P.tree_refill = std::max(1, std::min(32, atoi(env)));
    P.tree_inner_min = 12;
    if (const char* env = getenv("HRT_TREE_INNER")) P.tree_inner_min = std::max(1, std::min(33, atoi(env)));
    P.live_out = nullptr;
    // Two partitions: enough to fill one's end-of-iteration tails with the other's work, and the fewer the partitions the
    // larger (and the fewer) the launches (C5 at 2048 spp: 1 / 2 / 3 / 4 / 8 partitions of 4 Mi slots in all = 443 / 447 /
    // 436 / 430 / 410 Mpaths/s).  Slots in flight: every iteration costs a pass over ALL slots, also while a render ramps up
    // and drains (~50 iterations whatever its size), so a mid-size job gets fewer of them (C5 at 64 spp = 41 M paths: 404
    // Mpaths/s with 4 Mi slots, 337 with 8 Mi; at 2048 spp: 521 against 529).
    int parts = 2;
    if (const char* env = getenv("HRT_WAVE_PARTS")) parts = std::max(1, std::min(hrt::kWaveParts, atoi(env)));
    long long budget = W.n_slots;
    if (!getenv("HRT_WAVE_SLOTS")) budget = std::min<long long>(budget, std::max<long long>(4ll << 20, (long long)(P.total_paths / 128)));
    // the buffers are split evenly among the partitions in use; slot ids share a word with a texture id (wave_noise_kernel)
    const int cap = (int)std::min<long long>(budget / parts, 1 << 22) / kWaveBlock * kWaveBlock;
    const unsigned long long share = (P.total_paths + parts - 1) / parts;
    const unsigned long long want = (share + kWaveBlock - 1) / kWaveBlock * kWaveBlock;
    P.n_slots = (int)std::min<unsigned long long>((unsigned long long)cap, std::max<unsigned long long>(want, kWaveBlock));
    const int grid = (P.n_slots + kWaveBlock - 1) / kWaveBlock;
    L.grid = grid;
    L.block = kWaveBlock;
    L.launches = 0;
    cudaError_t e;
    if ((e = cudaMemsetAsync(W.acc64, 0, sizeof(double) * 4 * (size_t)P.n_pixels, stream)) != cudaSuccess) return e;
    if ((e = cudaMemsetAsync(W.tq_count, 0, sizeof(int) * hrt::kWaveCounters * hrt::kWaveParts, stream)) != cudaSuccess) return e;
    // every slot starts free
    if ((e = cudaMemsetAsync(W.state, 0xff, sizeof(float) * (size_t)WF_WORDS * (size_t)cap * parts, stream)) != cudaSuccess) return e;
    if ((e = cudaEventRecord(W.ev_begin, stream)) != cudaSuccess) return e;
    // persistent tree-walk warps: as many blocks as are resident at once
    int tree_blocks_per_sm = 1;
    if ((e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&tree_blocks_per_sm, wave_tree_kernel, kWaveBlock, 0)) != cudaSuccess) return e;
    if (const char* env = getenv("HRT_TREE_BLOCKS")) tree_blocks_per_sm = std::max(1, std::min(tree_blocks_per_sm, atoi(env)));
    const int tree_grid = std::min(grid, num_sms * std::max(1, tree_blocks_per_sm));
    // deferred noise evaluations: a few hundred thousand per iteration, each a long dependent chain (7 octaves of table
    // lookups) — several blocks per SM to hide it (HRT_NOISE_BLOCKS: diagnostic)
    int noise_blocks_per_sm = 4;
    if (const char* env = getenv("HRT_NOISE_BLOCKS")) noise_blocks_per_sm = std::max(1, std::min(8, atoi(env)));
    const int noise_grid = std::min(grid, num_sms * noise_blocks_per_sm);
    WaveParams PP[hrt::kWaveParts];
    bool running[hrt::kWaveParts];
    for (int p = 0; p < parts; ++p) {
        PP[p] = P;
        PP[p].st = W.state + (size_t)p * WF_WORDS * cap;
        PP[p].tq = reinterpret_cast<float4*>(W.tq) + (size_t)p * 2 * kMaxPreTrees * cap;
        PP[p].pre_res = reinterpret_cast<float2*>(W.pre) + (size_t)p * kMaxPreTrees * cap;
        PP[p].tq_count = W.tq_count + p * hrt::kWaveCounters;
        PP[p].xq = reinterpret_cast<float4*>(W.xq) + (size_t)p * cap;
        running[p] = true;
        if ((e = cudaStreamWaitEvent(W.streams[p], W.ev_begin, 0)) != cudaSuccess) return e;
    }
    const int kBatch = 16;
    int n_running = parts;
    for (int b = 0; n_running > 0; ++b) {
        for (int p = 0; p < parts; ++p) {
            if (!running[p]) continue;
            cudaStream_t sp = W.streams[p];
            int* live = W.d_live + 2 * p + (b & 1);
            if ((e = cudaMemsetAsync(live, 0, sizeof(int), sp)) != cudaSuccess) return e;
            for (int i = 0; i < kBatch; ++i) {
                PP[p].live_out = i == kBatch - 1 ? live : nullptr;
                wave_logic_kernel<<<grid * (kWaveBlock / kLogicBlock), kLogicBlock, 0, sp>>>(PP[p]);
                if (P.S.n_noise > 0) wave_noise_kernel<<<noise_grid, kWaveBlock, 0, sp>>>(PP[p]);
                if (one_tree_launch && P.n_pre > 1) {
                    wave_tree_kernel<<<tree_grid, kWaveBlock, 0, sp>>>(PP[p], -1);
                    L.launches += 1;
                } else {
                    for (int t = 0; t < P.n_pre; ++t) wave_tree_kernel<<<tree_grid, kWaveBlock, 0, sp>>>(PP[p], t);
                    L.launches += P.n_pre;
                }
                wave_trace_kernel<<<grid, kWaveBlock, 0, sp>>>(PP[p]);
                L.launches += 2 + (P.S.n_noise > 0 ? 1 : 0);
            }
            if ((e = cudaGetLastError()) != cudaSuccess) return e;
            if ((e = cudaMemcpyAsync(W.h_live + 2 * p + (b & 1), live, sizeof(int), cudaMemcpyDeviceToHost, sp)) != cudaSuccess) return e;
            if ((e = cudaEventRecord(W.ev[2 * p + (b & 1)], sp)) != cudaSuccess) return e;
        }
        if (b >= 1) {
            for (int p = 0; p < parts; ++p) {
                if (!running[p]) continue;
                if ((e = cudaEventSynchronize(W.ev[2 * p + ((b - 1) & 1)])) != cudaSuccess) return e;
                if (W.h_live[2 * p + ((b - 1) & 1)] == 0) {  // (the batch enqueued meanwhile finds nothing to do)
                    running[p] = false;
                    n_running--;
                }
            }
        }
    }
    for (int p = 0; p < parts; ++p) {
        if ((e = cudaEventRecord(W.ev_end[p], W.streams[p])) != cudaSuccess) return e;
        if ((e = cudaStreamWaitEvent(stream, W.ev_end[p], 0)) != cudaSuccess) return e;
    }
    wave_finish_kernel<<<(P.n_pixels + 255) / 256, 256, 0, stream>>>(W.acc64, P.n_pixels, P.accum);
    L.launches += 1;
    return cudaGetLastError();
}

cudaError_t launch_render(hrt::RenderLaunch& L, int num_sms, cudaStream_t stream) {
    RenderParams P;
    P.S = to_device_scene(L.scene);
    P.cam = to_camera(L.cam);
    P.width = L.width; P.height = L.height; P.depth = L.depth;
    P.bg[0] = L.background[0]; P.bg[1] = L.background[1]; P.bg[2] = L.background[2];
    P.k0 = L.key0; P.k1 = L.key1;
    P.sample_begin = L.sample_begin; P.sample_count = L.sample_count;
    P.tiles_x = (L.width + 7) / 8;
    P.tiles_y = (L.height + 3) / 4;
    P.n_tiles = P.tiles_x * P.tiles_y;
    int blocks_per_sm = 0;
    const bool uniform = L.interpreter != 1;
    cudaError_t e = uniform ? cudaOccupancyMaxActiveBlocksPerMultiprocessor(&blocks_per_sm, render_interp_kernel<true>, kBlock, 0)
                            : cudaOccupancyMaxActiveBlocksPerMultiprocessor(&blocks_per_sm, render_interp_kernel<false>, kBlock, 0);
    if (e != cudaSuccess) return e;
    if (blocks_per_sm < 1) blocks_per_sm = 1;
    const int grid = num_sms * blocks_per_sm;
    // Samples per work item: large enough to amortise the per-item flush and the end-of-item tail, small
    // enough that the dynamic cursor balances the last wave (>= ~8 items per resident warp when possible).
    int chunk = L.chunk;
    if (chunk <= 0) {
        chunk = 64;
        const long long resident_warps = (long long)grid * kWarpsPerBlock;
        while (chunk > 4 && (long long)P.n_tiles * ((L.sample_count + chunk - 1) / chunk) < 8 * resident_warps) chunk /= 2;
    }
    if (chunk > L.sample_count) chunk = L.sample_count;
    if (chunk < 1) chunk = 1;
    P.chunk = chunk;
    {
        // the last ~2 chunks' worth of samples is handed out in shrinking pieces: each a third of what is left (>= 16)
        int remaining = L.sample_count;
        P.n_big = 0;
        while (remaining >= 3 * chunk) { remaining -= chunk; P.n_big++; }
        int n_tail = 0, begin = P.n_big * chunk, piece = chunk;
        while (remaining > 0) {
            if (n_tail == kMaxTailChunks - 1) piece = remaining;
            else { piece = (remaining + 2) / 3; if (piece < 16) piece = 16; if (piece > chunk) piece = chunk; if (piece > remaining) piece = remaining; }
            P.tail_begin[n_tail] = begin;
            P.tail_size[n_tail] = piece;
            begin += piece;
            remaining -= piece;
            n_tail++;
        }
        for (int i = n_tail; i < kMaxTailChunks; ++i) { P.tail_begin[i] = begin; P.tail_size[i] = 0; }
        P.n_chunks = P.n_big + n_tail;
    }
    P.n_items = P.n_tiles * P.n_chunks;
    P.reference_boxes = L.reference_boxes;
    P.counters = L.counters;
    P.accum = reinterpret_cast<float4*>(L.accum);
    L.grid = grid;
    L.block = kBlock;
    L.chunk = chunk;
    if (uniform) render_interp_kernel<true><<<grid, kBlock, 0, stream>>>(P);
    else render_interp_kernel<false><<<grid, kBlock, 0, stream>>>(P);
    return cudaGetLastError();
}

cudaError_t launch_trace_hits(const hrt::DeviceSceneHost& S, const hrt_ray* d_rays, int n, const float* d_xi, hrt_hit* d_out,
                              int reference_boxes, cudaStream_t stream) {
    if (n <= 0) return cudaSuccess;
    if (reference_boxes & 4)  // bit 2: the warp-uniform walk
        trace_hits_uniform_kernel<<<(n + 127) / 128, 128, 0, stream>>>(to_device_scene(S), d_rays, n, d_xi, d_out, reference_boxes & 1);
    else
        trace_hits_kernel<<<(n + 127) / 128, 128, 0, stream>>>(to_device_scene(S), d_rays, n, d_xi, d_out, reference_boxes & 1);
    return cudaGetLastError();
}
cudaError_t launch_tex_value(const hrt::DeviceSceneHost& S, int tex, const float* d_uvp, int n, float* d_out,
                             cudaStream_t stream) {
    if (n <= 0) return cudaSuccess;
    tex_value_kernel<<<(n + kBlock - 1) / kBlock, kBlock, 0, stream>>>(to_device_scene(S), tex, d_uvp, n, d_out);
    return cudaGetLastError();
}
cudaError_t launch_scatter(const hrt::DeviceSceneHost& S, const hrt_ray* d_rays, const hrt_hit* d_hits, const float* d_u4, int n,
                           hrt_scatter_out* d_out, cudaStream_t stream) {
    if (n <= 0) return cudaSuccess;
    scatter_kernel<<<(n + kBlock - 1) / kBlock, kBlock, 0, stream>>>(to_device_scene(S), d_rays, d_hits, d_u4, n, d_out);
    return cudaGetLastError();
}
cudaError_t launch_camera_rays(const hrt_camera_state& cam, const float* d_stuuu, int n, hrt_ray* d_out, cudaStream_t stream) {
    if (n <= 0) return cudaSuccess;
    camera_rays_kernel<<<(n + kBlock - 1) / kBlock, kBlock, 0, stream>>>(to_camera(cam), d_stuuu, n, d_out);
    return cudaGetLastError();
}
cudaError_t launch_resolve(const float* d_accum, int n_pixels, int samples, float* d_out, cudaStream_t stream) {
    if (n_pixels <= 0) return cudaSuccess;
    const float scale = 1.0f / (float)samples;  // application.rs:403
    resolve_kernel<<<(n_pixels + 255) / 256, 256, 0, stream>>>(reinterpret_cast<const float4*>(d_accum), n_pixels, scale,
                                                              reinterpret_cast<float4*>(d_out));
    return cudaGetLastError();
}
void philox_uniforms(uint64_t seed, uint32_t pixel, uint32_t sample, uint32_t bounce, uint32_t block, float out[4]) {
    RngKey k;
    k.k0 = (uint32_t)seed; k.k1 = (uint32_t)(seed >> 32); k.pixel = pixel; k.sample = sample;
    rng_block(k, bounce, block, out);
}

}  // namespace HRT_NS
