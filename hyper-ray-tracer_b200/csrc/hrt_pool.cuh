// hrt_pool.cuh — render kernel with a warp-private RAY POOL in shared memory.
//
// Measured on the in-register warp scheduler (profiles/r01_render_kernel_summary.md): every record class runs with only
// ~11 of 32 lanes, whatever the voting policy, because a parked ray occupies its lane.  Here a warp owns kPoolSlots
// (3 x 32) rays whose state lives in shared memory; a round
//   1. counts the pool's rays per class (each lane looks after three "home" slots),
//   2. GATHERS up to 32 slots of the chosen class (ballot/popc compaction into a slot list),
//   3. loads those rays into registers, runs the class body (several box steps while most lanes stay at a box),
//   4. SCATTERS the rays back and records their new class.
// Parked rays cost no lanes, so rounds run nearly full.  Per-ray traversal order is unchanged (the reference's), and the
// class bodies are the very same step_* functions as in hrt_machine.cuh, so each ray's result is identical.
// The kernel (render_pool_kernel, hrt_kernels.cu) runs one 16-warp block per SM; behind the 16 pools the same dynamic
// shared memory holds the fp16 box table (hrt_types.h Box16) the box rounds read, and the perlin tables.
#pragma once
#include "hrt_machine.cuh"

namespace HRT_NS {

#ifndef HRT_POOL_SLOTS
#define HRT_POOL_SLOTS 96
#endif
constexpr int kPoolSlots = HRT_POOL_SLOTS;   // rays per warp (32 x home slots per lane)
constexpr int kPoolHomes = kPoolSlots / 32;
enum PoolField {
    PF_WOX, PF_WOY, PF_WOZ, PF_WDX, PF_WDY, PF_WDZ, PF_TIME,  // world-space ray segment
    PF_COX, PF_COY, PF_COZ, PF_CDX, PF_CDY, PF_CDZ,           // ray in the current context
    PF_PC, PF_CLOSEST, PF_BEST_PC, PF_BEST_FC, PF_CTX,        // traversal state (BEST_FC = face | ctx << 8)
    PF_TX, PF_TY, PF_TZ, PF_PIXEL, PF_SAMPLE, PF_BOUNCE_PL,   // path state (BOUNCE_PL = bounce | pixel-lane << 16)
    PF_WORDS
};
constexpr int kPoolWarpWords = PF_WORDS * kPoolSlots + 32 /* class bytes, 4 per lane */ + 32 /* gather list */;
#ifndef HRT_POOL_BOX_STEPS
#define HRT_POOL_BOX_STEPS 8
#endif
#ifndef HRT_POOL_BOX_KEEP
#define HRT_POOL_BOX_KEEP 12
#endif
constexpr int kPoolMaxBoxSteps = HRT_POOL_BOX_STEPS;  // box steps per gather while enough lanes stay at a box
constexpr int kPoolBoxKeep = HRT_POOL_BOX_KEEP;       // ... "enough" lanes

struct PoolWarp {  // views into this warp's slice of dynamic shared memory
    float* f;          // [PF_WORDS][kPoolSlots]
    uint32_t* cls_w;   // [32]: byte j of word h = class of slot h + 32 j
    int* list;         // [32]
    __device__ __forceinline__ float& at(int field, int slot) const { return f[field * kPoolSlots + slot]; }
    __device__ __forceinline__ void set_cls(int slot, int c) const {
        reinterpret_cast<uint8_t*>(cls_w)[(slot & 31) * 4 + (slot >> 5)] = (uint8_t)c;
    }
};

template <class PoolT>
__device__ __forceinline__ void pool_load_traversal(const PoolT& W, int s, Lane& L) {
    L.cur.o = v3(W.at(PF_COX, s), W.at(PF_COY, s), W.at(PF_COZ, s));
    L.cur.d = v3(W.at(PF_CDX, s), W.at(PF_CDY, s), W.at(PF_CDZ, s));
    L.cur.time = W.at(PF_TIME, s);
    L.pc = __float_as_int(W.at(PF_PC, s));
    L.closest = W.at(PF_CLOSEST, s);
    L.best_pc = __float_as_int(W.at(PF_BEST_PC, s));
    const int fc = __float_as_int(W.at(PF_BEST_FC, s));
    L.best_face = fc & 0xff;
    L.best_ctx = fc >> 8;
    L.ctx = __float_as_int(W.at(PF_CTX, s));
}
template <class PoolT>
__device__ __forceinline__ void pool_store_traversal(const PoolT& W, int s, const Lane& L, bool cur_changed) {
    W.at(PF_PC, s) = __int_as_float(L.pc);
    W.at(PF_CLOSEST, s) = L.closest;
    W.at(PF_BEST_PC, s) = __int_as_float(L.best_pc);
    W.at(PF_BEST_FC, s) = __int_as_float(L.best_face | (L.best_ctx << 8));
    if (cur_changed) {
        W.at(PF_COX, s) = L.cur.o.x; W.at(PF_COY, s) = L.cur.o.y; W.at(PF_COZ, s) = L.cur.o.z;
        W.at(PF_CDX, s) = L.cur.d.x; W.at(PF_CDY, s) = L.cur.d.y; W.at(PF_CDZ, s) = L.cur.d.z;
        W.at(PF_CTX, s) = __int_as_float(L.ctx);
    }
}
template <class PoolT>
__device__ __forceinline__ Ray pool_load_world(const PoolT& W, int s) {
    Ray w;
    w.o = v3(W.at(PF_WOX, s), W.at(PF_WOY, s), W.at(PF_WOZ, s));
    w.d = v3(W.at(PF_WDX, s), W.at(PF_WDY, s), W.at(PF_WDZ, s));
    w.time = W.at(PF_TIME, s);
    return w;
}
// A new ray segment starts: world == current ray, traversal state reset (lane_start).
template <class PoolT>
__device__ __forceinline__ void pool_store_segment(const PoolT& W, int s, const Ray& w) {
    W.at(PF_WOX, s) = w.o.x; W.at(PF_WOY, s) = w.o.y; W.at(PF_WOZ, s) = w.o.z;
    W.at(PF_WDX, s) = w.d.x; W.at(PF_WDY, s) = w.d.y; W.at(PF_WDZ, s) = w.d.z;
    W.at(PF_TIME, s) = w.time;
    W.at(PF_COX, s) = w.o.x; W.at(PF_COY, s) = w.o.y; W.at(PF_COZ, s) = w.o.z;
    W.at(PF_CDX, s) = w.d.x; W.at(PF_CDY, s) = w.d.y; W.at(PF_CDZ, s) = w.d.z;
    W.at(PF_PC, s) = __int_as_float(0);
    W.at(PF_CLOSEST, s) = CUDART_INF_F;
    W.at(PF_BEST_PC, s) = __int_as_float(-1);
    W.at(PF_BEST_FC, s) = __int_as_float(0);
    W.at(PF_CTX, s) = __int_as_float(0);
}

// Class populations of the pool: each lane reads the classes of its three home slots (one word).
struct PoolCounts {
    int n[6];  // box, sphere, rect, misc, done, new
};
__device__ __forceinline__ PoolCounts pool_count(uint32_t my_cls_word) {
    // two REDUX.ADD over 10-bit packed counters (three classes per word; populations <= 96)
    unsigned a = 0u, b = 0u;
#pragma unroll
    for (int j = 0; j < kPoolHomes; ++j) {
        const int c = (int)((my_cls_word >> (8 * j)) & 0xffu);
        if (c < 3) a += 1u << (10 * c);
        else if (c < 6) b += 1u << (10 * (c - 3));
    }
    a = __reduce_add_sync(0xffffffffu, a);
    b = __reduce_add_sync(0xffffffffu, b);
    PoolCounts p;
    p.n[0] = (int)(a & 1023u); p.n[1] = (int)((a >> 10) & 1023u); p.n[2] = (int)((a >> 20) & 1023u);
    p.n[3] = (int)(b & 1023u); p.n[4] = (int)((b >> 10) & 1023u); p.n[5] = (int)((b >> 20) & 1023u);
    return p;
}
// Compact up to 32 slots of class `c` into W.list; returns how many.  `start` rotates which home row is scanned first
// so that no row is starved.
__device__ __forceinline__ int pool_gather(const PoolWarp& W, uint32_t my_cls_word, int c, int lane, int start) {
    const unsigned lt = (1u << lane) - 1u;
    int base = 0;
#pragma unroll
    for (int jj = 0; jj < kPoolHomes; ++jj) {
        int j = jj + start;
        if (j >= kPoolHomes) j -= kPoolHomes;
        const bool mine = (int)((my_cls_word >> (8 * j)) & 0xffu) == c;
        const unsigned m = __ballot_sync(0xffffffffu, mine);
        const int pos = base + __popc(m & lt);
        if (mine && pos < 32) W.list[pos] = lane + 32 * j;
        base += __popc(m);
    }
    __syncwarp();
    return base < 32 ? base : 32;
}

}  // namespace HRT_NS
