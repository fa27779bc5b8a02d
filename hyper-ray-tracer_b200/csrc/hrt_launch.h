// hrt_launch.h — host-callable launchers exported by each compilation of hrt_kernels.cu
// (namespace hrt_exact: --fmad=false parity build; namespace hrt_fast: production build).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/hrt.h"
#include "hrt_types.h"

namespace hrt {

struct DeviceSceneHost {  // mirrors HRT_NS::DeviceScene field for field (checked by static_assert in the .cu)
    const void* ops;
    const void* nodes;
    const void* ctxs;
    const void* mats;
    const void* texs;
    const void* noise;
    cudaTextureObject_t images[kMaxImages];
    int32_t n_ops, n_noise, n_media;
    float ln_e;
};

struct RenderLaunch {
    DeviceSceneHost scene;
    hrt_camera_state cam;
    int32_t width, height, depth;
    float background[3];
    uint32_t key0, key1;
    int32_t sample_begin, sample_count;
    int32_t chunk;        // samples per work item
    int32_t reference_boxes;
    int32_t n_nodes;      // tree nodes of the scene form being rendered
    int32_t n_pre;        // OP_BVH trees the wavefront render walks in its own stage (<= kMaxPreTrees)
    hrt::PreTree pre[hrt::kMaxPreTrees];
    int32_t interpreter;  // render variant: 1 persistent kernel, per-lane interpreter; 3 persistent kernel, warp-uniform
                          // walk; 5 wavefront render
    unsigned long long* counters;  // device: [0] work-item cursor, [1] rays, [2] paths
    float* accum;                  // device: width*height*4 f32, added into
    int32_t grid, block;           // out: launch configuration actually used
    int32_t launches;              // out: kernels launched
};

// Device buffers of the wavefront render (hrt_kernels.cu launch_render_wave), owned per device.
constexpr int kWaveParts = 8;  // most partitions of the path slots, each iterating on its own stream (default 2 in use)
constexpr int kWaveCounters = 8;  // queue counters per partition (2 per pre-walked tree, 1 for deferred noise, padding)
struct WaveBuffers {
    float* state = nullptr;   // [kWaveParts][WF_WORDS][n_slots / kWaveParts]
    int32_t n_slots = 0;      // all partitions together
    float* tq = nullptr;      // tree-walk queues: [kWaveParts][kMaxPreTrees][n_slots / kWaveParts] entries of 8 words
    float* pre = nullptr;     // tree-walk results: [kWaveParts][n_slots / kWaveParts][kMaxPreTrees] x {t, code}
    float* xq = nullptr;      // deferred noise-texture evaluations: [kWaveParts][n_slots / kWaveParts] x 4 words
    int* tq_count = nullptr;  // per partition: kWaveCounters queue counters
    double* acc64 = nullptr;  // [acc_pixels][4]
    size_t acc_pixels = 0;
    int* d_live = nullptr;    // per partition: 2 counters
    int* h_live = nullptr;    // pinned mirror
    cudaStream_t streams[kWaveParts] = {};
    cudaEvent_t ev[2 * kWaveParts] = {};
    cudaEvent_t ev_end[kWaveParts] = {};
    cudaEvent_t ev_begin = nullptr;
};
constexpr int kWaveStateWords = 16;  // == WF_WORDS

}  // namespace hrt

#define HRT_DECLARE_LAUNCHERS(NS)                                                                                        \
    namespace NS {                                                                                                       \
    cudaError_t launch_render(hrt::RenderLaunch& L, int num_sms, cudaStream_t stream);                                        \
    cudaError_t launch_render_wave(hrt::RenderLaunch& L, hrt::WaveBuffers& W, int num_sms, cudaStream_t stream);              \
    cudaError_t launch_trace_hits(const hrt::DeviceSceneHost& S, const hrt_ray* d_rays, int n, const float* d_xi,             \
                                  hrt_hit* d_out, int reference_boxes, cudaStream_t stream);                             \
    cudaError_t launch_tex_value(const hrt::DeviceSceneHost& S, int tex, const float* d_uvp, int n, float* d_out,             \
                                 cudaStream_t stream);                                                                   \
    cudaError_t launch_scatter(const hrt::DeviceSceneHost& S, const hrt_ray* d_rays, const hrt_hit* d_hits, const float* d_u4, \
                               int n, hrt_scatter_out* d_out, cudaStream_t stream);                                      \
    cudaError_t launch_camera_rays(const hrt_camera_state& cam, const float* d_stuuu, int n, hrt_ray* d_out,             \
                                   cudaStream_t stream);                                                                 \
    cudaError_t launch_resolve(const float* d_accum, int n_pixels, int samples, float* d_out, cudaStream_t stream);      \
    void philox_uniforms(uint64_t seed, uint32_t pixel, uint32_t sample, uint32_t bounce, uint32_t block, float out[4]); \
    }

namespace hrt_fast {
cudaError_t launch_fma_peak(float* d_sink, int grid, int iters, cudaStream_t stream);
cudaError_t launch_reduce_resolve(const float* const* d_accums, int n, int n_pixels, int samples, float* d_out, float* d_sum_out,
                                  cudaStream_t stream);
cudaError_t launch_l2_read(const float4* d_buf, size_t n_vec, int repeats, float* d_sink, int grid, cudaStream_t stream);
}

HRT_DECLARE_LAUNCHERS(hrt_exact)
HRT_DECLARE_LAUNCHERS(hrt_fast)
