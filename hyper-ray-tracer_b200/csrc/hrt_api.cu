// hrt_api.cu — C-ABI compute entry points: device upload of the flattened scene, render / resolve /
// parity launches.  No CPU fallback: every function here returns HRT_ERR_CUDA when the CUDA runtime has no
// usable device.
#include <cuda_runtime.h>

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

#include "../../include/hrt.h"
#include "hrt_launch.h"
#include "hrt_scene.hpp"

namespace hrt {

constexpr int kCounterWords = 32;  // [0] work cursor / next path index, [1] rays, [2] paths
constexpr int kLaunchSlots = 8;

// What one render launch owns on the device: its counter block (the kernel's work cursor lives there) and its timing
// events.  hrt_render_accum_device returns before its kernel finishes, so two renders of one scene can be in flight on
// one device (different streams): each takes the next slot of a ring, and re-using a slot first waits (on the new
// launch's stream) for the slot's previous launch to finish.
struct LaunchSlot {
    unsigned long long* counters = nullptr;
    cudaEvent_t t0 = nullptr, t1 = nullptr, done = nullptr;
    bool used = false;
};

struct DeviceState {
    int device = -1;
    int num_sms = 0;
    // [0] reference form, [1] fast form of the flattened scene (hrt_scene.hpp), [2] the fast form's wave variant (ops only)
    void* d_ops[3] = {nullptr, nullptr, nullptr};
    void* d_ctxs[2] = {nullptr, nullptr};
    void* d_nodes = nullptr;
    void* d_mats = nullptr;
    void* d_texs = nullptr;
    void* d_noise = nullptr;
    std::vector<cudaArray_t> arrays;
    std::vector<cudaTextureObject_t> texobjs;
    unsigned long long* d_counters = nullptr;  // kLaunchSlots blocks of kCounterWords
    LaunchSlot slots[kLaunchSlots];
    int next_slot = 0;
    WaveBuffers wave;  // path slots of the wavefront render, allocated on first use
    // scratch for the host-buffer entry points
    float* d_accum = nullptr;
    float* d_rgba = nullptr;
    size_t accum_pixels = 0;
    cudaEvent_t ev[6] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    DeviceSceneHost view[3];
};

void release_device_state(DeviceState* d) {
    if (!d) return;
    int prev = 0;
    if (cudaGetDevice(&prev) == cudaSuccess && cudaSetDevice(d->device) == cudaSuccess) {
        for (auto t : d->texobjs) cudaDestroyTextureObject(t);
        for (auto a : d->arrays) cudaFreeArray(a);
        for (int k = 0; k < 2; ++k) { cudaFree(d->d_ops[k]); cudaFree(d->d_ctxs[k]); }
        cudaFree(d->d_ops[2]);
        cudaFree(d->d_nodes); cudaFree(d->d_mats); cudaFree(d->d_texs); cudaFree(d->d_noise);
        cudaFree(d->d_counters); cudaFree(d->d_accum); cudaFree(d->d_rgba);
        for (auto e : d->ev) if (e) cudaEventDestroy(e);
        cudaFree(d->wave.state); cudaFree(d->wave.d_live); cudaFree(d->wave.acc64);
        cudaFree(d->wave.tq); cudaFree(d->wave.pre); cudaFree(d->wave.tq_count); cudaFree(d->wave.xq);
        if (d->wave.h_live) cudaFreeHost(d->wave.h_live);
        for (auto e : d->wave.ev) if (e) cudaEventDestroy(e);
        for (auto e : d->wave.ev_end) if (e) cudaEventDestroy(e);
        if (d->wave.ev_begin) cudaEventDestroy(d->wave.ev_begin);
        for (auto st : d->wave.streams) if (st) cudaStreamDestroy(st);
        for (auto& sl : d->slots) {
            if (sl.t0) cudaEventDestroy(sl.t0);
            if (sl.t1) cudaEventDestroy(sl.t1);
            if (sl.done) cudaEventDestroy(sl.done);
        }
        cudaSetDevice(prev);
    }
    delete d;
}

static int32_t cuda_fail(cudaError_t e, const char* what) {
    return fail(HRT_ERR_CUDA, std::string(what) + ": " + cudaGetErrorString(e));
}
#define HRT_CUDA(call)                                      \
    do {                                                    \
        cudaError_t _e = (call);                            \
        if (_e != cudaSuccess) return cuda_fail(_e, #call); \
    } while (0)

static int32_t upload_table(void** dst, const void* src, size_t bytes) {
    if (bytes == 0) bytes = 16;  // keep pointers valid
    HRT_CUDA(cudaMalloc(dst, bytes));
    HRT_CUDA(cudaMemset(*dst, 0, bytes));
    if (src) HRT_CUDA(cudaMemcpy(*dst, src, bytes, cudaMemcpyHostToDevice));
    return HRT_OK;
}

static DeviceState* find_state(hrt_scene* s, int device) {
    for (DeviceState* d : s->devices)
        if (d && d->device == device) return d;
    return nullptr;
}

static int32_t ensure_wave(DeviceState* d, size_t pixels) {
    if (d->wave.acc_pixels < pixels) {
        cudaFree(d->wave.acc64);
        d->wave.acc64 = nullptr;
        d->wave.acc_pixels = 0;
        HRT_CUDA(cudaMalloc((void**)&d->wave.acc64, sizeof(double) * 4 * pixels));
        d->wave.acc_pixels = pixels;
    }
    if (d->wave.state) return HRT_OK;
    int n = 8 << 20;  // most path slots in flight, all partitions together (64 B of state + 96 B of stage buffers each: 1.3 GB)
    if (const char* env = getenv("HRT_WAVE_SLOTS")) n = atoi(env);
    if (n < 256 * kWaveParts) n = 256 * kWaveParts;
    if (n > (4 << 20) * kWaveParts) n = (4 << 20) * kWaveParts;  // slot ids share a word with a texture id (wave_noise_kernel)
    n = (n + 256 * kWaveParts - 1) / (256 * kWaveParts) * (256 * kWaveParts);
    HRT_CUDA(cudaMalloc((void**)&d->wave.state, sizeof(float) * (size_t)kWaveStateWords * (size_t)n));
    HRT_CUDA(cudaMalloc((void**)&d->wave.d_live, 2 * kWaveParts * sizeof(int)));
    HRT_CUDA(cudaMalloc((void**)&d->wave.tq, sizeof(float) * 8 * (size_t)kMaxPreTrees * (size_t)n));
    HRT_CUDA(cudaMalloc((void**)&d->wave.pre, sizeof(float) * 2 * (size_t)kMaxPreTrees * (size_t)n));
    HRT_CUDA(cudaMalloc((void**)&d->wave.tq_count, kWaveCounters * kWaveParts * sizeof(int)));
    HRT_CUDA(cudaMalloc((void**)&d->wave.xq, sizeof(float) * 4 * (size_t)n));
    HRT_CUDA(cudaMallocHost((void**)&d->wave.h_live, 2 * kWaveParts * sizeof(int)));
    for (auto& e : d->wave.ev) HRT_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    for (auto& e : d->wave.ev_end) HRT_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    HRT_CUDA(cudaEventCreateWithFlags(&d->wave.ev_begin, cudaEventDisableTiming));
    for (auto& st : d->wave.streams) HRT_CUDA(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
    d->wave.n_slots = n;
    return HRT_OK;
}

static int32_t ensure_scratch(DeviceState* d, size_t pixels) {
    if (d->accum_pixels >= pixels && d->d_accum) return HRT_OK;
    cudaFree(d->d_accum); cudaFree(d->d_rgba);
    d->d_accum = d->d_rgba = nullptr;
    d->accum_pixels = 0;
    HRT_CUDA(cudaMalloc((void**)&d->d_accum, pixels * 16));
    HRT_CUDA(cudaMalloc((void**)&d->d_rgba, pixels * 16));
    d->accum_pixels = pixels;
    return HRT_OK;
}

template <typename T>
struct DevBuf {
    T* p = nullptr;
    ~DevBuf() { cudaFree(p); }
    cudaError_t alloc(size_t n) { return cudaMalloc((void**)&p, (n ? n : 1) * sizeof(T)); }
};

static const FlatScene& flat_of(const hrt_scene* s, int k) { return k == 0 ? s->ref : s->fast; }

static int32_t fill_device_state(hrt_scene* s, DeviceState* d) {
    int32_t rc;
    for (int k = 0; k < 2; ++k) {
        const FlatScene& f = flat_of(s, k);
        if ((rc = upload_table(&d->d_ops[k], f.ops.data(), f.ops.size() * sizeof(Op))) != HRT_OK) return rc;
        if ((rc = upload_table(&d->d_ctxs[k], f.ctxs.data(), f.ctxs.size() * sizeof(Ctx))) != HRT_OK) return rc;
    }
    if (!s->fast.wave_ops.empty() &&
        (rc = upload_table(&d->d_ops[2], s->fast.wave_ops.data(), s->fast.wave_ops.size() * sizeof(Op))) != HRT_OK) return rc;
    if ((rc = upload_table(&d->d_nodes, s->fast.nodes.data(), s->fast.nodes.size() * sizeof(Bvh2Node))) != HRT_OK) return rc;
    if ((rc = upload_table(&d->d_mats, s->materials.data(), s->materials.size() * sizeof(Material))) != HRT_OK) return rc;
    if ((rc = upload_table(&d->d_texs, s->textures.data(), s->textures.size() * sizeof(Texture))) != HRT_OK) return rc;
    if ((rc = upload_table(&d->d_noise, s->noise_tables.data(), s->noise_tables.size() * sizeof(NoiseTable))) != HRT_OK) return rc;
    HRT_CUDA(cudaMalloc((void**)&d->d_counters, kLaunchSlots * kCounterWords * sizeof(unsigned long long)));
    for (auto& ev : d->ev) HRT_CUDA(cudaEventCreate(&ev));
    for (int i = 0; i < kLaunchSlots; ++i) {
        LaunchSlot& sl = d->slots[i];
        sl.counters = d->d_counters + (size_t)i * kCounterWords;
        HRT_CUDA(cudaEventCreate(&sl.t0));
        HRT_CUDA(cudaEventCreate(&sl.t1));
        HRT_CUDA(cudaEventCreateWithFlags(&sl.done, cudaEventDisableTiming));
    }
    // Image textures: RGBA8 CUDA arrays bound as point-sampled, unnormalised texture objects
    // (nearest texel, no filtering — image_texture.rs:44-62).
    for (const ImageData& img : s->images) {
        cudaChannelFormatDesc desc = cudaCreateChannelDesc<uchar4>();
        cudaArray_t arr = nullptr;
        HRT_CUDA(cudaMallocArray(&arr, &desc, img.width, img.height));
        d->arrays.push_back(arr);
        HRT_CUDA(cudaMemcpy2DToArray(arr, 0, 0, img.rgba.data(), (size_t)img.width * 4, (size_t)img.width * 4, img.height,
                                     cudaMemcpyHostToDevice));
        cudaResourceDesc rd;
        std::memset(&rd, 0, sizeof(rd));
        rd.resType = cudaResourceTypeArray;
        rd.res.array.array = arr;
        cudaTextureDesc td;
        std::memset(&td, 0, sizeof(td));
        td.addressMode[0] = td.addressMode[1] = cudaAddressModeClamp;
        td.filterMode = cudaFilterModePoint;
        td.readMode = cudaReadModeElementType;
        td.normalizedCoords = 0;
        cudaTextureObject_t tex = 0;
        HRT_CUDA(cudaCreateTextureObject(&tex, &rd, &td, nullptr));
        d->texobjs.push_back(tex);
    }
    for (int k = 0; k < 2; ++k) {
        const FlatScene& f = flat_of(s, k);
        DeviceSceneHost& v = d->view[k];
        std::memset(&v, 0, sizeof(v));
        v.ops = d->d_ops[k]; v.ctxs = d->d_ctxs[k];
        v.nodes = d->d_nodes;  // only the fast form has OP_BVH records
        v.mats = d->d_mats; v.texs = d->d_texs; v.noise = d->d_noise;
        for (size_t i = 0; i < d->texobjs.size() && i < (size_t)kMaxImages; ++i) v.images[i] = d->texobjs[i];
        v.n_ops = (int32_t)f.ops.size();
        v.n_noise = (int32_t)s->noise_tables.size();
        v.n_media = f.n_media;
        v.ln_e = logf(2.71828182845904523536f);
    }
    d->view[2] = d->view[1];
    if (d->d_ops[2]) d->view[2].ops = d->d_ops[2];
    return HRT_OK;
}

}  // namespace hrt

using namespace hrt;

extern "C" {

int32_t hrt_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    return n;
}

int32_t hrt_scene_upload(hrt_scene* s, int32_t device) {
    if (!s) return fail(HRT_ERR_INVALID, "null scene");
    if (!s->committed) return fail(HRT_ERR_STATE, "scene not committed");
    if (find_state(s, device)) return HRT_OK;
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0)
        return fail(HRT_ERR_CUDA, std::string("no CUDA device available (there is no CPU fallback): ") +
                                      (e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0"));
    if (device < 0 || device >= n) return fail(HRT_ERR_INVALID, "device index out of range");
    HRT_CUDA(cudaSetDevice(device));
    cudaDeviceProp prop;
    HRT_CUDA(cudaGetDeviceProperties(&prop, device));
    if (prop.major != 10 || prop.minor != 0)  // the library ships an sm_100a cubin and no PTX
        return fail(HRT_ERR_CUDA, std::string("device '") + prop.name + "' is sm_" + std::to_string(prop.major) +
                                      std::to_string(prop.minor) + "; this library ships sm_100a code only");
    // Built locally and registered only when complete: a failure below must not leave a half-built state behind that a
    // later call would find and launch kernels on.
    DeviceState* d = new DeviceState();
    d->device = device;
    d->num_sms = prop.multiProcessorCount;
    const int32_t rc = fill_device_state(s, d);
    if (rc != HRT_OK) {
        release_device_state(d);
        return rc;
    }
    s->devices.push_back(d);
    return HRT_OK;
}

int32_t hrt_scene_evict(hrt_scene* s, int32_t device) {
    if (!s) return fail(HRT_ERR_INVALID, "null scene");
    for (size_t i = 0; i < s->devices.size(); ++i)
        if (s->devices[i] && s->devices[i]->device == device) {
            release_device_state(s->devices[i]);
            s->devices.erase(s->devices.begin() + (long)i);
            return HRT_OK;
        }
    return HRT_OK;
}

int32_t hrt_scene_refresh(hrt_scene* s, int32_t device) {
    if (!s) return fail(HRT_ERR_INVALID, "null scene");
    if (!s->committed) return fail(HRT_ERR_STATE, "scene not committed");
    DeviceState* d = find_state(s, device);
    if (!d) return hrt_scene_upload(s, device);
    HRT_CUDA(cudaSetDevice(device));
    for (int k = 0; k < 2; ++k) {
        const FlatScene& f = flat_of(s, k);
        HRT_CUDA(cudaMemcpyAsync(d->d_ops[k], f.ops.data(), f.ops.size() * sizeof(Op), cudaMemcpyHostToDevice, 0));
        HRT_CUDA(cudaMemcpyAsync(d->d_ctxs[k], f.ctxs.data(), f.ctxs.size() * sizeof(Ctx), cudaMemcpyHostToDevice, 0));
    }
    if (d->d_ops[2] && !s->fast.wave_ops.empty())
        HRT_CUDA(cudaMemcpyAsync(d->d_ops[2], s->fast.wave_ops.data(), s->fast.wave_ops.size() * sizeof(Op), cudaMemcpyHostToDevice, 0));
    if (!s->fast.nodes.empty())
        HRT_CUDA(cudaMemcpyAsync(d->d_nodes, s->fast.nodes.data(), s->fast.nodes.size() * sizeof(Bvh2Node), cudaMemcpyHostToDevice, 0));
    if (!s->materials.empty())
        HRT_CUDA(cudaMemcpyAsync(d->d_mats, s->materials.data(), s->materials.size() * sizeof(Material), cudaMemcpyHostToDevice, 0));
    if (!s->textures.empty())
        HRT_CUDA(cudaMemcpyAsync(d->d_texs, s->textures.data(), s->textures.size() * sizeof(Texture), cudaMemcpyHostToDevice, 0));
    if (!s->noise_tables.empty())
        HRT_CUDA(cudaMemcpyAsync(d->d_noise, s->noise_tables.data(), s->noise_tables.size() * sizeof(NoiseTable), cudaMemcpyHostToDevice, 0));
    for (size_t i = 0; i < s->images.size() && i < d->arrays.size(); ++i) {
        const ImageData& img = s->images[i];
        HRT_CUDA(cudaMemcpy2DToArrayAsync(d->arrays[i], 0, 0, img.rgba.data(), (size_t)img.width * 4, (size_t)img.width * 4, img.height,
                                          cudaMemcpyHostToDevice, 0));
    }
    return HRT_OK;
}

int64_t hrt_scene_device_bytes(const hrt_scene* s) {
    if (!s || !s->committed) return fail(HRT_ERR_STATE, "scene not committed");
    int64_t b = (int64_t)(s->materials.size() * sizeof(Material) + s->textures.size() * sizeof(Texture) +
                          s->noise_tables.size() * sizeof(NoiseTable) + s->fast.nodes.size() * sizeof(Bvh2Node));
    for (int k = 0; k < 2; ++k) {
        const FlatScene& f = flat_of(s, k);
        b += (int64_t)(f.ops.size() * sizeof(Op) + f.ctxs.size() * sizeof(Ctx));
    }
    b += (int64_t)(s->fast.wave_ops.size() * sizeof(Op));
    for (const ImageData& img : s->images) b += (int64_t)img.rgba.size();
    return b;
}

int32_t hrt_measure_peaks(int32_t device, hrt_peaks* out) {
    if (!out) return fail(HRT_ERR_INVALID, "null output");
    int n = 0;
    cudaError_t ce = cudaGetDeviceCount(&n);
    if (ce != cudaSuccess || n == 0) return fail(HRT_ERR_CUDA, "no CUDA device available (there is no CPU fallback)");
    HRT_CUDA(cudaSetDevice(device));
    int sms = 0, khz = 0;
    HRT_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device));
    HRT_CUDA(cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, device));
    DevBuf<float> sink;
    DevBuf<float4> buf;
    const size_t n_vec = (32u << 20) / sizeof(float4);
    HRT_CUDA(sink.alloc(4));
    HRT_CUDA(buf.alloc(n_vec));
    HRT_CUDA(cudaMemset(buf.p, 0, n_vec * sizeof(float4)));
    cudaEvent_t e0, e1;
    HRT_CUDA(cudaEventCreate(&e0));
    HRT_CUDA(cudaEventCreate(&e1));
    const int grid = sms * 8, iters = 4096, repeats = 64;
    float best_fma = 1e30f, best_l2 = 1e30f;
    for (int rep = 0; rep < 4; ++rep) {  // first pass warms clocks / L2
        float ms = 0.0f;
        HRT_CUDA(cudaEventRecord(e0, 0));
        HRT_CUDA(hrt_fast::launch_fma_peak(sink.p, grid, iters, 0));
        HRT_CUDA(cudaEventRecord(e1, 0));
        HRT_CUDA(cudaEventSynchronize(e1));
        HRT_CUDA(cudaEventElapsedTime(&ms, e0, e1));
        if (rep > 0 && ms < best_fma) best_fma = ms;
        HRT_CUDA(cudaEventRecord(e0, 0));
        HRT_CUDA(hrt_fast::launch_l2_read(buf.p, n_vec, repeats, sink.p, grid, 0));
        HRT_CUDA(cudaEventRecord(e1, 0));
        HRT_CUDA(cudaEventSynchronize(e1));
        HRT_CUDA(cudaEventElapsedTime(&ms, e0, e1));
        if (rep > 0 && ms < best_l2) best_l2 = ms;
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    const double flops = (double)grid * 256.0 * (double)iters * 64.0 * 2.0;
    out->fp32_tflops = (float)(flops / (best_fma * 1e-3) / 1e12);
    out->l2_read_gbs = (float)((double)n_vec * 16.0 * repeats / (best_l2 * 1e-3) / 1e9);
    out->fma_ms = best_fma;
    out->l2_ms = best_l2;
    out->sm_count = sms;
    out->clock_khz = khz;
    return HRT_OK;
}

static int32_t get_state(hrt_scene* s, int32_t device, DeviceState** out) {
    int32_t rc = hrt_scene_upload(s, device);
    if (rc != HRT_OK) return rc;
    *out = find_state(s, device);
    HRT_CUDA(cudaSetDevice(device));
    return HRT_OK;
}

static int32_t check_render_args(const hrt_scene* s, const hrt_camera_desc* cam, const hrt_render_desc* rd) {
    if (!cam || !rd) return fail(HRT_ERR_INVALID, "null camera/render descriptor");
    if (rd->width <= 0 || rd->height <= 0 || rd->samples <= 0 || rd->depth < 0)
        return fail(HRT_ERR_INVALID, "render: width/height/samples must be positive, depth non-negative");
    if (cam->width != rd->width || cam->height != rd->height)
        return fail(HRT_ERR_INVALID, "render: camera size differs from render size (Camera::resize uses the image size)");
    if ((long long)rd->width * rd->height > (1ll << 31) - 1) return fail(HRT_ERR_INVALID, "render: image too large");
    // (the persistent kernels of round 1 kept the bounce index in 16 bits; the limit stays part of the interface)
    if (rd->depth > 65535) return fail(HRT_ERR_INVALID, "render: depth above 65535 is not supported");
    (void)s;
    return HRT_OK;
}

// The part every render variant shares: launch the path-trace kernel for [sample_begin, +count) into d_accum.
static int32_t render_into(hrt_scene* s, DeviceState* d, const hrt_camera_desc* cam, const hrt_render_desc* rd,
                           float* d_accum, cudaStream_t stream, hrt_stats* stats, LaunchSlot** slot_out) {
    int32_t rc = check_render_args(s, cam, rd);
    if (rc != HRT_OK) return rc;
    RenderLaunch L;
    std::memset(&L, 0, sizeof(L));
    if ((rc = hrt_camera_init(cam, &L.cam)) != HRT_OK) return rc;
    L.width = rd->width; L.height = rd->height; L.depth = rd->depth;
    std::memcpy(L.background, rd->background, 12);
    L.key0 = (uint32_t)rd->seed; L.key1 = (uint32_t)(rd->seed >> 32);
    L.sample_begin = rd->sample_count > 0 ? rd->sample_begin : 0;
    L.sample_count = rd->sample_count > 0 ? rd->sample_count : rd->samples;
    if (L.sample_begin < 0 || L.sample_begin + L.sample_count > rd->samples)
        return fail(HRT_ERR_INVALID, "render: sample slice outside [0, samples)");
    // Moving-sphere boxes only cover the BVH build interval (moving_sphere.rs:98-110): a shutter outside it
    // makes those boxes unsound, so fall back to the reference's own box test everywhere.
    bool ref_boxes = (rd->flags & HRT_FLAG_REFERENCE_TRAVERSAL) != 0;
    if (s->any_bvh && (std::fmin(cam->time0, cam->time1) < s->time_min || std::fmax(cam->time0, cam->time1) > s->time_max))
        ref_boxes = true;
    L.reference_boxes = ref_boxes ? 1 : 0;
    L.scene = d->view[ref_boxes ? 0 : 1];  // the reference form of the stream goes with the reference's box test
    L.n_nodes = ref_boxes ? 0 : (int32_t)s->fast.nodes.size();
    L.n_pre = 0;
    if (!ref_boxes && !getenv("HRT_NO_TREE_STAGE"))
        for (const PreTree& t : s->fast.trees)
            if (L.n_pre < kMaxPreTrees) L.pre[L.n_pre++] = t;
    {
        const char* env = getenv("HRT_KERNEL");  // diagnostic override: "interp" | "uniform" | "wave"
        // default: the wavefront render for big jobs on scenes with OP_BVH trees (its compacted tree stage and small
        // kernels win there: `final` 435 vs 270 Mpaths/s), the persistent uniform-walk kernel otherwise (no per-iteration
        // launches, no ramp-up and drain of a wave: Cornell 880 vs 650, `random` at 100 spp 1280 vs 580)
        const long long job_paths = (long long)rd->width * rd->height * L.sample_count;
        int variant = (L.n_pre > 0 && job_paths >= (32ll << 20)) ? 5 : 3;
        if (rd->flags & HRT_FLAG_WAVEFRONT) variant = 5;
        if (rd->flags & HRT_FLAG_UNIFORM) variant = 3;
        if (rd->flags & HRT_FLAG_INTERPRETER) variant = 1;
        if (env && env[0] == 'i') variant = 1;
        if (env && env[0] == 'u') variant = 3;
        if (env && env[0] == 'w') variant = 5;
        L.interpreter = variant;
    }
    LaunchSlot& slot = d->slots[d->next_slot];
    d->next_slot = (d->next_slot + 1) % kLaunchSlots;
    if (slot.used) HRT_CUDA(cudaStreamWaitEvent(stream, slot.done, 0));
    slot.used = true;
    if (slot_out) *slot_out = &slot;
    L.counters = slot.counters;
    L.accum = d_accum;
    L.chunk = 0;
    if (const char* env = getenv("HRT_CHUNK")) L.chunk = atoi(env);  // diagnostic: samples per work item
    HRT_CUDA(cudaMemsetAsync(slot.counters, 0, kCounterWords * sizeof(unsigned long long), stream));
    HRT_CUDA(cudaEventRecord(slot.t0, stream));
    cudaError_t e;
    if (L.interpreter == 5) {
        if ((rc = ensure_wave(d, (size_t)rd->width * rd->height)) != HRT_OK) return rc;
        // the stream walk meets the trees walked ahead where their spans begin (HRT_NO_SPANS: diagnostic, at their OP_BVH records)
        if (!ref_boxes && L.n_pre > 0 && !getenv("HRT_NO_SPANS")) L.scene = d->view[2];
        e = (rd->flags & HRT_FLAG_EXACT_MATH) ? hrt_exact::launch_render_wave(L, d->wave, d->num_sms, stream)
                                              : hrt_fast::launch_render_wave(L, d->wave, d->num_sms, stream);
    } else {
        L.launches = 1;
        e = (rd->flags & HRT_FLAG_EXACT_MATH) ? hrt_exact::launch_render(L, d->num_sms, stream)
                                              : hrt_fast::launch_render(L, d->num_sms, stream);
    }
    if (e != cudaSuccess) return cuda_fail(e, "render_kernel launch");
    HRT_CUDA(cudaEventRecord(slot.t1, stream));
    HRT_CUDA(cudaEventRecord(slot.done, stream));
    if (stats) {
        stats->launches += L.launches;
        stats->grid = L.grid;
        stats->block = L.block;
    }
    return HRT_OK;
}

static int32_t finish_stats(LaunchSlot* slot, cudaStream_t stream, hrt_stats* stats) {
    if (!stats || !slot) return HRT_OK;
    unsigned long long c[kCounterWords];
    HRT_CUDA(cudaMemcpyAsync(c, slot->counters, sizeof(c), cudaMemcpyDeviceToHost, stream));
    HRT_CUDA(cudaStreamSynchronize(stream));
    stats->rays = c[1];
    stats->paths = c[2];
    float ms = 0.0f;
    HRT_CUDA(cudaEventElapsedTime(&ms, slot->t0, slot->t1));
    stats->kernel_ms = ms;
    return HRT_OK;
}

static int32_t render_host(hrt_scene* s, int32_t device, const hrt_camera_desc* cam, const hrt_render_desc* rd, float* out,
                           hrt_stats* stats, bool resolve) {
    if (!out) return fail(HRT_ERR_INVALID, "null output buffer");
    DeviceState* d = nullptr;
    int32_t rc = get_state(s, device, &d);
    if (rc != HRT_OK) return rc;
    if ((rc = check_render_args(s, cam, rd)) != HRT_OK) return rc;
    const size_t pixels = (size_t)rd->width * rd->height;
    if ((rc = ensure_scratch(d, pixels)) != HRT_OK) return rc;
    hrt_stats local;
    std::memset(&local, 0, sizeof(local));
    cudaStream_t stream = 0;
    HRT_CUDA(cudaMemsetAsync(d->d_accum, 0, pixels * 16, stream));
    LaunchSlot* slot = nullptr;
    if ((rc = render_into(s, d, cam, rd, d->d_accum, stream, &local, &slot)) != HRT_OK) return rc;
    const float* src = d->d_accum;
    if (resolve) {
        HRT_CUDA(cudaEventRecord(d->ev[2], stream));
        cudaError_t e = hrt_fast::launch_resolve(d->d_accum, (int)pixels, rd->samples, d->d_rgba, stream);
        if (e != cudaSuccess) return cuda_fail(e, "resolve_kernel launch");
        HRT_CUDA(cudaEventRecord(d->ev[3], stream));
        local.launches += 1;
        src = d->d_rgba;
    }
    HRT_CUDA(cudaEventRecord(d->ev[4], stream));
    HRT_CUDA(cudaMemcpyAsync(out, src, pixels * 16, cudaMemcpyDeviceToHost, stream));
    HRT_CUDA(cudaEventRecord(d->ev[5], stream));
    if ((rc = finish_stats(slot, stream, &local)) != HRT_OK) return rc;
    if (resolve) HRT_CUDA(cudaEventElapsedTime(&local.resolve_ms, d->ev[2], d->ev[3]));
    HRT_CUDA(cudaEventElapsedTime(&local.d2h_ms, d->ev[4], d->ev[5]));
    if (stats) *stats = local;
    return HRT_OK;
}

int32_t hrt_render(hrt_scene* s, int32_t device, const hrt_camera_desc* cam, const hrt_render_desc* rd, float* out_rgba,
                   hrt_stats* stats) {
    return render_host(s, device, cam, rd, out_rgba, stats, true);
}
int32_t hrt_render_accum(hrt_scene* s, int32_t device, const hrt_camera_desc* cam, const hrt_render_desc* rd, float* out_sum,
                         hrt_stats* stats) {
    return render_host(s, device, cam, rd, out_sum, stats, false);
}

// Progressive delivery (application.rs:284-306: the reference shows tiles as they finish) and early exit
// (application.rs:357-391: a resize abandons the frame).  The frame is rendered in batches of `batch_samples` samples —
// disjoint slices of ONE render's sample set, accumulated on the device — and after every batch the frame of the samples
// so far is resolved on the device, copied to `out_rgba` and handed to `on_frame`; a non-zero return cancels the render.
int32_t hrt_render_progressive(hrt_scene* s, int32_t device, const hrt_camera_desc* cam, const hrt_render_desc* rd,
                               int32_t batch_samples, hrt_progress_fn on_frame, void* user, float* out_rgba, hrt_stats* stats) {
    if (!out_rgba) return fail(HRT_ERR_INVALID, "null output buffer");
    if (batch_samples <= 0) return fail(HRT_ERR_INVALID, "render_progressive: batch_samples must be positive");
    DeviceState* d = nullptr;
    int32_t rc = get_state(s, device, &d);
    if (rc != HRT_OK) return rc;
    if ((rc = check_render_args(s, cam, rd)) != HRT_OK) return rc;
    const size_t pixels = (size_t)rd->width * rd->height;
    if ((rc = ensure_scratch(d, pixels)) != HRT_OK) return rc;
    const int begin = rd->sample_count > 0 ? rd->sample_begin : 0;
    const int total = rd->sample_count > 0 ? rd->sample_count : rd->samples;
    hrt_stats agg;
    std::memset(&agg, 0, sizeof(agg));
    cudaStream_t stream = 0;
    HRT_CUDA(cudaMemsetAsync(d->d_accum, 0, pixels * 16, stream));
    int done = 0;
    while (done < total) {
        hrt_render_desc slice = *rd;
        slice.sample_begin = begin + done;
        slice.sample_count = std::min(batch_samples, total - done);
        hrt_stats local;
        std::memset(&local, 0, sizeof(local));
        LaunchSlot* slot = nullptr;
        if ((rc = render_into(s, d, cam, &slice, d->d_accum, stream, &local, &slot)) != HRT_OK) return rc;
        done += slice.sample_count;
        // resolve what has been accumulated so far: sqrt(sum / samples_done) (application.rs:451-456)
        cudaError_t e = hrt_fast::launch_resolve(d->d_accum, (int)pixels, done, d->d_rgba, stream);
        if (e != cudaSuccess) return cuda_fail(e, "resolve_kernel launch");
        HRT_CUDA(cudaMemcpyAsync(out_rgba, d->d_rgba, pixels * 16, cudaMemcpyDeviceToHost, stream));
        if ((rc = finish_stats(slot, stream, &local)) != HRT_OK) return rc;  // synchronises the stream
        agg.paths += local.paths; agg.rays += local.rays; agg.kernel_ms += local.kernel_ms;
        agg.launches += local.launches + 1; agg.grid = local.grid; agg.block = local.block;
        if (on_frame && on_frame(user, done, total, out_rgba) != 0) break;  // cancelled by the front end
    }
    if (stats) *stats = agg;
    return done < total ? HRT_CANCELLED : HRT_OK;
}

// Single-process multi-GPU render (the reference is ONE process): device k renders the k-th sample slice into its own
// accumulator, all devices run concurrently, and the first device sums the others' accumulators over NVLink peer memory
// inside the resolve kernel.
static int32_t render_multi(hrt_scene* s, const int32_t* devices, int32_t n, const hrt_camera_desc* cam, const hrt_render_desc* rd,
                            float* out, hrt_stats* stats, bool resolve) {
    if (!devices || n < 1 || n > 8) return fail(HRT_ERR_INVALID, "render_multi: 1..8 devices");
    if (!out) return fail(HRT_ERR_INVALID, "null output buffer");
    int32_t rc = check_render_args(s, cam, rd);
    if (rc != HRT_OK) return rc;
    for (int i = 0; i < n; ++i)
        for (int j = 0; j < i; ++j)
            if (devices[i] == devices[j]) return fail(HRT_ERR_INVALID, "render_multi: duplicate device");
    const size_t pixels = (size_t)rd->width * rd->height;
    DeviceState* st[8];
    const float* accums[8];
    for (int i = 0; i < n; ++i) {
        if ((rc = get_state(s, devices[i], &st[i])) != HRT_OK) return rc;
        if ((rc = ensure_scratch(st[i], pixels)) != HRT_OK) return rc;
        accums[i] = st[i]->d_accum;
    }
    HRT_CUDA(cudaSetDevice(devices[0]));
    for (int i = 1; i < n; ++i) {
        int can = 0;
        HRT_CUDA(cudaDeviceCanAccessPeer(&can, devices[0], devices[i]));
        if (!can) return fail(HRT_ERR_CUDA, "render_multi: device " + std::to_string(devices[0]) + " cannot map peer memory of device " +
                                                std::to_string(devices[i]));
        cudaError_t pe = cudaDeviceEnablePeerAccess(devices[i], 0);
        if (pe != cudaSuccess && pe != cudaErrorPeerAccessAlreadyEnabled) return cuda_fail(pe, "cudaDeviceEnablePeerAccess");
        cudaGetLastError();
    }
    hrt_stats local[8];
    LaunchSlot* slots[8] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    const int total = rd->sample_count > 0 ? rd->sample_count : rd->samples;
    const int base = rd->sample_count > 0 ? rd->sample_begin : 0;
    // One host thread per device: the wavefront render drives its iterations from the host (hrt_kernels.cu
    // launch_render_wave) and returns when its slice is complete, so the devices only render concurrently when their
    // loops run concurrently.
    int32_t rcs[8];
    std::string errs[8];
    {
        std::vector<std::thread> workers;
        for (int i = 0; i < n; ++i) {
            std::memset(&local[i], 0, sizeof(hrt_stats));
            rcs[i] = HRT_OK;
            workers.emplace_back([&, i]() {
                auto body = [&]() -> int32_t {
                    HRT_CUDA(cudaSetDevice(devices[i]));
                    HRT_CUDA(cudaMemsetAsync(st[i]->d_accum, 0, pixels * 16, 0));
                    hrt_render_desc slice = *rd;
                    const int q = total / n, r = total % n;
                    slice.sample_begin = base + i * q + (i < r ? i : r);
                    slice.sample_count = q + (i < r ? 1 : 0);
                    if (slice.sample_count > 0) {
                        const int32_t rc_i = render_into(s, st[i], cam, &slice, st[i]->d_accum, 0, &local[i], &slots[i]);
                        if (rc_i != HRT_OK) return rc_i;
                    }
                    HRT_CUDA(cudaEventRecord(st[i]->ev[2], 0));  // "my slice is in my accumulator"
                    return HRT_OK;
                };
                rcs[i] = body();
                if (rcs[i] != HRT_OK) errs[i] = hrt_last_error();  // the message is thread-local
            });
        }
        for (auto& w : workers) w.join();
    }
    for (int i = 0; i < n; ++i)
        if (rcs[i] != HRT_OK) return fail(rcs[i], errs[i]);
    HRT_CUDA(cudaSetDevice(devices[0]));
    for (int i = 1; i < n; ++i) HRT_CUDA(cudaStreamWaitEvent(0, st[i]->ev[2], 0));
    HRT_CUDA(cudaEventRecord(st[0]->ev[3], 0));
    cudaError_t e = hrt_fast::launch_reduce_resolve(accums, n, (int)pixels, rd->samples, resolve ? st[0]->d_rgba : nullptr,
                                                    resolve ? nullptr : st[0]->d_rgba, 0);
    if (e != cudaSuccess) return cuda_fail(e, "reduce_resolve_kernel launch");
    HRT_CUDA(cudaEventRecord(st[0]->ev[4], 0));
    HRT_CUDA(cudaMemcpyAsync(out, st[0]->d_rgba, pixels * 16, cudaMemcpyDeviceToHost, 0));
    HRT_CUDA(cudaEventRecord(st[0]->ev[5], 0));
    hrt_stats agg;
    std::memset(&agg, 0, sizeof(agg));
    for (int i = 0; i < n; ++i) {
        HRT_CUDA(cudaSetDevice(devices[i]));
        if ((rc = finish_stats(slots[i], 0, &local[i])) != HRT_OK) return rc;
        agg.paths += local[i].paths;
        agg.rays += local[i].rays;
        agg.launches += local[i].launches;
        if (local[i].kernel_ms > agg.kernel_ms) agg.kernel_ms = local[i].kernel_ms;
        agg.grid = local[i].grid;
        agg.block = local[i].block;
    }
    HRT_CUDA(cudaSetDevice(devices[0]));
    HRT_CUDA(cudaStreamSynchronize(0));
    HRT_CUDA(cudaEventElapsedTime(&agg.resolve_ms, st[0]->ev[3], st[0]->ev[4]));
    HRT_CUDA(cudaEventElapsedTime(&agg.d2h_ms, st[0]->ev[4], st[0]->ev[5]));
    agg.launches += 1;
    if (stats) *stats = agg;
    return HRT_OK;
}

int32_t hrt_render_multi(hrt_scene* s, const int32_t* devices, int32_t n_devices, const hrt_camera_desc* cam,
                         const hrt_render_desc* rd, float* out_rgba, hrt_stats* stats) {
    return render_multi(s, devices, n_devices, cam, rd, out_rgba, stats, true);
}
int32_t hrt_render_accum_multi(hrt_scene* s, const int32_t* devices, int32_t n_devices, const hrt_camera_desc* cam,
                               const hrt_render_desc* rd, float* out_sum, hrt_stats* stats) {
    return render_multi(s, devices, n_devices, cam, rd, out_sum, stats, false);
}

int32_t hrt_render_accum_device(hrt_scene* s, int32_t device, const hrt_camera_desc* cam, const hrt_render_desc* rd,
                                void* d_accum, void* stream_ptr, hrt_stats* stats) {
    if (!d_accum) return fail(HRT_ERR_INVALID, "null device accumulator");
    DeviceState* d = nullptr;
    int32_t rc = get_state(s, device, &d);
    if (rc != HRT_OK) return rc;
    cudaStream_t stream = (cudaStream_t)stream_ptr;
    hrt_stats local;
    std::memset(&local, 0, sizeof(local));
    LaunchSlot* slot = nullptr;
    if ((rc = render_into(s, d, cam, rd, (float*)d_accum, stream, &local, &slot)) != HRT_OK) return rc;
    if (stats) {
        if ((rc = finish_stats(slot, stream, &local)) != HRT_OK) return rc;
        *stats = local;
    }
    return HRT_OK;
}

int32_t hrt_resolve_device(int32_t device, const void* d_accum, int32_t width, int32_t height, int32_t samples,
                           void* d_out_rgba, void* stream_ptr) {
    if (!d_accum || !d_out_rgba || width <= 0 || height <= 0 || samples <= 0)
        return fail(HRT_ERR_INVALID, "resolve: bad argument");
    HRT_CUDA(cudaSetDevice(device));
    cudaError_t e = hrt_fast::launch_resolve((const float*)d_accum, width * height, samples, (float*)d_out_rgba,
                                             (cudaStream_t)stream_ptr);
    if (e != cudaSuccess) return cuda_fail(e, "resolve_kernel launch");
    return HRT_OK;
}

// ---- parity entry points (host buffers in, host buffers out) -------------------------------------------
int32_t hrt_trace_hits(hrt_scene* s, int32_t device, const hrt_ray* rays, int32_t n, const float* xi, hrt_hit* out,
                       uint32_t flags) {
    if (n < 0 || (n > 0 && (!rays || !out))) return fail(HRT_ERR_INVALID, "trace_hits: bad argument");
    DeviceState* d = nullptr;
    int32_t rc = get_state(s, device, &d);
    if (rc != HRT_OK) return rc;
    if (n == 0) return HRT_OK;
    DevBuf<hrt_ray> dr;
    DevBuf<float> dx;
    DevBuf<hrt_hit> dh;
    HRT_CUDA(dr.alloc(n));
    HRT_CUDA(dh.alloc(n));
    HRT_CUDA(cudaMemcpy(dr.p, rays, sizeof(hrt_ray) * (size_t)n, cudaMemcpyHostToDevice));
    if (xi) {
        HRT_CUDA(dx.alloc(n));
        HRT_CUDA(cudaMemcpy(dx.p, xi, sizeof(float) * (size_t)n, cudaMemcpyHostToDevice));
    }
    // the reference form of the stream goes with the reference's box test (as in render_into)
    const DeviceSceneHost& view = d->view[(flags & HRT_FLAG_REFERENCE_TRAVERSAL) ? 0 : 1];
    const int ref = ((flags & HRT_FLAG_REFERENCE_TRAVERSAL) ? 1 : 0) | ((flags & HRT_FLAG_UNIFORM) ? 4 : 0);
    cudaError_t e = (flags & HRT_FLAG_EXACT_MATH) ? hrt_exact::launch_trace_hits(view, dr.p, n, dx.p, dh.p, ref, 0)
                                                  : hrt_fast::launch_trace_hits(view, dr.p, n, dx.p, dh.p, ref, 0);
    if (e != cudaSuccess) return cuda_fail(e, "trace_hits_kernel launch");
    HRT_CUDA(cudaMemcpy(out, dh.p, sizeof(hrt_hit) * (size_t)n, cudaMemcpyDeviceToHost));
    return HRT_OK;
}

int32_t hrt_tex_value(hrt_scene* s, int32_t device, int32_t tex, const float* uvp, int32_t n, float* out, uint32_t flags) {
    if (!s || n < 0 || (n > 0 && (!uvp || !out))) return fail(HRT_ERR_INVALID, "tex_value: bad argument");
    if (tex < 0 || (size_t)tex >= s->textures.size()) return fail(HRT_ERR_INVALID, "tex_value: unknown texture id");
    DeviceState* d = nullptr;
    int32_t rc = get_state(s, device, &d);
    if (rc != HRT_OK) return rc;
    if (n == 0) return HRT_OK;
    DevBuf<float> di, dout;
    HRT_CUDA(di.alloc((size_t)n * 5));
    HRT_CUDA(dout.alloc((size_t)n * 3));
    HRT_CUDA(cudaMemcpy(di.p, uvp, sizeof(float) * 5 * (size_t)n, cudaMemcpyHostToDevice));
    cudaError_t e = (flags & HRT_FLAG_EXACT_MATH) ? hrt_exact::launch_tex_value(d->view[0], tex, di.p, n, dout.p, 0)
                                                  : hrt_fast::launch_tex_value(d->view[0], tex, di.p, n, dout.p, 0);
    if (e != cudaSuccess) return cuda_fail(e, "tex_value_kernel launch");
    HRT_CUDA(cudaMemcpy(out, dout.p, sizeof(float) * 3 * (size_t)n, cudaMemcpyDeviceToHost));
    return HRT_OK;
}

int32_t hrt_scatter(hrt_scene* s, int32_t device, const hrt_ray* rays, const hrt_hit* hits, const float* u4, int32_t n,
                    hrt_scatter_out* out, uint32_t flags) {
    if (!s || n < 0 || (n > 0 && (!rays || !hits || !u4 || !out))) return fail(HRT_ERR_INVALID, "scatter: bad argument");
    for (int i = 0; i < n; ++i)
        if (hits[i].hit && (hits[i].material_id < 0 || (size_t)hits[i].material_id >= s->materials.size()))
            return fail(HRT_ERR_INVALID, "scatter: unknown material id");
    DeviceState* d = nullptr;
    int32_t rc = get_state(s, device, &d);
    if (rc != HRT_OK) return rc;
    if (n == 0) return HRT_OK;
    DevBuf<hrt_ray> dr;
    DevBuf<hrt_hit> dh;
    DevBuf<float> du;
    DevBuf<hrt_scatter_out> dout;
    HRT_CUDA(dr.alloc(n)); HRT_CUDA(dh.alloc(n)); HRT_CUDA(du.alloc((size_t)n * 4)); HRT_CUDA(dout.alloc(n));
    HRT_CUDA(cudaMemcpy(dr.p, rays, sizeof(hrt_ray) * (size_t)n, cudaMemcpyHostToDevice));
    HRT_CUDA(cudaMemcpy(dh.p, hits, sizeof(hrt_hit) * (size_t)n, cudaMemcpyHostToDevice));
    HRT_CUDA(cudaMemcpy(du.p, u4, sizeof(float) * 4 * (size_t)n, cudaMemcpyHostToDevice));
    cudaError_t e = (flags & HRT_FLAG_EXACT_MATH) ? hrt_exact::launch_scatter(d->view[0], dr.p, dh.p, du.p, n, dout.p, 0)
                                                  : hrt_fast::launch_scatter(d->view[0], dr.p, dh.p, du.p, n, dout.p, 0);
    if (e != cudaSuccess) return cuda_fail(e, "scatter_kernel launch");
    HRT_CUDA(cudaMemcpy(out, dout.p, sizeof(hrt_scatter_out) * (size_t)n, cudaMemcpyDeviceToHost));
    return HRT_OK;
}

int32_t hrt_camera_rays(int32_t device, const hrt_camera_desc* cam, const float* stuuu, int32_t n, hrt_ray* out,
                        uint32_t flags) {
    if (!cam || n < 0 || (n > 0 && (!stuuu || !out))) return fail(HRT_ERR_INVALID, "camera_rays: bad argument");
    int cnt = 0;
    cudaError_t ce = cudaGetDeviceCount(&cnt);
    if (ce != cudaSuccess || cnt == 0) return fail(HRT_ERR_CUDA, "no CUDA device available (there is no CPU fallback)");
    HRT_CUDA(cudaSetDevice(device));
    hrt_camera_state st;
    int32_t rc = hrt_camera_init(cam, &st);
    if (rc != HRT_OK) return rc;
    if (n == 0) return HRT_OK;
    DevBuf<float> di;
    DevBuf<hrt_ray> dout;
    HRT_CUDA(di.alloc((size_t)n * 5));
    HRT_CUDA(dout.alloc(n));
    HRT_CUDA(cudaMemcpy(di.p, stuuu, sizeof(float) * 5 * (size_t)n, cudaMemcpyHostToDevice));
    cudaError_t e = (flags & HRT_FLAG_EXACT_MATH) ? hrt_exact::launch_camera_rays(st, di.p, n, dout.p, 0)
                                                  : hrt_fast::launch_camera_rays(st, di.p, n, dout.p, 0);
    if (e != cudaSuccess) return cuda_fail(e, "camera_rays_kernel launch");
    HRT_CUDA(cudaMemcpy(out, dout.p, sizeof(hrt_ray) * (size_t)n, cudaMemcpyDeviceToHost));
    return HRT_OK;
}

int32_t hrt_philox_uniforms(uint64_t seed, uint32_t pixel, uint32_t sample, uint32_t bounce, uint32_t block, float out4[4]) {
    if (!out4) return fail(HRT_ERR_INVALID, "null output");
    hrt_fast::philox_uniforms(seed, pixel, sample, bounce, block, out4);
    return HRT_OK;
}

}  // extern "C"
