// hrt_device.cuh — device-side restatement of the reference hot path for sm_100a.
//
// Compiled TWICE from hrt_kernels.cu:
//   * HRT_EXACT=1, nvcc --fmad=false  -> namespace hrt_exact : IEEE f32 in the reference's operation order,
//     accurate libm.  t / point / normal / front_face come out bit-identical to the CPU oracle; used by the
//     parity entry points (hrt_trace_hits & co.) and by HRT_FLAG_EXACT_MATH renders.
//   * HRT_EXACT=0, default fmad      -> namespace hrt_fast  : FMA contraction, reciprocal multiplies,
//     fast intrinsics where the image tolerance allows.  This is the production render path.
//
// Reference citations are /root/reference paths.
#pragma once
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <math_constants.h>
#include <stdint.h>

#include "hrt_types.h"

#ifndef HRT_EXACT
#error "define HRT_EXACT to 0 or 1"
#endif
#if HRT_EXACT
#define HRT_NS hrt_exact
#else
#define HRT_NS hrt_fast
#endif

namespace HRT_NS {

using namespace hrt;

#define HRT_PI 3.14159265358979323846f

struct V3 {
    float x, y, z;
};
__device__ __forceinline__ V3 v3(float x, float y, float z) { return V3{x, y, z}; }
__device__ __forceinline__ V3 operator+(V3 a, V3 b) { return v3(a.x + b.x, a.y + b.y, a.z + b.z); }
__device__ __forceinline__ V3 operator-(V3 a, V3 b) { return v3(a.x - b.x, a.y - b.y, a.z - b.z); }
__device__ __forceinline__ V3 operator-(V3 a) { return v3(-a.x, -a.y, -a.z); }
__device__ __forceinline__ V3 operator*(float s, V3 a) { return v3(s * a.x, s * a.y, s * a.z); }
__device__ __forceinline__ V3 operator*(V3 a, float s) { return v3(a.x * s, a.y * s, a.z * s); }
__device__ __forceinline__ V3 operator*(V3 a, V3 b) { return v3(a.x * b.x, a.y * b.y, a.z * b.z); }
__device__ __forceinline__ V3 operator/(V3 a, float s) { return v3(a.x / s, a.y / s, a.z / s); }
// cgmath 0.18 dot: (x*x + y*y) + z*z
__device__ __forceinline__ float dot(V3 a, V3 b) { return (a.x * b.x + a.y * b.y) + a.z * b.z; }
__device__ __forceinline__ float dot_rn(V3 a, V3 b);
__device__ __forceinline__ float length(V3 a) { return sqrtf(dot(a, a)); }
// cgmath normalize: v * (1 / |v|)
__device__ __forceinline__ V3 normalize(V3 a) {
#if HRT_EXACT
    return a * (1.0f / length(a));
#else
    return a * rsqrtf(dot(a, a));
#endif
}
__device__ __forceinline__ float comp(V3 a, int i) { return i == 0 ? a.x : (i == 1 ? a.y : a.z); }

struct Ray {
    V3 o, d;
    float time;
};
__device__ __forceinline__ V3 ray_at(const Ray& r, float t) { return r.o + t * r.d; }  // ray.rs:25-27

struct DeviceScene {
    const float4* ops;     // 2 x float4 per record
    const uint4* nodes;    // 2 x uint4 per OP_BVH tree node (hrt_types.h Bvh2Node)
    const Ctx* ctxs;
    const Material* mats;
    const Texture* texs;
    const NoiseTable* noise;
    cudaTextureObject_t images[kMaxImages];
    int32_t n_ops, n_noise, n_media;
    float ln_e;            // logf(E_f32) as computed by the host libm (f32::log(self, E) = ln(x)/ln(E))
};

// ------------------------------------------------------------------------------------------------
// Philox4x32-10 counter-based RNG (Salmon et al. 2011).  key = render seed; counter =
// (pixel, sample, bounce<<8 | block, stream).  uniform = top 24 bits * 2^-24 in [0,1) — the same
// distribution as rand 0.8.5's gen::<f32>().
// ------------------------------------------------------------------------------------------------
__host__ __device__ __forceinline__ uint32_t hrt_mulhi32(uint32_t a, uint32_t b) {
#ifdef __CUDA_ARCH__
    return __umulhi(a, b);
#else
    return (uint32_t)(((uint64_t)a * (uint64_t)b) >> 32);
#endif
}
struct U4 {
    uint32_t x, y, z, w;
};
__host__ __device__ __forceinline__ U4 philox4x32_10(U4 c, uint32_t k0, uint32_t k1) {
    const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
    for (int i = 0; i < 10; ++i) {
        uint32_t hi0 = hrt_mulhi32(M0, c.x), lo0 = M0 * c.x;
        uint32_t hi1 = hrt_mulhi32(M1, c.z), lo1 = M1 * c.z;
        U4 n;
        n.x = hi1 ^ c.y ^ k0;
        n.y = lo1;
        n.z = hi0 ^ c.w ^ k1;
        n.w = lo0;
        c = n;
        k0 += W0;
        k1 += W1;
    }
    return c;
}
__host__ __device__ __forceinline__ float u32_to_unit(uint32_t x) { return (float)(x >> 8) * (1.0f / 16777216.0f); }

enum { RNG_BLOCK_SCATTER = 0, RNG_BLOCK_MEDIA0 = 1, RNG_BLOCK_CAMERA = 255 };
struct RngKey {
    uint32_t k0, k1, pixel, sample;
};
#ifdef __CUDACC__
// One out-of-line copy per kernel: the 10 unrolled rounds are ~100 instructions and there are four call sites; inlining
// them all bloats the hot loop past the instruction cache (profiles/r01_v1_render_kernel.txt: stall_no_instruction).
__device__ __noinline__ uint4 philox_call(uint32_t pixel, uint32_t sample, uint32_t z, uint32_t k0, uint32_t k1) {
    U4 c;
    c.x = pixel; c.y = sample; c.z = z; c.w = 0x68727421u;
    U4 r = philox4x32_10(c, k0, k1);
    return make_uint4(r.x, r.y, r.z, r.w);
}
#endif
__host__ __device__ __forceinline__ void rng_block(const RngKey& k, uint32_t bounce, uint32_t block, float out[4]) {
#ifdef __CUDA_ARCH__
    const uint4 q = philox_call(k.pixel, k.sample, (bounce << 8) | block, k.k0, k.k1);
    U4 r;
    r.x = q.x; r.y = q.y; r.z = q.z; r.w = q.w;
#else
    U4 c;
    c.x = k.pixel; c.y = k.sample; c.z = (bounce << 8) | block; c.w = 0x68727421u;
    U4 r = philox4x32_10(c, k.k0, k.k1);
#endif
    out[0] = u32_to_unit(r.x); out[1] = u32_to_unit(r.y); out[2] = u32_to_unit(r.z); out[3] = u32_to_unit(r.w);
}

// Source of the one uniform a ConstantMedium draws (constant_medium.rs:59).
struct MediumXi {
    RngKey key;
    uint32_t bounce;
    float injected;   // used when inject
    bool inject;
    __device__ __forceinline__ float draw(int medium_index) const {
        if (inject) return injected;
        float u[4];
        rng_block(key, bounce, RNG_BLOCK_MEDIA0 + ((uint32_t)medium_index >> 2), u);
        int j = medium_index & 3;
        return j == 0 ? u[0] : (j == 1 ? u[1] : (j == 2 ? u[2] : u[3]));
    }
};

// ------------------------------------------------------------------------------------------------
// Fixed-draw samplers of the distributions the reference rejection-samples (math.rs:12-40):
// uniform on S^2, uniform in the unit ball, uniform in the unit disk.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void sincos_2pi(float u, float& s, float& c) {
    float phi = (2.0f * HRT_PI) * u;
#if HRT_EXACT
    s = sinf(phi);
    c = cosf(phi);
#else
    __sincosf(phi, &s, &c);
#endif
}
__device__ __forceinline__ V3 sample_unit_vector(float u1, float u2) {
    float z = 1.0f - 2.0f * u1;
    float r = sqrtf(fmaxf(0.0f, 1.0f - z * z));
    float s, c;
    sincos_2pi(u2, s, c);
    return v3(r * c, r * s, z);
}
__device__ __forceinline__ V3 sample_in_unit_sphere(float u1, float u2, float u3) {
    return cbrtf(u3) * sample_unit_vector(u1, u2);
}
__device__ __forceinline__ V3 sample_in_unit_disk(float u1, float u2) {
    float r = sqrtf(u1);
    float s, c;
    sincos_2pi(u2, s, c);
    return v3(r * c, r * s, 0.0f);
}

// ------------------------------------------------------------------------------------------------
// Ray-space contexts
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void load_op(const DeviceScene& S, int pc, float4& A, float4& B) {
    A = __ldg(S.ops + 2 * pc);
    B = __ldg(S.ops + 2 * pc + 1);
}
__device__ __forceinline__ void load_node(const DeviceScene& S, int node, uint4& L, uint4& R) {
    L = __ldg(S.nodes + 2 * (size_t)node);
    R = __ldg(S.nodes + 2 * (size_t)node + 1);
}

// translation.rs:25-29
__device__ __forceinline__ void apply_translate(Ray& r, float4 A) { r.o = r.o - v3(A.x, A.y, A.z); }
// rotation.rs:103-116.  (r,a,b): X -> (0,1,2), Y -> (1,2,0), Z -> (2,0,1).  Branch-free component selection keeps each
// inlined copy small (the three-way branchy version was 734 instructions of the render kernel).
__device__ __forceinline__ float sel3(int i, float x, float y, float z) { return i == 0 ? x : (i == 1 ? y : z); }
__device__ __forceinline__ V3 put2(V3 p, int ia, float va, int ib, float vb) {
    p.x = ia == 0 ? va : (ib == 0 ? vb : p.x);
    p.y = ia == 1 ? va : (ib == 1 ? vb : p.y);
    p.z = ia == 2 ? va : (ib == 2 ? vb : p.z);
    return p;
}
__device__ __forceinline__ void apply_rotate(Ray& r, float4 A) {
    const float sn = A.x, cs = A.y;
    const int axis = __float_as_int(A.z);
    const int ia = axis == 2 ? 0 : axis + 1, ib = axis == 0 ? 2 : axis - 1;
    const float oa = sel3(ia, r.o.x, r.o.y, r.o.z), ob = sel3(ib, r.o.x, r.o.y, r.o.z);
    const float da = sel3(ia, r.d.x, r.d.y, r.d.z), db = sel3(ib, r.d.x, r.d.y, r.d.z);
    // individually rounded in both builds: with |o| ~ 1e3 an FMA-contracted rotation moves t by ~1e-4 relative
    const float noa = __fadd_rn(__fmul_rn(cs, oa), __fmul_rn(sn, ob)), nob = __fadd_rn(__fmul_rn(-sn, oa), __fmul_rn(cs, ob));
    const float nda = __fadd_rn(__fmul_rn(cs, da), __fmul_rn(sn, db)), ndb = __fadd_rn(__fmul_rn(-sn, da), __fmul_rn(cs, db));
    r.o = put2(r.o, ia, noa, ib, nob);
    r.d = put2(r.d, ia, nda, ib, ndb);
}
// rotation.rs:119-131 (object -> parent space for a point or a normal)
__device__ __forceinline__ V3 unrotate(V3 p, float4 A) {
    const float sn = A.x, cs = A.y;
    const int axis = __float_as_int(A.z);
    const int ia = axis == 2 ? 0 : axis + 1, ib = axis == 0 ? 2 : axis - 1;
    const float pa = sel3(ia, p.x, p.y, p.z), pb = sel3(ib, p.x, p.y, p.z);
    const float na = __fsub_rn(__fmul_rn(cs, pa), __fmul_rn(sn, pb)), nb = __fadd_rn(__fmul_rn(sn, pa), __fmul_rn(cs, pb));
    return put2(p, ia, na, ib, nb);
}
// Map the world ray into context `ctx` by replaying its push records, outermost first.
__device__ __noinline__ Ray ray_in_ctx(const DeviceScene& S, const Ray& world, int ctx) {
    Ray r = world;
    if (ctx == 0) return r;
    const Ctx c = S.ctxs[ctx];
#pragma unroll 1
    for (int i = 0; i < c.depth; ++i) {
        float4 A, B;
        load_op(S, c.op_pc[i], A, B);
        if ((__float_as_uint(B.w) & 0xffu) == OP_TRANSLATE) apply_translate(r, A);
        else apply_rotate(r, A);
    }
    return r;
}

// Per-context ray constants.
struct RayK {
    V3 inv;    // 1/d per component (aabb.rs:22)
    float dd;  // d·d (sphere.rs:42)
};
__device__ __forceinline__ float fast_rcp(float x) {
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}
// Out of line: three IEEE divisions (each with its slow path) were inlined at nine sites of the render kernel.
__device__ __noinline__ float4 rayk_call(float dx, float dy, float dz) {
    // __frcp_rn is the correctly rounded reciprocal == the IEEE quotient 1.0f / d (aabb.rs:22) without the generic
    // division's slow path
#if HRT_EXACT
    return make_float4(__frcp_rn(dx), __frcp_rn(dy), __frcp_rn(dz),
#else
    // production build: MUFU.RCP (1 ulp).  The reciprocals only feed box culling and the rect plane distance, both of
    // which already differ from the reference by an ulp through FMA contraction.
    return make_float4(fast_rcp(dx), fast_rcp(dy), fast_rcp(dz),
#endif
                       __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz)));
}
__device__ __forceinline__ RayK make_rayk(const Ray& r) {
    const float4 q = rayk_call(r.d.x, r.d.y, r.d.z);
    RayK k;
    k.inv = v3(q.x, q.y, q.z);
    k.dd = q.w;
    return k;
}

// ------------------------------------------------------------------------------------------------
// Box tests
// ------------------------------------------------------------------------------------------------
// aabb.rs:20-47, verbatim semantics: per-axis interval vs (tmin,tmax), the three never intersected (Q1).
__device__ __forceinline__ bool box_hit_reference(float4 A, float4 B, const Ray& r, const RayK& k, float tmin, float tmax) {
    bool ok = true;
    {
        float t0 = (A.x - r.o.x) * k.inv.x, t1 = (B.x - r.o.x) * k.inv.x;
        float lo = k.inv.x < 0.0f ? t1 : t0, hi = k.inv.x < 0.0f ? t0 : t1;
        float a = lo > tmin ? lo : tmin, b = hi < tmax ? hi : tmax;
        ok = ok && !(b <= a);
    }
    {
        float t0 = (A.y - r.o.y) * k.inv.y, t1 = (B.y - r.o.y) * k.inv.y;
        float lo = k.inv.y < 0.0f ? t1 : t0, hi = k.inv.y < 0.0f ? t0 : t1;
        float a = lo > tmin ? lo : tmin, b = hi < tmax ? hi : tmax;
        ok = ok && !(b <= a);
    }
    {
        float t0 = (A.z - r.o.z) * k.inv.z, t1 = (B.z - r.o.z) * k.inv.z;
        float lo = k.inv.z < 0.0f ? t1 : t0, hi = k.inv.z < 0.0f ? t0 : t1;
        float a = lo > tmin ? lo : tmin, b = hi < tmax ? hi : tmax;
        ok = ok && !(b <= a);
    }
    return ok;
}
// Intersected slab test: same per-axis arithmetic and NaN behaviour, intervals intersected across axes.
// On a box that contains its contents it can never cull a hit the reference test would let through.
__device__ __forceinline__ bool box_hit_tight(float4 A, float4 B, const Ray& r, const RayK& k, float tmin, float tmax) {
    float t0x = (A.x - r.o.x) * k.inv.x, t1x = (B.x - r.o.x) * k.inv.x;
    float t0y = (A.y - r.o.y) * k.inv.y, t1y = (B.y - r.o.y) * k.inv.y;
    float t0z = (A.z - r.o.z) * k.inv.z, t1z = (B.z - r.o.z) * k.inv.z;
    float lox = k.inv.x < 0.0f ? t1x : t0x, hix = k.inv.x < 0.0f ? t0x : t1x;
    float loy = k.inv.y < 0.0f ? t1y : t0y, hiy = k.inv.y < 0.0f ? t0y : t1y;
    float loz = k.inv.z < 0.0f ? t1z : t0z, hiz = k.inv.z < 0.0f ? t0z : t1z;
    // `x > m ? x : m` == fmaxf(x, m) including NaN x (keeps m)
    float lo = fmaxf(fmaxf(lox, loy), fmaxf(loz, tmin));
    float hi = fminf(fminf(hix, hiy), fminf(hiz, tmax));
    return !(hi <= lo);
}

// ------------------------------------------------------------------------------------------------
// Primitive tests (closest-hit form: accept tmin <= t <= closest, as the reference's range checks do)
// ------------------------------------------------------------------------------------------------
// sphere.rs:41-58 / moving_sphere.rs:62-79.  NaN discriminants/roots are accepted exactly as the reference
// accepts them (Q15).
__device__ __forceinline__ float dot_rn(V3 a, V3 b) {  // (x*x + y*y) + z*z with every op rounded (never contracted)
    return __fadd_rn(__fadd_rn(__fmul_rn(a.x, b.x), __fmul_rn(a.y, b.y)), __fmul_rn(a.z, b.z));
}
__device__ __forceinline__ bool sphere_test(V3 center, float radius, const Ray& r, const RayK& k, float tmin, float closest,
                                            float& t_out) {
    // The quadratic's coefficients are cancellation-prone (r = 1000 ground sphere, r = 5000 fog boundary, origins that
    // lie ON the sphere): an FMA-contracted oc.oc - r^2 moves t by ~1e-5 relative.  They are therefore computed with
    // individually rounded operations in BOTH builds (6 extra instructions), which keeps the production build's t within
    // an ulp or two of the reference's.
    V3 oc = v3(__fsub_rn(r.o.x, center.x), __fsub_rn(r.o.y, center.y), __fsub_rn(r.o.z, center.z));
    float a = k.dd;
    float half_b = dot_rn(oc, r.d);
    float c = __fsub_rn(dot_rn(oc, oc), __fmul_rn(radius, radius));
    float disc = __fsub_rn(__fmul_rn(half_b, half_b), __fmul_rn(a, c));
    if (disc < 0.0f) return false;
    float sqrtd = sqrtf(disc);
    // true divisions in both builds (only reached when the discriminant is non-negative): a reciprocal multiply is
    // an ulp off, which is enough to flip `root < t_min` for rays that start on the sphere
    float root = __fdiv_rn(-half_b - sqrtd, a);
    if (root < tmin || closest < root) {
        root = __fdiv_rn(-half_b + sqrtd, a);
        if (root < tmin || closest < root) return false;
    }
    t_out = root;
    return true;
}
// moving_sphere.rs:53-57
__device__ __forceinline__ V3 msphere_center(V3 c0, V3 c1, float t0, float t1, float time) {
    const float f = __fdiv_rn(time - t0, t1 - t0);  // individually rounded in both builds (centre ~ 1e2..1e3 units)
    return v3(__fadd_rn(c0.x, __fmul_rn(f, c1.x - c0.x)), __fadd_rn(c0.y, __fmul_rn(f, c1.y - c0.y)),
              __fadd_rn(c0.z, __fmul_rn(f, c1.z - c0.z)));
}
// rect.rs:60-69 with (k,a,b) already resolved to scalars
__device__ __forceinline__ bool rect_test(float ok, float dk, float invk, float oa, float da, float ob, float db, float a0,
                                          float a1, float b0, float b1, float kk, float tmin, float closest, float& t_out) {
#if HRT_EXACT
    float t = (kk - ok) / dk;
#else
    float t = (kk - ok) * invk;
#endif
    if (t < tmin || t > closest) return false;
    float a = oa + t * da;
    float b = ob + t * db;
    if (a < a0 || a > a1 || b < b0 || b > b1) return false;
    t_out = t;
    return true;
}
// rect.rs:54-69 for any plane: (k, a, b) = XY (z, x, y), YZ (x, y, z), ZX (y, z, x)
__device__ __forceinline__ bool rect_any(uint32_t opc, float4 A, float kk, const Ray& c, const RayK& k, float tmin, float closest,
                                         float& t) {
    const int ik = opc == OP_RECT_XY ? 2 : (opc == OP_RECT_YZ ? 0 : 1);
    const int ia = ik == 2 ? 0 : ik + 1, ib = ik == 0 ? 2 : ik - 1;
    return rect_test(sel3(ik, c.o.x, c.o.y, c.o.z), sel3(ik, c.d.x, c.d.y, c.d.z), sel3(ik, k.inv.x, k.inv.y, k.inv.z),
                     sel3(ia, c.o.x, c.o.y, c.o.z), sel3(ia, c.d.x, c.d.y, c.d.z), sel3(ib, c.o.x, c.o.y, c.o.z),
                     sel3(ib, c.d.x, c.d.y, c.d.z), A.x, A.y, A.z, A.w, kk, tmin, closest, t);
}

// cuboid.rs:30-96 + list.rs:20-31: six rects in construction order with closest-so-far narrowing:
//   0: XY @ max.z   1: XY @ min.z   2: ZX @ max.y   3: ZX @ min.y   4: YZ @ max.x   5: YZ @ min.x
// i.e. (k, a, b) = (z, x, y), (y, z, x), (x, y, z): each axis step rotates the components by one.  A rolled loop: the
// unrolled form is 6 inlined rect tests per call site, and code size is what bounds the render kernel.
__device__ __forceinline__ bool cuboid_test(V3 mn, V3 mx, const Ray& r, const RayK& k, float tmin, float closest,
                                            float& t_out, int& face_out) {
    float ok = r.o.z, dk = r.d.z, ik = k.inv.z, oa = r.o.x, da = r.d.x, ob = r.o.y, db = r.d.y;
    float k0 = mn.z, k1 = mx.z, a0 = mn.x, a1 = mx.x, b0 = mn.y, b1 = mx.y;
    float ia = k.inv.x, ib = k.inv.y;
    bool any = false;
#pragma unroll 1
    for (int axis = 0; axis < 3; ++axis) {
        float t;
        if (rect_test(ok, dk, ik, oa, da, ob, db, a0, a1, b0, b1, k1, tmin, closest, t)) { closest = t; face_out = 2 * axis; any = true; }
        if (rect_test(ok, dk, ik, oa, da, ob, db, a0, a1, b0, b1, k0, tmin, closest, t)) { closest = t; face_out = 2 * axis + 1; any = true; }
        // next axis: k <- b, a <- k, b <- a
        float x;
        x = ob; ob = oa; oa = ok; ok = x;
        x = db; db = da; da = dk; dk = x;
        x = ib; ib = ia; ia = ik; ik = x;
        x = b0; b0 = a0; a0 = k0; k0 = x;
        x = b1; b1 = a1; a1 = k1; k1 = x;
    }
    t_out = closest;
    return any;
}

// ------------------------------------------------------------------------------------------------
// OP_BVH: a sound BvhNode as a two-child tree (hrt_types.h Bvh2Node), walked per ray with a stack, nearer child first.
// ------------------------------------------------------------------------------------------------
// Entry / exit distances of the ray against an fp16 box packed as three half2 words (min.x min.y | min.z max.x |
// max.y max.z); same per-axis arithmetic and NaN behaviour as box_hit_tight.  The box is missed iff hi < lo: unlike
// box_hit_tight (hi <= lo) a zero-width overlap still counts, so a primitive whose hit distance EQUALS the current
// closest is always reached and the tie rule below decides, not the visit order.
__device__ __forceinline__ void slab16(uint32_t w0, uint32_t w1, uint32_t w2, const Ray& r, const RayK& k, float tmin, float closest,
                                       float& lo, float& hi) {
    const float2 a = __half22float2(*reinterpret_cast<const __half2*>(&w0));  // min.x, min.y
    const float2 b = __half22float2(*reinterpret_cast<const __half2*>(&w1));  // min.z, max.x
    const float2 c = __half22float2(*reinterpret_cast<const __half2*>(&w2));  // max.y, max.z
    const float t0x = (a.x - r.o.x) * k.inv.x, t1x = (b.y - r.o.x) * k.inv.x;
    const float t0y = (a.y - r.o.y) * k.inv.y, t1y = (c.x - r.o.y) * k.inv.y;
    const float t0z = (b.x - r.o.z) * k.inv.z, t1z = (c.y - r.o.z) * k.inv.z;
    const float lox = k.inv.x < 0.0f ? t1x : t0x, hix = k.inv.x < 0.0f ? t0x : t1x;
    const float loy = k.inv.y < 0.0f ? t1y : t0y, hiy = k.inv.y < 0.0f ? t0y : t1y;
    const float loz = k.inv.z < 0.0f ? t1z : t0z, hiz = k.inv.z < 0.0f ? t0z : t1z;
    lo = fmaxf(fmaxf(lox, loy), fmaxf(loz, tmin));
    hi = fminf(fminf(hix, hiy), fminf(hiz, closest));
}

struct TreeHit {
    float t;   // closest so far (in: the running t_max; out: narrowed)
    int pc;    // record of the closest hit so far (in: the caller's, -1 none; out: unchanged or a leaf of this tree)
    int face;  // cuboid side when pc changed
};

// An EXACT tie between two surfaces (same t) inside an OP_BVH tree, decided as the reference decides it.  The reference
// visits a BvhNode's leaves in ITS depth-first order with a running t_max (bvh_node.rs:110-124): the earlier leaf is hit
// first, t_max becomes t, and the later leaf replaces it (`t <= t_max`) only if it is still REACHED — i.e. if its
// bounding box passes Aabb::hit with t_max = t (aabb.rs:20-47: rejected when `min(t1, t_max) <= max(t0, t_min)` on an
// axis, which is exactly what happens when the ray ENTERS that box at the tie point, as on the shared face of two
// adjacent cuboids seen from inside one of them).  The leaf records keep the reference's order, so "later" is the larger
// pc; this evaluates the reference's test on the later leaf's own reference box (sphere.rs:77-83,
// moving_sphere.rs:98-110 over the BvhNode's time interval, rect.rs:88-103, cuboid.rs:104-106).  Cold: ties only.
__device__ __noinline__ bool tie_goes_to_later(const DeviceScene& S, int later_pc, Ray cur, float tmin, float t, float ts, float te) {
    const RayK k = make_rayk(cur);
    float4 A, B;
    load_op(S, later_pc, A, B);
    const uint32_t opc = __float_as_uint(B.w) & 0xffu;
    float4 mn, mx;
    mn.w = mx.w = 0.0f;
    if (opc == OP_SPHERE) {
        mn.x = A.x - A.w; mn.y = A.y - A.w; mn.z = A.z - A.w;
        mx.x = A.x + A.w; mx.y = A.y + A.w; mx.z = A.z + A.w;
    } else if (opc == OP_MSPHERE) {
        float4 C, D;
        load_op(S, later_pc + 1, C, D);
        const V3 ca = msphere_center(v3(A.x, A.y, A.z), v3(C.x, C.y, C.z), C.w, D.x, ts);
        const V3 cb = msphere_center(v3(A.x, A.y, A.z), v3(C.x, C.y, C.z), C.w, D.x, te);
        mn.x = fminf(ca.x - A.w, cb.x - A.w); mn.y = fminf(ca.y - A.w, cb.y - A.w); mn.z = fminf(ca.z - A.w, cb.z - A.w);
        mx.x = fmaxf(ca.x + A.w, cb.x + A.w); mx.y = fmaxf(ca.y + A.w, cb.y + A.w); mx.z = fmaxf(ca.z + A.w, cb.z + A.w);
    } else if (opc == OP_CUBOID) {
        mn = A; mx = B;
    } else {
        const float lo = B.x - 0.0001f, hi = B.x + 0.0001f;
        if (opc == OP_RECT_XY) { mn.x = A.x; mn.y = A.z; mn.z = lo; mx.x = A.y; mx.y = A.w; mx.z = hi; }
        else if (opc == OP_RECT_YZ) { mn.x = lo; mn.y = A.x; mn.z = A.z; mx.x = hi; mx.y = A.y; mx.z = A.w; }
        else { mn.x = A.x; mn.y = lo; mn.z = A.z; mx.x = A.y; mx.y = hi; mx.z = A.w; }  // rect.rs:98-101 as written
    }
    return box_hit_reference(mn, mx, cur, k, tmin, t);
}

// The walk of the tree whose first node is `base`, one step at a time.  `ref` >= 0: an inner node (relative to base),
// < 0: ~pc of a leaf record; start with ref = 0, sp = 0.  Box tests only prune (every box is sound), so any visit order
// finds the same closest t; exact ties are settled by tie_goes_to_later.  (ts, te): the BvhNode's time interval.
//
// Next postponed child that can still hold a closer or equal hit; true when the walk is over.
__device__ __forceinline__ bool bvh2_pop(int& ref, int& sp, const int* stack_ref, const float* stack_t, const TreeHit& h) {
    do {
        if (sp == 0) return true;
        --sp;
        ref = stack_ref[sp];
    } while (stack_t[sp] > h.t);
    return false;
}
// Inner node: both children's boxes, the nearer one first, the other postponed on the stack.
__device__ __forceinline__ bool bvh2_inner(const DeviceScene& S, int base, const Ray& cur, const RayK& k, float tmin, int& ref, int& sp,
                                           int* stack_ref, float* stack_t, const TreeHit& h) {
    uint4 L, R;
    load_node(S, base + ref, L, R);
    float llo, lhi, rlo, rhi;
    slab16(L.x, L.y, L.z, cur, k, tmin, h.t, llo, lhi);
    slab16(R.x, R.y, R.z, cur, k, tmin, h.t, rlo, rhi);
    const bool hl = !(lhi < llo), hr = !(rhi < rlo);
    if (hl && hr) {
        const bool right_first = rlo < llo;
        stack_ref[sp] = right_first ? (int)L.w : (int)R.w;
        stack_t[sp] = right_first ? llo : rlo;
        sp++;
        ref = right_first ? (int)R.w : (int)L.w;
        return false;
    }
    if (hl || hr) {
        ref = hl ? (int)L.w : (int)R.w;
        return false;
    }
    return bvh2_pop(ref, sp, stack_ref, stack_t, h);
}
// Leaf: a primitive record of the op stream.
__device__ __forceinline__ bool bvh2_leaf(const DeviceScene& S, const Ray& cur, const RayK& k, float tmin, float ts, float te, int& ref,
                                          int& sp, const int* stack_ref, const float* stack_t, TreeHit& h) {
    const int pc = ~ref;
    float4 A, B;
    load_op(S, pc, A, B);
    const uint32_t opc = __float_as_uint(B.w) & 0xffu;
    float t = 0.0f;
    int face = 0;
    bool hit;
    if (opc == OP_SPHERE || opc == OP_MSPHERE) {
        V3 ctr = v3(A.x, A.y, A.z);
        if (opc == OP_MSPHERE) {
            float4 C, D;
            load_op(S, pc + 1, C, D);
            ctr = msphere_center(ctr, v3(C.x, C.y, C.z), C.w, D.x, cur.time);
        }
        hit = sphere_test(ctr, A.w, cur, k, tmin, h.t, t);
    } else if (opc == OP_CUBOID) {
        hit = cuboid_test(v3(A.x, A.y, A.z), v3(B.x, B.y, B.z), cur, k, tmin, h.t, t, face);
    } else {
        hit = rect_any(opc, A, B.x, cur, k, tmin, h.t, t);
    }
    if (hit) {  // the primitive tests accept t <= closest
        bool take = t < h.t || h.pc < 0;
        if (!take && !(t > h.t)) {  // exact tie (or a NaN, Q15) with the hit held so far
            const bool later_wins = tie_goes_to_later(S, pc > h.pc ? pc : h.pc, cur, tmin, t, ts, te);
            take = (pc > h.pc) == later_wins;
        }
        if (take) { h.t = t; h.pc = pc; h.face = face; }
    }
    return bvh2_pop(ref, sp, stack_ref, stack_t, h);
}
__device__ __forceinline__ bool bvh2_step(const DeviceScene& S, int base, const Ray& cur, const RayK& k, float tmin, float ts, float te,
                                          int& ref, int& sp, int* stack_ref, float* stack_t, TreeHit& h) {
    return ref >= 0 ? bvh2_inner(S, base, cur, k, tmin, ref, sp, stack_ref, stack_t, h)
                    : bvh2_leaf(S, cur, k, tmin, ts, te, ref, sp, stack_ref, stack_t, h);
}
// Closest hit of a whole tree for t in [tmin, closest]; `best_pc`: the record of the hit held so far (-1 none).  Out of
// line: one copy per kernel.
__device__ __noinline__ TreeHit bvh2_walk(const DeviceScene& S, int base, Ray cur, float tmin, float closest, int best_pc,
                                          float ts, float te) {
    const RayK k = make_rayk(cur);
    TreeHit h;
    h.t = closest; h.pc = best_pc; h.face = 0;
    int stack_ref[kBvh2Stack];
    float stack_t[kBvh2Stack];
    int sp = 0, ref = 0;
    while (!bvh2_step(S, base, cur, k, tmin, ts, te, ref, sp, stack_ref, stack_t, h)) {}
    return h;
}

// ------------------------------------------------------------------------------------------------
// Traversal: a forward-only interpreter over the op stream.
// ------------------------------------------------------------------------------------------------
struct Best {
    float t;
    int pc;    // record that produced the hit (-1: none)
    int face;  // cuboid side
    int ctx;   // ray space the record lives in
};

// Closest hit of the sub-stream [pc, pc_end) for t in [tmin, closest]; returns true and narrows `closest`
// when something was hit.  kInner: boundary query of a ConstantMedium (only t is needed; media cannot
// nest).  `world` is the world-space ray, `cur`/`cur_ctx` the ray in the space the sub-stream starts in.
template <bool kInner>
__device__ __forceinline__ bool traverse(const DeviceScene& S, int pc, const int pc_end, const Ray& world, Ray cur,
                                         int cur_ctx, const float tmin, float& closest, Best& best,
                                         const bool reference_boxes, const MediumXi& xi) {
    RayK k = make_rayk(cur);
    bool any = false;
    while (pc < pc_end) {
        float4 A, B;
        load_op(S, pc, A, B);
        const uint32_t w7 = __float_as_uint(B.w);
        const uint32_t opc = w7 & 0xffu;
        if (opc == OP_BOX || opc == OP_BOX_LOOSE) {
            bool hit = (opc == OP_BOX_LOOSE || reference_boxes) ? box_hit_reference(A, B, cur, k, tmin, closest)
                                                                : box_hit_tight(A, B, cur, k, tmin, closest);
            pc = hit ? pc + 1 : (int)(w7 >> 8);
            continue;
        }
        switch (opc) {
            case OP_SPHERE: {
                float t;
                if (sphere_test(v3(A.x, A.y, A.z), A.w, cur, k, tmin, closest, t)) {
                    closest = t; any = true;
                    if (!kInner) { best.t = t; best.pc = pc; best.face = 0; best.ctx = cur_ctx; }
                }
                pc += 1;
                break;
            }
            case OP_MSPHERE: {
                float4 C, D;
                load_op(S, pc + 1, C, D);
                V3 ctr = msphere_center(v3(A.x, A.y, A.z), v3(C.x, C.y, C.z), C.w, D.x, cur.time);
                float t;
                if (sphere_test(ctr, A.w, cur, k, tmin, closest, t)) {
                    closest = t; any = true;
                    if (!kInner) { best.t = t; best.pc = pc; best.face = 0; best.ctx = cur_ctx; }
                }
                pc += 2;
                break;
            }
            case OP_RECT_XY: case OP_RECT_YZ: case OP_RECT_ZX: {
                float t;
                bool h;
                if (opc == OP_RECT_XY) h = rect_test(cur.o.z, cur.d.z, k.inv.z, cur.o.x, cur.d.x, cur.o.y, cur.d.y, A.x, A.y, A.z, A.w, B.x, tmin, closest, t);
                else if (opc == OP_RECT_YZ) h = rect_test(cur.o.x, cur.d.x, k.inv.x, cur.o.y, cur.d.y, cur.o.z, cur.d.z, A.x, A.y, A.z, A.w, B.x, tmin, closest, t);
                else h = rect_test(cur.o.y, cur.d.y, k.inv.y, cur.o.z, cur.d.z, cur.o.x, cur.d.x, A.x, A.y, A.z, A.w, B.x, tmin, closest, t);
                if (h) {
                    closest = t; any = true;
                    if (!kInner) { best.t = t; best.pc = pc; best.face = 0; best.ctx = cur_ctx; }
                }
                pc += 1;
                break;
            }
            case OP_CUBOID: {
                float t;
                int face = 0;
                if (cuboid_test(v3(A.x, A.y, A.z), v3(B.x, B.y, B.z), cur, k, tmin, closest, t, face)) {
                    closest = t; any = true;
                    if (!kInner) { best.t = t; best.pc = pc; best.face = face; best.ctx = cur_ctx; }
                }
                pc += 1;
                break;
            }
            case OP_TRANSLATE: case OP_ROTATE: case OP_POP: {
                // enter / leave a (run of) ray space(s): map the world ray through the target context's push records
                cur_ctx = __float_as_int(A.w);
                cur = ray_in_ctx(S, world, cur_ctx);
                k = make_rayk(cur);
                const int run = (int)(w7 >> 8);
                pc += run > 0 ? run : 1;
                break;
            }
            case OP_BVH: {
                const TreeHit th = bvh2_walk(S, __float_as_int(A.x), cur, tmin, closest, kInner ? -1 : best.pc, B.x, B.y);
                if (th.pc != (kInner ? -1 : best.pc)) {
                    closest = th.t; any = true;
                    if (!kInner) { best.t = th.t; best.pc = th.pc; best.face = th.face; best.ctx = cur_ctx; }
                }
                pc = (int)(w7 >> 8);
                break;
            }
            case OP_MEDIUM: case OP_MEDIUM_SPHERE: case OP_MEDIUM_CUBOID: {
                const int end = (int)(w7 >> 8);
                if (!kInner) {
                    // constant_medium.rs:34-76
                    Best dummy;
                    float t1 = CUDART_INF_F;
                    bool h1 = traverse<true>(S, pc + 1, end, world, cur, cur_ctx, -CUDART_INF_F, t1, dummy, reference_boxes, xi);
                    if (h1) {
                        float t2 = CUDART_INF_F;
                        bool h2 = traverse<true>(S, pc + 1, end, world, cur, cur_ctx, t1 + 0.0001f, t2, dummy, reference_boxes, xi);
                        if (h2) {
                            if (t1 < tmin) t1 = tmin;
                            if (t2 > closest) t2 = closest;
                            if (!(t1 >= t2)) {
                                if (t1 < 0.0f) t1 = 0.0f;
                                float ray_length = sqrtf(k.dd);
                                float dist_inside = (t2 - t1) * ray_length;
                                float u = xi.draw(__float_as_int(A.z));
#if HRT_EXACT
                                float hit_distance = A.x * (logf(u) / S.ln_e);
                                float t = t1 + hit_distance / ray_length;
#else
                                // accurate logf here too: __logf's absolute error near u = 1 is a relative error of up to
                                // ~1e-3 in the free-flight distance; one call per medium query is noise in the profile
                                float hit_distance = A.x * logf(u);
                                float t = t1 + hit_distance / ray_length;
#endif
                                if (!(hit_distance > dist_inside)) {
                                    closest = t; any = true;
                                    best.t = t; best.pc = pc; best.face = 0; best.ctx = cur_ctx;
                                }
                            }
                        }
                    }
                }
                pc = end;
                break;
            }
            default:  // OP_END (or a stray record): stop
                pc = pc_end;
                break;
        }
    }
    return any;
}

// ConstantMedium boundary query (constant_medium.rs:37-38): closest hit of the boundary sub-stream [pc, end) in
// [tmin, +inf).  Out of line so that the hot loop carries ONE copy of the inner interpreter instead of two inlined ones.
__device__ __noinline__ float boundary_hit(const DeviceScene& S, int pc, int end, Ray world, Ray cur, int ctx, float tmin,
                                           bool reference_boxes) {
    Best dummy;
    MediumXi none;
    none.key.k0 = 0; none.key.k1 = 0; none.key.pixel = 0; none.key.sample = 0;
    none.bounce = 0; none.injected = 0.5f; none.inject = true;
    float t = CUDART_INF_F;
    const bool hit = traverse<true>(S, pc, end, world, cur, ctx, tmin, t, dummy, reference_boxes, none);
    return hit ? t : CUDART_NAN_F;  // NaN = miss (a genuine NaN t cannot be told apart and is treated as a miss)
}

// The tail of ConstantMedium::hit once both boundary hits are known (constant_medium.rs:40-75): true and the
// scattering distance `t_out` when the medium scatters inside [tmin, closest].  A = the medium record's first half.
__device__ __noinline__ bool medium_sample(const DeviceScene& S, float4 A, float dd, float t1, float t2, float tmin,
                                              float closest, const MediumXi& xi, float& t_out) {
    if (t1 < tmin) t1 = tmin;
    if (t2 > closest) t2 = closest;
    if (t1 >= t2) return false;
    if (t1 < 0.0f) t1 = 0.0f;
    const float ray_length = sqrtf(dd);
    const float dist_inside = (t2 - t1) * ray_length;
    const float u = xi.draw(__float_as_int(A.z));
#if HRT_EXACT
    const float hit_distance = A.x * (logf(u) / S.ln_e);
#else
    // accurate logf here too: __logf's absolute error near u = 1 is a relative error of up to ~1e-3 in the free-flight
    // distance
    const float hit_distance = A.x * logf(u);
#endif
    if (hit_distance > dist_inside) return false;
    t_out = t1 + hit_distance / ray_length;
    return true;
}

// The two boundary queries of a ConstantMedium whose boundary is one cuboid (constant_medium.rs:37-38): (t1, t2), t2 = NaN
// when either misses (or is NaN, which the walk of the sub-stream treats as a miss too).  Out of line: cold for every
// scene without such a medium.
__device__ __noinline__ float2 cuboid_boundary(float4 C, float4 D, Ray r, float4 kq) {
    RayK k;
    k.inv = v3(kq.x, kq.y, kq.z);
    k.dd = kq.w;
    float t1, t2;
    int face;
    if (cuboid_test(v3(C.x, C.y, C.z), v3(D.x, D.y, D.z), r, k, -CUDART_INF_F, CUDART_INF_F, t1, face) && t1 == t1 &&
        cuboid_test(v3(C.x, C.y, C.z), v3(D.x, D.y, D.z), r, k, t1 + 0.0001f, CUDART_INF_F, t2, face))
        return make_float2(t1, t2);
    return make_float2(0.0f, CUDART_NAN_F);
}

// ------------------------------------------------------------------------------------------------
// Warp-uniform traversal of the op stream.
//
// The stream is forward-only (a missed box jumps FORWARD to its skip link, everything else falls through), so the 32
// rays of a warp can walk it together: every step executes the record at the SMALLEST pc any lane still has to visit
// (one REDUX.MIN into a uniform register), for exactly the lanes that are at it; lanes that are further ahead wait.
// The record fetch is one broadcast load, and the opcode — hence every branch of the interpreter — is warp-uniform: no
// divergence between record kinds, no per-ray bookkeeping, no votes.  Per ray the visit order and the arithmetic are
// those of `traverse<>`, so results are identical.  The price is that a warp walks the UNION of its rays' records;
// that is what the short streams of the fast form want (Cornell: 34 records; `final`: a 21-node top level around two
// OP_BVH trees, which each ray walks on its own with a stack).
//
// ONE copy of the interpreter serves the world ray and the two boundary queries of a ConstantMedium
// (constant_medium.rs:37-38): a medium whose boundary is not a plain sphere switches the walk — warp-uniformly — into
// query mode over its sub-stream [pc+1, end): the medium's lanes restart there with the query's own range, (-inf, +inf)
// and then (t1 + 1e-4, +inf), and afterwards go on behind the sub-stream with the world ray's state restored.  Code
// size is what bounds this kernel (instruction fetch was the top stall of its first version, profiles/r02_u1_*), hence
// also the compact forms of the rect / cuboid tests.
// Every lane of the warp must call it (`active` = this lane carries a ray).
// ------------------------------------------------------------------------------------------------
// `pre` / `n_pre`: results of the first n_pre OP_BVH trees of the stream walked AHEAD for this ray over [tmin, +inf)
// (the wavefront render's tree stage): {t, code} with code = kPreNone: no hit, else leaf pc | side << 24.
constexpr int kPreNone = -1;
// kKeepSpace: keep the ray space entered last in registers (11 of them) — what the persistent kernel wants (a
// ConstantMedium enters its boundary's space twice per ray); the wavefront's trace kernel runs at 80 registers and is
// better off recomputing.
template <bool kKeepSpace = true>
__device__ __forceinline__ bool traverse_uniform(const DeviceScene& S, const int pc_begin, const int pc_end, const bool active,
                                                 const Ray& world, Ray cur, int cur_ctx, const float tmin_world, float& closest_io,
                                                 Best& best, const bool reference_boxes, const MediumXi& xi,
                                                 const float2* pre = nullptr, const int n_pre = 0) {
    const unsigned kAll = 0xffffffffu;
    int pc = active ? pc_begin : pc_end;
    RayK k = make_rayk(cur);
    const int entry_ctx = cur_ctx;  // the ray space the walk starts in (the world's, for a whole stream)
    const Ray entry_ray = cur;
    const RayK entry_k = k;
    int kept_ctx = -1;              // the other ray space entered last
    Ray kept_ray = cur;
    RayK kept_k = k;
    float closest = closest_io;
    float tmin = tmin_world;  // t_min of the query this lane is in
    bool hitf = false;        // the query this lane is in has hit something
    // query mode (warp-uniform): 0 = world ray; 1 / 2 = first / second boundary query of the medium record at m_pc
    int mode = 0, m_pc = 0, range_end = pc_end;
    bool in_q = false;        // this lane takes part in the medium's queries
    bool saved_hitf = false;
    float saved_closest = 0.0f, q_t1 = 0.0f;
    for (;;) {
        const int upc = __reduce_min_sync(kAll, pc);  // warp-uniform
        if (upc >= range_end) {
            if (mode == 0) break;
            // ---- a boundary query of the medium at m_pc is over for every lane that took part ----
            float4 A, B;
            load_op(S, m_pc, A, B);
            if (in_q) {
                const bool q_hit = hitf && closest == closest;  // a NaN boundary t counts as a miss
                if (mode == 1 && q_hit) {  // second query: (t1 + 1e-4, +inf)
                    q_t1 = closest;
                    pc = m_pc + 1; tmin = q_t1 + 0.0001f; closest = CUDART_INF_F; hitf = false;
                } else {  // the medium is decided: back to the world ray, behind the boundary sub-stream
                    const float t2 = closest;
                    closest = saved_closest; tmin = tmin_world; hitf = saved_hitf;
                    float t;
                    if (mode == 2 && q_hit && medium_sample(S, A, k.dd, q_t1, t2, tmin_world, closest, xi, t)) {
                        closest = t; hitf = true;
                        best.t = t; best.pc = m_pc; best.face = 0; best.ctx = cur_ctx;
                    }
                    pc = (int)(__float_as_uint(B.w) >> 8);
                    in_q = false;
                }
            }
            if (mode == 1 && __any_sync(kAll, in_q)) { mode = 2; continue; }
            mode = 0;
            range_end = pc_end;
            continue;
        }
        float4 A, B;
        load_op(S, upc, A, B);  // one address for the whole warp
        const uint32_t w7 = __float_as_uint(B.w);
        const uint32_t opc = w7 & 0xffu;
        const bool me = pc == upc;
        if (opc == OP_BOX || opc == OP_BOX_LOOSE) {
            if (me) {
                const bool hit = (opc == OP_BOX_LOOSE || reference_boxes) ? box_hit_reference(A, B, cur, k, tmin, closest)
                                                                          : box_hit_tight(A, B, cur, k, tmin, closest);
                pc = hit ? upc + 1 : (int)(w7 >> 8);
            }
            continue;
        }
        // primitive records share one acceptance tail
        bool prim = false, hit = false;
        float t = 0.0f;
        int face = 0, next = upc + 1;
        switch (opc) {
            case OP_SPHERE: case OP_MSPHERE: {
                prim = true;
                V3 ctr = v3(A.x, A.y, A.z);
                if (opc == OP_MSPHERE) {
                    float4 C, D;
                    load_op(S, upc + 1, C, D);
                    ctr = msphere_center(ctr, v3(C.x, C.y, C.z), C.w, D.x, cur.time);
                    next = upc + 2;
                }
                if (me) hit = sphere_test(ctr, A.w, cur, k, tmin, closest, t);
                break;
            }
            case OP_RECT_XY: case OP_RECT_YZ: case OP_RECT_ZX:
                prim = true;
                if (me) hit = rect_any(opc, A, B.x, cur, k, tmin, closest, t);
                break;
            case OP_CUBOID:
                prim = true;
                if (me) hit = cuboid_test(v3(A.x, A.y, A.z), v3(B.x, B.y, B.z), cur, k, tmin, closest, t, face);
                break;
            case OP_BVH: {  // a sound BvhNode as a two-child tree: every ray walks it on its own (bvh2_walk)
                const int tree = __float_as_int(B.z);
                if (mode == 0 && tree >= 0 && tree < n_pre) {
                    // walked ahead over [tmin, +inf): its closest hit t* is the tree's closest hit in [tmin, closest] iff
                    // t* <= closest (the tree is sound), and an exact tie with the hit held so far — an earlier record —
                    // is settled the reference's way
                    if (me) {
                        const float2 r = pre[tree];
                        const int code = __float_as_int(r.y);
                        if (code != kPreNone) {
                            const float tt = r.x;
                            const int leaf = code & 0xffffff;
                            bool take = tt < closest;
                            if (!take && !(tt > closest)) take = best.pc < 0 || tie_goes_to_later(S, leaf, cur, tmin, tt, B.x, B.y);
                            if (take) {
                                closest = tt; hitf = true;
                                best.t = tt; best.pc = leaf; best.face = code >> 24; best.ctx = cur_ctx;
                            }
                        }
                        pc = (int)(w7 >> 8);
                    }
                    break;
                }
                if (me) {
                    const int held = mode == 0 ? best.pc : -1;
                    const TreeHit th = bvh2_walk(S, __float_as_int(A.x), cur, tmin, closest, held, B.x, B.y);
                    if (th.pc != held) {
                        closest = th.t; hitf = true;
                        if (mode == 0) { best.t = th.t; best.pc = th.pc; best.face = th.face; best.ctx = cur_ctx; }
                    }
                    pc = (int)(w7 >> 8);
                }
                break;
            }
            case OP_BVH_PRE: {
                // (wave form of the stream) A tree walked ahead, met where the records that exist only for it begin: its
                // answer is merged as above and the walk goes on behind them — the sound boxes in between only prune what
                // `t* < closest` rejects anyway, and the tree's ray space is entered only to settle an exact tie.
                if (me) {
                    const float2 r = pre[__float_as_int(A.x)];
                    const int code = __float_as_int(r.y);
                    if (code != kPreNone) {
                        const float tt = r.x;
                        const int leaf = code & 0xffffff;
                        const int tctx = __float_as_int(A.y);
                        bool take = tt < closest;
                        if (!take && !(tt > closest))
                            take = best.pc < 0 ||
                                   tie_goes_to_later(S, leaf, tctx == cur_ctx ? cur : ray_in_ctx(S, world, tctx), tmin, tt, B.x, B.y);
                        if (take) {
                            closest = tt; hitf = true;
                            best.t = tt; best.pc = leaf; best.face = code >> 24; best.ctx = tctx;
                        }
                    }
                    pc = (int)(w7 >> 8);
                }
                break;
            }
            case OP_TRANSLATE: case OP_ROTATE: case OP_POP: {
                // A ray space is a pure function of (world ray, context): leaving to the world space restores the world
                // ray, and the space entered last is kept — a ConstantMedium enters its boundary's space twice per ray
                // (two queries), every Translation / Rotation is left again right after its child.
                if (me) {
                    cur_ctx = __float_as_int(A.w);
                    if (cur_ctx == entry_ctx) {
                        cur = entry_ray;
                        k = kKeepSpace ? entry_k : make_rayk(cur);
                    } else if (kKeepSpace && cur_ctx == kept_ctx) {
                        cur = kept_ray; k = kept_k;
                    } else {
                        cur = ray_in_ctx(S, world, cur_ctx);
                        k = make_rayk(cur);
                        if (kKeepSpace) { kept_ctx = cur_ctx; kept_ray = cur; kept_k = k; }
                    }
                    const int run = (int)(w7 >> 8);
                    pc = upc + (run > 0 ? run : 1);
                }
                break;
            }
            case OP_MEDIUM: case OP_MEDIUM_SPHERE: case OP_MEDIUM_CUBOID: {  // constant_medium.rs:34-76; media do not nest: mode == 0
                const int end = (int)(w7 >> 8);
                bool generic = me;
                if (opc == OP_MEDIUM_CUBOID) {
                    // Boundary = one cuboid in its own ray space: both boundary queries here, with the very calls the
                    // sub-stream walk would make (cuboid_test over (-inf, +inf), then over (t1 + 1e-4, +inf)) — one step
                    // instead of two walks over push / cuboid / pop records.
                    float4 P0, P1;
                    load_op(S, upc + 1, P0, P1);
                    const uint32_t first = __float_as_uint(P1.w);
                    const bool pushed = (first & 0xffu) != OP_CUBOID;
                    const int box_pc = upc + 1 + (pushed ? (int)(first >> 8) : 0);
                    float4 C, D;
                    load_op(S, box_pc, C, D);
                    generic = false;
                    if (me) {
                        Ray r = cur;
                        RayK kr = k;
                        if (pushed) {
                            const int target = __float_as_int(P0.w);
                            if (kKeepSpace && target == kept_ctx) {
                                r = kept_ray; kr = kept_k;
                            } else {
                                r = ray_in_ctx(S, world, target);
                                kr = make_rayk(r);
                                if (kKeepSpace) { kept_ctx = target; kept_ray = r; kept_k = kr; }
                            }
                        }
                        const float2 tt = cuboid_boundary(C, D, r, make_float4(kr.inv.x, kr.inv.y, kr.inv.z, kr.dd));
                        if (tt.y == tt.y) {
                            float tm;
                            if (medium_sample(S, A, k.dd, tt.x, tt.y, tmin, closest, xi, tm)) {
                                closest = tm; hitf = true;
                                best.t = tm; best.pc = upc; best.face = 0; best.ctx = cur_ctx;
                            }
                        }
                    }
                }
                if (opc == OP_MEDIUM_SPHERE) {
                    // Boundary = one plain sphere: both boundary queries in closed form with sphere_test's arithmetic
                    // (query 1 over (-inf, +inf) always takes the near root; query 2 over (t1 + 1e-4, +inf) takes the
                    // near root again when the f32 sum t1 + 1e-4 == t1, else the far root).
                    float4 C, D;
                    load_op(S, upc + 1, C, D);
                    generic = false;
                    if (me) {
                        const V3 oc = v3(__fsub_rn(cur.o.x, C.x), __fsub_rn(cur.o.y, C.y), __fsub_rn(cur.o.z, C.z));
                        const float a = k.dd;
                        const float half_b = dot_rn(oc, cur.d);
                        const float c = __fsub_rn(dot_rn(oc, oc), __fmul_rn(C.w, C.w));
                        const float disc = __fsub_rn(__fmul_rn(half_b, half_b), __fmul_rn(a, c));
                        if (!(disc < 0.0f)) {
                            const float sqrtd = sqrtf(disc);
                            const float t1 = __fdiv_rn(-half_b - sqrtd, a);
                            const float t2 = __fdiv_rn(-half_b + sqrtd, a);
                            const float lo = t1 + 0.0001f;
                            if (t1 == t1 && t2 == t2) {
                                float tm;
                                bool h = false;
                                if (!(t1 < lo)) h = medium_sample(S, A, k.dd, t1, t1, tmin, closest, xi, tm);
                                else if (!(t2 < lo)) h = medium_sample(S, A, k.dd, t1, t2, tmin, closest, xi, tm);
                                if (h) {
                                    closest = tm; hitf = true;
                                    best.t = tm; best.pc = upc; best.face = 0; best.ctx = cur_ctx;
                                }
                            } else {
                                generic = true;  // NaN roots take the generic path
                            }
                        }
                    }
                }
                if (me && !generic) pc = end;
                if (__any_sync(kAll, generic)) {  // query mode over [upc + 1, end): first query (-inf, +inf)
                    in_q = generic;
                    if (generic) {
                        saved_closest = closest; saved_hitf = hitf;
                        pc = upc + 1; tmin = -CUDART_INF_F; closest = CUDART_INF_F; hitf = false;
                    }
                    mode = 1; m_pc = upc; range_end = end;
                }
                break;
            }
            default:  // OP_END (or a stray record): stop
                if (me) pc = range_end;
                break;
        }
        if (prim && me) {
            if (hit) {
                closest = t; hitf = true;
                if (mode == 0) { best.t = t; best.pc = upc; best.face = face; best.ctx = cur_ctx; }
            }
            pc = next;
        }
    }
    if (hitf) closest_io = closest;
    return hitf;
}

// ------------------------------------------------------------------------------------------------
// Hit record
// ------------------------------------------------------------------------------------------------
struct HitRec {
    V3 p, n;
    float t, u, v;
    bool front_face;
    int mat, prim, face;
};

// hit_record.rs:22-29
__device__ __forceinline__ void set_face_normal(HitRec& h, V3 ray_d, V3 outward) {
    h.front_face = dot(ray_d, outward) < 0.0f;
    h.n = h.front_face ? outward : -outward;
}
// sphere.rs:31-36
__device__ __forceinline__ void sphere_uv(V3 p, float& u, float& v) {
    float theta = acosf(-p.y);
    float phi = atan2f(-p.z, p.x) + HRT_PI;
    u = phi / (2.0f * HRT_PI);
    v = theta / HRT_PI;
}

// Rebuild the full reference HitRecord for the winning record: evaluate the primitive in its own ray space
// at best.t, then undo the enclosing Rotation/Translation records innermost-first
// (rotation.rs:119-131, translation.rs:33-34).
__device__ __noinline__ void make_hit_record(const DeviceScene& S, const Ray& world, const Best& best, bool want_uv, HitRec& h) {
    // forward: world -> the record's ray space
    const int depth = best.ctx != 0 ? S.ctxs[best.ctx].depth : 0;
    const Ray r = depth > 0 ? ray_in_ctx(S, world, best.ctx) : world;
    float4 A, B;
    load_op(S, best.pc, A, B);
    const uint32_t w7 = __float_as_uint(B.w);
    const uint32_t opc = w7 & 0xffu;
    h.t = best.t;
    h.u = 0.0f; h.v = 0.0f;
    h.face = 0;
    h.p = ray_at(r, best.t);
    if (opc == OP_SPHERE || opc == OP_MSPHERE) {
        V3 ctr = v3(A.x, A.y, A.z);
        if (opc == OP_MSPHERE) {
            float4 C, D;
            load_op(S, best.pc + 1, C, D);
            ctr = msphere_center(ctr, v3(C.x, C.y, C.z), C.w, D.x, r.time);
        }
        const V3 outward = (h.p - ctr) / A.w;
        h.mat = __float_as_int(B.x);
        // sphere.rs:61 computes (u,v) on every hit; only image-textured materials ever read them
        if (want_uv || (S.mats[h.mat].flags & MATF_NEEDS_UV)) sphere_uv(outward, h.u, h.v);
        set_face_normal(h, r.d, outward);
        h.prim = __float_as_int(B.y);
    } else if (opc == OP_MEDIUM || opc == OP_MEDIUM_SPHERE || opc == OP_MEDIUM_CUBOID) {  // constant_medium.rs:67-75
        h.n = v3(0.0f, 0.0f, 0.0f);
        h.front_face = false;
        h.mat = __float_as_int(A.y);
        h.prim = __float_as_int(A.w);
    } else {
        // rect.rs:71-84 / the cuboid's side `face` (cuboid.rs:30-96): plane axis k, in-plane axes (a, b)
        int ik;
        float a0, a1, b0, b1;
        if (opc == OP_CUBOID) {
            const int f = best.face;
            ik = f < 2 ? 2 : (f < 4 ? 1 : 0);
            const int ia = ik == 2 ? 0 : ik + 1, ib = ik == 0 ? 2 : ik - 1;
            a0 = sel3(ia, A.x, A.y, A.z); a1 = sel3(ia, B.x, B.y, B.z);
            b0 = sel3(ib, A.x, A.y, A.z); b1 = sel3(ib, B.x, B.y, B.z);
            h.mat = __float_as_int(A.w);
            h.prim = (int)(w7 >> 8);
            h.face = f;
        } else {
            ik = opc == OP_RECT_XY ? 2 : (opc == OP_RECT_YZ ? 0 : 1);
            a0 = A.x; a1 = A.y; b0 = A.z; b1 = A.w;
            h.mat = __float_as_int(B.y);
            h.prim = __float_as_int(B.z);
        }
        const int ia = ik == 2 ? 0 : ik + 1, ib = ik == 0 ? 2 : ik - 1;
        h.u = (sel3(ia, h.p.x, h.p.y, h.p.z) - a0) / (a1 - a0);  // rect.rs:75-76
        h.v = (sel3(ib, h.p.x, h.p.y, h.p.z) - b0) / (b1 - b0);
        set_face_normal(h, r.d, v3(ik == 0 ? 1.0f : 0.0f, ik == 1 ? 1.0f : 0.0f, ik == 2 ? 1.0f : 0.0f));
    }
    // backward: the record's space -> world, innermost Rotation / Translation first (rotation.rs:119-131,
    // translation.rs:33-34).  Translation re-faces the normal against ITS moved ray (Q4), whose direction is the world
    // direction mapped through the records outside it.
#pragma unroll 1
    for (int i = depth - 1; i >= 0; --i) {
        const int op_pc = S.ctxs[best.ctx].op_pc[i];
        float4 PA, PB;
        load_op(S, op_pc, PA, PB);
        if ((__float_as_uint(PB.w) & 0xffu) == OP_TRANSLATE) {
            Ray rr = world;
#pragma unroll 1
            for (int j = 0; j < i; ++j) {
                float4 QA, QB;
                load_op(S, S.ctxs[best.ctx].op_pc[j], QA, QB);
                if ((__float_as_uint(QB.w) & 0xffu) == OP_TRANSLATE) apply_translate(rr, QA);
                else apply_rotate(rr, QA);
            }
            h.p = h.p + v3(PA.x, PA.y, PA.z);
            set_face_normal(h, rr.d, h.n);
        } else {
            h.p = unrotate(h.p, PA);
            h.n = unrotate(h.n, PA);
        }
    }
}

// ------------------------------------------------------------------------------------------------
// Textures (src/textures/*.rs, src/perlin_noise.rs)
// ------------------------------------------------------------------------------------------------
struct NoiseView {  // where the tables of noise texture `i` live (global or shared memory)
    const float4* ranvec;
    const uint8_t* perm;  // 3 x 256
};

// perlin_noise.rs:80-123 (Q6: smoothed u,v,w inside the weight vector too)
__device__ __forceinline__ float perlin_noise(const NoiseView& N, V3 p) {
    float fx = floorf(p.x), fy = floorf(p.y), fz = floorf(p.z);
    int i = (int)fx, j = (int)fy, k = (int)fz;
    float u = p.x - fx, v = p.y - fy, w = p.z - fz;
    u = u * u * (3.0f - 2.0f * u);
    v = v * v * (3.0f - 2.0f * v);
    w = w * w * (3.0f - 2.0f * w);
    uint32_t px0 = N.perm[(i)&255], px1 = N.perm[(i + 1) & 255];
    uint32_t py0 = N.perm[256 + ((j)&255)], py1 = N.perm[256 + ((j + 1) & 255)];
    uint32_t pz0 = N.perm[512 + ((k)&255)], pz1 = N.perm[512 + ((k + 1) & 255)];
    float acc = 0.0f;
#pragma unroll
    for (int idx = 0; idx < 8; ++idx) {
        const int x = idx >> 2, y = (idx >> 1) & 1, z = idx & 1;
        float4 g = N.ranvec[(x ? px1 : px0) ^ (y ? py1 : py0) ^ (z ? pz1 : pz0)];
        float wx = u - (float)x, wy = v - (float)y, wz = w - (float)z;
        float bx = x ? u : (1.0f - u), by = y ? v : (1.0f - v), bz = z ? w : (1.0f - w);
        // (x*u + (1-x)*(1-u)) reduces exactly to u or (1-u): the other term is an exact 0
        acc += ((bx * by) * bz) * ((g.x * wx + g.y * wy) + g.z * wz);
    }
    return acc;
}
// perlin_noise.rs:66-78
__device__ __forceinline__ float perlin_turbulence(const NoiseView& N, V3 p, int depth) {
    float acc = 0.0f, weight = 1.0f;
#pragma unroll 1
    for (int d = 0; d < depth; ++d) {
        acc += weight * perlin_noise(N, p);
        weight *= 0.5f;
        p = p * 2.0f;
    }
    return fabsf(acc);
}

// Per-kernel texture environment.  The first kMaxNoiseTablesShared perlin tables are staged in shared memory by the
// kernels that shade (stage_noise in hrt_kernels.cu); `n_shared_noise` says how many.
struct TexEnv {
    const NoiseTable* sh_noise;  // shared-memory copies of the first n_shared_noise tables
    int n_shared_noise;
};
__device__ __forceinline__ NoiseView noise_view(const DeviceScene& S, const TexEnv E, int table) {
    NoiseView nv;
    if (table < E.n_shared_noise) {
        nv.ranvec = reinterpret_cast<const float4*>(E.sh_noise[table].ranvec);
        nv.perm = &E.sh_noise[table].perm[0][0];
    } else {
        nv.ranvec = reinterpret_cast<const float4*>(S.noise[table].ranvec);
        nv.perm = &S.noise[table].perm[0][0];
    }
    return nv;
}

// Out of line (one copy per kernel): perlin turbulence + image fetch + libm sinf slow paths are cold, bulky code.
__device__ __noinline__ V3 texture_value(const DeviceScene& S, const TexEnv E, int tex, float u, float v, V3 p) {
    Texture T = S.texs[tex];
    // checker_texture.rs:22-30 — select and descend (checkers may nest)
    while (T.kind == TEX_CHECKER) {
        float sines = 1.0f;  // (sin(10x) * sin(10y)) * sin(10z); 1.0f * s is exact
#pragma unroll 1
        for (int a = 0; a < 3; ++a) sines *= sinf(10.0f * comp(p, a));
        T = S.texs[sines < 0.0f ? T.i0 : T.i1];
    }
    if (T.kind == TEX_SOLID) return v3(T.v[0], T.v[1], T.v[2]);  // solid_color.rs:21-23
    if (T.kind == TEX_NOISE) {                                   // noise_texture.rs:25-31 (Q7)
        const float scale = T.v[0];
        NoiseView nv = noise_view(S, E, T.i0);
        float arg = (scale * p.z) + (10.0f * perlin_turbulence(nv, scale * p, 7));
        float s = sinf(arg);
        float g = 0.5f * (1.0f + s);
        return v3(g, g, g);
    }
    // image_texture.rs:36-63
    if (T.i0 < 0) return v3(1.0f, 0.0f, 1.0f);
    float uc = (u < 0.0f) ? 0.0f : ((u > 1.0f) ? 1.0f : u);  // f32::clamp keeps NaN
    float vc = (v < 0.0f) ? 0.0f : ((v > 1.0f) ? 1.0f : v);
    vc = 1.0f - vc;
    const uint32_t W = (uint32_t)T.i1, H = (uint32_t)T.i2;
    uint32_t i = __float2uint_rz(uc * (float)W);  // saturating, NaN -> 0: same as Rust `as u32`
    uint32_t j = __float2uint_rz(vc * (float)H);
    if (i >= W) i = W - 1;
    if (j >= H) j = H - 1;
    uchar4 c = tex2D<uchar4>(S.images[T.i0], (float)i + 0.5f, (float)j + 0.5f);
    const float color_scale = 1.0f / 255.0f;
    return v3(color_scale * (float)c.x, color_scale * (float)c.y, color_scale * (float)c.z);
}

// ------------------------------------------------------------------------------------------------
// Materials (src/materials/*.rs, src/math.rs)
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ bool near_zero(V3 v) {  // math.rs:42-45
    const float S = 1e-8f;
    return (fabsf(v.x) < S) && (fabsf(v.y) < S) && (fabsf(v.z) < S);
}
__device__ __forceinline__ V3 reflect(V3 v, V3 n) { return v - (2.0f * dot(v, n)) * n; }  // math.rs:47-49
__device__ __forceinline__ V3 refract(V3 uv, V3 n, float eta) {                            // math.rs:51-56
    float cos_theta = fminf(dot(-uv, n), 1.0f);
    V3 perp = eta * (uv + cos_theta * n);
    V3 par = (-sqrtf(fabsf(1.0f - dot(perp, perp)))) * n;
    return perp + par;
}
__device__ __forceinline__ float reflectance(float cosine, float ref_idx) {  // math.rs:58-62
    float r0 = (1.0f - ref_idx) / (1.0f + ref_idx);
    r0 = r0 * r0;
#if HRT_EXACT
    return r0 + (1.0f - r0) * powf(1.0f - cosine, 5.0f);
#else
    float m = 1.0f - cosine;
    float m2 = m * m;
    return r0 + (1.0f - r0) * (m2 * m2 * m);
#endif
}

// Material::emitted (materials/mod.rs:18; only DiffuseLight is non-zero, diffuse_light.rs:25-27)
__device__ __forceinline__ V3 material_emitted(const DeviceScene& S, const TexEnv E, const Material& m, const HitRec& h) {
    if (m.kind == MAT_DIFFUSE_LIGHT) return texture_value(S, E, m.tex, h.u, h.v, h.p);
    return v3(0.0f, 0.0f, 0.0f);
}
// Material::scatter with uniforms u[0..3].  Returns false for "None".
// `defer_noise` (wavefront render): when the albedo is a NoiseTexture, do not evaluate it — 7 octaves of perlin noise
// for the few lanes of a warp that hit such a surface — but report its id; the caller queues (slot, texture, point) and
// a compacted kernel multiplies the path's throughput by the value afterwards.
__device__ __forceinline__ bool material_scatter(const DeviceScene& S, const TexEnv E, const Material& m, const Ray& ray,
                                                 const HitRec& h, const float u[4], V3& attenuation, Ray& scattered,
                                                 int* defer_noise = nullptr) {
    switch (m.kind) {
        case MAT_LAMBERTIAN: {  // lambertian.rs:27-38
            V3 dir = h.n + sample_unit_vector(u[0], u[1]);
            if (near_zero(dir)) dir = h.n;
            if (defer_noise && m.tex < 1024 && S.texs[m.tex].kind == TEX_NOISE) { *defer_noise = m.tex; attenuation = v3(1.0f, 1.0f, 1.0f); }
            else attenuation = texture_value(S, E, m.tex, h.u, h.v, h.p);
            scattered = Ray{h.p, dir, ray.time};
            return true;
        }
        case MAT_METAL: {  // metal.rs:29-42
            V3 reflected = reflect(normalize(ray.d), h.n);
            V3 dir = reflected + m.param * sample_in_unit_sphere(u[0], u[1], u[2]);
            scattered = Ray{h.p, dir, ray.time};
            if (dot(dir, h.n) > 0.0f) {
                attenuation = v3(m.albedo[0], m.albedo[1], m.albedo[2]);
                return true;
            }
            return false;
        }
        case MAT_DIELECTRIC: {  // dielectric.rs:31-55
            float ratio = h.front_face ? (1.0f / m.param) : m.param;
            V3 unit = normalize(ray.d);
            float cos_theta = fminf(dot(-unit, h.n), 1.0f);
            float sin_theta = sqrtf(1.0f - cos_theta * cos_theta);
            bool cannot_refract = (ratio * sin_theta) > 1.0f;
            V3 dir;
            if (cannot_refract || reflectance(cos_theta, ratio) > u[0]) dir = reflect(unit, h.n);
            else dir = refract(unit, h.n, ratio);
            attenuation = v3(1.0f, 1.0f, 1.0f);
            scattered = Ray{h.p, dir, ray.time};
            return true;
        }
        case MAT_ISOTROPIC: {  // isotropic.rs:27-33
            if (defer_noise && m.tex < 1024 && S.texs[m.tex].kind == TEX_NOISE) { *defer_noise = m.tex; attenuation = v3(1.0f, 1.0f, 1.0f); }
            else attenuation = texture_value(S, E, m.tex, h.u, h.v, h.p);
            scattered = Ray{h.p, sample_in_unit_sphere(u[0], u[1], u[2]), ray.time};
            return true;
        }
        default:  // MAT_DIFFUSE_LIGHT: diffuse_light.rs:21-23
            return false;
    }
}

// ------------------------------------------------------------------------------------------------
// Camera (src/camera.rs:85-95)
// ------------------------------------------------------------------------------------------------
struct CameraK {
    V3 origin, lower_left_corner, horizontal, vertical, u, v;
    float lens_radius, time0, time1;
};
__device__ __forceinline__ Ray camera_get_ray(const CameraK& c, float s, float t, float u_lens1, float u_lens2, float u_time) {
    V3 rd = c.lens_radius * sample_in_unit_disk(u_lens1, u_lens2);
    V3 offset = c.u * rd.x + c.v * rd.y;
    Ray r;
    r.o = c.origin + offset;
    r.d = c.lower_left_corner + s * c.horizontal + t * c.vertical - c.origin - offset;
    r.time = c.time0 + (c.time1 - c.time0) * u_time;
    return r;
}

}  // namespace HRT_NS
