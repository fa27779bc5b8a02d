// oracle.cpp — CPU ORACLE for the hyper-ray-tracer path-tracing hot path.
//
// THIS IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only tests/, __graft_entry__.smoke()
// and bench.py's cpu_baseline / --impl reference legs may load it.  The product path
// (libhrt.so, hyper-ray-tracer_b200/) never links, imports or calls anything in here.
//
// What it is: a plain C++17 restatement of the reference's (Rust, CPU) algorithm for the
// per-pixel sample loop and everything beneath it, *including the reference's behavioural
// quirks* (SURVEY.md §8a Q1..Q15).  Every function cites the /root/reference file:line it
// follows.  Arithmetic is IEEE f32 without FMA contraction (build with -ffp-contract=off),
// libm calls are glibc sinf/cosf/tanf/acosf/atan2f/sqrtf/powf/logf/floorf, which is what
// rustc lowers f32 methods to on linux-gnu.
//
// Parity status: the reference has no tests, golden vectors or fixtures and
// cannot be built here (no Rust toolchain), so this oracle is pinned against
// known-answer vectors authored from the reference source (tests/test_oracle_kat.py):
// BVH topology fixtures, analytic hit records, the Q1/Q2 light-clipping vectors, perlin and
// image-texel KATs.  Third-party arithmetic (rand 0.8.5 / rand_chacha 0.3.1 streams, cgmath
// 0.18.0, image 0.24.5 + jpeg-decoder 0.3.0) is "parity unpinned": rand streams are
// OS-seeded and unreproducible by design — only the distributions are restated.
//
// Build: see oracle/Makefile (g++ -O3 -ffp-contract=off -pthread -shared -fPIC).

#include <atomic>
#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <limits>
#include <memory>
#include <thread>
#include <vector>

namespace orc {

// ---------------------------------------------------------------------------------------------
// Vec3 = cgmath::Vector3<f32>  (src/math.rs:10).  cgmath 0.18.0 semantics: dot = (x*x + y*y) + z*z,
// magnitude = sqrt(dot), normalize = v * (1 / magnitude), v / s = per-component divide.
// ---------------------------------------------------------------------------------------------
struct Vec3 {
    float x, y, z;
    Vec3() : x(0), y(0), z(0) {}
    Vec3(float a, float b, float c) : x(a), y(b), z(c) {}
    float operator[](int i) const { return i == 0 ? x : (i == 1 ? y : z); }
    float& operator[](int i) { return i == 0 ? x : (i == 1 ? y : z); }
};
static inline Vec3 operator+(Vec3 a, Vec3 b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
static inline Vec3 operator-(Vec3 a, Vec3 b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
static inline Vec3 operator-(Vec3 a) { return {-a.x, -a.y, -a.z}; }
static inline Vec3 operator*(float s, Vec3 a) { return {s * a.x, s * a.y, s * a.z}; }
static inline Vec3 operator*(Vec3 a, float s) { return {a.x * s, a.y * s, a.z * s}; }
static inline Vec3 operator/(Vec3 a, float s) { return {a.x / s, a.y / s, a.z / s}; }
static inline Vec3 mul_element_wise(Vec3 a, Vec3 b) { return {a.x * b.x, a.y * b.y, a.z * b.z}; }
static inline float dot(Vec3 a, Vec3 b) { return (a.x * b.x + a.y * b.y) + a.z * b.z; }
static inline Vec3 cross(Vec3 a, Vec3 b) {
    return {a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x};
}
static inline float magnitude(Vec3 a) { return std::sqrt(dot(a, a)); }
static inline Vec3 normalize(Vec3 a) { return a * (1.0f / magnitude(a)); }

static const float PI_F = 3.14159265358979323846f;
static const float E_F = 2.71828182845904523536f;
static const float INF_F = std::numeric_limits<float>::infinity();

// ---------------------------------------------------------------------------------------------
// RNG.  The reference draws from rand::thread_rng() (ChaCha12, OS-seeded) — unreproducible.
// The oracle keeps the *distributions* (rand 0.8.5: gen::<f32>() = 24-bit [0,1);
// gen_range(a..b) = 23-bit mantissa * (b-a) + a, half-open) on a seeded xoshiro128++.
// In trace_hits mode the medium draw is injected so that the CUDA path can be compared.
// ---------------------------------------------------------------------------------------------
struct Rng {
    uint32_t s[4];
    bool injected = false;
    float injected_value = 0.5f;
    static inline uint32_t rotl(uint32_t x, int k) { return (x << k) | (x >> (32 - k)); }
    void seed(uint64_t seed) {
        // splitmix64 expansion
        uint64_t z = seed;
        for (int i = 0; i < 4; i += 2) {
            z += 0x9E3779B97F4A7C15ull;
            uint64_t r = z;
            r = (r ^ (r >> 30)) * 0xBF58476D1CE4E5B9ull;
            r = (r ^ (r >> 27)) * 0x94D049BB133111EBull;
            r = r ^ (r >> 31);
            s[i] = (uint32_t)r;
            s[i + 1] = (uint32_t)(r >> 32);
        }
        if ((s[0] | s[1] | s[2] | s[3]) == 0) s[0] = 1;
    }
    uint32_t next_u32() {
        uint32_t result = rotl(s[0] + s[3], 7) + s[0];
        uint32_t t = s[1] << 9;
        s[2] ^= s[0];
        s[3] ^= s[1];
        s[1] ^= s[2];
        s[0] ^= s[3];
        s[2] ^= t;
        s[3] = rotl(s[3], 11);
        return result;
    }
    // rand::Rng::gen::<f32>()
    float gen_f32() {
        if (injected) return injected_value;
        return (float)(next_u32() >> 8) * (1.0f / 16777216.0f);
    }
    // rand::Rng::gen_range(lo..hi) for f32 (half-open)
    float gen_range(float lo, float hi) {
        float scale = hi - lo;
        for (;;) {
            float v01 = (float)(next_u32() >> 9) * (1.0f / 8388608.0f);
            float res = v01 * scale + lo;
            if (res < hi) return res;
        }
    }
};

struct Counters {
    uint64_t paths = 0, rays = 0, aabb_tests = 0, sphere_tests = 0, rect_tests = 0, medium_queries = 0,
             noise_evals = 0, scatters = 0;
    void add(const Counters& o) {
        paths += o.paths; rays += o.rays; aabb_tests += o.aabb_tests; sphere_tests += o.sphere_tests;
        rect_tests += o.rect_tests; medium_queries += o.medium_queries; noise_evals += o.noise_evals;
        scatters += o.scatters;
    }
};

static thread_local Rng* tls_rng = nullptr;
static thread_local Counters tls_cnt;
// 0 = reference per-axis-independent slab test (src/aabb.rs:20-47); 1 = intersected ("tight") slab
// test, used ONLY to count the traversal work a sound-box traversal needs.
static int g_aabb_mode = 0;
// WORK-MODEL STUDIES ONLY (tools/work_model.py --bvh-study): ORC_BVH_STUDY=sah builds every BvhNode with a full-sweep
// surface-area heuristic instead of the reference's longest-axis median split, ORC_BVH_STUDY=sah+near additionally
// visits the child whose box the ray enters first.  Neither is the reference's algorithm; they only answer "how many
// box tests would a better tree / order save" (SURVEY.md §8f N4).  Unset (the default) = reference behaviour.
static int bvh_study_mode() {
    static const int mode = [] {
        const char* e = std::getenv("ORC_BVH_STUDY");
        if (!e) return 0;
        return std::strcmp(e, "sah+near") == 0 ? 2 : (std::strcmp(e, "sah") == 0 ? 1 : (std::strcmp(e, "near") == 0 ? 3 : 0));
    }();
    return mode;
}

// ---------------------------------------------------------------------------------------------
// src/math.rs
// ---------------------------------------------------------------------------------------------
// src/math.rs:16-30
static Vec3 random_in_unit_sphere() {
    Rng& r = *tls_rng;
    for (;;) {
        float a = r.gen_range(-1.0f, 1.0f);
        float b = r.gen_range(-1.0f, 1.0f);
        float c = r.gen_range(-1.0f, 1.0f);
        Vec3 p(a, b, c);
        if (dot(p, p) < 1.0f) return p;
    }
}
// src/math.rs:12-14
static Vec3 random_unit_vector() { return normalize(random_in_unit_sphere()); }
// src/math.rs:32-40
static Vec3 random_in_unit_disk() {
    Rng& r = *tls_rng;
    for (;;) {
        float a = r.gen_range(-1.0f, 1.0f);
        float b = r.gen_range(-1.0f, 1.0f);
        Vec3 p(a, b, 0.0f);
        if (dot(p, p) < 1.0f) return p;
    }
}
// src/math.rs:42-45
static bool near_zero(Vec3 v) {
    const float S = 1e-8f;
    return (std::fabs(v.x) < S) && (std::fabs(v.y) < S) && (std::fabs(v.z) < S);
}
// src/math.rs:47-49
static Vec3 reflect(Vec3 v, Vec3 n) { return v - (2.0f * dot(v, n)) * n; }
// src/math.rs:51-56
static Vec3 refract(Vec3 uv, Vec3 n, float etai_over_etat) {
    float cos_theta = std::fmin(dot(-uv, n), 1.0f);
    Vec3 r_out_perp = etai_over_etat * (uv + cos_theta * n);
    Vec3 r_out_parallel = (-std::sqrt(std::fabs(1.0f - dot(r_out_perp, r_out_perp)))) * n;
    return r_out_perp + r_out_parallel;
}
// src/math.rs:58-62
static float reflectance(float cosine, float refraction_index) {
    float r0 = (1.0f - refraction_index) / (1.0f + refraction_index);
    r0 = r0 * r0;
    return r0 + (1.0f - r0) * powf(1.0f - cosine, 5.0f);
}

// Direct (fixed-draw-count) samplers of the SAME distributions as the rejection loops above.
// These are NOT in the reference; they restate the mapping the CUDA path uses so that scatter
// arithmetic can be compared 1:1 under injected uniforms (orc_scatter_direct), and so that a test
// can check rejection ≡ direct in distribution.
static Vec3 direct_unit_vector(float u1, float u2) {
    float z = 1.0f - 2.0f * u1;
    float r = std::sqrt(std::fmax(0.0f, 1.0f - z * z));
    float phi = (2.0f * PI_F) * u2;
    return Vec3(r * cosf(phi), r * sinf(phi), z);
}
static Vec3 direct_in_unit_sphere(float u1, float u2, float u3) { return cbrtf(u3) * direct_unit_vector(u1, u2); }
static Vec3 direct_in_unit_disk(float u1, float u2) {
    float r = std::sqrt(u1);
    float phi = (2.0f * PI_F) * u2;
    return Vec3(r * cosf(phi), r * sinf(phi), 0.0f);
}

// ---------------------------------------------------------------------------------------------
// src/ray.rs, src/hit_record.rs, src/aabb.rs
// ---------------------------------------------------------------------------------------------
struct Ray {
    Vec3 origin, direction;
    float time;
    Vec3 at(float t) const { return origin + t * direction; }  // src/ray.rs:25-27
};

struct Material;
struct HitRecord {
    Vec3 point, normal;
    float t = 0, u = 0, v = 0;
    bool front_face = false;
    const Material* material = nullptr;
    int prim_id = -1;  // oracle bookkeeping (object id of the primitive that produced the record)
    int face = 0;      // cuboid side index 0..5
    // src/hit_record.rs:22-29
    void set_face_normal(const Ray& ray, Vec3 outward_normal) {
        front_face = dot(ray.direction, outward_normal) < 0.0f;
        normal = front_face ? outward_normal : -outward_normal;
    }
};

struct Aabb {
    Vec3 minimum, maximum;
    // src/aabb.rs:20-47 — NOTE (Q1): t_min/t_max are fresh locals per axis; the three slab intervals
    // are never intersected with each other.
    bool hit(const Ray& ray, float time_min, float time_max) const {
        tls_cnt.aabb_tests++;
        if (g_aabb_mode == 1) return hit_tight(ray, time_min, time_max);
        for (int a = 0; a < 3; ++a) {
            float inverse_direction = 1.0f / ray.direction[a];
            float time_start = (minimum[a] - ray.origin[a]) * inverse_direction;
            float time_end = (maximum[a] - ray.origin[a]) * inverse_direction;
            if (inverse_direction < 0.0f) std::swap(time_start, time_end);
            float t_min = time_start > time_min ? time_start : time_min;
            float t_max = time_end < time_max ? time_end : time_max;
            if (t_max <= t_min) return false;
        }
        return true;
    }
    // entry distance of the ray into the box (study modes only; +inf when missed)
    float entry_tight(const Ray& ray, float time_min, float time_max) const {
        for (int a = 0; a < 3; ++a) {
            float inverse_direction = 1.0f / ray.direction[a];
            float time_start = (minimum[a] - ray.origin[a]) * inverse_direction;
            float time_end = (maximum[a] - ray.origin[a]) * inverse_direction;
            if (inverse_direction < 0.0f) std::swap(time_start, time_end);
            time_min = time_start > time_min ? time_start : time_min;
            time_max = time_end < time_max ? time_end : time_max;
            if (time_max <= time_min) return std::numeric_limits<float>::infinity();
        }
        return time_min;
    }
    float half_area() const {
        Vec3 d = maximum - minimum;
        return d.x * d.y + d.y * d.z + d.z * d.x;
    }
    bool hit_tight(const Ray& ray, float time_min, float time_max) const {
        for (int a = 0; a < 3; ++a) {
            float inverse_direction = 1.0f / ray.direction[a];
            float time_start = (minimum[a] - ray.origin[a]) * inverse_direction;
            float time_end = (maximum[a] - ray.origin[a]) * inverse_direction;
            if (inverse_direction < 0.0f) std::swap(time_start, time_end);
            time_min = time_start > time_min ? time_start : time_min;
            time_max = time_end < time_max ? time_end : time_max;
            if (time_max <= time_min) return false;
        }
        return true;
    }
    // src/aabb.rs:49-63
    static Aabb surrounding_box(const Aabb& b0, const Aabb& b1) {
        Vec3 small(std::fmin(b0.minimum.x, b1.minimum.x), std::fmin(b0.minimum.y, b1.minimum.y),
                   std::fmin(b0.minimum.z, b1.minimum.z));
        Vec3 big(std::fmax(b0.maximum.x, b1.maximum.x), std::fmax(b0.maximum.y, b1.maximum.y),
                 std::fmax(b0.maximum.z, b1.maximum.z));
        return Aabb{small, big};
    }
};

// ---------------------------------------------------------------------------------------------
// src/perlin_noise.rs — tables are INPUT (the reference draws them from thread_rng, :23-64)
// ---------------------------------------------------------------------------------------------
struct PerlinNoise {
    Vec3 random_vectors[256];
    uint32_t permutation_x[256], permutation_y[256], permutation_z[256];

    // src/perlin_noise.rs:104-123 — NOTE (Q6): the Hermite-smoothed u,v,w are used both for the
    // blend weights AND inside the weight vector.
    static float trilinear_interpolation(const Vec3 c[2][2][2], float u, float v, float w) {
        u = u * u * (3.0f - 2.0f * u);
        v = v * v * (3.0f - 2.0f * v);
        w = w * w * (3.0f - 2.0f * w);
        float accumulator = 0.0f;
        for (int i = 0; i < 8; ++i) {
            int x = i / 4, y = (i / 2) % 2, z = i % 2;
            Vec3 weight(u - (float)x, v - (float)y, w - (float)z);
            accumulator += ((float)x * u + (float)(1 - x) * (1.0f - u)) * ((float)y * v + (float)(1 - y) * (1.0f - v)) *
                           ((float)z * w + (float)(1 - z) * (1.0f - w)) * dot(c[x][y][z], weight);
        }
        return accumulator;
    }
    // src/perlin_noise.rs:80-102
    float noise(Vec3 point) const {
        tls_cnt.noise_evals++;
        int i = (int)std::floor(point.x);
        int j = (int)std::floor(point.y);
        int k = (int)std::floor(point.z);
        Vec3 c[2][2][2];
        for (int index = 0; index < 8; ++index) {
            int i_x = index / 4, i_y = (index / 2) % 2, i_z = index % 2;
            uint32_t x = permutation_x[(i + i_x) & 255];
            uint32_t y = permutation_y[(j + i_y) & 255];
            uint32_t z = permutation_z[(k + i_z) & 255];
            c[i_x][i_y][i_z] = random_vectors[x ^ y ^ z];
        }
        float u = point.x - std::floor(point.x);
        float v = point.y - std::floor(point.y);
        float w = point.z - std::floor(point.z);
        return trilinear_interpolation(c, u, v, w);
    }
    // src/perlin_noise.rs:66-78
    float turbulence(Vec3 point, uint32_t depth) const {
        float accumulator = 0.0f;
        float weight = 1.0f;
        for (uint32_t d = 0; d < depth; ++d) {
            accumulator += weight * noise(point);
            weight *= 0.5f;
            point = point * 2.0f;
        }
        return std::fabs(accumulator);
    }
};

// ---------------------------------------------------------------------------------------------
// src/textures/*
// ---------------------------------------------------------------------------------------------
struct Texture {
    virtual ~Texture() {}
    virtual Vec3 value(float u, float v, Vec3 point) const = 0;  // src/textures/mod.rs:14-16
};
struct SolidColor : Texture {
    Vec3 color;
    Vec3 value(float, float, Vec3) const override { return color; }  // src/textures/solid_color.rs:21-23
};
struct CheckerTexture : Texture {
    const Texture *odd, *even;
    // src/textures/checker_texture.rs:22-30
    Vec3 value(float u, float v, Vec3 point) const override {
        float sines = sinf(10.0f * point.x) * sinf(10.0f * point.y) * sinf(10.0f * point.z);
        return sines < 0.0f ? odd->value(u, v, point) : even->value(u, v, point);
    }
};
struct NoiseTexture : Texture {
    PerlinNoise noise;
    float scale;
    // src/textures/noise_texture.rs:25-31 — NOTE (Q7): turbulence is fed scale*point.
    Vec3 value(float, float, Vec3 point) const override {
        float s = sinf((scale * point.z) + (10.0f * noise.turbulence(scale * point, 7)));
        return (Vec3(1.0f, 1.0f, 1.0f) * 0.5f) * (1.0f + s);
    }
};
struct ImageTexture : Texture {
    std::vector<uint8_t> data;
    uint32_t components = 0, width = 0, height = 0, bytes_per_scanline = 0;
    // src/textures/image_texture.rs:36-63 — nearest texel, no filtering, no sRGB decode.
    Vec3 value(float u, float v, Vec3) const override {
        if (data.empty()) return Vec3(1.0f, 0.0f, 1.0f);
        // f32::clamp: NaN stays NaN
        u = (u < 0.0f) ? 0.0f : ((u > 1.0f) ? 1.0f : u);
        float vc = (v < 0.0f) ? 0.0f : ((v > 1.0f) ? 1.0f : v);
        v = 1.0f - vc;
        // Rust `as u32` saturates; NaN -> 0
        auto as_u32 = [](float f) -> uint32_t {
            if (!(f == f)) return 0u;
            if (f <= 0.0f) return 0u;
            if (f >= 4294967296.0f) return 0xFFFFFFFFu;
            return (uint32_t)f;
        };
        uint32_t i = as_u32(u * (float)width);
        uint32_t j = as_u32(v * (float)height);
        if (i >= width) i = width - 1;
        if (j >= height) j = height - 1;
        const float color_scale = 1.0f / 255.0f;
        size_t offset = (size_t)j * bytes_per_scanline + (size_t)i * components;
        return Vec3(color_scale * (float)data[offset], color_scale * (float)data[offset + 1],
                    color_scale * (float)data[offset + 2]);
    }
};

// ---------------------------------------------------------------------------------------------
// src/materials/*
// ---------------------------------------------------------------------------------------------
struct Material {
    int id = -1;
    virtual ~Material() {}
    // src/materials/mod.rs:15-19.  `u4` != nullptr selects the direct samplers with injected uniforms.
    virtual bool scatter(const Ray& ray, const HitRecord& rec, Vec3& attenuation, Ray& scattered,
                         const float* u4) const = 0;
    virtual Vec3 emitted(float, float, Vec3) const { return Vec3(0, 0, 0); }
};
struct Lambertian : Material {
    const Texture* albedo;
    // src/materials/lambertian.rs:27-38
    bool scatter(const Ray& ray, const HitRecord& rec, Vec3& attenuation, Ray& scattered,
                 const float* u4) const override {
        Vec3 scatter_direction = rec.normal + (u4 ? direct_unit_vector(u4[0], u4[1]) : random_unit_vector());
        if (near_zero(scatter_direction)) scatter_direction = rec.normal;
        attenuation = albedo->value(rec.u, rec.v, rec.point);
        scattered = Ray{rec.point, scatter_direction, ray.time};
        return true;
    }
};
struct Metal : Material {
    Vec3 albedo;
    float fuzz;
    // src/materials/metal.rs:29-42 — NOTE (Q9): fuzz is not clamped.
    bool scatter(const Ray& ray, const HitRecord& rec, Vec3& attenuation, Ray& scattered,
                 const float* u4) const override {
        Vec3 reflected = reflect(normalize(ray.direction), rec.normal);
        Vec3 fz = u4 ? direct_in_unit_sphere(u4[0], u4[1], u4[2]) : random_in_unit_sphere();
        scattered = Ray{rec.point, reflected + fuzz * fz, ray.time};
        if (dot(scattered.direction, rec.normal) > 0.0f) {
            attenuation = albedo;
            return true;
        }
        return false;
    }
};
struct Dielectric : Material {
    float index_of_refraction;
    // src/materials/dielectric.rs:31-55
    bool scatter(const Ray& ray, const HitRecord& rec, Vec3& attenuation, Ray& scattered,
                 const float* u4) const override {
        float refraction_ratio = rec.front_face ? (1.0f / index_of_refraction) : index_of_refraction;
        Vec3 unit_direction = normalize(ray.direction);
        float cos_theta = std::fmin(dot(-unit_direction, rec.normal), 1.0f);
        float sin_theta = std::sqrt(1.0f - cos_theta * cos_theta);
        bool cannot_refract = (refraction_ratio * sin_theta) > 1.0f;
        Vec3 direction;
        // short-circuit OR: the draw happens only when refraction is possible (:45)
        if (cannot_refract || reflectance(cos_theta, refraction_ratio) > (u4 ? u4[0] : tls_rng->gen_f32()))
            direction = reflect(unit_direction, rec.normal);
        else
            direction = refract(unit_direction, rec.normal, refraction_ratio);
        attenuation = Vec3(1.0f, 1.0f, 1.0f);
        scattered = Ray{rec.point, direction, ray.time};
        return true;
    }
};
struct DiffuseLight : Material {
    const Texture* emit;
    // src/materials/diffuse_light.rs:21-27
    bool scatter(const Ray&, const HitRecord&, Vec3&, Ray&, const float*) const override { return false; }
    Vec3 emitted(float u, float v, Vec3 p) const override { return emit->value(u, v, p); }
};
struct Isotropic : Material {
    const Texture* albedo;
    // src/materials/isotropic.rs:27-33 — direction is uniform IN the unit ball, not normalised.
    bool scatter(const Ray& ray, const HitRecord& rec, Vec3& attenuation, Ray& scattered,
                 const float* u4) const override {
        attenuation = albedo->value(rec.u, rec.v, rec.point);
        scattered = Ray{rec.point, u4 ? direct_in_unit_sphere(u4[0], u4[1], u4[2]) : random_in_unit_sphere(),
                        ray.time};
        return true;
    }
};

// ---------------------------------------------------------------------------------------------
// src/hittable/*
// ---------------------------------------------------------------------------------------------
struct Hittable {
    int id = -1;
    virtual ~Hittable() {}
    // src/hittable/mod.rs:19-25
    virtual bool hit(const Ray& ray, float time_min, float time_max, HitRecord& out) const = 0;
    virtual bool bounding_box(float time_start, float time_end, Aabb& out) const = 0;
    virtual uint32_t count() const = 0;
};

// src/hittable/sphere.rs:31-36 (and moving_sphere.rs:44-49)
static inline void sphere_uv(Vec3 p, float& u, float& v) {
    float theta = acosf(-p.y);
    float phi = atan2f(-p.z, p.x) + PI_F;
    u = phi / (2.0f * PI_F);
    v = theta / PI_F;
}

struct Sphere : Hittable {
    Vec3 center;
    float radius;
    const Material* material;
    // src/hittable/sphere.rs:40-75
    bool hit(const Ray& ray, float time_min, float time_max, HitRecord& rec) const override {
        tls_cnt.sphere_tests++;
        Vec3 oc = ray.origin - center;
        float a = dot(ray.direction, ray.direction);
        float half_b = dot(oc, ray.direction);
        float c = dot(oc, oc) - radius * radius;
        float discriminant = half_b * half_b - a * c;
        if (discriminant < 0.0f) return false;
        float sqrtd = std::sqrt(discriminant);
        float root = (-half_b - sqrtd) / a;
        if (root < time_min || time_max < root) {
            root = (-half_b + sqrtd) / a;
            if (root < time_min || time_max < root) return false;
        }
        Vec3 outward_normal = (ray.at(root) - center) / radius;
        rec = HitRecord();
        sphere_uv(outward_normal, rec.u, rec.v);
        rec.point = ray.at(root);
        rec.t = root;
        rec.material = material;
        rec.prim_id = id;
        rec.set_face_normal(ray, outward_normal);
        return true;
    }
    // src/hittable/sphere.rs:77-83
    bool bounding_box(float, float, Aabb& out) const override {
        Vec3 rv(radius, radius, radius);
        out = Aabb{center - rv, center + rv};
        return true;
    }
    uint32_t count() const override { return 1; }
};

struct MovingSphere : Hittable {
    Vec3 center_start, center_end;
    float time_start, time_end, radius;
    const Material* material;
    // src/hittable/moving_sphere.rs:53-57
    Vec3 center(float time) const {
        return center_start + ((time - time_start) / (time_end - time_start)) * (center_end - center_start);
    }
    // src/hittable/moving_sphere.rs:61-96
    bool hit(const Ray& ray, float time_min, float time_max, HitRecord& rec) const override {
        tls_cnt.sphere_tests++;
        Vec3 oc = ray.origin - center(ray.time);
        float a = dot(ray.direction, ray.direction);
        float half_b = dot(oc, ray.direction);
        float c = dot(oc, oc) - radius * radius;
        float discriminant = half_b * half_b - a * c;
        if (discriminant < 0.0f) return false;
        float sqrtd = std::sqrt(discriminant);
        float root = (-half_b - sqrtd) / a;
        if (root < time_min || time_max < root) {
            root = (-half_b + sqrtd) / a;
            if (root < time_min || time_max < root) return false;
        }
        Vec3 outward_normal = (ray.at(root) - center(ray.time)) / radius;
        rec = HitRecord();
        sphere_uv(outward_normal, rec.u, rec.v);
        rec.point = ray.at(root);
        rec.t = root;
        rec.material = material;
        rec.prim_id = id;
        rec.set_face_normal(ray, outward_normal);
        return true;
    }
    // src/hittable/moving_sphere.rs:98-110
    bool bounding_box(float t0, float t1, Aabb& out) const override {
        Vec3 rv(radius, radius, radius);
        Aabb b0{center(t0) - rv, center(t0) + rv};
        Aabb b1{center(t1) - rv, center(t1) + rv};
        out = Aabb::surrounding_box(b0, b1);
        return true;
    }
    uint32_t count() const override { return 1; }
};

enum Plane { PLANE_XY = 0, PLANE_YZ = 1, PLANE_ZX = 2 };  // src/hittable/rect.rs:13-17

struct Rect : Hittable {
    int plane;
    float a0, a1, b0, b1, k;
    const Material* material;
    // src/hittable/rect.rs:53-86 — NOTE (Q12): ZX maps (k=y, a=z, b=x).
    bool hit(const Ray& ray, float time_min, float time_max, HitRecord& rec) const override {
        tls_cnt.rect_tests++;
        int k_axis, a_axis, b_axis;
        switch (plane) {
            case PLANE_XY: k_axis = 2; a_axis = 0; b_axis = 1; break;
            case PLANE_YZ: k_axis = 0; a_axis = 1; b_axis = 2; break;
            default:       k_axis = 1; a_axis = 2; b_axis = 0; break;
        }
        float t = (k - ray.origin[k_axis]) / ray.direction[k_axis];
        if (t < time_min || t > time_max) return false;  // NaN passes (Q15)
        float a = ray.origin[a_axis] + t * ray.direction[a_axis];
        float b = ray.origin[b_axis] + t * ray.direction[b_axis];
        if (a < a0 || a > a1 || b < b0 || b > b1) return false;
        rec = HitRecord();
        rec.point = ray.at(t);
        rec.t = t;
        rec.u = (a - a0) / (a1 - a0);
        rec.v = (b - b0) / (b1 - b0);
        rec.material = material;
        rec.prim_id = id;
        Vec3 outward_normal(0, 0, 0);
        outward_normal[k_axis] = 1.0f;
        rec.set_face_normal(ray, outward_normal);
        return true;
    }
    // src/hittable/rect.rs:88-103 — NOTE (Q2): the ZX box is built as x∈[a0,a1], z∈[b0,b1] although
    // hit() treats a as z and b as x.
    bool bounding_box(float, float, Aabb& out) const override {
        switch (plane) {
            case PLANE_XY: out = Aabb{Vec3(a0, b0, k - 0.0001f), Vec3(a1, b1, k + 0.0001f)}; break;
            case PLANE_YZ: out = Aabb{Vec3(k - 0.0001f, a0, b0), Vec3(k + 0.0001f, a1, b1)}; break;
            default:       out = Aabb{Vec3(a0, k - 0.0001f, b0), Vec3(a1, k + 0.0001f, b1)}; break;
        }
        return true;
    }
    uint32_t count() const override { return 1; }
};

struct List : Hittable {
    std::vector<const Hittable*> objects;
    // src/hittable/list.rs:20-31
    bool hit(const Ray& ray, float time_min, float time_max, HitRecord& rec) const override {
        float closest = time_max;
        bool hit_anything = false;
        HitRecord tmp;
        for (size_t i = 0; i < objects.size(); ++i) {
            if (objects[i]->hit(ray, time_min, closest, tmp)) {
                closest = tmp.t;
                rec = tmp;
                rec.face = (int)i;
                hit_anything = true;
            }
        }
        return hit_anything;
    }
    // src/hittable/list.rs:33-44
    bool bounding_box(float t0, float t1, Aabb& out) const override {
        if (objects.empty()) return false;
        Aabb acc;
        if (!objects[0]->bounding_box(t0, t1, acc)) return false;
        for (size_t i = 1; i < objects.size(); ++i) {
            Aabb b;
            if (!objects[i]->bounding_box(t0, t1, b)) return false;
            acc = Aabb::surrounding_box(acc, b);
        }
        out = acc;
        return true;
    }
    uint32_t count() const override {
        uint32_t n = 0;
        for (auto* o : objects) n += o->count();
        return n;
    }
};

struct Cuboid : Hittable {
    Vec3 box_min, box_max;
    List sides;
    std::vector<std::unique_ptr<Rect>> owned;
    // src/hittable/cuboid.rs:30-96 — side order: XY@max.z, XY@min.z, ZX@max.y, ZX@min.y, YZ@max.x, YZ@min.x
    Cuboid(Vec3 mn, Vec3 mx, const Material* m) : box_min(mn), box_max(mx) {
        auto add = [&](int plane, float a0, float a1, float b0, float b1, float k) {
            auto r = std::make_unique<Rect>();
            r->plane = plane; r->a0 = a0; r->a1 = a1; r->b0 = b0; r->b1 = b1; r->k = k; r->material = m;
            sides.objects.push_back(r.get());
            owned.push_back(std::move(r));
        };
        add(PLANE_XY, mn.x, mx.x, mn.y, mx.y, mx.z);
        add(PLANE_XY, mn.x, mx.x, mn.y, mx.y, mn.z);
        add(PLANE_ZX, mn.z, mx.z, mn.x, mx.x, mx.y);
        add(PLANE_ZX, mn.z, mx.z, mn.x, mx.x, mn.y);
        add(PLANE_YZ, mn.y, mx.y, mn.z, mx.z, mx.x);
        add(PLANE_YZ, mn.y, mx.y, mn.z, mx.z, mn.x);
    }
    // src/hittable/cuboid.rs:100-102
    bool hit(const Ray& ray, float time_min, float time_max, HitRecord& rec) const override {
        if (!sides.hit(ray, time_min, time_max, rec)) return false;
        rec.prim_id = id;
        return true;
    }
    // src/hittable/cuboid.rs:104-106
    bool bounding_box(float, float, Aabb& out) const override {
        out = Aabb{box_min, box_max};
        return true;
    }
    uint32_t count() const override { return sides.count(); }
};

struct Translation : Hittable {
    const Hittable* hittable;
    Vec3 displacement;
    // src/hittable/translation.rs:24-37 — NOTE (Q4): set_face_normal is re-applied to the already
    // face-forwarded normal, so front_face comes out true (false only when d·n >= 0, e.g. medium n=0).
    bool hit(const Ray& ray, float time_min, float time_max, HitRecord& rec) const override {
        Ray moved{ray.origin - displacement, ray.direction, ray.time};
        if (!hittable->hit(moved, time_min, time_max, rec)) return false;
        rec.point = rec.point + displacement;
        rec.set_face_normal(moved, rec.normal);
        return true;
    }
    // src/hittable/translation.rs:39-48
    bool bounding_box(float t0, float t1, Aabb& out) const override {
        Aabb b;
        if (!hittable->bounding_box(t0, t1, b)) return false;
        out = Aabb{b.minimum + displacement, b.maximum + displacement};
        return true;
    }
    uint32_t count() const override { return hittable->count(); }
};

struct Rotation : Hittable {
    const Hittable* hittable;
    float sin_theta, cos_theta;
    bool has_box;
    Aabb bbox;
    int axis;  // 0 X, 1 Y, 2 Z
    // src/hittable/rotation.rs:20-26
    static void get_axes(int axis, int& r, int& a, int& b) {
        switch (axis) {
            case 0: r = 0; a = 1; b = 2; break;
            case 1: r = 1; a = 2; b = 0; break;
            default: r = 2; a = 0; b = 1; break;
        }
    }
    // src/hittable/rotation.rs:38-98
    Rotation(int axis_, const Hittable* h, float angle) : hittable(h), axis(axis_) {
        int r_axis, a_axis, b_axis;
        get_axes(axis, r_axis, a_axis, b_axis);
        float radians = (PI_F / 180.0f) * angle;
        sin_theta = sinf(radians);
        cos_theta = cosf(radians);
        Aabb bb;
        has_box = h->bounding_box(0.0f, 1.0f, bb);
        if (has_box) {
            const float FMAX = std::numeric_limits<float>::max();
            Vec3 mn(FMAX, FMAX, FMAX), mx(-FMAX, -FMAX, -FMAX);
            for (int i = 0; i < 2; ++i)
                for (int j = 0; j < 2; ++j)
                    for (int k = 0; k < 2; ++k) {
                        float r = (float)k * bb.maximum[r_axis] + (float)(1 - k) * bb.minimum[r_axis];
                        float a = (float)i * bb.maximum[a_axis] + (float)(1 - i) * bb.minimum[a_axis];
                        float b = (float)j * bb.maximum[b_axis] + (float)(1 - j) * bb.minimum[b_axis];
                        float new_a = cos_theta * a - sin_theta * b;
                        float new_b = sin_theta * a + cos_theta * b;
                        if (new_a < mn[a_axis]) mn[a_axis] = new_a;
                        if (new_b < mn[b_axis]) mn[b_axis] = new_b;
                        if (r < mn[r_axis]) mn[r_axis] = r;
                        if (new_a > mx[a_axis]) mx[a_axis] = new_a;
                        if (new_b > mx[b_axis]) mx[b_axis] = new_b;
                        if (r > mx[r_axis]) mx[r_axis] = r;
                    }
            bbox = Aabb{mn, mx};
        }
    }
    // src/hittable/rotation.rs:102-134 — front_face is left as computed in object space (Q5).
    bool hit(const Ray& ray, float time_min, float time_max, HitRecord& rec) const override {
        int r_axis, a_axis, b_axis;
        get_axes(axis, r_axis, a_axis, b_axis);
        Vec3 origin = ray.origin, direction = ray.direction;
        origin[a_axis] = cos_theta * ray.origin[a_axis] + sin_theta * ray.origin[b_axis];
        origin[b_axis] = -sin_theta * ray.origin[a_axis] + cos_theta * ray.origin[b_axis];
        direction[a_axis] = cos_theta * ray.direction[a_axis] + sin_theta * ray.direction[b_axis];
        direction[b_axis] = -sin_theta * ray.direction[a_axis] + cos_theta * ray.direction[b_axis];
        Ray rotated{origin, direction, ray.time};
        if (!hittable->hit(rotated, time_min, time_max, rec)) return false;
        Vec3 point = rec.point, normal = rec.normal;
        point[a_axis] = cos_theta * rec.point[a_axis] - sin_theta * rec.point[b_axis];
        point[b_axis] = sin_theta * rec.point[a_axis] + cos_theta * rec.point[b_axis];
        normal[a_axis] = cos_theta * rec.normal[a_axis] - sin_theta * rec.normal[b_axis];
        normal[b_axis] = sin_theta * rec.normal[a_axis] + cos_theta * rec.normal[b_axis];
        rec.point = point;
        rec.normal = normal;
        return true;
    }
    // src/hittable/rotation.rs:136-138
    bool bounding_box(float, float, Aabb& out) const override {
        if (!has_box) return false;
        out = bbox;
        return true;
    }
    uint32_t count() const override { return 1; }  // src/hittable/rotation.rs:140-142
};

struct ConstantMedium : Hittable {
    const Hittable* boundary;
    float negative_inverse_density;
    const Isotropic* phase_function;
    // src/hittable/constant_medium.rs:34-76 — NOTE (Q8): normal (0,0,0), u=v=0, front_face=false,
    // second boundary query starts at t1 + 0.0001.
    bool hit(const Ray& ray, float time_min, float time_max, HitRecord& rec) const override {
        tls_cnt.medium_queries++;
        HitRecord r1, r2;
        if (!boundary->hit(ray, -INF_F, INF_F, r1)) return false;
        if (!boundary->hit(ray, r1.t + 0.0001f, INF_F, r2)) return false;
        if (r1.t < time_min) r1.t = time_min;
        if (r2.t > time_max) r2.t = time_max;
        if (r1.t >= r2.t) return false;
        if (r1.t < 0.0f) r1.t = 0.0f;
        float ray_length = magnitude(ray.direction);
        float distance_inside_boundary = (r2.t - r1.t) * ray_length;
        float xi = tls_rng->gen_f32();
        float hit_distance = negative_inverse_density * (logf(xi) / logf(E_F));  // f32::log(self, E)
        if (hit_distance > distance_inside_boundary) return false;
        float t = r1.t + hit_distance / ray_length;
        rec = HitRecord();
        rec.point = ray.at(t);
        rec.normal = Vec3(0, 0, 0);
        rec.t = t;
        rec.u = 0.0f;
        rec.v = 0.0f;
        rec.front_face = false;
        rec.material = phase_function;
        rec.prim_id = id;
        return true;
    }
    // src/hittable/constant_medium.rs:78-80
    bool bounding_box(float t0, float t1, Aabb& out) const override { return boundary->bounding_box(t0, t1, out); }
    uint32_t count() const override { return boundary->count(); }
};

struct BvhNode : Hittable {
    // src/hittable/bvh_node.rs:11-24
    std::unique_ptr<BvhNode> left, right;  // Branch
    const Hittable* leaf = nullptr;        // Leaf
    Aabb bbox;

    // src/hittable/bvh_node.rs:83-100
    static float axis_range(const std::vector<const Hittable*>& objects, float t0, float t1, int axis) {
        float mn = std::numeric_limits<float>::max(), mx = -std::numeric_limits<float>::max();
        for (auto* o : objects) {
            Aabb b;
            if (!o->bounding_box(t0, t1, b)) continue;
            mn = std::fmin(mn, b.minimum[axis]);
            mx = std::fmax(mx, b.maximum[axis]);
        }
        return mx - mn;
    }
    // src/hittable/bvh_node.rs:27-63.  sort_unstable_by: for n <= 20 rustc's implementation is an
    // insertion sort (= stable order); for larger n the tie order is toolchain-defined.  The oracle uses
    // a stable sort throughout ("parity unpinned" for tie order at n > 20; ties cannot change hit
    // results on sound boxes — SURVEY.md §8a a25).
    BvhNode(std::vector<const Hittable*> objects, float t0, float t1) {
        const int study = bvh_study_mode();
        if ((study == 1 || study == 2) && objects.size() > 2) {
            build_sah(std::move(objects), t0, t1);
            return;
        }
        std::pair<int, float> ranges[3];
        for (int a = 0; a < 3; ++a) ranges[a] = {a, axis_range(objects, t0, t1, a)};
        std::stable_sort(ranges, ranges + 3, [](const std::pair<int, float>& a, const std::pair<int, float>& b) {
            return a.second > b.second;
        });
        int axis = ranges[0].first;
        // src/hittable/bvh_node.rs:65-81
        std::stable_sort(objects.begin(), objects.end(), [&](const Hittable* a, const Hittable* b) {
            Aabb ba, bb;
            a->bounding_box(t0, t1, ba);
            b->bounding_box(t0, t1, bb);
            float ac = ba.minimum[axis] + ba.maximum[axis];
            float bc = bb.minimum[axis] + bb.maximum[axis];
            return ac < bc;
        });
        size_t len = objects.size();
        if (len == 1) {
            leaf = objects[0];
            leaf->bounding_box(t0, t1, bbox);
        } else {
            std::vector<const Hittable*> r(objects.begin() + len / 2, objects.end());
            std::vector<const Hittable*> l(objects.begin(), objects.begin() + len / 2);
            right = std::make_unique<BvhNode>(std::move(r), t0, t1);
            left = std::make_unique<BvhNode>(std::move(l), t0, t1);
            bbox = Aabb::surrounding_box(left->bbox, right->bbox);
        }
    }
    // study only (see bvh_study_mode): full-sweep SAH over the three centroid orders
    void build_sah(std::vector<const Hittable*> objects, float t0, float t1) {
        const size_t n = objects.size();
        float best_cost = std::numeric_limits<float>::infinity();
        int best_axis = 0;
        size_t best_split = n / 2;
        std::vector<Aabb> boxes(n);
        std::vector<float> right_area(n + 1);
        for (int axis = 0; axis < 3; ++axis) {
            std::stable_sort(objects.begin(), objects.end(), [&](const Hittable* a, const Hittable* b) {
                Aabb ba, bb;
                a->bounding_box(t0, t1, ba);
                b->bounding_box(t0, t1, bb);
                return ba.minimum[axis] + ba.maximum[axis] < bb.minimum[axis] + bb.maximum[axis];
            });
            for (size_t i = 0; i < n; ++i) objects[i]->bounding_box(t0, t1, boxes[i]);
            Aabb acc = boxes[n - 1];
            right_area[n - 1] = acc.half_area();
            for (size_t i = n - 1; i-- > 0;) { acc = Aabb::surrounding_box(acc, boxes[i]); right_area[i] = acc.half_area(); }
            acc = boxes[0];
            for (size_t i = 1; i < n; ++i) {  // left = [0, i), right = [i, n)
                const float cost = acc.half_area() * (float)i + right_area[i] * (float)(n - i);
                if (cost < best_cost) { best_cost = cost; best_axis = axis; best_split = i; }
                acc = Aabb::surrounding_box(acc, boxes[i]);
            }
        }
        std::stable_sort(objects.begin(), objects.end(), [&](const Hittable* a, const Hittable* b) {
            Aabb ba, bb;
            a->bounding_box(t0, t1, ba);
            b->bounding_box(t0, t1, bb);
            return ba.minimum[best_axis] + ba.maximum[best_axis] < bb.minimum[best_axis] + bb.maximum[best_axis];
        });
        std::vector<const Hittable*> r(objects.begin() + (long)best_split, objects.end());
        std::vector<const Hittable*> l(objects.begin(), objects.begin() + (long)best_split);
        right = std::make_unique<BvhNode>(std::move(r), t0, t1);
        left = std::make_unique<BvhNode>(std::move(l), t0, t1);
        bbox = Aabb::surrounding_box(left->bbox, right->bbox);
    }
    // src/hittable/bvh_node.rs:104-127 — left first, narrow t_max, right wins if it hits (Q3).
    bool hit(const Ray& ray, float time_min, float time_max, HitRecord& rec) const override {
        if (!bbox.hit(ray, time_min, time_max)) return false;
        if (leaf) return leaf->hit(ray, time_min, time_max, rec);
        if (bvh_study_mode() >= 2) {  // study only: nearer child first (not the reference's order)
            const BvhNode* a = left.get();
            const BvhNode* b = right.get();
            if (b->bbox.entry_tight(ray, time_min, time_max) < a->bbox.entry_tight(ray, time_min, time_max)) std::swap(a, b);
            HitRecord arec, brec;
            const bool ahit = a->hit(ray, time_min, time_max, arec);
            if (ahit) time_max = arec.t;
            const bool bhit = b->hit(ray, time_min, time_max, brec);
            if (bhit) { rec = brec; return true; }
            if (ahit) { rec = arec; return true; }
            return false;
        }
        HitRecord lrec;
        bool lhit = left->hit(ray, time_min, time_max, lrec);
        if (lhit) time_max = lrec.t;
        HitRecord rrec;
        bool rhit = right->hit(ray, time_min, time_max, rrec);
        if (rhit) { rec = rrec; return true; }
        if (lhit) { rec = lrec; return true; }
        return false;
    }
    bool bounding_box(float, float, Aabb& out) const override { out = bbox; return true; }
    uint32_t count() const override { return leaf ? leaf->count() : left->count() + right->count(); }
    void leaf_order(std::vector<int>& out) const {
        if (leaf) out.push_back(leaf->id);
        else { left->leaf_order(out); right->leaf_order(out); }
    }
    uint32_t node_count() const { return leaf ? 1 : 1 + left->node_count() + right->node_count(); }
};

// ---------------------------------------------------------------------------------------------
// src/camera.rs
// ---------------------------------------------------------------------------------------------
struct Camera {
    Vec3 origin, lower_left_corner, horizontal, vertical, look_from, look_at, w, u, v;
    float fov, focus_dist, lens_radius, time_0, time_1;
    // src/camera.rs:34-83
    Camera(Vec3 from, Vec3 at, float fov_, float aperture, float focus, float t0, float t1, int width, int height)
        : look_from(from), look_at(at), fov(fov_), focus_dist(focus), lens_radius(aperture / 2.0f), time_0(t0),
          time_1(t1) {
        resize(width, height);
    }
    void resize(int width, int height) {
        float aspect_ratio = (float)width / (float)height;
        float theta = fov * (PI_F / 180.0f);  // f32::to_radians
        float h = tanf(theta / 2.0f);
        float viewport_height = 2.0f * h;
        float viewport_width = aspect_ratio * viewport_height;
        w = normalize(look_from - look_at);
        u = normalize(cross(Vec3(0.0f, 1.0f, 0.0f), w));
        v = cross(w, u);
        origin = look_from;
        horizontal = (focus_dist * viewport_width) * u;
        vertical = (focus_dist * viewport_height) * v;
        lower_left_corner = origin - horizontal / 2.0f - vertical / 2.0f - focus_dist * w;
    }
    // src/camera.rs:85-95.  `u4` != nullptr: direct disk sampler + injected time uniform.
    Ray get_ray(float s, float t, const float* u3 = nullptr) const {
        Vec3 rd = lens_radius * (u3 ? direct_in_unit_disk(u3[0], u3[1]) : random_in_unit_disk());
        Vec3 offset = u * rd.x + v * rd.y;
        float time = u3 ? (time_0 + (time_1 - time_0) * u3[2]) : tls_rng->gen_range(time_0, time_1);
        return Ray{origin + offset, lower_left_corner + s * horizontal + t * vertical - origin - offset, time};
    }
};

// ---------------------------------------------------------------------------------------------
// src/application.rs:477-495  ray_color (recursive, as in the reference)
// ---------------------------------------------------------------------------------------------
static Vec3 ray_color(const Ray& ray, Vec3 background, const Hittable* world, uint32_t depth) {
    if (depth == 0) return Vec3(0, 0, 0);
    tls_cnt.rays++;
    HitRecord rec;
    if (!world->hit(ray, 0.001f, INF_F, rec)) return background;
    Vec3 emitted = rec.material->emitted(rec.u, rec.v, rec.point);
    Vec3 attenuation;
    Ray scattered;
    if (!rec.material->scatter(ray, rec, attenuation, scattered, nullptr)) return emitted;
    tls_cnt.scatters++;
    Vec3 c = ray_color(scattered, background, world, depth - 1);
    return mul_element_wise(attenuation, c) + emitted;
}

// ---------------------------------------------------------------------------------------------
// Scene container + C API
// ---------------------------------------------------------------------------------------------
struct Scene {
    std::vector<std::unique_ptr<Texture>> textures;
    std::vector<std::unique_ptr<Material>> materials;
    std::vector<std::unique_ptr<Hittable>> objects;
    const Hittable* world = nullptr;
    bool tex_ok(int i) const { return i >= 0 && (size_t)i < textures.size(); }
    bool mat_ok(int i) const { return i >= 0 && (size_t)i < materials.size(); }
    bool obj_ok(int i) const { return i >= 0 && (size_t)i < objects.size(); }
    int add_obj(Hittable* h) {
        h->id = (int)objects.size();
        objects.emplace_back(h);
        return h->id;
    }
    int add_mat(Material* m) {
        m->id = (int)materials.size();
        materials.emplace_back(m);
        return m->id;
    }
};

}  // namespace orc

using namespace orc;

extern "C" {

struct orc_camera_desc {
    float look_from[3], look_at[3];
    float vfov, aperture, focus_dist, time0, time1;
    int32_t width, height;
};
struct orc_camera_state {
    float origin[3], lower_left_corner[3], horizontal[3], vertical[3], u[3], v[3], w[3];
    float lens_radius, time0, time1;
};
struct orc_render_desc {
    int32_t width, height, samples, depth;
    float background[3];
    int32_t tile_size;
    uint64_t seed;
    int32_t threads;
    int32_t aabb_mode;  // 0 reference-loose, 1 tight (counter studies only)
    int32_t verbose;
};
struct orc_counters {
    uint64_t paths, rays, aabb_tests, sphere_tests, rect_tests, medium_queries, noise_evals, scatters;
    double seconds;
};
struct orc_ray {
    float o[3], d[3], time, tmin, tmax;
};
struct orc_hit {
    int32_t hit;
    float t, p[3], n[3], u, v;
    int32_t front_face, material_id, prim_id, face;
};

void* orc_scene_create() { return new Scene(); }
void orc_scene_destroy(void* s) { delete (Scene*)s; }

int orc_tex_solid(void* sp, const float* rgb) {
    Scene* s = (Scene*)sp;
    auto* t = new SolidColor();
    t->color = Vec3(rgb[0], rgb[1], rgb[2]);
    s->textures.emplace_back(t);
    return (int)s->textures.size() - 1;
}
int orc_tex_checker(void* sp, int odd, int even) {
    Scene* s = (Scene*)sp;
    if (!s->tex_ok(odd) || !s->tex_ok(even)) return -1;
    auto* t = new CheckerTexture();
    t->odd = s->textures[odd].get();
    t->even = s->textures[even].get();
    s->textures.emplace_back(t);
    return (int)s->textures.size() - 1;
}
int orc_tex_noise(void* sp, float scale, const float* ranvec, const uint32_t* px, const uint32_t* py,
                  const uint32_t* pz) {
    Scene* s = (Scene*)sp;
    auto* t = new NoiseTexture();
    t->scale = scale;
    for (int i = 0; i < 256; ++i) {
        t->noise.random_vectors[i] = Vec3(ranvec[3 * i], ranvec[3 * i + 1], ranvec[3 * i + 2]);
        t->noise.permutation_x[i] = px[i];
        t->noise.permutation_y[i] = py[i];
        t->noise.permutation_z[i] = pz[i];
    }
    s->textures.emplace_back(t);
    return (int)s->textures.size() - 1;
}
int orc_tex_image(void* sp, const uint8_t* data, uint32_t w, uint32_t h, uint32_t comps) {
    Scene* s = (Scene*)sp;
    auto* t = new ImageTexture();
    if (data && w && h && comps) t->data.assign(data, data + (size_t)w * h * comps);
    t->components = comps;
    t->width = w;
    t->height = h;
    t->bytes_per_scanline = comps * w;
    s->textures.emplace_back(t);
    return (int)s->textures.size() - 1;
}
int orc_mat_lambertian(void* sp, int tex) {
    Scene* s = (Scene*)sp;
    if (!s->tex_ok(tex)) return -1;
    auto* m = new Lambertian();
    m->albedo = s->textures[tex].get();
    return s->add_mat(m);
}
int orc_mat_metal(void* sp, const float* rgb, float fuzz) {
    Scene* s = (Scene*)sp;
    auto* m = new Metal();
    m->albedo = Vec3(rgb[0], rgb[1], rgb[2]);
    m->fuzz = fuzz;
    return s->add_mat(m);
}
int orc_mat_dielectric(void* sp, float ior) {
    Scene* s = (Scene*)sp;
    auto* m = new Dielectric();
    m->index_of_refraction = ior;
    return s->add_mat(m);
}
int orc_mat_diffuse_light(void* sp, int tex) {
    Scene* s = (Scene*)sp;
    if (!s->tex_ok(tex)) return -1;
    auto* m = new DiffuseLight();
    m->emit = s->textures[tex].get();
    return s->add_mat(m);
}
int orc_sphere(void* sp, const float* c, float r, int mat) {
    Scene* s = (Scene*)sp;
    if (!s->mat_ok(mat)) return -1;
    auto* o = new Sphere();
    o->center = Vec3(c[0], c[1], c[2]);
    o->radius = r;
    o->material = s->materials[mat].get();
    return s->add_obj(o);
}
int orc_moving_sphere(void* sp, const float* c0, const float* c1, float t0, float t1, float r, int mat) {
    Scene* s = (Scene*)sp;
    if (!s->mat_ok(mat)) return -1;
    auto* o = new MovingSphere();
    o->center_start = Vec3(c0[0], c0[1], c0[2]);
    o->center_end = Vec3(c1[0], c1[1], c1[2]);
    o->time_start = t0;
    o->time_end = t1;
    o->radius = r;
    o->material = s->materials[mat].get();
    return s->add_obj(o);
}
int orc_rect(void* sp, int plane, float a0, float a1, float b0, float b1, float k, int mat) {
    Scene* s = (Scene*)sp;
    if (!s->mat_ok(mat) || plane < 0 || plane > 2) return -1;
    auto* o = new Rect();
    o->plane = plane; o->a0 = a0; o->a1 = a1; o->b0 = b0; o->b1 = b1; o->k = k;
    o->material = s->materials[mat].get();
    return s->add_obj(o);
}
int orc_cuboid(void* sp, const float* mn, const float* mx, int mat) {
    Scene* s = (Scene*)sp;
    if (!s->mat_ok(mat)) return -1;
    return s->add_obj(new Cuboid(Vec3(mn[0], mn[1], mn[2]), Vec3(mx[0], mx[1], mx[2]), s->materials[mat].get()));
}
int orc_translate(void* sp, int child, const float* d) {
    Scene* s = (Scene*)sp;
    if (!s->obj_ok(child)) return -1;
    auto* o = new Translation();
    o->hittable = s->objects[child].get();
    o->displacement = Vec3(d[0], d[1], d[2]);
    return s->add_obj(o);
}
int orc_rotate(void* sp, int axis, int child, float degrees) {
    Scene* s = (Scene*)sp;
    if (!s->obj_ok(child) || axis < 0 || axis > 2) return -1;
    return s->add_obj(new Rotation(axis, s->objects[child].get(), degrees));
}
int orc_constant_medium(void* sp, int boundary, float density, int tex) {
    Scene* s = (Scene*)sp;
    if (!s->obj_ok(boundary) || !s->tex_ok(tex)) return -1;
    auto* iso = new Isotropic();
    iso->albedo = s->textures[tex].get();
    s->add_mat(iso);
    auto* o = new ConstantMedium();
    o->boundary = s->objects[boundary].get();
    o->negative_inverse_density = -1.0f / density;  // src/hittable/constant_medium.rs:27
    o->phase_function = iso;
    return s->add_obj(o);
}
int orc_list(void* sp, const int* ids, int n) {
    Scene* s = (Scene*)sp;
    auto* o = new List();
    for (int i = 0; i < n; ++i) {
        if (!s->obj_ok(ids[i])) { delete o; return -1; }
        o->objects.push_back(s->objects[ids[i]].get());
    }
    return s->add_obj(o);
}
int orc_bvh(void* sp, const int* ids, int n, float t0, float t1) {
    Scene* s = (Scene*)sp;
    if (n <= 0) return -1;  // reference panics "no elements in scene" (bvh_node.rs:38)
    std::vector<const Hittable*> objs;
    for (int i = 0; i < n; ++i) {
        if (!s->obj_ok(ids[i])) return -1;
        objs.push_back(s->objects[ids[i]].get());
    }
    return s->add_obj(new BvhNode(std::move(objs), t0, t1));
}
int orc_scene_commit(void* sp, int root) {
    Scene* s = (Scene*)sp;
    if (!s->obj_ok(root)) return -1;
    s->world = s->objects[root].get();
    return 0;
}
uint32_t orc_scene_count(void* sp) { return ((Scene*)sp)->world ? ((Scene*)sp)->world->count() : 0; }

// DFS left->right leaf object ids of a BvhNode object (topology fixture).
int orc_bvh_leaf_order(void* sp, int bvh, int* out, int cap) {
    Scene* s = (Scene*)sp;
    if (!s->obj_ok(bvh)) return -1;
    auto* b = dynamic_cast<const BvhNode*>(s->objects[bvh].get());
    if (!b) return -1;
    std::vector<int> v;
    b->leaf_order(v);
    for (size_t i = 0; i < v.size() && (int)i < cap; ++i) out[i] = v[i];
    return (int)v.size();
}
int orc_bvh_node_count(void* sp, int bvh) {
    Scene* s = (Scene*)sp;
    if (!s->obj_ok(bvh)) return -1;
    auto* b = dynamic_cast<const BvhNode*>(s->objects[bvh].get());
    return b ? (int)b->node_count() : -1;
}
int orc_bounding_box(void* sp, int obj, float* out6) {
    Scene* s = (Scene*)sp;
    if (!s->obj_ok(obj)) return -1;
    Aabb b;
    if (!s->objects[obj]->bounding_box(0.0f, 1.0f, b)) return -2;
    out6[0] = b.minimum.x; out6[1] = b.minimum.y; out6[2] = b.minimum.z;
    out6[3] = b.maximum.x; out6[4] = b.maximum.y; out6[5] = b.maximum.z;
    return 0;
}

static Camera make_camera(const orc_camera_desc* d) {
    return Camera(Vec3(d->look_from[0], d->look_from[1], d->look_from[2]),
                  Vec3(d->look_at[0], d->look_at[1], d->look_at[2]), d->vfov, d->aperture, d->focus_dist, d->time0,
                  d->time1, d->width, d->height);
}
static void put3(float* o, Vec3 v) { o[0] = v.x; o[1] = v.y; o[2] = v.z; }

void orc_camera_init(const orc_camera_desc* d, orc_camera_state* out) {
    Camera c = make_camera(d);
    put3(out->origin, c.origin);
    put3(out->lower_left_corner, c.lower_left_corner);
    put3(out->horizontal, c.horizontal);
    put3(out->vertical, c.vertical);
    put3(out->u, c.u);
    put3(out->v, c.v);
    put3(out->w, c.w);
    out->lens_radius = c.lens_radius;
    out->time0 = c.time_0;
    out->time1 = c.time_1;
}

// Camera::get_ray with injected uniforms (direct disk sampler): rays[i] from (s,t,u_lens1,u_lens2,u_time).
void orc_camera_rays(const orc_camera_desc* d, const float* stuuu, int n, orc_ray* out) {
    Camera c = make_camera(d);
    for (int i = 0; i < n; ++i) {
        const float* p = stuuu + 5 * i;
        Ray r = c.get_ray(p[0], p[1], p + 2);
        put3(out[i].o, r.origin);
        put3(out[i].d, r.direction);
        out[i].time = r.time;
        out[i].tmin = 0.001f;
        out[i].tmax = INF_F;
    }
}

// world.hit on explicit rays; xi[i] is the uniform every ConstantMedium draw on ray i returns.
int orc_trace_hits(void* sp, const orc_ray* rays, int n, const float* xi, orc_hit* out, int aabb_mode) {
    Scene* s = (Scene*)sp;
    if (!s->world) return -1;
    Rng rng;
    rng.seed(1);
    rng.injected = true;
    tls_rng = &rng;
    g_aabb_mode = aabb_mode;
    for (int i = 0; i < n; ++i) {
        rng.injected_value = xi ? xi[i] : 0.5f;
        Ray r{Vec3(rays[i].o[0], rays[i].o[1], rays[i].o[2]), Vec3(rays[i].d[0], rays[i].d[1], rays[i].d[2]),
              rays[i].time};
        HitRecord rec;
        bool h = s->world->hit(r, rays[i].tmin, rays[i].tmax, rec);
        orc_hit& o = out[i];
        std::memset(&o, 0, sizeof(o));
        o.material_id = -1;
        o.prim_id = -1;
        if (h) {
            o.hit = 1;
            o.t = rec.t;
            put3(o.p, rec.point);
            put3(o.n, rec.normal);
            o.u = rec.u;
            o.v = rec.v;
            o.front_face = rec.front_face ? 1 : 0;
            o.material_id = rec.material->id;
            o.prim_id = rec.prim_id;
            o.face = rec.face;
        }
    }
    g_aabb_mode = 0;
    tls_rng = nullptr;
    return 0;
}

// Texture::value on explicit (u,v,p) tuples: uvp = n x 5 floats, out = n x 3.
int orc_tex_value(void* sp, int tex, const float* uvp, int n, float* out) {
    Scene* s = (Scene*)sp;
    if (!s->tex_ok(tex)) return -1;
    for (int i = 0; i < n; ++i) {
        const float* q = uvp + 5 * i;
        put3(out + 3 * i, s->textures[tex]->value(q[0], q[1], Vec3(q[2], q[3], q[4])));
    }
    return 0;
}

// Material::scatter / emitted with the direct samplers under injected uniforms u4[4].
// in: ray (o,d,time), hit record; out: did_scatter, attenuation, scattered (o,d,time), emitted.
struct orc_scatter_out {
    int32_t scattered;
    float attenuation[3], o[3], d[3], time, emitted[3];
};
int orc_scatter_direct(void* sp, const orc_ray* rays, const orc_hit* hits, const float* u4, int n,
                       orc_scatter_out* out) {
    Scene* s = (Scene*)sp;
    for (int i = 0; i < n; ++i) {
        std::memset(&out[i], 0, sizeof(out[i]));
        if (!hits[i].hit || !s->mat_ok(hits[i].material_id)) continue;
        const Material* m = s->materials[hits[i].material_id].get();
        Ray r{Vec3(rays[i].o[0], rays[i].o[1], rays[i].o[2]), Vec3(rays[i].d[0], rays[i].d[1], rays[i].d[2]),
              rays[i].time};
        HitRecord rec;
        rec.point = Vec3(hits[i].p[0], hits[i].p[1], hits[i].p[2]);
        rec.normal = Vec3(hits[i].n[0], hits[i].n[1], hits[i].n[2]);
        rec.t = hits[i].t; rec.u = hits[i].u; rec.v = hits[i].v;
        rec.front_face = hits[i].front_face != 0;
        rec.material = m;
        Vec3 att;
        Ray sc{Vec3(), Vec3(), 0.0f};
        put3(out[i].emitted, m->emitted(rec.u, rec.v, rec.point));
        bool did = m->scatter(r, rec, att, sc, u4 + 4 * i);
        out[i].scattered = did ? 1 : 0;
        if (did) {
            put3(out[i].attenuation, att);
            put3(out[i].o, sc.origin);
            put3(out[i].d, sc.direction);
            out[i].time = sc.time;
        }
    }
    return 0;
}

// Sampler moments for the "rejection ≡ direct" distribution test: kind 0 ball, 1 unit vector, 2 disk;
// mode 0 rejection (reference), 1 direct.  out = n x 3.
void orc_sample(int kind, int mode, uint64_t seed, int n, float* out) {
    Rng rng;
    rng.seed(seed);
    tls_rng = &rng;
    for (int i = 0; i < n; ++i) {
        Vec3 p;
        if (mode == 0) p = kind == 0 ? random_in_unit_sphere() : (kind == 1 ? random_unit_vector() : random_in_unit_disk());
        else {
            float a = rng.gen_f32(), b = rng.gen_f32(), c = rng.gen_f32();
            p = kind == 0 ? direct_in_unit_sphere(a, b, c) : (kind == 1 ? direct_unit_vector(a, b) : direct_in_unit_disk(a, b));
        }
        put3(out + 3 * i, p);
    }
    tls_rng = nullptr;
}

// Application::render (src/application.rs:393-475): tile -> pixel -> sample loop.  One worker thread per
// host core pulls 80x80 tiles (the reference spawns one tokio task per tile on a multi-thread runtime).
// Output: sum_rgb = linear f32 per-pixel sum of ray_color over `samples` (accumulated in f32 in sample
// order, as the reference does); the reference's pixel is sqrt(sum * (1/samples)), alpha 1.
// Rows are bottom-up (row 0 = v≈0), as in the reference (Q11).  sumsq_rgb (optional) = per-channel sum of
// squares accumulated in f64 (for noise-floor estimates; not in the reference).
int orc_render(void* sp, const orc_camera_desc* cd, const orc_render_desc* rd, float* sum_rgb, float* sumsq_rgb,
               orc_counters* counters) {
    Scene* s = (Scene*)sp;
    if (!s->world) return -1;
    const int width = rd->width, height = rd->height;
    const uint32_t tile_size = (uint32_t)(rd->tile_size > 0 ? rd->tile_size : 80);
    // src/application.rs:363-364
    const uint32_t tile_x_count = (uint32_t)std::ceil((float)width / (float)tile_size);
    const uint32_t tile_y_count = (uint32_t)std::ceil((float)height / (float)tile_size);
    const uint32_t tiles = tile_x_count * tile_y_count;
    Camera camera = make_camera(cd);
    const Vec3 background(rd->background[0], rd->background[1], rd->background[2]);
    const Hittable* world = s->world;
    const uint32_t sample_count = (uint32_t)rd->samples;
    const uint32_t depth = (uint32_t)rd->depth;
    g_aabb_mode = rd->aabb_mode;

    std::atomic<uint32_t> next_tile{0};
    std::atomic<uint32_t> done_tiles{0};
    std::vector<Counters> per_thread;
    int nthreads = rd->threads > 0 ? rd->threads : (int)std::thread::hardware_concurrency();
    if (nthreads < 1) nthreads = 1;
    per_thread.resize(nthreads);
    auto t_begin = std::chrono::steady_clock::now();

    auto worker = [&](int tid) {
        tls_cnt = Counters();
        for (;;) {
            uint32_t i = next_tile.fetch_add(1);
            if (i >= tiles) break;
            uint32_t x = i % tile_x_count, y = i / tile_x_count;
            uint32_t local_x = x * tile_size, local_y = y * tile_size;
            // src/application.rs:416-430 (edge tiles).  The reference derives the edge size through f32
            // fractions; for sizes that are exact in f32 this equals width % tile_size.
            uint32_t tile_width = tile_size, tile_height = tile_size;
            if ((uint32_t)width % tile_size != 0 && x == tile_x_count - 1) {
                float q = (float)width / (float)tile_size;
                tile_width = (uint32_t)((q - std::floor(q)) * (float)tile_size);
                // guard the reference's own f32 round-down hazard so no pixel column is dropped
                if (local_x + tile_width != (uint32_t)width) tile_width = (uint32_t)width - local_x;
            }
            if ((uint32_t)height % tile_size != 0 && y == tile_y_count - 1) {
                float q = (float)height / (float)tile_size;
                tile_height = (uint32_t)((q - std::floor(q)) * (float)tile_size);
                if (local_y + tile_height != (uint32_t)height) tile_height = (uint32_t)height - local_y;
            }
            Rng rng;
            rng.seed(rd->seed * 0x9E3779B97F4A7C15ull + (uint64_t)i * 0xD1B54A32D192ED03ull + 0x1234567ull);
            tls_rng = &rng;
            for (uint32_t p = 0; p < tile_width * tile_height; ++p) {
                uint32_t px = (p % tile_width) + local_x;
                uint32_t py = (p / tile_width) + local_y;
                Vec3 pixel_color(0, 0, 0);
                double sq[3] = {0, 0, 0};
                for (uint32_t sidx = 0; sidx < sample_count; ++sidx) {
                    // src/application.rs:444-448
                    float u = ((float)px + rng.gen_f32()) / ((float)width - 1.0f);
                    float v = ((float)py + rng.gen_f32()) / ((float)height - 1.0f);
                    Ray ray = camera.get_ray(u, v);
                    Vec3 c = ray_color(ray, background, world, depth);
                    tls_cnt.paths++;
                    pixel_color = pixel_color + c;
                    sq[0] += (double)c.x * c.x; sq[1] += (double)c.y * c.y; sq[2] += (double)c.z * c.z;
                }
                size_t o = ((size_t)py * width + px) * 3;
                sum_rgb[o] = pixel_color.x; sum_rgb[o + 1] = pixel_color.y; sum_rgb[o + 2] = pixel_color.z;
                if (sumsq_rgb) { sumsq_rgb[o] = (float)sq[0]; sumsq_rgb[o + 1] = (float)sq[1]; sumsq_rgb[o + 2] = (float)sq[2]; }
            }
            uint32_t d = done_tiles.fetch_add(1) + 1;
            if (rd->verbose) {
                double el = std::chrono::duration<double>(std::chrono::steady_clock::now() - t_begin).count();
                std::fprintf(stderr, "[oracle] tile %u/%u  %.1fs\n", d, tiles, el);
            }
        }
        per_thread[tid] = tls_cnt;
        tls_rng = nullptr;
    };
    std::vector<std::thread> pool;
    for (int t = 0; t < nthreads; ++t) pool.emplace_back(worker, t);
    for (auto& t : pool) t.join();
    double seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t_begin).count();
    g_aabb_mode = 0;
    if (counters) {
        Counters tot;
        for (auto& c : per_thread) tot.add(c);
        counters->paths = tot.paths; counters->rays = tot.rays; counters->aabb_tests = tot.aabb_tests;
        counters->sphere_tests = tot.sphere_tests; counters->rect_tests = tot.rect_tests;
        counters->medium_queries = tot.medium_queries; counters->noise_evals = tot.noise_evals;
        counters->scatters = tot.scatters; counters->seconds = seconds;
    }
    return 0;
}

// The reference's gamma resolve (src/application.rs:451-456): out RGBA = (sqrt(sum*scale), 1.0).
void orc_resolve(const float* sum_rgb, int n_pixels, int samples, float* out_rgba) {
    float scale = 1.0f / (float)samples;  // src/application.rs:403
    for (int i = 0; i < n_pixels; ++i) {
        out_rgba[4 * i + 0] = std::sqrt(sum_rgb[3 * i + 0] * scale);
        out_rgba[4 * i + 1] = std::sqrt(sum_rgb[3 * i + 1] * scale);
        out_rgba[4 * i + 2] = std::sqrt(sum_rgb[3 * i + 2] * scale);
        out_rgba[4 * i + 3] = 1.0f;
    }
}

const char* orc_version() { return "hrt-oracle 1"; }

}  // extern "C"
