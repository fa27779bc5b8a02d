// hrt_scenes.cpp — the reference's scene library behind the C ABI:
//   * hrt_make_scene: the eight scene generators of src/application.rs:497-935 and the per-scene camera / background
//     table of :132-197, with an explicit seed (the reference draws from an unseeded thread_rng, :509,820 and
//     perlin_noise.rs:24,59, so no two of its runs render the same `random` / `final` / perlin scene);
//   * hrt_scene_save / hrt_scene_load: a scene INSTANCE — every builder call, the root, camera and background — as one
//     flat file, so that a render can be repeated on the same world (the reference never stores one).
// Everything goes through the public builder calls of include/hrt.h, in the order a depth-first walk of the reference's
// object graph constructs them, all geometry arithmetic in f32 like the reference's.
#include <cmath>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/hrt.h"
#include "hrt_rng.hpp"
#include "hrt_scene.hpp"

using namespace hrt;

namespace {

struct Gen {
    hrt_scene* s;
    SceneRng rng;
    int32_t err = HRT_OK;
    Gen(hrt_scene* sc, uint64_t seed) : s(sc), rng(seed) {}
    int32_t ck(int32_t r) {
        if (r < 0 && err == HRT_OK) err = r;
        return r;
    }
    int32_t solid(float r, float g, float b) {
        const float c[3] = {r, g, b};
        return ck(hrt_tex_solid(s, c));
    }
    int32_t lambertian_solid(float r, float g, float b) { return ck(hrt_mat_lambertian(s, solid(r, g, b))); }
    int32_t light(float r, float g, float b) { return ck(hrt_mat_diffuse_light(s, solid(r, g, b))); }
    int32_t metal(float r, float g, float b, float fuzz) {
        const float c[3] = {r, g, b};
        return ck(hrt_mat_metal(s, c, fuzz));
    }
    int32_t sphere(float x, float y, float z, float r, int32_t mat) {
        const float c[3] = {x, y, z};
        return ck(hrt_sphere(s, c, r, mat));
    }
    int32_t cuboid(float x0, float y0, float z0, float x1, float y1, float z1, int32_t mat) {
        const float mn[3] = {x0, y0, z0}, mx[3] = {x1, y1, z1};
        return ck(hrt_cuboid(s, mn, mx, mat));
    }
    int32_t translate(int32_t child, float x, float y, float z) {
        const float d[3] = {x, y, z};
        return ck(hrt_translate(s, child, d));
    }
    int32_t bvh(const std::vector<int32_t>& ids) { return ck(hrt_bvh(s, ids.data(), (int32_t)ids.size(), 0.0f, 1.0f)); }
    // PerlinNoise::new (perlin_noise.rs:23-64): 256 normalised vectors from [-1,1)^3, then three Sattolo-style shuffles
    // (`gen_range(0..i)` is exclusive)
    int32_t noise(float scale) {
        std::vector<float> ranvec(256 * 3);
        for (int i = 0; i < 256; ++i) {
            const float x = rng.gen_range(-1.0f, 1.0f), y = rng.gen_range(-1.0f, 1.0f), z = rng.gen_range(-1.0f, 1.0f);
            const float inv = 1.0f / std::sqrt((x * x + y * y) + z * z);  // cgmath normalize
            ranvec[3 * i] = x * inv; ranvec[3 * i + 1] = y * inv; ranvec[3 * i + 2] = z * inv;
        }
        uint32_t perm[3][256];
        for (int k = 0; k < 3; ++k) {
            for (int i = 0; i < 256; ++i) perm[k][i] = (uint32_t)i;
            for (int i = 255; i > 0; --i) {
                const uint32_t target = rng.gen_index((uint32_t)i);
                std::swap(perm[k][i], perm[k][target]);
            }
        }
        return ck(hrt_tex_noise(s, scale, ranvec.data(), perm[0], perm[1], perm[2]));
    }
};

void set_view(hrt_scene_view* v, float fx, float fy, float fz, float ax, float ay, float az, float vfov, float aperture, float br,
              float bg, float bb) {
    if (!v) return;
    v->look_from[0] = fx; v->look_from[1] = fy; v->look_from[2] = fz;
    v->look_at[0] = ax; v->look_at[1] = ay; v->look_at[2] = az;
    v->vfov = vfov; v->aperture = aperture;
    v->focus_dist = 10.0f; v->time0 = 0.0f; v->time1 = 1.0f;  // application.rs:206-208
    v->background[0] = br; v->background[1] = bg; v->background[2] = bb;
}

// application.rs:497-565 — the *Next Week* variant: checker ground, moving diffuse spheres (Q13)
int32_t gen_random(Gen& g) {
    std::vector<int32_t> objs;
    {
        const int32_t odd = g.solid(0.2f, 0.3f, 0.1f), even = g.solid(0.9f, 0.9f, 0.9f);
        const int32_t ground = g.ck(hrt_mat_lambertian(g.s, g.ck(hrt_tex_checker(g.s, odd, even))));
        objs.push_back(g.sphere(0.0f, -1000.0f, 0.0f, 1000.0f, ground));
    }
    for (int a = -11; a < 11; ++a)
        for (int b = -11; b < 11; ++b) {
            const float choose = g.rng.gen();
            const float cx = (float)a + 0.9f * g.rng.gen();
            const float cz = (float)b + 0.9f * g.rng.gen();
            const float dx = cx - 4.0f, dy = 0.2f - 0.2f, dz = cz - 0.0f;
            if (std::sqrt((dx * dx + dy * dy) + dz * dz) > 0.9f) {
                if (choose < 0.8f) {
                    const float r = g.rng.gen(), gg = g.rng.gen(), bb = g.rng.gen();
                    const float c0[3] = {cx, 0.2f, cz};
                    const float c1[3] = {cx + 0.0f, 0.2f + g.rng.gen_range(0.0f, 0.5f), cz + 0.0f};
                    const int32_t m = g.lambertian_solid(r, gg, bb);
                    objs.push_back(g.ck(hrt_moving_sphere(g.s, c0, c1, 0.0f, 1.0f, 0.2f, m)));
                } else if (choose < 0.95f) {
                    const float r = g.rng.gen_range(0.5f, 1.0f), gg = g.rng.gen_range(0.5f, 1.0f), bb = g.rng.gen_range(0.5f, 1.0f);
                    const float fuzz = g.rng.gen_range(0.0f, 0.5f);
                    objs.push_back(g.sphere(cx, 0.2f, cz, 0.2f, g.metal(r, gg, bb, fuzz)));
                } else {
                    objs.push_back(g.sphere(cx, 0.2f, cz, 0.2f, g.ck(hrt_mat_dielectric(g.s, 1.5f))));
                }
            }
        }
    objs.push_back(g.sphere(0.0f, 1.0f, 0.0f, 1.0f, g.ck(hrt_mat_dielectric(g.s, 1.5f))));
    objs.push_back(g.sphere(-4.0f, 1.0f, 0.0f, 1.0f, g.lambertian_solid(0.4f, 0.2f, 0.1f)));
    objs.push_back(g.sphere(4.0f, 1.0f, 0.0f, 1.0f, g.metal(0.7f, 0.6f, 0.5f, 0.0f)));
    return g.bvh(objs);
}

// application.rs:567-587
int32_t gen_two_spheres(Gen& g) {
    const int32_t odd = g.solid(0.2f, 0.3f, 0.1f), even = g.solid(0.9f, 0.9f, 0.9f);
    const int32_t checker = g.ck(hrt_mat_lambertian(g.s, g.ck(hrt_tex_checker(g.s, odd, even))));
    return g.bvh({g.sphere(0.0f, -10.0f, 0.0f, 10.0f, checker), g.sphere(0.0f, 10.0f, 0.0f, 10.0f, checker)});
}

// application.rs:589-602
int32_t gen_two_perlin_spheres(Gen& g) {
    const int32_t m = g.ck(hrt_mat_lambertian(g.s, g.noise(4.0f)));
    return g.bvh({g.sphere(0.0f, -1000.0f, 0.0f, 1000.0f, m), g.sphere(0.0f, 2.0f, 0.0f, 2.0f, m)});
}

// application.rs:604-612
int32_t gen_earth(Gen& g, const uint8_t* img, uint32_t w, uint32_t h, uint32_t comps) {
    const int32_t tex = g.ck(hrt_tex_image(g.s, img, w, h, comps));
    return g.bvh({g.sphere(0.0f, 0.0f, 0.0f, 2.0f, g.ck(hrt_mat_lambertian(g.s, tex)))});
}

// application.rs:614-637
int32_t gen_simple_light(Gen& g) {
    const int32_t m = g.ck(hrt_mat_lambertian(g.s, g.noise(4.0f)));
    const int32_t a = g.sphere(0.0f, -1000.0f, 0.0f, 1000.0f, m), b = g.sphere(0.0f, 2.0f, 0.0f, 2.0f, m);
    const int32_t r = g.ck(hrt_rect(g.s, HRT_PLANE_XY, 3.0f, 5.0f, 1.0f, 3.0f, -2.0f, g.light(4.0f, 4.0f, 4.0f)));
    return g.bvh({a, b, r});
}

// the five walls and the light of application.rs:639-703 / :723-787; `white` is shared with the boxes
void cornell_walls(Gen& g, std::vector<int32_t>& objs, int32_t& white) {
    objs.push_back(g.ck(hrt_rect(g.s, HRT_PLANE_YZ, 0.0f, 555.0f, 0.0f, 555.0f, 555.0f, g.lambertian_solid(0.12f, 0.45f, 0.15f))));
    objs.push_back(g.ck(hrt_rect(g.s, HRT_PLANE_YZ, 0.0f, 555.0f, 0.0f, 555.0f, 0.0f, g.lambertian_solid(0.65f, 0.05f, 0.05f))));
    objs.push_back(g.ck(hrt_rect(g.s, HRT_PLANE_ZX, 213.0f, 343.0f, 227.0f, 332.0f, 554.0f, g.light(15.0f, 15.0f, 15.0f))));
    white = g.lambertian_solid(0.73f, 0.73f, 0.73f);
    objs.push_back(g.ck(hrt_rect(g.s, HRT_PLANE_ZX, 0.0f, 555.0f, 0.0f, 555.0f, 0.0f, white)));
    objs.push_back(g.ck(hrt_rect(g.s, HRT_PLANE_ZX, 0.0f, 555.0f, 0.0f, 555.0f, 555.0f, white)));
    objs.push_back(g.ck(hrt_rect(g.s, HRT_PLANE_XY, 0.0f, 555.0f, 0.0f, 555.0f, 555.0f, white)));
}
int32_t cornell_block(Gen& g, float sx, float sy, float sz, float angle, float tx, float ty, float tz, int32_t white) {
    const int32_t c = g.cuboid(0.0f, 0.0f, 0.0f, sx, sy, sz, white);
    return g.translate(g.ck(hrt_rotate(g.s, HRT_AXIS_Y, c, angle)), tx, ty, tz);
}
// application.rs:639-721
int32_t gen_cornell(Gen& g) {
    std::vector<int32_t> objs;
    int32_t white;
    cornell_walls(g, objs, white);
    objs.push_back(cornell_block(g, 165.0f, 330.0f, 165.0f, 15.0f, 265.0f, 0.0f, 295.0f, white));
    objs.push_back(cornell_block(g, 165.0f, 165.0f, 165.0f, -18.0f, 130.0f, 0.0f, 65.0f, white));
    return g.bvh(objs);
}
// application.rs:723-815
int32_t gen_cornell_smoke(Gen& g) {
    std::vector<int32_t> objs;
    int32_t white;
    cornell_walls(g, objs, white);
    const int32_t b1 = cornell_block(g, 165.0f, 330.0f, 165.0f, 15.0f, 265.0f, 0.0f, 295.0f, white);
    const int32_t black = g.solid(0.0f, 0.0f, 0.0f);
    objs.push_back(g.ck(hrt_constant_medium(g.s, b1, 0.01f, black)));
    const int32_t b2 = cornell_block(g, 165.0f, 165.0f, 165.0f, -18.0f, 130.0f, 0.0f, 65.0f, white);
    const int32_t fog = g.solid(1.0f, 1.0f, 1.0f);
    objs.push_back(g.ck(hrt_constant_medium(g.s, b2, 0.01f, fog)));
    return g.bvh(objs);
}

// application.rs:817-935 — 20 x 20 = 400 ground boxes, 1000 small spheres, the top level is a BvhNode (Q14)
int32_t gen_final(Gen& g, const uint8_t* img, uint32_t w, uint32_t h, uint32_t comps) {
    // the reference draws while it constructs: box heights, then the perlin tables, then the sphere centres
    float y1[400];
    for (int i = 0; i < 400; ++i) y1[i] = g.rng.gen_range(1.0f, 101.0f);
    std::vector<int32_t> objs;
    {
        const int32_t ground = g.lambertian_solid(0.48f, 0.83f, 0.53f);
        std::vector<int32_t> boxes;
        for (int i = 0; i < 20; ++i)
            for (int j = 0; j < 20; ++j) {
                const float wd = 100.0f;
                const float x0 = -1000.0f + (float)i * wd, z0 = -1000.0f + (float)j * wd;
                boxes.push_back(g.cuboid(x0, 0.0f, z0, x0 + wd, y1[i * 20 + j], z0 + wd, ground));
            }
        objs.push_back(g.bvh(boxes));
    }
    objs.push_back(g.ck(hrt_rect(g.s, HRT_PLANE_ZX, 123.0f, 423.0f, 147.0f, 412.0f, 554.0f, g.light(7.0f, 7.0f, 7.0f))));
    {
        const float c0[3] = {400.0f, 400.0f, 200.0f}, c1[3] = {400.0f + 30.0f, 400.0f + 0.0f, 200.0f + 0.0f};
        objs.push_back(g.ck(hrt_moving_sphere(g.s, c0, c1, 0.0f, 1.0f, 50.0f, g.lambertian_solid(0.7f, 0.3f, 0.1f))));
    }
    objs.push_back(g.sphere(260.0f, 150.0f, 45.0f, 50.0f, g.ck(hrt_mat_dielectric(g.s, 1.5f))));
    objs.push_back(g.sphere(0.0f, 150.0f, 145.0f, 50.0f, g.metal(0.8f, 0.8f, 0.9f, 1.0f)));
    const int32_t glass = g.ck(hrt_mat_dielectric(g.s, 1.5f));
    objs.push_back(g.sphere(360.0f, 150.0f, 145.0f, 70.0f, glass));
    {
        const int32_t boundary = g.sphere(360.0f, 150.0f, 145.0f, 70.0f, glass);
        const int32_t blue = g.solid(0.2f, 0.4f, 0.9f);
        objs.push_back(g.ck(hrt_constant_medium(g.s, boundary, 0.2f, blue)));
    }
    {
        const int32_t boundary = g.sphere(0.0f, 0.0f, 0.0f, 5000.0f, g.ck(hrt_mat_dielectric(g.s, 1.5f)));
        const int32_t white = g.solid(1.0f, 1.0f, 1.0f);
        objs.push_back(g.ck(hrt_constant_medium(g.s, boundary, 0.0001f, white)));
    }
    objs.push_back(g.sphere(400.0f, 200.0f, 400.0f, 100.0f, g.ck(hrt_mat_lambertian(g.s, g.ck(hrt_tex_image(g.s, img, w, h, comps))))));
    objs.push_back(g.sphere(220.0f, 280.0f, 300.0f, 80.0f, g.ck(hrt_mat_lambertian(g.s, g.noise(0.1f)))));
    {
        // (the 1000 centres are drawn after the perlin tables, which g.noise just drew)
        std::vector<int32_t> balls;
        int32_t white = -1;
        for (int i = 0; i < 1000; ++i) {
            const float cx = g.rng.gen_range(0.0f, 165.0f), cy = g.rng.gen_range(0.0f, 165.0f), cz = g.rng.gen_range(0.0f, 165.0f);
            if (white < 0) white = g.lambertian_solid(0.73f, 0.73f, 0.73f);
            balls.push_back(g.sphere(cx, cy, cz, 10.0f, white));
        }
        objs.push_back(g.translate(g.ck(hrt_rotate(g.s, HRT_AXIS_Y, g.bvh(balls), 15.0f)), -100.0f, 270.0f, 395.0f));
    }
    return g.bvh(objs);
}

// ---- scene-instance files ----------------------------------------------------------------------------------------
const char kMagic[8] = {'H', 'R', 'T', 'S', 'C', 'N', '1', 0};
struct FileHeader {
    char magic[8];
    uint32_t n_textures, n_materials, n_objects, n_noise, n_images;
    int32_t root, bvh_builder;
    hrt_scene_view view;
};
template <typename T>
bool put(FILE* f, const T* p, size_t n) { return n == 0 || fwrite(p, sizeof(T), n, f) == n; }
template <typename T>
bool get(FILE* f, T* p, size_t n) { return n == 0 || fread(p, sizeof(T), n, f) == n; }
struct ObjRecord {  // fixed-size part of an Obj (children follow)
    int32_t kind, plane_or_axis, mat, child;
    float c0[3], c1[3], r, t0, t1, a0, a1, b0, b1, k, density_or_angle;
    uint32_t n_children;
};

}  // namespace

extern "C" {

int32_t hrt_make_scene(hrt_scene* s, const char* name, uint64_t seed, const uint8_t* image, uint32_t image_width,
                       uint32_t image_height, uint32_t image_components, int32_t* root_out, hrt_scene_view* view_out) {
    if (!s || !name || !root_out) return fail(HRT_ERR_INVALID, "make_scene: null argument");
    if (s->committed) return fail(HRT_ERR_STATE, "scene is already committed (immutable)");
    Gen g(s, seed);
    const std::string n(name);
    int32_t root = -1;
    // camera and background per scene: application.rs:132-197
    if (n == "random") { root = gen_random(g); set_view(view_out, 13, 2, 3, 0, 0, 0, 20.0f, 0.1f, 0.7f, 0.8f, 1.0f); }
    else if (n == "two-spheres") { root = gen_two_spheres(g); set_view(view_out, 13, 2, 3, 0, 0, 0, 20.0f, 0.0f, 0.7f, 0.8f, 1.0f); }
    else if (n == "two-perlin-spheres") { root = gen_two_perlin_spheres(g); set_view(view_out, 13, 2, 3, 0, 0, 0, 20.0f, 0.0f, 0.7f, 0.8f, 1.0f); }
    else if (n == "earth") { root = gen_earth(g, image, image_width, image_height, image_components); set_view(view_out, 13, 2, 3, 0, 0, 0, 20.0f, 0.0f, 0.7f, 0.8f, 1.0f); }
    else if (n == "simple-light") { root = gen_simple_light(g); set_view(view_out, 26, 3, 6, 0, 2, 0, 20.0f, 0.0f, 0, 0, 0); }
    else if (n == "cornell") { root = gen_cornell(g); set_view(view_out, 278, 278, -800, 278, 278, 0, 40.0f, 0.0f, 0, 0, 0); }
    else if (n == "cornell-smoke") { root = gen_cornell_smoke(g); set_view(view_out, 278, 278, -800, 278, 278, 0, 40.0f, 0.0f, 0, 0, 0); }
    else if (n == "final") { root = gen_final(g, image, image_width, image_height, image_components); set_view(view_out, 478, 278, -600, 278, 278, 0, 40.0f, 0.0f, 0, 0, 0); }
    else return fail(HRT_ERR_INVALID, "make_scene: unknown scene '" + n + "' (src/arguments.rs:10-19: random, two-spheres, two-perlin-spheres, earth, simple-light, cornell, cornell-smoke, final)");
    if (g.err != HRT_OK) return g.err;  // hrt_last_error() holds the builder's message
    *root_out = root;
    return HRT_OK;
}

int32_t hrt_scene_save(const hrt_scene* s, int32_t root, const hrt_scene_view* view, const char* path) {
    if (!s || !path) return fail(HRT_ERR_INVALID, "scene_save: null argument");
    if (root < 0 || (size_t)root >= s->objects.size()) return fail(HRT_ERR_INVALID, "scene_save: unknown root id");
    FILE* f = fopen(path, "wb");
    if (!f) return fail(HRT_ERR_INVALID, std::string("scene_save: cannot open ") + path);
    FileHeader h;
    std::memset(&h, 0, sizeof(h));
    std::memcpy(h.magic, kMagic, 8);
    h.n_textures = (uint32_t)s->textures.size(); h.n_materials = (uint32_t)s->materials.size();
    h.n_objects = (uint32_t)s->objects.size(); h.n_noise = (uint32_t)s->noise_tables.size(); h.n_images = (uint32_t)s->images.size();
    h.root = root; h.bvh_builder = s->bvh_builder;
    if (view) h.view = *view;
    bool ok = put(f, &h, 1) && put(f, s->textures.data(), s->textures.size()) && put(f, s->materials.data(), s->materials.size()) &&
              put(f, s->noise_tables.data(), s->noise_tables.size());
    for (const ImageData& img : s->images) {
        const uint32_t wh[2] = {img.width, img.height};
        ok = ok && put(f, wh, 2) && put(f, img.rgba.data(), img.rgba.size());
    }
    for (const Obj& o : s->objects) {
        ObjRecord r;
        std::memset(&r, 0, sizeof(r));
        r.kind = o.kind; r.plane_or_axis = o.plane_or_axis; r.mat = o.mat; r.child = o.child;
        std::memcpy(r.c0, o.c0, 12); std::memcpy(r.c1, o.c1, 12);
        r.r = o.r; r.t0 = o.t0; r.t1 = o.t1; r.a0 = o.a0; r.a1 = o.a1; r.b0 = o.b0; r.b1 = o.b1; r.k = o.k;
        r.density_or_angle = o.kind == OBJ_ROTATE ? o.angle_degrees : (o.kind == OBJ_MEDIUM ? o.density : 0.0f);
        r.n_children = (uint32_t)o.children.size();
        ok = ok && put(f, &r, 1) && put(f, o.children.data(), o.children.size());
    }
    ok = (fclose(f) == 0) && ok;
    return ok ? HRT_OK : fail(HRT_ERR_INVALID, std::string("scene_save: write to ") + path + " failed");
}

int32_t hrt_scene_load(const char* path, hrt_scene** out, int32_t* root_out, hrt_scene_view* view_out) {
    if (!path || !out || !root_out) return fail(HRT_ERR_INVALID, "scene_load: null argument");
    FILE* f = fopen(path, "rb");
    if (!f) return fail(HRT_ERR_INVALID, std::string("scene_load: cannot open ") + path);
    FileHeader h;
    hrt_scene* s = nullptr;
    auto bail = [&](const std::string& why) {
        fclose(f);
        if (s) hrt_scene_destroy(s);
        return fail(HRT_ERR_INVALID, "scene_load: " + why);
    };
    if (!get(f, &h, 1) || std::memcmp(h.magic, kMagic, 8) != 0) return bail("not a scene-instance file");
    if (h.n_textures > (1u << 24) || h.n_materials > (1u << 24) || h.n_objects > (1u << 24) || h.n_noise > 4096 || h.n_images > (uint32_t)kMaxImages)
        return bail("implausible table sizes");
    if (hrt_scene_create(&s) != HRT_OK) { fclose(f); return HRT_ERR_INVALID; }
    s->bvh_builder = h.bvh_builder;
    // textures / materials / perlin tables / images are plain tables: restore them as they were (ids are positions)
    s->textures.resize(h.n_textures); s->materials.resize(h.n_materials); s->noise_tables.resize(h.n_noise);
    if (!get(f, s->textures.data(), h.n_textures) || !get(f, s->materials.data(), h.n_materials) || !get(f, s->noise_tables.data(), h.n_noise))
        return bail("truncated tables");
    for (uint32_t i = 0; i < h.n_images; ++i) {
        uint32_t wh[2];
        if (!get(f, wh, 2) || wh[0] > 32768 || wh[1] > 32768) return bail("bad image header");
        ImageData img;
        img.width = wh[0]; img.height = wh[1];
        img.rgba.resize((size_t)wh[0] * wh[1] * 4);
        if (!get(f, img.rgba.data(), img.rgba.size())) return bail("truncated image");
        s->images.push_back(std::move(img));
    }
    for (const Texture& t : s->textures) {
        const bool ok = (t.kind == TEX_SOLID) || (t.kind == TEX_CHECKER && t.i0 >= 0 && t.i1 >= 0 && (uint32_t)t.i0 < h.n_textures && (uint32_t)t.i1 < h.n_textures) ||
                        (t.kind == TEX_NOISE && t.i0 >= 0 && (uint32_t)t.i0 < h.n_noise) || (t.kind == TEX_IMAGE && t.i0 < (int32_t)h.n_images);
        if (!ok) return bail("texture table refers outside the file");
    }
    for (const Material& m : s->materials)
        if (m.kind < MAT_LAMBERTIAN || m.kind > MAT_ISOTROPIC || (m.kind != MAT_METAL && m.kind != MAT_DIELECTRIC && (m.tex < 0 || (uint32_t)m.tex >= h.n_textures)))
            return bail("material table refers outside the file");
    // hittables are REPLAYED through the builder calls, in creation order: every bounding box and every BvhNode is
    // rebuilt exactly as when the scene was first described
    for (uint32_t i = 0; i < h.n_objects; ++i) {
        ObjRecord r;
        if (!get(f, &r, 1) || r.n_children > h.n_objects) return bail("truncated object table");
        std::vector<int32_t> children(r.n_children);
        if (!get(f, children.data(), children.size())) return bail("truncated object table");
        int32_t id = HRT_ERR_INVALID;
        switch (r.kind) {
            case OBJ_SPHERE: id = hrt_sphere(s, r.c0, r.r, r.mat); break;
            case OBJ_MSPHERE: id = hrt_moving_sphere(s, r.c0, r.c1, r.t0, r.t1, r.r, r.mat); break;
            case OBJ_RECT: id = hrt_rect(s, r.plane_or_axis, r.a0, r.a1, r.b0, r.b1, r.k, r.mat); break;
            case OBJ_CUBOID: id = hrt_cuboid(s, r.c0, r.c1, r.mat); break;
            case OBJ_TRANSLATE: id = hrt_translate(s, r.child, r.c0); break;
            case OBJ_ROTATE: id = hrt_rotate(s, r.plane_or_axis, r.child, r.density_or_angle); break;
            case OBJ_MEDIUM: {
                // hrt_constant_medium allocated the Isotropic material itself; it is already in the restored table
                if (r.mat < 0 || (uint32_t)r.mat >= h.n_materials || r.child < 0 || (uint32_t)r.child >= i) return bail("bad medium record");
                id = hrt::add_medium_with_material(s, r.child, r.density_or_angle, r.mat);
                break;
            }
            case OBJ_LIST: id = hrt_list(s, children.data(), (int32_t)children.size()); break;
            case OBJ_BVH: id = hrt_bvh(s, children.data(), (int32_t)children.size(), r.t0, r.t1); break;
            default: return bail("unknown object kind");
        }
        if (id != (int32_t)i) return bail(std::string("object ") + std::to_string(i) + " could not be rebuilt: " + hrt_last_error());
    }
    fclose(f);
    if (h.root < 0 || (uint32_t)h.root >= h.n_objects) { hrt_scene_destroy(s); return fail(HRT_ERR_INVALID, "scene_load: bad root"); }
    *out = s;
    *root_out = h.root;
    if (view_out) *view_out = h.view;
    return HRT_OK;
}

}  // extern "C"
