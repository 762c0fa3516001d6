// world.hpp — host-side mirror of the reference's scene-construction API and scene library,
// written against the C-ABI of include/rtb200.h.
//
// The reference builds scenes with Rust constructors (Sphere::new, HittableList::add,
// BvhNode::from_list, Lambertian::new, Checker::from_colors, Camera::new, ...; SURVEY.md
// Appendix C) and a scene switch get_world_cam(id) (src/world.rs:876-1179).  `Builder` keeps those
// names and argument orders; each call records the object through one RTB_FN(...) entry point and
// returns its id (the analogue of the Arc<Box<dyn Trait>>).  Compiled into librtb200.so it targets
// rt_*; compiled into the oracle (-DRTB_PREFIX_ORC) it targets orc_*, so both sides receive the
// same call sequence and the same scene randomness (SceneRng, seeded — the reference is unseeded).
#pragma once
#include <cmath>
#include <cstdio>
#include <string>
#include <vector>

#include "rtb200.h"
#include "scene_rng.hpp"

namespace rtb {

struct V3 {
    double x, y, z;
    V3() : x(0), y(0), z(0) {}
    V3(double a, double b, double c) : x(a), y(b), z(c) {}
};
inline V3 operator+(V3 a, V3 b) { return V3(a.x + b.x, a.y + b.y, a.z + b.z); }
inline V3 operator-(V3 a, V3 b) { return V3(a.x - b.x, a.y - b.y, a.z - b.z); }
inline V3 operator*(V3 a, V3 b) { return V3(a.x * b.x, a.y * b.y, a.z * b.z); }
inline double length(V3 a) { return std::sqrt(a.x * a.x + a.y * a.y + a.z * a.z); }

// HittableList::new() / add()  [ref: src/hit.rs:646-652]
struct HittableList {
    std::vector<int32_t> objects;
    void add(int32_t id) { objects.push_back(id); }
};

struct Builder {
    rt_scene* s;
    int32_t err = 0; // first negative status seen
    explicit Builder(rt_scene* scene) : s(scene) {}
    int32_t chk(int32_t r) { if (r < 0 && err == 0) err = r; return r; }

    // ---- textures [ref: src/texture.rs]
    int32_t SolidColor_new(V3 c) { const double a[3] = {c.x, c.y, c.z}; return chk(RTB_FN(tex_solid)(s, a)); }
    int32_t Checker_new(int32_t even, int32_t odd) { return chk(RTB_FN(tex_checker)(s, even, odd)); }
    int32_t Checker_from_colors(V3 even, V3 odd) { return Checker_new(SolidColor_new(even), SolidColor_new(odd)); }
    int32_t Noise_new(double scale, uint64_t table_seed) {
        PerlinTables t;
        perlin_generate(table_seed, t);
        return chk(RTB_FN(tex_noise)(s, scale, t.ranvec, t.perm_x, t.perm_y, t.perm_z, 0));
    }
    int32_t Image_from_texels(int32_t w, int32_t h, const std::vector<double>& rgb) { return chk(RTB_FN(tex_image)(s, w, h, rgb.data())); }
    int32_t Image_from_ppm(const char* name) { return chk(RTB_FN(tex_image_ppm)(s, name)); }

    // ---- materials [ref: src/hit.rs:992-1152]
    int32_t Lambertian_from_pointer(int32_t tex) { return chk(RTB_FN(mat_lambertian)(s, tex)); }
    int32_t Lambertian_new(V3 albedo) { return Lambertian_from_pointer(SolidColor_new(albedo)); }
    int32_t Metal_new(V3 albedo, double fuzz) { const double a[3] = {albedo.x, albedo.y, albedo.z}; return chk(RTB_FN(mat_metal)(s, a, fuzz)); }
    int32_t Dielectric_new(double ir) { return chk(RTB_FN(mat_dielectric)(s, ir)); }
    int32_t DiffuseLight_from_pointer(int32_t tex) { return chk(RTB_FN(mat_diffuse_light)(s, tex)); }
    int32_t DiffuseLight_new(V3 c) { return DiffuseLight_from_pointer(SolidColor_new(c)); }

    // ---- hittables [ref: src/hit.rs, src/bvh.rs, src/model.rs]
    int32_t Sphere_new(V3 c, double r, int32_t mat) { const double a[3] = {c.x, c.y, c.z}; return chk(RTB_FN(sphere)(s, a, r, mat)); }
    int32_t MovingSphere_new(V3 c0, V3 c1, double t0, double t1, double r, int32_t mat) {
        const double a[3] = {c0.x, c0.y, c0.z}, b[3] = {c1.x, c1.y, c1.z};
        return chk(RTB_FN(moving_sphere)(s, a, b, t0, t1, r, mat));
    }
    int32_t GravitySphere_new(V3 start, double t0, double r, int32_t mat) {
        const double a[3] = {start.x, start.y, start.z};
        return chk(RTB_FN(gravity_sphere)(s, a, t0, r, mat));
    }
    int32_t XyRect_new(double x0, double x1, double y0, double y1, double k, int32_t mat) { return chk(RTB_FN(xy_rect)(s, x0, x1, y0, y1, k, mat)); }
    int32_t XzRect_new(double x0, double x1, double z0, double z1, double k, int32_t mat) { return chk(RTB_FN(xz_rect)(s, x0, x1, z0, z1, k, mat)); }
    int32_t YzRect_new(double y0, double y1, double z0, double z1, double k, int32_t mat) { return chk(RTB_FN(yz_rect)(s, y0, y1, z0, z1, k, mat)); }
    int32_t RectPrism_new(V3 p0, V3 p1, int32_t mat) {
        const double a[3] = {p0.x, p0.y, p0.z}, b[3] = {p1.x, p1.y, p1.z};
        return chk(RTB_FN(box)(s, a, b, mat));
    }
    int32_t Triangle_new(V3 v0, V3 v1, V3 v2, int32_t mat) {
        const double a[3] = {v0.x, v0.y, v0.z}, b[3] = {v1.x, v1.y, v1.z}, c[3] = {v2.x, v2.y, v2.z};
        return chk(RTB_FN(triangle)(s, a, b, c, mat));
    }
    int32_t List(const HittableList& l) { return chk(RTB_FN(list)(s, l.objects.data(), (int32_t)l.objects.size())); }
    int32_t BvhNode_from_list(const HittableList& l, double t0, double t1) {
        return chk(RTB_FN(bvh)(s, l.objects.data(), (int32_t)l.objects.size(), t0, t1));
    }
    int32_t Translate_new(V3 off, int32_t obj) { const double a[3] = {off.x, off.y, off.z}; return chk(RTB_FN(translate)(s, a, obj)); }
    int32_t RotateY_new(double angle, int32_t obj) { return chk(RTB_FN(rotate_y)(s, angle, obj)); }
    int32_t ConstantMedium_from_color(V3 c, double d, int32_t boundary) {
        const double a[3] = {c.x, c.y, c.z};
        return chk(RTB_FN(constant_medium)(s, a, d, boundary));
    }
    // TriangleModel::load_from_file(path, scale).to_hittable()  [ref: src/model.rs:13-76]
    int32_t TriangleModel_load(const char* path, double scale, int32_t mat) { return chk(RTB_FN(ply_load)(s, path, scale, mat)); }

    // Camera::new  [ref: src/camera.rs:20-30]
    void Camera_new(V3 lookfrom, V3 lookat, V3 vup, double vfov, double aspect, double aperture, double focus, double t1, double t2) {
        const double a[3] = {lookfrom.x, lookfrom.y, lookfrom.z}, b[3] = {lookat.x, lookat.y, lookat.z}, c[3] = {vup.x, vup.y, vup.z};
        chk(RTB_FN(scene_set_camera)(s, a, b, c, vfov, aspect, aperture, focus, t1, t2));
    }
    void background(V3 c) { const double a[3] = {c.x, c.y, c.z}; chk(RTB_FN(scene_set_background)(s, a)); }
    void root(int32_t id) { chk(RTB_FN(scene_set_root)(s, id)); }
};

// vec3.rs:273-285 with the scene stream
inline V3 scene_random(SceneRng& r) { const double a = r.gen(), b = r.gen(), c = r.gen(); return V3(a, b, c); }
inline V3 scene_random_range(SceneRng& r, double mn, double mx) {
    const double a = r.gen_range(mn, mx), b = r.gen_range(mn, mx), c = r.gen_range(mn, mx);
    return V3(a, b, c);
}

// ---------------------------------------------------------------- scene builders (src/world.rs:95-874)

// gen_random_scene (world.rs:95-167), the crate as shipped ("C1b"): checker ground, thresholds
// 0.3/0.6, choose_mat < 0.8 => MovingSphere(c, c+(0,5,0), 0, 10), BVH over [0,10].
inline int32_t gen_random_scene(Builder& b, SceneRng& rng) {
    HittableList list;
    const int32_t ground = b.Lambertian_from_pointer(b.Checker_from_colors(V3(0.2, 0.3, 0.1), V3(0.9, 0.9, 0.9)));
    list.add(b.Sphere_new(V3(0, -1000, -1), 1000.0, ground));
    for (int a = -11; a < 11; ++a) {
        for (int bb = -11; bb < 11; ++bb) {
            const double choose_mat = rng.gen();
            const double cx = (double)a + 0.9 * rng.gen();
            const double cz = (double)bb + 0.9 * rng.gen();
            const V3 center(cx, 0.2, cz);
            if (length(center - V3(4, 0.2, 0)) > 0.9) {
                int32_t sphere_material;
                if (choose_mat < 0.3) {
                    const V3 r1 = scene_random(rng), r2 = scene_random(rng);
                    sphere_material = b.Lambertian_new(r1 * r2);
                } else if (choose_mat < 0.6) {
                    const V3 albedo = scene_random_range(rng, 0.5, 1.0);
                    const double fuzz = rng.gen_range(0.0, 0.5);
                    sphere_material = b.Metal_new(albedo, fuzz);
                } else {
                    sphere_material = b.Dielectric_new(1.5);
                }
                if (choose_mat < 0.8) {
                    list.add(b.MovingSphere_new(center, center + V3(0, 5, 0), 0.0, 10.0, 0.2, sphere_material));
                    continue;
                }
                list.add(b.Sphere_new(center, 0.2, sphere_material));
            }
        }
    }
    list.add(b.Sphere_new(V3(0, 1, 0), 1.0, b.Dielectric_new(1.5)));
    list.add(b.Sphere_new(V3(-4, 1, 0), 1.0, b.Lambertian_new(V3(0.4, 0.2, 0.1))));
    list.add(b.Sphere_new(V3(4, 1, 0), 1.0, b.Metal_new(V3(0.7, 0.6, 0.5), 0.0)));
    return b.BvhNode_from_list(list, 0.0, 10.0);
}

// "C1a": the RTIOW book-1 final scene the README numbers were taken on (README.md:11-23,
// images/book1.png): grey ground at (0,-1000,0), static spheres, thresholds 0.8/0.95.
inline int32_t gen_book1_classic(Builder& b, SceneRng& rng) {
    HittableList list;
    list.add(b.Sphere_new(V3(0, -1000, 0), 1000.0, b.Lambertian_new(V3(0.5, 0.5, 0.5))));
    for (int a = -11; a < 11; ++a) {
        for (int bb = -11; bb < 11; ++bb) {
            const double choose_mat = rng.gen();
            const double cx = (double)a + 0.9 * rng.gen();
            const double cz = (double)bb + 0.9 * rng.gen();
            const V3 center(cx, 0.2, cz);
            if (length(center - V3(4, 0.2, 0)) > 0.9) {
                int32_t m;
                if (choose_mat < 0.8) {
                    const V3 r1 = scene_random(rng), r2 = scene_random(rng);
                    m = b.Lambertian_new(r1 * r2);
                } else if (choose_mat < 0.95) {
                    const V3 albedo = scene_random_range(rng, 0.5, 1.0);
                    const double fuzz = rng.gen_range(0.0, 0.5);
                    m = b.Metal_new(albedo, fuzz);
                } else {
                    m = b.Dielectric_new(1.5);
                }
                list.add(b.Sphere_new(center, 0.2, m));
            }
        }
    }
    list.add(b.Sphere_new(V3(0, 1, 0), 1.0, b.Dielectric_new(1.5)));
    list.add(b.Sphere_new(V3(-4, 1, 0), 1.0, b.Lambertian_new(V3(0.4, 0.2, 0.1))));
    list.add(b.Sphere_new(V3(4, 1, 0), 1.0, b.Metal_new(V3(0.7, 0.6, 0.5), 0.0)));
    return b.BvhNode_from_list(list, 0.0, 1.0);
}

// gen_random_scene_moving (world.rs:169-244): GravitySphere everywhere (choose_mat < 1.0 always)
inline int32_t gen_random_scene_moving(Builder& b, SceneRng& rng) {
    const double max_time = 100.0;
    HittableList list;
    list.add(b.Sphere_new(V3(0, -1000, -1), 1000.0, b.Lambertian_from_pointer(b.SolidColor_new(V3(0.8, 0.8, 0.8)))));
    for (int a = -11; a < 11; ++a) {
        for (int bb = -11; bb < 11; ++bb) {
            if (std::abs(a - 0) <= 1 && std::abs(bb - 0) <= 1) continue;
            if (std::abs(a - 4) <= 1 && std::abs(bb - 0) <= 1) continue;
            const double choose_mat = rng.gen();
            const double cx = (double)a + 0.9 * rng.gen();
            const double cy = 1.7 + rng.gen_range(0.0, 2.0);
            const double cz = (double)bb + 0.9 * rng.gen();
            const V3 center(cx, cy, cz);
            if (length(center - V3(4, 0.2, 0)) > 0.9) {
                int32_t m;
                if (choose_mat < 0.3) {
                    const V3 r1 = scene_random(rng), r2 = scene_random(rng);
                    m = b.Lambertian_new(r1 * r2);
                } else if (choose_mat < 0.6) {
                    const V3 albedo = scene_random_range(rng, 0.5, 1.0);
                    const double fuzz = rng.gen_range(0.0, 0.5);
                    m = b.Metal_new(albedo, fuzz);
                } else {
                    m = b.Dielectric_new(1.5);
                }
                list.add(b.GravitySphere_new(center, 0.0, 0.2, m));
            }
        }
    }
    list.add(b.Sphere_new(V3(0, 1, 0), 1.0, b.Dielectric_new(1.5)));
    list.add(b.Sphere_new(V3(-4, 1, 0), 1.0, b.Lambertian_new(V3(0.4, 0.2, 0.1))));
    list.add(b.Sphere_new(V3(4, 1, 0), 1.0, b.Metal_new(V3(0.7, 0.6, 0.5), 0.0)));
    return b.BvhNode_from_list(list, 0.0, max_time);
}

// gen_checkered_sphere (world.rs:246-265)
inline int32_t gen_checkered_sphere(Builder& b) {
    HittableList list;
    const int32_t ground = b.Lambertian_from_pointer(b.Checker_from_colors(V3(0.2, 0.3, 0.1), V3(0.9, 0.9, 0.9)));
    list.add(b.Sphere_new(V3(0, -10, 0), 10.0, ground));
    list.add(b.Sphere_new(V3(0, 10, 0), 10.0, ground));
    return b.List(list);
}

// gen_two_perlin (world.rs:267-285)
inline int32_t gen_two_perlin(Builder& b, uint64_t seed) {
    HittableList list;
    const int32_t ground = b.Lambertian_from_pointer(b.Noise_new(4.0, seed));
    list.add(b.Sphere_new(V3(0, -1000, 0), 1000.0, ground));
    list.add(b.Sphere_new(V3(0, 2, 0), 2.0, ground));
    return b.List(list);
}

// Procedural stand-in for the absent "earthshit.ppm" (world.rs:290,580; .gitignore:8): a
// deterministic 1024x512 lat/long map (sea, banded continents, ice caps) on the 0..255 scale.
inline void earth_standin_texels(int32_t& w, int32_t& h, std::vector<double>& rgb) {
    w = 1024; h = 512;
    rgb.resize((size_t)w * h * 3);
    for (int j = 0; j < h; ++j) {
        for (int i = 0; i < w; ++i) {
            const double lon = 2.0 * 3.14159265358979323846 * (double)i / w;
            const double lat = 3.14159265358979323846 * ((double)j / h - 0.5);
            const double f = std::sin(3.0 * lon + 1.3) * std::cos(2.0 * lat) + 0.6 * std::sin(7.0 * lon - 2.0 * lat) * std::sin(5.0 * lat + 0.7) +
                             0.3 * std::cos(13.0 * lon + 4.0 * lat);
            double r, g, bl;
            if (std::fabs(lat) > 1.25) { r = 235; g = 240; bl = 245; }
            else if (f > 0.35) { r = 60 + 80 * (f - 0.35); g = 120 + 60 * std::cos(lat); bl = 50; }
            else { r = 20; g = 60 + 30 * (f + 1.0); bl = 140 + 40 * (f + 1.0); }
            const size_t o = ((size_t)j * w + i) * 3;
            rgb[o] = std::floor(r); rgb[o + 1] = std::floor(g); rgb[o + 2] = std::floor(bl);
        }
    }
}
inline int32_t earth_texture(Builder& b) {
    int32_t w, h;
    std::vector<double> rgb;
    earth_standin_texels(w, h, rgb);
    return b.Image_from_texels(w, h, rgb);
}

// earth (world.rs:287-305)
inline int32_t earth(Builder& b) {
    HittableList list;
    const int32_t ground = b.Lambertian_from_pointer(earth_texture(b));
    list.add(b.Sphere_new(V3(0, -1000, 0), 1000.0, ground));
    list.add(b.Sphere_new(V3(0, 2, 0), 2.0, ground));
    return b.List(list);
}

// gen_simple_light (world.rs:307-342)
inline int32_t gen_simple_light(Builder& b, uint64_t seed) {
    HittableList list;
    const int32_t ground = b.Lambertian_from_pointer(b.Noise_new(4.0, seed));
    list.add(b.Sphere_new(V3(0, -1000, 0), 1000.0, ground));
    list.add(b.Sphere_new(V3(0, 2, 0), 2.0, ground));
    const int32_t difflight = b.DiffuseLight_new(V3(10, 10, 10));
    list.add(b.XyRect_new(3.0, 5.0, 1.0, 3.0, -2.0, difflight));
    list.add(b.Sphere_new(V3(0, 10, 0), 3.0, difflight));
    return b.List(list);
}

// the five walls + light shared by cornell_box / cornell_smoke / triangular_prism
inline void cornell_walls(Builder& b, HittableList& list, int32_t& white) {
    const int32_t red = b.Lambertian_new(V3(0.65, 0.05, 0.05));
    white = b.Lambertian_new(V3(0.73, 0.73, 0.73));
    const int32_t green = b.Lambertian_new(V3(0.12, 0.45, 0.15));
    const int32_t light = b.DiffuseLight_new(V3(15, 15, 15));
    list.add(b.YzRect_new(0.0, 555.0, 0.0, 555.0, 555.0, green));
    list.add(b.YzRect_new(0.0, 555.0, 0.0, 555.0, 0.0, red));
    list.add(b.XzRect_new(213.0, 343.0, 227.0, 332.0, 554.0, light));
    list.add(b.XzRect_new(0.0, 555.0, 0.0, 555.0, 0.0, white));
    list.add(b.XzRect_new(0.0, 555.0, 0.0, 555.0, 555.0, white));
    list.add(b.XyRect_new(0.0, 555.0, 0.0, 555.0, 555.0, white));
}

// cornell_box (world.rs:344-413)
inline int32_t cornell_box(Builder& b) {
    HittableList list;
    int32_t white;
    cornell_walls(b, list, white);
    list.add(b.Translate_new(V3(265, 0, 295), b.RotateY_new(15.0, b.RectPrism_new(V3(0, 0, 0), V3(165, 330, 165), white))));
    list.add(b.Translate_new(V3(130, 0, 65), b.RotateY_new(-18.0, b.RectPrism_new(V3(0, 0, 0), V3(165, 165, 165), white))));
    return b.List(list);
}

// cornell_smoke (world.rs:415-492)
inline int32_t cornell_smoke(Builder& b) {
    HittableList list;
    int32_t white;
    cornell_walls(b, list, white);
    list.add(b.ConstantMedium_from_color(
        V3(0, 0, 0), 0.01, b.Translate_new(V3(265, 0, 295), b.RotateY_new(15.0, b.RectPrism_new(V3(0, 0, 0), V3(165, 330, 165), white)))));
    list.add(b.ConstantMedium_from_color(
        V3(1, 1, 1), 0.01, b.Translate_new(V3(130, 0, 65), b.RotateY_new(-18.0, b.RectPrism_new(V3(0, 0, 0), V3(165, 165, 165), white)))));
    return b.List(list);
}

// final_scene (world.rs:494-616).  Differences from the book that the reference has and this keeps:
// flat top-level list, light x1 = 432, the r=5000 boundary sphere is also a visible dielectric,
// moving-sphere albedo (0.7,0.3,1), box heights gen_range(1..101).
inline int32_t final_scene(Builder& b, SceneRng& rng, uint64_t seed) {
    HittableList list, boxes1;
    const int32_t ground = b.Lambertian_from_pointer(b.SolidColor_new(V3(0.48, 0.83, 0.53)));
    const int boxes_per_side = 20;
    for (int i = 0; i < boxes_per_side; ++i) {
        for (int j = 0; j < boxes_per_side; ++j) {
            const double w = 100.0;
            const double x0 = -1000.0 + (double)i * w;
            const double z0 = -1000.0 + (double)j * w;
            const double y0 = 0.0;
            const double x1 = x0 + w;
            const double y1 = rng.gen_range(1.0, 101.0);
            const double z1 = z0 + w;
            boxes1.add(b.RectPrism_new(V3(x0, y0, z0), V3(x1, y1, z1), ground));
        }
    }
    list.add(b.BvhNode_from_list(boxes1, 0.0, 1.0));
    list.add(b.XzRect_new(123.0, 432.0, 147.0, 412.0, 554.0, b.DiffuseLight_new(V3(7, 7, 7))));
    const V3 center1(400, 400, 400);
    const V3 center2 = center1 + V3(30, 0, 0);
    list.add(b.MovingSphere_new(center1, center2, 0.0, 1.0, 50.0, b.Lambertian_new(V3(0.7, 0.3, 1))));
    list.add(b.Sphere_new(V3(260, 150, 45), 50.0, b.Dielectric_new(1.5)));
    list.add(b.Sphere_new(V3(0, 150, 145), 50.0, b.Metal_new(V3(0.8, 0.8, 0.9), 1.0)));
    list.add(b.Sphere_new(V3(360, 150, 145), 70.0, b.Dielectric_new(1.5)));
    list.add(b.ConstantMedium_from_color(V3(0.2, 0.4, 0.9), 0.2, b.Sphere_new(V3(360, 150, 145), 70.0, b.Dielectric_new(1.5))));
    list.add(b.Sphere_new(V3(0, 0, 0), 5000.0, b.Dielectric_new(1.5)));
    list.add(b.ConstantMedium_from_color(V3(1, 1, 1), 0.0001, b.Sphere_new(V3(0, 0, 0), 5000.0, b.Dielectric_new(1.5))));
    list.add(b.Sphere_new(V3(400, 200, 400), 100.0, b.Lambertian_from_pointer(earth_texture(b))));
    list.add(b.Sphere_new(V3(220, 280, 300), 80.0, b.Lambertian_from_pointer(b.Noise_new(0.1, seed))));
    const int32_t white = b.Lambertian_new(V3(0.73, 0.73, 0.73));
    HittableList boxes2;
    const int ns = 1000;
    for (int k = 0; k < ns; ++k) boxes2.add(b.Sphere_new(scene_random_range(rng, 0.0, 165.0), 10.0, white));
    list.add(b.Translate_new(V3(-100, 270, 395), b.RotateY_new(15.0, b.BvhNode_from_list(boxes2, 0.0, 1.0))));
    return b.List(list);
}

// gen_moving_test (world.rs:618-647)
inline int32_t gen_moving_test(Builder& b) {
    HittableList list;
    const int32_t ground = b.Lambertian_from_pointer(b.Checker_from_colors(V3(0.2, 0.3, 0.1), V3(0.9, 0.9, 0.9)));
    list.add(b.Sphere_new(V3(0, -1000, -1), 1000.0, ground));
    list.add(b.MovingSphere_new(V3(2, -1, 2), V3(2, 7, 2), 0.0, 10.0, 1.0, b.Lambertian_new(V3(1, 0, 0))));
    return b.BvhNode_from_list(list, 0.0, 10.0);
}

// benchmark_test_scene (world.rs:649-663): one sphere nested in 20 HittableLists
inline int32_t benchmark_test_scene(Builder& b) {
    HittableList amit;
    amit.add(b.Sphere_new(V3(0, 0, 0), 4.0, b.Lambertian_new(V3(0.5, 0.5, 0.5))));
    for (int i = 0; i < 19; ++i) {
        HittableList tramit;
        tramit.add(b.List(amit));
        amit = tramit;
    }
    return b.List(amit);
}

// triangle_test (world.rs:665-679)
inline int32_t triangle_test(Builder& b) {
    HittableList list;
    list.add(b.Triangle_new(V3(0, 5, 0), V3(5, 0, 0), V3(0, 0, 0), b.Lambertian_new(V3(1, 0, 0))));
    list.add(b.Sphere_new(V3(5, 0, 0), 1.0, b.Lambertian_new(V3(0, 1, 0))));
    return b.List(list);
}

// Synthetic stand-in for models/dragon_recon/dragon_vrip_res2.ply (absent, .gitignore:8): an
// n x n-quad displaced parametric "blob" (2 n^2 triangles, (n+1)^2 vertices; n = 660 gives
// 871 200 triangles, the size of dragon_vrip.ply), in the dragon's raw units (extent ~0.2, the
// scene scales by 100), written as the ASCII-PLY subset that model.rs:13-62 parses.
inline bool write_synthetic_mesh_ply(const char* path, int n, uint64_t seed) {
    FILE* f = std::fopen(path, "w");
    if (!f) return false;
    SceneRng rng(seed);
    const int nv = (n + 1) * (n + 1), nf = 2 * n * n;
    std::fprintf(f, "ply\nformat ascii 1.0\nelement vertex %d\nproperty float x\nproperty float y\nproperty float z\n", nv);
    std::fprintf(f, "element face %d\nproperty list uchar int vertex_indices\nend_header\n", nf);
    const double ph1 = rng.gen_range(0.0, 6.28), ph2 = rng.gen_range(0.0, 6.28), ph3 = rng.gen_range(0.0, 6.28);
    for (int j = 0; j <= n; ++j) {
        // heavy-tailed parameter spacing => heavy-tailed triangle sizes
        const double tv = (double)j / n;
        const double v = 0.5 - 0.5 * std::cos(3.14159265358979323846 * tv); // clusters rows near the poles
        for (int i = 0; i <= n; ++i) {
            const double tu = (double)i / n;
            const double u = tu + 0.15 * std::sin(2.0 * 3.14159265358979323846 * tu) / (2.0 * 3.14159265358979323846);
            const double th = 2.0 * 3.14159265358979323846 * u, ph = 3.14159265358979323846 * (0.02 + 0.96 * v);
            const double bump = 1.0 + 0.25 * std::sin(5.0 * th + ph1) * std::sin(4.0 * ph + ph2) + 0.08 * std::sin(23.0 * th + 17.0 * ph + ph3) +
                                0.03 * std::sin(61.0 * th) * std::sin(47.0 * ph);
            const double r = 0.072 * bump;
            // dragon extent x100: x in [-11,10], y in [5.3,19.8], z in [-4.5,4.7]
            const double x = -0.005 + 1.45 * r * std::sin(ph) * std::cos(th);
            const double y = 0.1255 + 1.0 * r * std::cos(ph);
            const double z = 0.001 + 0.63 * r * std::sin(ph) * std::sin(th);
            std::fprintf(f, "%.9g %.9g %.9g\n", (double)(float)x, (double)(float)y, (double)(float)z);
        }
    }
    for (int j = 0; j < n; ++j) {
        for (int i = 0; i < n; ++i) {
            const int a = j * (n + 1) + i, b2 = a + 1, c = a + (n + 1), d = c + 1;
            std::fprintf(f, "3 %d %d %d\n3 %d %d %d\n", a, b2, d, a, d, c);
        }
    }
    std::fclose(f);
    return true;
}

// stanford_dragon (world.rs:681-751) with the synthetic mesh; `quads` = n above.
inline int32_t stanford_dragon(Builder& b, uint64_t seed, int quads, const char* ply_path_or_null) {
    HittableList list;
    std::string path;
    if (ply_path_or_null) {
        path = ply_path_or_null;
    } else {
        char buf[256];
        std::snprintf(buf, sizeof buf, "/tmp/rtb200_mesh_%llx_%d.ply", (unsigned long long)seed, quads);
        path = buf;
        FILE* probe = std::fopen(path.c_str(), "r");
        if (probe) std::fclose(probe);
        else if (!write_synthetic_mesh_ply(path.c_str(), quads, seed)) { b.err = RT_ERR_IO; return RT_ERR_IO; }
    }
    // model.rs:68-73 allocates one Lambertian(0.2,0.2,0.2) per triangle; one shared material here
    const int32_t dragon_list = b.TriangleModel_load(path.c_str(), 100.0, b.Lambertian_new(V3(0.2, 0.2, 0.2)));
    if (dragon_list < 0) return dragon_list;
    HittableList dl;
    dl.add(dragon_list);
    const int32_t dragon = b.BvhNode_from_list(dl, 0.0, 1.0);
    const int32_t light = b.DiffuseLight_new(V3(4, 4, 4));
    const int32_t backdrop = b.XyRect_new(-100.0, 100.0, -100.0, 100.0, -20.0, b.Lambertian_new(V3(0.8, 0.3, 0.3)));
    const int32_t backwall = b.XyRect_new(-100.0, 100.0, -100.0, 100.0, 20.0, b.Lambertian_new(V3(1, 1, 1)));
    const int32_t ground = b.XzRect_new(-40.0, 40.0, -40.0, 40.0, 5.0, b.Metal_new(V3(0.3, 0.3, 0.3), 0.02));
    const int32_t ceiling = b.XzRect_new(-100.0, 100.0, -100.0, 100.0, 55.0, b.Metal_new(V3(1, 1, 1), 0.0));
    const int32_t left_wall = b.YzRect_new(-100.0, 100.0, -100.0, 100.0, -30.0, b.Lambertian_new(V3(0.3, 0.8, 0.3)));
    const int32_t right_wall = b.YzRect_new(-100.0, 100.0, -100.0, 100.0, 30.0, b.Lambertian_new(V3(0.3, 0.3, 0.8)));
    const int32_t ceiling_light = b.XzRect_new(-100.0, 100.0, -100.0, 100.0, 55.0, light);
    list.add(dragon);
    list.add(backdrop);
    list.add(backwall);
    list.add(ground);
    list.add(ceiling);
    list.add(left_wall);
    list.add(right_wall);
    list.add(ceiling_light);
    return b.List(list);
}

// triangular_prism (world.rs:753-874): Cornell walls + one triangle + one rect
inline int32_t triangular_prism(Builder& b) {
    HittableList list;
    int32_t white;
    cornell_walls(b, list, white);
    list.add(b.Triangle_new(V3(200, 0, 200), V3(300, 0, 200), V3(250, 250, 200), white));
    list.add(b.XyRect_new(0.0, 300.0, 0.0, 150.0, 201.0, white));
    return b.List(list);
}

// get_world_cam (world.rs:876-1179): scene + camera preset + background.  Ids as the reference;
// 13 = C1a classic book-1 (3:2 camera), 14 = same room as 11 (alias), others = gen_random_scene.
// A BvhNode::from_list over a single TriangleModel list (world.rs:687) is expressed as
// bvh([list]) — semantically the same closest-hit set.
inline int32_t build_world(rt_scene* s, int32_t scene_id, uint64_t seed, int32_t param) {
    Builder b(s);
    SceneRng rng(seed);
    const double aspect_ratio = 16.0 / 9.0;       // world.rs:878
    const V3 background(0.7, 0.8, 1.0);           // world.rs:879
    const V3 vup(0, 1, 0);
    int32_t root = -1;
    switch (scene_id) {
    case 0: root = gen_checkered_sphere(b); b.Camera_new(V3(13, 2, 3), V3(0, 0, 0), vup, 20.0, aspect_ratio, 0.0, 10.0, 0.0, 1.0); b.background(background); break;
    case 1: root = gen_two_perlin(b, seed); b.Camera_new(V3(13, 2, 3), V3(0, 0, 0), vup, 20.0, aspect_ratio, 0.0, 10.0, 0.0, 1.0); b.background(background); break;
    case 2: root = earth(b); b.Camera_new(V3(13, 2, 3), V3(0, 0, 0), vup, 20.0, aspect_ratio, 0.0, 10.0, 0.0, 1.0); b.background(background); break;
    case 3: root = gen_simple_light(b, seed); b.Camera_new(V3(26, 3, 6), V3(0, 2, 0), vup, 20.0, aspect_ratio, 0.0, 10.0, 0.0, 1.0); b.background(V3(0, 0, 0)); break;
    case 4: root = cornell_box(b); b.Camera_new(V3(278, 278, -800), V3(278, 278, 0), vup, 40.0, 1.0, 0.0, 10.0, 0.0, 1.0); b.background(V3(0, 0, 0)); break;
    case 5: root = cornell_smoke(b); b.Camera_new(V3(278, 278, -800), V3(278, 278, 0), vup, 40.0, 1.0, 0.0, 10.0, 0.0, 1.0); b.background(V3(0, 0, 0)); break;
    case 6: root = final_scene(b, rng, seed); b.Camera_new(V3(478, 278, -600), V3(278, 278, 0), vup, 40.0, 1.0, 0.0, 10.0, 0.0, 1.0); b.background(V3(0, 0, 0)); break;
    case 7: root = gen_moving_test(b); b.Camera_new(V3(13, 2, 3), V3(0, 0, 0), vup, 20.0, aspect_ratio, 0.1, 10.0, 2.0, 2.5); b.background(background); break;
    case 8: root = gen_random_scene_moving(b, rng); b.Camera_new(V3(13, 2, 3), V3(0, 0, 0), vup, 20.0, aspect_ratio, 0.1, 10.0, 0.0, 10.0); b.background(background); break;
    case 9: root = benchmark_test_scene(b); b.Camera_new(V3(13, 2, 3), V3(0, 0, 0), vup, 20.0, aspect_ratio, 0.1, 10.0, 0.0, 10.0); b.background(background); break;
    case 10: root = triangle_test(b); b.Camera_new(V3(0, 0, 20), V3(0, 0, 0), vup, 20.0, aspect_ratio, 0.1, 10.0, 0.0, 10.0); b.background(background); break;
    case 11:
    case 14:
        root = stanford_dragon(b, seed, param > 0 ? param : 660, nullptr);
        // world.rs:1114-1134 uses the 16/9 camera; id 14 uses aspect 1.0 to match a square image
        b.Camera_new(V3(0, 20, 20), V3(0, 11, 0), vup, 60.0, scene_id == 11 ? aspect_ratio : 1.0, 0.0, 40.0, 0.0, 10.0);
        b.background(background);
        break;
    case 12: root = triangular_prism(b); b.Camera_new(V3(278, 278, -800), V3(278, 278, 0), vup, 40.0, 1.0, 0.0, 10.0, 0.0, 1.0); b.background(V3(0, 0, 0)); break;
    case 13: root = gen_book1_classic(b, rng); b.Camera_new(V3(13, 2, 3), V3(0, 0, 0), vup, 20.0, 3.0 / 2.0, 0.1, 10.0, 0.0, 1.0); b.background(background); break;
    default: root = gen_random_scene(b, rng); b.Camera_new(V3(13, 2, 3), V3(0, 0, 0), vup, 20.0, aspect_ratio, 0.1, 10.0, 0.0, 10.0); b.background(background); break;
    }
    if (root < 0) return root;
    b.root(root);
    return b.err;
}

} // namespace rtb
