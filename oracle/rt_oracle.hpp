// rt_oracle.hpp — CPU f64 restatement of the reference's path-tracing hot path.
//
// TEST INFRASTRUCTURE ONLY.  Nothing in the product (ray_tracing_series_rust_b200/, include/)
// may include, link or call this.  Only tests/, __graft_entry__.smoke() and bench.py's
// cpu_baseline / --impl reference legs use it, as the checker and the CPU baseline.
//
// Parity pin: the reference (patrickzbhe/ray-tracing-series-rust) is pure Rust and cannot be
// built here (no cargo/rustc).  What the reference itself holds for this path, and what pins
// this restatement to it: (1) its only unit tests, src/vec3.rs:343-428 (replicated exactly in
// tests/test_oracle_kat.py); (2) the three renders it ships, images/book1.png, book2.png,
// stanford_dragon.png: unrendered black rows, saturated light, the analytic book-1 sky and the
// mean radiance of every region the scene code fixes (tests/test_reference_images.py against
// tests/golden/reference_images.json, made by tools/make_reference_image_fixture.py).  The
// renders pin the path statistically (the reference's RNG is unseeded); the exact arithmetic
// of single functions beyond vec3 has no reference-owned vector and is pinned by the
// hand-derived known-answer vectors of SURVEY.md Appendix B and by structural invariants.
//
// Every function cites the reference file:line it follows.  Data structures deliberately keep
// the reference's shape (pointer-based object graph, virtual dispatch, per-node reciprocal
// divides, recursive median-split BVH over {x,y} axes) because this is also the CPU baseline.
//
// Randomness: the reference draws from an unseeded rand::thread_rng() (ChaCha12).  Only the
// distributions matter; the oracle draws from Philox-4x32-10 keyed (seed, path id) with one
// counter step per draw so the GPU path can be compared sample by sample.
#pragma once
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <limits>
#include <memory>
#include <string>
#include <vector>

namespace orc {

constexpr double INF = std::numeric_limits<double>::infinity();
constexpr double PI = 3.14159265358979323846264338327950288; // std::f64::consts::PI

// ------------------------------------------------------------------ Philox-4x32-10
struct Philox {
    static inline void round(uint32_t c[4], const uint32_t k[2]) {
        const uint64_t p0 = (uint64_t)0xD2511F53u * c[0];
        const uint64_t p1 = (uint64_t)0xCD9E8D57u * c[2];
        const uint32_t hi0 = (uint32_t)(p0 >> 32), lo0 = (uint32_t)p0;
        const uint32_t hi1 = (uint32_t)(p1 >> 32), lo1 = (uint32_t)p1;
        const uint32_t n0 = hi1 ^ c[1] ^ k[0];
        const uint32_t n2 = hi0 ^ c[3] ^ k[1];
        c[0] = n0; c[1] = lo1; c[2] = n2; c[3] = lo0;
    }
    static inline void block(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]) {
        uint32_t c[4] = {ctr[0], ctr[1], ctr[2], ctr[3]};
        uint32_t k[2] = {key[0], key[1]};
        for (int r = 0; r < 10; ++r) {
            if (r) { k[0] += 0x9E3779B9u; k[1] += 0xBB67AE85u; }
            round(c, k);
        }
        out[0] = c[0]; out[1] = c[1]; out[2] = c[2]; out[3] = c[3];
    }
};

// ------------------------------------------------------------------ event counters
struct Counters {
    uint64_t paths = 0, segments = 0, box_tests = 0, medium_queries = 0;
    uint64_t prim_tests[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    uint64_t scatters[5] = {0, 0, 0, 0, 0};
    void add(const Counters& o) {
        paths += o.paths; segments += o.segments; box_tests += o.box_tests;
        medium_queries += o.medium_queries;
        for (int i = 0; i < 8; ++i) prim_tests[i] += o.prim_tests[i];
        for (int i = 0; i < 5; ++i) scatters[i] += o.scatters[i];
    }
};
enum PrimType { PT_SPHERE = 0, PT_MOVING = 1, PT_GRAVITY = 2, PT_RECT = 3, PT_BOX = 4, PT_TRI = 5, PT_MEDIUM = 6 };
enum MatType { MT_LAMBERTIAN = 0, MT_METAL = 1, MT_DIELECTRIC = 2, MT_LIGHT = 3, MT_ISOTROPIC = 4 };

// ------------------------------------------------------------------ per-path RNG context
// Stands in for every `thread_rng()` call site of the reference.  Draw k of a path is word k&3
// of Philox block (path_lo, path_hi, k>>2, 0) under key (seed_lo, seed_hi); xi = u32 * 2^-32.
// Each Material::scatter starts on a block boundary (begin_event).
// ConstantMedium draws come from the order-independent sub-stream
// (path_lo, path_hi, medium prim id, 0x80000000 | segment).
struct PathCtx {
    uint32_t key[2] = {0, 0};
    uint32_t path[2] = {0, 0};
    uint32_t draw = 0;
    uint32_t cached_block = 0xffffffffu;
    uint32_t blk[4] = {0, 0, 0, 0};
    uint32_t segment = 0;
    bool media_enabled = true;
    Counters cnt;

    void begin_path(uint64_t seed, uint64_t path_id) {
        key[0] = (uint32_t)seed; key[1] = (uint32_t)(seed >> 32);
        path[0] = (uint32_t)path_id; path[1] = (uint32_t)(path_id >> 32);
        draw = 0; cached_block = 0xffffffffu; segment = 0;
    }
    uint32_t next_u32() {
        const uint32_t b = draw >> 2;
        if (b != cached_block) {
            const uint32_t ctr[4] = {path[0], path[1], b, 0u};
            Philox::block(ctr, key, blk);
            cached_block = b;
        }
        return blk[draw++ & 3u];
    }
    // Every Material::scatter call starts at the next multiple-of-4 draw index, i.e. on a fresh Philox
    // block (part of the RNG contract shared with the GPU path: the block a scatter needs first is known
    // before its rejection loop starts; at most 3 words per bounce are skipped).
    void begin_event() { draw = (draw + 3u) & ~3u; }
    // rng.gen::<f64>()  -> uniform [0,1)
    double gen() { return (double)next_u32() * (1.0 / 4294967296.0); }
    // rng.gen_range(a..b) on f64
    double gen_range(double a, double b) { return a + gen() * (b - a); }
    double medium_xi(uint32_t medium_prim_id) const {
        const uint32_t ctr[4] = {path[0], path[1], medium_prim_id, 0x80000000u | segment};
        uint32_t out[4];
        Philox::block(ctr, key, out);
        return (double)out[0] * (1.0 / 4294967296.0);
    }
};
inline PathCtx*& tls_ctx() {
    static thread_local PathCtx* p = nullptr;
    return p;
}
inline PathCtx& ctx() { return *tls_ctx(); }

// ------------------------------------------------------------------ vec3.rs
struct Vec3 {
    double x = 0, y = 0, z = 0;
    Vec3() {}
    Vec3(double a, double b, double c) : x(a), y(b), z(c) {}
    double length_squared() const { return dot(*this); }                 // vec3.rs:35-37
    double length() const { return std::sqrt(length_squared()); }        // vec3.rs:39-41
    double dot(const Vec3& o) const { return x * o.x + y * o.y + z * o.z; } // vec3.rs:43-45
    Vec3 cross(const Vec3& o) const {                                    // vec3.rs:47-53
        return Vec3(y * o.z - z * o.y, z * o.x - x * o.z, x * o.y - y * o.x);
    }
    Vec3 unit() const;                                                   // vec3.rs:55-57
    bool near_zero() const {                                             // vec3.rs:59-62
        const double s = 1e-8;
        return std::fabs(x) < s && std::fabs(y) < s && std::fabs(z) < s;
    }
    Vec3 reflect(const Vec3& n) const;                                   // vec3.rs:64-66
    double axis(int a) const { return a == 0 ? x : (a == 1 ? y : z); }
};
inline Vec3 operator*(const Vec3& a, const Vec3& b) { return Vec3(a.x * b.x, a.y * b.y, a.z * b.z); } // vec3.rs:139-149
inline Vec3 operator-(const Vec3& a) { return Vec3(a.x * -1.0, a.y * -1.0, a.z * -1.0); }             // vec3.rs:151-161
inline Vec3 operator*(const Vec3& a, double s) { return Vec3(a.x * s, a.y * s, a.z * s); }             // vec3.rs:163-173
inline Vec3 operator*(double s, const Vec3& a) { return Vec3(s * a.x, s * a.y, s * a.z); }             // vec3.rs:175-185
inline Vec3 operator/(const Vec3& a, double s) { return Vec3(a.x / s, a.y / s, a.z / s); }             // vec3.rs:187-197
inline Vec3 operator+(const Vec3& a, const Vec3& b) { return Vec3(a.x + b.x, a.y + b.y, a.z + b.z); } // vec3.rs:199-209
inline Vec3 operator-(const Vec3& a, const Vec3& b) { return Vec3(a.x - b.x, a.y - b.y, a.z - b.z); } // vec3.rs:211-221
inline Vec3& operator+=(Vec3& a, const Vec3& b) { a.x += b.x; a.y += b.y; a.z += b.z; return a; }      // vec3.rs:223-229
inline Vec3& operator*=(Vec3& a, double s) { a.x *= s; a.y *= s; a.z *= s; return a; }                 // vec3.rs:231-237
inline Vec3& operator/=(Vec3& a, double s) { a *= (1.0 / s); return a; }                               // vec3.rs:239-243
inline Vec3& operator*=(Vec3& a, const Vec3& b) { a.x *= b.x; a.y *= b.y; a.z *= b.z; return a; }      // vec3.rs:245-251
inline bool operator==(const Vec3& a, const Vec3& b) { return a.x == b.x && a.y == b.y && a.z == b.z; } // vec3.rs:253-259
inline Vec3 Vec3::unit() const { return *this / length(); }
inline Vec3 Vec3::reflect(const Vec3& n) const { return *this - 2.0 * this->dot(n) * n; }
typedef Vec3 Point3;
typedef Vec3 Color;

inline double clamp(double x, double mn, double mx) { // mutil.rs:1-9
    if (x < mn) return mn;
    if (x > mx) return mx;
    return x;
}

// Rust `f64 as i32`: truncation toward zero, saturating, NaN -> 0
inline int32_t f64_as_i32(double v) {
    if (v != v) return 0;
    if (v >= 2147483647.0) return 2147483647;
    if (v <= -2147483648.0) return (-2147483647 - 1);
    return (int32_t)v;
}

// vec3.rs:89-107 get_normalized_color
inline Color get_normalized_color(const Color& sum, uint32_t samples_per_pixel) {
    const double COLOR_MAX = 255.9; // vec3.rs:10
    double r = sum.x, g = sum.y, b = sum.z;
    const double scale = 1.0 / (double)samples_per_pixel;
    r *= scale; g *= scale; b *= scale;
    r = std::sqrt(r); g = std::sqrt(g); b = std::sqrt(b);
    return Color((double)f64_as_i32(COLOR_MAX * clamp(r, 0.0, 1.0)),
                 (double)f64_as_i32(COLOR_MAX * clamp(g, 0.0, 1.0)),
                 (double)f64_as_i32(COLOR_MAX * clamp(b, 0.0, 1.0)));
}

// vec3.rs:116-121 refract
inline Vec3 refract(const Vec3& uv, const Vec3& n, double etai_over_etat) {
    const double cos_theta = std::fmin((-uv).dot(n), 1.0);
    const Vec3 r_out_perp = etai_over_etat * (uv + cos_theta * n);
    const Vec3 r_out_parallel = -(std::sqrt(std::fabs(1.0 - r_out_perp.length_squared()))) * n;
    return r_out_perp + r_out_parallel;
}

// vec3.rs:273-322 samplers
inline Vec3 random_vec() { // vec3.rs:273-276
    const double a = ctx().gen(), b = ctx().gen(), c = ctx().gen();
    return Vec3(a, b, c);
}
inline Vec3 random_range(double mn, double mx) { // vec3.rs:278-285
    const double a = ctx().gen_range(mn, mx), b = ctx().gen_range(mn, mx), c = ctx().gen_range(mn, mx);
    return Vec3(a, b, c);
}
inline Vec3 random_in_unit_sphere() { // vec3.rs:287-295
    for (;;) {
        const Vec3 p = random_range(-1.0, 1.0);
        if (p.length_squared() < 1.0) return p;
    }
}
inline Vec3 random_unit_vector() { return random_in_unit_sphere().unit(); } // vec3.rs:297-299
inline Vec3 random_in_unit_disk() { // vec3.rs:310-322
    for (;;) {
        const double a = ctx().gen_range(-1.0, 1.0), b = ctx().gen_range(-1.0, 1.0);
        const Vec3 p(a, b, 0.0);
        if (p.length_squared() < 1.0) return p;
    }
}

// ------------------------------------------------------------------ ray.rs
struct Ray {
    Point3 origin;
    Vec3 direction;
    double time = 0;
    Ray() {}
    Ray(const Point3& o, const Vec3& d, double t) : origin(o), direction(d), time(t) {}
    Point3 at(double t) const { return origin + direction * t; } // ray.rs:31-33
};

// ------------------------------------------------------------------ aabb.rs
struct Aabb {
    Point3 minimum, maximum;
    Aabb() {}
    Aabb(const Point3& a, const Point3& b) : minimum(a), maximum(b) {}
    // aabb.rs:23-61
    bool hit(const Ray& r, double t_min, double t_max) const {
        ctx().cnt.box_tests++;
        for (int a = 0; a < 3; ++a) {
            const double inv_d = 1.0 / r.direction.axis(a);
            double t0 = (minimum.axis(a) - r.origin.axis(a)) * inv_d;
            double t1 = (maximum.axis(a) - r.origin.axis(a)) * inv_d;
            if (inv_d < 0.0) std::swap(t0, t1);
            t_min = t0 > t_min ? t0 : t_min;
            t_max = t1 < t_max ? t1 : t_max;
            if (t_max <= t_min) return false;
        }
        return true;
    }
    // aabb.rs:63-77
    static Aabb surrounding_box(const Aabb& b0, const Aabb& b1) {
        const Point3 small(std::fmin(b0.minimum.x, b1.minimum.x), std::fmin(b0.minimum.y, b1.minimum.y),
                           std::fmin(b0.minimum.z, b1.minimum.z));
        const Point3 big(std::fmax(b0.maximum.x, b1.maximum.x), std::fmax(b0.maximum.y, b1.maximum.y),
                         std::fmax(b0.maximum.z, b1.maximum.z));
        return Aabb(small, big);
    }
};

// ------------------------------------------------------------------ perlin.rs
struct Perlin {
    std::vector<Vec3> ranvec;
    std::vector<int32_t> perm_x, perm_y, perm_z;
    // perlin.rs:28-52
    double noise(const Point3& p) const {
        const double u = p.x - std::floor(p.x);
        const double v = p.y - std::floor(p.y);
        const double w = p.z - std::floor(p.z);
        const int32_t i = f64_as_i32(std::floor(p.x));
        const int32_t j = f64_as_i32(std::floor(p.y));
        const int32_t k = f64_as_i32(std::floor(p.z));
        Vec3 c[2][2][2];
        for (int di = 0; di < 2; ++di)
            for (int dj = 0; dj < 2; ++dj)
                for (int dk = 0; dk < 2; ++dk)
                    c[di][dj][dk] = ranvec[(size_t)(perm_x[(size_t)((i + di) & 255)] ^
                                                    perm_y[(size_t)((j + dj) & 255)] ^
                                                    perm_z[(size_t)((k + dk) & 255)])];
        return trilinear_interp(c, u, v, w);
    }
    // perlin.rs:54-66
    double turbulence(const Point3& p, size_t depth) const {
        double accum = 0.0;
        Point3 temp_p = p;
        double weight = 1.0;
        for (size_t o = 0; o < depth; ++o) {
            accum += weight * noise(temp_p);
            weight *= 0.5;
            temp_p = temp_p * 2.0;
        }
        return std::fabs(accum);
    }
    // perlin.rs:85-106
    static double trilinear_interp(const Vec3 c[2][2][2], double u, double v, double w) {
        const double uu = u * u * (3.0 - 2.0 * u);
        const double vv = v * v * (3.0 - 2.0 * v);
        const double ww = w * w * (3.0 - 2.0 * w);
        double accum = 0.0;
        for (int i = 0; i < 2; ++i)
            for (int j = 0; j < 2; ++j)
                for (int k = 0; k < 2; ++k) {
                    const double i1 = i, j1 = j, k1 = k;
                    const Vec3 weight_v(u - i1, v - j1, w - k1);
                    accum += (i1 * uu + (1.0 - i1) * (1.0 - uu)) * (j1 * vv + (1.0 - j1) * (1.0 - vv)) *
                             (k1 * ww + (1.0 - k1) * (1.0 - ww)) * c[i][j][k].dot(weight_v);
                }
        return accum;
    }
    // perlin.rs:14-26, 68-83 with a seeded generator (the reference's is unseeded): gradients
    // U[-1,1)^3, identity permutation shuffled for i = 254..1 with target U{0..i}.
    static Perlin generate(uint64_t seed);
};

// ------------------------------------------------------------------ texture.rs
struct ImageData { // screen.rs Screen as a texel store: pixels[row*width+col], row 0 = first file row
    size_t width = 0, height = 0;
    std::vector<Color> pixels;
};
struct Texture {
    virtual ~Texture() {}
    virtual Color value(double u, double v, const Point3& p) const = 0; // texture.rs:7-9
};
struct SolidColor : Texture {
    Color color_value;
    explicit SolidColor(const Color& c) : color_value(c) {}
    Color value(double, double, const Point3&) const override { return color_value; } // texture.rs:27-31
};
struct Checker : Texture {
    std::shared_ptr<Texture> even, odd;
    Checker(std::shared_ptr<Texture> e, std::shared_ptr<Texture> o) : even(e), odd(o) {}
    Color value(double u, double v, const Point3& p) const override { // texture.rs:54-64
        const double sines = std::sin(10.0 * p.x) * std::sin(10.0 * p.y) * std::sin(10.0 * p.z);
        if (sines < 0.0) return odd->value(u, v, p);
        return even->value(u, v, p);
    }
};
struct Noise : Texture {
    Perlin noise;
    double scale;
    Noise(const Perlin& p, double s) : noise(p), scale(s) {}
    Color value(double, double, const Point3& p) const override { // texture.rs:80-88
        return Color(1, 1, 1) * 0.5 * (1.0 + std::sin(scale * p.z + 10.0 * noise.turbulence(p, 7)));
    }
};
struct Image : Texture {
    ImageData data;
    explicit Image(ImageData d) : data(std::move(d)) {}
    Color value(double u, double v, const Point3&) const override { // texture.rs:102-121
        u = clamp(u, 0.0, 1.0);
        v = 1.0 - clamp(v, 0.0, 1.0);
        int32_t i = f64_as_i32(u * (double)data.width);
        int32_t j = f64_as_i32(v * (double)data.height);
        i = std::min(i, (int32_t)data.width - 1);
        j = std::min(j, (int32_t)data.height - 1);
        const double color_scale = 1.0 / 255.0;
        const Color& pixel = data.pixels[(size_t)j * data.width + (size_t)i];
        return Color(color_scale * pixel.x, color_scale * pixel.y, color_scale * pixel.z);
    }
};

// ------------------------------------------------------------------ hit.rs: HitRecord, traits
struct Material;
struct HitRecord { // hit.rs:9-18 (+ prim_id: which leaf produced the record, for the parity hook)
    Point3 p;
    Vec3 normal;
    double t = 0, u = 0, v = 0;
    bool front_face = false;
    const Material* mat_ptr = nullptr;
    int32_t prim_id = -1;
    // hit.rs:69-79
    static void create_normal_face(const Ray& r, const Vec3& outward_normal, Vec3& normal, bool& front_face) {
        front_face = r.direction.dot(outward_normal) < 0.0;
        normal = front_face ? outward_normal : -outward_normal;
    }
};

struct Material {
    int32_t mat_id = -1;
    virtual ~Material() {}
    virtual bool scatter(const Ray& r_in, const HitRecord& rec, Ray& scattered, Color& attenuation) const = 0; // hit.rs:1014
    virtual Color emitted(double, double, const Point3&) const { return Color(0, 0, 0); }                      // hit.rs:1015-1017
};

struct Hittable {
    int32_t prim_id = -1; // depth-first leaf number, assigned at commit
    virtual ~Hittable() {}
    virtual bool hit(const Ray& r, double t_min, double t_max, HitRecord& rec) const = 0; // hit.rs:83
    virtual bool bounding_box(double time0, double time1, Aabb& out) const = 0;          // hit.rs:84
    // depth-first leaf numbering (not in the reference; defines rt_hit.prim_id)
    virtual void number_leaves(int32_t& next) { if (prim_id < 0) prim_id = next++; }
};
typedef std::shared_ptr<Hittable> HittablePtr;
typedef std::shared_ptr<Material> MaterialPtr;
typedef std::shared_ptr<Texture> TexturePtr;

// ------------------------------------------------------------------ hit.rs:87-178 Triangle
struct Triangle : Hittable {
    Point3 v0, v1, v2;
    Vec3 normal;
    MaterialPtr mat_ptr;
    Triangle(const Point3& a, const Point3& b, const Point3& c, MaterialPtr m) : v0(a), v1(b), v2(c), mat_ptr(m) {
        const Vec3 e0 = v1 - v0, e1 = v2 - v0; // hit.rs:96-107
        normal = e0.cross(e1).unit();
    }
    bool hit(const Ray& r, double t_min, double t_max, HitRecord& rec) const override { // hit.rs:111-162
        ctx().cnt.prim_tests[PT_TRI]++;
        if (std::fabs(normal.dot(r.direction)) < 0.0001) return false;
        const double d = -normal.dot(v0);
        const double t = -(normal.dot(r.origin) + d) / normal.dot(r.direction);
        if (t < t_min || t > t_max) return false;
        const Point3 p = r.at(t);
        const Vec3 edge0 = v1 - v0, vp0 = p - v0;
        if (normal.dot(edge0.cross(vp0)) < 0.0) return false;
        const Vec3 edge1 = v2 - v1, vp1 = p - v1;
        if (normal.dot(edge1.cross(vp1)) < 0.0) return false;
        const Vec3 edge2 = v0 - v2, vp2 = p - v2;
        if (normal.dot(edge2.cross(vp2)) < 0.0) return false;
        HitRecord::create_normal_face(r, normal, rec.normal, rec.front_face);
        rec.p = r.at(t); rec.t = t; rec.u = 1.0; rec.v = 1.0;
        rec.mat_ptr = mat_ptr.get(); rec.prim_id = prim_id;
        return true;
    }
    bool bounding_box(double, double, Aabb& out) const override { // hit.rs:164-177
        Point3 mn(INF, INF, INF), mx(-INF, -INF, -INF);
        for (const Point3* v : {&v0, &v1, &v2}) {
            mn.x = std::fmin(mn.x, v->x); mn.y = std::fmin(mn.y, v->y); mn.z = std::fmin(mn.z, v->z);
            mx.x = std::fmax(mx.x, v->x); mx.y = std::fmax(mx.y, v->y); mx.z = std::fmax(mx.z, v->z);
        }
        out = Aabb(mn, mx);
        return true;
    }
};

// shared by Sphere / MovingSphere / GravitySphere::hit (hit.rs:204-238, 282-316, 398-432)
inline bool sphere_roots(const Ray& r, const Point3& center, double radius, double t_min, double t_max, double& t) {
    const Vec3 oc = r.origin - center;
    const double a = r.direction.length_squared();
    const double half_b = oc.dot(r.direction);
    const double c = oc.length_squared() - radius * radius;
    const double discriminant = half_b * half_b - a * c;
    if (discriminant < 0.0) return false;
    const double sqrtd = std::sqrt(discriminant);
    double root = (-half_b - sqrtd) / a;
    if (root < t_min || t_max < root) {
        root = (-half_b + sqrtd) / a;
        if (root < t_min || t_max < root) return false;
    }
    t = root;
    return true;
}

// ------------------------------------------------------------------ hit.rs:180-245 Sphere
struct Sphere : Hittable {
    Point3 center;
    double radius;
    MaterialPtr mat_ptr;
    Sphere(const Point3& c, double r, MaterialPtr m) : center(c), radius(r), mat_ptr(m) {}
    static void get_sphere_uv(const Point3& p, double& u, double& v) { // hit.rs:195-200
        const double theta = std::acos(-p.y);
        const double phi = std::atan2(-p.z, p.x) + PI;
        u = phi / (2.0 * PI);
        v = theta / PI;
    }
    bool hit(const Ray& r, double t_min, double t_max, HitRecord& rec) const override { // hit.rs:204-238
        ctx().cnt.prim_tests[PT_SPHERE]++;
        double t;
        if (!sphere_roots(r, center, radius, t_min, t_max, t)) return false;
        const Point3 p = r.at(t);
        const Vec3 outward_normal = (p - center) / radius;
        HitRecord::create_normal_face(r, outward_normal, rec.normal, rec.front_face);
        get_sphere_uv(outward_normal, rec.u, rec.v);
        rec.p = p; rec.t = t; rec.mat_ptr = mat_ptr.get(); rec.prim_id = prim_id;
        return true;
    }
    bool bounding_box(double, double, Aabb& out) const override { // hit.rs:239-244
        out = Aabb(center - Point3(radius, radius, radius), center + Point3(radius, radius, radius));
        return true;
    }
};

// ------------------------------------------------------------------ hit.rs:247-328 MovingSphere
struct MovingSphere : Hittable {
    Point3 center0, center1;
    double time0, time1, radius;
    MaterialPtr mat_ptr;
    MovingSphere(const Point3& c0, const Point3& c1, double t0, double t1, double r, MaterialPtr m)
        : center0(c0), center1(c1), time0(t0), time1(t1), radius(r), mat_ptr(m) {}
    Point3 get_center(double time) const { // hit.rs:275-278
        return center0 + ((time - time0) / (time1 - time0)) * (center1 - center0);
    }
    bool hit(const Ray& r, double t_min, double t_max, HitRecord& rec) const override { // hit.rs:282-316
        ctx().cnt.prim_tests[PT_MOVING]++;
        const Point3 cur = get_center(r.time);
        double t;
        if (!sphere_roots(r, cur, radius, t_min, t_max, t)) return false;
        const Point3 p = r.at(t);
        const Vec3 outward_normal = (p - cur) / radius;
        HitRecord::create_normal_face(r, outward_normal, rec.normal, rec.front_face);
        rec.p = p; rec.t = t; rec.u = 0.0; rec.v = 0.0; rec.mat_ptr = mat_ptr.get(); rec.prim_id = prim_id;
        return true;
    }
    bool bounding_box(double t0, double t1, Aabb& out) const override { // hit.rs:317-327
        const Point3 rr(radius, radius, radius);
        const Aabb box0(get_center(t0) - rr, get_center(t0) + rr);
        const Aabb box1(get_center(t1) - rr, get_center(t1) + rr);
        out = Aabb::surrounding_box(box0, box1);
        return true;
    }
};

// ------------------------------------------------------------------ hit.rs:330-444 GravitySphere
struct GravitySphere : Hittable {
    Point3 start;
    double time0, radius;
    MaterialPtr mat_ptr;
    std::vector<double> stored;
    GravitySphere(const Point3& s, double t0, double r, MaterialPtr m) : start(s), time0(t0), radius(r), mat_ptr(m) {
        // hit.rs:346-359
        stored.push_back(start.y);
        const double incr = 0.001;
        double t = time0;
        Point3 cur_pos = start;
        double vel = 0.0;
        while (t < 100.0) {
            t += incr;
            vel -= 0.000001;
            if (cur_pos.y - 1.0 * radius <= 0.0) vel *= -0.92;
            cur_pos.y = std::fmax(1.0 * radius, cur_pos.y + vel);
            stored.push_back(cur_pos.y);
        }
    }
    Point3 get_center(double time) const { // hit.rs:370-394
        const double incr = 0.001;
        const double q = time / incr;
        // Rust `as usize`: saturating, NaN -> 0, negative -> 0
        size_t idx = 0;
        if (q == q && q > 0.0) idx = (q >= 1.8446744073709552e19) ? (size_t)-1 : (size_t)q;
        if (idx != (size_t)-1 && idx + 1 <= stored.size()) return Vec3(start.x, stored[idx], start.z);
        double t = time0;
        Point3 cur_pos = start;
        double vel = 0.0;
        while (t < time) {
            t += incr;
            vel -= 0.000001;
            if (cur_pos.y - 2.0 * radius <= 0.0) vel *= -0.8;
            cur_pos.y = std::fmax(2.0 * radius, cur_pos.y + vel);
        }
        return cur_pos;
    }
    bool hit(const Ray& r, double t_min, double t_max, HitRecord& rec) const override { // hit.rs:398-432
        ctx().cnt.prim_tests[PT_GRAVITY]++;
        const Point3 cur = get_center(r.time);
        double t;
        if (!sphere_roots(r, cur, radius, t_min, t_max, t)) return false;
        const Point3 p = r.at(t);
        const Vec3 outward_normal = (p - cur) / radius;
        HitRecord::create_normal_face(r, outward_normal, rec.normal, rec.front_face);
        rec.p = p; rec.t = t; rec.u = 0.0; rec.v = 0.0; rec.mat_ptr = mat_ptr.get(); rec.prim_id = prim_id;
        return true;
    }
    bool bounding_box(double t0, double t1, Aabb& out) const override { // hit.rs:433-443
        const Point3 rr(radius, radius, radius);
        const Aabb box0(get_center(t0) - rr, get_center(t0) + rr);
        const Aabb box1(get_center(t1) - rr, get_center(t1) + rr);
        out = Aabb::surrounding_box(box0, box1);
        return true;
    }
};

// ------------------------------------------------------------------ hit.rs:446-639 axis rects
// axis = the constant axis: 2 -> XyRect, 1 -> XzRect, 0 -> YzRect.  (a, b) are the in-plane axes
// in the reference's field order: Xy (x,y), Xz (x,z), Yz (y,z).
struct AxisRect : Hittable {
    int axis;
    double a0, a1, b0, b1, k;
    MaterialPtr mat_ptr;
    AxisRect(int ax, double a0_, double a1_, double b0_, double b1_, double k_, MaterialPtr m)
        : axis(ax), a0(a0_), a1(a1_), b0(b0_), b1(b1_), k(k_), mat_ptr(m) {}
    bool hit(const Ray& r, double t_min, double t_max, HitRecord& rec) const override { // hit.rs:476-501, 541-566, 606-631
        ctx().cnt.prim_tests[PT_RECT]++;
        const int ia = axis == 0 ? 1 : 0;
        const int ib = axis == 2 ? 1 : 2;
        const double t = (k - r.origin.axis(axis)) / r.direction.axis(axis);
        if (t < t_min || t > t_max) return false;
        const double x = r.origin.axis(ia) + t * r.direction.axis(ia);
        const double y = r.origin.axis(ib) + t * r.direction.axis(ib);
        if (x < a0 || x > a1 || y < b0 || y > b1) return false;
        rec.u = (x - a0) / (a1 - a0);
        rec.v = (y - b0) / (b1 - b0);
        const Vec3 outward_normal(axis == 0 ? 1.0 : 0.0, axis == 1 ? 1.0 : 0.0, axis == 2 ? 1.0 : 0.0);
        HitRecord::create_normal_face(r, outward_normal, rec.normal, rec.front_face);
        rec.p = r.at(t); rec.t = t; rec.mat_ptr = mat_ptr.get(); rec.prim_id = prim_id;
        return true;
    }
    bool bounding_box(double, double, Aabb& out) const override { // hit.rs:503-508, 568-573, 633-638
        if (axis == 2) out = Aabb(Point3(a0, b0, k - 0.0001), Point3(a1, b1, k + 0.0001));
        else if (axis == 1) out = Aabb(Point3(a0, k - 0.0001, b0), Point3(a1, k + 0.0001, b1));
        else out = Aabb(Point3(k - 0.0001, a0, b0), Point3(k + 0.0001, a1, b1));
        return true;
    }
};

// ------------------------------------------------------------------ hit.rs:641-711 HittableList
struct HittableList : Hittable {
    std::vector<HittablePtr> objects;
    void add(HittablePtr o) { objects.push_back(o); } // hit.rs:650-652
    bool hit(const Ray& r, double t_min, double t_max, HitRecord& rec) const override { // hit.rs:660-690
        bool hit_anything = false;
        double closest_so_far = t_max;
        HitRecord temp;
        for (const HittablePtr& object : objects) {
            if (object->hit(r, t_min, closest_so_far, temp)) {
                hit_anything = true;
                closest_so_far = temp.t;
                rec = temp;
            }
        }
        return hit_anything;
    }
    bool bounding_box(double t0, double t1, Aabb& out) const override { // hit.rs:691-710
        if (objects.empty()) return false;
        Aabb temp_box;
        if (!objects[0]->bounding_box(t0, t1, temp_box)) return false;
        for (size_t i = 1; i < objects.size(); ++i) {
            Aabb other;
            if (!objects[i]->bounding_box(t0, t1, other)) return false;
            temp_box = Aabb::surrounding_box(temp_box, other);
        }
        out = temp_box;
        return true;
    }
    void number_leaves(int32_t& next) override {
        for (HittablePtr& o : objects) o->number_leaves(next);
    }
};

// ------------------------------------------------------------------ hit.rs:713-785 RectPrism
struct RectPrism : Hittable {
    Point3 box_min, box_max;
    HittableList sides;
    RectPrism(const Point3& p0, const Point3& p1, MaterialPtr mat) : box_min(p0), box_max(p1) { // hit.rs:720-775
        sides.add(std::make_shared<AxisRect>(2, p0.x, p1.x, p0.y, p1.y, p1.z, mat));
        sides.add(std::make_shared<AxisRect>(2, p0.x, p1.x, p0.y, p1.y, p0.z, mat));
        sides.add(std::make_shared<AxisRect>(1, p0.x, p1.x, p0.z, p1.z, p1.y, mat));
        sides.add(std::make_shared<AxisRect>(1, p0.x, p1.x, p0.z, p1.z, p0.y, mat));
        sides.add(std::make_shared<AxisRect>(0, p0.y, p1.y, p0.z, p1.z, p1.x, mat));
        sides.add(std::make_shared<AxisRect>(0, p0.y, p1.y, p0.z, p1.z, p0.x, mat));
    }
    bool hit(const Ray& r, double t_min, double t_max, HitRecord& rec) const override { // hit.rs:779-781
        return sides.hit(r, t_min, t_max, rec);
    }
    bool bounding_box(double, double, Aabb& out) const override { // hit.rs:782-784
        out = Aabb(box_min, box_max);
        return true;
    }
    void number_leaves(int32_t& next) override { sides.number_leaves(next); }
};

// ------------------------------------------------------------------ hit.rs:787-833 Translate
struct Translate : Hittable {
    HittablePtr obj;
    Vec3 offset;
    Translate(const Vec3& off, HittablePtr o) : obj(o), offset(off) {}
    bool hit(const Ray& r, double t_min, double t_max, HitRecord& rec) const override { // hit.rs:802-823
        const Ray moved_r(r.origin - offset, r.direction, r.time);
        HitRecord in;
        if (!obj->hit(moved_r, t_min, t_max, in)) return false;
        Vec3 normal; bool front_face;
        HitRecord::create_normal_face(moved_r, in.normal, normal, front_face);
        rec = in;
        rec.p = in.p + offset; rec.normal = normal; rec.front_face = front_face;
        return true;
    }
    bool bounding_box(double t0, double t1, Aabb& out) const override { // hit.rs:824-832
        Aabb a;
        if (!obj->bounding_box(t0, t1, a)) return false;
        out = Aabb(a.minimum + offset, a.maximum + offset);
        return true;
    }
    void number_leaves(int32_t& next) override { obj->number_leaves(next); }
};

// ------------------------------------------------------------------ hit.rs:835-936 RotateY
struct RotateY : Hittable {
    HittablePtr obj;
    double sin_theta, cos_theta;
    bool has_box;
    Aabb bbox;
    RotateY(double angle_deg, HittablePtr o) : obj(o) { // hit.rs:843-888
        const double angle = angle_deg * (PI / 180.0); // f64::to_radians
        sin_theta = std::sin(angle);
        cos_theta = std::cos(angle);
        // the reference computes the rotated corner min/max and then stores the UN-rotated box
        // (hit.rs:886 uses `bbox`, not Aabb::new(min,max)); reproduced.
        has_box = obj->bounding_box(0.0, 1.0, bbox);
    }
    bool hit(const Ray& r, double t_min, double t_max, HitRecord& rec) const override { // hit.rs:892-931
        const Vec3 origin(cos_theta * r.origin.x - sin_theta * r.origin.z, r.origin.y,
                          sin_theta * r.origin.x + cos_theta * r.origin.z);
        const Vec3 direction(cos_theta * r.direction.x - sin_theta * r.direction.z, r.direction.y,
                             sin_theta * r.direction.x + cos_theta * r.direction.z);
        const Ray rotated_r(origin, direction, r.time);
        HitRecord in;
        if (!obj->hit(rotated_r, t_min, t_max, in)) return false;
        const Vec3 p(cos_theta * in.p.x + sin_theta * in.p.z, in.p.y, -sin_theta * in.p.x + cos_theta * in.p.z);
        const Vec3 normal(cos_theta * in.normal.x + sin_theta * in.normal.z, in.normal.y,
                          -sin_theta * in.normal.x + cos_theta * in.normal.z);
        Vec3 n2; bool front_face;
        // face-forwarded against the OBJECT-space ray (book quirk, hit.rs:921)
        HitRecord::create_normal_face(rotated_r, normal, n2, front_face);
        rec = in;
        rec.p = p; rec.normal = n2; rec.front_face = front_face;
        return true;
    }
    bool bounding_box(double, double, Aabb& out) const override { // hit.rs:933-935
        if (!has_box) return false;
        out = bbox;
        return true;
    }
    void number_leaves(int32_t& next) override { obj->number_leaves(next); }
};

// ------------------------------------------------------------------ materials hit.rs:992-1152
struct Isotropic : Material {
    TexturePtr albedo;
    explicit Isotropic(TexturePtr a) : albedo(a) {}
    bool scatter(const Ray& r_in, const HitRecord& rec, Ray& scattered, Color& attenuation) const override { // hit.rs:1004-1011
        ctx().cnt.scatters[MT_ISOTROPIC]++;
        ctx().begin_event();
        scattered = Ray(rec.p, random_in_unit_sphere(), r_in.time);
        attenuation = albedo->value(rec.u, rec.v, rec.p);
        return true;
    }
};
struct Lambertian : Material {
    TexturePtr albedo;
    explicit Lambertian(TexturePtr a) : albedo(a) {}
    bool scatter(const Ray& r_in, const HitRecord& rec, Ray& scattered, Color& attenuation) const override { // hit.rs:1039-1051
        ctx().cnt.scatters[MT_LAMBERTIAN]++;
        ctx().begin_event();
        Vec3 scatter_direction = rec.normal + random_unit_vector();
        if (scatter_direction.near_zero()) scatter_direction = rec.normal;
        scattered = Ray(rec.p, scatter_direction, r_in.time);
        attenuation = albedo->value(rec.u, rec.v, rec.p);
        return true;
    }
};
struct Metal : Material {
    Color albedo;
    double fuzz;
    Metal(const Color& a, double f) : albedo(a), fuzz(f < 1.0 ? f : 1.0) {} // hit.rs:1060-1065
    bool scatter(const Ray& r_in, const HitRecord& rec, Ray& scattered, Color& attenuation) const override { // hit.rs:1069-1083
        ctx().cnt.scatters[MT_METAL]++;
        ctx().begin_event();
        const Vec3 reflected = r_in.direction.unit().reflect(rec.normal);
        scattered = Ray(rec.p, reflected + fuzz * random_in_unit_sphere(), r_in.time);
        if (scattered.direction.dot(rec.normal) > 0.0) {
            attenuation = albedo;
            return true;
        }
        return false;
    }
};
struct Dielectric : Material {
    double ir;
    explicit Dielectric(double i) : ir(i) {}
    static double reflectance(double cosine, double ref_idx) { // hit.rs:1095-1099
        double r0 = (1.0 - ref_idx) / (1.0 + ref_idx);
        r0 = r0 * r0;
        const double x = 1.0 - cosine; // f64::powi(x, 5)
        return r0 + (1.0 - r0) * (x * x * x * x * x);
    }
    bool scatter(const Ray& r_in, const HitRecord& rec, Ray& scattered, Color& attenuation) const override { // hit.rs:1103-1126
        ctx().cnt.scatters[MT_DIELECTRIC]++;
        ctx().begin_event();
        attenuation = Vec3(1, 1, 1);
        const double refraction_ratio = rec.front_face ? 1.0 / ir : ir;
        const Vec3 unit_direction = r_in.direction.unit();
        const double cos_theta = std::fmin((-unit_direction).dot(rec.normal), 1.0);
        const double sin_theta = std::sqrt(1.0 - cos_theta * cos_theta);
        const bool cannot_refract = refraction_ratio * sin_theta > 1.0;
        Vec3 direction;
        // `||` short-circuits: the random number is drawn only when refraction is possible
        if (cannot_refract || reflectance(cos_theta, refraction_ratio) > ctx().gen())
            direction = unit_direction.reflect(rec.normal);
        else
            direction = refract(unit_direction, rec.normal, refraction_ratio);
        scattered = Ray(rec.p, direction, r_in.time);
        return true;
    }
};
struct DiffuseLight : Material {
    TexturePtr emit;
    explicit DiffuseLight(TexturePtr e) : emit(e) {}
    bool scatter(const Ray&, const HitRecord&, Ray&, Color&) const override { // hit.rs:1146-1148
        ctx().cnt.scatters[MT_LIGHT]++;
        return false;
    }
    Color emitted(double u, double v, const Point3& p) const override { return emit->value(u, v, p); } // hit.rs:1149-1151
};

// ------------------------------------------------------------------ hit.rs:938-990 ConstantMedium
struct ConstantMedium : Hittable {
    HittablePtr boundary;
    MaterialPtr phase_function;
    double neg_inv_density;
    ConstantMedium(MaterialPtr phase, double d, HittablePtr b) : boundary(b), phase_function(phase), neg_inv_density(-1.0 / d) {}
    bool hit(const Ray& r, double t_min, double t_max, HitRecord& rec) const override { // hit.rs:955-986
        if (!ctx().media_enabled) return false; // geometric parity batches skip media (rt_trace_batch flag)
        ctx().cnt.medium_queries++;
        HitRecord rec1, rec2;
        if (!boundary->hit(r, -INF, INF, rec1)) return false;
        if (!boundary->hit(r, rec1.t + 0.0001, INF, rec2)) return false;
        double t1 = std::fmax(rec1.t, t_min);
        const double t2 = std::fmin(rec2.t, t_max);
        if (t1 >= t2) return false;
        if (t1 < 0.0) t1 = 0.0;
        const double ray_length = r.direction.length();
        const double distance_inside_boundary = (t2 - t1) * ray_length;
        const double hit_distance = neg_inv_density * std::log(ctx().medium_xi((uint32_t)prim_id));
        if (hit_distance > distance_inside_boundary) return false;
        const double t = t1 + hit_distance / ray_length;
        rec.p = r.at(t); rec.normal = Vec3(0, 0, 0); rec.t = t; rec.u = 0.0; rec.v = 0.0;
        rec.front_face = true; rec.mat_ptr = phase_function.get(); rec.prim_id = prim_id;
        return true;
    }
    bool bounding_box(double t0, double t1, Aabb& out) const override { return boundary->bounding_box(t0, t1, out); } // hit.rs:987-989
    void number_leaves(int32_t& next) override {
        if (prim_id < 0) prim_id = next++;
        boundary->number_leaves(next);
    }
};

// ------------------------------------------------------------------ bvh.rs
struct SplitMix64 { // stands in for thread_rng() at bvh.rs:21 (axis choice) — topology only
    uint64_t s;
    explicit SplitMix64(uint64_t seed) : s(seed) {}
    uint64_t next() {
        uint64_t z = (s += 0x9E3779B97F4A7C15ull);
        z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
        z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
        return z ^ (z >> 31);
    }
    double uniform() { return (double)(next() >> 11) * (1.0 / 9007199254740992.0); }
};

struct BvhNode : Hittable {
    HittablePtr left, right;
    Aabb bbox;
    // bvh.rs:14-83.  The reference clones the object vector at every node and sorts the clone's
    // [start,end) span; children only ever read their own sub-span of their parent's clone, so
    // sorting one shared vector in place yields the same tree.  Returns false for a child with
    // no bounding box (the reference panics, bvh.rs:73,76).
    BvhNode() {}
    static bool build(std::vector<HittablePtr>& objects, size_t start, size_t end, double time0, double time1,
                      SplitMix64& rng, std::shared_ptr<BvhNode>& out) {
        std::shared_ptr<BvhNode> node = std::make_shared<BvhNode>();
        const int axis = (int)(rng.next() % 2); // gen_range(0..2): x or y, never z (bvh.rs:24)
        auto less = [axis](const HittablePtr& a, const HittablePtr& b) { // bvh.rs:25-46
            Aabb ba, bb;
            a->bounding_box(0.0, 0.0, ba);
            b->bounding_box(0.0, 0.0, bb);
            return ba.minimum.axis(axis) < bb.minimum.axis(axis);
        };
        const size_t object_span = end - start;
        if (object_span == 1) {
            node->left = objects[start];
            node->right = objects[start];
        } else if (object_span == 2) {
            if (less(objects[start], objects[start + 1])) {
                node->left = objects[start];
                node->right = objects[start + 1];
            } else {
                node->left = objects[start + 1];
                node->right = objects[start];
            }
        } else {
            // Rust's sort_by is a stable merge sort; the comparator (never Equal) orders by min[axis]
            std::vector<std::pair<double, HittablePtr>> keyed;
            keyed.reserve(object_span);
            for (size_t i = start; i < end; ++i) {
                Aabb b;
                objects[i]->bounding_box(0.0, 0.0, b);
                keyed.emplace_back(b.minimum.axis(axis), objects[i]);
            }
            std::stable_sort(keyed.begin(), keyed.end(),
                             [](const std::pair<double, HittablePtr>& a, const std::pair<double, HittablePtr>& b) { return a.first < b.first; });
            for (size_t i = 0; i < object_span; ++i) objects[start + i] = keyed[i].second;
            const size_t mid = start + object_span / 2;
            std::shared_ptr<BvhNode> l, r;
            if (!build(objects, start, mid, time0, time1, rng, l)) return false;
            if (!build(objects, mid, end, time0, time1, rng, r)) return false;
            node->left = l;
            node->right = r;
        }
        Aabb lb, rb;
        if (!node->left->bounding_box(time0, time1, lb)) return false;
        if (!node->right->bounding_box(time0, time1, rb)) return false;
        node->bbox = Aabb::surrounding_box(lb, rb);
        out = node;
        return true;
    }
    bool hit(const Ray& r, double t_min, double t_max, HitRecord& rec) const override { // bvh.rs:97-112
        if (!bbox.hit(r, t_min, t_max)) return false;
        // A node over a single object stores it as both children (bvh.rs:53-55) and tests it twice;
        // the event counters count that test once ("minus the duplicate single-leaf tests").
        const bool dup = left.get() == right.get();
        HitRecord l;
        if (left->hit(r, t_min, t_max, l)) {
            HitRecord rr;
            const Counters saved = ctx().cnt;
            const bool rh = right->hit(r, t_min, l.t, rr);
            if (dup) ctx().cnt = saved;
            if (rh) { rec = rr; return true; }
            rec = l;
            return true;
        }
        const Counters saved = ctx().cnt;
        const bool rh = right->hit(r, t_min, t_max, rec);
        if (dup) ctx().cnt = saved;
        return rh;
    }
    bool bounding_box(double, double, Aabb& out) const override { out = bbox; return true; } // bvh.rs:113-116
    void number_leaves(int32_t&) override {} // numbered by BvhGroup in list order
};

// BvhNode::from_list(list, time0, time1) (bvh.rs:85-93): keeps the list so that leaves are
// numbered in the list's construction order (the tree order depends on the random axes).
struct BvhGroup : Hittable {
    std::vector<HittablePtr> original;
    std::shared_ptr<BvhNode> root;
    bool hit(const Ray& r, double t_min, double t_max, HitRecord& rec) const override { return root->hit(r, t_min, t_max, rec); }
    bool bounding_box(double t0, double t1, Aabb& out) const override { return root->bounding_box(t0, t1, out); }
    void number_leaves(int32_t& next) override {
        for (HittablePtr& o : original) o->number_leaves(next);
    }
};

// ------------------------------------------------------------------ camera.rs
struct Camera {
    Point3 origin, lower_left_corner;
    Vec3 horizontal, vertical, u, v, w;
    double lens_radius = 0, time1 = 0, time2 = 0;
    Camera() {}
    Camera(const Point3& lookfrom, const Point3& lookat, const Vec3& vup, double vfov, double aspect_ratio, double aperture,
           double focus_dist, double t1, double t2) { // camera.rs:20-57
        const double theta = vfov * (PI / 180.0);
        const double h = std::tan(theta / 2.0);
        const double viewport_height = 2.0 * h;
        const double viewport_width = aspect_ratio * viewport_height;
        w = (lookfrom - lookat).unit();
        u = vup.cross(w).unit();
        v = w.cross(u);
        origin = lookfrom;
        horizontal = focus_dist * viewport_width * u;
        vertical = focus_dist * viewport_height * v;
        lower_left_corner = origin - horizontal / 2.0 - vertical / 2.0 - focus_dist * w;
        lens_radius = aperture / 2.0;
        time1 = t1; time2 = t2;
    }
    Ray get_ray(double s, double t) const { // camera.rs:59-71
        const Vec3 rd = lens_radius * random_in_unit_disk();
        const Vec3 offset = u * rd.x + v * rd.y;
        const Vec3 dir = lower_left_corner + s * horizontal + t * vertical - origin - offset;
        const double time = ctx().gen_range(time1, time2);
        return Ray(origin + offset, dir, time);
    }
};

// ------------------------------------------------------------------ world.rs:52-93 ray_color
// Background of a miss.  world.rs:86-89 multiplies by one constant `background`; the earlier revision that rendered the shipped
// images/book1.png used the book-1 sky instead (unit(d).y blended white -> blue), kept here as an option so that image can pin the port.
struct Background {
    Color c0 = Color(0, 0, 0), c1 = Color(0, 0, 0);
    bool gradient = false;
    Color at(const Ray& r) const {
        if (!gradient) return c0;
        const double t = 0.5 * (r.direction.unit().y + 1.0);
        return (1.0 - t) * c0 + t * c1;
    }
};
inline Color ray_color(const Ray& r, const Background& background, const Hittable& world, int32_t depth) {
    Vec3 product(1, 1, 1), output(0, 0, 0);
    Ray current_ray = r;
    PathCtx& c = ctx();
    for (;;) {
        depth -= 1;
        if (depth < 0) break;
        HitRecord rec;
        c.cnt.segments++;
        const bool got = world.hit(current_ray, 0.001, INF, rec);
        c.segment++;
        if (got) {
            Ray scattered; Color attenuation;
            if (rec.mat_ptr->scatter(current_ray, rec, scattered, attenuation)) {
                const Color emitted = rec.mat_ptr->emitted(rec.u, rec.v, rec.p);
                output += emitted * product;
                product *= attenuation;
                current_ray = scattered;
            } else {
                const Color emitted = rec.mat_ptr->emitted(rec.u, rec.v, rec.p);
                output += emitted * product;
                break;
            }
        } else {
            output += product * background.at(current_ray);
            break;
        }
    }
    return output;
}

} // namespace orc
