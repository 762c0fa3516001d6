// oracle_capi.cpp — the oracle behind the same C-ABI as the product (include/rtb200.h compiled
// with -DRTB_PREFIX_ORC => orc_*), plus orc_kat_* hooks that expose single reference functions
// to the known-answer tests.
//
// TEST INFRASTRUCTURE ONLY (see rt_oracle.hpp).  Also serves as the CPU baseline: orc_render is
// the reference's render_scene (src/world.rs:1181-1247): `threads` OS threads, each owning a
// static contiguous band of rows, per-pixel sample loop, get_normalized_color per pixel.
#define RTB_PREFIX_ORC 1
#include "rtb200.h"

#include <atomic>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <sstream>
#include <thread>

#include "rt_oracle.hpp"
#include "world.hpp"

using namespace orc;

namespace {
thread_local std::string g_err;
int32_t fail(int32_t code, const std::string& msg) {
    g_err = msg;
    return code;
}
} // namespace

struct rt_scene {
    std::vector<TexturePtr> tex;
    std::vector<MaterialPtr> mats;
    std::vector<HittablePtr> objs;
    HittablePtr root;
    Camera cam;
    bool has_cam = false;
    Background background;
    int32_t n_prims = 0;
    bool committed = false;
    uint64_t bvh_seed = 0x0B5EEDull;
};

#define CHECK_SCENE(s) \
    if (!(s)) return fail(RT_ERR_INVALID, "null scene")
#define CHECK_TEX(s, id) \
    if ((id) < 0 || (size_t)(id) >= (s)->tex.size()) return fail(RT_ERR_INVALID, "texture id out of range")
#define CHECK_MAT(s, id) \
    if ((id) < 0 || (size_t)(id) >= (s)->mats.size()) return fail(RT_ERR_INVALID, "material id out of range")
#define CHECK_OBJ(s, id) \
    if ((id) < 0 || (size_t)(id) >= (s)->objs.size()) return fail(RT_ERR_INVALID, "hittable id out of range")

static int32_t add_tex(rt_scene* s, TexturePtr t) {
    s->tex.push_back(t);
    return (int32_t)s->tex.size() - 1;
}
static int32_t add_mat(rt_scene* s, MaterialPtr m) {
    m->mat_id = (int32_t)s->mats.size();
    s->mats.push_back(m);
    return m->mat_id;
}
static int32_t add_obj(rt_scene* s, HittablePtr o) {
    s->objs.push_back(o);
    s->committed = false;
    return (int32_t)s->objs.size() - 1;
}

// screen.rs:61-95 Screen::from_ppm_p3
static bool read_ppm_p3(const char* name, ImageData& img, std::string& err) {
    std::ifstream f(name, std::ios::binary);
    if (!f) { err = std::string("Couldn't open the file ") + name; return false; }
    std::stringstream ss;
    ss << f.rdbuf();
    const std::string contents = ss.str();
    std::vector<std::string> lines;
    {
        size_t pos = 0;
        for (;;) {
            const size_t nl = contents.find('\n', pos);
            if (nl == std::string::npos) { lines.push_back(contents.substr(pos)); break; }
            lines.push_back(contents.substr(pos, nl - pos));
            pos = nl + 1;
        }
    }
    if (lines.size() < 3) { err = "ppm: short header"; return false; }
    size_t width = 0, height = 0;
    {
        const std::string& wh = lines[1];
        const size_t sp = wh.find(' ');
        if (sp == std::string::npos) { err = "ppm: bad size line"; return false; }
        width = (size_t)std::strtoull(wh.substr(0, sp).c_str(), nullptr, 10);
        height = (size_t)std::strtoull(wh.substr(sp + 1).c_str(), nullptr, 10);
    }
    if (width == 0 || height == 0) { err = "ppm: zero size"; return false; }
    std::vector<double> nums;
    nums.reserve(width * height * 3);
    for (size_t li = 3; li < lines.size(); ++li) {
        std::istringstream ls(lines[li]);
        std::string tok;
        while (ls >> tok) nums.push_back(std::strtod(tok.c_str(), nullptr));
    }
    if (nums.size() < width * height * 3) { err = "ppm: not enough samples"; return false; }
    img.width = width;
    img.height = height;
    img.pixels.resize(width * height);
    size_t it = 0;
    for (size_t j = 0; j < height; ++j)
        for (size_t i = 0; i < width; ++i) {
            img.pixels[j * width + i] = Color(nums[it], nums[it + 1], nums[it + 2]);
            it += 3;
        }
    return true;
}

// model.rs:13-62 TriangleModel::load_from_file
static bool load_ply(const char* path, double scale, std::vector<Point3>& vertices, std::vector<uint32_t>& faces, std::string& err) {
    std::ifstream f(path, std::ios::binary);
    if (!f) { err = std::string("Couldn't open the file ") + path; return false; }
    std::string line;
    long vertex_count = 0, face_count = 0;
    bool header_done = false;
    while (std::getline(f, line, '\n')) {
        if (line == "end_header") { header_done = true; break; }
        // line.split(" ")
        std::vector<std::string> parts;
        size_t pos = 0;
        for (;;) {
            const size_t sp = line.find(' ', pos);
            if (sp == std::string::npos) { parts.push_back(line.substr(pos)); break; }
            parts.push_back(line.substr(pos, sp - pos));
            pos = sp + 1;
        }
        if (parts[0] == "element" && parts.size() >= 3) {
            if (parts[1] == "vertex") vertex_count = std::strtol(parts[2].c_str(), nullptr, 10);
            if (parts[1] == "face") face_count = std::strtol(parts[2].c_str(), nullptr, 10);
        }
    }
    if (!header_done) { err = "ply: no end_header"; return false; }
    vertices.reserve((size_t)vertex_count);
    for (long i = 0; i < vertex_count; ++i) {
        if (!std::getline(f, line, '\n')) { err = "ply: truncated vertex list"; return false; }
        const char* p = line.c_str();
        char* e;
        const double x = std::strtod(p, &e); p = e;
        const double y = std::strtod(p, &e); p = e;
        const double z = std::strtod(p, &e);
        vertices.push_back(Point3(x * scale, y * scale, z * scale));
    }
    faces.reserve((size_t)face_count * 3);
    for (long i = 0; i < face_count; ++i) {
        if (!std::getline(f, line, '\n')) { err = "ply: truncated face list"; return false; }
        const char* p = line.c_str();
        char* e;
        (void)std::strtoul(p, &e, 10); p = e; // token 0 (the count) is ignored, model.rs:53-57
        const unsigned long a = std::strtoul(p, &e, 10); p = e;
        const unsigned long b = std::strtoul(p, &e, 10); p = e;
        const unsigned long c = std::strtoul(p, &e, 10);
        if (a >= vertices.size() || b >= vertices.size() || c >= vertices.size()) { err = "ply: vertex index out of range"; return false; }
        faces.push_back((uint32_t)a); faces.push_back((uint32_t)b); faces.push_back((uint32_t)c);
    }
    return true;
}

extern "C" {

rt_scene* orc_scene_create(void) { return new rt_scene(); }
void orc_scene_destroy(rt_scene* s) { delete s; }
const char* orc_last_error(void) { return g_err.c_str(); }
const char* orc_version(void) { return "oracle f64 restatement of ray-tracing-series-rust (CPU)"; }

int32_t orc_tex_solid(rt_scene* s, const double rgb[3]) {
    CHECK_SCENE(s);
    return add_tex(s, std::make_shared<SolidColor>(Color(rgb[0], rgb[1], rgb[2])));
}
int32_t orc_tex_checker(rt_scene* s, int32_t even, int32_t odd) {
    CHECK_SCENE(s); CHECK_TEX(s, even); CHECK_TEX(s, odd);
    return add_tex(s, std::make_shared<Checker>(s->tex[(size_t)even], s->tex[(size_t)odd]));
}
int32_t orc_tex_noise(rt_scene* s, double scale, const double* ranvec, const int32_t* px, const int32_t* py, const int32_t* pz, uint64_t seed) {
    CHECK_SCENE(s);
    rtb::PerlinTables t;
    if (!ranvec || !px || !py || !pz) {
        rtb::perlin_generate(seed, t);
        ranvec = t.ranvec; px = t.perm_x; py = t.perm_y; pz = t.perm_z;
    }
    Perlin p;
    p.ranvec.resize(256);
    for (int i = 0; i < 256; ++i) p.ranvec[(size_t)i] = Vec3(ranvec[3 * i], ranvec[3 * i + 1], ranvec[3 * i + 2]);
    p.perm_x.assign(px, px + 256);
    p.perm_y.assign(py, py + 256);
    p.perm_z.assign(pz, pz + 256);
    for (int i = 0; i < 256; ++i)
        if ((p.perm_x[(size_t)i] | p.perm_y[(size_t)i] | p.perm_z[(size_t)i]) & ~255) return fail(RT_ERR_INVALID, "perm entry outside 0..255");
    return add_tex(s, std::make_shared<Noise>(p, scale));
}
int32_t orc_tex_image(rt_scene* s, int32_t w, int32_t h, const double* rgb) {
    CHECK_SCENE(s);
    if (w <= 0 || h <= 0 || !rgb) return fail(RT_ERR_INVALID, "bad image");
    ImageData img;
    img.width = (size_t)w; img.height = (size_t)h;
    img.pixels.resize((size_t)w * h);
    for (size_t i = 0; i < (size_t)w * h; ++i) img.pixels[i] = Color(rgb[3 * i], rgb[3 * i + 1], rgb[3 * i + 2]);
    return add_tex(s, std::make_shared<Image>(std::move(img)));
}
int32_t orc_tex_image_ppm(rt_scene* s, const char* path) {
    CHECK_SCENE(s);
    ImageData img;
    std::string err;
    if (!path || !read_ppm_p3(path, img, err)) return fail(RT_ERR_IO, err);
    return add_tex(s, std::make_shared<Image>(std::move(img)));
}

int32_t orc_mat_lambertian(rt_scene* s, int32_t tex) { CHECK_SCENE(s); CHECK_TEX(s, tex); return add_mat(s, std::make_shared<Lambertian>(s->tex[(size_t)tex])); }
int32_t orc_mat_metal(rt_scene* s, const double a[3], double fuzz) { CHECK_SCENE(s); return add_mat(s, std::make_shared<Metal>(Color(a[0], a[1], a[2]), fuzz)); }
int32_t orc_mat_dielectric(rt_scene* s, double ir) { CHECK_SCENE(s); return add_mat(s, std::make_shared<Dielectric>(ir)); }
int32_t orc_mat_diffuse_light(rt_scene* s, int32_t tex) { CHECK_SCENE(s); CHECK_TEX(s, tex); return add_mat(s, std::make_shared<DiffuseLight>(s->tex[(size_t)tex])); }
int32_t orc_mat_isotropic(rt_scene* s, int32_t tex) { CHECK_SCENE(s); CHECK_TEX(s, tex); return add_mat(s, std::make_shared<Isotropic>(s->tex[(size_t)tex])); }

int32_t orc_sphere(rt_scene* s, const double c[3], double r, int32_t mat) {
    CHECK_SCENE(s); CHECK_MAT(s, mat);
    return add_obj(s, std::make_shared<Sphere>(Point3(c[0], c[1], c[2]), r, s->mats[(size_t)mat]));
}
int32_t orc_moving_sphere(rt_scene* s, const double c0[3], const double c1[3], double t0, double t1, double r, int32_t mat) {
    CHECK_SCENE(s); CHECK_MAT(s, mat);
    return add_obj(s, std::make_shared<MovingSphere>(Point3(c0[0], c0[1], c0[2]), Point3(c1[0], c1[1], c1[2]), t0, t1, r, s->mats[(size_t)mat]));
}
int32_t orc_gravity_sphere(rt_scene* s, const double st[3], double t0, double r, int32_t mat) {
    CHECK_SCENE(s); CHECK_MAT(s, mat);
    return add_obj(s, std::make_shared<GravitySphere>(Point3(st[0], st[1], st[2]), t0, r, s->mats[(size_t)mat]));
}
int32_t orc_xy_rect(rt_scene* s, double x0, double x1, double y0, double y1, double k, int32_t mat) {
    CHECK_SCENE(s); CHECK_MAT(s, mat);
    return add_obj(s, std::make_shared<AxisRect>(2, x0, x1, y0, y1, k, s->mats[(size_t)mat]));
}
int32_t orc_xz_rect(rt_scene* s, double x0, double x1, double z0, double z1, double k, int32_t mat) {
    CHECK_SCENE(s); CHECK_MAT(s, mat);
    return add_obj(s, std::make_shared<AxisRect>(1, x0, x1, z0, z1, k, s->mats[(size_t)mat]));
}
int32_t orc_yz_rect(rt_scene* s, double y0, double y1, double z0, double z1, double k, int32_t mat) {
    CHECK_SCENE(s); CHECK_MAT(s, mat);
    return add_obj(s, std::make_shared<AxisRect>(0, y0, y1, z0, z1, k, s->mats[(size_t)mat]));
}
int32_t orc_box(rt_scene* s, const double p0[3], const double p1[3], int32_t mat) {
    CHECK_SCENE(s); CHECK_MAT(s, mat);
    return add_obj(s, std::make_shared<RectPrism>(Point3(p0[0], p0[1], p0[2]), Point3(p1[0], p1[1], p1[2]), s->mats[(size_t)mat]));
}
int32_t orc_triangle(rt_scene* s, const double a[3], const double b[3], const double c[3], int32_t mat) {
    CHECK_SCENE(s); CHECK_MAT(s, mat);
    return add_obj(s, std::make_shared<Triangle>(Point3(a[0], a[1], a[2]), Point3(b[0], b[1], b[2]), Point3(c[0], c[1], c[2]), s->mats[(size_t)mat]));
}
int32_t orc_triangle_mesh(rt_scene* s, const double* verts, int64_t nv, const uint32_t* idx, int64_t nt, int32_t mat) {
    CHECK_SCENE(s); CHECK_MAT(s, mat);
    if (!verts || !idx || nv <= 0 || nt < 0) return fail(RT_ERR_INVALID, "bad mesh");
    auto list = std::make_shared<HittableList>(); // model.rs:64-76
    list->objects.reserve((size_t)nt);
    for (int64_t t = 0; t < nt; ++t) {
        const uint32_t a = idx[3 * t], b = idx[3 * t + 1], c = idx[3 * t + 2];
        if (a >= nv || b >= nv || c >= nv) return fail(RT_ERR_INVALID, "mesh index out of range");
        list->add(std::make_shared<Triangle>(Point3(verts[3 * a], verts[3 * a + 1], verts[3 * a + 2]), Point3(verts[3 * b], verts[3 * b + 1], verts[3 * b + 2]),
                                             Point3(verts[3 * c], verts[3 * c + 1], verts[3 * c + 2]), s->mats[(size_t)mat]));
    }
    return add_obj(s, list);
}
int32_t orc_ply_load(rt_scene* s, const char* path, double scale, int32_t mat) {
    CHECK_SCENE(s); CHECK_MAT(s, mat);
    std::vector<Point3> v;
    std::vector<uint32_t> f;
    std::string err;
    if (!path || !load_ply(path, scale, v, f, err)) return fail(RT_ERR_IO, err);
    auto list = std::make_shared<HittableList>();
    list->objects.reserve(f.size() / 3);
    for (size_t t = 0; t + 2 < f.size(); t += 3) list->add(std::make_shared<Triangle>(v[f[t]], v[f[t + 1]], v[f[t + 2]], s->mats[(size_t)mat]));
    return add_obj(s, list);
}
int32_t orc_list(rt_scene* s, const int32_t* ids, int32_t n) {
    CHECK_SCENE(s);
    if (n < 0 || (n > 0 && !ids)) return fail(RT_ERR_INVALID, "bad list");
    auto list = std::make_shared<HittableList>();
    for (int32_t i = 0; i < n; ++i) { CHECK_OBJ(s, ids[i]); list->add(s->objs[(size_t)ids[i]]); }
    return add_obj(s, list);
}
int32_t orc_bvh(rt_scene* s, const int32_t* ids, int32_t n, double t0, double t1) {
    CHECK_SCENE(s);
    if (n < 0 || (n > 0 && !ids)) return fail(RT_ERR_INVALID, "bad list");
    auto g = std::make_shared<BvhGroup>();
    // BvhNode::from_list takes list.get_objects(): a single HittableList argument contributes its
    // elements (world.rs:687 passes the TriangleModel list itself)
    for (int32_t i = 0; i < n; ++i) { CHECK_OBJ(s, ids[i]); g->original.push_back(s->objs[(size_t)ids[i]]); }
    if (g->original.size() == 1) {
        if (HittableList* inner = dynamic_cast<HittableList*>(g->original[0].get())) {
            std::vector<HittablePtr> flat = inner->objects;
            g->original.swap(flat);
        }
    }
    if (g->original.empty()) return fail(RT_ERR_EMPTY, "BvhNode over an empty list (the reference panics, bvh.rs:27-28)");
    std::vector<HittablePtr> work = g->original;
    SplitMix64 rng(s->bvh_seed + 0x9E37u * (uint64_t)s->objs.size());
    if (!BvhNode::build(work, 0, work.size(), t0, t1, rng, g->root)) return fail(RT_ERR_EMPTY, "No bounding box in bvh node constructor..");
    return add_obj(s, g);
}
int32_t orc_translate(rt_scene* s, const double off[3], int32_t child) {
    CHECK_SCENE(s); CHECK_OBJ(s, child);
    return add_obj(s, std::make_shared<Translate>(Vec3(off[0], off[1], off[2]), s->objs[(size_t)child]));
}
int32_t orc_rotate_y(rt_scene* s, double angle, int32_t child) {
    CHECK_SCENE(s); CHECK_OBJ(s, child);
    return add_obj(s, std::make_shared<RotateY>(angle, s->objs[(size_t)child]));
}
int32_t orc_constant_medium(rt_scene* s, const double rgb[3], double density, int32_t boundary) {
    CHECK_SCENE(s); CHECK_OBJ(s, boundary);
    // ConstantMedium::from_color builds its own Isotropic(SolidColor) (hit.rs:945-951); its texture
    // and material take the next texture / material ids, as in the product library.
    TexturePtr solid = std::make_shared<SolidColor>(Color(rgb[0], rgb[1], rgb[2]));
    add_tex(s, solid);
    MaterialPtr phase = std::make_shared<Isotropic>(solid);
    add_mat(s, phase);
    return add_obj(s, std::make_shared<ConstantMedium>(phase, density, s->objs[(size_t)boundary]));
}

int32_t orc_scene_set_root(rt_scene* s, int32_t id) {
    CHECK_SCENE(s); CHECK_OBJ(s, id);
    s->root = s->objs[(size_t)id];
    s->committed = false;
    return RT_OK;
}
int32_t orc_scene_set_camera(rt_scene* s, const double lf[3], const double la[3], const double vup[3], double vfov, double aspect, double aperture,
                             double focus, double t1, double t2) {
    CHECK_SCENE(s);
    s->cam = Camera(Point3(lf[0], lf[1], lf[2]), Point3(la[0], la[1], la[2]), Vec3(vup[0], vup[1], vup[2]), vfov, aspect, aperture, focus, t1, t2);
    s->has_cam = true;
    return RT_OK;
}
int32_t orc_scene_set_camera_fields(rt_scene* s, const double f[24]) {
    CHECK_SCENE(s);
    if (!f) return fail(RT_ERR_INVALID, "null camera fields");
    Camera c;
    Vec3* v[7] = {&c.origin, &c.lower_left_corner, &c.horizontal, &c.vertical, &c.u, &c.v, &c.w};
    for (int i = 0; i < 7; ++i) *v[i] = Vec3(f[3 * i], f[3 * i + 1], f[3 * i + 2]);
    c.lens_radius = f[21]; c.time1 = f[22]; c.time2 = f[23];
    s->cam = c;
    s->has_cam = true;
    return RT_OK;
}
int32_t orc_scene_set_background(rt_scene* s, const double rgb[3]) {
    CHECK_SCENE(s);
    s->background = Background();
    s->background.c0 = Color(rgb[0], rgb[1], rgb[2]);
    return RT_OK;
}
int32_t orc_scene_set_background_gradient(rt_scene* s, const double horizon_rgb[3], const double zenith_rgb[3]) {
    CHECK_SCENE(s);
    if (!horizon_rgb || !zenith_rgb) return fail(RT_ERR_INVALID, "null colour");
    s->background.c0 = Color(horizon_rgb[0], horizon_rgb[1], horizon_rgb[2]);
    s->background.c1 = Color(zenith_rgb[0], zenith_rgb[1], zenith_rgb[2]);
    s->background.gradient = true;
    return RT_OK;
}
int32_t orc_scene_commit(rt_scene* s) {
    CHECK_SCENE(s);
    if (!s->root) return fail(RT_ERR_STATE, "no root set");
    // a second commit keeps the numbering (ids are assigned once per object)
    for (HittablePtr& o : s->objs) (void)o;
    int32_t next = s->n_prims;
    s->root->number_leaves(next);
    s->n_prims = next;
    s->committed = true;
    return RT_OK;
}
int32_t orc_world_build(rt_scene* s, int32_t scene_id, uint64_t seed, int32_t param) {
    CHECK_SCENE(s);
    const int32_t r = rtb::build_world(s, scene_id, seed, param);
    return r;
}
int32_t orc_scene_num_prims(rt_scene* s) {
    CHECK_SCENE(s);
    return s->n_prims;
}

int32_t orc_image_height(const rt_render_config* cfg) {
    if (!cfg || cfg->image_width <= 0 || !(cfg->aspect_ratio > 0)) return RT_ERR_INVALID;
    return f64_as_i32((double)cfg->image_width / cfg->aspect_ratio); // world.rs:1192
}

// render_scene (world.rs:1181-1247)
int32_t orc_render(rt_scene* s, const rt_render_config* cfg, double* out_screen, int64_t* out_accum, rt_stats* stats) {
    CHECK_SCENE(s);
    if (!cfg) return fail(RT_ERR_INVALID, "null config");
    if (!s->committed) return fail(RT_ERR_STATE, "scene not committed");
    if (!s->has_cam) return fail(RT_ERR_STATE, "no camera");
    // Config::new asserts (world.rs:36-40)
    if (cfg->image_width <= 0 || cfg->samples_per_pixel <= 0 || cfg->max_depth <= 0) return fail(RT_ERR_INVALID, "Config assert");
    const int32_t W = cfg->image_width;
    const int32_t H = orc_image_height(cfg);
    if (H <= 0) return fail(RT_ERR_INVALID, "image height <= 0");
    const int32_t spp = cfg->samples_per_pixel;
    const int32_t s0 = cfg->sample_begin;
    const int32_t s1 = cfg->sample_end == 0 ? spp : cfg->sample_end;
    if (s0 < 0 || s1 > spp || s0 > s1) return fail(RT_ERR_INVALID, "bad sample range");
    int32_t rows = H;
    if (cfg->compat_threads > 0) rows = (H / cfg->compat_threads) * cfg->compat_threads; // world.rs:1198-1202
    int32_t threads = cfg->threads > 0 ? cfg->threads : (int32_t)std::thread::hardware_concurrency();
    if (threads < 1) threads = 1;
    const auto t_begin = std::chrono::steady_clock::now();

    std::vector<Color> sums((size_t)W * H, Color(0, 0, 0));
    std::vector<Counters> cnts((size_t)threads);
    const Hittable& world = *s->root;
    const Camera& cam = s->cam;
    const Background background = s->background;
    const int32_t max_depth = cfg->max_depth;
    const uint64_t seed = cfg->seed;
    const int32_t tile_count = RT_RENDER_TILE_COUNT(cfg->flags) > 1 ? RT_RENDER_TILE_COUNT(cfg->flags) : 1;
    const int32_t tile_rank = tile_count > 1 ? RT_RENDER_TILE_RANK(cfg->flags) : 0;
    if (tile_rank >= tile_count) return fail(RT_ERR_INVALID, "tile shard rank >= count");
    // static contiguous row bands, one per thread (world.rs:1198-1227)
    const int32_t chunk = (rows + threads - 1) / threads;
    std::vector<std::thread> pool;
    for (int32_t t = 0; t < threads; ++t) {
        const int32_t start = t * chunk;
        const int32_t end = std::min(start + chunk, rows);
        pool.emplace_back([&, t, start, end]() {
            PathCtx c;
            tls_ctx() = &c;
            for (int32_t j = start; j < end; ++j) {
                if (tile_count > 1 && (j / RT_TILE_ROWS) % tile_count != tile_rank) continue; // RT_RENDER_TILE_SHARD: not this shard's band
                for (int32_t i = 0; i < W; ++i) {
                    Vec3 pixel(0, 0, 0);
                    for (int32_t sidx = s0; sidx < s1; ++sidx) {
                        const uint64_t path_id = ((uint64_t)j * (uint64_t)W + (uint64_t)i) * (uint64_t)spp + (uint64_t)sidx;
                        c.begin_path(seed, path_id);
                        c.cnt.paths++;
                        const double u = ((double)i + c.gen()) / (double)(W - 1); // world.rs:1212
                        const double v = ((double)j + c.gen()) / (double)(H - 1); // world.rs:1213
                        const Ray r = cam.get_ray(u, v);
                        pixel += ray_color(r, background, world, max_depth);
                    }
                    sums[(size_t)j * W + i] = pixel;
                }
            }
            cnts[(size_t)t] = c.cnt;
            tls_ctx() = nullptr;
        });
    }
    for (std::thread& th : pool) th.join();

    if (out_screen) {
        for (int32_t j = 0; j < H; ++j)
            for (int32_t i = 0; i < W; ++i) {
                const size_t o = ((size_t)j * W + i) * 3;
                if (j < rows) {
                    const Color c = get_normalized_color(sums[(size_t)j * W + i], (uint32_t)spp); // world.rs:1221
                    out_screen[o] = c.x; out_screen[o + 1] = c.y; out_screen[o + 2] = c.z;
                } else {
                    out_screen[o] = out_screen[o + 1] = out_screen[o + 2] = 0.0; // Screen::new zeros (screen.rs:18)
                }
            }
    }
    if (out_accum) {
        const double scale = 4294967296.0;
        for (size_t p = 0; p < (size_t)W * H; ++p) {
            out_accum[3 * p] = (int64_t)std::llround(sums[p].x * scale);
            out_accum[3 * p + 1] = (int64_t)std::llround(sums[p].y * scale);
            out_accum[3 * p + 2] = (int64_t)std::llround(sums[p].z * scale);
        }
    }
    if (stats) {
        std::memset(stats, 0, sizeof *stats);
        Counters total;
        for (const Counters& c : cnts) total.add(c);
        stats->paths = total.paths; stats->segments = total.segments; stats->box_tests = total.box_tests;
        stats->medium_queries = total.medium_queries;
        for (int i = 0; i < 8; ++i) stats->prim_tests[i] = total.prim_tests[i];
        for (int i = 0; i < 5; ++i) stats->scatters[i] = total.scatters[i];
        stats->ms_total = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_begin).count();
    }
    return RT_OK;
}

// Screen::write_to_ppm / write_to_ppm_file (screen.rs:40-59); `{}` on an integer-valued f64 prints no decimal point
int32_t orc_write_ppm(const char* path, const double* screen, int32_t width, int32_t height) {
    if (!screen || width <= 0 || height <= 0) return fail(RT_ERR_INVALID, "bad screen");
    std::string out;
    out.reserve((size_t)width * height * 12 + 32);
    char buf[128];
    std::snprintf(buf, sizeof buf, "P3\n%d %d\n255\n", width, height);
    out += buf;
    auto fmt = [](double v, char* b, size_t n) {
        if (v == std::floor(v) && std::fabs(v) < 1e15) std::snprintf(b, n, "%lld", (long long)v);
        else std::snprintf(b, n, "%.17g", v);
    };
    for (int32_t j = height - 1; j >= 0; --j)
        for (int32_t i = 0; i < width; ++i) {
            const double* p = screen + ((size_t)j * width + i) * 3;
            char a[40], b[40], c[40];
            fmt(p[0], a, sizeof a); fmt(p[1], b, sizeof b); fmt(p[2], c, sizeof c);
            out += a; out += ' '; out += b; out += ' '; out += c; out += '\n';
        }
    FILE* f = path ? std::fopen(path, "wb") : stdout;
    if (!f) return fail(RT_ERR_IO, "cannot open output");
    std::fwrite(out.data(), 1, out.size(), f);
    if (path) std::fclose(f); else std::fflush(f);
    return RT_OK;
}

#define PFX(name) orc_##name
// render_scene_with_time (world.rs:1249-1330)
int32_t PFX(render_scene_with_time)(rt_scene* s, double t0, double t1, const char* path, const rt_render_config* cfg_in, double* out_screen, rt_stats* stats) {
    CHECK_SCENE(s);
    rt_render_config cfg;
    if (cfg_in) {
        cfg = *cfg_in;
    } else {
        std::memset(&cfg, 0, sizeof cfg);
        cfg.image_width = 500; cfg.aspect_ratio = 1.0; cfg.samples_per_pixel = 500; cfg.max_depth = 50; // world.rs:1253-1257
        cfg.compat_threads = 11;                                                                          // world.rs:18, 1281
        cfg.seed = 1;
    }
    const double lookfrom[3] = {13, 2, 3}, lookat[3] = {0, 0, 0}, vup[3] = {0, 1, 0}, bg[3] = {0.7, 0.8, 1.0}; // world.rs:1252, 1259-1263
    int32_t rc = PFX(scene_set_camera)(s, lookfrom, lookat, vup, 20.0, cfg.aspect_ratio, 0.1, 10.0, t0, t1);
    if (rc != RT_OK) return rc;
    if ((rc = PFX(scene_set_background)(s, bg)) != RT_OK) return rc;
    if ((rc = PFX(scene_commit)(s)) != RT_OK) return rc;
    const int32_t H = PFX(image_height)(&cfg);
    if (H <= 0) return fail(RT_ERR_INVALID, "image height <= 0");
    std::vector<double> local;
    double* screen = out_screen;
    if (!screen) { local.resize((size_t)cfg.image_width * H * 3); screen = local.data(); }
    if ((rc = PFX(render)(s, &cfg, screen, nullptr, stats)) != RT_OK) return rc;
    if (path) return PFX(write_ppm)(path, screen, cfg.image_width, H);
    return RT_OK;
}
#undef PFX

// world.hit(ray, t_min, t_max) per ray (world.rs:68)
int32_t orc_trace_batch(rt_scene* s, const rt_ray* rays, int64_t n, double t_min, double t_max, int32_t flags, uint64_t seed, rt_hit* out) {
    CHECK_SCENE(s);
    if (!s->committed) return fail(RT_ERR_STATE, "scene not committed");
    if (n < 0 || (n > 0 && (!rays || !out))) return fail(RT_ERR_INVALID, "bad batch");
    int32_t threads = (int32_t)std::thread::hardware_concurrency();
    if (threads < 1) threads = 1;
    if (n < 4096) threads = 1;
    const Hittable& world = *s->root;
    std::vector<std::thread> pool;
    const int64_t chunk = (n + threads - 1) / threads;
    for (int32_t t = 0; t < threads; ++t) {
        const int64_t a = t * chunk, b = std::min<int64_t>(n, a + chunk);
        pool.emplace_back([&, a, b]() {
            PathCtx c;
            tls_ctx() = &c;
            c.media_enabled = (flags & RT_TRACE_SEEDED_MEDIA) != 0;
            for (int64_t i = a; i < b; ++i) {
                c.begin_path(seed, (uint64_t)i);
                const rt_ray& rr = rays[i];
                const Ray r(Point3(rr.o[0], rr.o[1], rr.o[2]), Vec3(rr.d[0], rr.d[1], rr.d[2]), rr.time);
                HitRecord rec;
                rt_hit& h = out[i];
                std::memset(&h, 0, sizeof h);
                if (world.hit(r, t_min, t_max, rec)) {
                    h.prim_id = rec.prim_id;
                    h.mat_id = rec.mat_ptr ? rec.mat_ptr->mat_id : -1;
                    h.t = rec.t;
                    h.p[0] = rec.p.x; h.p[1] = rec.p.y; h.p[2] = rec.p.z;
                    h.normal[0] = rec.normal.x; h.normal[1] = rec.normal.y; h.normal[2] = rec.normal.z;
                    h.u = rec.u; h.v = rec.v;
                    h.front_face = rec.front_face ? 1 : 0;
                } else {
                    h.prim_id = -1;
                    h.mat_id = -1;
                }
            }
            tls_ctx() = nullptr;
        });
    }
    for (std::thread& th : pool) th.join();
    return RT_OK;
}

// ------------------------------------------------------------------ known-answer hooks (oracle only)
#define KAT __attribute__((visibility("default")))

KAT void orc_kat_philox(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]) { Philox::block(ctr, key, out); }

// op: 0 add, 1 sub, 2 mul(vec), 3 mul(scalar s), 4 div(scalar s), 5 cross, 6 neg, 7 unit, 8 add_assign, 9 mul_assign(s),
// 10 div_assign(s), 11 mul_assign(vec), 12 s*vec;  dot/length via orc_kat_vec3_scalar
KAT void orc_kat_vec3(int32_t op, const double a[3], const double b[3], double sc, double out[3]) {
    Vec3 x(a[0], a[1], a[2]), y(b[0], b[1], b[2]), r;
    switch (op) {
    case 0: r = x + y; break;
    case 1: r = x - y; break;
    case 2: r = x * y; break;
    case 3: r = x * sc; break;
    case 4: r = x / sc; break;
    case 5: r = x.cross(y); break;
    case 6: r = -x; break;
    case 7: r = x.unit(); break;
    case 8: r = x; r += y; break;
    case 9: r = x; r *= sc; break;
    case 10: r = x; r /= sc; break;
    case 11: r = x; r *= y; break;
    case 12: r = sc * x; break;
    default: break;
    }
    out[0] = r.x; out[1] = r.y; out[2] = r.z;
}
// op: 0 dot, 1 length, 2 length_squared, 3 near_zero
KAT double orc_kat_vec3_scalar(int32_t op, const double a[3], const double b[3]) {
    Vec3 x(a[0], a[1], a[2]), y(b[0], b[1], b[2]);
    switch (op) {
    case 0: return x.dot(y);
    case 1: return x.length();
    case 2: return x.length_squared();
    case 3: return x.near_zero() ? 1.0 : 0.0;
    default: return 0.0;
    }
}
KAT int32_t orc_kat_aabb_hit(const double mn[3], const double mx[3], const double o[3], const double d[3], double t_min, double t_max) {
    PathCtx c;
    tls_ctx() = &c;
    const Aabb box(Point3(mn[0], mn[1], mn[2]), Point3(mx[0], mx[1], mx[2]));
    const bool h = box.hit(Ray(Point3(o[0], o[1], o[2]), Vec3(d[0], d[1], d[2]), 0.0), t_min, t_max);
    tls_ctx() = nullptr;
    return h ? 1 : 0;
}
KAT double orc_kat_reflectance(double cosine, double ref_idx) { return Dielectric::reflectance(cosine, ref_idx); }
KAT void orc_kat_refract(const double uv[3], const double n[3], double ratio, double out[3]) {
    const Vec3 r = refract(Vec3(uv[0], uv[1], uv[2]), Vec3(n[0], n[1], n[2]), ratio);
    out[0] = r.x; out[1] = r.y; out[2] = r.z;
}
KAT void orc_kat_reflect(const double v[3], const double n[3], double out[3]) {
    const Vec3 r = Vec3(v[0], v[1], v[2]).reflect(Vec3(n[0], n[1], n[2]));
    out[0] = r.x; out[1] = r.y; out[2] = r.z;
}
KAT void orc_kat_normalized_color(const double sum[3], uint32_t spp, double out[3]) {
    const Color c = get_normalized_color(Color(sum[0], sum[1], sum[2]), spp);
    out[0] = c.x; out[1] = c.y; out[2] = c.z;
}
KAT void orc_kat_moving_center(const double c0[3], const double c1[3], double t0, double t1, double time, double out[3]) {
    MovingSphere m(Point3(c0[0], c0[1], c0[2]), Point3(c1[0], c1[1], c1[2]), t0, t1, 1.0, nullptr);
    const Point3 c = m.get_center(time);
    out[0] = c.x; out[1] = c.y; out[2] = c.z;
}
// camera fields: origin, lower_left_corner, horizontal, vertical, u, v, w (7 x 3), lens_radius, time1, time2
KAT int32_t orc_kat_camera(rt_scene* s, double out[24]) {
    CHECK_SCENE(s);
    if (!s->has_cam) return fail(RT_ERR_STATE, "no camera");
    const Camera& c = s->cam;
    const Vec3* v[7] = {&c.origin, &c.lower_left_corner, &c.horizontal, &c.vertical, &c.u, &c.v, &c.w};
    for (int i = 0; i < 7; ++i) { out[3 * i] = v[i]->x; out[3 * i + 1] = v[i]->y; out[3 * i + 2] = v[i]->z; }
    out[21] = c.lens_radius; out[22] = c.time1; out[23] = c.time2;
    return RT_OK;
}
// camera ray for path `path_id` exactly as the render loop draws it; also returns u, v
KAT int32_t orc_kat_camera_ray(rt_scene* s, uint64_t seed, uint64_t path_id, int32_t i, int32_t j, int32_t W, int32_t H, rt_ray* out) {
    CHECK_SCENE(s);
    PathCtx c;
    tls_ctx() = &c;
    c.begin_path(seed, path_id);
    const double u = ((double)i + c.gen()) / (double)(W - 1);
    const double v = ((double)j + c.gen()) / (double)(H - 1);
    const Ray r = s->cam.get_ray(u, v);
    tls_ctx() = nullptr;
    out->o[0] = r.origin.x; out->o[1] = r.origin.y; out->o[2] = r.origin.z;
    out->d[0] = r.direction.x; out->d[1] = r.direction.y; out->d[2] = r.direction.z;
    out->time = r.time;
    return RT_OK;
}
KAT int32_t orc_kat_texture_value(rt_scene* s, int32_t tex, double u, double v, const double p[3], double out[3]) {
    CHECK_SCENE(s); CHECK_TEX(s, tex);
    const Color c = s->tex[(size_t)tex]->value(u, v, Point3(p[0], p[1], p[2]));
    out[0] = c.x; out[1] = c.y; out[2] = c.z;
    return RT_OK;
}
KAT int32_t orc_kat_perlin(rt_scene* s, int32_t tex, const double p[3], double* noise, double* turb) {
    CHECK_SCENE(s); CHECK_TEX(s, tex);
    const Noise* n = dynamic_cast<const Noise*>(s->tex[(size_t)tex].get());
    if (!n) return fail(RT_ERR_INVALID, "not a noise texture");
    *noise = n->noise.noise(Point3(p[0], p[1], p[2]));
    *turb = n->noise.turbulence(Point3(p[0], p[1], p[2]), 7);
    return RT_OK;
}
KAT int32_t orc_kat_gravity_table(double start_y, double time0, double radius, int32_t n, double* out, int32_t* total) {
    GravitySphere g(Point3(0, start_y, 0), time0, radius, nullptr);
    *total = (int32_t)g.stored.size();
    for (int32_t i = 0; i < n && (size_t)i < g.stored.size(); ++i) out[i] = g.stored[(size_t)i];
    return RT_OK;
}
// samplers: kind 0 random_in_unit_sphere, 1 random_unit_vector, 2 random_in_unit_disk, 3 gen()
KAT void orc_kat_sampler(int32_t kind, uint64_t seed, uint64_t path_id, int32_t n, double* out3n) {
    PathCtx c;
    tls_ctx() = &c;
    c.begin_path(seed, path_id);
    for (int32_t i = 0; i < n; ++i) {
        Vec3 v;
        if (kind == 0) v = random_in_unit_sphere();
        else if (kind == 1) v = random_unit_vector();
        else if (kind == 2) v = random_in_unit_disk();
        else v = Vec3(c.gen(), 0, 0);
        out3n[3 * i] = v.x; out3n[3 * i + 1] = v.y; out3n[3 * i + 2] = v.z;
    }
    tls_ctx() = nullptr;
}
// material scatter on a synthetic hit record: returns 1 if scattered
KAT int32_t orc_kat_scatter(rt_scene* s, int32_t mat, uint64_t seed, uint64_t path_id, const rt_ray* r_in, const rt_hit* rec_in, rt_ray* scattered,
                            double attenuation[3], double emitted[3]) {
    CHECK_SCENE(s); CHECK_MAT(s, mat);
    PathCtx c;
    tls_ctx() = &c;
    c.begin_path(seed, path_id);
    HitRecord rec;
    rec.p = Point3(rec_in->p[0], rec_in->p[1], rec_in->p[2]);
    rec.normal = Vec3(rec_in->normal[0], rec_in->normal[1], rec_in->normal[2]);
    rec.t = rec_in->t; rec.u = rec_in->u; rec.v = rec_in->v; rec.front_face = rec_in->front_face != 0;
    const Ray rin(Point3(r_in->o[0], r_in->o[1], r_in->o[2]), Vec3(r_in->d[0], r_in->d[1], r_in->d[2]), r_in->time);
    Ray sc; Color att(0, 0, 0);
    const bool ok = s->mats[(size_t)mat]->scatter(rin, rec, sc, att);
    const Color em = s->mats[(size_t)mat]->emitted(rec.u, rec.v, rec.p);
    tls_ctx() = nullptr;
    scattered->o[0] = sc.origin.x; scattered->o[1] = sc.origin.y; scattered->o[2] = sc.origin.z;
    scattered->d[0] = sc.direction.x; scattered->d[1] = sc.direction.y; scattered->d[2] = sc.direction.z;
    scattered->time = sc.time;
    attenuation[0] = att.x; attenuation[1] = att.y; attenuation[2] = att.z;
    emitted[0] = em.x; emitted[1] = em.y; emitted[2] = em.z;
    return ok ? 1 : 0;
}
// radiance of single paths (linear, un-quantised): out = n x 3
KAT int32_t orc_kat_path_radiance(rt_scene* s, const rt_render_config* cfg, const uint64_t* path_ids, int32_t n, double* out3n) {
    CHECK_SCENE(s);
    if (!s->committed || !s->has_cam) return fail(RT_ERR_STATE, "scene not ready");
    const int32_t W = cfg->image_width, H = orc_image_height(cfg), spp = cfg->samples_per_pixel;
    PathCtx c;
    tls_ctx() = &c;
    for (int32_t k = 0; k < n; ++k) {
        const uint64_t pid = path_ids[k];
        const uint64_t pix = pid / (uint64_t)spp;
        const int32_t j = (int32_t)(pix / (uint64_t)W), i = (int32_t)(pix % (uint64_t)W);
        c.begin_path(cfg->seed, pid);
        const double u = ((double)i + c.gen()) / (double)(W - 1);
        const double v = ((double)j + c.gen()) / (double)(H - 1);
        const Ray r = s->cam.get_ray(u, v);
        const Color col = ray_color(r, s->background, *s->root, cfg->max_depth);
        out3n[3 * k] = col.x; out3n[3 * k + 1] = col.y; out3n[3 * k + 2] = col.z;
    }
    tls_ctx() = nullptr;
    return RT_OK;
}

} // extern "C"
