// rt_types.h — flattened scene layout shared by the host flattener and the device kernels.
//
// The reference's scene is a pointer graph of Arc<Box<dyn Hittable>> (src/hit.rs:82-85) walked by
// virtual calls.  Here it is flattened at rt_scene_commit into:
//   * typed primitive buffers in BVH-leaf order (SoA by type, "type tag" = which buffer),
//   * one array of 32-byte BVH nodes (sibling pairs adjacent => one 64 B fetch tests both children),
//   * a short list of instances = (transform chain, BVH root): every leaf is reached through
//     exactly one chain of Translate / RotateY wrappers (hit.rs:787-936); leaves that share a chain
//     share an instance, lists and BvhNodes merge into the enclosing instance (closest-hit is
//     topology independent, bvh.rs:97-112 vs hit.rs:660-690),
//   * media (ConstantMedium, hit.rs:938-990), each with its boundary as its own instance range,
//   * material / texture tables, Perlin tables, image texels, the camera block.
//
// Precision: rays, hit points, normals and primitive tests are f64 (B200's FP64 pipe runs at half
// the FP32 rate, so exactness against the f64 reference is affordable); BVH slab tests are f32 on
// outward-padded boxes (conservative); colours / throughput are f32; mesh triangles are stored f32.
#pragma once
#include <stdint.h>
#include <vector_types.h> // float4

namespace rtb {

#define RT_LEAF_FLAG 0x80000000u
#define RT_BVH_MAX_DEPTH 60 // traversal stack is 64 entries (rt_device.cuh RT_STACK)

// 4-wide BVH node (128 bytes = one cache line, eight float4), the binary tree collapsed by surface area
// (host/bvh_wide.hpp): q0 = lo.x of the four children, q1 = hi.x, q2 = lo.y, q3 = hi.y, q4 = lo.z, q5 = hi.z,
// q6 = four child references, q7 = the four binary child node indices (host bookkeeping, never read on the device).  A reference is the index of a 4-wide node (bit 31 clear), a leaf
// RT_LEAF_FLAG | type << 28 | (n - 1) << 25 | first (n <= 8 primitives of one type, first < 2^25), or
// RT_WIDE_EMPTY for an unused slot.  Half as many dependent fetches per ray as the sibling-pair walk.
#define RT_WIDE_EMPTY 0xffffffffu
#define RT_WIDE_STACK 96 // 64-bit entries (entry distance, reference); a visit pushes at most three
#define RT_WIDE_MAX_DEPTH 30

enum PrimType : uint32_t {
    PRIM_SPHERE = 0, PRIM_MOVING = 1, PRIM_GRAVITY = 2, PRIM_RECT = 3, PRIM_BOX = 4, PRIM_TRI = 5,
    PRIM_TYPE_COUNT = 6, PRIM_MEDIUM = 6
};
enum MatType : uint32_t { MAT_LAMBERTIAN = 0, MAT_METAL = 1, MAT_DIELECTRIC = 2, MAT_LIGHT = 3, MAT_ISOTROPIC = 4, MAT_TYPE_COUNT = 5 };
enum TexType : uint32_t { TEX_SOLID = 0, TEX_CHECKER = 1, TEX_NOISE = 2, TEX_IMAGE = 3 };
enum XformType : uint32_t { XF_TRANSLATE = 0, XF_ROTATE_Y = 1 };

// 32-byte BVH node.  Interior: count == 0, first = index of the left child (right = first + 1).
// Leaf: count = 0x80000000 | type << 24 | n: n primitives of one type, `first` = index into that
// type's buffer.  An instance root is the first node of a pair whose second node is an empty leaf.
struct alignas(32) BvhNode32 {
    float min[3];
    uint32_t first;
    float max[3];
    uint32_t count; // 0 = interior, else RT_LEAF_FLAG | (type << 24) | n
};

struct alignas(16) DSphere { double cx, cy, cz, r; };                                  // hit.rs:180-184
struct alignas(16) DMoving { double c0[3]; double dc[3]; double t0, dt, r, pad_; };   // hit.rs:247-254 (dc = center1 - center0, dt = time1 - time0)
struct alignas(8) DGravity { double x, z, r; int32_t table_off; int32_t idx0; int32_t n; int32_t pad_; }; // hit.rs:330-336, window of `stored`
struct alignas(16) DRect { double a0, a1, b0, b1, k; int32_t axis; int32_t pad_; };    // hit.rs:446-453 (axis 2 = Xy, 1 = Xz, 0 = Yz)
struct alignas(16) DBox { double p0[3]; double p1[3]; };                              // hit.rs:713-717
// hit.rs:87-93.  64 bytes = two whole sectors: f32 vertices and f32 unit normal n (hit.rs:96-107, computed in f64, then rounded), and the
// plane offset dd = -(n . v0) in f64, evaluated from the ROUNDED normal and the EXACT f64 vertex: the plane n . x + dd = 0 then passes
// through the reference's v0 exactly and only its tilt (6e-8 rad) differs, so t agrees to ~1e-7 relative for rays of any length (with
// dd recomputed from the f32 vertex the plane itself was displaced by the vertex rounding, ~1e-6 absolute: 2e-5 relative on t = 0.05).
// The f32 vertices serve the three edge tests only.
struct alignas(16) DTri { float v0[3], v1[3], v2[3], n[3]; double dd; double pad_; };

struct PrimMeta { uint32_t mat_id; uint32_t prim_id; }; // per primitive, same order as its typed buffer

struct XformOp { // one wrapper of the chain, outermost first
    uint32_t type;
    uint32_t pad_;
    double a, b, c; // translate: offset xyz; rotate-y: sin_theta, cos_theta, unused
};

struct Instance {
    uint32_t root;      // node index of this instance's BVH root
    uint32_t chain_off; // into ops[]
    uint32_t chain_len;
    uint32_t root4;     // reference of this instance's 4-wide root in DeviceScene::nodes4, RT_WIDE_EMPTY if it has none
    float bmin[3], bmax[3]; // padded bounds in the instance's own (innermost) space
};

struct Medium { // hit.rs:938-951
    uint32_t inst_begin, inst_end; // the boundary's instances
    uint32_t chain_off, chain_len; // wrappers around the ConstantMedium itself (usually 0)
    uint32_t mat_id, prim_id;
    double neg_inv_density;
    // fast path: the boundary is ONE sphere (1) or ONE box (2) reached through one wrapper chain
    uint32_t fast_type, fast_idx, fast_chain_off, fast_chain_len;
};

struct DMaterial {
    uint32_t type;
    uint32_t tex;     // albedo / emit texture id (lambertian, light, isotropic)
    uint32_t flags;   // bit 0: texture chain reads (u, v) -> the hit must carry sphere uv
    uint32_t pad_;
    float albedo[3];  // metal
    float pad2_;
    double fuzz_or_ir; // metal fuzz (clamped <= 1) / dielectric ir
};

struct DTexture {
    uint32_t type;
    uint32_t a, b;    // checker: even, odd texture ids; noise: perlin table index; image: texel offset
    uint32_t w, h;    // image size
    uint32_t pad_;
    float rgb[3];     // solid
    float pad2_;
    double scale;     // noise
};

struct PerlinTable { // perlin.rs:8-13 (gradients f64 as given; perms as bytes)
    double ranvec[256][3];
    uint8_t perm_x[256], perm_y[256], perm_z[256];
};

struct DCamera { // camera.rs:6-17
    double origin[3], lower_left_corner[3], horizontal[3], vertical[3], u[3], v[3], w[3];
    double lens_radius, time1, time2;
};

struct DeviceScene {
    const BvhNode32* nodes;
    const DSphere* spheres;
    const DMoving* movings;
    const DGravity* gravities;
    const double* gravity_table;
    const DRect* rects;
    const DBox* boxes;
    const DTri* tris;
    const PrimMeta* meta[PRIM_TYPE_COUNT];
    const Instance* instances;
    const XformOp* ops;
    const Medium* media;
    const DMaterial* materials;
    const DTexture* textures;
    const PerlinTable* perlin;
    const float4* texels; // rgba f32, row 0 = top of file
    // Motion-interpolated boxes for scenes with MovingSpheres (nullptr otherwise): node i = mnodes[2i] (box at the shutter's start,
    // same first / count as nodes[i]) and mnodes[2i + 1] (box at the shutter's end minus box at its start).  The box at ray time t
    // is box0 + s * delta with s = (t - motion_t0) * motion_inv_dt: exact for the reference's linear motion (hit.rs:275-278), so the
    // spheres-and-moving-spheres kernels walk tight boxes instead of the union over the shutter that `nodes` holds (hit.rs:317-327).
    const BvhNode32* mnodes;
    // 4-wide collapse of the main world's trees (nullptr unless built, see RT_WIDE_EMPTY above); per instance: Instance::root4
    const float4* nodes4;
    // motion form of nodes4 for MovingSphere scenes (nullptr otherwise): 16 float4 per wide node = the six box rows of the children at the
    // shutter's start, the references, one unused row, then the six rows of (box at the shutter's end - box at its start) and two unused rows
    const float4* mnodes4;
    double motion_t0, motion_inv_dt;
    uint32_t prim_mask; // bit t set: primitives of PrimType t exist
    uint32_t root4;     // = instances[0].root4, for the single-instance fused kernels (valid when nodes4 != nullptr)
    uint32_t n_main_instances;
    uint32_t n_media;
    uint32_t n_prims;
    uint32_t flags; // bit 0: the scene has axis rects / boxes (per-ray inverse directions are needed); bit 1: every medium has the fast path; bit 2: large triangle mesh; bit 3: Perlin-noise textures; bit 4: image textures; bit 5: Translate / RotateY wrappers present
    DCamera cam;
    float background[3];    // world.rs:86-89 constant background (gradient sky: colour at the horizon side, t = 0)
    uint32_t bg_gradient;   // 1: book-1 sky, (1 - t) * background + t * background_top, t = 0.5 * (unit(d).y + 1)  (rt_scene_set_background_gradient)
    float background_top[3];
    float pad2_;
};

} // namespace rtb
