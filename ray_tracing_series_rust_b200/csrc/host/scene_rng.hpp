// scene_rng.hpp — deterministic host-side randomness for SCENE CONSTRUCTION.
//
// The reference draws scene randomness (sphere placement, box heights, Perlin tables) from an
// unseeded rand::thread_rng() (src/world.rs:96,171,192,510; src/perlin.rs:15,78).  Here the same
// draws come from a seeded SplitMix64 stream so that the GPU library and the CPU oracle build
// bit-identical scenes from one seed.  Render-time randomness is Philox (csrc/cuda/philox.cuh).
#pragma once
#include <cstdint>
#include <vector>

namespace rtb {

struct SceneRng {
    uint64_t s;
    explicit SceneRng(uint64_t seed) : s(seed) {}
    uint64_t next_u64() {
        uint64_t z = (s += 0x9E3779B97F4A7C15ull);
        z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
        z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
        return z ^ (z >> 31);
    }
    // rng.gen::<f64>(): 53-bit uniform [0,1)
    double gen() { return (double)(next_u64() >> 11) * (1.0 / 9007199254740992.0); }
    // rng.gen_range(a..b) for f64
    double gen_range(double a, double b) { return a + gen() * (b - a); }
    // rng.gen_range(0..n) for integers
    uint64_t below(uint64_t n) { return next_u64() % n; }
};

// Perlin::new() (src/perlin.rs:14-26) + perlin_generate_perm/permute (src/perlin.rs:68-83):
// 256 gradients U[-1,1)^3 (NOT normalised), three identity permutations shuffled with
// `for i in (1..len-1).rev() { swap(i, gen_range(0..i+1)) }`, i.e. i = 254..1, so perm[255]==255.
struct PerlinTables {
    double ranvec[768];
    int32_t perm_x[256], perm_y[256], perm_z[256];
};

inline void perlin_permute(SceneRng& rng, int32_t* p) {
    for (int i = 0; i < 256; ++i) p[i] = i;
    for (int i = 254; i >= 1; --i) {
        const int target = (int)rng.below((uint64_t)i + 1);
        const int32_t tmp = p[i];
        p[i] = p[target];
        p[target] = tmp;
    }
}

inline void perlin_generate(uint64_t seed, PerlinTables& t) {
    SceneRng rng(seed ^ 0x5045524C494E0000ull);
    for (int i = 0; i < 768; ++i) t.ranvec[i] = rng.gen_range(-1.0, 1.0);
    perlin_permute(rng, t.perm_x);
    perlin_permute(rng, t.perm_y);
    perlin_permute(rng, t.perm_z);
}

} // namespace rtb
