// cuda_shim.h — TEST INFRASTRUCTURE.  Lets g++ compile csrc/cuda/rt_device.cuh for the host, so that the device functions
// themselves (not a restatement of them) run on machines without a GPU against the host-flattened scene arrays
// (rt_debug_host_scene).  A "warp" here is one lane.  Nothing in the product includes this file.
#pragma once
#include <cuda_runtime.h> // vector types, make_float4; __device__ / __forceinline__ expand to host-compatible forms under g++

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>

#ifndef __noinline__
#define __noinline__ __attribute__((noinline))
#endif
using std::max;
using std::min;
template <class T> static inline T __ldg(const T* p) { return *p; }
static inline uint32_t __float_as_uint(float f) { uint32_t u; std::memcpy(&u, &f, 4); return u; }
static inline uint32_t __umulhi(uint32_t a, uint32_t b) { return (uint32_t)(((uint64_t)a * (uint64_t)b) >> 32); }
static inline int __popc(uint32_t a) { return __builtin_popcount(a); }
static inline uint32_t __ballot_sync(uint32_t, bool p) { return p ? 1u : 0u; }
static inline double __hiloint2double(int hi, int lo) { const uint64_t v = ((uint64_t)(uint32_t)hi << 32) | (uint32_t)lo; double d; std::memcpy(&d, &v, 8); return d; }
static inline double __longlong_as_double(long long v) { double d; std::memcpy(&d, &v, 8); return d; }
static inline float __double2float_ru(double v) { float f = (float)v; if ((double)f < v) f = std::nextafterf(f, INFINITY); return f; }
static inline float __double2float_rd(double v) { float f = (float)v; if ((double)f > v) f = std::nextafterf(f, -INFINITY); return f; }
