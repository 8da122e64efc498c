/*
 * tests/hostsim/cuda_runtime.h -- TEST INFRASTRUCTURE, never shipped.
 *
 * Stand-in for <cuda_runtime.h> so that the per-ray device headers
 * (rayoptics_b200/csrc/rt_device.cuh, rt_lean.cuh) compile with g++ for the
 * host.  tests/test_hostsim.py runs the very same source expressions on CPU and
 * compares them bit for bit with the oracle: a check of the ALGEBRA of the exact
 * shortcuts (shared-reciprocal division, sqrt sequence, aperture band, ...)
 * that needs no GPU.  It is not a CPU implementation of the product: nothing in
 * rayoptics_b200/ can load it, and the MUFU seeds are emulated, not reproduced
 * (any sufficiently accurate seed must give the IEEE result where the sequences'
 * own fast-path tests pass -- that is the property under test).
 *
 * Compile with: g++ -O2 -ffp-contract=off -mfma -DRT_HOSTSIM -I tests/hostsim
 */
#pragma once
#include <cmath>
#include <cstdint>
#include <cstring>

#define __device__
#define __host__
#define __global__
#define __forceinline__ inline __attribute__((always_inline))
#define __noinline__ __attribute__((noinline))
#define __restrict__ __restrict
#define CUDART_INF (__builtin_inf())

static const struct { int x, y, z; } threadIdx = {0, 0, 0}, blockDim = {1, 1, 1};

static inline double __fma_rn(double a, double b, double c) { return std::fma(a, b, c); }
static inline double __dmul_rn(double a, double b) { return a*b; }
static inline long long __double_as_longlong(double a) { long long u; std::memcpy(&u, &a, 8); return u; }
static inline double __longlong_as_double(long long u) { double a; std::memcpy(&a, &u, 8); return a; }
static inline int __double2hiint(double a) { return (int)(__double_as_longlong(a) >> 32); }
static inline int __double2loint(double a) { return (int)(__double_as_longlong(a) & 0xffffffffLL); }
static inline double __hiloint2double(int hi, int lo)
{
    return __longlong_as_double((long long)(((unsigned long long)(unsigned)hi << 32) | (unsigned)lo));
}
static inline float __int_as_float(int i) { float f; std::memcpy(&f, &i, 4); return f; }

/* MUFU.RCP64H / MUFU.RSQ64H stand-ins: ~20 good bits in the high word, low word 0 */
static inline double hostsim_rcp64h(double b)
{
    double r = 1.0/b;
    return __hiloint2double(__double2hiint(r), 0);
}
static inline double hostsim_rsq64h(double x)
{
    double r = 1.0/std::sqrt(x);
    return __hiloint2double(__double2hiint(r), 0);
}

using std::sqrt;
using std::fabs;
using std::copysign;
using std::isnan;
using std::isinf;
