// rt_hd.h — small FP32 vector toolkit shared by every kernel.
//
// All functions are RT_HD (__host__ __device__) so that tests/emul/ can run the very same
// arithmetic on the CPU to debug logic without a GPU; the product only ever calls them from
// device code.  Restates the parts of Vector3D / Color the hot path uses
// (/root/reference/Serial/vector3D.cpp:32-113, color.cpp:19-43) in float.
#pragma once

#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

#if defined(__CUDACC__)
#define RT_HD __host__ __device__ __forceinline__
#else
#define RT_HD inline
#endif

struct f3 {
    float x, y, z;
};

RT_HD f3 mk3(float x, float y, float z) { f3 r; r.x = x; r.y = y; r.z = z; return r; }
RT_HD f3 mk3(float4 v) { return mk3(v.x, v.y, v.z); }
RT_HD f3 operator+(f3 a, f3 b) { return mk3(a.x + b.x, a.y + b.y, a.z + b.z); }
RT_HD f3 operator-(f3 a, f3 b) { return mk3(a.x - b.x, a.y - b.y, a.z - b.z); }
RT_HD f3 operator-(f3 a) { return mk3(-a.x, -a.y, -a.z); }
RT_HD f3 operator*(f3 a, float s) { return mk3(a.x * s, a.y * s, a.z * s); }
RT_HD f3 operator*(float s, f3 a) { return mk3(a.x * s, a.y * s, a.z * s); }
RT_HD f3 operator*(f3 a, f3 b) { return mk3(a.x * b.x, a.y * b.y, a.z * b.z); }
RT_HD float dot(f3 a, f3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
RT_HD f3 cross(f3 a, f3 b) { return mk3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x); }
RT_HD float length(f3 a) { return sqrtf(dot(a, a)); }
// Vector3D::normalize divides by the length (vector3D.cpp:94-95); keep the division so that a
// zero vector becomes NaN exactly like the reference (world.cpp:83 relies on it).
RT_HD f3 normalize(f3 a) {
    float l = length(a);
    return mk3(a.x / l, a.y / l, a.z / l);
}
RT_HD f3 fma3(f3 a, float s, f3 b) { return mk3(fmaf(a.x, s, b.x), fmaf(a.y, s, b.y), fmaf(a.z, s, b.z)); }

RT_HD float as_float(uint32_t u) {
#if defined(__CUDA_ARCH__)
    return __uint_as_float(u);
#else
    union { uint32_t u; float f; } c; c.u = u; return c.f;
#endif
}
RT_HD uint32_t as_uint(float f) {
#if defined(__CUDA_ARCH__)
    return __float_as_uint(f);
#else
    union { uint32_t u; float f; } c; c.f = f; return c.u;
#endif
}

template <typename T>
RT_HD T ldg(const T* p) {
#if defined(__CUDA_ARCH__)
    return __ldg(p);
#else
    return *p;
#endif
}

RT_HD int clz32(uint32_t x) {
#if defined(__CUDA_ARCH__)
    return __clz((int)x);
#else
    return x ? __builtin_clz(x) : 32;
#endif
}
RT_HD int clz64(uint64_t x) {
#if defined(__CUDA_ARCH__)
    return __clzll((long long)x);
#else
    return x ? __builtin_clzll(x) : 64;
#endif
}

// 1/x for hit distances and barycentrics: one MUFU.RCP-based approximate division on the device
// (<= 2 ulp, far inside the 1e-4 relative tolerance on t), IEEE division in the host emulation.
RT_HD float rt_rcp(float x) {
#if defined(__CUDA_ARCH__)
    return __fdividef(1.0f, x);
#else
    return 1.0f / x;
#endif
}

#define RT_FLT_MAX 3.402823466e+38f
