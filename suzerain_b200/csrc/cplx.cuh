// cplx.cuh -- minimal FP64 complex arithmetic for device code.
#pragma once
#include <cuda_runtime.h>

namespace szb {

struct __align__(16) cplx {
    double x, y;
    __host__ __device__ cplx() {}
    __host__ __device__ constexpr cplx(double re, double im = 0.0) : x(re), y(im) {}
};

__host__ __device__ __forceinline__ cplx operator+(cplx a, cplx b) { return cplx(a.x + b.x, a.y + b.y); }
__host__ __device__ __forceinline__ cplx operator-(cplx a, cplx b) { return cplx(a.x - b.x, a.y - b.y); }
__host__ __device__ __forceinline__ cplx operator-(cplx a)         { return cplx(-a.x, -a.y); }
__host__ __device__ __forceinline__ cplx operator*(cplx a, cplx b) { return cplx(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x); }
__host__ __device__ __forceinline__ cplx operator*(cplx a, double s) { return cplx(a.x * s, a.y * s); }
__host__ __device__ __forceinline__ cplx operator*(double s, cplx a) { return cplx(a.x * s, a.y * s); }
__host__ __device__ __forceinline__ cplx &operator+=(cplx &a, cplx b) { a.x += b.x; a.y += b.y; return a; }
__host__ __device__ __forceinline__ cplx &operator-=(cplx &a, cplx b) { a.x -= b.x; a.y -= b.y; return a; }
__host__ __device__ __forceinline__ bool is_zero(cplx a) { return a.x == 0.0 && a.y == 0.0; }
// |re| + |im|: the izamax / zgbtf2 pivot magnitude
__host__ __device__ __forceinline__ double cabs1(cplx a) { return fabs(a.x) + fabs(a.y); }

// a -= b*c  (the LU rank-1 update kernel operation)
__device__ __forceinline__ void submul(cplx &a, cplx b, cplx c)
{
    a.x = fma(-b.x, c.x, a.x); a.x = fma( b.y, c.y, a.x);
    a.y = fma(-b.x, c.y, a.y); a.y = fma(-b.y, c.x, a.y);
}
// a += b*c
__device__ __forceinline__ void addmul(cplx &a, cplx b, cplx c)
{
    a.x = fma(b.x, c.x, a.x); a.x = fma(-b.y, c.y, a.x);
    a.y = fma(b.x, c.y, a.y); a.y = fma( b.y, c.x, a.y);
}
// a += c * s  (real s)
__device__ __forceinline__ void addmul(cplx &a, cplx c, double s)
{
    a.x = fma(c.x, s, a.x); a.y = fma(c.y, s, a.y);
}

// 1/z, robust (Smith's algorithm), as compilers implement ONE / z in zgbtf2
__host__ __device__ __forceinline__ cplx recip(cplx z)
{
    if (fabs(z.x) >= fabs(z.y)) {
        const double r = z.y / z.x, d = z.x + z.y * r;
        return cplx(1.0 / d, -r / d);
    } else {
        const double r = z.x / z.y, d = z.x * r + z.y;
        return cplx(r / d, -1.0 / d);
    }
}
// a / b (Smith)
__host__ __device__ __forceinline__ cplx cdiv(cplx a, cplx b)
{
    if (fabs(b.x) >= fabs(b.y)) {
        const double r = b.y / b.x, d = b.x + b.y * r;
        return cplx((a.x + a.y * r) / d, (a.y - a.x * r) / d);
    } else {
        const double r = b.x / b.y, d = b.x * r + b.y;
        return cplx((a.x * r + a.y) / d, (a.y * r - a.x) / d);
    }
}

}  // namespace szb
