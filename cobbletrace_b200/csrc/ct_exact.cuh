// ct_exact.cuh -- device primitives that reproduce the reference's mixed fp64/fp32
// arithmetic bit for bit (SURVEY 0.2, 7 hard part #1).
//
// The reference computes vectors in double (v3_t, mymath.h:24-28), rounds every dot
// product / magnitude to float on return (mymath.h:229,213) and does its scalar work in
// float.  Every operation below is an explicit round-to-nearest intrinsic (__dmul_rn,
// __dadd_rn, __fmul_rn, ...): those are never contracted into FMAs, so the rounding points
// are the reference's whatever -fmad says.  B200 runs fp64 at half the fp32 rate, so exact
// evaluation is affordable; nothing here needs tensor cores.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

namespace ct {

#define CT_DEV __device__ __forceinline__

constexpr float kRayTInit = 1e30f;              // ray_t.t of primary/shadow rays (raythread.cpp:508,304)
constexpr float kFinf = 4294967296.0f;          // FINF raythread.cpp:58
constexpr float kEps = 0.0001f;                 // bvh.cpp:152,161

struct V3 { double x, y, z; };

CT_DEV V3 vsub(V3 a, V3 b) { return {__dsub_rn(a.x, b.x), __dsub_rn(a.y, b.y), __dsub_rn(a.z, b.z)}; }   // mymath.h:147
CT_DEV V3 vadd(V3 a, V3 b) { return {__dadd_rn(a.x, b.x), __dadd_rn(a.y, b.y), __dadd_rn(a.z, b.z)}; }   // mymath.h:129
CT_DEV V3 vscale(double s, V3 a) { return {__dmul_rn(s, a.x), __dmul_rn(s, a.y), __dmul_rn(s, a.z)}; }   // mymath.h:47
CT_DEV V3 vneg(V3 a) { return {-a.x, -a.y, -a.z}; }                                                      // mymath.h:118
CT_DEV V3 vcross(V3 a, V3 b) {                                                                           // mymath.h:234
    return {__dsub_rn(__dmul_rn(a.y, b.z), __dmul_rn(a.z, b.y)),
            __dsub_rn(__dmul_rn(a.z, b.x), __dmul_rn(a.x, b.z)),
            __dsub_rn(__dmul_rn(a.x, b.y), __dmul_rn(a.y, b.x))};
}
// DotProduct mymath.h:229: ((ax*bx + ay*by) + az*bz) in double, rounded to float on return.
CT_DEV float vdot(V3 a, V3 b) {
    return __double2float_rn(__dadd_rn(__dadd_rn(__dmul_rn(a.x, b.x), __dmul_rn(a.y, b.y)), __dmul_rn(a.z, b.z)));
}
// Magnitude mymath.h:213: double sqrt, rounded to float on return.
CT_DEV float vmag(V3 a) {
    return __double2float_rn(__dsqrt_rn(__dadd_rn(__dadd_rn(__dmul_rn(a.x, a.x), __dmul_rn(a.y, a.y)), __dmul_rn(a.z, a.z))));
}

// mymath.h:11-17 macros: a NaN operand makes the comparison false and yields the SECOND operand.
CT_DEV float macro_min(float a, float b) { return (a < b) ? a : b; }
CT_DEV float macro_max(float a, float b) { return (a > b) ? a : b; }

struct Ray {
    V3 o, d;        // ray_t scenefile.h:104-108
    V3 rd;          // 1/d per axis (correctly rounded), only used by box_times' guarded fast path
    float t;
    bool exact_div; // some |d| is so large/small that 1/d is inf/denormal: always divide
};

CT_DEV void ray_finish(Ray &r) {
    r.rd = {__ddiv_rn(1.0, r.d.x), __ddiv_rn(1.0, r.d.y), __ddiv_rn(1.0, r.d.z)};
    auto odd = [](double d) {
        uint32_t e = ((uint32_t)__double2hiint(d) >> 20) & 0x7ffu;   // biased exponent
        // zero/inf/NaN directions are fine (IEEE gives the same inf/NaN quotients both ways);
        // only finite non-zero d whose reciprocal leaves the normal range needs true division.
        bool zero = (((uint32_t)__double2hiint(d) & 0x7fffffffu) | (uint32_t)__double2loint(d)) == 0u;
        return !zero && e != 0x7ffu && (e < 24u || e > 2022u);
    };
    r.exact_div = odd(r.d.x) || odd(r.d.y) || odd(r.d.z);
}

// mymath.h:11-17 on doubles (same NaN behaviour as the float macros).
CT_DEV double macro_min_d(double a, double b) { return (a < b) ? a : b; }
CT_DEV double macro_max_d(double a, double b) { return (a > b) ? a : b; }

// Cold path: the literal IntersectAABB arithmetic of bvh.cpp:166-177 -- six fp64 divisions, each rounded to
// float on assignment, macro min/max in float.
struct BoxTimes { float tmin, tmax; };
__device__ __noinline__ BoxTimes box_times_exact(double ox, double oy, double oz, double dx, double dy, double dz,
                                                 double nx, double ny, double nz, double mx, double my, double mz) {
    float tx1 = __double2float_rn(__ddiv_rn(__dsub_rn(nx, ox), dx));
    float tx2 = __double2float_rn(__ddiv_rn(__dsub_rn(mx, ox), dx));
    float tmin = macro_min(tx1, tx2);
    float tmax = macro_max(tx1, tx2);
    float ty1 = __double2float_rn(__ddiv_rn(__dsub_rn(ny, oy), dy));
    float ty2 = __double2float_rn(__ddiv_rn(__dsub_rn(my, oy), dy));
    tmin = macro_max(tmin, macro_min(ty1, ty2));
    tmax = macro_min(tmax, macro_max(ty1, ty2));
    float tz1 = __double2float_rn(__ddiv_rn(__dsub_rn(nz, oz), dz));
    float tz2 = __double2float_rn(__ddiv_rn(__dsub_rn(mz, oz), dz));
    tmin = macro_max(tmin, macro_min(tz1, tz2));
    tmax = macro_min(tmax, macro_max(tz1, tz2));
    return {tmin, tmax};
}

// True when float(x) may differ from float(y) for some y within a few fp64 ulps of x: x sits within +-16 units of
// a float rounding boundary (the 29 dropped mantissa bits ~ 0x10000000), or x is non-zero with |x| < 2^-125 (float
// subnormal range, where the boundaries are elsewhere and the fp64 product may itself have lost bits).
CT_DEV bool float_rounding_unsafe(double x) {
    uint32_t lo = (uint32_t)__double2loint(x);
    uint32_t hi = (uint32_t)__double2hiint(x) & 0x7fffffffu;
    bool near_mid = ((lo & 0x1fffffffu) - 0x0ffffff0u) <= 0x20u;
    bool tiny = (hi < 0x38200000u) && ((hi | lo) != 0u);
    return near_mid | tiny;
}

// The two floats IntersectAABB (bvh.cpp:165-177) compares: tmin and tmax of the slab test, bit-exact (up to
// the sign of a zero, which no comparison sees).
//
// The reference rounds each of the six quotients (b - o)/d to float and then takes macro min/max in float.
// Rounding to float is monotonic and NaN-preserving, so selecting with the same macros on the unrounded
// doubles and rounding the two survivors gives the same float VALUES: whenever the double comparison picks
// a different operand than the float comparison would, the two operands round to the same float.
// The hot path also replaces the fp64 division by a multiplication with the correctly rounded 1/d: every
// quotient, and therefore (min/max being monotone selections) each survivor, is then within a few fp64 ulps
// of the reference's double.  float() of a survivor can differ from the reference only when it lies that
// close to a float rounding boundary (probability ~2^-24 per test); those cases, and rays whose 1/d leaves the
// normal range, take box_times_exact().  Everything else is proven equal.
CT_DEV void box_times(const Ray &r, const double bmin[3], const double bmax[3], float &tmin_f, float &tmax_f) {
    double x1 = __dmul_rn(__dsub_rn(bmin[0], r.o.x), r.rd.x), x2 = __dmul_rn(__dsub_rn(bmax[0], r.o.x), r.rd.x);
    double y1 = __dmul_rn(__dsub_rn(bmin[1], r.o.y), r.rd.y), y2 = __dmul_rn(__dsub_rn(bmax[1], r.o.y), r.rd.y);
    double z1 = __dmul_rn(__dsub_rn(bmin[2], r.o.z), r.rd.z), z2 = __dmul_rn(__dsub_rn(bmax[2], r.o.z), r.rd.z);
    double tmin = macro_min_d(x1, x2);
    double tmax = macro_max_d(x1, x2);
    tmin = macro_max_d(tmin, macro_min_d(y1, y2));
    tmax = macro_min_d(tmax, macro_max_d(y1, y2));
    tmin = macro_max_d(tmin, macro_min_d(z1, z2));
    tmax = macro_min_d(tmax, macro_max_d(z1, z2));
    if (float_rounding_unsafe(tmin) | float_rounding_unsafe(tmax) | r.exact_div) {
        BoxTimes e = box_times_exact(r.o.x, r.o.y, r.o.z, r.d.x, r.d.y, r.d.z, bmin[0], bmin[1], bmin[2], bmax[0], bmax[1], bmax[2]);
        tmin_f = e.tmin; tmax_f = e.tmax;
        return;
    }
    tmin_f = __double2float_rn(tmin);
    tmax_f = __double2float_rn(tmax);
}

// IntersectAABB bvh.cpp:165-179.
CT_DEV bool intersect_aabb(const Ray &r, const double bmin[3], const double bmax[3]) {
    float tmin, tmax;
    box_times(r, bmin, bmax, tmin, tmax);
    return tmax >= tmin && tmin < r.t && tmax > 0.0f;
}

// IntersectTriangle bvh.cpp:147-163 with edge1 = p2-p1, edge2 = p3-p1 precomputed at upload (the same
// two fp64 subtractions the reference redoes per call).  Returns the barycentric verdict; *t_out is
// the candidate distance (only meaningful when true).  The caller applies `if (t > 0.0001f) ray.t =
// min(ray.t, t)` -- the function returns true whatever the sign/size of t.
CT_DEV bool intersect_triangle(const Ray &r, V3 p1, V3 e1, V3 e2, float *t_out) {
    V3 h = vcross(r.d, e2);
    float a = vdot(e1, h);
    if (a > -kEps && a < kEps) return false;
    float f = __fdiv_rn(1.0f, a);
    V3 s = vsub(r.o, p1);
    float u = __fmul_rn(f, vdot(s, h));
    if (u < 0.0f || u > 1.0f) return false;
    V3 q = vcross(s, e1);
    float v = __fmul_rn(f, vdot(r.d, q));
    if (v < 0.0f || __fadd_rn(u, v) > 1.0f) return false;
    *t_out = __fmul_rn(f, vdot(e2, q));
    return true;
}

// ReflectRay raythread.cpp:270-273: 2.0*normal*Dot(normal,ray) - ray
CT_DEV V3 reflect_ray(V3 rv, V3 n) {
    double d = (double)vdot(n, rv);
    return vsub(vscale(d, vscale(2.0, n)), rv);
}

// ---- color.h ------------------------------------------------------------------------------------
CT_DEV uint32_t to_u8(float x) { return (uint32_t)__float2int_rz(x) & 0xffu; }   // float -> uint8_t argument conversion

// ColorToHsv (color.h:114) + `hsv.v = intensity` (raythread.cpp:365) + HsvToColor (color.h:122)
CT_DEV uint32_t shade_color(uint32_t color, float intensity) {
    float r = __fdiv_rn((float)(color & 0xffu), 255.0f);
    float g = __fdiv_rn((float)((color >> 8) & 0xffu), 255.0f);
    float b = __fdiv_rn((float)((color >> 16) & 0xffu), 255.0f);
    float max_c = macro_max(r, macro_max(g, b));
    float min_c = macro_min(r, macro_min(g, b));
    float delta = __fsub_rn(max_c, min_c);
    float s, h;
    if (max_c != 0.0f) {
        s = __fdiv_rn(delta, max_c);
        if (r == max_c) h = __fdiv_rn(__fsub_rn(g, b), delta);
        else if (g == max_c) h = __fadd_rn(2.0f, __fdiv_rn(__fsub_rn(b, r), delta));
        else h = __fadd_rn(4.0f, __fdiv_rn(__fsub_rn(r, g), delta));
        h = __double2float_rn(__dmul_rn((double)h, 60.0));
        if (h < 0.0f) h = __double2float_rn(__dadd_rn((double)h, 360.0));
    } else {
        s = 0.0f; h = -1.0f;
    }
    float v = intensity;
    float R, G, B;
    if (s == 0.0f) {
        R = G = B = v;
    } else {
        h = __double2float_rn(__ddiv_rn((double)h, 60.0));
        int i = (int)floor((double)h);
        float f = __fsub_rn(h, (float)i);
        float aa = __fmul_rn(v, __fsub_rn(1.0f, s));
        float bb = __fmul_rn(v, __fsub_rn(1.0f, __fmul_rn(s, f)));
        float cc = __fmul_rn(v, __fsub_rn(1.0f, __fmul_rn(s, __fsub_rn(1.0f, f))));
        R = G = B = 0.0f;
        switch (i) {
        case 0: R = v;  G = cc; B = aa; break;
        case 1: R = bb; G = v;  B = aa; break;
        case 2: R = aa; G = v;  B = cc; break;
        case 3: R = aa; G = bb; B = v;  break;
        case 4: R = cc; G = aa; B = v;  break;
        case 5: R = v;  G = aa; B = bb; break;
        }
    }
    uint32_t r8 = to_u8(macro_min(__fmul_rn(R, 255.0f), 255.0f));
    uint32_t g8 = to_u8(macro_min(__fmul_rn(G, 255.0f), 255.0f));
    uint32_t b8 = to_u8(macro_min(__fmul_rn(B, 255.0f), 255.0f));
    return (b8 << 16) | (g8 << 8) | r8;
}

// raythread.cpp:375-379: local*(1-r) + reflected*r per channel in double, truncated to uint8_t
CT_DEV uint32_t blend_color(uint32_t lc, uint32_t rc, float reflection) {
    double wl = (double)__fsub_rn(1.0f, reflection), wr = (double)reflection;
    uint32_t out = 0;
#pragma unroll
    for (int sh = 0; sh <= 16; sh += 8) {
        double l = (double)((lc >> sh) & 0xffu), r = (double)((rc >> sh) & 0xffu);
        double c = __dadd_rn(__dmul_rn(wl, l), __dmul_rn(wr, r));
        out |= ((uint32_t)__double2int_rz(c) & 0xffu) << sh;
    }
    return out;
}

}  // namespace ct
