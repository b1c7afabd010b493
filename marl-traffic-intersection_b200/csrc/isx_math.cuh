// isx_math.cuh — float math with the SAME BITS as the libm the reference binds.
//
// The reference sim core imports exactly sincosf, tanf, atan2f, hypotf, fmodf (+ expf, host only)
// from glibc (SURVEY.md §7 hard part 1).  Collision flags and lidar hit indices are knife-edge
// functions of poses, so "within 1 ulp" device intrinsics are not good enough for bit-exact parity.
// These are restatements of the PUBLISHED algorithms glibc 2.39 ships for those entry points:
//
//   sincosf : ARM Optimized Routines sincosf (Szabolcs Nagy, 2018; glibc sysdeps/ieee754/flt-32/
//             s_sincosf.{c,h}, s_sincosf_data.c): double-precision range reduction by pi/2 and two
//             degree-(7,8) polynomials on [-pi/4, pi/4].  glibc dispatches to an FMA build on every
//             FMA-capable x86-64 CPU; the operation pairing below (which products are fused) is that
//             build's, so results agree bit-for-bit with it.
//   tanf    : fdlibm k_tanf.c (Sun, 1993; float port by Ian Taylor) + the sincosf reduction.
//   atan2f  : fdlibm e_atan2f.c / s_atanf.c.
//   hypotf  : glibc >= 2.35 e_hypotf.c: sqrt in double of the exact squares.
//   fmodf   : exact by definition (CUDA fmodf has 0 ulp error); used as is.
//
// Every function is __host__ __device__: the host build is swept against this machine's libm in
// tests/test_math_host.py (exhaustively for the 1-argument functions), the device build is compared
// with the host build in tests/test_gpu_math.py.  REQUIRES float/double contraction OFF
// (nvcc -fmad=false, g++ -ffp-contract=off): fused operations are written explicitly with fma().
#pragma once
#include <stdint.h>
#include <math.h>
#include <string.h>

#if defined(__CUDACC__)
#define ISX_HD __host__ __device__ __forceinline__
#define ISX_HDM __host__ __device__ __forceinline__     // member functions
// Big functions that are called from many places of the dynamics kernel: one out-of-line copy per translation unit,
// otherwise k_dynamics inlines them into >100 KB of SASS and stalls on instruction fetch (ncu: stall_no_instruction).
#define ISX_HD_NOINL static __host__ __device__ __noinline__
#else
#define ISX_HD static inline
#define ISX_HDM inline
#define ISX_HD_NOINL static inline
#endif

namespace isx {

constexpr float PI_F = 3.14159265358979323846f;       // 0x40490fdb, as Car.cpp:7
constexpr float TWO_PI_F = 2.0f * PI_F;

ISX_HD uint32_t f2u(float f) {
#if defined(__CUDA_ARCH__)
    return __float_as_uint(f);
#else
    uint32_t u; memcpy(&u, &f, 4); return u;
#endif
}
ISX_HD float u2f(uint32_t u) {
#if defined(__CUDA_ARCH__)
    return __uint_as_float(u);
#else
    float f; memcpy(&f, &u, 4); return f;
#endif
}
ISX_HD int f2i_rz(float f) {      // C's (int)f for in-range values
#if defined(__CUDA_ARCH__)
    return __float2int_rz(f);
#else
    return (int)f;
#endif
}
ISX_HD int d2i_rz(double d) {
#if defined(__CUDA_ARCH__)
    return __double2int_rz(d);
#else
    return (int)d;
#endif
}
ISX_HD float fsqrt_rn(float x) {
#if defined(__CUDA_ARCH__)
    return __fsqrt_rn(x);
#else
    return sqrtf(x);
#endif
}
ISX_HD float fdiv_rn(float a, float b) {
#if defined(__CUDA_ARCH__)
    return __fdiv_rn(a, b);
#else
    return a / b;
#endif
}
ISX_HD double dsqrt_rn(double x) {
#if defined(__CUDA_ARCH__)
    return __dsqrt_rn(x);
#else
    return sqrt(x);
#endif
}

// ---------------------------------------------------------------- sincosf
// 4/pi as 24 overlapping 32-bit words (192 bits), for |x| >= 120.  Table-free (the words are windows of one bit
// string) so that the rare large-argument path costs no stack frame in the kernels.
ISX_HD uint32_t inv_pio4_word(int i) {
    // bit string of 4/pi, most significant first, as 27 bytes: a2 f9 83 6e 4e 44 15 29 fc 27 57 d1 f5 34 dd c0 db 62 95 99 3c 43 90 41 ..
    // word i = bytes [i-3 .. i] of that string (zero-padded on the left)
    const uint64_t b0 = 0xa2f9836e4e441529ull, b1 = 0xfc2757d1f534ddc0ull, b2 = 0xdb6295993c439041ull;
    // extract 32 bits ending at byte index i (0-based) of the 24-byte string b0|b1|b2
    const int end_bit = 8 * (i + 1);          // number of leading bits included
    // value = (string >> (192 - end_bit)) & 0xffffffff, with the string zero-extended on the left
    const int sh = 192 - end_bit;             // 0..184
    uint64_t lo, hi;                          // 128-bit window [hi:lo] = string >> sh (low 64 bits suffice)
    if (sh >= 128) { lo = b0 >> (sh - 128); }
    else if (sh >= 64) { const int t = sh - 64; lo = t ? ((b1 >> t) | (b0 << (64 - t))) : b1; }
    else { lo = sh ? ((b2 >> sh) | (b1 << (64 - sh))) : b2; }
    (void)hi;
    return (uint32_t)lo;
}

// Reduce y to xr in [-pi/4, pi/4] and quadrant n; returns false for inf/nan.  `sgn_extra` is the
// sign bit that the large-argument path folds into the quadrant (0 elsewhere).
ISX_HD bool sincos_reduce(float y, double* xr, int* n, int* sgn_extra) {
    const double HPI_INV = 0x1.45F306DC9C883p+23;   // 2/pi * 2^24
    const double HPI = 0x1.921FB54442D18p0;
    const uint32_t iy = f2u(y);
    const uint32_t top = (iy >> 20) & 0x7ffu;
    const double x = (double)y;
    *sgn_extra = 0;
    if (top < 0x3f4u) {                 // |y| < 0.75 (top-12-bit compare with pi/4)
        *xr = x; *n = 0;
        return true;
    }
    if (top < 0x42fu) {                 // |y| < 120
        const double r = x * HPI_INV;
        const int q = (d2i_rz(r) + 0x800000) >> 24;
        *xr = fma(-(double)q, HPI, x);
        *n = q;
        return true;
    }
    if (top < 0x7f8u) {                 // finite: 192-bit 4/pi table
        const int idx = (int)((iy >> 26) & 15u);
        const int shift = (int)((iy >> 23) & 7u);
        uint32_t xi = (iy & 0xffffffu) | 0x800000u;
        xi <<= shift;
        uint64_t res0 = (uint64_t)(uint32_t)(xi * inv_pio4_word(idx));
        const uint64_t res1 = (uint64_t)xi * inv_pio4_word(idx + 4);
        const uint64_t res2 = (uint64_t)xi * inv_pio4_word(idx + 8);
        res0 = (res2 >> 32) | (res0 << 32);
        res0 += res1;
        const uint64_t q = (res0 + (1ULL << 61)) >> 62;
        res0 -= q << 62;
        *xr = (double)(int64_t)res0 * 0x1.921FB54442D18p-62;
        *n = (int)q;
        *sgn_extra = (int)(iy >> 31);
        return true;
    }
    return false;
}

ISX_HD void sincosf_(float y, float* sinp, float* cosp) {
    const double C0 = 0x1p0, C1 = -0x1.ffffffd0c621cp-2, C2 = 0x1.55553e1068f19p-5,
                 C3 = -0x1.6c087e89a359dp-10, C4 = 0x1.99343027bf8c3p-16;
    const double S1 = -0x1.555545995a603p-3, S2 = 0x1.1107605230bc4p-7, S3 = -0x1.994eb3774cf24p-13;
    const uint32_t top = (f2u(y) >> 20) & 0x7ffu;
    double xr;
    int n, q;
    if (top < 0x42fu) {
        // |y| < 120.  glibc branches on |y| < 0.75 first and skips the reduction there; the reduction below
        // yields n == 0 and xr == y exactly for every |y| < pi/4, so one straight-line path serves both
        // (keeps the lanes of a warp converged: beam angles span [-2pi, 2pi]).
        const double x = (double)y;
        n = (d2i_rz(x * 0x1.45F306DC9C883p+23) + 0x800000) >> 24;
        xr = fma(-(double)n, 0x1.921FB54442D18p0, x);
        q = n;
    } else {
        int se;
        if (!sincos_reduce(y, &xr, &n, &se)) {
            const float nan = y - y;
            *sinp = nan; *cosp = nan;
            return;
        }
        q = n + se;
    }
    // sign table {1,-1,-1,1}[q&3]; the second coefficient set (q&2) is the first with c0..c4 negated
    const double sg = (((q + 1) & 2) != 0) ? -1.0 : 1.0;
    const double xs = xr * sg;
    const double x2 = xr * xr;
    const double x3 = x2 * xs;
    const double x4 = x2 * x2;
    const double s1v = fma(x2, S3, S2);
    const double c2v = fma(x2, C4, C3);
    const double x5 = x2 * x3;
    const double x6 = x2 * x4;
    const double c1v = fma(x2, C1, C0);
    const double sv = fma(x3, S1, xs);
    const double cv = fma(x4, C2, c1v);
    float fs = (float)fma(s1v, x5, sv);
    float fc = (float)fma(c2v, x6, cv);
    if (q & 2) fc = -fc;
    if (n & 1) { const float t = fs; fs = fc; fc = t; }
    if (top < 0x398u) { fs = y; fc = 1.0f; }      // |y| < 2^-12: glibc returns (y, 1) without evaluating anything
    *sinp = fs; *cosp = fc;
}

// ---------------------------------------------------------------- tanf
ISX_HD float kernel_tanf(float x, float y, int iy) {
    const float T0 = u2f(0x3eaaaaabu), T1 = u2f(0x3e088889u), T2 = u2f(0x3d5d0dd1u), T3 = u2f(0x3cb327a4u),
                T4 = u2f(0x3c11371fu), T5 = u2f(0x3b6b6916u), T6 = u2f(0x3abede48u), T7 = u2f(0x3a1a26c8u),
                T8 = u2f(0x398137b9u), T9 = u2f(0x38a3f445u), T10 = u2f(0x3895c07au), T11 = u2f(0xb79bae5fu),
                T12 = u2f(0x37d95384u);
    const float pio4 = u2f(0x3f490fdau), pio4lo = u2f(0x33222168u);
    const uint32_t hx = f2u(x);
    const uint32_t ix = hx & 0x7fffffffu;
    if (ix < 0x39000000u) {                 // |x| < 2^-13
        if (f2i_rz(x) == 0) {
            if ((ix | (uint32_t)(iy + 1)) == 0u) return fdiv_rn(1.0f, fabsf(x));
            else if (iy == 1) return x;
            else return fdiv_rn(-1.0f, x);
        }
    }
    const bool big = ix >= 0x3f2ca140u;     // |x| >= 0.6744
    if (big) {
        if ((int32_t)hx < 0) { x = -x; y = -y; }
        const float z0 = pio4 - x;
        const float w0 = pio4lo - y;
        x = z0 + w0; y = 0.0f;
        if (fabsf(x) < 0x1p-13f) {
            const int sgn = 1 - (int)((hx >> 30) & 2u);
            return (float)(sgn * iy) * (1.0f - (float)(2 * iy) * x);
        }
    }
    const float z = x * x;
    const float w = z * z;
    const float r0 = T1 + w * (T3 + w * (T5 + w * (T7 + w * (T9 + w * T11))));
    const float v0 = z * (T2 + w * (T4 + w * (T6 + w * (T8 + w * (T10 + w * T12)))));
    const float s = z * x;
    float r = y + z * (s * (r0 + v0) + y);
    r += T0 * s;
    const float ww = x + r;
    if (big) {
        const float v = (float)iy;
        const int sgn = 1 - (int)((hx >> 30) & 2u);
        return (float)sgn * (v - 2.0f * (x - (fdiv_rn(ww * ww, ww + v) - r)));
    }
    if (iy == 1) return ww;
    // -1/(x+r) with extra care
    const float zz = u2f(f2u(ww) & 0xfffff000u);
    const float vv = r - (zz - x);
    const float a = fdiv_rn(-1.0f, ww);
    const float t = u2f(f2u(a) & 0xfffff000u);
    const float ss = 1.0f + t * zz;
    return t + a * (ss + t * vv);
}

ISX_HD float tanf_(float x) {
    const uint32_t ix = f2u(x) & 0x7fffffffu;
    if (ix <= 0x3f490fdau) return kernel_tanf(x, 0.0f, 1);    // |x| <= pi/4
    if (ix >= 0x7f800000u) return x - x;                       // inf/nan
    double xr; int n, se;
    sincos_reduce(x, &xr, &n, &se);       // top>=0x3f4 here, so always the reducing branches
    if (se) xr = -xr;                     // large path reduces |x|
    // the medium branch of glibc's tanf does NOT fuse x - n*hpi (baseline build): redo it unfused
    if (((f2u(x) >> 20) & 0x7ffu) < 0x42fu) {
        const double HPI = 0x1.921FB54442D18p0;
        xr = (double)x - (double)n * HPI;
    }
    const float y0 = (float)xr;
    const float y1 = (float)(xr - (double)y0);
    return kernel_tanf(y0, y1, 1 - ((n & 1) << 1));
}

// ---------------------------------------------------------------- atanf / atan2f
ISX_HD float atanf_(float x) {
    const float aT0 = u2f(0x3eaaaaabu), aT1 = u2f(0xbe4ccccdu), aT2 = u2f(0x3e124925u), aT3 = u2f(0xbde38e38u),
                aT4 = u2f(0x3dba2e6eu), aT5 = u2f(0xbd9d8795u), aT6 = u2f(0x3d886b35u), aT7 = u2f(0xbd6ef16bu),
                aT8 = u2f(0x3d4bda59u), aT9 = u2f(0xbd15a221u), aT10 = u2f(0x3c8569d7u);
    const uint32_t hx = f2u(x);
    const uint32_t ix = hx & 0x7fffffffu;
    float hi = 0.0f, lo = 0.0f;
    int id;
    if (ix >= 0x4c000000u) {               // |x| >= 2^25
        if (ix > 0x7f800000u) return x + x;
        const float r = u2f(0x3fc90fdau) + u2f(0x33a22168u);
        return ((int32_t)hx > 0) ? r : -r;
    }
    if (ix < 0x3ee00000u) {                // |x| < 0.4375
        if (ix < 0x31000000u) return x;    // |x| < 2^-29
        id = -1;
    } else {
        x = fabsf(x);
        if (ix < 0x3f980000u) {            // |x| < 1.1875
            if (ix < 0x3f300000u) { id = 0; x = fdiv_rn(2.0f * x - 1.0f, 2.0f + x); hi = u2f(0x3eed6338u); lo = u2f(0x31ac3769u); }
            else                  { id = 1; x = fdiv_rn(x - 1.0f, x + 1.0f);        hi = u2f(0x3f490fdau); lo = u2f(0x33222168u); }
        } else {
            if (ix < 0x401c0000u) { id = 2; x = fdiv_rn(x - 1.5f, 1.0f + 1.5f * x); hi = u2f(0x3f7b985eu); lo = u2f(0x33140fb4u); }
            else                  { id = 3; x = fdiv_rn(-1.0f, x);                  hi = u2f(0x3fc90fdau); lo = u2f(0x33a22168u); }
        }
    }
    const float z = x * x;
    const float w = z * z;
    const float s1 = z * (aT0 + w * (aT2 + w * (aT4 + w * (aT6 + w * (aT8 + w * aT10)))));
    const float s2 = w * (aT1 + w * (aT3 + w * (aT5 + w * (aT7 + w * aT9))));
    if (id < 0) return x - x * (s1 + s2);
    const float zz = hi - ((x * (s1 + s2) - lo) - x);
    return ((int32_t)hx < 0) ? -zz : zz;
}

ISX_HD float atan2f_(float y, float x) {
    const float tiny = u2f(0x0da24260u);   // 1.0e-30
    const float pi_o_4 = u2f(0x3f490fdbu), pi_o_2 = u2f(0x3fc90fdbu), pi = u2f(0x40490fdbu), pi_lo = u2f(0xb3bbbd2eu);
    const uint32_t hx = f2u(x), hy = f2u(y);
    const uint32_t ix = hx & 0x7fffffffu, iy = hy & 0x7fffffffu;
    if (ix > 0x7f800000u || iy > 0x7f800000u) return x + y;
    if (hx == 0x3f800000u) return atanf_(y);
    const int m = (int)((hy >> 31) & 1u) | (int)((hx >> 30) & 2u);
    if (iy == 0u) {
        switch (m) {
            case 0: case 1: return y;
            case 2: return pi + tiny;
            default: return -pi - tiny;
        }
    }
    if (ix == 0u) return ((int32_t)hy < 0) ? -pi_o_2 - tiny : pi_o_2 + tiny;
    if (ix == 0x7f800000u) {
        if (iy == 0x7f800000u) {
            switch (m) {
                case 0: return pi_o_4 + tiny;
                case 1: return -pi_o_4 - tiny;
                case 2: return 3.0f * pi_o_4 + tiny;
                default: return -3.0f * pi_o_4 - tiny;
            }
        } else {
            switch (m) {
                case 0: return 0.0f;
                case 1: return -0.0f;
                case 2: return pi + tiny;
                default: return -pi - tiny;
            }
        }
    }
    if (iy == 0x7f800000u) return ((int32_t)hy < 0) ? -pi_o_2 - tiny : pi_o_2 + tiny;
    const int k = ((int)iy - (int)ix) >> 23;
    float z;
    if (k > 60) z = pi_o_2 + 0.5f * pi_lo;
    else if ((int32_t)hx < 0 && k < -60) z = 0.0f;
    else z = atanf_(fabsf(fdiv_rn(y, x)));
    switch (m) {
        case 0: return z;
        case 1: return u2f(f2u(z) ^ 0x80000000u);
        case 2: return pi - (z - pi_lo);
        default: return (z - pi_lo) - pi;
    }
}

// ---------------------------------------------------------------- hypotf
ISX_HD float hypotf_(float x, float y) {
    const uint32_t ax = f2u(x) & 0x7fffffffu, ay = f2u(y) & 0x7fffffffu;
    if (ax >= 0x7f800000u || ay >= 0x7f800000u) {
        if (ax == 0x7f800000u || ay == 0x7f800000u) return u2f(0x7f800000u);
        return x + y;
    }
    const double dx = (double)x, dy = (double)y;
    return (float)dsqrt_rn(dx * dx + dy * dy);
}

// Historic out-of-line entry points (same arithmetic).  Round 2 measured them inline again: with the kernels leaner, the calls
// cost more than the code size (k_traffic 94 -> 91 us, k_ego / k_features unchanged); the names stay for the call sites.
ISX_HD void sincosf_nc(float y, float* sinp, float* cosp) { sincosf_(y, sinp, cosp); }
ISX_HD float tanf_nc(float x) { return tanf_(x); }
ISX_HD float atan2f_nc(float y, float x) { return atan2f_(y, x); }
ISX_HD float hypotf_nc(float x, float y) { return hypotf_(x, y); }

// ---------------------------------------------------------------- helpers used all over the sim
// fmodf is exact by definition, so any exact evaluation has libm's bits.  Every call site divides an angle by 2*pi
// and the quotient is almost always 0 or +-1: |a| < b returns a; b <= |a| < 2b returns a -+ b, which is exact by
// Sterbenz's lemma (b <= |a| <= 2b).  Anything else takes the generic routine.  (CUDA's generic fmodf is a ~100-cycle
// dependent chain and sat on the critical path of every small kernel.)
ISX_HD float fmodf_(float a, float b) {
    const float aa = fabsf(a), ab = fabsf(b);
    if (aa < ab) return a;
    if (aa < 2.0f * ab) {
        const float r = (a < 0.0f) ? a + ab : a - ab;
        return (r == 0.0f) ? copysignf(0.0f, a) : r;          // fmod's zero carries the sign of a
    }
    return fmodf(a, b);
}

// wrap to [-pi, pi): IntersectionEnv.cpp:9-13, TrafficFlow.cpp:8-12, Car.cpp:33-36
ISX_HD float wrap_angle(float a) {
    a = fmodf_(a + PI_F, TWO_PI_F);
    if (a < 0.0f) a += TWO_PI_F;
    return a - PI_F;
}

}  // namespace isx
