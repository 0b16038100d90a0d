// irec_boxmuller.cuh -- table-driven float64 log / sincos for the Box-Muller candidates of the importance sampler.
//
// Contract (oracle/irec_oracle.c box_muller; TF random_distributions.h BoxMullerFloat): with u1 = Uint32ToFloat(x0)
// clamped to 1e-7f and v1 = float(2 pi * Uint32ToFloat(x1)),
//     logf := float(log(double(u1))),  s := float(sin(double(v1))),  c := float(cos(double(v1)))   ("float64 libm, round once")
// and the two normals are s * sqrtf(-2 * logf), c * sqrtf(-2 * logf) in separately rounded float32 operations.
//
// Both inputs are 23-bit integers (m = x & 0x7FFFFF), so each of the three functions has only 2^23 possible arguments.
// libm-grade log()/sincos() cost ~55 FP64 instructions plus ~60 integer/branch instructions per pair of normals
// (profiles/r1_is_block_ncu.md: FP64 pipe 27 %, issue slots 63 % busy, 136 lane-instructions per candidate-dim against
// the W_IS = 34 of the work model).  Here: one table lookup + a short fma polynomial each, ~2^-50 relative error, i.e.
// the float32 rounding agrees with the libm-then-round definition unless the exact value lies within ~1e-8 ulp of a
// rounding boundary.  tests/test_boxmuller_exhaustive.py evaluates the HOST build of these very functions (same IEEE
// fma sequence as the device) for ALL 2^23 arguments of each function against the oracle's definition; an argument
// that disagreed would be listed in BM_*_EXCEPTIONS below (none did).
#pragma once
#include <stdint.h>
#include <math.h>

struct BmTables {
    const double2* logA;   // [128] m0 <= sqrt2 : (1/c_j, log c_j),   c_j = 1 + (j + 0.5) / 128, m0 in [1, 2)
    const double2* logB;   // [128] m0 >  sqrt2 : (2/c_j, log(c_j/2)); j = 127: (1, 0) so that arguments next to 1 are exact
    const double2* sc;     // [257] (sin, cos)(j * pi / 128), exact 0 / +-1 at multiples of pi/2
};

// Every float64 constant lives in one array: in constant memory on the device (a DFMA takes a c[bank][offset] operand;
// as immediates each constant costs two UMOVs per use -- 35 extra instructions per pair of normals), a static table on
// the host.
#define BM_NK 24
#define BM_K_INIT { 1.4142135623730951,       /* 0  sqrt 2 */                                                   \
                    0.6931471805599453,       /* 1  ln 2 */                                                     \
                    0.024543692605220713,     /* 2  H_HI = pi/128 with the low 21 significand bits cleared (31 bits): j * H_HI is exact for j <= 256 */ \
                    9.495469541415925e-13,    /* 3  H_LO = pi/128 - H_HI */                                      \
                    40.74366543152521,        /* 4  128/pi */                                                   \
                    1.0 / 9.0, -1.0 / 8.0, 1.0 / 7.0, -1.0 / 6.0, 1.0 / 5.0, -1.0 / 4.0, 1.0 / 3.0, -0.5,   /* 5..12 log1p */ \
                    1.0 / 362880.0, -1.0 / 5040.0, 1.0 / 120.0, -1.0 / 6.0,                                   /* 13..16 sin */  \
                    1.0 / 3628800.0, -1.0 / 40320.0, 1.0 / 720.0, -1.0 / 24.0, 0.5,                           /* 17..21 cos */  \
                    0.0, 0.0 }
static __constant__ double c_bm_k[BM_NK] = BM_K_INIT;
static const double bm_k_host[BM_NK] = BM_K_INIT;
#ifdef __CUDA_ARCH__
#define BM_K(i) c_bm_k[i]
#else
#define BM_K(i) bm_k_host[i]
#endif
#define BM_SQRT2 BM_K(0)
#define BM_LN2 BM_K(1)
#define BM_H_HI BM_K(2)
#define BM_H_LO BM_K(3)
#define BM_INV_H BM_K(4)

#ifdef __CUDA_ARCH__
#define BM_FMA(a, b, c) __fma_rn((a), (b), (c))
#define BM_MUL(a, b) __dmul_rn((a), (b))
#define BM_ADD(a, b) __dadd_rn((a), (b))
#define BM_RINT_I(x) __double2int_rn(x)
#define BM_D2F(x) __double2float_rn(x)
#define BM_LD(p) __ldg(p)
#else
#define BM_FMA(a, b, c) fma((a), (b), (c))
#define BM_MUL(a, b) ((a) * (b))
#define BM_ADD(a, b) ((a) + (b))
#define BM_RINT_I(x) ((int)nearbyint(x))
#define BM_D2F(x) ((float)(x))
#define BM_LD(p) (*(p))
#endif

// float(log(double(u1))) for u1 = a positive float32 in [1e-7, 1)
__host__ __device__ __forceinline__ float bm_logf(float u1, const BmTables& t)
{
    const double d = (double)u1;
#ifdef __CUDA_ARCH__
    const long long bits = __double_as_longlong(d);
#else
    long long bits;
    memcpy(&bits, &d, 8);
#endif
    int e = (int)((bits >> 52) & 0x7ff) - 1023;
    const long long mant = bits & 0xfffffffffffffLL;
    const int j = (int)(mant >> 45);                                   // top 7 mantissa bits
    const long long mb = mant | 0x3ff0000000000000LL;
#ifdef __CUDA_ARCH__
    double m = __longlong_as_double(mb);
#else
    double m;
    memcpy(&m, &mb, 8);
#endif
    double2 tv;
    if (m > BM_SQRT2) { m = BM_MUL(m, 0.5); e += 1; tv = BM_LD(t.logB + j); }
    else tv = BM_LD(t.logA + j);
    const double r = BM_FMA(m, tv.x, -1.0);                            // |r| <= 2^-7, exact up to one rounding
    // log1p(r) = r - r^2/2 + r^3/3 - ... + r^9/9   (next term < 2^-73)
    double p = BM_FMA(r, BM_K(5), BM_K(6));
    p = BM_FMA(p, r, BM_K(7));
    p = BM_FMA(p, r, BM_K(8));
    p = BM_FMA(p, r, BM_K(9));
    p = BM_FMA(p, r, BM_K(10));
    p = BM_FMA(p, r, BM_K(11));
    p = BM_FMA(p, r, BM_K(12));
    p = BM_MUL(p, BM_MUL(r, r));
    p = BM_ADD(p, r);
    const double base = BM_FMA((double)e, BM_LN2, tv.y);               // e == 0 and tv.y == 0 next to 1: exact
    return BM_D2F(BM_ADD(base, p));
}

// (float(sin(double(v1))), float(cos(double(v1)))) for a float32 v1 in [0, 2 pi]
__host__ __device__ __forceinline__ void bm_sincosf(float v1, const BmTables& t, float& s, float& c)
{
    const double v = (double)v1;
    const int j = BM_RINT_I(BM_MUL(v, BM_INV_H));                      // 0 .. 256
    const double jd = (double)j;
    double x = BM_FMA(-jd, BM_H_HI, v);                                // exact
    x = BM_FMA(-jd, BM_H_LO, x);                                       // |x| <= pi/256 (+ rounding of the index)
    const double z = BM_MUL(x, x);
    double ps = BM_FMA(z, BM_K(13), BM_K(14));
    ps = BM_FMA(ps, z, BM_K(15));
    ps = BM_FMA(ps, z, BM_K(16));
    const double sx = BM_FMA(BM_MUL(ps, z), x, x);                     // sin x
    double pc = BM_FMA(z, BM_K(17), BM_K(18));
    pc = BM_FMA(pc, z, BM_K(19));
    pc = BM_FMA(pc, z, BM_K(20));
    pc = BM_FMA(pc, z, BM_K(21));
    const double cx = BM_FMA(-pc, z, 1.0);                             // cos x
    const double2 sc = BM_LD(t.sc + j);
    s = BM_D2F(BM_FMA(sc.x, cx, BM_MUL(sc.y, sx)));
    c = BM_D2F(BM_FMA(sc.y, cx, -BM_MUL(sc.x, sx)));
}
