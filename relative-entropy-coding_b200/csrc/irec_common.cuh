// irec_common.cuh -- device primitives shared by every kernel of libirec.so (sm_100a).
//
// Everything numerics-critical is written with explicit round-to-nearest intrinsics
// (__fmul_rn / __fadd_rn / __fmaf_rn / __fdiv_rn / __fsqrt_rn) so that neither -fmad nor the
// optimiser can change a rounding; the library is additionally compiled with -fmad=false.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#define IREC_PRIME 10007u
#define IREC_CHUNK 32
#define IREC_GEN 5u          // smallest primitive root of 10007
#define IREC_ORD 10006u      // order of the multiplicative group

// ---------------------------------------------------------------------------------------------
// Philox4x32-10 (TF core/lib/random/philox_random.h).  One call = 4 consecutive stream elements.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint4 philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                                               uint32_t k0, uint32_t k1)
{
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
        const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
        const uint32_t n0 = hi1 ^ c1 ^ k0;
        const uint32_t n2 = hi0 ^ c3 ^ k1;
        c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    return make_uint4(c0, c1, c2, c3);
}

// TF seed pair of a random op -> Philox key / counter-high words.
// key = (lo32(seed1), hi32(seed1)); counter = (lo(group), hi(group), lo32(seed2), hi32(seed2)).
struct TfStream {
    uint32_t k0, k1, c2, c3;
};

__host__ __device__ __forceinline__ int64_t tf_truncate_seed(int64_t s)
{
    const int64_t m = 2147483647LL;   // python/framework/random_seed.py _MAXINT32
    int64_t r = s % m;
    if (r < 0) r += m;
    return r;
}

// seeded op after tf.random.set_seed(g): (g, op); (0,0) -> (0, 2^31-1)
__host__ __device__ __forceinline__ TfStream tf_stream_seeded(int64_t g, int64_t op)
{
    int64_t s1 = tf_truncate_seed(g), s2 = tf_truncate_seed(op);
    if (s1 == 0 && s2 == 0) s2 = 2147483647LL;
    TfStream st;
    st.k0 = (uint32_t)s1; st.k1 = (uint32_t)((uint64_t)s1 >> 32);
    st.c2 = (uint32_t)s2; st.c3 = (uint32_t)((uint64_t)s2 >> 32);
    return st;
}

// 4 consecutive stream elements of group g (elements 4g..4g+3)
__device__ __forceinline__ uint4 tf_stream_group(const TfStream& st, uint64_t g)
{
    return philox4x32_10((uint32_t)g, (uint32_t)(g >> 32), st.c2, st.c3, st.k0, st.k1);
}

// 4 consecutive stream elements starting at an arbitrary element j0 (two groups when unaligned)
__device__ __forceinline__ uint4 tf_stream_quad_at(const TfStream& st, uint64_t j0)
{
    const uint64_t g = j0 >> 2;
    const uint32_t a = (uint32_t)(j0 & 3);
    uint4 x = tf_stream_group(st, g);
    if (a == 0) return x;
    uint4 y = tf_stream_group(st, g + 1);
    uint32_t v[8] = { x.x, x.y, x.z, x.w, y.x, y.y, y.z, y.w };
    uint4 o;
    // a in {1,2,3}
    o.x = a == 1 ? v[1] : (a == 2 ? v[2] : v[3]);
    o.y = a == 1 ? v[2] : (a == 2 ? v[3] : v[4]);
    o.z = a == 1 ? v[3] : (a == 2 ? v[4] : v[5]);
    o.w = a == 1 ? v[4] : (a == 2 ? v[5] : v[6]);
    return o;
}

// beam_search_coder.py:39-43  tf.random.uniform(minval=1, maxval=10007, int32): lo + u % (hi - lo)
__device__ __forceinline__ uint32_t beam_r_from_u32(uint32_t u) { return 1u + u % (IREC_PRIME - 1u); }

// beam_search_coder.py:45-47  k = floormod(r * h, 10007), r,h in [1,10006]  (product < 2^27)
__device__ __forceinline__ uint32_t beam_mix(uint32_t r, uint32_t h) { return (r * h) % IREC_PRIME; }

// beam_search_coder.py:33-35 simple_hash from the running wrapping int32 sum  sum_j idx_j * (69 + j)
__host__ __device__ __forceinline__ int32_t hash_from_sum(int32_t hsum)
{
    int32_t m = hsum % (int32_t)(IREC_PRIME - 1u);
    if (m < 0) m += (int32_t)(IREC_PRIME - 1u);
    return m + 1;
}
__host__ __device__ __forceinline__ int32_t hsum_extend(int32_t hsum, int32_t s, int t)
{
    return (int32_t)((uint32_t)hsum + (uint32_t)s * (uint32_t)(69 + t));
}

// TF random_distributions.h Uint32ToFloat
__device__ __forceinline__ float u32_to_float(uint32_t x)
{
    return __fadd_rn(__uint_as_float(0x3F800000u | (x & 0x7FFFFFu)), -1.0f);
}

// canonical float32 log / sin / cos: evaluate in float64, round once (DESIGN.md "float contracts")
__device__ __forceinline__ float c_logf(float x) { return (float)log((double)x); }

// TF random_distributions.h BoxMullerFloat (v1 = float(2.0f * M_PI(double) * u))
#ifdef IREC_BM_TABLES          // irec_is.cu: table-driven float64 log / sincos (irec_boxmuller.cuh), same float32 results
__device__ __forceinline__ void box_muller(uint32_t x0, uint32_t x1, float& f0, float& f1)
{
    float u1 = u32_to_float(x0);
    if (u1 < 1.0e-7f) u1 = 1.0e-7f;
    const float v1 = (float)(6.283185307179586 * (double)u32_to_float(x1));
    const BmTables t = IREC_BM_TABLES;
    const float u2 = __fsqrt_rn(__fmul_rn(-2.0f, bm_logf(u1, t)));
    float s, c;
    bm_sincosf(v1, t, s, c);
    f0 = __fmul_rn(s, u2);
    f1 = __fmul_rn(c, u2);
}
#else
__device__ __forceinline__ void box_muller(uint32_t x0, uint32_t x1, float& f0, float& f1)
{
    float u1 = u32_to_float(x0);
    if (u1 < 1.0e-7f) u1 = 1.0e-7f;
    const float v1 = (float)(6.283185307179586 * (double)u32_to_float(x1));
    const float u2 = __fsqrt_rn(__fmul_rn(-2.0f, c_logf(u1)));
    double sd, cd;
    sincos((double)v1, &sd, &cd);
    f0 = __fmul_rn((float)sd, u2);
    f1 = __fmul_rn((float)cd, u2);
}
#endif

// 4 consecutive N(0,1) stream elements of an ALIGNED group
__device__ __forceinline__ float4 tf_normal_group(const TfStream& st, uint64_t g)
{
    const uint4 u = tf_stream_group(st, g);
    float4 z;
    box_muller(u.x, u.y, z.x, z.y);
    box_muller(u.z, u.w, z.z, z.w);
    return z;
}

// single N(0,1) stream element j
__device__ __forceinline__ float tf_normal_elem(const TfStream& st, uint64_t j)
{
    const uint4 u = tf_stream_group(st, j >> 2);
    float f0, f1;
    if (j & 2) box_muller(u.z, u.w, f0, f1);
    else box_muller(u.x, u.y, f0, f1);
    return (j & 1) ? f1 : f0;
}

// ---------------------------------------------------------------------------------------------
// Packed FP32x2 arithmetic of sm_100 (SASS FMUL2 / FADD2 / FFMA2): one warp-instruction, two IEEE round-to-nearest float32
// operations on an aligned register pair -- the same bits as two scalar instructions, half the issue slots.
// ---------------------------------------------------------------------------------------------
typedef unsigned long long f32x2_t;
__device__ __forceinline__ f32x2_t f2_pack(float lo, float hi) { f32x2_t r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi)); return r; }
__device__ __forceinline__ void f2_unpack(f32x2_t v, float& lo, float& hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ f32x2_t f2_mul(f32x2_t a, f32x2_t b) { f32x2_t r; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ f32x2_t f2_add(f32x2_t a, f32x2_t b) { f32x2_t r; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ f32x2_t f2_fma(f32x2_t a, f32x2_t b, f32x2_t c) { f32x2_t r; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r; }

// ---------------------------------------------------------------------------------------------
// canonical reduction: leaves are 32-dim chunk sums; combine = balanced pairwise tree over the
// chunk index (pad with zeros to a power of two).  Helpers for the index mapping used in shared
// memory: dim d = 32*l + i is stored at   ((i>>2)*32 + l)*4 + (i&3)
// so that lane l (chunk l) reads a float4 of 4 consecutive dims conflict-free.
// ---------------------------------------------------------------------------------------------
__host__ __device__ __forceinline__ int chunk_interleave(int d)
{
    const int l = d >> 5, i = d & 31;
    return (((i >> 2) << 5) + l) * 4 + (i & 3);
}

__host__ __device__ __forceinline__ int next_pow2_int(int n)
{
    int p = 1;
    while (p < n) p <<= 1;
    return p;
}

// total order of candidates: higher score first, ties -> smaller flat index (tf.argsort DESCENDING
// == top_k: lowest index first)
__device__ __forceinline__ bool cand_better(float sa, int64_t fa, float sb, int64_t fb)
{
    return (sa > sb) || (sa == sb && fa < fb);
}
