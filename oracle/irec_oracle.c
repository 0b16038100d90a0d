/*
 * irec_oracle.c -- CPU ORACLE (test infrastructure, NOT product code).
 *
 * A plain-C restatement of the iREC encode/decode inner loop of
 * gergely-flamich/relative-entropy-coding (rec/coding), used ONLY by tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline leg as the checker for the
 * CUDA path in relative-entropy-coding_b200/csrc.  Nothing in the product imports it.
 *
 * PARITY STATUS: "parity unpinned" for emitted indices / absolute sample values --
 * the reference is TensorFlow 2.1 + TFP 0.9 (not installable here) and its own tests
 * are encode->decode round trips only.  What IS pinned against reference material:
 *   - Philox4x32-10 constants, key/counter layout, TF seed plumbing, Uint32ToFloat and
 *     the element<->counter mapping: notebooks/Discrete REC.ipynb:51,64-66 (100
 *     Bernoulli(0.7) draws after tf.random.set_seed(42)) -- see tests/test_oracle_kat.py.
 *   - S = int(exp(3)) = 20 (notebooks/scratch.ipynb:403-412).
 * Everything else restates the published TF 2.1 / TFP 0.9 algorithms (named below) and
 * DEFINES the float semantics those leave open (see "canonical" notes).
 *
 * Build: gcc -O2 -fPIC -shared -ffp-contract=off -fno-fast-math (oracle/build.py).
 * All float32 arithmetic below is IEEE round-to-nearest with NO fma contraction unless
 * fmaf() is written explicitly.
 *
 * Reference citations are relative to /root/reference.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define IREC_PRIME 10007          /* beam_search_coder.py:30 big_prime */
#define IREC_CHUNK 32             /* canonical reduction chunk (dims)   */

/* ------------------------------------------------------------------------------------------
 * Philox4x32-10  (TF core/lib/random/philox_random.h; Random123)
 * ---------------------------------------------------------------------------------------- */
static void philox4x32_10(uint32_t k0, uint32_t k1, const uint32_t ctr[4], uint32_t out[4])
{
    uint32_t c0 = ctr[0], c1 = ctr[1], c2 = ctr[2], c3 = ctr[3];
    for (int r = 0; r < 10; ++r) {
        uint64_t p0 = (uint64_t)0xD2511F53u * c0;
        uint64_t p1 = (uint64_t)0xCD9E8D57u * c2;
        uint32_t hi0 = (uint32_t)(p0 >> 32), lo0 = (uint32_t)p0;
        uint32_t hi1 = (uint32_t)(p1 >> 32), lo1 = (uint32_t)p1;
        uint32_t n0 = hi1 ^ c1 ^ k0;
        uint32_t n2 = hi0 ^ c3 ^ k1;
        c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

/* raw u32 element j of the TF op stream with seeds (seed1, seed2):
 * key=(lo32(seed1),hi32(seed1)), counter=(lo(j/4),hi(j/4),lo32(seed2),hi32(seed2)), lane j%4
 * (TF core/kernels/random_op.cc FillPhiloxRandom; GuardedPhiloxRandom::Init). */
static uint32_t tf_stream_u32(int64_t seed1, int64_t seed2, uint64_t j)
{
    uint64_t g = j >> 2;
    uint32_t ctr[4] = { (uint32_t)g, (uint32_t)(g >> 32), (uint32_t)seed2, (uint32_t)((uint64_t)seed2 >> 32) };
    uint32_t out[4];
    philox4x32_10((uint32_t)seed1, (uint32_t)((uint64_t)seed1 >> 32), ctr, out);
    return out[j & 3];
}

void orc_philox_raw(uint32_t k0, uint32_t k1, const uint32_t* ctr, uint32_t* out)
{
    philox4x32_10(k0, k1, ctr, out);
}

void orc_tf_stream_u32(int64_t seed1, int64_t seed2, int64_t start, int64_t n, uint32_t* out)
{
    for (int64_t i = 0; i < n; ++i) out[i] = tf_stream_u32(seed1, seed2, (uint64_t)(start + i));
}

/* TF python/framework/random_seed.py: _truncate_seed(seed) = seed % (2^31-1); (0,0)->(0,2^31-1) */
static int64_t tf_truncate_seed(int64_t s)
{
    const int64_t m = 2147483647LL;
    int64_t r = s % m;
    if (r < 0) r += m;
    return r;
}

/* ------------------------------------------------------------------------------------------
 * Python random.Random(seed).randint(0, 2**31-1): MT19937 seeded by init_by_array(abs(seed)
 * as little-endian 32-bit words), then getrandbits(32) rejection until < 2^31
 * (CPython Modules/_randommodule.c, Lib/random.py _randbelow_with_getrandbits).
 * This is TF's Context._internal_operation_seed() for the first unseeded op after set_seed.
 * ---------------------------------------------------------------------------------------- */
typedef struct { uint32_t mt[624]; int idx; } mt_t;
static void mt_init_genrand(mt_t* s, uint32_t seed)
{
    s->mt[0] = seed;
    for (int i = 1; i < 624; ++i)
        s->mt[i] = 1812433253u * (s->mt[i - 1] ^ (s->mt[i - 1] >> 30)) + (uint32_t)i;
    s->idx = 624;
}
static void mt_init_by_array(mt_t* s, const uint32_t* key, int len)
{
    mt_init_genrand(s, 19650218u);
    int i = 1, j = 0;
    int k = (624 > len ? 624 : len);
    for (; k; --k) {
        s->mt[i] = (s->mt[i] ^ ((s->mt[i - 1] ^ (s->mt[i - 1] >> 30)) * 1664525u)) + key[j] + (uint32_t)j;
        ++i; ++j;
        if (i >= 624) { s->mt[0] = s->mt[623]; i = 1; }
        if (j >= len) j = 0;
    }
    for (k = 623; k; --k) {
        s->mt[i] = (s->mt[i] ^ ((s->mt[i - 1] ^ (s->mt[i - 1] >> 30)) * 1566083941u)) - (uint32_t)i;
        ++i;
        if (i >= 624) { s->mt[0] = s->mt[623]; i = 1; }
    }
    s->mt[0] = 0x80000000u;
}
static uint32_t mt_genrand(mt_t* s)
{
    if (s->idx >= 624) {
        uint32_t* mt = s->mt;
        int kk;
        for (kk = 0; kk < 624 - 397; ++kk) {
            uint32_t y = (mt[kk] & 0x80000000u) | (mt[kk + 1] & 0x7fffffffu);
            mt[kk] = mt[kk + 397] ^ (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
        }
        for (; kk < 623; ++kk) {
            uint32_t y = (mt[kk] & 0x80000000u) | (mt[kk + 1] & 0x7fffffffu);
            mt[kk] = mt[kk + (397 - 624)] ^ (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
        }
        uint32_t y = (mt[623] & 0x80000000u) | (mt[0] & 0x7fffffffu);
        mt[623] = mt[396] ^ (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
        s->idx = 0;
    }
    uint32_t y = s->mt[s->idx++];
    y ^= (y >> 11);
    y ^= (y << 7) & 0x9d2c5680u;
    y ^= (y << 15) & 0xefc60000u;
    y ^= (y >> 18);
    return y;
}

int64_t orc_py_randint31(int64_t seed)
{
    uint64_t a = (uint64_t)(seed < 0 ? -seed : seed);
    uint32_t key[2] = { (uint32_t)a, (uint32_t)(a >> 32) };
    int len = key[1] ? 2 : 1;
    mt_t s;
    mt_init_by_array(&s, key, len);
    for (;;) {
        uint32_t r = mt_genrand(&s);       /* getrandbits(32) */
        if (r < 0x80000000u) return (int64_t)r;
    }
}

/* (seed1, seed2) of an op after tf.random.set_seed(g):  seeded op -> (g, op); unseeded -> (g, randint) */
static void tf_seeds_seeded(int64_t g, int64_t op, int64_t* s1, int64_t* s2)
{
    *s1 = tf_truncate_seed(g); *s2 = tf_truncate_seed(op);
    if (*s1 == 0 && *s2 == 0) *s2 = 2147483647LL;
}
static void tf_seeds_unseeded_first(int64_t g, int64_t* s1, int64_t* s2)
{
    *s1 = tf_truncate_seed(g); *s2 = tf_truncate_seed(orc_py_randint31(g));
    if (*s1 == 0 && *s2 == 0) *s2 = 2147483647LL;
}

/* TF random_distributions.h Uint32ToFloat: bits(0x3F800000 | (x & 0x7FFFFF)) - 1.0f */
static float u32_to_float(uint32_t x)
{
    uint32_t v = 0x3F800000u | (x & 0x7FFFFFu);
    float f;
    memcpy(&f, &v, 4);
    return f - 1.0f;
}

/* tf.random.uniform(float32) after set_seed(g), unseeded op: the notebook known-answer stream */
void orc_tf_uniform_f32_unseeded(int64_t g, int64_t n, float* out)
{
    int64_t s1, s2;
    tf_seeds_unseeded_first(g, &s1, &s2);
    for (int64_t j = 0; j < n; ++j) out[j] = u32_to_float(tf_stream_u32(s1, s2, (uint64_t)j));
}

/* beam_search_coder.py:38-43: set_seed(q); tf.random.uniform(minval=1, maxval=10007, seed=q, int32)
 * -> UniformDistribution<int32>: lo + u32 % (hi-lo)  */
void orc_beam_uniform_int(int64_t q, int64_t start, int64_t n, int32_t* out)
{
    int64_t s1, s2;
    tf_seeds_seeded(q, q, &s1, &s2);
    for (int64_t i = 0; i < n; ++i)
        out[i] = 1 + (int32_t)(tf_stream_u32(s1, s2, (uint64_t)(start + i)) % (uint32_t)(IREC_PRIME - 1));
}

/* canonical logf/sinf/cosf: evaluate in float64 libm, round once to float32 (SURVEY 8c policy) */
static float c_logf(float x) { return (float)log((double)x); }

/* TF random_distributions.h BoxMullerFloat; element j of tf.random.normal stream */
static void box_muller(uint32_t x0, uint32_t x1, float* f0, float* f1)
{
    const float epsilon = 1.0e-7f;
    float u1 = u32_to_float(x0);
    if (u1 < epsilon) u1 = epsilon;
    const float v1 = (float)(2.0 * 3.14159265358979323846 * (double)u32_to_float(x1));
    const float u2 = sqrtf(-2.0f * c_logf(u1));
    float s = (float)sin((double)v1), c = (float)cos((double)v1);
    *f0 = s * u2;
    *f1 = c * u2;
}

/* the three float32 functions behind box_muller, by definition (float64 libm, round once), for m = 23 mantissa bits */
void orc_bm_components(const uint32_t* m, int64_t n, float* logf_out, float* sin_out, float* cos_out)
{
    for (int64_t i = 0; i < n; ++i) {
        const float u = u32_to_float(m[i]);
        float u1 = u;
        if (u1 < 1.0e-7f) u1 = 1.0e-7f;
        const float v1 = (float)(2.0 * 3.14159265358979323846 * (double)u);
        if (logf_out) logf_out[i] = c_logf(u1);
        if (sin_out) sin_out[i] = (float)sin((double)v1);
        if (cos_out) cos_out[i] = (float)cos((double)v1);
    }
}

static float tf_stream_normal(int64_t s1, int64_t s2, uint64_t j)
{
    uint64_t g = j >> 2;
    uint32_t ctr[4] = { (uint32_t)g, (uint32_t)(g >> 32), (uint32_t)s2, (uint32_t)((uint64_t)s2 >> 32) };
    uint32_t o[4];
    philox4x32_10((uint32_t)s1, (uint32_t)((uint64_t)s1 >> 32), ctr, o);
    float f0, f1;
    int pair = (int)((j & 3) >> 1);
    box_muller(o[2 * pair], o[2 * pair + 1], &f0, &f1);
    return (j & 1) ? f1 : f0;
}

/* importance_sampling.py:38,54: set_seed(seed); Normal(0,1).sample(S) (unseeded op) */
void orc_is_normal_stream(int64_t seed, int64_t start, int64_t n, float* out)
{
    int64_t s1, s2;
    tf_seeds_unseeded_first(seed, &s1, &s2);
    for (int64_t i = 0; i < n; ++i) out[i] = tf_stream_normal(s1, s2, (uint64_t)(start + i));
}

/* ------------------------------------------------------------------------------------------
 * TFP 0.9 special_math._ndtri in float32 (Cephes rational approximations, Horner with a
 * separate rounded multiply and add per step).  Only used to build the 10006-entry table
 * T[k] = ndtri_f32(float32(k)/float32(10007))  (beam_search_coder.py:48-49).
 * ---------------------------------------------------------------------------------------- */
static const double P0d[5] = { -5.99633501014107895267E1, 9.80010754185999661536E1, -5.66762857469070293439E1,
                               1.39312609387279679503E1, -1.23916583867381258016E0 };
static const double Q0d[9] = { 1.0, 1.95448858338141759834E0, 4.67627912898881538453E0, 8.63602421390890590575E1,
                               -2.25462687854119370527E2, 2.00260212380060660359E2, -8.20372256168333339912E1,
                               1.59056225126211695515E1, -1.18331621121330003142E0 };
static const double P1d[9] = { 4.05544892305962419923E0, 3.15251094599893866154E1, 5.71628192246421288162E1,
                               4.40805073893200834700E1, 1.46849561928858024014E1, 2.18663306850790267539E0,
                               -1.40256079171354495875E-1, -3.50424626827848203418E-2, -8.57456785154685413611E-4 };
static const double Q1d[9] = { 1.0, 1.57799883256466749731E1, 4.53907635128879210584E1, 4.13172038254672030440E1,
                               1.50425385692907503408E1, 2.50464946208309415979E0, -1.42182922854787788574E-1,
                               -3.80806407691578277194E-2, -9.33259480895457427372E-4 };
static const double P2d[9] = { 3.23774891776946035970E0, 6.91522889068984211695E0, 3.93881025292474443415E0,
                               1.33303460815807542389E0, 2.01485389549179081538E-1, 1.23716634817820021358E-2,
                               3.01581553508235416007E-4, 2.65806974686737550832E-6, 6.23974539184983293730E-9 };
static const double Q2d[9] = { 1.0, 6.02427039364742014255E0, 3.67983563856160859403E0, 1.37702099489081330271E0,
                               2.16236993594496635890E-1, 1.34204006088543189037E-2, 3.28014464682127739104E-4,
                               2.89247864745380683936E-6, 6.79019408009981274425E-9 };

/* coefficients are listed highest order first (Cephes); TFP reverses them and evaluates
 * c[0] + (c[1] + (...)*x)*x, i.e. Horner from the FIRST listed coefficient down. */
static float horner_f32(const double* c_hi_first, int n, float x)
{
    float r = (float)c_hi_first[0];
    for (int i = 1; i < n; ++i) {
        float m = r * x;
        r = (float)c_hi_first[i] + m;
    }
    return r;
}

float orc_ndtri_f32(float p)
{
    const float one_minus_em2 = (float)0.8646647167633873;   /* -np.expm1(-2.) */
    const float em2 = (float)0.1353352832366127;             /* np.exp(-2.)    */
    float mcp = (p > one_minus_em2) ? (1.0f - p) : p;
    float s = (mcp <= 0.0f) ? 0.5f : mcp;
    /* p > exp(-2) branch */
    float w = s - 0.5f;
    float ww = w * w;
    float ratio0 = horner_f32(P0d, 5, ww) / horner_f32(Q0d, 9, ww);
    float xb = w + (w * ww) * ratio0;
    xb = xb * (float)(-2.5066282746310002);                   /* -np.sqrt(2*pi) */
    /* tails */
    float z = sqrtf(-2.0f * c_logf(s));
    float first = z - c_logf(z) / z;
    float iz = 1.0f / z;
    float second_small = (horner_f32(P2d, 9, iz) / horner_f32(Q2d, 9, iz)) / z;
    float second_other = (horner_f32(P1d, 9, iz) / horner_f32(Q1d, 9, iz)) / z;
    float x;
    if (s > em2) x = xb;
    else if (z >= 8.0f) x = first - second_small;
    else x = first - second_other;
    x = (p > one_minus_em2) ? x : -x;
    if (p <= 0.0f) return -INFINITY;
    if (p >= 1.0f) return INFINITY;
    return x;
}

/* T[0] is unused (k ranges over 1..10006) and set to 0 */
void orc_ndtri_table(float* T)
{
    T[0] = 0.0f;
    for (int k = 1; k < IREC_PRIME; ++k) T[k] = orc_ndtri_f32((float)k / (float)IREC_PRIME);
}

/* ------------------------------------------------------------------------------------------
 * coder.py:16,218-220  ratio(i) = float32(np.power(i + 1., -0.7864636765648174))
 * ---------------------------------------------------------------------------------------- */
/* coder.py:197-231: learned ratios (extrapolate_auxiliary_ratios=False) replace the power law; per-thread override,
 * set by orc_set_aux_ratios(host table, n) and cleared with (NULL, 0) */
#define ORC_MAX_LEARNED 4096
static __thread float g_learned[ORC_MAX_LEARNED];
static __thread int g_n_learned = 0;
int orc_set_aux_ratios(const float* ratios, int n)
{
    if (n < 0 || n > ORC_MAX_LEARNED || (n > 0 && !ratios)) return -1;
    for (int i = 0; i < n; ++i) g_learned[i] = ratios[i];
    g_n_learned = n;
    return 0;
}
float orc_aux_ratio(int i)
{
    if (g_n_learned > 0) return (i >= 0 && i < g_n_learned) ? g_learned[i] : 0.0f;
    return (float)pow((double)i + 1.0, -0.7864636765648174);
}

/* ------------------------------------------------------------------------------------------
 * canonical chunk/tree reduction helpers
 *   leaves: chunk sums c_i over dims [32 i, 32 i + 32) accumulated in ascending d;
 *   combine: pad to P = next_pow2(#chunks) with zeros, then
 *            for (stride = 1; stride < P; stride *= 2) v[i] += v[i + stride]  (i multiple of 2*stride)
 * ---------------------------------------------------------------------------------------- */
static int next_pow2(int n) { int p = 1; while (p < n) p <<= 1; return p; }

static float tree_sum_f32(float* v, int n)
{
    int P = next_pow2(n);
    for (int stride = 1; stride < P; stride <<= 1)
        for (int i = 0; i + stride < n; i += 2 * stride) v[i] = v[i] + v[i + stride];
    /* entries >= n are zeros: v[i] + 0 == v[i], so skipping them is exact */
    return v[0];
}
static double tree_sum_f64(double* v, int n)
{
    int P = next_pow2(n);
    for (int stride = 1; stride < P; stride <<= 1)
        for (int i = 0; i + stride < n; i += 2 * stride) v[i] = v[i] + v[i + stride];
    return v[0];
}

/* ------------------------------------------------------------------------------------------
 * KL and partition count  (coder.py:499-501, beam_search_coder.py:57-59; TFP kl_normal_normal)
 * canonical: per-dim terms in float64 from the float32 inputs, chunk/tree sum in float64,
 * rounded once to float32; n_aux = ceil(KL32 / omega32) in float32.
 * ---------------------------------------------------------------------------------------- */
float orc_kl(const float* t_loc, const float* t_scale, const float* p_loc, const float* p_scale, int D)
{
    int nch = (D + IREC_CHUNK - 1) / IREC_CHUNK;
    double* c = (double*)calloc((size_t)nch, sizeof(double));
    for (int i = 0; i < nch; ++i) {
        double acc = 0.0;
        int hi = (i + 1) * IREC_CHUNK < D ? (i + 1) * IREC_CHUNK : D;
        for (int d = i * IREC_CHUNK; d < hi; ++d) {
            double sp = (double)p_scale[d];
            double dl = log((double)t_scale[d]) - log(sp);
            double dm = (double)t_loc[d] / sp - (double)p_loc[d] / sp;
            double kl = 0.5 * (dm * dm) + 0.5 * expm1(2.0 * dl) - dl;
            acc = acc + kl;
        }
        c[i] = acc;
    }
    double kl = tree_sum_f64(c, nch);
    free(c);
    return (float)kl;
}

int orc_n_aux(float kl, float omega)
{
    float q = kl / omega;
    if (!(q == q) || isinf(q)) return -1;
    return (int)ceilf(q);
}

/* ------------------------------------------------------------------------------------------
 * Beam-search schedule, partition t  (beam_search_coder.py:64-77,108-109; coder.py:141-154)
 * float32 exactly as Appendix A.3 of SURVEY.md; pow(x,2) := x*x.
 * Per dim state: cum[d] (cumulative auxiliary variance, starts 0).
 * Outputs for partition t: sa = sqrt(v_t) (scale of the candidates), the auxiliary-target mean m,
 * and the canonical CENTRED quadratic log-ratio coefficients  A = 0.5*(1/tot - 1/s2),  E = m/tot
 * (float64 from the float32 schedule values, rounded once to float32), so that with d = x - m
 *     log q_aux(x) - log p_cum(x) = A d^2 + E d + const.
 * Centring at m keeps the form free of cancellation when the target is narrow (s2 << tot, the
 * expanded A x^2 + B x loses all digits there).  The per-partition constant does not influence
 * the selection and is dropped.
 * ---------------------------------------------------------------------------------------- */
static void beam_schedule_step(const float* t_loc, const float* t_scale, const float* p_loc, const float* p_scale,
                               int D, float ratio, float* cum, float* sa, float* A, float* E, float* M,
                               float* s2_out, float* tot_out)
{
    for (int d = 0; d < D; ++d) {
        float cv = p_scale[d] * p_scale[d];
        float tv = t_scale[d] * t_scale[d];
        float v = ratio * (cv - cum[d]);
        float tot = v + cum[d];
        float m = ((t_loc[d] - p_loc[d]) * tot) / cv;
        float s2 = (tv * (tot * tot)) / (cv * cv) + (tot * (cv - tot)) / cv;
        sa[d] = sqrtf(v);
        double a = 0.5 * (1.0 / (double)tot - 1.0 / (double)s2);
        double e = (double)m / (double)tot;
        A[d] = (float)a;
        E[d] = (float)e;
        M[d] = m;
        if (s2_out) { s2_out[d] = s2; tot_out[d] = tot; }
        cum[d] = cum[d] + v;
    }
}

/* whole table for tests: sa/A/E/M are [n_aux x D] row-major */
void orc_beam_schedule(const float* t_loc, const float* t_scale, const float* p_loc, const float* p_scale,
                       int D, int n_aux, float* sa, float* A, float* E, float* M)
{
    float* cum = (float*)calloc((size_t)D, sizeof(float));
    for (int t = 0; t < n_aux; ++t)
        beam_schedule_step(t_loc, t_scale, p_loc, p_scale, D, orc_aux_ratio(n_aux - 1 - t), cum,
                           sa + (size_t)t * D, A + (size_t)t * D, E + (size_t)t * D, M + (size_t)t * D, NULL, NULL);
    free(cum);
}

/* beam_search_coder.py:33-35 simple_hash from the running int32 (wrapping) sum */
static int32_t hash_from_sum(int32_t hsum)
{
    int32_t m = hsum % (IREC_PRIME - 1);
    if (m < 0) m += (IREC_PRIME - 1);
    return m + 1;
}
static int32_t hsum_extend(int32_t hsum, int32_t s, int t)
{
    return (int32_t)((uint32_t)hsum + (uint32_t)s * (uint32_t)(69 + t));
}
int32_t orc_simple_hash(const int32_t* idx, int n)
{
    int32_t h = 0;
    for (int j = 0; j < n; ++j) h = hsum_extend(h, idx[j], j);
    return hash_from_sum(h);
}

/* candidate value a = T[(r*h) mod 10007] * sa   (beam_search_coder.py:45-49) */
static inline float beam_candidate(const float* T, int32_t r, int32_t h, float sa)
{
    int32_t k = (int32_t)(((int64_t)r * (int64_t)h) % IREC_PRIME);
    return T[k] * sa;
}

/* canonical score of candidate (s,b): chunk sums of  acc = fma(fma(A,d,E), d, acc),  d = x - m,
 * x = beam + a (rounded add of the rounded product = the actual candidate value), tree-combined.
 * r_row = r[s, 0..D) */
static float beam_score(const float* T, const int32_t* r_row, int32_t h, const float* beam /*may be NULL = zeros*/,
                        const float* sa, const float* A, const float* E, const float* M, int D, float* scratch)
{
    int nch = (D + IREC_CHUNK - 1) / IREC_CHUNK;
    for (int i = 0; i < nch; ++i) {
        float acc = 0.0f;
        int hi = (i + 1) * IREC_CHUNK < D ? (i + 1) * IREC_CHUNK : D;
        for (int d = i * IREC_CHUNK; d < hi; ++d) {
            float a = beam_candidate(T, r_row[d], h, sa[d]);
            float x = (beam ? beam[d] : 0.0f) + a;
            float dd = x - M[d];
            float tt = fmaf(A[d], dd, E[d]);
            acc = fmaf(tt, dd, acc);
        }
        scratch[i] = acc;
    }
    return tree_sum_f32(scratch, nch);
}

/* scores for one partition given explicit state (teacher-forced mode).
 * beams [Bc x D] (NULL => single zero beam), hsum [Bcur]; scores [S x Bcur] (flat f = s*Bcur + b) */
void orc_beam_scores(const float* T, const float* sa, const float* A, const float* E, const float* M,
                     const float* beams, const int32_t* hsum, int D, int S, int Bcur, int64_t q, float* scores)
{
    int32_t* r = (int32_t*)malloc(sizeof(int32_t) * (size_t)D);
    float* scratch = (float*)malloc(sizeof(float) * (size_t)((D + IREC_CHUNK - 1) / IREC_CHUNK));
    for (int s = 0; s < S; ++s) {
        orc_beam_uniform_int(q, (int64_t)s * D, D, r);
        for (int b = 0; b < Bcur; ++b)
            scores[(size_t)s * Bcur + b] =
                beam_score(T, r, hash_from_sum(hsum[b]), beams ? beams + (size_t)b * D : NULL, sa, A, E, M, D, scratch);
    }
    free(r);
    free(scratch);
}

/* top-k by (score desc, flat index asc)  == tf.argsort(DESCENDING)[:k] (top_k: lowest index first on ties) */
static void top_k_desc(const float* v, int64_t n, int k, int64_t* out)
{
    /* simple O(n*k) selection; n*k is small in the oracle's use */
    char* used = (char*)calloc((size_t)n, 1);
    for (int j = 0; j < k; ++j) {
        int64_t best = -1;
        for (int64_t i = 0; i < n; ++i) {
            if (used[i]) continue;
            if (best < 0 || v[i] > v[best]) best = i;
        }
        used[best] = 1;
        out[j] = best;
    }
    free(used);
}
void orc_top_k_desc(const float* v, int64_t n, int k, int64_t* out) { top_k_desc(v, n, k, out); }

/* ------------------------------------------------------------------------------------------
 * BeamSearchCoder.encode_block  (beam_search_coder.py:53-122)
 * returns 0 ok, 1 = KL not finite / n_aux <= 0, 2 = n_aux > max_aux
 * trace_* (optional, may be NULL): per partition, the kept beams' (score, s, parent b): [n_aux x B]
 * ---------------------------------------------------------------------------------------- */
int orc_beam_encode(const float* t_loc, const float* t_scale, const float* p_loc, const float* p_scale, int D,
                    float omega, int S, int B, int64_t seed, int32_t* out_indices, int max_aux, int32_t* out_n_aux,
                    float* out_sample, float* out_kl, float* trace_score, int32_t* trace_s, int32_t* trace_b,
                    int32_t* trace_nbeams)
{
    static float T[IREC_PRIME];
    static int T_ready = 0;
    if (!T_ready) { orc_ndtri_table(T); T_ready = 1; }

    float kl = orc_kl(t_loc, t_scale, p_loc, p_scale, D);
    if (out_kl) *out_kl = kl;
    int n_aux = orc_n_aux(kl, omega);
    *out_n_aux = n_aux;
    if (n_aux <= 0) return 1;
    if (n_aux > max_aux) return 2;

    float* cum = (float*)calloc((size_t)D, sizeof(float));
    float* sa = (float*)malloc(sizeof(float) * D);
    float* A = (float*)malloc(sizeof(float) * D);
    float* E = (float*)malloc(sizeof(float) * D);
    float* M = (float*)malloc(sizeof(float) * D);
    float* beams = (float*)calloc((size_t)B * D, sizeof(float));
    float* nbeams = (float*)calloc((size_t)B * D, sizeof(float));
    int32_t* hsum = (int32_t*)calloc((size_t)B, sizeof(int32_t));
    int32_t* nhsum = (int32_t*)calloc((size_t)B, sizeof(int32_t));
    int32_t* hist = (int32_t*)calloc((size_t)B * n_aux, sizeof(int32_t));
    int32_t* nhist = (int32_t*)calloc((size_t)B * n_aux, sizeof(int32_t));
    float* scores = (float*)malloc(sizeof(float) * (size_t)S * B);
    int64_t* top = (int64_t*)malloc(sizeof(int64_t) * B);
    int32_t* r = (int32_t*)malloc(sizeof(int32_t) * D);
    int Bcur = 1;                       /* t = 0: one empty beam (zeros, hash 1) */

    for (int t = 0; t < n_aux; ++t) {
        beam_schedule_step(t_loc, t_scale, p_loc, p_scale, D, orc_aux_ratio(n_aux - 1 - t), cum, sa, A, E, M,
                           NULL, NULL);
        int64_t q = seed + t;
        orc_beam_scores(T, sa, A, E, M, t == 0 ? NULL : beams, hsum, D, S, Bcur, q, scores);
        int64_t ncand = (int64_t)S * Bcur;
        int keep = (int)(ncand < B ? ncand : B);
        top_k_desc(scores, ncand, keep, top);
        for (int j = 0; j < keep; ++j) {
            int s = (int)(top[j] / Bcur), b = (int)(top[j] % Bcur);
            orc_beam_uniform_int(q, (int64_t)s * D, D, r);
            int32_t h = hash_from_sum(hsum[b]);
            for (int d = 0; d < D; ++d) {
                float a = beam_candidate(T, r[d], h, sa[d]);
                nbeams[(size_t)j * D + d] = (t == 0 ? 0.0f : beams[(size_t)b * D + d]) + a;
            }
            memcpy(nhist + (size_t)j * n_aux, hist + (size_t)b * n_aux, sizeof(int32_t) * (size_t)t);
            nhist[(size_t)j * n_aux + t] = s;
            nhsum[j] = hsum_extend(hsum[b], s, t);
            if (trace_score) {
                trace_score[(size_t)t * B + j] = scores[top[j]];
                trace_s[(size_t)t * B + j] = s;
                trace_b[(size_t)t * B + j] = b;
            }
        }
        if (trace_nbeams) trace_nbeams[t] = keep;
        float* tf_ = beams; beams = nbeams; nbeams = tf_;
        int32_t* ti = hist; hist = nhist; nhist = ti;
        ti = hsum; hsum = nhsum; nhsum = ti;
        Bcur = keep;
    }
    for (int t = 0; t < n_aux; ++t) out_indices[t] = hist[t];
    for (int d = 0; d < D; ++d) out_sample[d] = beams[d] + p_loc[d];

    free(cum); free(sa); free(A); free(E); free(M); free(beams); free(nbeams); free(hsum); free(nhsum);
    free(hist); free(nhist); free(scores); free(top); free(r);
    return 0;
}

/* ------------------------------------------------------------------------------------------
 * BeamSearchCoder.decode_block  (beam_search_coder.py:124-148). indices are in partition order
 * t = 0..n_aux-1 (the order encode returns them; the reference reverses its list in place and
 * then reads it back to front, which is the same thing).
 * ---------------------------------------------------------------------------------------- */
int orc_beam_decode(const float* p_loc, const float* p_scale, int D, int S, int64_t seed, const int32_t* indices,
                    int n_aux, float* out_sample)
{
    static float T[IREC_PRIME];
    static int T_ready = 0;
    if (!T_ready) { orc_ndtri_table(T); T_ready = 1; }
    (void)S;
    float* cum = (float*)calloc((size_t)D, sizeof(float));
    float* sample = (float*)calloc((size_t)D, sizeof(float));
    int32_t* r = (int32_t*)malloc(sizeof(int32_t) * D);
    int32_t hsum = 0;
    for (int t = 0; t < n_aux; ++t) {
        float ratio = orc_aux_ratio(n_aux - 1 - t);
        int64_t q = seed + t;
        int32_t s = indices[t];
        int32_t h = hash_from_sum(hsum);
        orc_beam_uniform_int(q, (int64_t)s * D, D, r);
        for (int d = 0; d < D; ++d) {
            float cv = p_scale[d] * p_scale[d];
            float v = ratio * (cv - cum[d]);
            float a = beam_candidate(T, r[d], h, sqrtf(v));
            sample[d] = sample[d] + a;
            cum[d] = cum[d] + v;
        }
        hsum = hsum_extend(hsum, s, t);
    }
    for (int d = 0; d < D; ++d) out_sample[d] = sample[d] + p_loc[d];
    free(cum); free(sample); free(r);
    return 0;
}

/* ------------------------------------------------------------------------------------------
 * Reference-FORM log-weight of one candidate vector x at partition t (for the tolerance tests):
 * sum_d [ lp(x; m, sqrt(s2)) - lp(x; 0, sqrt(tot)) ] with TFP's float32 _log_prob,
 * accumulated in float64 (so that it is a tight estimate of what the float32 reference targets).
 * ---------------------------------------------------------------------------------------- */
static float tfp_log_prob(float x, float loc, float scale)
{
    float a = x / scale, b = loc / scale;
    float d = a - b;
    float lu = -0.5f * (d * d);
    float ln = (float)(0.5 * log(2.0 * 3.14159265358979323846)) + c_logf(scale);
    return lu - ln;
}
double orc_beam_refform_logw(const float* t_loc, const float* t_scale, const float* p_loc, const float* p_scale,
                             int D, int n_aux, int t, const float* x)
{
    float* cum = (float*)calloc((size_t)D, sizeof(float));
    float* sa = (float*)malloc(sizeof(float) * D);
    float* A = (float*)malloc(sizeof(float) * D);
    float* Bc = (float*)malloc(sizeof(float) * D);
    float* m = (float*)malloc(sizeof(float) * D);
    float* s2 = (float*)malloc(sizeof(float) * D);
    float* tot = (float*)malloc(sizeof(float) * D);
    for (int u = 0; u <= t; ++u)
        beam_schedule_step(t_loc, t_scale, p_loc, p_scale, D, orc_aux_ratio(n_aux - 1 - u), cum, sa, A, Bc, m, s2, tot);
    double acc = 0.0;
    for (int d = 0; d < D; ++d)
        acc += (double)(tfp_log_prob(x[d], m[d], sqrtf(s2[d])) - tfp_log_prob(x[d], 0.0f, sqrtf(tot[d])));
    free(cum); free(sa); free(A); free(Bc); free(m); free(s2); free(tot);
    return acc;
}
/* the dropped per-partition constant: log-ratio at x = m,  sum_d [ 0.5 m^2/tot - 0.5 log s2 + 0.5 log tot ]  (float64) */
double orc_beam_score_constant(const float* t_loc, const float* t_scale, const float* p_loc, const float* p_scale,
                               int D, int n_aux, int t)
{
    float* cum = (float*)calloc((size_t)D, sizeof(float));
    float* sa = (float*)malloc(sizeof(float) * D);
    float* A = (float*)malloc(sizeof(float) * D);
    float* Bc = (float*)malloc(sizeof(float) * D);
    float* m = (float*)malloc(sizeof(float) * D);
    float* s2 = (float*)malloc(sizeof(float) * D);
    float* tot = (float*)malloc(sizeof(float) * D);
    for (int u = 0; u <= t; ++u)
        beam_schedule_step(t_loc, t_scale, p_loc, p_scale, D, orc_aux_ratio(n_aux - 1 - u), cum, sa, A, Bc, m, s2, tot);
    double acc = 0.0;
    for (int d = 0; d < D; ++d)
        acc += 0.5 * (double)m[d] * (double)m[d] / (double)tot[d] - 0.5 * log((double)s2[d]) + 0.5 * log((double)tot[d]);
    free(cum); free(sa); free(A); free(Bc); free(m); free(s2); free(tot);
    return acc;
}

/* ------------------------------------------------------------------------------------------
 * Population study: the canonical log-weight (centred quadratic, fixed reduction tree) against the reference's OWN
 * form (beam_search_coder.py:79-106): per candidate  reduce_sum_d( lp(x; m, sqrt(s2)) - lp(x; 0, sqrt(tot)) )  in
 * float32 with TFP's Normal._log_prob, then argsort DESCENDING.  TensorFlow's reduce_sum order over the last axis is
 * an Eigen implementation detail, so the sum is evaluated in three ways (sum_mode):
 *   0 = sequential float32,  1 = pairwise float32 (8 lanes, blocks of 128: NumPy's scheme),  2 = float64 accumulator.
 * Two comparisons per coder-block:
 *   free-running   : the reference-form coder runs on its own beam state; its indices vs the canonical indices;
 *   teacher-forced : at every partition of the CANONICAL run, reference-form scores of the same state; the kept set
 *                    (top-B) vs the canonical kept set.  When they differ, gap = the largest |w_in - w_out| over the
 *                    swapped candidates, relative to max |w| of the kept set, measured in the reference form (float64
 *                    accumulation): north_star's 1e-5 tolerance applies to it.
 * ---------------------------------------------------------------------------------------- */
typedef struct {
    int32_t n_aux;
    int32_t free_identical;          /* reference-form indices == canonical indices */
    int32_t free_first_diff;         /* first differing partition (or -1) */
    int32_t tf_partitions;           /* partitions compared teacher-forced */
    int32_t tf_set_mismatch;         /* partitions whose kept SET differs */
    int32_t tf_order_mismatch;       /* partitions whose kept set is equal but ordered differently */
    int32_t tf_best_mismatch;        /* partitions whose FIRST (best) candidate differs */
    int32_t pad;
    double tf_max_rel_gap;           /* worst relative gap over the set mismatches (0 when none) */
    double max_rel_score_dev;        /* max |canonical + const - refform64| / max(1, |refform64|) over all candidates */
    double max_dev_canon_exact;      /* max |canonical + const - exact| / max(1, |exact|): exact = the log-ratio in float64 throughout */
    double max_dev_ref32_exact;      /* max |reference float32 form (sum_mode) + const' - exact| / max(1, |exact|) */
} orc_refform_stats_t;

static float refform_sum(const float* term, int D, int sum_mode)
{
    if (sum_mode == 0) {
        float a = 0.0f;
        for (int d = 0; d < D; ++d) a = a + term[d];
        return a;
    }
    if (sum_mode == 2) {
        double a = 0.0;
        for (int d = 0; d < D; ++d) a += (double)term[d];
        return (float)a;
    }
    /* NumPy pairwise_sum for float32 (npy PW_BLOCKSIZE 128, 8 accumulators) */
    float stack[64];
    int sp = 0, cnt = 0;
    for (int base = 0; base < D; base += 128) {
        int n = D - base < 128 ? D - base : 128;
        float blk;
        if (n < 8) {
            blk = 0.0f;
            for (int i = 0; i < n; ++i) blk = blk + term[base + i];
        } else {
            float r[8];
            for (int j = 0; j < 8; ++j) r[j] = term[base + j];
            int i;
            for (i = 8; i + 8 <= n; i += 8)
                for (int j = 0; j < 8; ++j) r[j] = r[j] + term[base + i + j];
            blk = ((r[0] + r[1]) + (r[2] + r[3])) + ((r[4] + r[5]) + (r[6] + r[7]));
            for (; i < n; ++i) blk = blk + term[base + i];
        }
        stack[sp++] = blk;
        ++cnt;
        for (int c = cnt; (c & 1) == 0; c >>= 1) { stack[sp - 2] = stack[sp - 2] + stack[sp - 1]; --sp; }
    }
    while (sp > 1) { stack[sp - 2] = stack[sp - 2] + stack[sp - 1]; --sp; }
    return sp ? stack[0] : 0.0f;
}

/* reference-form scores of one partition for an explicit state; ref64 (may be NULL): the same sums accumulated in double */
static void refform_scores(const float* T, const float* sa, const float* m, const float* s2, const float* tot,
                           const float* beams, const int32_t* hsum, int D, int S, int Bcur, int64_t q, int sum_mode,
                           float* scores, double* ref64, double* exact64)
{
    int32_t* r = (int32_t*)malloc(sizeof(int32_t) * (size_t)D);
    float* term = (float*)malloc(sizeof(float) * (size_t)D);
    float* qs = (float*)malloc(sizeof(float) * (size_t)D);
    float* cs = (float*)malloc(sizeof(float) * (size_t)D);
    float* qb = (float*)malloc(sizeof(float) * (size_t)D);
    float* qn = (float*)malloc(sizeof(float) * (size_t)D);
    float* cn = (float*)malloc(sizeof(float) * (size_t)D);
    const float half_log_2pi = (float)(0.5 * log(2.0 * 3.14159265358979323846));
    for (int d = 0; d < D; ++d) {
        qs[d] = sqrtf(s2[d]); cs[d] = sqrtf(tot[d]);
        qb[d] = m[d] / qs[d];
        qn[d] = half_log_2pi + c_logf(qs[d]);
        cn[d] = half_log_2pi + c_logf(cs[d]);
    }
    for (int s = 0; s < S; ++s) {
        orc_beam_uniform_int(q, (int64_t)s * D, D, r);
        for (int b = 0; b < Bcur; ++b) {
            const int32_t h = hash_from_sum(hsum[b]);
            const float* beam = beams ? beams + (size_t)b * D : NULL;
            double a64 = 0.0, e64 = 0.0;
            for (int d = 0; d < D; ++d) {
                float x = (beam ? beam[d] : 0.0f) + beam_candidate(T, r[d], h, sa[d]);
                if (exact64) {       /* log N(x; m, s2) - log N(x; 0, tot), float64 throughout from the float32 schedule values */
                    double dx = (double)x - (double)m[d];
                    e64 += -0.5 * dx * dx / (double)s2[d] + 0.5 * (double)x * (double)x / (double)tot[d]
                           - 0.5 * log((double)s2[d]) + 0.5 * log((double)tot[d]);
                }
                float dq = x / qs[d] - qb[d];
                float dc = x / cs[d] - 0.0f / cs[d];
                float lq = -0.5f * (dq * dq) - qn[d];
                float lc = -0.5f * (dc * dc) - cn[d];
                term[d] = lq - lc;
                a64 += (double)term[d];
            }
            scores[(size_t)s * Bcur + b] = refform_sum(term, D, sum_mode);
            if (ref64) ref64[(size_t)s * Bcur + b] = a64;
            if (exact64) exact64[(size_t)s * Bcur + b] = e64;
        }
    }
    free(r); free(term); free(qs); free(cs); free(qb); free(qn); free(cn);
}

static void commit_winners(const float* T, const float* sa, const int64_t* top, int keep, int Bcur, int D, int n_aux, int t,
                           int64_t q, float** beams, float** nbeams, int32_t** hsum, int32_t** nhsum, int32_t** hist,
                           int32_t** nhist, int32_t* r)
{
    for (int j = 0; j < keep; ++j) {
        int s = (int)(top[j] / Bcur), b = (int)(top[j] % Bcur);
        orc_beam_uniform_int(q, (int64_t)s * D, D, r);
        int32_t h = hash_from_sum((*hsum)[b]);
        for (int d = 0; d < D; ++d)
            (*nbeams)[(size_t)j * D + d] = (t == 0 ? 0.0f : (*beams)[(size_t)b * D + d]) + beam_candidate(T, r[d], h, sa[d]);
        memcpy(*nhist + (size_t)j * n_aux, *hist + (size_t)b * n_aux, sizeof(int32_t) * (size_t)t);
        (*nhist)[(size_t)j * n_aux + t] = s;
        (*nhsum)[j] = hsum_extend((*hsum)[b], s, t);
    }
    float* tf_ = *beams; *beams = *nbeams; *nbeams = tf_;
    int32_t* ti = *hist; *hist = *nhist; *nhist = ti;
    ti = *hsum; *hsum = *nhsum; *nhsum = ti;
}

int orc_beam_refform_study(const float* t_loc, const float* t_scale, const float* p_loc, const float* p_scale, int D,
                           float omega, int S, int B, int64_t seed, int sum_mode, int32_t* idx_canon, int32_t* idx_ref,
                           int max_aux, orc_refform_stats_t* st)
{
    static float T[IREC_PRIME];
    static int T_ready = 0;
    if (!T_ready) { orc_ndtri_table(T); T_ready = 1; }
    memset(st, 0, sizeof(*st));
    st->free_first_diff = -1;
    const float kl = orc_kl(t_loc, t_scale, p_loc, p_scale, D);
    const int n_aux = orc_n_aux(kl, omega);
    st->n_aux = n_aux;
    if (n_aux <= 0) return 1;
    if (n_aux > max_aux) return 2;

    float* sa = (float*)malloc(sizeof(float) * D); float* A = (float*)malloc(sizeof(float) * D);
    float* E = (float*)malloc(sizeof(float) * D); float* M = (float*)malloc(sizeof(float) * D);
    float* s2 = (float*)malloc(sizeof(float) * D); float* tot = (float*)malloc(sizeof(float) * D);
    float* scores = (float*)malloc(sizeof(float) * (size_t)S * B);
    float* rscores = (float*)malloc(sizeof(float) * (size_t)S * B);
    double* r64 = (double*)malloc(sizeof(double) * (size_t)S * B);
    double* x64 = (double*)malloc(sizeof(double) * (size_t)S * B);
    int64_t* top = (int64_t*)malloc(sizeof(int64_t) * B);
    int64_t* rtop = (int64_t*)malloc(sizeof(int64_t) * B);
    int32_t* r = (int32_t*)malloc(sizeof(int32_t) * D);

    for (int pass = 0; pass < 2; ++pass) {          /* 0: canonical run + teacher-forced comparison, 1: free-running reference form */
        float* cum = (float*)calloc((size_t)D, sizeof(float));
        float* beams = (float*)calloc((size_t)B * D, sizeof(float));
        float* nbeams = (float*)calloc((size_t)B * D, sizeof(float));
        int32_t* hsum = (int32_t*)calloc((size_t)B, sizeof(int32_t));
        int32_t* nhsum = (int32_t*)calloc((size_t)B, sizeof(int32_t));
        int32_t* hist = (int32_t*)calloc((size_t)B * n_aux, sizeof(int32_t));
        int32_t* nhist = (int32_t*)calloc((size_t)B * n_aux, sizeof(int32_t));
        int Bcur = 1;
        for (int t = 0; t < n_aux; ++t) {
            beam_schedule_step(t_loc, t_scale, p_loc, p_scale, D, orc_aux_ratio(n_aux - 1 - t), cum, sa, A, E, M, s2, tot);
            const int64_t q = seed + t;
            const int64_t ncand = (int64_t)S * Bcur;
            const int keep = (int)(ncand < B ? ncand : B);
            if (pass == 0) {
                orc_beam_scores(T, sa, A, E, M, t == 0 ? NULL : beams, hsum, D, S, Bcur, q, scores);
                refform_scores(T, sa, M, s2, tot, t == 0 ? NULL : beams, hsum, D, S, Bcur, q, sum_mode, rscores, r64, x64);
                top_k_desc(scores, ncand, keep, top);
                top_k_desc(rscores, ncand, keep, rtop);
                /* canonical = reference form - per-partition constant: compare through the differences to the best */
                double wmax = 0.0;
                for (int j = 0; j < keep; ++j) { double w = fabs(r64[top[j]]); if (w > wmax) wmax = w; }
                if (wmax < 1.0) wmax = 1.0;
                {
                    const double c = r64[top[0]] - (double)scores[top[0]];
                    for (int64_t i = 0; i < ncand; ++i) {
                        if (!(scores[i] == scores[i]) || isinf(scores[i])) continue;
                        double dev = fabs(((double)scores[i] + c) - r64[i]) / (fabs(r64[i]) > 1.0 ? fabs(r64[i]) : 1.0);
                        if (dev > st->max_rel_score_dev) st->max_rel_score_dev = dev;
                    }
                    /* against the exact value: the canonical form drops a per-partition constant (fixed through the best
                       candidate); the float32 reference form carries its own constant */
                    const double ce = x64[top[0]] - (double)scores[top[0]];
                    const double cr = x64[top[0]] - (double)rscores[top[0]];
                    for (int64_t i = 0; i < ncand; ++i) {
                        if (!(scores[i] == scores[i]) || isinf(scores[i])) continue;
                        const double den = fabs(x64[i]) > 1.0 ? fabs(x64[i]) : 1.0;
                        double d1 = fabs(((double)scores[i] + ce) - x64[i]) / den;
                        double d2 = fabs(((double)rscores[i] + cr) - x64[i]) / den;
                        if (d1 > st->max_dev_canon_exact) st->max_dev_canon_exact = d1;
                        if (d2 > st->max_dev_ref32_exact) st->max_dev_ref32_exact = d2;
                    }
                }
                st->tf_partitions++;
                int same_order = 1, same_set = 1;
                for (int j = 0; j < keep; ++j) if (top[j] != rtop[j]) same_order = 0;
                if (!same_order) {
                    double gap = 0.0;
                    for (int j = 0; j < keep; ++j) {
                        int in_r = 0, in_c = 0;
                        for (int k = 0; k < keep; ++k) { if (rtop[k] == top[j]) in_r = 1; if (top[k] == rtop[j]) in_c = 1; }
                        if (!in_r || !in_c) same_set = 0;
                    }
                    if (!same_set) {
                        /* swapped candidates: those kept by exactly one side; the gap is the spread of their reference-form weights */
                        double lo = 1e300, hi = -1e300;
                        for (int j = 0; j < keep; ++j) {
                            int in_r = 0, in_c = 0;
                            for (int k = 0; k < keep; ++k) { if (rtop[k] == top[j]) in_r = 1; if (top[k] == rtop[j]) in_c = 1; }
                            if (!in_r) { double w = r64[top[j]]; if (w < lo) lo = w; if (w > hi) hi = w; }
                            if (!in_c) { double w = r64[rtop[j]]; if (w < lo) lo = w; if (w > hi) hi = w; }
                        }
                        gap = (hi - lo) / wmax;
                        if (gap > st->tf_max_rel_gap) st->tf_max_rel_gap = gap;
                        st->tf_set_mismatch++;
                    } else {
                        st->tf_order_mismatch++;
                    }
                    if (top[0] != rtop[0]) st->tf_best_mismatch++;
                }
                commit_winners(T, sa, top, keep, Bcur, D, n_aux, t, q, &beams, &nbeams, &hsum, &nhsum, &hist, &nhist, r);
            } else {
                refform_scores(T, sa, M, s2, tot, t == 0 ? NULL : beams, hsum, D, S, Bcur, q, sum_mode, rscores, NULL, NULL);
                top_k_desc(rscores, ncand, keep, rtop);
                commit_winners(T, sa, rtop, keep, Bcur, D, n_aux, t, q, &beams, &nbeams, &hsum, &nhsum, &hist, &nhist, r);
            }
            Bcur = keep;
        }
        int32_t* out = pass == 0 ? idx_canon : idx_ref;
        for (int t = 0; t < n_aux; ++t) out[t] = hist[t];
        free(cum); free(beams); free(nbeams); free(hsum); free(nhsum); free(hist); free(nhist);
    }
    st->free_identical = 1;
    for (int t = 0; t < n_aux; ++t)
        if (idx_canon[t] != idx_ref[t]) { st->free_identical = 0; st->free_first_diff = t; break; }
    free(sa); free(A); free(E); free(M); free(s2); free(tot); free(scores); free(rscores); free(r64); free(x64); free(top); free(rtop); free(r);
    return 0;
}

/* KL in float32 the way TFP 0.9 kl_normal_normal evaluates it (coder.py:499, beam_search_coder.py:57), summed in float32
 * (sum_mode as above) -- to count the blocks whose n_aux differs from the canonical float64 KL */
float orc_kl_f32(const float* t_loc, const float* t_scale, const float* p_loc, const float* p_scale, int D, int sum_mode)
{
    float* term = (float*)malloc(sizeof(float) * (size_t)D);
    for (int d = 0; d < D; ++d) {
        float dl = c_logf(t_scale[d]) - c_logf(p_scale[d]);
        float a = t_loc[d] / p_scale[d], b = p_loc[d] / p_scale[d];
        float sq = (a - b) * (a - b);
        float em = (float)expm1((double)(2.0f * dl));
        term[d] = (0.5f * sq + 0.5f * em) - dl;
    }
    float r = refform_sum(term, D, sum_mode);
    free(term);
    return r;
}

/* ------------------------------------------------------------------------------------------
 * Importance sampler  (importance_sampling.py:9-103) and GaussianCoder blocks (coder.py:493-584)
 * ---------------------------------------------------------------------------------------- */
/* importance_sampling.py:51  S = int32(ceil(exp(coding_bits * log(2.)))) -- canonical: float32 product,
 * exp in float64 rounded once to float32, ceil. */
int32_t orc_is_num_samples(float coding_bits)
{
    float prod = coding_bits * (float)0.6931471805599453;
    float e = (float)exp((double)prod);
    return (int32_t)ceilf(e);
}

/* canonical IS score of sample s: chunk/tree sum of (A d + E) d, d = z - mu', z from the normal stream;
 * A = 0.5 (1 - 1/sigma'^2), E = mu'  (log N(z;mu',sigma') - log N(z;0,1) = A d^2 + E d + const) */
static float is_score(int64_t s1, int64_t s2, int64_t s, const float* A, const float* E, const float* M, int D,
                      float* scratch)
{
    int nch = (D + IREC_CHUNK - 1) / IREC_CHUNK;
    for (int i = 0; i < nch; ++i) {
        float acc = 0.0f;
        int hi = (i + 1) * IREC_CHUNK < D ? (i + 1) * IREC_CHUNK : D;
        for (int d = i * IREC_CHUNK; d < hi; ++d) {
            float z = tf_stream_normal(s1, s2, (uint64_t)(s * D + d));
            float dd = z - M[d];
            float tt = fmaf(A[d], dd, E[d]);
            acc = fmaf(tt, dd, acc);
        }
        scratch[i] = acc;
    }
    return tree_sum_f32(scratch, nch);
}

/* encode_gaussian_importance_sample, alpha = inf (importance_sampling.py:9-79).
 * scores_out optional [S]. */
int64_t orc_is_coded_sample(const float* t_loc, const float* t_scale, const float* p_loc, const float* p_scale,
                            int D, int32_t S, int64_t seed, float* out_sample, float* scores_out)
{
    int64_t s1, s2;
    tf_seeds_unseeded_first(seed, &s1, &s2);
    float* A = (float*)malloc(sizeof(float) * D);
    float* E = (float*)malloc(sizeof(float) * D);
    float* M = (float*)malloc(sizeof(float) * D);
    float* scratch = (float*)malloc(sizeof(float) * (size_t)((D + IREC_CHUNK - 1) / IREC_CHUNK));
    for (int d = 0; d < D; ++d) {
        float mu = (t_loc[d] - p_loc[d]) / p_scale[d];
        float sg = t_scale[d] / p_scale[d];
        double s2d = (double)sg * (double)sg;
        A[d] = (float)(0.5 * (1.0 - 1.0 / s2d));
        E[d] = mu;
        M[d] = mu;
    }
    int64_t best = 0;
    float bestv = 0.0f;
    for (int64_t s = 0; s < S; ++s) {
        float v = is_score(s1, s2, s, A, E, M, D, scratch);
        if (scores_out) scores_out[s] = v;
        if (s == 0 || v > bestv) { best = s; bestv = v; }
    }
    for (int d = 0; d < D; ++d) {
        float z = tf_stream_normal(s1, s2, (uint64_t)(best * D + d));
        out_sample[d] = p_scale[d] * z + p_loc[d];
    }
    free(A); free(E); free(M); free(scratch);
    return best;
}

/* decode_gaussian_importance_sample (importance_sampling.py:82-103) */
void orc_is_decode_sample(const float* p_loc, const float* p_scale, int D, int64_t index, int64_t seed, float* out)
{
    int64_t s1, s2;
    tf_seeds_unseeded_first(seed, &s1, &s2);
    for (int d = 0; d < D; ++d) {
        float z = tf_stream_normal(s1, s2, (uint64_t)(index * D + d));
        out[d] = p_scale[d] * z + p_loc[d];
    }
}

/* GaussianCoder.encode_block with an ImportanceSampler (coder.py:493-559, 141-171) */
int orc_is_encode_block(const float* t_loc_in, const float* t_scale_in, const float* p_loc_in, const float* p_scale_in,
                        int D, float omega, int32_t S, int64_t seed, int64_t* out_indices, int max_aux,
                        int32_t* out_n_idx, float* out_sample, float* out_kl)
{
    float kl = orc_kl(t_loc_in, t_scale_in, p_loc_in, p_scale_in, D);
    if (out_kl) *out_kl = kl;
    int n_aux = orc_n_aux(kl, omega);
    if (n_aux < 0) { *out_n_idx = n_aux; return 1; }
    int n_idx = n_aux > 1 ? n_aux : 1;
    *out_n_idx = n_idx;
    if (n_idx > max_aux) return 2;
    size_t bytes = sizeof(float) * (size_t)D;
    float* tl = (float*)malloc(bytes); float* ts = (float*)malloc(bytes);
    float* pl = (float*)malloc(bytes); float* ps = (float*)malloc(bytes);
    float* al = (float*)malloc(bytes); float* as = (float*)malloc(bytes);
    float* zl = (float*)calloc((size_t)D, sizeof(float)); float* sv = (float*)malloc(bytes);
    float* a = (float*)malloc(bytes); float* vv = (float*)malloc(bytes);
    memcpy(tl, t_loc_in, bytes); memcpy(ts, t_scale_in, bytes);
    memcpy(pl, p_loc_in, bytes); memcpy(ps, p_scale_in, bytes);
    int n = 0;
    for (int i = n_aux - 1; i >= 1; --i) {
        float ratio = orc_aux_ratio(i);
        for (int d = 0; d < D; ++d) {
            float cv = ps[d] * ps[d], tv = ts[d] * ts[d];
            float v = ratio * cv;
            vv[d] = v;
            al[d] = ((tl[d] - pl[d]) * v) / cv;
            as[d] = sqrtf((tv * (v * v)) / (cv * cv) + (v * (cv - v)) / cv);
            sv[d] = sqrtf(v);
        }
        out_indices[n++] = orc_is_coded_sample(al, as, zl, sv, D, S, seed, a, NULL);
        seed += 1;
        for (int d = 0; d < D; ++d) {
            float cv = ps[d] * ps[d], tv = ts[d] * ts[d], v = vv[d];
            float num = (a[d] * tv) * cv + ((tl[d] - pl[d]) * (cv - v)) * cv;
            float den = tv * v + cv * (cv - v);
            float nm = pl[d] + num / den;
            float nv = ((tv * cv) * (cv - v)) / (v * tv + cv * (cv - v));
            tl[d] = nm;
            ts[d] = sqrtf(nv);
            pl[d] = pl[d] + a[d];
            ps[d] = sqrtf(cv - v);
        }
    }
    out_indices[n++] = orc_is_coded_sample(tl, ts, pl, ps, D, S, seed, out_sample, NULL);
    free(tl); free(ts); free(pl); free(ps); free(al); free(as); free(zl); free(sv); free(a); free(vv);
    return 0;
}

/* GaussianCoder.decode_block (coder.py:561-584); indices in the order encode emitted them */
int orc_is_decode_block(const float* p_loc_in, const float* p_scale_in, int D, int64_t seed, const int64_t* indices,
                        int n_idx, float* out_sample)
{
    size_t bytes = sizeof(float) * (size_t)D;
    float* pl = (float*)malloc(bytes); float* ps = (float*)malloc(bytes);
    float* zl = (float*)calloc((size_t)D, sizeof(float)); float* sv = (float*)malloc(bytes);
    float* a = (float*)malloc(bytes);
    memcpy(pl, p_loc_in, bytes); memcpy(ps, p_scale_in, bytes);
    int n = 0;
    for (int i = n_idx - 1; i >= 1; --i) {
        float ratio = orc_aux_ratio(i);
        for (int d = 0; d < D; ++d) sv[d] = sqrtf(ratio * (ps[d] * ps[d]));
        orc_is_decode_sample(zl, sv, D, indices[n++], seed, a);
        seed += 1;
        for (int d = 0; d < D; ++d) {
            float cv = ps[d] * ps[d];
            float v = ratio * cv;
            pl[d] = pl[d] + a[d];
            ps[d] = sqrtf(cv - v);
        }
    }
    orc_is_decode_sample(pl, ps, D, indices[n++], seed, out_sample);
    free(pl); free(ps); free(zl); free(sv); free(a);
    return 0;
}

/* ------------------------------------------------------------------------------------------
 * Coder.split permutation (coder.py:60-67): set_seed(seed); tf.random.shuffle(range(n)) (unseeded op)
 * TF core/kernels/random_shuffle_op.cc: forward Fisher-Yates, one u32 per step,
 *   for i in 0..n-2: swap(a[i], a[i + u32 % (n - i)])
 * ---------------------------------------------------------------------------------------- */
void orc_shuffle_perm(int64_t n, int64_t seed, int64_t* perm)
{
    int64_t s1, s2;
    tf_seeds_unseeded_first(seed, &s1, &s2);
    for (int64_t i = 0; i < n; ++i) perm[i] = i;
    uint64_t j = 0;
    for (int64_t i = 0; i + 1 < n; ++i) {
        uint32_t u = tf_stream_u32(s1, s2, j++);
        int64_t k = i + (int64_t)(u % (uint32_t)(n - i));
        int64_t tmp = perm[i]; perm[i] = perm[k]; perm[k] = tmp;
    }
}
