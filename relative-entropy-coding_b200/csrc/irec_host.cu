// irec_host.cu -- library state, one-time device tables, TensorFlow seed plumbing restated on the
// host (MT19937 op seeds, RandomShuffle permutation), error reporting.
#include <math.h>
#include <stdio.h>
#include <string.h>
#include <atomic>
#include <mutex>
#include <algorithm>
#include <cstdlib>

#include "irec_common.cuh"
#include "irec_host.h"

static thread_local char g_err[512] = "";
static std::atomic<int64_t> g_launches{0};
static IrecDevice g_dev[64];
static std::mutex g_mu;

int irec_fail(int code, const char* msg)
{
    snprintf(g_err, sizeof(g_err), "%s", msg);
    return code;
}
int irec_check_launch(const char* what)
{
    const cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
        snprintf(g_err, sizeof(g_err), "%s: %s", what, cudaGetErrorString(e));
        return IREC_E_CUDA;
    }
    return IREC_OK;
}
void irec_count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }
bool irec_force_general()
{
    const char* e = getenv("IREC_FORCE_GENERAL");
    return e && e[0] == '1';
}
const IrecDevice& irec_device()
{
    int d = 0;
    cudaGetDevice(&d);
    return g_dev[d & 63];
}

// learned auxiliary ratios (rec/coding/coder.py:218-231): a per-thread override of the power-law table, installed by
// irec_set_thread_aux_ratios for the calls that follow on the same host thread
static thread_local const float* tl_ratio = nullptr;
static thread_local int tl_ratio_len = 0;
const float* irec_ratio_tab() { return tl_ratio ? tl_ratio : irec_device().d_ratio; }
// SMs the persistent batch kernels leave free for the calling thread's launches (include/irec.h: irec_set_thread_reserved_sms)
static thread_local int tl_reserved_sms = -1;
int irec_reserved_sms()
{
    if (tl_reserved_sms >= 0) return tl_reserved_sms;
    const char* e = getenv("IREC_RESERVE_SMS");
    return e ? std::max(0, atoi(e)) : 0;
}
int irec_ratio_len() { return tl_ratio ? tl_ratio_len : irec_device().ratio_len; }

// ---------------------------------------------------------------------------------------------
// quantile table: TFP 0.9 special_math._ndtri in float32 (Cephes P0/Q0, P1/Q1, P2/Q2; Horner with
// a separately rounded multiply and add per step), evaluated at p = float32(k) / float32(10007)
// (rec/coding/beam_search_coder.py:48-49).  logf := float64 log rounded once.
// ---------------------------------------------------------------------------------------------
__constant__ float c_P0[5], c_Q0[9], c_P1[9], c_Q1[9], c_P2[9], c_Q2[9];

__device__ __forceinline__ float horner_f32(const float* c, int n, float x)
{
    float r = c[0];
    for (int i = 1; i < n; ++i) r = __fadd_rn(c[i], __fmul_rn(r, x));
    return r;
}

__device__ float ndtri_f32(float p)
{
    const float one_minus_em2 = (float)0.8646647167633873;
    const float em2 = (float)0.1353352832366127;
    const float mcp = (p > one_minus_em2) ? __fadd_rn(1.0f, -p) : p;
    const float s = (mcp <= 0.0f) ? 0.5f : mcp;
    const float w = __fadd_rn(s, -0.5f);
    const float ww = __fmul_rn(w, w);
    const float ratio0 = __fdiv_rn(horner_f32(c_P0, 5, ww), horner_f32(c_Q0, 9, ww));
    float xb = __fadd_rn(w, __fmul_rn(__fmul_rn(w, ww), ratio0));
    xb = __fmul_rn(xb, (float)(-2.5066282746310002));
    const float z = __fsqrt_rn(__fmul_rn(-2.0f, c_logf(s)));
    const float first = __fadd_rn(z, -__fdiv_rn(c_logf(z), z));
    const float iz = __fdiv_rn(1.0f, z);
    const float second_small = __fdiv_rn(__fdiv_rn(horner_f32(c_P2, 9, iz), horner_f32(c_Q2, 9, iz)), z);
    const float second_other = __fdiv_rn(__fdiv_rn(horner_f32(c_P1, 9, iz), horner_f32(c_Q1, 9, iz)), z);
    float x;
    if (s > em2) x = xb;
    else if (z >= 8.0f) x = __fadd_rn(first, -second_small);
    else x = __fadd_rn(first, -second_other);
    x = (p > one_minus_em2) ? x : -x;
    if (p <= 0.0f) return __int_as_float(0xff800000);
    if (p >= 1.0f) return __int_as_float(0x7f800000);
    return x;
}

__global__ void k_build_ndtri_table(float* T)
{
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k == 0) T[0] = 0.f;
    else if (k < (int)IREC_PRIME) T[k] = ndtri_f32(__fdiv_rn((float)k, (float)IREC_PRIME));
}

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

template <int N>
static cudaError_t upload_coeffs(const float (&sym)[N], const double* src)
{
    float tmp[N];
    for (int i = 0; i < N; ++i) tmp[i] = (float)src[i];
    return cudaMemcpyToSymbol(sym, tmp, sizeof(tmp));
}

extern "C" {

int irec_version(void) { return 100; }
const char* irec_last_error_string(void) { return g_err; }
int64_t irec_launch_count(void) { return g_launches.load(); }

int irec_set_thread_aux_ratios(const float* dev_ratios, int n)
{
    if ((dev_ratios == nullptr) != (n == 0) || n < 0) return irec_fail(IREC_E_INVALID, "irec_set_thread_aux_ratios: need (ptr, n > 0) or (NULL, 0)");
    tl_ratio = dev_ratios;
    tl_ratio_len = n;
    return IREC_OK;
}

int irec_set_thread_reserved_sms(int k)
{
    tl_reserved_sms = k < 0 ? -1 : k;
    return IREC_OK;
}

int irec_aux_ratio_len(void)
{
    if (tl_ratio) return tl_ratio_len;
    if (irec_init() != IREC_OK) return 0;
    return irec_device().ratio_len;
}

float irec_aux_ratio(int i)
{
    // rec/coding/coder.py:16,218-220: np.power(i + 1., -0.7864636765648174) (float64), cast to float32
    return (float)pow((double)i + 1.0, -0.7864636765648174);
}

int irec_init(void)
{
    int d = 0;
    if (cudaGetDevice(&d) != cudaSuccess) return irec_fail(IREC_E_CUDA, "irec_init: no CUDA device");
    if (d < 0 || d >= 64) return irec_fail(IREC_E_INVALID, "irec_init: device ordinal out of range");
    if (g_dev[d].ready) return IREC_OK;
    std::lock_guard<std::mutex> lock(g_mu);
    if (g_dev[d].ready) return IREC_OK;
    IrecDevice dev{};
    dev.device = d;
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, d) != cudaSuccess) return irec_fail(IREC_E_CUDA, "irec_init: cudaGetDeviceProperties failed");
    if (prop.major < 10) return irec_fail(IREC_E_CUDA, "irec_init: libirec.so is built for sm_100a (B200) only");
    dev.sm_count = prop.multiProcessorCount;
    dev.max_smem_optin = (int)prop.sharedMemPerBlockOptin;
    if (cudaMalloc(&dev.d_T, sizeof(float) * 10008) != cudaSuccess ||
        cudaMalloc(&dev.d_ratio, sizeof(float) * IREC_RATIO_LEN) != cudaSuccess)
        return irec_fail(IREC_E_CUDA, "irec_init: cudaMalloc failed");
    if (upload_coeffs(c_P0, P0d) != cudaSuccess || upload_coeffs(c_Q0, Q0d) != cudaSuccess ||
        upload_coeffs(c_P1, P1d) != cudaSuccess || upload_coeffs(c_Q1, Q1d) != cudaSuccess ||
        upload_coeffs(c_P2, P2d) != cudaSuccess || upload_coeffs(c_Q2, Q2d) != cudaSuccess)
        return irec_fail(IREC_E_CUDA, "irec_init: constant upload failed (is this an sm_100a device?)");
    k_build_ndtri_table<<<(10007 + 255) / 256, 256>>>(dev.d_T);
    irec_count_launch();
    std::vector<float> ratios(IREC_RATIO_LEN);
    for (int i = 0; i < IREC_RATIO_LEN; ++i) ratios[i] = irec_aux_ratio(i);
    if (cudaMemcpy(dev.d_ratio, ratios.data(), sizeof(float) * IREC_RATIO_LEN, cudaMemcpyHostToDevice) != cudaSuccess)
        return irec_fail(IREC_E_CUDA, "irec_init: ratio upload failed");
    dev.ratio_len = IREC_RATIO_LEN;
    if (cudaDeviceSynchronize() != cudaSuccess || cudaGetLastError() != cudaSuccess)
        return irec_fail(IREC_E_CUDA, "irec_init: table kernel failed");
    // Exponent-indexed view of the same table.  10007 is prime and g = 5 generates Z_10007^*, so with
    // a = dlog(r), c = dlog(h):  T[(r * h) mod 10007] = T2[a + c]  where T2[e] = T[g^(e mod 10006)];
    // the modular product of beam_search_coder.py:45-47 becomes one integer add.  T2 holds the same
    // float32 bit patterns as T.  It is a permutation, stored three times: a + c needs no wrap, and every
    // value lives at two word offsets a and a + 10006 whose shared-memory banks differ by 22 -- the per-launch
    // exponent table (k_r2_exps) picks per gather instruction the copy that keeps bank multiplicity at <= 2.
    {
        std::vector<float> T(10008), T2(IREC_T2_LEN, 0.f);
        std::vector<uint16_t> dl4(10006);
        if (cudaMemcpy(T.data(), dev.d_T, sizeof(float) * 10007, cudaMemcpyDeviceToHost) != cudaSuccess)
            return irec_fail(IREC_E_CUDA, "irec_init: table read-back failed");
        uint32_t pw = 1;
        for (int e = 0; e < 10006; ++e) {
            for (int rep = 0; rep < 3; ++rep) T2[e + 10006 * rep] = T[pw];
            dl4[pw - 1] = (uint16_t)(4 * e);
            pw = (pw * IREC_GEN) % IREC_PRIME;
        }
        if (pw != 1) return irec_fail(IREC_E_INVALID, "irec_init: generator check failed");
        if (cudaMalloc(&dev.d_T2, sizeof(float) * IREC_T2_LEN) != cudaSuccess ||
            cudaMalloc(&dev.d_dl4, sizeof(uint16_t) * 10008) != cudaSuccess ||
            cudaMemcpy(dev.d_T2, T2.data(), sizeof(float) * IREC_T2_LEN, cudaMemcpyHostToDevice) != cudaSuccess ||
            cudaMemcpy(dev.d_dl4, dl4.data(), sizeof(uint16_t) * 10006, cudaMemcpyHostToDevice) != cudaSuccess)
            return irec_fail(IREC_E_CUDA, "irec_init: exponent table upload failed");
    }
    dev.ready = true;
    g_dev[d] = dev;
    return IREC_OK;
}

int irec_get_ndtri_table(float* host_out)
{
    IREC_ENSURE_INIT();
    if (cudaMemcpy(host_out, irec_device().d_T, sizeof(float) * 10007, cudaMemcpyDeviceToHost) != cudaSuccess)
        return irec_fail(IREC_E_CUDA, "irec_get_ndtri_table: copy failed");
    return IREC_OK;
}

// ---------------------------------------------------------------------------------------------
// random.Random(seed).randint(0, 2**31 - 1): MT19937 + init_by_array + getrandbits(32) rejection
// (CPython Modules/_randommodule.c; Lib/random.py).  TF uses it for the op seed of unseeded ops.
// ---------------------------------------------------------------------------------------------
struct MT {
    uint32_t mt[624];
    int idx;
    void init_genrand(uint32_t s)
    {
        mt[0] = s;
        for (int i = 1; i < 624; ++i) mt[i] = 1812433253u * (mt[i - 1] ^ (mt[i - 1] >> 30)) + (uint32_t)i;
        idx = 624;
    }
    void init_by_array(const uint32_t* key, int len)
    {
        init_genrand(19650218u);
        int i = 1, j = 0;
        for (int k = (624 > len ? 624 : len); k; --k) {
            mt[i] = (mt[i] ^ ((mt[i - 1] ^ (mt[i - 1] >> 30)) * 1664525u)) + key[j] + (uint32_t)j;
            if (++i >= 624) { mt[0] = mt[623]; i = 1; }
            if (++j >= len) j = 0;
        }
        for (int k = 623; k; --k) {
            mt[i] = (mt[i] ^ ((mt[i - 1] ^ (mt[i - 1] >> 30)) * 1566083941u)) - (uint32_t)i;
            if (++i >= 624) { mt[0] = mt[623]; i = 1; }
        }
        mt[0] = 0x80000000u;
    }
    uint32_t next()
    {
        if (idx >= 624) {
            for (int kk = 0; kk < 624; ++kk) {
                const uint32_t y = (mt[kk] & 0x80000000u) | (mt[(kk + 1) % 624] & 0x7fffffffu);
                mt[kk] = mt[(kk + 397) % 624] ^ (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
            }
            idx = 0;
        }
        uint32_t y = mt[idx++];
        y ^= (y >> 11);
        y ^= (y << 7) & 0x9d2c5680u;
        y ^= (y << 15) & 0xefc60000u;
        y ^= (y >> 18);
        return y;
    }
};

int64_t irec_tf_op_seed(int64_t seed)
{
    const uint64_t a = (uint64_t)(seed < 0 ? -seed : seed);
    const uint32_t key[2] = { (uint32_t)a, (uint32_t)(a >> 32) };
    MT m;
    m.init_by_array(key, key[1] ? 2 : 1);
    for (;;) {
        const uint32_t r = m.next();
        if (r < 0x80000000u) return (int64_t)r;
    }
}

// host Philox (same rounds as the device one) for the sequential Fisher-Yates of Coder.split
static void philox_host(uint32_t k0, uint32_t k1, uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t out[4])
{
    for (int r = 0; r < 10; ++r) {
        const uint64_t p0 = (uint64_t)0xD2511F53u * c0, p1 = (uint64_t)0xCD9E8D57u * c2;
        const uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0, n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
        c0 = n0; c1 = (uint32_t)p1; c2 = n2; c3 = (uint32_t)p0;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

int irec_split_permutation(int64_t n, int64_t seed, int64_t* perm)
{
    // rec/coding/coder.py:60-67; TF core/kernels/random_shuffle_op.cc: forward Fisher-Yates,
    // one uint32 per step from the op's Philox stream (unseeded op after set_seed(seed))
    if (n < 0 || !perm) return irec_fail(IREC_E_INVALID, "split_permutation: bad arguments");
    if (n >= (1LL << 32)) return irec_fail(IREC_E_CAPACITY, "split_permutation: n must be < 2^32");
    const TfStream st = tf_stream_seeded(seed, irec_tf_op_seed(seed));
    for (int64_t i = 0; i < n; ++i) perm[i] = i;
    uint32_t buf[4];
    uint64_t j = 0;
    for (int64_t i = 0; i + 1 < n; ++i, ++j) {
        if ((j & 3) == 0) philox_host(st.k0, st.k1, (uint32_t)(j >> 2), (uint32_t)((j >> 2) >> 32), st.c2, st.c3, buf);
        const int64_t k = i + (int64_t)(buf[j & 3] % (uint32_t)(n - i));
        const int64_t tmp = perm[i]; perm[i] = perm[k]; perm[k] = tmp;
    }
    return IREC_OK;
}

}  // extern "C"
