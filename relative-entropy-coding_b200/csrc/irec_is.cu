// irec_is.cu -- importance-sampling sampler and the GaussianCoder auxiliary-variable loop (sm_100a).
//
// Reference path being replaced: rec/coding/importance_sampling.py:9-103 (alpha = inf branch),
// rec/coding/samplers.py:61-101 (ImportanceSampler) and rec/coding/coder.py:493-584
// (GaussianCoder.encode_block / decode_block with the closed-form conditioning of coder.py:141-171).
//
// Candidates are the float32 N(0,1) stream of tf.random.normal (Philox4x32-10 + Box-Muller), element
// j = s*D + d.  Canonical score of sample s:  sum_d (A d + E) d,  d = z - mu',  A = 0.5 (1 - 1/sigma'^2),
// E = mu'  (= log N(z; mu', sigma') - log N(z; 0, 1) up to a constant), 32-dim chunk sums combined by a
// pairwise tree -- the same reduction contract as the beam path.
#include <string.h>
#include "irec_boxmuller.cuh"
// device copies of the Box-Muller tables (filled once per device by bm_ensure_tables)
__device__ double2 g_bm_logA[128];
__device__ double2 g_bm_logB[128];
__device__ double2 g_bm_sc[257];
#ifndef IREC_BM_LIBM            // -DIREC_BM_LIBM: the libm-grade log()/sincos() path (A/B runs)
#define IREC_BM_TABLES (BmTables{ g_bm_logA, g_bm_logB, g_bm_sc })
#endif
#include "irec_beam.cuh"
#include "irec_host.h"
#include <mutex>
#include <unordered_map>
#include <vector>

// per-lane chunk sum: 32 dims starting at chunk's first dim; A4/M4 in CI layout (q0 = first quad index)
__device__ __forceinline__ float is_score_chunk(const float4* __restrict__ A4, const float4* __restrict__ M4, int P, int q0,
                                                const TfStream& st, uint64_t j_base)
{
    float acc = 0.f;
    const bool aligned = (j_base & 3) == 0;
#ifndef IREC_IS_UNROLL
#define IREC_IS_UNROLL 1
#endif
#define IREC_PRAGMA_(x) _Pragma(#x)
#define IREC_UNROLL_(n) IREC_PRAGMA_(unroll n)
    IREC_UNROLL_(IREC_IS_UNROLL)
    for (int iq = 0; iq < 8; ++iq) {
        float4 z;
        if (aligned) {
            z = tf_normal_group(st, (j_base >> 2) + iq);
        } else {
            const uint64_t j = j_base + 4 * iq;
            z.x = tf_normal_elem(st, j); z.y = tf_normal_elem(st, j + 1);
            z.z = tf_normal_elem(st, j + 2); z.w = tf_normal_elem(st, j + 3);
        }
        const float4 A = A4[q0 + iq * P], M = M4[q0 + iq * P];
        float d, t;
        d = __fadd_rn(z.x, -M.x); t = __fmaf_rn(A.x, d, M.x); acc = __fmaf_rn(t, d, acc);
        d = __fadd_rn(z.y, -M.y); t = __fmaf_rn(A.y, d, M.y); acc = __fmaf_rn(t, d, acc);
        d = __fadd_rn(z.z, -M.z); t = __fmaf_rn(A.z, d, M.z); acc = __fmaf_rn(t, d, acc);
        d = __fadd_rn(z.w, -M.w); t = __fmaf_rn(A.w, d, M.w); acc = __fmaf_rn(t, d, acc);
    }
    return acc;
}

// canonical score of sample s for the P lanes of its group (all lanes of the warp must call)
__device__ __forceinline__ float is_score_sample(const float4* A4, const float4* M4, const BeamGeom& g, int lg,
                                                 const TfStream& st, uint64_t s)
{
    if (g.nslots == 1) {
        float acc = is_score_chunk(A4, M4, g.P, lg, st, s * (uint64_t)g.D + (uint64_t)(32 * lg));
        for (int stride = 1; stride < g.P; stride <<= 1) acc = __fadd_rn(acc, __shfl_xor_sync(0xffffffffu, acc, stride));
        return acc;
    }
    float stack[12];
    float acc = 0.f;
    for (int m = 0; m < g.nslots; ++m) {
        acc = is_score_chunk(A4, M4, 32, m * 256 + lg, st, s * (uint64_t)g.D + (uint64_t)(1024 * m + 32 * lg));
        for (int stride = 1; stride < 32; stride <<= 1) acc = __fadd_rn(acc, __shfl_xor_sync(0xffffffffu, acc, stride));
        int lvl = 0;
        while ((m >> lvl) & 1) { acc = __fadd_rn(stack[lvl], acc); ++lvl; }
        stack[lvl] = acc;
    }
    bool have = false;
    for (int lvl = 0; lvl < 12; ++lvl)
        if ((g.nslots >> lvl) & 1) { acc = have ? __fadd_rn(stack[lvl], acc) : stack[lvl]; have = true; }
    return acc;
}

// standardised target -> canonical coefficients.  mu' = (tl - pl)/ps, sigma' = ts/ps  (importance_sampling.py:41-42)
__device__ __forceinline__ void is_coeffs(float tl, float ts, float pl, float ps, float& A, float& M)
{
    const float mu = __fdiv_rn(__fadd_rn(tl, -pl), ps);
    const float sg = __fdiv_rn(ts, ps);
    const double s2 = __dmul_rn((double)sg, (double)sg);
    A = (float)__dmul_rn(0.5, __dsub_rn(1.0, __ddiv_rn(1.0, s2)));
    M = mu;
}

// CTA-wide arg-best of (score desc, s asc); every thread passes its own best; result broadcast
__device__ __forceinline__ void block_argbest(float& v, int64_t& s, float* s_v, int64_t* s_s)
{
    for (int stride = 16; stride >= 1; stride >>= 1) {
        const float ov = __shfl_xor_sync(0xffffffffu, v, stride);
        const int64_t os = __shfl_xor_sync(0xffffffffu, s, stride);
        if (cand_better(ov, os, v, s)) { v = ov; s = os; }
    }
    const int warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
    __syncthreads();
    if ((threadIdx.x & 31) == 0) { s_v[warp] = v; s_s[warp] = s; }
    __syncthreads();
    v = s_v[0]; s = s_s[0];
    for (int w = 1; w < nw; ++w)
        if (cand_better(s_v[w], s_s[w], v, s)) { v = s_v[w]; s = s_s[w]; }
}

#define IS_NEG_INF __int_as_float(0xff800000)
#define IS_NO_SAMPLE 0x7fffffffffffffffLL

// ---------------------------------------------------------------------------------------------
// single partition, grid-wide (ImportanceSampler.coded_sample, samplers.py:74-84)
// workspace: A[DP], M[DP] (CI layout), per-CTA (score, s) records
// ---------------------------------------------------------------------------------------------
__global__ void k_is_params(const float* __restrict__ t_loc, const float* __restrict__ t_scale,
                            const float* __restrict__ p_loc, const float* __restrict__ p_scale, int D, int P, int DP,
                            float* A, float* M)
{
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < DP; i += gridDim.x * blockDim.x) { A[i] = 0.f; M[i] = 0.f; }
    // (same thread writes the real entry after the zero fill only if it owns it: use a second pass)
    __syncthreads();
    for (int d = blockIdx.x * blockDim.x + threadIdx.x; d < D; d += gridDim.x * blockDim.x) {
        float a, m;
        is_coeffs(t_loc[d], t_scale[d], p_loc[d], p_scale[d], a, m);
        const int ci = ci_index(d, P);
        A[ci] = a; M[ci] = m;
    }
}

__global__ void __launch_bounds__(256) k_is_score_grid(const float* __restrict__ A, const float* __restrict__ M, int D,
                                                       int64_t S, TfStream st, float* out_v, int64_t* out_s)
{
    __shared__ float s_v[8];
    __shared__ int64_t s_s[8];
    const BeamGeom g = make_geom(D);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
    const int lg = lane & (g.P - 1);
    const float4* A4 = reinterpret_cast<const float4*>(A);
    const float4* M4 = reinterpret_cast<const float4*>(M);
    const int64_t nsg = (S + g.SPW - 1) / g.SPW;
    float bv = IS_NEG_INF;
    int64_t bs = IS_NO_SAMPLE;
    for (int64_t sg = (int64_t)blockIdx.x * nw + warp; sg < nsg; sg += (int64_t)gridDim.x * nw) {
        const int64_t s = sg * g.SPW + lane / g.P;
        const bool valid = s < S;
        float v = is_score_sample(A4, M4, g, lg, st, (uint64_t)(valid ? s : 0));
        v = (v == v) ? v : IS_NEG_INF;
        if (valid && cand_better(v, s, bv, bs)) { bv = v; bs = s; }
    }
    block_argbest(bv, bs, s_v, s_s);
    if (threadIdx.x == 0) { out_v[blockIdx.x] = bv; out_s[blockIdx.x] = bs; }
}

// reduce the per-CTA records, emit index and sample = p_scale * z[idx] + p_loc (importance_sampling.py:74-77)
__global__ void __launch_bounds__(256) k_is_finish(const float* __restrict__ rec_v, const int64_t* __restrict__ rec_s, int nrec,
                                                   const float* __restrict__ p_loc, const float* __restrict__ p_scale, int D,
                                                   TfStream st, int64_t* out_index, float* out_sample)
{
    __shared__ float s_v[8];
    __shared__ int64_t s_s[8];
    float bv = IS_NEG_INF;
    int64_t bs = IS_NO_SAMPLE;
    for (int i = threadIdx.x; i < nrec; i += blockDim.x)
        if (cand_better(rec_v[i], rec_s[i], bv, bs)) { bv = rec_v[i]; bs = rec_s[i]; }
    block_argbest(bv, bs, s_v, s_s);
    if (bs == IS_NO_SAMPLE) bs = 0;
    if (threadIdx.x == 0) *out_index = bs;
    for (int d = threadIdx.x; d < D; d += blockDim.x) {
        const float z = tf_normal_elem(st, (uint64_t)bs * (uint64_t)D + (uint64_t)d);
        out_sample[d] = __fadd_rn(__fmul_rn(p_scale[d], z), p_loc[d]);
    }
}

// decode_gaussian_importance_sample (importance_sampling.py:82-103)
__global__ void k_is_decode_sample(const float* __restrict__ p_loc, const float* __restrict__ p_scale, int D,
                                   const int64_t* __restrict__ index, TfStream st, float* out_sample)
{
    const int64_t idx = *index;
    for (int d = blockIdx.x * blockDim.x + threadIdx.x; d < D; d += gridDim.x * blockDim.x) {
        const float z = tf_normal_elem(st, (uint64_t)idx * (uint64_t)D + (uint64_t)d);
        out_sample[d] = __fadd_rn(__fmul_rn(p_scale[d], z), p_loc[d]);
    }
}

__global__ void k_is_normal_stream(TfStream st, int64_t start, int64_t n, float* out)
{
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        out[i] = tf_normal_elem(st, (uint64_t)(start + i));
}

// ---------------------------------------------------------------------------------------------
// GaussianCoder.encode_block with an ImportanceSampler, one CTA per coder-block (coder.py:493-559).
// Running target/coder live in shared memory (plain layout), the per-partition coefficients in CI layout.
//
// Candidate table.  The candidates of partition k are z[s, d] = element s*D + d of the stream seeded with
// `seed + k` (importance_sampling.py:38,54; coder.py:523,538: the seed does not depend on the coder-block), so every
// block of a launch with the same size D scores the SAME S x D standard normals.  k_is_ztab evaluates them ONCE per
// launch ([size][k][s][DP] float32, rows in CI layout) and the blocks stream them from L2 -- one coalesced 16-byte load
// per quad instead of a Philox4x32-10 group and two float64-accurate Box-Muller pairs (~100 lane-instructions per
// candidate-dim, profiles/r1_is_block_b_ncu.md).  Sizes without a table (a third distinct block size, or a table
// beyond IS_TAB_MAX_BYTES) fall back to generating in place; both give the same bits.
// ---------------------------------------------------------------------------------------------
#define IS_MAX_D 4096
#define IS_MAX_SIZES 2
#define IS_TAB_MAX_BYTES ((size_t)512 << 20)
struct IsPlan {
    int32_t n_sizes;
    int32_t D[IS_MAX_SIZES];
    int32_t pad;
};
struct IsBlockArgs {
    const float* t_loc; const float* t_scale; const float* p_loc; const float* p_scale;
    const int64_t* gidx; const int64_t* offs; int nb;
    float omega; int64_t S; int64_t seed;
    const TfStream* streams;        // [max_aux] stream of coding seed `seed + k`
    const int64_t* in_indices;      // decode only
    const int32_t* in_n_idx;        // decode only
    int64_t* out_indices; int max_aux; int32_t* out_n_idx; int32_t* out_status; float* out_sample;
    const float* ratio_tab; int ratio_len;
    int DPmax;
    const IsPlan* plan;             // encode: distinct block sizes with a candidate table (nullptr: none)
    const float* ztab;              // [IS_MAX_SIZES][max_aux][S][DPmax]
};

// distinct block sizes of the launch (single CTA; plan zeroed by the host)
__global__ void __launch_bounds__(1024) k_is_plan(const int64_t* __restrict__ offs, int nb, IsPlan* plan)
{
    for (int b = threadIdx.x; b < nb; b += blockDim.x) {
        const int64_t D64 = offs[b + 1] - offs[b];
        if (D64 <= 0 || D64 > IS_MAX_D) continue;
        const int D = (int)D64;
        for (int k = 0; k < IS_MAX_SIZES; ++k) {
            const int old = atomicCAS(&plan->D[k], 0, D);
            if (old == 0 || old == D) break;
        }
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        int n = 0;
        for (int k = 0; k < IS_MAX_SIZES; ++k) n += plan->D[k] != 0;
        plan->n_sizes = n;
    }
}

// one thread per (size, partition, sample, physical quad of the CI row)
__global__ void __launch_bounds__(256) k_is_ztab(const IsPlan* __restrict__ plan, const TfStream* __restrict__ streams, int S,
                                                 int max_aux, int DPmax, float* __restrict__ tab)
{
    const int nq = DPmax >> 2;
    const int64_t per_size = (int64_t)max_aux * S * nq;
    const int64_t total = per_size * IS_MAX_SIZES;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int k = (int)(i / per_size);
        const int D = plan->D[k];
        if (D == 0) continue;
        int64_t r = i - (int64_t)k * per_size;
        const int q = (int)(r % nq); r /= nq;
        const int s = (int)(r % S);
        const int t = (int)(r / S);
        const BeamGeom g = make_geom(D);
        if (q >= (g.DP >> 2)) continue;
        const int slot = q / (8 * g.P), qs = q - slot * 8 * g.P;        // CI(P): quad qs of a slot = iqd * P + l
        const int iqd = qs / g.P, l = qs - iqd * g.P;
        const int d0 = slot * 32 * g.P + 32 * l + 4 * iqd;
        float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
        if (d0 < D) {
            const TfStream st = streams[t];
            const uint64_t j0 = (uint64_t)s * (uint64_t)D + (uint64_t)d0;
            if ((j0 & 3) == 0) {
                z = tf_normal_group(st, j0 >> 2);
            } else {
                z.x = tf_normal_elem(st, j0); z.y = tf_normal_elem(st, j0 + 1);
                z.z = tf_normal_elem(st, j0 + 2); z.w = tf_normal_elem(st, j0 + 3);
            }
            if (d0 + 1 >= D) z.y = 0.f;      // dims beyond D carry A = M = 0 and add exactly 0 to a score either way
            if (d0 + 2 >= D) z.z = 0.f;
            if (d0 + 3 >= D) z.w = 0.f;
        }
        reinterpret_cast<float4*>(tab)[i] = z;
    }
}

// table version of is_score_chunk / is_score_sample: zrow = the sample's CI row
__device__ __forceinline__ float is_score_chunk_tab(const float4* __restrict__ A4, const float4* __restrict__ M4, int P, int q0,
                                                    const float4* __restrict__ zrow)
{
    float acc = 0.f;
    float4 z[8];
#pragma unroll
    for (int iq = 0; iq < 8; ++iq) z[iq] = __ldg(zrow + q0 + iq * P);
#pragma unroll
    for (int iq = 0; iq < 8; ++iq) {
        const float4 A = A4[q0 + iq * P], M = M4[q0 + iq * P];
        float d, t;
        d = __fadd_rn(z[iq].x, -M.x); t = __fmaf_rn(A.x, d, M.x); acc = __fmaf_rn(t, d, acc);
        d = __fadd_rn(z[iq].y, -M.y); t = __fmaf_rn(A.y, d, M.y); acc = __fmaf_rn(t, d, acc);
        d = __fadd_rn(z[iq].z, -M.z); t = __fmaf_rn(A.z, d, M.z); acc = __fmaf_rn(t, d, acc);
        d = __fadd_rn(z[iq].w, -M.w); t = __fmaf_rn(A.w, d, M.w); acc = __fmaf_rn(t, d, acc);
    }
    return acc;
}

__device__ __forceinline__ float is_score_sample_tab(const float4* A4, const float4* M4, const BeamGeom& g, int lg,
                                                     const float4* __restrict__ zrow)
{
    if (g.nslots == 1) {
        float acc = is_score_chunk_tab(A4, M4, g.P, lg, zrow);
        for (int stride = 1; stride < g.P; stride <<= 1) acc = __fadd_rn(acc, __shfl_xor_sync(0xffffffffu, acc, stride));
        return acc;
    }
    float stack[12];
    float acc = 0.f;
    for (int m = 0; m < g.nslots; ++m) {
        acc = is_score_chunk_tab(A4, M4, 32, m * 256 + lg, zrow);
        for (int stride = 1; stride < 32; stride <<= 1) acc = __fadd_rn(acc, __shfl_xor_sync(0xffffffffu, acc, stride));
        int lvl = 0;
        while ((m >> lvl) & 1) { acc = __fadd_rn(stack[lvl], acc); ++lvl; }
        stack[lvl] = acc;
    }
    bool have = false;
    for (int lvl = 0; lvl < 12; ++lvl)
        if ((g.nslots >> lvl) & 1) { acc = have ? __fadd_rn(stack[lvl], acc) : stack[lvl]; have = true; }
    return acc;
}

__global__ void __launch_bounds__(256, 3) k_is_block(const IsBlockArgs a)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ double s_kl[IS_MAX_D / 32];
    __shared__ float s_bv[8];
    __shared__ int64_t s_bs[8];
    const int tid = threadIdx.x, nt = blockDim.x, lane = tid & 31, warp = tid >> 5, nw = nt >> 5;
    const int DPm = a.DPmax;
    float* s_tl = reinterpret_cast<float*>(smem_raw);
    float* s_ts = s_tl + DPm; float* s_pl = s_ts + DPm; float* s_ps = s_pl + DPm;
    float* s_A = s_ps + DPm; float* s_M = s_A + DPm; float* s_sv = s_M + DPm; float* s_v = s_sv + DPm;

    for (int blk = blockIdx.x; blk < a.nb; blk += gridDim.x) {
        __syncthreads();
        const int64_t off = a.offs[blk];
        const int D = (int)(a.offs[blk + 1] - off);
        const BeamGeom g = make_geom(D);
        const int lg = lane & (g.P - 1);
        const float* ztab_blk = nullptr;            // candidate table of this block size (if it has one)
        if (a.ztab) {
#pragma unroll
            for (int k = 0; k < IS_MAX_SIZES; ++k)
                if (a.plan->D[k] == D) ztab_blk = a.ztab + (size_t)k * a.max_aux * a.S * DPm;
        }
        for (int i = tid; i < g.DP; i += nt) { s_A[i] = 0.f; s_M[i] = 0.f; }
        for (int d = tid; d < D; d += nt) {
            const int64_t gi = a.gidx ? a.gidx[off + d] : off + d;
            s_pl[d] = a.p_loc[gi]; s_ps[d] = a.p_scale[gi];
            s_tl[d] = a.t_loc[gi]; s_ts[d] = a.t_scale[gi];
        }
        __syncthreads();
        // KL and n_aux (coder.py:499-501), canonical float64 chunk/tree sum
        for (int c = tid; c < g.nch; c += nt) {
            double acc = 0.0;
            const int hi = min(D, 32 * c + 32);
            for (int d = 32 * c; d < hi; ++d) acc = __dadd_rn(acc, kl_dim(s_tl[d], s_ts[d], s_pl[d], s_ps[d]));
            s_kl[c] = acc;
        }
        const float klf = (float)block_tree_sum_f64(s_kl, g.nch);
        const int n_aux = n_aux_from_kl(klf, a.omega);
        int status = IREC_BLK_OK;
        if (n_aux < 0) status = IREC_BLK_BAD_KL;
        const int n_idx = n_aux > 1 ? n_aux : 1;
        if (status == IREC_BLK_OK && (n_idx > a.max_aux || n_idx > a.ratio_len)) status = IREC_BLK_TOO_LONG;
        if (tid == 0) { a.out_n_idx[blk] = n_idx; a.out_status[blk] = status; }
        if (status != IREC_BLK_OK) continue;
        int64_t* out_idx = a.out_indices + (size_t)blk * a.max_aux;

        // partitions k = 0 .. n_idx-1 ; k < n_idx-1 are auxiliary variables i = n_idx-1-k, the last is final
        for (int k = 0; k < n_idx; ++k) {
            const bool final_part = (k == n_idx - 1);
            const TfStream st = a.streams[k];
            const float* ztab_k = ztab_blk ? ztab_blk + (size_t)k * a.S * DPm : nullptr;
            // --- parameters of this partition's (target, coder) pair ---
            const float ratio = final_part ? 0.f : a.ratio_tab[n_idx - 1 - k];
            for (int d = tid; d < D; d += nt) {
                const float ps = s_ps[d];
                float sv, v = 0.f;
                float tl_e, ts_e, pl_e;
                if (!final_part) {
                    // get_auxiliary_coder / get_auxiliary_target (coder.py:141-154), coder loc of the auxiliary coder is 0
                    const float cv = __fmul_rn(ps, ps), tv = __fmul_rn(s_ts[d], s_ts[d]);
                    v = __fmul_rn(ratio, cv);
                    sv = __fsqrt_rn(v);
                    tl_e = __fdiv_rn(__fmul_rn(__fadd_rn(s_tl[d], -s_pl[d]), v), cv);
                    const float var = __fadd_rn(__fdiv_rn(__fmul_rn(tv, __fmul_rn(v, v)), __fmul_rn(cv, cv)),
                                                __fdiv_rn(__fmul_rn(v, __fadd_rn(cv, -v)), cv));
                    ts_e = __fsqrt_rn(var);
                    pl_e = 0.f;
                } else {
                    sv = ps;
                    tl_e = s_tl[d]; ts_e = s_ts[d]; pl_e = s_pl[d];
                }
                s_sv[d] = sv; s_v[d] = v;
                float A, M;
                is_coeffs(tl_e, ts_e, pl_e, sv, A, M);
                const int ci = ci_index(d, g.P);
                s_A[ci] = A; s_M[ci] = M;
            }
            __syncthreads();
            // --- choose the index: arg max of the importance weights (importance_sampling.py:60-72, alpha = inf) ---
            const float4* A4 = reinterpret_cast<const float4*>(s_A);
            const float4* M4 = reinterpret_cast<const float4*>(s_M);
            const int64_t nsg = (a.S + g.SPW - 1) / g.SPW;
            float bv = IS_NEG_INF;
            int64_t bs = IS_NO_SAMPLE;
            for (int64_t sg = warp; sg < nsg; sg += nw) {
                const int64_t s = sg * g.SPW + lane / g.P;
                const bool valid = s < a.S;
                const int64_t sc = valid ? s : 0;
                float v = ztab_k ? is_score_sample_tab(A4, M4, g, lg, reinterpret_cast<const float4*>(ztab_k + (size_t)sc * DPm))
                                 : is_score_sample(A4, M4, g, lg, st, (uint64_t)sc);
                v = (v == v) ? v : IS_NEG_INF;
                if (valid && cand_better(v, s, bv, bs)) { bv = v; bs = s; }
            }
            block_argbest(bv, bs, s_bv, s_bs);
            const int64_t idx = (bs == IS_NO_SAMPLE) ? 0 : bs;
            if (tid == 0) out_idx[k] = idx;
            // --- sample of this partition and conditioning (coder.py:157-171, 533-540) ---
            const float* zrow = ztab_k ? ztab_k + (size_t)idx * DPm : nullptr;
            for (int d = tid; d < D; d += nt) {
                const float z = zrow ? __ldg(zrow + ci_index(d, g.P)) : tf_normal_elem(st, (uint64_t)idx * (uint64_t)D + (uint64_t)d);
                if (final_part) {
                    const int64_t gi = a.gidx ? a.gidx[off + d] : off + d;
                    a.out_sample[gi] = __fadd_rn(__fmul_rn(s_ps[d], z), s_pl[d]);
                } else {
                    const float av = __fadd_rn(__fmul_rn(s_sv[d], z), 0.f);
                    const float ps = s_ps[d], pl = s_pl[d], v = s_v[d];
                    const float cv = __fmul_rn(ps, ps);
                    const float tl = s_tl[d], ts = s_ts[d];
                    const float tv = __fmul_rn(ts, ts);
                    const float cmv = __fadd_rn(cv, -v);
                    const float num = __fadd_rn(__fmul_rn(__fmul_rn(av, tv), cv),
                                                __fmul_rn(__fmul_rn(__fadd_rn(tl, -pl), cmv), cv));
                    const float den = __fadd_rn(__fmul_rn(tv, v), __fmul_rn(cv, cmv));
                    s_tl[d] = __fadd_rn(pl, __fdiv_rn(num, den));
                    const float nvar = __fdiv_rn(__fmul_rn(__fmul_rn(tv, cv), cmv),
                                                 __fadd_rn(__fmul_rn(v, tv), __fmul_rn(cv, cmv)));
                    s_ts[d] = __fsqrt_rn(nvar);
                    s_pl[d] = __fadd_rn(pl, av);
                    s_ps[d] = __fsqrt_rn(__fadd_rn(cv, -v));
                }
            }
            __syncthreads();
        }
    }
}

// ---------------------------------------------------------------------------------------------
// GaussianCoder.decode_block with an ImportanceSampler (coder.py:561-584): replay of the chosen candidates.
// Every thread owns quads of dims and carries their running coder (loc, scale) through all partitions in
// registers -- O(n_idx * D) per block, no barrier, one Philox group + two Box-Muller pairs per quad and
// partition (k_is_block<false>, which this replaces, regenerated every element on its own: 4x the Philox
// work, and one CTA barrier per partition).  Blocks whose index count is out of range (< 1, > max_aux,
// > the ratio table: the reference raises IndexError / CodingError there) are filled with NaN and flagged.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_is_decode(const IsBlockArgs a)
{
    for (int blk = blockIdx.x; blk < a.nb; blk += gridDim.x) {
        const int64_t off = a.offs[blk];
        const int D = (int)(a.offs[blk + 1] - off);
        const int n_idx = a.in_n_idx[blk];
        const bool bad = n_idx < 1 || n_idx > a.max_aux || n_idx > a.ratio_len;
        if (threadIdx.x == 0 && a.out_status) a.out_status[blk] = bad ? IREC_BLK_TOO_LONG : IREC_BLK_OK;
        const int64_t* in_idx = a.in_indices + (size_t)blk * a.max_aux;
        const int nq = (D + 3) >> 2;
        for (int q = threadIdx.x; q < nq; q += blockDim.x) {
            const int d0 = 4 * q;
            float ps[4], pl[4];
            int64_t gi[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const int d = min(d0 + e, D - 1);
                gi[e] = a.gidx ? a.gidx[off + d] : off + d;
                ps[e] = a.p_scale[gi[e]]; pl[e] = a.p_loc[gi[e]];
            }
            if (bad) {
#pragma unroll
                for (int e = 0; e < 4; ++e)
                    if (d0 + e < D) a.out_sample[gi[e]] = __int_as_float(0x7fc00000);
                continue;
            }
            for (int k = 0; k < n_idx; ++k) {
                const TfStream st = a.streams[k];
                const uint64_t j0 = (uint64_t)in_idx[k] * (uint64_t)D + (uint64_t)d0;
                float z[4];
                if ((j0 & 3) == 0) {
                    const float4 zz = tf_normal_group(st, j0 >> 2);
                    z[0] = zz.x; z[1] = zz.y; z[2] = zz.z; z[3] = zz.w;
                } else {
#pragma unroll
                    for (int e = 0; e < 4; ++e) z[e] = tf_normal_elem(st, j0 + e);
                }
                if (k == n_idx - 1) {
#pragma unroll
                    for (int e = 0; e < 4; ++e)
                        if (d0 + e < D) a.out_sample[gi[e]] = __fadd_rn(__fmul_rn(ps[e], z[e]), pl[e]);
                } else {
                    const float ratio = a.ratio_tab[n_idx - 1 - k];
#pragma unroll
                    for (int e = 0; e < 4; ++e) {       // auxiliary sample and conditioning of the coder (coder.py:157-171, 575-580)
                        const float cv = __fmul_rn(ps[e], ps[e]);
                        const float v = __fmul_rn(ratio, cv);
                        const float av = __fadd_rn(__fmul_rn(__fsqrt_rn(v), z[e]), 0.f);
                        pl[e] = __fadd_rn(pl[e], av);
                        ps[e] = __fsqrt_rn(__fadd_rn(cv, -v));
                    }
                }
            }
        }
    }
}

// =============================================================================================
// host side
// =============================================================================================
// tf.random.set_seed(seed) followed by an unseeded op: the op seed is random.Random(seed).randint(0, 2**31-1) (a Mersenne
// twister initialisation per seed) -- memoised, a coder asks for the same seed + k sequence on every call
static TfStream is_stream_for_seed(int64_t seed)
{
    static std::mutex mu;
    static std::unordered_map<int64_t, int64_t> memo;
    int64_t op;
    {
        std::lock_guard<std::mutex> lock(mu);
        auto it = memo.find(seed);
        if (it != memo.end()) {
            op = it->second;
        } else {
            op = irec_tf_op_seed(seed);
            if (memo.size() > (1u << 20)) memo.clear();
            memo.emplace(seed, op);
        }
    }
    return tf_stream_seeded(seed, op);
}

// ---- Box-Muller tables (irec_boxmuller.cuh): built in long double on the host, one upload per device ----
static double2 h_bm_logA[128], h_bm_logB[128], h_bm_sc[257];
static std::once_flag bm_host_once;
static void bm_build_host_tables()
{
    for (int j = 0; j < 128; ++j) {
        const long double c = 1.0L + ((long double)j + 0.5L) / 128.0L;
        const double invA = (double)(1.0L / c), invB = (double)(2.0L / c);
        h_bm_logA[j] = make_double2(invA, (double)(-logl((long double)invA)));
        h_bm_logB[j] = make_double2(invB, (double)(-logl((long double)invB)));
    }
    h_bm_logB[127] = make_double2(1.0, 0.0);                 // arguments next to 1: r = m - 1 exactly, no table term
    const long double pi = 3.14159265358979323846264338327950288L;
    for (int j = 0; j <= 256; ++j) {
        const long double a = (long double)j * pi / 128.0L;
        h_bm_sc[j] = make_double2((double)sinl(a), (double)cosl(a));
    }
    const double q[5][2] = { { 0, 1 }, { 1, 0 }, { 0, -1 }, { -1, 0 }, { 0, 1 } };     // exact at multiples of pi/2
    for (int k = 0; k < 5; ++k) h_bm_sc[64 * k] = make_double2(q[k][0], q[k][1]);
}
static BmTables bm_host_tables()
{
    std::call_once(bm_host_once, bm_build_host_tables);
    return BmTables{ h_bm_logA, h_bm_logB, h_bm_sc };
}
static int bm_ensure_tables()
{
    static std::mutex mu;
    static bool done[64] = { false };
    int d = 0;
    if (cudaGetDevice(&d) != cudaSuccess || d < 0 || d >= 64) return irec_fail(IREC_E_CUDA, "bm_ensure_tables: no CUDA device");
    if (done[d]) return IREC_OK;
    std::lock_guard<std::mutex> lock(mu);
    if (done[d]) return IREC_OK;
    bm_host_tables();
    if (cudaMemcpyToSymbol(g_bm_logA, h_bm_logA, sizeof(h_bm_logA)) != cudaSuccess ||
        cudaMemcpyToSymbol(g_bm_logB, h_bm_logB, sizeof(h_bm_logB)) != cudaSuccess ||
        cudaMemcpyToSymbol(g_bm_sc, h_bm_sc, sizeof(h_bm_sc)) != cudaSuccess)
        return irec_fail(IREC_E_CUDA, "bm_ensure_tables: table upload failed");
    done[d] = true;
    return IREC_OK;
}
#define IREC_ENSURE_BM()                          \
    do {                                          \
        const int rc_bm_ = bm_ensure_tables();    \
        if (rc_bm_ != IREC_OK) return rc_bm_;     \
    } while (0)

extern "C" {

int irec_bm_components_host(const uint32_t* m, int64_t n, float* logf_out, float* sin_out, float* cos_out)
{
    // HOST evaluation of the device's table-driven functions (same IEEE operation sequence), for exhaustive checks:
    // logf_out[i] = bm_logf(Uint32ToFloat(m[i]) clamped), (sin_out, cos_out)[i] = bm_sincosf(float(2 pi Uint32ToFloat(m[i])))
    const BmTables t = bm_host_tables();
    for (int64_t i = 0; i < n; ++i) {
        const uint32_t b = 0x3F800000u | (m[i] & 0x7FFFFFu);
        float f;
        memcpy(&f, &b, 4);
        const float u = f - 1.0f;
        float u1 = u;
        if (u1 < 1.0e-7f) u1 = 1.0e-7f;
        if (logf_out) logf_out[i] = bm_logf(u1, t);
        const float v1 = (float)(6.283185307179586 * (double)u);
        float s, c;
        bm_sincosf(v1, t, s, c);
        if (sin_out) sin_out[i] = s;
        if (cos_out) cos_out[i] = c;
    }
    return IREC_OK;
}

int irec_is_normal_stream(int64_t seed, int64_t start, int64_t n, float* out, void* stream)
{
    IREC_ENSURE_INIT();
    IREC_ENSURE_BM();
    if (n <= 0) return IREC_OK;
    k_is_normal_stream<<<(int)std::min<int64_t>((n + 255) / 256, 2048), 256, 0, (cudaStream_t)stream>>>(is_stream_for_seed(seed), start, n, out);
    irec_count_launch();
    return irec_check_launch("k_is_normal_stream");
}

int irec_normal_stream_seeded(int64_t global_seed, int64_t op_seed, int64_t start, int64_t n, float* out, void* stream)
{
    IREC_ENSURE_INIT();
    IREC_ENSURE_BM();
    if (n <= 0) return IREC_OK;
    k_is_normal_stream<<<(int)std::min<int64_t>((n + 255) / 256, 2048), 256, 0, (cudaStream_t)stream>>>(
        tf_stream_seeded(global_seed, op_seed), start, n, out);
    irec_count_launch();
    return irec_check_launch("k_is_normal_stream");
}

size_t irec_is_workspace_bytes(int D)
{
    if (irec_init() != IREC_OK) return 0;
    const BeamGeom g = make_geom(D);
    const int grid_max = irec_device().sm_count * 8;
    return sizeof(float) * 2 * (size_t)g.DP + (sizeof(float) + sizeof(int64_t)) * (size_t)grid_max + 256;
}

int irec_is_coded_sample(const float* t_loc, const float* t_scale, const float* p_loc, const float* p_scale,
                         int D, int64_t S, int64_t seed, int64_t* out_index, float* out_sample,
                         void* workspace, size_t workspace_bytes, void* stream)
{
    IREC_ENSURE_INIT();
    IREC_ENSURE_BM();
    cudaStream_t s = (cudaStream_t)stream;
    if (D <= 0 || S <= 0) return irec_fail(IREC_E_INVALID, "is_coded_sample: bad sizes");
    if (workspace_bytes < irec_is_workspace_bytes(D)) return irec_fail(IREC_E_CAPACITY, "is_coded_sample: workspace too small");
    const BeamGeom g = make_geom(D);
    const int grid_max = irec_device().sm_count * 8;
    unsigned char* w = reinterpret_cast<unsigned char*>(workspace);
    int64_t* rec_s = reinterpret_cast<int64_t*>(w);
    float* A = reinterpret_cast<float*>(rec_s + grid_max + 8);
    float* M = A + g.DP;
    float* rec_v = M + g.DP;
    const TfStream st = is_stream_for_seed(seed);
    k_is_params<<<1, 1024, 0, s>>>(t_loc, t_scale, p_loc, p_scale, D, g.P, g.DP, A, M);
    irec_count_launch();
    const int64_t nsg = (S + g.SPW - 1) / g.SPW;
    const int grid = (int)std::max<int64_t>(1, std::min<int64_t>((nsg + 7) / 8, grid_max));
    k_is_score_grid<<<grid, 256, 0, s>>>(A, M, D, S, st, rec_v, rec_s);
    irec_count_launch();
    k_is_finish<<<1, 256, 0, s>>>(rec_v, rec_s, grid, p_loc, p_scale, D, st, out_index, out_sample);
    irec_count_launch();
    return irec_check_launch("irec_is_coded_sample");
}

int irec_is_decode_sample(const float* p_loc, const float* p_scale, int D, const int64_t* index, int64_t seed,
                          float* out_sample, void* stream)
{
    IREC_ENSURE_INIT();
    IREC_ENSURE_BM();
    if (D <= 0) return irec_fail(IREC_E_INVALID, "is_decode_sample: bad sizes");
    k_is_decode_sample<<<std::max(1, std::min((D + 255) / 256, 256)), 256, 0, (cudaStream_t)stream>>>(
        p_loc, p_scale, D, index, is_stream_for_seed(seed), out_sample);
    irec_count_launch();
    return irec_check_launch("k_is_decode_sample");
}

size_t irec_is_block_workspace_bytes(int max_aux) { return sizeof(TfStream) * (size_t)std::max(max_aux, 1) + 256; }

// candidate table of an encode launch: [IS_MAX_SIZES][max_aux][S][DP] float32, or 0 when it would be too large
static size_t is_ztab_bytes(int64_t max_block_dim, int64_t S, int max_aux)
{
    if (max_block_dim <= 0 || max_block_dim > IS_MAX_D || S <= 0 || max_aux <= 0) return 0;
    const char* e = getenv("IREC_IS_NO_TABLE");       // tests / A-B runs: candidates generated in place
    if (e && e[0] == '1') return 0;
    const BeamGeom g = make_geom((int)max_block_dim);
    const double b = (double)sizeof(float) * IS_MAX_SIZES * (double)max_aux * (double)S * (double)g.DP;
    return b <= (double)IS_TAB_MAX_BYTES ? (size_t)b : 0;
}
// workspace of irec_is_encode: stream table | IsPlan | candidate table
size_t irec_is_encode_workspace_bytes(int nb, int64_t max_block_dim, int64_t S, int max_aux)
{
    (void)nb;
    return irec_is_block_workspace_bytes(max_aux) + 256 + is_ztab_bytes(max_block_dim, S, max_aux);
}

static int is_upload_streams(int max_aux, int64_t seed, void* workspace, cudaStream_t s)
{
    std::vector<TfStream> streams((size_t)max_aux);
    for (int k = 0; k < max_aux; ++k) streams[k] = is_stream_for_seed(seed + k);
    // pageable source: cudaMemcpyAsync returns once the buffer has been staged, so `streams` may go out of scope and the
    // copy is ordered before the kernels on `s` -- no host synchronisation
    if (cudaMemcpyAsync(workspace, streams.data(), sizeof(TfStream) * (size_t)max_aux, cudaMemcpyHostToDevice, s) != cudaSuccess)
        return irec_fail(IREC_E_CUDA, "is block: stream table upload failed");
    return IREC_OK;
}

int irec_is_encode(const float* t_loc, const float* t_scale, const float* p_loc, const float* p_scale,
                   const int64_t* gather_idx, const int64_t* block_offsets, int nb, int64_t max_block_dim,
                   float omega, int64_t S, int64_t seed,
                   int64_t* out_indices, int max_aux, int32_t* out_n_idx, int32_t* out_status,
                   float* out_sample, void* workspace, size_t workspace_bytes, void* stream)
{
    IREC_ENSURE_INIT();
    IREC_ENSURE_BM();
    if (nb <= 0) return IREC_OK;
    if (S <= 0 || !(omega > 0.f)) return irec_fail(IREC_E_INVALID, "is_encode: bad arguments");
    cudaStream_t s = (cudaStream_t)stream;
    if (max_block_dim <= 0 || max_block_dim > IS_MAX_D)
        return irec_fail(IREC_E_CAPACITY, "importance-sampler blocks support at most 4096 dims per block (use block_size)");
    if (max_aux <= 0) return irec_fail(IREC_E_INVALID, "is_encode: max_aux must be > 0");
    if (workspace_bytes < irec_is_block_workspace_bytes(max_aux)) return irec_fail(IREC_E_CAPACITY, "is_encode: workspace too small");
    IsBlockArgs a{};
    a.t_loc = t_loc; a.t_scale = t_scale; a.p_loc = p_loc; a.p_scale = p_scale; a.gidx = gather_idx; a.offs = block_offsets;
    a.nb = nb; a.omega = omega; a.S = S; a.seed = seed; a.out_indices = out_indices; a.max_aux = max_aux;
    a.out_n_idx = out_n_idx; a.out_status = out_status; a.out_sample = out_sample;
    const int rc = is_upload_streams(max_aux, seed, workspace, s);
    if (rc != IREC_OK) return rc;
    a.streams = reinterpret_cast<const TfStream*>(workspace);
    a.ratio_tab = irec_ratio_tab(); a.ratio_len = irec_ratio_len();
    const BeamGeom g = make_geom((int)max_block_dim);
    a.DPmax = g.DP;
    // candidate table, when the caller's workspace has room for it (irec_is_encode_workspace_bytes)
    unsigned char* w = reinterpret_cast<unsigned char*>(workspace);
    const size_t head = irec_is_block_workspace_bytes(max_aux);
    const size_t tab_bytes = is_ztab_bytes(max_block_dim, S, max_aux);
    if (tab_bytes && workspace_bytes >= head + 256 + tab_bytes) {
        IsPlan* dplan = reinterpret_cast<IsPlan*>(w + head);
        float* tab = reinterpret_cast<float*>(w + head + 256);
        if (cudaMemsetAsync(dplan, 0, sizeof(IsPlan), s) != cudaSuccess) return irec_fail(IREC_E_CUDA, "is_encode: memset failed");
        k_is_plan<<<1, 1024, 0, s>>>(block_offsets, nb, dplan);
        irec_count_launch();
        const int64_t items = (int64_t)IS_MAX_SIZES * max_aux * S * (g.DP >> 2);
        const int grid = (int)std::min<int64_t>((items + 255) / 256, (int64_t)irec_device().sm_count * 16);
        k_is_ztab<<<grid, 256, 0, s>>>(dplan, a.streams, (int)S, max_aux, g.DP, tab);
        irec_count_launch();
        a.plan = dplan; a.ztab = tab;
    }
    const size_t smem = sizeof(float) * 8 * (size_t)g.DP;
    // 3 CTAs of 8 warps per SM (launch bounds)
    const int grid = std::min(a.nb, irec_device().sm_count * (smem * 3 <= (size_t)200 * 1024 ? 3 : 2));
    static bool attr = false;
    if (!attr) { cudaFuncSetAttribute(k_is_block, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(sizeof(float) * 8 * IS_MAX_D)); attr = true; }
    k_is_block<<<grid, 256, smem, s>>>(a);
    irec_count_launch();
    return irec_check_launch("k_is_block");
}

int irec_is_decode(const float* p_loc, const float* p_scale, const int64_t* gather_idx,
                   const int64_t* block_offsets, int nb, int64_t max_block_dim, int64_t seed,
                   const int64_t* indices, int max_aux, const int32_t* n_idx,
                   float* out_sample, int32_t* out_status, void* workspace, size_t workspace_bytes, void* stream)
{
    IREC_ENSURE_INIT();
    IREC_ENSURE_BM();
    if (nb <= 0) return IREC_OK;
    cudaStream_t s = (cudaStream_t)stream;
    if (max_block_dim <= 0) return irec_fail(IREC_E_INVALID, "is_decode: bad sizes");
    if (max_aux <= 0) return irec_fail(IREC_E_INVALID, "is_decode: max_aux must be > 0");
    if (workspace_bytes < irec_is_block_workspace_bytes(max_aux)) return irec_fail(IREC_E_CAPACITY, "is_decode: workspace too small");
    IsBlockArgs a{};
    a.p_loc = p_loc; a.p_scale = p_scale; a.gidx = gather_idx; a.offs = block_offsets; a.nb = nb;
    a.seed = seed; a.in_indices = indices; a.in_n_idx = n_idx; a.max_aux = max_aux; a.out_sample = out_sample;
    a.out_status = out_status;
    const int rc = is_upload_streams(max_aux, seed, workspace, s);
    if (rc != IREC_OK) return rc;
    a.streams = reinterpret_cast<const TfStream*>(workspace);
    a.ratio_tab = irec_ratio_tab(); a.ratio_len = irec_ratio_len();
    k_is_decode<<<std::min(a.nb, irec_device().sm_count * 8), 256, 0, s>>>(a);
    irec_count_launch();
    return irec_check_launch("k_is_decode");
}

}  // extern "C"
