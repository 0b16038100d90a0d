// irec_wide.cu -- beam-search encode for WIDE beams (32 < n_beams <= IREC_WIDE_MAX_BEAMS).
//
// The reference puts no limit on n_beams (rec/coding/beam_search_coder.py:15-30 keeps `n_beams` rows of the argsort,
// :95-101); its examples use 10 and 20, which is what the persistent kernels (irec_tmem.cu, irec_resident2.cuh,
// irec_beam.cu) are shaped for: beams in tensor / shared memory, top-B scratch of 32.  This kernel covers the rest of the
// range with the same arithmetic (per candidate-dim: k = (r h_b) mod 10007, z = T[k] sigma_aux, x = beam_b + z, centred
// quadratic, 32-dim sequential chunk sums combined by a pairwise tree -- score_sample of irec_beam.cuh, so the scores are
// the oracle's bit for bit), one persistent CTA per coder-block, with everything that scales with n_beams in global
// memory (L2-resident): both beam buffers [2][B][DP], the scores [S * B], the back-pointer history [max_aux][B].
//
//   * scoring: work items (sample group, page of WIDE_PB beam slots) spread over the warps; the Philox stream of a
//     sample is regenerated once per page (24 / WIDE_PB lane-instructions per candidate-dim instead of 24 / B);
//   * top-B (beam_search_coder.py:86-101: tf.argsort(DESCENDING) keeps (score desc, flat index asc)): every candidate gets
//     the 64-bit key (ordered score bits << 32 | ~flat index) -- all keys distinct -- the B-th largest key is found by an
//     8-pass radix select (256-bin shared-memory histograms), the B survivors are ranked by counting;
//   * winners re-materialised from the counter-based stream into the OTHER beam buffer (no in-place hazard), reference
//     order `beam + (T[k] * sigma)`.
#include <algorithm>
#include <cstdint>
#include <cuda_runtime.h>

#include "../../include/irec.h"
#include "irec_beam.cuh"
#include "irec_host.h"

#define WIDE_THREADS 512
#define WIDE_PB 8                   // beam slots per scoring page

struct WideArgs {
    const float* t_loc; const float* t_scale; const float* p_loc; const float* p_scale;
    const int64_t* gidx; const int64_t* offs; int nb;
    float omega; int S; int B; int Bp; int64_t seed;      // Bp: B padded to a multiple of WIDE_PB (row capacity of every per-beam array)
    int32_t* out_indices; int max_aux; int32_t* out_n_aux; int32_t* out_status; float* out_sample;
    const float* T; const float* ratio_tab; int ratio_len;
    int* work_counter;
    int DPmax;             // padded dims capacity (multiple of 32)
    float* g_beams;        // [grid][2][Bp][DPmax]
    float* g_scores;       // [grid][S * Bp]
    int2* g_hist;          // [grid][max_aux][Bp]
    int32_t* g_hsum;       // [grid][2][Bp]
    unsigned long long* g_keys;   // [grid][Bp] survivors of the radix select
};

__device__ __forceinline__ uint32_t wide_ordered(float v)
{
    const uint32_t u = __float_as_uint(v);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);          // larger float -> larger unsigned
}
__device__ __forceinline__ unsigned long long wide_key(float v, uint32_t flat)
{
    return ((unsigned long long)wide_ordered(v) << 32) | (unsigned long long)(0xffffffffu - flat);
}

// The K = min(B, n) best of the n scores, best first, by (score desc, flat index asc): out_id[rank] = flat index.
// Scratch: s_hist[256] (shared), s_ctl[4] (shared), g_keys[K] (global).  All threads of the CTA call it.
__device__ int wide_topk(const float* __restrict__ sc, int n, int K, int32_t* s_wid, unsigned long long* g_keys,
                         int32_t* s_hist, unsigned long long* s_sel)
{
    const int tid = threadIdx.x, nt = blockDim.x;
    const int Kout = K < n ? K : n;
    if (Kout <= 0) return 0;
    // ---- radix select of the Kout-th largest key, 8 bits per pass from the top ----
    unsigned long long prefix = 0ull;      // the selected high bits so far
    int want = Kout;                       // rank (1-based, from the top) of the wanted key among the keys matching `prefix`
    for (int pass = 0; pass < 8; ++pass) {
        const int shift = 56 - 8 * pass;
        for (int i = tid; i < 256; i += nt) s_hist[i] = 0;
        __syncthreads();
        for (int i = tid; i < n; i += nt) {
            const unsigned long long k = wide_key(sc[i], (uint32_t)i);
            if (pass == 0 || (k >> (shift + 8)) == (prefix >> (shift + 8))) atomicAdd(&s_hist[(int)((k >> shift) & 0xffull)], 1);
        }
        __syncthreads();
        if (tid == 0) {
            int acc = 0, bin = 255;
            for (; bin > 0; --bin) {
                if (acc + s_hist[bin] >= want) break;
                acc += s_hist[bin];
            }
            s_sel[0] = prefix | ((unsigned long long)bin << shift);
            s_sel[1] = (unsigned long long)(want - acc);
        }
        __syncthreads();
        prefix = s_sel[0];
        want = (int)s_sel[1];
        __syncthreads();
    }
    // `prefix` is now the Kout-th largest key; keys are distinct, so exactly Kout keys are >= prefix
    if (tid == 0) s_hist[0] = 0;
    __syncthreads();
    for (int i = tid; i < n; i += nt) {
        const unsigned long long k = wide_key(sc[i], (uint32_t)i);
        if (k >= prefix) g_keys[atomicAdd(&s_hist[0], 1)] = k;
    }
    __syncthreads();
    // ---- exact rank among the survivors ----
    for (int e = tid; e < Kout; e += nt) {
        const unsigned long long mine = g_keys[e];
        int rank = 0;
        for (int j = 0; j < Kout; ++j) rank += g_keys[j] > mine;
        s_wid[rank] = (int32_t)(0xffffffffu - (uint32_t)(mine & 0xffffffffull));
    }
    __syncthreads();
    return Kout;
}

__global__ void __launch_bounds__(WIDE_THREADS, 1) k_beam_encode_wide(const WideArgs a)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int tid = threadIdx.x, nt = blockDim.x, lane = tid & 31, warp = tid >> 5, nwarps = nt >> 5;
    const int DPm = a.DPmax, B = a.Bp;             // B: row capacity (strides); a.B: beams kept

    double* s_kl = reinterpret_cast<double*>(smem_raw);                  // [32]
    unsigned long long* s_sel = reinterpret_cast<unsigned long long*>(s_kl + 32);   // [2]
    float* s_T = reinterpret_cast<float*>(s_sel + 2);                    // [10008]
    float* s_sa = s_T + 10008;                                           // [DPm] x 8
    float* s_A = s_sa + DPm; float* s_E = s_A + DPm; float* s_M = s_E + DPm;
    float* s_cv = s_M + DPm; float* s_tv = s_cv + DPm; float* s_dmu = s_tv + DPm; float* s_cum = s_dmu + DPm;
    int32_t* s_wid = reinterpret_cast<int32_t*>(s_cum + DPm);            // [B] winners' flat ids, best first
    int32_t* s_hist = s_wid + B;                                         // [256]
    int32_t* s_misc = s_hist + 256;                                      // [4]

    for (int i = tid; i < 10007; i += nt) s_T[i] = a.T[i];
    float* beams = a.g_beams + (size_t)blockIdx.x * 2 * B * DPm;
    float* scores = a.g_scores + (size_t)blockIdx.x * (size_t)a.S * B;
    int2* hist = a.g_hist + (size_t)blockIdx.x * a.max_aux * B;
    int32_t* hsum = a.g_hsum + (size_t)blockIdx.x * 2 * B;
    unsigned long long* keys = a.g_keys + (size_t)blockIdx.x * B;

    for (;;) {
        __syncthreads();
        if (tid == 0) s_misc[0] = atomicAdd(a.work_counter, 1);
        __syncthreads();
        const int blk = s_misc[0];
        if (blk >= a.nb) break;
        const int64_t off = a.offs[blk];
        const int D = (int)(a.offs[blk + 1] - off);
        const BeamGeom g = make_geom(D);
        const int bstride = g.DP;                  // floats per beam row

        // ---- load + KL (coder.py:499-501) ----
        for (int i = tid; i < g.DP; i += nt) {
            s_cv[i] = 0.f; s_tv[i] = 0.f; s_dmu[i] = 0.f; s_cum[i] = 0.f;
            s_sa[i] = 0.f; s_A[i] = 0.f; s_E[i] = 0.f; s_M[i] = 0.f;
        }
        for (int i = tid; i < 2 * B * g.DP; i += nt) beams[i] = 0.f;
        for (int i = tid; i < 2 * B; i += nt) hsum[i] = 0;
        __syncthreads();
        for (int c = tid; c < g.nch; c += nt) {
            double acc = 0.0;
            const int hi = min(D, 32 * c + 32);
            for (int d = 32 * c; d < hi; ++d) {
                const int64_t gi = a.gidx ? a.gidx[off + d] : off + d;
                const float tl = a.t_loc[gi], ts = a.t_scale[gi], pl = a.p_loc[gi], ps = a.p_scale[gi];
                acc = __dadd_rn(acc, kl_dim(tl, ts, pl, ps));
                const int ci = ci_index(d, g.P);
                s_cv[ci] = __fmul_rn(ps, ps);
                s_tv[ci] = __fmul_rn(ts, ts);
                s_dmu[ci] = __fadd_rn(tl, -pl);
            }
            s_kl[c] = acc;
        }
        const double kld = block_tree_sum_f64(s_kl, g.nch);
        const int n_aux = n_aux_from_kl((float)kld, a.omega);
        int status = IREC_BLK_OK;
        if (n_aux <= 0) status = IREC_BLK_BAD_KL;
        else if (n_aux > a.max_aux || n_aux > a.ratio_len) status = IREC_BLK_TOO_LONG;
        if (tid == 0) { a.out_n_aux[blk] = n_aux; a.out_status[blk] = status; }
        if (status != IREC_BLK_OK) continue;

        int Bcur = 1, cur = 0;
        const float4* sa4 = reinterpret_cast<const float4*>(s_sa);
        const float4* A4 = reinterpret_cast<const float4*>(s_A);
        const float4* E4 = reinterpret_cast<const float4*>(s_E);
        const float4* M4 = reinterpret_cast<const float4*>(s_M);
        const int nsg = (a.S + g.SPW - 1) / g.SPW;
        const int lg = lane & (g.P - 1);
        const int nq = g.DP >> 2;

        for (int t = 0; t < n_aux; ++t) {
            // ---- schedule (beam_search_coder.py:64-77) ----
            const float ratio = a.ratio_tab[n_aux - 1 - t];
            for (int i = tid; i < g.DP; i += nt) {
                const float cv = s_cv[i];
                if (cv != 0.f) {
                    const SchedOut o = beam_sched_dim(cv, s_tv[i], s_dmu[i], s_cum[i], ratio);
                    s_sa[i] = o.sa; s_A[i] = o.A; s_E[i] = o.E; s_M[i] = o.M; s_cum[i] = o.cum_next;
                }
            }
            __syncthreads();

            // ---- score all S * Bcur candidates (beam_search_coder.py:79-84), pages of WIDE_PB beam slots ----
            const TfStream st = tf_stream_seeded(a.seed + t, a.seed + t);
            const float* bcur = beams + (size_t)cur * B * bstride;
            const int32_t* hs = hsum + cur * B;
            const int npages = (Bcur + WIDE_PB - 1) / WIDE_PB;
            for (int item = warp; item < nsg * npages; item += nwarps) {
                const int sg = item / npages, pg = item - sg * npages;
                const int s = sg * g.SPW + lane / g.P;
                BeamHash h[WIDE_PB];
#pragma unroll
                for (int b = 0; b < WIDE_PB; ++b) {
                    const int bb = pg * WIDE_PB + b;
                    h[b].h = bb < Bcur ? (uint32_t)hash_from_sum(hs[bb]) : 0u;
                    h[b].h4 = 4u * h[b].h;
                }
                float acc[WIDE_PB];
                // slots beyond Bcur inside the last page read rows that exist (B is padded to a multiple of WIDE_PB) and are ignored
                score_sample<WIDE_PB, false>(s_T, sa4, A4, E4, M4, reinterpret_cast<const float4*>(bcur + (size_t)pg * WIDE_PB * bstride),
                                             g, lg, st, (uint64_t)min(s, a.S - 1), h, acc);
                if (lg == 0 && s < a.S) {
#pragma unroll
                    for (int b = 0; b < WIDE_PB; ++b) {
                        const int bb = pg * WIDE_PB + b;
                        if (bb < Bcur) scores[(size_t)s * Bcur + bb] = (acc[b] == acc[b]) ? acc[b] : __int_as_float(0xff800000);
                    }
                }
            }
            __syncthreads();

            // ---- top-B (beam_search_coder.py:86-101) ----
            const int Kout = wide_topk(scores, a.S * Bcur, a.B, s_wid, keys, s_hist, s_sel);

            // ---- history + hash sums of the new beams ----
            int32_t* hs_new = hsum + (cur ^ 1) * B;
            for (int j = tid; j < Kout; j += nt) {
                const int f = s_wid[j];
                const int sj = f / Bcur, bj = f - sj * Bcur;
                hist[(size_t)t * B + j] = make_int2(sj, bj);
                hs_new[j] = hsum_extend(hs[bj], sj, t);
            }
            // ---- re-materialise the winners into the other buffer: beam_j <- beam_{b_j} + T[k] * sigma (reference order) ----
            float* bnew = beams + (size_t)(cur ^ 1) * B * bstride;
            for (int task = tid; task < Kout * nq; task += nt) {
                const int j = task / nq, qq = task - j * nq;
                const int slot = qq / (8 * g.P), rem = qq - slot * 8 * g.P;
                const int iq = rem / g.P, l = rem - iq * g.P;
                const int d0 = slot * 32 * g.P + 32 * l + 4 * iq;
                const int f = s_wid[j];
                const int sj = f / Bcur, bj = f - sj * Bcur;
                float4 o = make_float4(0.f, 0.f, 0.f, 0.f);
                if (d0 < D) {
                    const uint32_t h = (uint32_t)hash_from_sum(hs[bj]);
                    const uint4 u = tf_stream_quad_at(st, (uint64_t)sj * (uint64_t)D + (uint64_t)d0);
                    const float4 sa = sa4[qq];
                    const float4 ob = reinterpret_cast<const float4*>(bcur + (size_t)bj * bstride)[qq];
                    o.x = __fadd_rn(ob.x, __fmul_rn(s_T[beam_mix(beam_r_from_u32(u.x), h)], sa.x));
                    o.y = __fadd_rn(ob.y, __fmul_rn(s_T[beam_mix(beam_r_from_u32(u.y), h)], sa.y));
                    o.z = __fadd_rn(ob.z, __fmul_rn(s_T[beam_mix(beam_r_from_u32(u.z), h)], sa.z));
                    o.w = __fadd_rn(ob.w, __fmul_rn(s_T[beam_mix(beam_r_from_u32(u.w), h)], sa.w));
                }
                reinterpret_cast<float4*>(bnew + (size_t)j * bstride)[qq] = o;
            }
            __syncthreads();
            Bcur = Kout;
            cur ^= 1;
        }

        // ---- emit: indices of the best beam (back-pointers) and its sample ----
        if (tid == 0) {
            int j = 0;
            int32_t* oi = a.out_indices + (size_t)blk * a.max_aux;
            for (int t = n_aux - 1; t >= 0; --t) {
                const int2 e = hist[(size_t)t * B + j];
                oi[t] = e.x;
                j = e.y;
            }
        }
        const float* best = beams + (size_t)cur * B * bstride;
        for (int d = tid; d < D; d += nt) {
            const int64_t gi = a.gidx ? a.gidx[off + d] : off + d;
            a.out_sample[gi] = __fadd_rn(best[ci_index(d, g.P)], a.p_loc[gi]);
        }
    }
}

// =============================================================================================
// host side
// =============================================================================================
static int wide_bpad(int B) { return (B + WIDE_PB - 1) / WIDE_PB * WIDE_PB; }
static size_t wide_smem_bytes(int DPmax, int Bp)
{
    return sizeof(double) * 32 + sizeof(unsigned long long) * 2 + sizeof(float) * (10008 + 8 * (size_t)DPmax) +
           sizeof(int32_t) * ((size_t)Bp + 256 + 4) + 16;
}
static size_t wide_up(size_t x) { return (x + 255) / 256 * 256; }
// global state of one CTA (the regions are laid out region by region: [grid][...] each, region starts 256-byte aligned)
static size_t wide_per_cta_bytes(int DPmax, int S, int Bp, int max_aux)
{
    return sizeof(float) * 2 * (size_t)Bp * DPmax + sizeof(float) * (size_t)S * Bp + sizeof(int2) * (size_t)max_aux * Bp +
           sizeof(int32_t) * 2 * (size_t)Bp + sizeof(unsigned long long) * (size_t)Bp;
}
// CTAs of a launch: one per SM, fewer if the per-CTA state (2 beam buffers of B x DP floats) would exceed IREC_WIDE_WS_MAX in total
#define IREC_WIDE_WS_MAX ((size_t)4 << 30)
static int wide_grid(int nb, int DPmax, int S, int Bp, int max_aux)
{
    const size_t per = wide_per_cta_bytes(DPmax, S, Bp, max_aux);
    const size_t fit = std::max<size_t>(1, IREC_WIDE_WS_MAX / per);
    return (int)std::max<size_t>(1, std::min<size_t>(std::min<size_t>((size_t)nb, (size_t)irec_device().sm_count), fit));
}

bool irec_wide_supported(int nb, int max_D, int S, int B)
{
    if (B <= 32 || B > IREC_WIDE_MAX_BEAMS || max_D < 1 || max_D > 1024 || S < 1 || nb < 1) return false;
    if ((int64_t)S * wide_bpad(B) >= (1LL << 31)) return false;
    return wide_smem_bytes(make_geom(max_D).DP, wide_bpad(B)) <= (size_t)irec_device().max_smem_optin;
}

size_t irec_wide_workspace_bytes(int nb, int max_D, int S, int B, int max_aux)
{
    if (!irec_wide_supported(nb, max_D, S, B)) return 0;
    const int Bp = wide_bpad(B), DP = make_geom(max_D).DP;
    return 256 + (size_t)wide_grid(nb, DP, S, Bp, max_aux) * wide_per_cta_bytes(DP, S, Bp, max_aux) + 5 * 256;
}

int irec_launch_wide(const float* t_loc, const float* t_scale, const float* p_loc, const float* p_scale,
                     const int64_t* gidx, const int64_t* offs, int nb, int max_D, float omega, int S, int B, int64_t seed,
                     int32_t* out_indices, int max_aux, int32_t* out_n_aux, int32_t* out_status, float* out_sample,
                     void* workspace, cudaStream_t s)
{
    const int Bp = wide_bpad(B), DP = make_geom(max_D).DP;
    const int grid = wide_grid(nb, DP, S, Bp, max_aux);
    unsigned char* w = reinterpret_cast<unsigned char*>(workspace);
    WideArgs a;
    a.t_loc = t_loc; a.t_scale = t_scale; a.p_loc = p_loc; a.p_scale = p_scale; a.gidx = gidx; a.offs = offs; a.nb = nb;
    a.omega = omega; a.S = S; a.B = B; a.Bp = Bp; a.seed = seed;
    a.out_indices = out_indices; a.max_aux = max_aux; a.out_n_aux = out_n_aux; a.out_status = out_status; a.out_sample = out_sample;
    a.T = irec_device().d_T; a.ratio_tab = irec_ratio_tab(); a.ratio_len = irec_ratio_len();
    a.work_counter = reinterpret_cast<int*>(w); a.DPmax = DP;
    w += 256;
    a.g_beams = reinterpret_cast<float*>(w); w += wide_up((size_t)grid * sizeof(float) * 2 * (size_t)Bp * DP);
    a.g_scores = reinterpret_cast<float*>(w); w += wide_up((size_t)grid * sizeof(float) * (size_t)S * Bp);
    a.g_hist = reinterpret_cast<int2*>(w); w += wide_up((size_t)grid * sizeof(int2) * (size_t)max_aux * Bp);
    a.g_hsum = reinterpret_cast<int32_t*>(w); w += wide_up((size_t)grid * sizeof(int32_t) * 2 * (size_t)Bp);
    a.g_keys = reinterpret_cast<unsigned long long*>(w);
    if (cudaMemsetAsync(a.work_counter, 0, 256, s) != cudaSuccess) return irec_fail(IREC_E_CUDA, "beam_encode (wide): memset failed");
    const size_t smem = wide_smem_bytes(DP, Bp);
    if (cudaFuncSetAttribute(k_beam_encode_wide, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess)
        return irec_fail(IREC_E_CUDA, "beam_encode (wide): shared memory attribute failed");
    k_beam_encode_wide<<<grid, WIDE_THREADS, smem, s>>>(a);
    irec_count_launch();
    return irec_check_launch("k_beam_encode_wide");
}
