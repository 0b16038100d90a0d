// irec_beam.cu -- beam-search coder kernels (sm_100a) behind the C ABI of include/irec.h.
//
// Reference path being replaced: rec/coding/beam_search_coder.py:33-151 driven from
// rec/coding/coder.py:412-491 (GaussianCoder.encode/decode over coder-blocks).
//
// Kernels
//   k_kl_naux              per block KL(q||p) and n_aux                       (coder.py:499-501)
//   k_beam_encode_resident K1a: one CTA owns one coder-block for all of its partitions: schedule,
//                          candidate scoring, top-B, beam update all in shared memory / registers
//   k_beam_decode          K3: replay of the chosen candidates, O(n_aux * D)
//   k_gp_* / k_topb_merge  K1b/K5: the same step split per partition over the whole grid and over
//                          candidate ranges (large S, large D, multi-GPU candidate sharding)
#include "irec_beam.cuh"
#include "irec_host.h"
#include <string.h>
#include <mutex>
#include <vector>
#include <type_traits>
#include <cstdio>
#include <cstdlib>

// =============================================================================================
// k_kl_naux
// =============================================================================================
#define KL_MAX_CHUNKS 4096
__global__ void k_kl_naux(const float* __restrict__ t_loc, const float* __restrict__ t_scale,
                          const float* __restrict__ p_loc, const float* __restrict__ p_scale,
                          const int64_t* __restrict__ gidx, const int64_t* __restrict__ offs, int nb, float omega,
                          float* out_kl, int32_t* out_n_aux)
{
    __shared__ double cs[KL_MAX_CHUNKS];
    for (int blk = blockIdx.x; blk < nb; blk += gridDim.x) {
        const int64_t off = offs[blk];
        const int D = (int)(offs[blk + 1] - off);
        const int nch = (D + 31) >> 5;
        for (int c = threadIdx.x; c < nch; c += blockDim.x) {
            double acc = 0.0;
            const int hi = min(D, 32 * c + 32);
            for (int d = 32 * c; d < hi; ++d) {
                const int64_t gi = gidx ? gidx[off + d] : off + d;
                acc = __dadd_rn(acc, kl_dim(t_loc[gi], t_scale[gi], p_loc[gi], p_scale[gi]));
            }
            cs[c] = acc;
        }
        const double kl = block_tree_sum_f64(cs, nch);
        if (threadIdx.x == 0) {
            const float klf = (float)kl;
            if (out_kl) out_kl[blk] = klf;
            if (out_n_aux) out_n_aux[blk] = n_aux_from_kl(klf, omega);
        }
        __syncthreads();
    }
}

// =============================================================================================
// K3: k_beam_decode  (beam_search_coder.py:124-148)
// each thread owns quads of dims and replays all partitions for them; no synchronisation needed
// =============================================================================================
__global__ void k_beam_decode(const float* __restrict__ p_loc, const float* __restrict__ p_scale,
                              const int64_t* __restrict__ gidx, const int64_t* __restrict__ offs, int nb,
                              int64_t seed, const int32_t* __restrict__ indices, int max_aux,
                              const int32_t* __restrict__ n_aux_arr, const float* __restrict__ T,
                              const float* __restrict__ ratio_tab, int ratio_len, float* __restrict__ out,
                              int32_t* __restrict__ out_status)
{
    for (int blk = blockIdx.x; blk < nb; blk += gridDim.x) {
        const int64_t off = offs[blk];
        const int D = (int)(offs[blk + 1] - off);
        int n_aux = n_aux_arr[blk];
        // an index list longer than the (learned) ratio table or than the row: the reference raises CodingError
        // (coder.py:226-231); here the block decodes to NaN and is flagged
        const bool bad = n_aux < 0 || n_aux > max_aux || n_aux > ratio_len;
        if (threadIdx.x == 0 && out_status) out_status[blk] = bad ? IREC_BLK_TOO_LONG : IREC_BLK_OK;
        if (bad) n_aux = 0;
        const int32_t* idx = indices + (size_t)blk * max_aux;
        const int nq = (D + 3) >> 2;
        for (int q = threadIdx.x; q < nq; q += blockDim.x) {
            const int d0 = 4 * q;
            float ps[4], pl[4], cum[4], smp[4];
            int64_t gi[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const int d = min(d0 + e, D - 1);
                gi[e] = gidx ? gidx[off + d] : off + d;
                ps[e] = p_scale[gi[e]]; pl[e] = p_loc[gi[e]];
                cum[e] = 0.f; smp[e] = 0.f;
            }
            int32_t hsum = 0;
            for (int t = 0; t < n_aux; ++t) {
                const float ratio = ratio_tab[n_aux - 1 - t];
                const int32_t s = idx[t];
                const uint32_t h = (uint32_t)hash_from_sum(hsum);
                const TfStream st = tf_stream_seeded(seed + t, seed + t);
                const uint4 u = tf_stream_quad_at(st, (uint64_t)s * (uint64_t)D + (uint64_t)d0);
                const uint32_t uu[4] = { u.x, u.y, u.z, u.w };
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    const float cv = __fmul_rn(ps[e], ps[e]);
                    const float v = __fmul_rn(ratio, __fadd_rn(cv, -cum[e]));
                    const float a = __fmul_rn(T[beam_mix(beam_r_from_u32(uu[e]), h)], __fsqrt_rn(v));
                    smp[e] = __fadd_rn(smp[e], a);
                    cum[e] = __fadd_rn(cum[e], v);
                }
                hsum = hsum_extend(hsum, s, t);
            }
#pragma unroll
            for (int e = 0; e < 4; ++e)
                if (d0 + e < D) out[gi[e]] = bad ? __int_as_float(0x7fc00000) : __fadd_rn(smp[e], pl[e]);
        }
    }
}

// =============================================================================================
// K1a: k_beam_encode_resident
// =============================================================================================
#define TOPK_CAP 1024
#define UPD_MAX 8

struct ResidentArgs {
    const float* t_loc; const float* t_scale; const float* p_loc; const float* p_scale;
    const int64_t* gidx; const int64_t* offs; int nb;
    float omega; int S; int B; int64_t seed;
    int32_t* out_indices; int max_aux; int32_t* out_n_aux; int32_t* out_status; float* out_sample;
    const float* T; const float* ratio_tab; int ratio_len;
    int2* hist;            // [gridDim.x][max_aux][BMAX]
    int* work_counter;     // dynamic block queue
    int DPmax;             // padded dims capacity of the shared arrays (multiple of 32)
    int NC;                // capacity of the score array (>= S * BMAX)
};

template <int BMAX>
constexpr int resident_max_threads()
{
#ifdef IREC_RES_THREADS
    return IREC_RES_THREADS;
#else
    return BMAX <= 10 ? 640 : (BMAX <= 20 ? 576 : 512);
#endif
}

template <int BMAX>
__global__ void __launch_bounds__(resident_max_threads<BMAX>(), 1) k_beam_encode_resident(const ResidentArgs a)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int tid = threadIdx.x, nt = blockDim.x, lane = tid & 31, warp = tid >> 5, nwarps = nt >> 5;
    const int DPm = a.DPmax;

    // ---- shared memory carve-up (all float4-aligned) ----
    double* s_kl = reinterpret_cast<double*>(smem_raw);                  // [32]
    float* s_T = reinterpret_cast<float*>(s_kl + 32);                    // [10008]
    float* s_beams = s_T + 10008;                                        // [BMAX][DPm]
    float* s_sa = s_beams + (size_t)BMAX * DPm;                          // [DPm] x 8
    float* s_A = s_sa + DPm; float* s_E = s_A + DPm; float* s_M = s_E + DPm;
    float* s_cv = s_M + DPm; float* s_tv = s_cv + DPm; float* s_dmu = s_tv + DPm; float* s_cum = s_dmu + DPm;
    float* s_scores = s_cum + DPm;                                       // [NC]
    float* s_gmax = s_scores + a.NC;                                     // [1024]
    float* s_wsc = s_gmax + 1024;                                        // [32] winners' scores
    int32_t* s_wid = reinterpret_cast<int32_t*>(s_wsc + 32);             // [32] winners' flat ids
    int32_t* s_list = s_wid + 32;                                        // [TOPK_CAP]
    int32_t* s_ctl = s_list + TOPK_CAP;                                  // [4]
    int32_t* s_hsum = s_ctl + 4;                                         // [2][32]
    int32_t* s_misc = s_hsum + 64;                                       // [4]: blk, n_aux, status

    for (int i = tid; i < 10007; i += nt) s_T[i] = a.T[i];
    int2* hist = a.hist + (size_t)blockIdx.x * a.max_aux * BMAX;

    for (;;) {
        __syncthreads();
        if (tid == 0) s_misc[0] = atomicAdd(a.work_counter, 1);
        __syncthreads();
        const int blk = s_misc[0];
        if (blk >= a.nb) break;
        const int64_t off = a.offs[blk];
        const int D = (int)(a.offs[blk + 1] - off);
        const BeamGeom g = make_geom(D);

        // ---- load + KL ----
        for (int i = tid; i < g.DP; i += nt) {
            s_cv[i] = 0.f; s_tv[i] = 0.f; s_dmu[i] = 0.f; s_cum[i] = 0.f;
            s_sa[i] = 0.f; s_A[i] = 0.f; s_E[i] = 0.f; s_M[i] = 0.f;
        }
        for (int i = tid; i < BMAX * g.DP; i += nt) s_beams[i] = 0.f;   // rows use stride g.DP
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

        if (tid < 64) s_hsum[tid] = 0;
        int Bcur = 1, hb = 0;                      // hb: which half of s_hsum is current
        const float4* sa4 = reinterpret_cast<const float4*>(s_sa);
        const float4* A4 = reinterpret_cast<const float4*>(s_A);
        const float4* E4 = reinterpret_cast<const float4*>(s_E);
        const float4* M4 = reinterpret_cast<const float4*>(s_M);
        const float4* beams4 = reinterpret_cast<const float4*>(s_beams);
        const int nsg = (a.S + g.SPW - 1) / g.SPW;
        const int lg = lane & (g.P - 1);

        for (int t = 0; t < n_aux; ++t) {
            // ---- schedule ----
            const float ratio = a.ratio_tab[n_aux - 1 - t];
            for (int i = tid; i < g.DP; i += nt) {
                const float cv = s_cv[i];
                if (cv != 0.f) {                   // padding stays zero
                    const SchedOut o = beam_sched_dim(cv, s_tv[i], s_dmu[i], s_cum[i], ratio);
                    s_sa[i] = o.sa; s_A[i] = o.A; s_E[i] = o.E; s_M[i] = o.M; s_cum[i] = o.cum_next;
                }
            }
            __syncthreads();

            // ---- score all S * Bcur candidates ----
            const TfStream st = tf_stream_seeded(a.seed + t, a.seed + t);
            const int32_t* hs = s_hsum + 32 * hb;
            if (Bcur == 1) {
                BeamHash h[1];
                h[0].h = (uint32_t)hash_from_sum(hs[0]); h[0].h4 = 4u * h[0].h;
                float acc[1];
                for (int sg = warp; sg < nsg; sg += nwarps) {
                    const int s = sg * g.SPW + lane / g.P;
                    score_sample<1, false>(s_T, sa4, A4, E4, M4, beams4, g, lg, st, (uint64_t)min(s, a.S - 1), h, acc);
                    if (lg == 0 && s < a.S) s_scores[s] = (acc[0] == acc[0]) ? acc[0] : __int_as_float(0xff800000);
                }
            } else {
                BeamHash h[BMAX];
#pragma unroll
                for (int b = 0; b < BMAX; ++b) {
                    h[b].h = b < Bcur ? (uint32_t)hash_from_sum(hs[b]) : 0u;
                    h[b].h4 = 4u * h[b].h;
                }
                float acc[BMAX];
                for (int sg = warp; sg < nsg; sg += nwarps) {
                    const int s = sg * g.SPW + lane / g.P;
                    score_sample<BMAX, false>(s_T, sa4, A4, E4, M4, beams4, g, lg, st, (uint64_t)min(s, a.S - 1), h, acc);
                    if (lg == 0 && s < a.S) {
#pragma unroll
                        for (int b = 0; b < BMAX; ++b)
                            if (b < Bcur)
                                s_scores[s * Bcur + b] = (acc[b] == acc[b]) ? acc[b] : __int_as_float(0xff800000);
                    }
                }
            }
            __syncthreads();

            // ---- top-B ----
            const int Kout = block_topk(s_scores, nullptr, a.S * Bcur, a.B, s_wsc, s_wid, s_gmax, s_list, TOPK_CAP, s_ctl);

            // ---- history + hash sums of the new beams ----
            int32_t* hs_new = s_hsum + 32 * (hb ^ 1);
            if (tid < Kout) {
                const int f = s_wid[tid];
                const int sj = f / Bcur, bj = f - sj * Bcur;
                hist[(size_t)t * BMAX + tid] = make_int2(sj, bj);
                hs_new[tid] = hsum_extend(hs[bj], sj, t);
            }

            // ---- re-materialise the winners: beam_j <- beam_{b_j} + a(s_j, b_j), in place, two phases ----
            const int nq = g.DP >> 2;
            const int Qr = max(1, (nt * UPD_MAX) / Kout);
            for (int qa = 0; qa < nq; qa += Qr) {
                const int ntask = min(Qr, nq - qa) * Kout;
                float4 nv[UPD_MAX];
#pragma unroll
                for (int k = 0; k < UPD_MAX; ++k) {
                    const int task = tid + k * nt;
                    if (task < ntask) {
                        const int qq = qa + task / Kout, j = task % Kout;
                        // physical quad qq -> first dim
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
                            const float4 ob = beams4[bj * (g.DP >> 2) + qq];
                            o.x = __fadd_rn(ob.x, __fmul_rn(s_T[beam_mix(beam_r_from_u32(u.x), h)], sa.x));
                            o.y = __fadd_rn(ob.y, __fmul_rn(s_T[beam_mix(beam_r_from_u32(u.y), h)], sa.y));
                            o.z = __fadd_rn(ob.z, __fmul_rn(s_T[beam_mix(beam_r_from_u32(u.z), h)], sa.z));
                            o.w = __fadd_rn(ob.w, __fmul_rn(s_T[beam_mix(beam_r_from_u32(u.w), h)], sa.w));
                            // dims beyond D inside the last quad: sa = 0 there, so o stays the (zero) padding
                        }
                        nv[k] = o;
                    }
                }
                __syncthreads();
#pragma unroll
                for (int k = 0; k < UPD_MAX; ++k) {
                    const int task = tid + k * nt;
                    if (task < ntask) {
                        const int qq = qa + task / Kout, j = task % Kout;
                        reinterpret_cast<float4*>(s_beams)[j * (g.DP >> 2) + qq] = nv[k];
                    }
                }
                __syncthreads();
            }
            Bcur = Kout;
            hb ^= 1;
        }

        // ---- emit: indices of the best beam (trace the back-pointers) and its sample ----
        __syncthreads();
        if (tid == 0) {
            int j = 0;
            int32_t* oi = a.out_indices + (size_t)blk * a.max_aux;
            for (int t = n_aux - 1; t >= 0; --t) {
                const int2 e = hist[(size_t)t * BMAX + j];
                oi[t] = e.x;
                j = e.y;
            }
        }
        for (int d = tid; d < D; d += nt) {
            const int64_t gi = a.gidx ? a.gidx[off + d] : off + d;
            a.out_sample[gi] = __fadd_rn(s_beams[ci_index(d, g.P)], a.p_loc[gi]);
        }
    }
}

#include "irec_resident2.cuh"

// =============================================================================================
// K1b / K5: general path -- one partition per launch, candidates spread over the grid
// Device state of one block ("BeamState", owned by the caller, laid out by the library):
// =============================================================================================
struct BeamStateHdr {
    int32_t D, P, nslots, DP;
    int32_t B, S, max_aux, n_aux;
    int32_t Bcur, cur;          // cur: which beams/hsum buffer is current
    int32_t status, pad;
    float kl, omega;
    int64_t seed;
};
// float arrays after the header (each DP long): cv, tv, dmu, cum, sa, A, E, M, pl(p_loc), beams[2][B][DP]
// then int32 hsum[2][32], int2 hist[max_aux][32]
__host__ __device__ inline size_t state_off_floats() { return 256; }
__host__ __device__ inline float* st_arr(void* state, int which, int DP)
{
    return reinterpret_cast<float*>(reinterpret_cast<unsigned char*>(state) + state_off_floats()) + (size_t)which * DP;
}
__host__ __device__ inline float* st_beams(void* state, int buf, int B, int DP)
{
    return st_arr(state, 9, DP) + (size_t)buf * B * DP;
}
__host__ __device__ inline int32_t* st_hsum(void* state, int buf, int B, int DP)
{
    return reinterpret_cast<int32_t*>(st_arr(state, 9, DP) + (size_t)2 * B * DP) + 32 * buf;
}
__host__ __device__ inline int2* st_hist(void* state, int B, int DP)
{
    return reinterpret_cast<int2*>(st_hsum(state, 0, B, DP) + 64);
}
static size_t state_bytes(int D, int B, int max_aux)
{
    const BeamGeom g = make_geom(D);
    return state_off_floats() + sizeof(float) * ((size_t)9 * g.DP + (size_t)2 * B * g.DP) + sizeof(int32_t) * 64 +
           sizeof(int2) * (size_t)max_aux * 32 + 64;
}

// init: gather the block's parameters, KL, n_aux, zero the beams (single CTA)
__global__ void k_gp_init(void* state, const float* __restrict__ t_loc, const float* __restrict__ t_scale,
                          const float* __restrict__ p_loc, const float* __restrict__ p_scale,
                          const int64_t* __restrict__ gidx, int64_t off, int D, int B, int S, int max_aux,
                          int ratio_len, float omega, int64_t seed)
{
    __shared__ double cs[KL_MAX_CHUNKS];
    const BeamGeom g = make_geom(D);
    BeamStateHdr* hdr = reinterpret_cast<BeamStateHdr*>(state);
    const int tid = threadIdx.x, nt = blockDim.x;
    float* cv = st_arr(state, 0, g.DP); float* tv = st_arr(state, 1, g.DP); float* dmu = st_arr(state, 2, g.DP);
    float* plc = st_arr(state, 8, g.DP);
    for (int i = tid; i < 9 * g.DP; i += nt) st_arr(state, 0, g.DP)[i] = 0.f;
    for (int i = tid; i < 2 * B * g.DP; i += nt) st_beams(state, 0, B, g.DP)[i] = 0.f;
    if (tid < 64) st_hsum(state, 0, B, g.DP)[tid] = 0;
    __syncthreads();
    for (int c = tid; c < g.nch; c += nt) {
        double acc = 0.0;
        const int hi = min(D, 32 * c + 32);
        for (int d = 32 * c; d < hi; ++d) {
            const int64_t gi = gidx ? gidx[off + d] : off + d;
            const float tl = t_loc[gi], ts = t_scale[gi], pl = p_loc[gi], ps = p_scale[gi];
            acc = __dadd_rn(acc, kl_dim(tl, ts, pl, ps));
            const int ci = ci_index(d, g.P);
            cv[ci] = __fmul_rn(ps, ps); tv[ci] = __fmul_rn(ts, ts); dmu[ci] = __fadd_rn(tl, -pl); plc[ci] = pl;
        }
        cs[c] = acc;
    }
    const double kld = block_tree_sum_f64(cs, g.nch);
    if (tid == 0) {
        const int n_aux = n_aux_from_kl((float)kld, omega);
        hdr->D = D; hdr->P = g.P; hdr->nslots = g.nslots; hdr->DP = g.DP;
        hdr->B = B; hdr->S = S; hdr->max_aux = max_aux; hdr->n_aux = n_aux;
        hdr->Bcur = 1; hdr->cur = 0;
        hdr->status = n_aux <= 0 ? IREC_BLK_BAD_KL : ((n_aux > max_aux || n_aux > ratio_len) ? IREC_BLK_TOO_LONG : IREC_BLK_OK);
        hdr->kl = (float)kld; hdr->omega = omega; hdr->seed = seed;
    }
}

// schedule of partition t
__global__ void k_gp_params(void* state, int t, const float* __restrict__ ratio_tab)
{
    BeamStateHdr* hdr = reinterpret_cast<BeamStateHdr*>(state);
    if (hdr->status != IREC_BLK_OK || t >= hdr->n_aux) return;
    const int DP = hdr->DP;
    const float ratio = ratio_tab[hdr->n_aux - 1 - t];
    const float* cv = st_arr(state, 0, DP); const float* tv = st_arr(state, 1, DP); const float* dmu = st_arr(state, 2, DP);
    float* cum = st_arr(state, 3, DP); float* sa = st_arr(state, 4, DP); float* A = st_arr(state, 5, DP);
    float* E = st_arr(state, 6, DP); float* M = st_arr(state, 7, DP);
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < DP; i += gridDim.x * blockDim.x) {
        const float c = cv[i];
        if (c != 0.f) {
            const SchedOut o = beam_sched_dim(c, tv[i], dmu[i], cum[i], ratio);
            sa[i] = o.sa; A[i] = o.A; E[i] = o.E; M[i] = o.M; cum[i] = o.cum_next;
        }
    }
}

// score candidates s in [s_begin, s_end) x beams, keep the CTA's best B, write them to
// out_rec[blockIdx.x * B ...] (+ count).  Dynamic smem: T + candidate buffer + top-k scratch.
struct ScoreArgs {
    void* state; int t; const float* T;
    int64_t s_begin, s_end;
    irec_record_t* out_rec; int32_t* out_cnt;
    int cand_cap;           // capacity of the per-CTA candidate buffer (>= nwarps*SPW*BMAX + 32)
};

template <int BMAX>
__global__ void __launch_bounds__(256) k_gp_score_topb(const ScoreArgs a)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    BeamStateHdr* hdr = reinterpret_cast<BeamStateHdr*>(a.state);
    const int tid = threadIdx.x, nt = blockDim.x, lane = tid & 31, warp = tid >> 5, nwarps = nt >> 5;
    if (hdr->status != IREC_BLK_OK || a.t >= hdr->n_aux) {
        if (tid == 0) a.out_cnt[blockIdx.x] = 0;
        return;
    }
    BeamGeom g;
    g.D = hdr->D; g.P = hdr->P; g.nslots = hdr->nslots; g.DP = hdr->DP; g.nch = (g.D + 31) >> 5; g.SPW = 32 / g.P;
    const int B = hdr->B, Bcur = hdr->Bcur, cur = hdr->cur;
    const int cap = a.cand_cap;

    float* s_T = reinterpret_cast<float*>(smem_raw);                 // [10008]
    float* s_csc = s_T + 10008;                                      // [cap]
    int32_t* s_cid = reinterpret_cast<int32_t*>(s_csc + cap);        // [cap]
    float* s_gmax = reinterpret_cast<float*>(s_cid + cap);           // [256]
    float* s_wsc = s_gmax + 256;                                     // [32]
    int32_t* s_wid = reinterpret_cast<int32_t*>(s_wsc + 32);         // [32]
    int32_t* s_list = s_wid + 32;                                    // [TOPK_CAP]
    int32_t* s_ctl = s_list + TOPK_CAP;                              // [4]
    int32_t* s_cnt = s_ctl + 4;                                      // [1] candidate count
    float* s_tau = reinterpret_cast<float*>(s_cnt + 1);              // [1]

    for (int i = tid; i < 10007; i += nt) s_T[i] = a.T[i];
    if (tid == 0) { *s_cnt = 0; *s_tau = __int_as_float(0xff800000); }

    const float4* sa4 = reinterpret_cast<const float4*>(st_arr(a.state, 4, g.DP));
    const float4* A4 = reinterpret_cast<const float4*>(st_arr(a.state, 5, g.DP));
    const float4* E4 = reinterpret_cast<const float4*>(st_arr(a.state, 6, g.DP));
    const float4* M4 = reinterpret_cast<const float4*>(st_arr(a.state, 7, g.DP));
    const float4* beams4 = reinterpret_cast<const float4*>(st_beams(a.state, cur, B, g.DP));
    const int32_t* hs = st_hsum(a.state, cur, B, g.DP);
    const TfStream st = tf_stream_seeded(hdr->seed + a.t, hdr->seed + a.t);
    const int lg = lane & (g.P - 1);

    BeamHash h[BMAX];
#pragma unroll
    for (int b = 0; b < BMAX; ++b) {
        h[b].h = b < Bcur ? (uint32_t)hash_from_sum(hs[b]) : 0u;
        h[b].h4 = 4u * h[b].h;
    }

    // contiguous range of sample groups per CTA
    const int64_t nsg = (a.s_end - a.s_begin + g.SPW - 1) / g.SPW;
    const int64_t per = (nsg + gridDim.x - 1) / gridDim.x;
    const int64_t sg0 = (int64_t)blockIdx.x * per, sg1 = min(nsg, sg0 + per);
    __syncthreads();

    for (int64_t base = sg0; base < sg1; base += nwarps) {
        const int64_t sg = base + warp;
        const int64_t s = a.s_begin + sg * g.SPW + lane / g.P;
        const bool valid = sg < sg1 && s < a.s_end;
        const uint64_t s_eff = (uint64_t)(valid ? s : a.s_begin);
        float acc[BMAX];
        if (g.nslots > 1)
            score_sample<BMAX, true>(s_T, sa4, A4, E4, M4, beams4, g, lg, st, s_eff, h, acc);
        else
            score_sample<BMAX, false>(s_T, sa4, A4, E4, M4, beams4, g, lg, st, s_eff, h, acc);
        const float tau = *s_tau;
        if (valid && lg == 0) {
#pragma unroll
            for (int b = 0; b < BMAX; ++b) {
                if (b < Bcur) {
                    const float v = (acc[b] == acc[b]) ? acc[b] : __int_as_float(0xff800000);
                    if (v >= tau) {
                        const int pos = atomicAdd(s_cnt, 1);
                        s_csc[pos] = v;
                        s_cid[pos] = (int32_t)(s * Bcur + b);
                    }
                }
            }
        }
        __syncthreads();
        const int cnt = *s_cnt;
        if (cnt > B) {
            const int Kout = block_topk(s_csc, s_cid, cnt, B, s_wsc, s_wid, s_gmax, s_list, TOPK_CAP, s_ctl);
            if (tid < Kout) { s_csc[tid] = s_wsc[tid]; s_cid[tid] = s_wid[tid]; }
            if (tid == 0) { *s_cnt = Kout; if (Kout == B) *s_tau = s_wsc[Kout - 1]; }
            __syncthreads();
        }
    }
    __syncthreads();
    // final ordering of whatever is left (<= B entries, possibly unsorted if never compacted)
    const int cnt = *s_cnt;
    const int Kout = block_topk(s_csc, s_cid, cnt, B, s_wsc, s_wid, s_gmax, s_list, TOPK_CAP, s_ctl);
    if (tid < Kout) {
        irec_record_t r;
        const int f = s_wid[tid];
        r.score = s_wsc[tid]; r.s = f / Bcur; r.b = f - r.s * Bcur; r.pad = 0;
        a.out_rec[(size_t)blockIdx.x * B + tid] = r;
    }
    if (tid == 0) a.out_cnt[blockIdx.x] = Kout;
}

// ---- second generation of the general-path scoring kernel (D <= 1024): discrete-log addressing with the exponent
// computed in place (S is far too large for a table here: 2^24 samples x 64 dims), in-place two-choice bank assignment
// (r2_spread_banks), beams in passes of R2Pass::HB.  First generation (profiles/r1_gp_score_ncu.md): 255 registers,
// three IMADs per candidate-dim, 3.5 shared-memory wavefronts per gather, data pipe 74 % busy.  Same scores bit for bit.
#ifndef GP2_THREADS
#define GP2_THREADS 384
#endif
// In-place two-choice bank assignment of the quantile gathers (r2_spread_banks: one match.any per exponent) -- OFF.  It lowers
// the gathers from 3.5 to 2.7 shared-memory wavefronts, but MATCH.ANY itself is the slower resource: A/B of k_gp_fused<20> on
// one B200, C5 at S = 2^24: 117.0 ms with, 89.9 ms without (profiles/r2_gp_variants.log; the consumers of the match held 30 %
// of the warp samples in profiles/r1_gp_score2_ncu.md).  Generating the exponents one quad ahead did not hide it (125 ms).
#ifndef GP2_SPREAD
#define GP2_SPREAD false
#endif
#define GP_DL4_LEN 10008     // uint16 entries of the discrete-log table copied to shared memory (d_dl4 is allocated with 10008)
#ifndef GP2_NS
#define GP2_NS 2            // sample groups per warp and round: a beam quad read from global memory serves 8 candidate-dims
#endif            // (A/B at S = 2^24, one GPU: 512 threads x 1 group 118 ms, 384 x 2 108 ms, 512 x 2 109 ms, 256 x 3 137 ms)
struct Score2Args {
    void* state; int t; const float* T2; const uint16_t* dl4;
    int64_t s_begin, s_end;
    irec_record_t* out_rec; int32_t* out_cnt;
    int cand_cap;
};
// Candidate buffer of a CTA: everything that beats `tau` (the B-th best score at the last compaction) is appended; the buffer
// is compacted to its best B (exact block_topk) only when more than GP2_SLACK entries have piled up, not after every round --
// after the first rounds of a variable a round adds a handful of entries, and a compaction costs several CTA barriers
// (block_topk held 41 % of the warp samples of k_gp_fused when it ran after every round).  Capacity: one full round + GP2_SLACK.
#define GP2_SLACK 512
struct Gp2Sink {
    float* s_csc; int32_t* s_cid; int32_t* s_cnt;
    float tau; int Bcur, boff; int s_hi;      // s_hi: first sample beyond this CTA's range
    __device__ __forceinline__ void operator()(int sk, int b, float x, bool dup) const      // called by all 32 lanes of a warp
    {
        b += boff;
        const float v = (x == x) ? x : __int_as_float(0xff800000);
        const int pos = warp_append_pos(s_cnt, !dup && b < Bcur && sk < s_hi && v >= tau);
        if (pos >= 0) {
            s_csc[pos] = v;
            s_cid[pos] = sk * Bcur + b;
        }
    }
};

template <int BMAX>
__global__ void __launch_bounds__(GP2_THREADS, 1) k_gp_score_topb2(const Score2Args a)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    BeamStateHdr* hdr = reinterpret_cast<BeamStateHdr*>(a.state);
    const int tid = threadIdx.x, nt = blockDim.x, lane = tid & 31, warp = tid >> 5, nwarps = nt >> 5;
    if (hdr->status != IREC_BLK_OK || a.t >= hdr->n_aux) {
        if (tid == 0) a.out_cnt[blockIdx.x] = 0;
        return;
    }
    BeamGeom g;
    g.D = hdr->D; g.P = hdr->P; g.nslots = hdr->nslots; g.DP = hdr->DP; g.nch = (g.D + 31) >> 5; g.SPW = 32 / g.P;
    const int B = hdr->B, Bcur = hdr->Bcur, cur = hdr->cur;
    const int cap = a.cand_cap;

    float* s_T2 = reinterpret_cast<float*>(smem_raw);                // [IREC_T2_LEN]
    float* s_csc = s_T2 + IREC_T2_LEN;                               // [cap]
    int32_t* s_cid = reinterpret_cast<int32_t*>(s_csc + cap);        // [cap]
    float* s_gmax = reinterpret_cast<float*>(s_cid + cap);           // [GP2_THREADS]
    float* s_wsc = s_gmax + GP2_THREADS;                             // [32]
    int32_t* s_wid = reinterpret_cast<int32_t*>(s_wsc + 32);         // [32]
    int32_t* s_list = s_wid + 32;                                    // [TOPK_CAP]
    int32_t* s_ctl = s_list + TOPK_CAP;                              // [4]
    int32_t* s_cnt = s_ctl + 4;                                      // [1] candidate count
    float* s_tau = reinterpret_cast<float*>(s_cnt + 1);              // [1]
    uint32_t* s_cb = reinterpret_cast<uint32_t*>(s_tau + 1) + 2;     // [32] 4 * dlog(h_b)
    uint16_t* s_dl4 = reinterpret_cast<uint16_t*>(s_cb + 32);        // [GP_DL4_LEN] the discrete-log table (r2_exp4<true>)

    {
        const float4* src = reinterpret_cast<const float4*>(a.T2);
        float4* dst = reinterpret_cast<float4*>(s_T2);
        for (int i = tid; i < IREC_T2_LEN / 4; i += nt) dst[i] = src[i];
        const uint32_t* dsrc = reinterpret_cast<const uint32_t*>(a.dl4);
        uint32_t* ddst = reinterpret_cast<uint32_t*>(s_dl4);
        for (int i = tid; i < GP_DL4_LEN / 2; i += nt) ddst[i] = dsrc[i];
    }
    if (tid == 0) { *s_cnt = 0; *s_tau = __int_as_float(0xff800000); }
    const int32_t* hs = st_hsum(a.state, cur, B, g.DP);
    if (tid < 32) s_cb[tid] = tid < Bcur ? (uint32_t)__ldg(a.dl4 + (hash_from_sum(hs[tid]) - 1)) : 0u;

    const float4* sa4 = reinterpret_cast<const float4*>(st_arr(a.state, 4, g.DP));
    const float4* A4 = reinterpret_cast<const float4*>(st_arr(a.state, 5, g.DP));
    const float4* E4 = reinterpret_cast<const float4*>(st_arr(a.state, 6, g.DP));
    const float4* M4 = reinterpret_cast<const float4*>(st_arr(a.state, 7, g.DP));
    const float4* beams4 = reinterpret_cast<const float4*>(st_beams(a.state, cur, B, g.DP));
    const TfStream st = tf_stream_seeded(hdr->seed + a.t, hdr->seed + a.t);
    const int lg = lane & (g.P - 1);
    const char* T2b = reinterpret_cast<const char*>(s_T2);
    constexpr int HB = R2Pass<BMAX>::HB;

    // contiguous range of sample groups per CTA
    const int64_t nsg = (a.s_end - a.s_begin + g.SPW - 1) / g.SPW;
    const int64_t per = (nsg + gridDim.x - 1) / gridDim.x;
    const int64_t sg0 = (int64_t)blockIdx.x * per, sg1 = min(nsg, sg0 + per);
    __syncthreads();

    const int64_t s_hi64 = min(a.s_end, a.s_begin + sg1 * g.SPW);
    const int s_hi = (int)s_hi64;
    for (int64_t base = sg0; base < sg1; base += (int64_t)nwarps * GP2_NS) {
        uint64_t jb[GP2_NS];
        uint32_t row[GP2_NS];
#pragma unroll
        for (int k = 0; k < GP2_NS; ++k) {
            const int64_t s = a.s_begin + (base + warp + (int64_t)k * nwarps) * g.SPW + lane / g.P;
            const uint64_t s_eff = (uint64_t)(s < s_hi64 ? s : a.s_begin);          // groups beyond the range: clamped, not stored
            jb[k] = s_eff * (uint64_t)g.D + (uint64_t)(32 * lg);
            row[k] = 0u;
        }
        const int s_first = (int)(a.s_begin + (base + warp) * g.SPW + lane / g.P);
        const float tau = *s_tau;
#pragma unroll 1
        for (int boff = 0; boff < BMAX; boff += HB) {
            if (boff >= Bcur) break;
            float acc[GP2_NS][HB];
#pragma unroll
            for (int k = 0; k < GP2_NS; ++k)
#pragma unroll
                for (int b = 0; b < HB; ++b) acc[k][b] = 0.f;
            r2_score_chunk<HB, GP2_NS, false, GP2_SPREAD, false, true>(T2b, s_dl4, s_cb + boff, sa4, A4, E4, M4, beams4 + boff * (g.DP >> 2),
                                                                       g.DP >> 2, g.P, lg, st, jb, nullptr, row, acc);
            float v[GP2_NS * HB];
#pragma unroll
            for (int k = 0; k < GP2_NS; ++k)
#pragma unroll
                for (int b = 0; b < HB; ++b) v[k * HB + b] = acc[k][b];
            const Gp2Sink sink{ s_csc, s_cid, s_cnt, tau, Bcur, boff, s_hi };
            r2_tree_store<GP2_NS * HB, GP2_NS * HB, HB, 0, Gp2Sink>(v, g.P, lane, s_first, nwarps * g.SPW, sink);
        }
        __syncthreads();
        const int cnt = *s_cnt;
        if (cnt > GP2_SLACK) {                     // room for the next full round must remain: cap = round + GP2_SLACK
            const int Kout = block_topk(s_csc, s_cid, cnt, B, s_wsc, s_wid, s_gmax, s_list, TOPK_CAP, s_ctl);
            if (tid < Kout) { s_csc[tid] = s_wsc[tid]; s_cid[tid] = s_wid[tid]; }
            if (tid == 0) { *s_cnt = Kout; if (Kout == B) *s_tau = s_wsc[Kout - 1]; }
            __syncthreads();
        }
    }
    __syncthreads();
    const int cnt = *s_cnt;
    const int Kout = block_topk(s_csc, s_cid, cnt, B, s_wsc, s_wid, s_gmax, s_list, TOPK_CAP, s_ctl);
    if (tid < Kout) {
        irec_record_t r;
        const int f = s_wid[tid];
        r.score = s_wsc[tid]; r.s = f / Bcur; r.b = f - r.s * Bcur; r.pad = 0;
        a.out_rec[(size_t)blockIdx.x * B + tid] = r;
    }
    if (tid == 0) a.out_cnt[blockIdx.x] = Kout;
}

// merge: records laid out as n_lists lists of stride `stride`, list i has cnt[i] valid entries
// (cnt == nullptr: every list is full with `stride` entries... use n_valid).  Output best B, sorted.
__global__ void __launch_bounds__(256) k_topb_merge(const irec_record_t* __restrict__ rec, const int32_t* __restrict__ cnt,
                                                    int n_lists, int stride, const int32_t* __restrict__ bcur_ptr, int bcur_val,
                                                    int B, irec_record_t* out_rec, int32_t* out_cnt, float* g_sc, int32_t* g_id)
{
    const int Bcur = bcur_ptr ? *bcur_ptr : bcur_val;
    __shared__ float s_gmax[256];
    __shared__ float s_wsc[32];
    __shared__ int32_t s_wid[32];
    __shared__ int32_t s_list[TOPK_CAP];
    __shared__ int32_t s_ctl[4];
    __shared__ int32_t s_n;
    const int tid = threadIdx.x, nt = blockDim.x;
    // compact valid records into (g_sc, g_id) -- order is irrelevant for the result
    if (tid == 0) s_n = 0;
    __syncthreads();
    for (int i = tid; i < n_lists * stride; i += nt) {
        const int li = i / stride, e = i - li * stride;
        const int c = cnt ? cnt[li] : stride;
        if (e < c) {
            const int pos = atomicAdd(&s_n, 1);
            g_sc[pos] = rec[i].score;
            g_id[pos] = rec[i].s * Bcur + rec[i].b;
        }
    }
    __syncthreads();
    const int n = s_n;
    const int Kout = block_topk(g_sc, g_id, n, B, s_wsc, s_wid, s_gmax, s_list, TOPK_CAP, s_ctl);
    if (tid < Kout) {
        irec_record_t r;
        const int f = s_wid[tid];
        r.score = s_wsc[tid]; r.s = f / Bcur; r.b = f - r.s * Bcur; r.pad = 0;
        out_rec[tid] = r;
    }
    if (tid == 0) *out_cnt = Kout;
}

// commit: winners -> new beams (other buffer), hash sums, history; flips the state
__global__ void k_gp_commit(void* state, int t, const irec_record_t* __restrict__ win, const int32_t* __restrict__ n_win,
                            const float* __restrict__ T)
{
    BeamStateHdr* hdr = reinterpret_cast<BeamStateHdr*>(state);
    if (hdr->status != IREC_BLK_OK || t >= hdr->n_aux) return;
    const int D = hdr->D, P = hdr->P, DP = hdr->DP, B = hdr->B, cur = hdr->cur;
    const int K = *n_win;
    const float4* sa4 = reinterpret_cast<const float4*>(st_arr(state, 4, DP));
    const float4* old4 = reinterpret_cast<const float4*>(st_beams(state, cur, B, DP));
    float4* new4 = reinterpret_cast<float4*>(st_beams(state, cur ^ 1, B, DP));
    const int32_t* hs = st_hsum(state, cur, B, DP);
    int32_t* hs_new = st_hsum(state, cur ^ 1, B, DP);
    int2* hist = st_hist(state, B, DP);
    const TfStream st = tf_stream_seeded(hdr->seed + t, hdr->seed + t);
    const int nq = DP >> 2;
    const int gtid = blockIdx.x * blockDim.x + threadIdx.x;
    for (int task = gtid; task < nq * K; task += gridDim.x * blockDim.x) {
        const int qq = task / K, j = task - qq * K;
        const int slot = qq / (8 * P), rem = qq - slot * 8 * P;
        const int iq = rem / P, l = rem - iq * P;
        const int d0 = slot * 32 * P + 32 * l + 4 * iq;
        const int sj = win[j].s, bj = win[j].b;
        float4 o = make_float4(0.f, 0.f, 0.f, 0.f);
        if (d0 < D) {
            const uint32_t h = (uint32_t)hash_from_sum(hs[bj]);
            const uint4 u = tf_stream_quad_at(st, (uint64_t)sj * (uint64_t)D + (uint64_t)d0);
            const float4 sa = sa4[qq];
            const float4 ob = old4[bj * nq + qq];
            o.x = __fadd_rn(ob.x, __fmul_rn(T[beam_mix(beam_r_from_u32(u.x), h)], sa.x));
            o.y = __fadd_rn(ob.y, __fmul_rn(T[beam_mix(beam_r_from_u32(u.y), h)], sa.y));
            o.z = __fadd_rn(ob.z, __fmul_rn(T[beam_mix(beam_r_from_u32(u.z), h)], sa.z));
            o.w = __fadd_rn(ob.w, __fmul_rn(T[beam_mix(beam_r_from_u32(u.w), h)], sa.w));
        }
        new4[j * nq + qq] = o;
    }
    if (gtid < K) {
        hist[(size_t)t * 32 + gtid] = make_int2(win[gtid].s, win[gtid].b);
        hs_new[gtid] = hsum_extend(hs[win[gtid].b], win[gtid].s, t);
    }
}
// flips cur / Bcur after commit (separate tiny launch so that every CTA of k_gp_commit saw the old header)
__global__ void k_gp_flip(void* state, int t, const int32_t* __restrict__ n_win)
{
    BeamStateHdr* hdr = reinterpret_cast<BeamStateHdr*>(state);
    if (hdr->status != IREC_BLK_OK || t >= hdr->n_aux) return;
    hdr->cur ^= 1;
    hdr->Bcur = *n_win;
}

// finish: indices (back-pointer trace) + sample of the best beam
__global__ void k_gp_finish(void* state, const int64_t* __restrict__ gidx, int64_t off, int32_t* out_indices,
                            int32_t* out_n_aux, int32_t* out_status, float* out_sample)
{
    BeamStateHdr* hdr = reinterpret_cast<BeamStateHdr*>(state);
    const int D = hdr->D, P = hdr->P, DP = hdr->DP, B = hdr->B;
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        if (out_n_aux) *out_n_aux = hdr->n_aux;
        if (out_status) *out_status = hdr->status;
        if (hdr->status == IREC_BLK_OK) {
            const int2* hist = st_hist(state, B, DP);
            int j = 0;
            for (int t = hdr->n_aux - 1; t >= 0; --t) {
                const int2 e = hist[(size_t)t * 32 + j];
                out_indices[t] = e.x;
                j = e.y;
            }
        }
    }
    if (hdr->status != IREC_BLK_OK) return;
    const float* beam0 = st_beams(state, hdr->cur, B, DP);
    const float* plc = st_arr(state, 8, DP);
    for (int d = blockIdx.x * blockDim.x + threadIdx.x; d < D; d += gridDim.x * blockDim.x) {
        const int64_t gi = gidx ? gidx[off + d] : off + d;
        const int ci = ci_index(d, P);
        out_sample[gi] = __fadd_rn(beam0[ci], plc[ci]);
    }
}

// =============================================================================================
// K5 fused: ALL auxiliary variables of one coder-block in ONE persistent launch, candidates spread over the grid and,
// across GPUs, over the ranks (beam_search_coder.py:66-109 with the argsort at :86 split at every level).
// The multi-launch path above spends ~65 us per auxiliary variable in seven dependent small launches (params, score,
// merge, exchange, merge, commit, flip) -- more than a GPU's share of the scoring below S ~ 10^6.  Here every CTA keeps
// its OWN replica of the block state in shared memory (schedule inputs, coefficients, both beam buffers, hash sums; the
// per-variable schedule and the winners' re-materialisation are computed redundantly by every CTA: a few thousand
// flops), so one variable needs a single all-to-all point:
//   score own sample groups -> CTA top-B -> publish records -> arrive
//   the LAST CTA to arrive merges the grid's lists (rank top-B), exchanges it with the peer ranks (stores into their
//   exchange buffers over NVLink + sequence flags, as k_p2p_exchange), merges the world's lists, publishes the winners
//   every CTA waits for the winners (acquire spin on a sequence word), commits them into its replica, goes on.
// Cooperative launch (all CTAs co-resident: the spin waits cannot deadlock).  Limits: the replica must fit beside the
// 120 KB quantile table -- D <= 256 at 20 beams; larger blocks take the multi-launch path.
// =============================================================================================
#define P2P_FLAG_OFF 64            // layout of a rank's exchange buffer: see k_p2p_exchange below
#define P2P_SLOT_OFF 128
struct FusedArgs {
    void* state;                     // BeamState (k_gp_init has filled it); receives the final beams / history / header
    const float* T2; const uint16_t* dl4; const float* ratio_tab;
    int64_t s_begin, s_end;          // this rank's candidate range
    irec_record_t* lists; int32_t* list_cnt;      // [grid][B], [grid]
    float* g_sc; int32_t* g_id;      // merge scratch [(grid + world) * B]
    irec_record_t* win; int32_t* nwin;            // [2][32], [2]   (parity = t & 1)
    int32_t* sync;                   // [0] arrivals, [1] published sequence (both zeroed before the launch)
    int32_t* const* peer_bufs; int rank, world;   // nullptr / 0 / 1 on a single GPU
    int cand_cap;
    long long* prof;                 // diagnostics (IREC_GP_PROFILE=1): [grid][8] phase cycles of thread 0, or nullptr
};
#define GPF_TICK(slot)                                                        \
    do {                                                                      \
        if (a.prof && tid == 0) {                                             \
            const long long now_ = clock64();                                 \
            a.prof[(size_t)blockIdx.x * 8 + (slot)] += now_ - prof_t;         \
            prof_t = now_;                                                    \
        }                                                                     \
    } while (0)

template <int BMAX>
__global__ void __launch_bounds__(GP2_THREADS, 1) k_gp_fused(const FusedArgs a)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    BeamStateHdr* hdr = reinterpret_cast<BeamStateHdr*>(a.state);
    const int tid = threadIdx.x, nt = blockDim.x, lane = tid & 31, warp = tid >> 5, nwarps = nt >> 5;
    if (hdr->status != IREC_BLK_OK) return;
    BeamGeom g;
    g.D = hdr->D; g.P = hdr->P; g.nslots = hdr->nslots; g.DP = hdr->DP; g.nch = (g.D + 31) >> 5; g.SPW = 32 / g.P;
    const int B = hdr->B, n_aux = hdr->n_aux, DP = g.DP;
    const int64_t seed = hdr->seed;
    const int cap = a.cand_cap;
    constexpr int HB = R2Pass<BMAX>::HB;

    float* s_T2 = reinterpret_cast<float*>(smem_raw);                // [IREC_T2_LEN]
    float* s_cv = s_T2 + IREC_T2_LEN;                                // replica: cv, tv, dmu, cum, sa, A, E, M  [8][DP]
    float* s_tv = s_cv + DP; float* s_dmu = s_tv + DP; float* s_cum = s_dmu + DP;
    float* s_sa = s_cum + DP; float* s_A = s_sa + DP; float* s_E = s_A + DP; float* s_M = s_E + DP;
    float* s_beams = s_M + DP;                                       // [2][BMAX][DP]
    float* s_csc = s_beams + 2 * BMAX * DP;                          // [cap]
    int32_t* s_cid = reinterpret_cast<int32_t*>(s_csc + cap);        // [cap]
    float* s_gmax = reinterpret_cast<float*>(s_cid + cap);           // [GP2_THREADS]
    float* s_wsc = s_gmax + GP2_THREADS;                             // [32]
    int32_t* s_wid = reinterpret_cast<int32_t*>(s_wsc + 32);         // [32]
    int32_t* s_list = s_wid + 32;                                    // [TOPK_CAP]
    int32_t* s_ctl = s_list + TOPK_CAP;                              // [4]
    int32_t* s_cnt = s_ctl + 4;                                      // [1] candidate count
    float* s_tau = reinterpret_cast<float*>(s_cnt + 1);              // [1]
    int32_t* s_flag = reinterpret_cast<int32_t*>(s_tau + 1);         // [2] last-arriver flag, exchange sequence
    uint32_t* s_cb = reinterpret_cast<uint32_t*>(s_flag + 2);        // [32] 4 * dlog(h_b)
    int32_t* s_hsum = reinterpret_cast<int32_t*>(s_cb + 32);         // [2][32]
    irec_record_t* s_win = reinterpret_cast<irec_record_t*>(s_hsum + 64);   // [32]
    uint16_t* s_dl4 = reinterpret_cast<uint16_t*>(s_win + 32);       // [GP_DL4_LEN] the discrete-log table (r2_exp4<true>)

    {
        const float4* src = reinterpret_cast<const float4*>(a.T2);
        float4* dst = reinterpret_cast<float4*>(s_T2);
        for (int i = tid; i < IREC_T2_LEN / 4; i += nt) dst[i] = src[i];
        const uint32_t* dsrc = reinterpret_cast<const uint32_t*>(a.dl4);
        uint32_t* ddst = reinterpret_cast<uint32_t*>(s_dl4);
        for (int i = tid; i < GP_DL4_LEN / 2; i += nt) ddst[i] = dsrc[i];
    }
    for (int i = tid; i < DP; i += nt) {
        s_cv[i] = st_arr(a.state, 0, DP)[i]; s_tv[i] = st_arr(a.state, 1, DP)[i]; s_dmu[i] = st_arr(a.state, 2, DP)[i];
        s_cum[i] = 0.f; s_sa[i] = 0.f; s_A[i] = 0.f; s_E[i] = 0.f; s_M[i] = 0.f;
    }
    for (int i = tid; i < 2 * BMAX * DP; i += nt) s_beams[i] = 0.f;
    if (tid < 64) s_hsum[tid] = 0;
    __syncthreads();

    const int lg = lane & (g.P - 1);
    const char* T2b = reinterpret_cast<const char*>(s_T2);
    const int nq = DP >> 2;
    // contiguous range of sample groups per CTA
    const int64_t nsg = (a.s_end - a.s_begin + g.SPW - 1) / g.SPW;
    const int64_t per = (nsg + gridDim.x - 1) / gridDim.x;
    const int64_t sg0 = (int64_t)blockIdx.x * per, sg1 = min(nsg, sg0 + per);
    const int64_t s_hi64 = min(a.s_end, a.s_begin + sg1 * g.SPW);
    const int s_hi = (int)s_hi64;
    int Bcur = 1, cur = 0;
    long long prof_t = a.prof ? clock64() : 0;

    for (int t = 0; t < n_aux; ++t) {
        // ---- schedule of this auxiliary variable (every CTA for itself) ----
        const float ratio = a.ratio_tab[n_aux - 1 - t];
        for (int i = tid; i < DP; i += nt) {
            const float c = s_cv[i];
            if (c != 0.f) {
                const SchedOut o = beam_sched_dim(c, s_tv[i], s_dmu[i], s_cum[i], ratio);
                s_sa[i] = o.sa; s_A[i] = o.A; s_E[i] = o.E; s_M[i] = o.M; s_cum[i] = o.cum_next;
            }
        }
        const int32_t* hs = s_hsum + 32 * cur;
        if (tid < 32) s_cb[tid] = tid < Bcur ? (uint32_t)__ldg(a.dl4 + (hash_from_sum(hs[tid]) - 1)) : 0u;
        if (tid == 0) { *s_cnt = 0; *s_tau = __int_as_float(0xff800000); }
        __syncthreads();
        GPF_TICK(0);                                                 // schedule
        const float4* sa4 = reinterpret_cast<const float4*>(s_sa);
        const float4* A4 = reinterpret_cast<const float4*>(s_A);
        const float4* E4 = reinterpret_cast<const float4*>(s_E);
        const float4* M4 = reinterpret_cast<const float4*>(s_M);
        const float4* beams4 = reinterpret_cast<const float4*>(s_beams + (size_t)cur * BMAX * DP);
        const TfStream st = tf_stream_seeded(seed + t, seed + t);

        // ---- score this CTA's sample groups, keep its best B (as k_gp_score_topb2).  A warp whose first group of a round lies
        //      beyond the CTA's range only joins the barriers: with a GPU's share of ~30 groups per CTA (S = 6e5 over 8 GPUs)
        //      a round of clamped duplicates would double the work ----
        for (int64_t base = sg0; base < sg1; base += (int64_t)nwarps * GP2_NS) {
            const float tau = *s_tau;
            // A warp scores GP2_NS sample groups of the round when all of them exist, otherwise only its first one (the tail
            // round of a variable: with ~30 groups per CTA -- S = 6e5 over 8 GPUs -- scoring a clamped duplicate as the second
            // group doubled the tail round's work)
            auto score_round = [&](auto ns_c) {
                constexpr int NS = decltype(ns_c)::value;
                uint64_t jb[NS];
                uint32_t row[NS];
#pragma unroll
                for (int k = 0; k < NS; ++k) {
                    const int64_t s = a.s_begin + (base + warp + (int64_t)k * nwarps) * g.SPW + lane / g.P;
                    const uint64_t s_eff = (uint64_t)(s < s_hi64 ? s : a.s_begin);          // samples beyond the range: clamped, not stored
                    jb[k] = s_eff * (uint64_t)g.D + (uint64_t)(32 * lg);
                    row[k] = 0u;
                }
                const int s_first = (int)(a.s_begin + (base + warp) * g.SPW + lane / g.P);
#pragma unroll 1
                for (int boff = 0; boff < BMAX; boff += HB) {
                    if (boff >= Bcur) break;
                    float acc[NS][HB];
#pragma unroll
                    for (int k = 0; k < NS; ++k)
#pragma unroll
                        for (int b = 0; b < HB; ++b) acc[k][b] = 0.f;
                    r2_score_chunk<HB, NS, false, GP2_SPREAD, false, true>(T2b, s_dl4, s_cb + boff, sa4, A4, E4, M4, beams4 + boff * nq, nq,
                                                                           g.P, lg, st, jb, nullptr, row, acc);
                    float v[NS * HB];
#pragma unroll
                    for (int k = 0; k < NS; ++k)
#pragma unroll
                        for (int b = 0; b < HB; ++b) v[k * HB + b] = acc[k][b];
                    const Gp2Sink sink{ s_csc, s_cid, s_cnt, tau, Bcur, boff, s_hi };
                    r2_tree_store<NS * HB, NS * HB, HB, 0, Gp2Sink>(v, g.P, lane, s_first, nwarps * g.SPW, sink);
                }
            };
            if (base + warp + (int64_t)(GP2_NS - 1) * nwarps < sg1) score_round(std::integral_constant<int, GP2_NS>{});
            else if (base + warp < sg1) score_round(std::integral_constant<int, 1>{});
            __syncthreads();
            const int cnt = *s_cnt;
            if (cnt > GP2_SLACK) {                 // lazy compaction (Gp2Sink)
                const int Kc = block_topk(s_csc, s_cid, cnt, B, s_wsc, s_wid, s_gmax, s_list, TOPK_CAP, s_ctl);
                if (tid < Kc) { s_csc[tid] = s_wsc[tid]; s_cid[tid] = s_wid[tid]; }
                if (tid == 0) { *s_cnt = Kc; if (Kc == B) *s_tau = s_wsc[Kc - 1]; }
                __syncthreads();
            }
        }
        __syncthreads();
        GPF_TICK(1);                                                 // scoring rounds (with their compactions)
        const int Kcta = block_topk(s_csc, s_cid, *s_cnt, B, s_wsc, s_wid, s_gmax, s_list, TOPK_CAP, s_ctl);

        // ---- publish, arrive; the last CTA merges, exchanges, publishes the winners ----
        if (tid < Kcta) {
            irec_record_t r;
            const int f = s_wid[tid];
            r.score = s_wsc[tid]; r.s = f / Bcur; r.b = f - r.s * Bcur; r.pad = 0;
            a.lists[(size_t)blockIdx.x * B + tid] = r;
        }
        if (tid == 0) a.list_cnt[blockIdx.x] = Kcta;
        __threadfence();
        __syncthreads();
        if (tid == 0) {
            const int ticket = atomicAdd(&a.sync[0], 1);
            s_flag[0] = (ticket == (t + 1) * (int)gridDim.x - 1) ? 1 : 0;
        }
        __syncthreads();
        GPF_TICK(2);                                                 // the CTA's own top-B, publish, arrive
        const int parity = t & 1;
        if (s_flag[0]) {
            __threadfence();
            // rank top-B over the grid's lists
            if (tid == 0) *s_cnt = 0;
            __syncthreads();
            // merge scratch: the CTA's own candidate buffer (free at this point) when the lists fit, else global memory
            float* m_sc = ((int)gridDim.x * B <= cap && a.world * B <= cap) ? s_csc : a.g_sc;
            int32_t* m_id = (m_sc == s_csc) ? s_cid : a.g_id;
            // The record loads are issued four rounds at a time BEFORE anything depends on them: the loop used to pay two
            // dependent L2 round trips (count, then record) per round of 384 records -- half of the 11 us this merge took per
            // variable (IREC_GP_PROFILE).  Unpublished slots of a list are skipped by their position (e >= count).
            const int n_rec = (int)gridDim.x * B;
            for (int i0 = 0; i0 < n_rec; i0 += 4 * nt) {              // whole warps stay in the loop (warp_append_pos)
                int4 v[4];
                bool have[4];
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const int i = i0 + k * nt + tid;
                    const int li = i / B, e = i - li * B;
                    have[k] = i < n_rec && e < __ldcg(a.list_cnt + li);
                    v[k] = make_int4(0, 0, 0, 0);
                    if (i < n_rec) v[k] = __ldcg(reinterpret_cast<const int4*>(a.lists) + i);   // (score bits, s, b, pad) straight from L2
                }
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const int pos = warp_append_pos(s_cnt, have[k]);
                    if (pos >= 0) {
                        m_sc[pos] = __int_as_float(v[k].x);
                        m_id[pos] = v[k].y * Bcur + v[k].z;
                    }
                }
            }
            __syncthreads();
            int K = block_topk(m_sc, m_id, *s_cnt, B, s_wsc, s_wid, s_gmax, s_list, TOPK_CAP, s_ctl);
            if (a.world > 1) {
                // exchange the rank lists over peer memory (protocol of k_p2p_exchange), then merge the world's lists
                int32_t* mine = a.peer_bufs[a.rank];
                if (tid == 0) s_flag[1] = mine[0] + 1;
                __syncthreads();
                const int seq = s_flag[1], W = (B + 1) * 4, par = seq & 1;
                for (int i = tid; i < a.world * W; i += nt) {
                    const int r = i / W, w = i - r * W;
                    int32_t val;
                    if (w < B * 4) {
                        const int e = w >> 2, c = w & 3;
                        const int f = e < K ? s_wid[e] : 0;
                        const int sj = f / Bcur;
                        val = e < K ? (c == 0 ? __float_as_int(s_wsc[e]) : (c == 1 ? sj : (c == 2 ? f - sj * Bcur : 0))) : 0;
                    } else {
                        val = (w == B * 4) ? K : 0;
                    }
                    a.peer_bufs[r][P2P_SLOT_OFF + (par * a.world + a.rank) * W + w] = val;
                }
                __threadfence_system();
                __syncthreads();
                if (tid < a.world) {
                    int32_t* flag = a.peer_bufs[tid] + P2P_FLAG_OFF + a.rank;
                    asm volatile("st.release.sys.global.s32 [%0], %1;" ::"l"(flag), "r"(seq) : "memory");
                    const int32_t* want = mine + P2P_FLAG_OFF + tid;
                    int got;
                    do {
                        asm volatile("ld.acquire.sys.global.s32 %0, [%1];" : "=r"(got) : "l"(want) : "memory");
                    } while (got - seq < 0);
                }
                __syncthreads();
                const volatile int32_t* slots = mine + P2P_SLOT_OFF + par * a.world * W;
                if (tid == 0) *s_cnt = 0;
                __syncthreads();
                for (int i = tid; i < a.world * B; i += nt) {
                    const int r = i / B, e = i - r * B;
                    if (e < slots[r * W + B * 4]) {
                        const int pos = atomicAdd(s_cnt, 1);
                        m_sc[pos] = __int_as_float(slots[r * W + e * 4]);
                        m_id[pos] = slots[r * W + e * 4 + 1] * Bcur + slots[r * W + e * 4 + 2];
                    }
                }
                __syncthreads();
                K = block_topk(m_sc, m_id, *s_cnt, B, s_wsc, s_wid, s_gmax, s_list, TOPK_CAP, s_ctl);
                if (tid == 0) mine[0] = seq;
            }
            if (tid < K) {
                irec_record_t r;
                const int f = s_wid[tid];
                r.score = s_wsc[tid]; r.s = f / Bcur; r.b = f - r.s * Bcur; r.pad = 0;
                a.win[parity * 32 + tid] = r;
            }
            if (tid == 0) a.nwin[parity] = K;
            __threadfence();
            __syncthreads();
            if (tid == 0) asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(a.sync + 1), "r"(t + 1) : "memory");
            GPF_TICK(3);                                             // last arriver: merge + exchange + publish
        }
        if (tid == 0) {
            int got;
            do {
                asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(got) : "l"(a.sync + 1) : "memory");
            } while (got < t + 1);
        }
        __syncthreads();
        GPF_TICK(4);                                                 // waiting for the winners
        const int K = __ldcg(a.nwin + parity);
        if (tid < K) {
            const int4 v = __ldcg(reinterpret_cast<const int4*>(a.win + parity * 32) + tid);
            irec_record_t r;
            r.score = __int_as_float(v.x); r.s = v.y; r.b = v.z; r.pad = 0;
            s_win[tid] = r;
        }
        __syncthreads();

        // ---- commit into this CTA's replica: new beams (other buffer), hash sums; CTA 0 keeps the history ----
        {
            const float4* old4 = reinterpret_cast<const float4*>(s_beams + (size_t)cur * BMAX * DP);
            float4* new4 = reinterpret_cast<float4*>(s_beams + (size_t)(cur ^ 1) * BMAX * DP);
            for (int task = tid; task < nq * K; task += nt) {
                const int j = task / nq, qq = task - j * nq;
                const int iq = qq / g.P, l = qq - iq * g.P;
                const int d0 = 32 * l + 4 * iq;
                const int sj = s_win[j].s, bj = s_win[j].b;
                float4 o = make_float4(0.f, 0.f, 0.f, 0.f);
                if (d0 < g.D) {
                    const uint32_t cb = s_cb[bj];
                    const uint4 u = tf_stream_quad_at(st, (uint64_t)sj * (uint64_t)g.D + (uint64_t)d0);
                    const float4 sa = sa4[qq];
                    const float4 ob = old4[bj * nq + qq];
                    o.x = __fadd_rn(ob.x, __fmul_rn(*reinterpret_cast<const float*>(T2b + r2_exp4(a.dl4, u.x) + cb), sa.x));
                    o.y = __fadd_rn(ob.y, __fmul_rn(*reinterpret_cast<const float*>(T2b + r2_exp4(a.dl4, u.y) + cb), sa.y));
                    o.z = __fadd_rn(ob.z, __fmul_rn(*reinterpret_cast<const float*>(T2b + r2_exp4(a.dl4, u.z) + cb), sa.z));
                    o.w = __fadd_rn(ob.w, __fmul_rn(*reinterpret_cast<const float*>(T2b + r2_exp4(a.dl4, u.w) + cb), sa.w));
                    // dims beyond D inside the last quad: sigma_aux = 0 and the parent is 0 there, so the padding stays zero
                }
                new4[j * nq + qq] = o;
            }
            int32_t* hs_new = s_hsum + 32 * (cur ^ 1);
            if (tid < K) {
                hs_new[tid] = hsum_extend(hs[s_win[tid].b], s_win[tid].s, t);
                if (blockIdx.x == 0) st_hist(a.state, B, DP)[(size_t)t * 32 + tid] = make_int2(s_win[tid].s, s_win[tid].b);
            }
        }
        __syncthreads();
        GPF_TICK(5);                                                 // commit
        Bcur = K;
        cur ^= 1;
    }
    // ---- CTA 0 hands the final state back (irec_beam_state_finish reads it) ----
    if (blockIdx.x == 0) {
        float* gb = st_beams(a.state, cur, B, DP);
        for (int i = tid; i < B * DP; i += nt) gb[i] = s_beams[(size_t)cur * BMAX * DP + i];
        if (tid == 0) { hdr->cur = cur; hdr->Bcur = Bcur; }
    }
}

// schedule export (include/irec.h: irec_schedule): per-partition per-dim coefficients of ONE block, exactly what the encode
// kernels compute on the fly (beam_sched_dim); thread d walks the auxiliary variables of its dim
__global__ void k_schedule(const float* __restrict__ t_loc, const float* __restrict__ t_scale, const float* __restrict__ p_loc,
                           const float* __restrict__ p_scale, int D, int max_aux, const int32_t* __restrict__ n_aux_ptr,
                           const float* __restrict__ ratio_tab, int ratio_len, float* __restrict__ out_sa, float* __restrict__ out_A,
                           float* __restrict__ out_E, float* __restrict__ out_M)
{
    const int n_aux = *n_aux_ptr;
    if (n_aux <= 0 || n_aux > max_aux || n_aux > ratio_len) return;
    for (int d = blockIdx.x * blockDim.x + threadIdx.x; d < D; d += gridDim.x * blockDim.x) {
        const float ps = p_scale[d], ts = t_scale[d];
        const float cv = __fmul_rn(ps, ps), tv = __fmul_rn(ts, ts), dmu = __fadd_rn(t_loc[d], -p_loc[d]);
        float cum = 0.f;
        for (int t = 0; t < n_aux; ++t) {
            SchedOut o;
            o.sa = 0.f; o.A = 0.f; o.E = 0.f; o.M = 0.f; o.cum_next = cum;
            if (cv != 0.f) o = beam_sched_dim(cv, tv, dmu, cum, ratio_tab[n_aux - 1 - t]);
            out_sa[(size_t)t * D + d] = o.sa; out_A[(size_t)t * D + d] = o.A;
            out_E[(size_t)t * D + d] = o.E; out_M[(size_t)t * D + d] = o.M;
            cum = o.cum_next;
        }
    }
}

// raw stream (tests)
__global__ void k_beam_uniform_int(int64_t q, int64_t start, int64_t n, int32_t* out)
{
    const TfStream st = tf_stream_seeded(q, q);
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const uint64_t j = (uint64_t)(start + i);
        const uint4 u = tf_stream_group(st, j >> 2);
        const uint32_t v[4] = { u.x, u.y, u.z, u.w };
        out[i] = (int32_t)beam_r_from_u32(v[j & 3]);
    }
}

// tf.random.uniform(shape, lo, hi, dtype=int32, seed=op_seed) after tf.random.set_seed(global_seed): lo + u32 % (hi - lo)
__global__ void k_uniform_int_stream(TfStream st, uint32_t lo, uint32_t range, int64_t start, int64_t n, int32_t* out)
{
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const uint64_t j = (uint64_t)(start + i);
        const uint4 u = tf_stream_group(st, j >> 2);
        const uint32_t v[4] = { u.x, u.y, u.z, u.w };
        out[i] = (int32_t)(lo + v[j & 3] % range);
    }
}

// =============================================================================================
// host side
// =============================================================================================
template <int BMAX>
static size_t resident_smem_bytes(int DPmax, int NC)
{
    size_t fl = 10008 + (size_t)BMAX * DPmax + (size_t)8 * DPmax + NC + 1024 + 32;
    size_t in = 32 + TOPK_CAP + 4 + 64 + 4;
    return 32 * sizeof(double) + fl * sizeof(float) + in * sizeof(int32_t) + 16;
}

static int pick_bmax(int B)
{
    const int opts[] = { 1, 2, 4, 8, 10, 16, 20, 32 };
    for (int o : opts) if (B <= o) return o;
    return -1;
}

struct ResidentPlan {
    bool ok; int bmax; int nthreads; size_t smem; int DPmax; int NC; int grid;
};

template <int BMAX>
static ResidentPlan plan_resident_t(int nb, int max_D, int S, int B)
{
    ResidentPlan p{};
    p.ok = false; p.bmax = BMAX;
    if (max_D > 1024 || (int64_t)S * BMAX > 32768) return p;
    const BeamGeom g = make_geom(max_D);
    p.DPmax = g.DP;
    p.NC = ((S * BMAX + 31) / 32) * 32;
    p.smem = resident_smem_bytes<BMAX>(p.DPmax, p.NC);
    const IrecDevice& dev = irec_device();
    if (p.smem > (size_t)dev.max_smem_optin) return p;
    cudaFuncAttributes fa;
    if (cudaFuncGetAttributes(&fa, k_beam_encode_resident<BMAX>) != cudaSuccess) return p;
    int max_threads = (65536 / (fa.numRegs > 0 ? fa.numRegs : 64)) / 32 * 32;
    if (max_threads > fa.maxThreadsPerBlock) max_threads = fa.maxThreadsPerBlock / 32 * 32;
    if (max_threads > 1024) max_threads = 1024;
    if (max_threads < 64) return p;
    // one sample group per warp and round; pick the warp count that balances the rounds
    const int nsg = (S + g.SPW - 1) / g.SPW;
    const int maxw = max_threads / 32;
    const int rounds = (nsg + maxw - 1) / maxw;
    int nw = (nsg + rounds - 1) / rounds;
    if (nw < 8) nw = std::min(8, maxw);           // keep enough threads for the elementwise phases
    p.nthreads = nw * 32;
    if (cudaFuncSetAttribute(k_beam_encode_resident<BMAX>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)p.smem) != cudaSuccess)
        return p;
    p.grid = std::min(nb, dev.sm_count);
    p.ok = true;
    return p;
}

static ResidentPlan plan_resident(int nb, int max_D, int S, int B)
{
    switch (pick_bmax(B)) {
        case 1: return plan_resident_t<1>(nb, max_D, S, B);
        case 2: return plan_resident_t<2>(nb, max_D, S, B);
        case 4: return plan_resident_t<4>(nb, max_D, S, B);
        case 8: return plan_resident_t<8>(nb, max_D, S, B);
        case 10: return plan_resident_t<10>(nb, max_D, S, B);
        case 16: return plan_resident_t<16>(nb, max_D, S, B);
        case 20: return plan_resident_t<20>(nb, max_D, S, B);
        case 32: return plan_resident_t<32>(nb, max_D, S, B);
    }
    ResidentPlan p{}; p.ok = false; return p;
}

template <int BMAX>
static void launch_resident_t(const ResidentPlan& p, const ResidentArgs& a, cudaStream_t s)
{
    k_beam_encode_resident<BMAX><<<p.grid, p.nthreads, p.smem, s>>>(a);
    irec_count_launch();
}
static void launch_resident(const ResidentPlan& p, const ResidentArgs& a, cudaStream_t s)
{
    switch (p.bmax) {
        case 1: launch_resident_t<1>(p, a, s); break;
        case 2: launch_resident_t<2>(p, a, s); break;
        case 4: launch_resident_t<4>(p, a, s); break;
        case 8: launch_resident_t<8>(p, a, s); break;
        case 10: launch_resident_t<10>(p, a, s); break;
        case 16: launch_resident_t<16>(p, a, s); break;
        case 20: launch_resident_t<20>(p, a, s); break;
        case 32: launch_resident_t<32>(p, a, s); break;
    }
}

// ---- second-generation resident kernel (irec_resident2.cuh) ----
// workspace: [0,256) block-queue counter | [256,512) R2Plan | back-pointer history | schedule scratch | block order | exponent table
#define R2_TABLE_MAX_BYTES ((size_t)1 << 30)
static size_t r2_ws_hist_bytes(int bmax, int max_aux)
{
    // 2 x: k_beam_encode_tmem keeps two coder-block contexts per CTA
    const size_t h = 512 + sizeof(int2) * 2 * (size_t)irec_device().sm_count * (size_t)max_aux * bmax;
    return (h + 255) / 256 * 256;
}
static size_t r2_ws_order_bytes(int nb) { return ((size_t)nb * sizeof(int32_t) + 255) / 256 * 256; }
static size_t r2_ws_sched_bytes()
{
    return sizeof(float) * 2 * 4 * 1024 * (size_t)irec_device().sm_count;      // [grid][(2 contexts)][4][DPmax <= 1024]
}
// ---- library-owned exponent table, reused across launches (R2TabKey) ----
// One table per device, shared by every stream.  Cross-stream protocol (host side under the cache mutex, which is HELD from
// r2_tab_acquire to r2_tab_release, i.e. while one launch is being enqueued):
//   * rows are append-only while the key (block sizes) stands, so launches of different streams may read the table
//     concurrently and may append concurrently (same values);
//   * `ev_w` is recorded after the table-building kernels of every launch; a launch on another stream waits for the latest
//     one (it precedes that stream's long main kernel, so this does not serialise the main kernels);
//   * `rd[]` holds one event per recently used stream, recorded after the main kernel.  The table is REWRITTEN (host-known
//     parameters changed: the stream first waits for every other reader; block sizes changed: decided on the device, allowed
//     only if all other readers are known to have finished -- otherwise that launch builds a private table in its own
//     workspace, R2_TAB_PRIVATE) only when nobody else can be reading it.
#define R2_CACHE_MAX_BYTES ((size_t)256 << 20)
#define R2_MAX_READERS 8
struct R2TabCache {
    std::mutex mu;
    void* buf = nullptr; size_t bytes = 0;
    R2TabKey* key = nullptr;                       // device
    int64_t seed = 0; int S = 0, row_stride = 0, cap_aux = 0;
    bool have_params = false;
    cudaEvent_t ev_w = nullptr; cudaStream_t last_w = nullptr; bool ev_w_valid = false;
    struct Reader { cudaStream_t s = nullptr; cudaEvent_t ev = nullptr; bool valid = false; } rd[R2_MAX_READERS];
};
static R2TabCache g_tab_cache[64];

struct R2TabUse { uint2* tab; int tab_aux; R2TabKey* key; R2TabCache* cache; uint2* priv; int allow_rekey; };

// Optional (IREC_R2_L2_WINDOW=1): keep the shared exponent table resident in L2 through a persisting access-policy window on the
// launching stream.  Built to test whether the 1.0 GB of DRAM reads per configs[3]-size launch were table misses -- they were not
// (966 MB with the window; the cause was the queue order, k_tm_order) -- and OFF by default: the 64 MB set-aside is taken from every
// other user of L2 (the importance sampler streams a 38 MB candidate table).
static void r2_tab_l2_window(cudaStream_t s, const void* ptr, size_t bytes)
{
    static int state[64] = { 0 };                  // per device: 0 = not tried, 1 = on, -1 = unavailable / off
    static size_t max_window[64] = { 0 };
    int d = 0;
    if (cudaGetDevice(&d) != cudaSuccess || d < 0 || d >= 64) return;
    if (state[d] == 0) {
        state[d] = -1;
        const char* e = getenv("IREC_R2_L2_WINDOW");
        cudaDeviceProp prop;
        if ((e && e[0] == '1') && cudaGetDeviceProperties(&prop, d) == cudaSuccess && prop.persistingL2CacheMaxSize > 0 &&
            prop.accessPolicyMaxWindowSize > 0) {
            const size_t want = std::min<size_t>((size_t)prop.persistingL2CacheMaxSize, (size_t)64 << 20);
            if (cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, want) == cudaSuccess) {
                state[d] = 1;
                max_window[d] = std::min<size_t>(want, (size_t)prop.accessPolicyMaxWindowSize);
            }
        }
        cudaGetLastError();
    }
    if (state[d] != 1 || !ptr || !bytes) return;
    cudaStreamAttrValue attr;
    memset(&attr, 0, sizeof(attr));
    attr.accessPolicyWindow.base_ptr = const_cast<void*>(ptr);
    attr.accessPolicyWindow.num_bytes = std::min(bytes, max_window[d]);
    attr.accessPolicyWindow.hitRatio = 1.0f;
    attr.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
    attr.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
    if (cudaStreamSetAttribute(s, cudaStreamAttributeAccessPolicyWindow, &attr) != cudaSuccess) cudaGetLastError();
}

static bool r2_event_done(cudaEvent_t ev)
{
    const cudaError_t e = cudaEventQuery(ev);
    if (e != cudaSuccess) cudaGetLastError();      // cudaErrorNotReady is not an error of ours
    return e == cudaSuccess;
}

// table for this launch: the per-device cache when it can be used (capacity, not capturing, IREC_R2_NO_CACHE unset),
// otherwise `ws_tab` inside the caller's workspace.  On success with a cache the mutex is HELD until r2_tab_release.
static R2TabUse r2_tab_acquire(uint2* ws_tab, int64_t seed, int S, int max_aux, int row_stride, cudaStream_t s)
{
    R2TabUse u{ ws_tab, max_aux, nullptr, nullptr, nullptr, 0 };
    const char* e = getenv("IREC_R2_NO_CACHE");
    if (e && e[0] == '1') return u;
    cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
    if (cudaStreamIsCapturing(s, &cs) != cudaSuccess || cs != cudaStreamCaptureStatusNone) return u;
    int d = 0;
    if (cudaGetDevice(&d) != cudaSuccess || d < 0 || d >= 64) return u;
    R2TabCache& c = g_tab_cache[d];
    const int cap_aux = std::max(128, (max_aux + 63) / 64 * 64);
    const size_t need = sizeof(uint2) * (size_t)R2_MAX_SIZES * (size_t)cap_aux * (size_t)S * (size_t)row_stride;
    if (need > R2_CACHE_MAX_BYTES) return u;
    c.mu.lock();
    bool fresh = false;
    if (!c.key) {
        bool ok = cudaMalloc(&c.key, sizeof(R2TabKey)) == cudaSuccess &&
                  cudaEventCreateWithFlags(&c.ev_w, cudaEventDisableTiming) == cudaSuccess;
        for (int i = 0; ok && i < R2_MAX_READERS; ++i)
            ok = cudaEventCreateWithFlags(&c.rd[i].ev, cudaEventDisableTiming) == cudaSuccess;
        if (!ok) { cudaGetLastError(); c.key = nullptr; c.mu.unlock(); return u; }
        fresh = true;
    }
    if (need > c.bytes) {
        if (c.buf) { cudaDeviceSynchronize(); cudaFree(c.buf); c.buf = nullptr; c.bytes = 0; }
        if (cudaMalloc(&c.buf, need) != cudaSuccess) { cudaGetLastError(); c.have_params = false; c.mu.unlock(); return u; }
        c.bytes = need;
        fresh = true;
    }
    const bool same = c.have_params && c.seed == seed && c.S == S && c.row_stride == row_stride && c.cap_aux >= cap_aux;
    int allow_rekey = 1;
    if (fresh || !same) {
        // host-known parameters changed: the table is rebuilt, so this stream queues behind every other reader
        for (auto& r : c.rd)
            if (r.valid && r.s != s && !r2_event_done(r.ev)) cudaStreamWaitEvent(s, r.ev, 0);
        if (c.ev_w_valid && c.last_w != s) cudaStreamWaitEvent(s, c.ev_w, 0);
        cudaMemsetAsync(c.key, 0, sizeof(R2TabKey), s);
        c.seed = seed; c.S = S; c.row_stride = row_stride; c.cap_aux = cap_aux; c.have_params = true;
    } else {
        if (c.ev_w_valid && c.last_w != s) cudaStreamWaitEvent(s, c.ev_w, 0);      // the table as the latest builder left it
        for (auto& r : c.rd)
            if (r.valid && r.s != s && !r2_event_done(r.ev)) allow_rekey = 0;      // somebody else may be reading
    }
    u.tab = reinterpret_cast<uint2*>(c.buf); u.tab_aux = c.cap_aux; u.key = c.key; u.cache = &c;
    u.priv = ws_tab; u.allow_rekey = allow_rekey;
    r2_tab_l2_window(s, c.buf, need);
    return u;
}
static void r2_tab_release(const R2TabUse& u, cudaStream_t s)
{
    if (!u.cache) return;
    R2TabCache& c = *u.cache;
    R2TabCache::Reader* slot = nullptr;
    for (auto& r : c.rd) if (r.valid && r.s == s) { slot = &r; break; }
    if (!slot) for (auto& r : c.rd) if (!r.valid || r2_event_done(r.ev)) { slot = &r; break; }
    if (!slot) { slot = &c.rd[0]; cudaEventSynchronize(slot->ev); }          // more than R2_MAX_READERS streams in flight
    cudaEventRecord(slot->ev, s);
    slot->s = s; slot->valid = true;
    c.mu.unlock();
}
// exponent rows of this launch (skipping what the cache already holds) + the cache bookkeeping kernel
static void r2_build_table(const R2TabUse& u, R2Plan* dplan, int64_t seed, int S, int max_aux, int row_stride, cudaStream_t s)
{
    const int64_t items = (int64_t)R2_MAX_SIZES * max_aux * S * 8;          // one thread per gather family
    const int grid = (int)std::min<int64_t>((items + 127) / 128, (int64_t)irec_device().sm_count * 32);
    k_r2_exps<<<grid, 128, 0, s>>>(dplan, irec_device().d_dl4, seed, S, max_aux, u.tab_aux, row_stride, u.tab, u.key, u.priv, u.allow_rekey);
    irec_count_launch();
    if (u.key) {
        k_r2_key_commit<<<1, 1, 0, s>>>(dplan, u.key, max_aux, u.allow_rekey);
        irec_count_launch();
        cudaEventRecord(u.cache->ev_w, s);
        u.cache->ev_w_valid = true; u.cache->last_w = s;
    }
}

static size_t r2_table_bytes(int max_D, int S, int max_aux)
{
    if (max_D > 1024) return 0;
    const BeamGeom g = make_geom(max_D);
    const size_t b = sizeof(uint2) * (size_t)R2_MAX_SIZES * (size_t)max_aux * (size_t)S * (size_t)(g.DP >> 2);
    return b <= R2_TABLE_MAX_BYTES ? b : 0;       // very large S * max_aux: exponents are generated in place instead
}

static int resident_choice()
{
    // IREC_RESIDENT=1 forces the first-generation kernel, =2 the second, =3 the tensor-memory kernel (tests / A-B runs);
    // default: 3 where its sizes are covered, else 2 if it fits
    const char* e = getenv("IREC_RESIDENT");
    if (e && e[0] == '1') return 1;
    if (e && e[0] == '2') return 2;
    if (e && e[0] == '3') return 3;
    return 0;
}

static bool r2_no_table()
{
    const char* e = getenv("IREC_R2_NO_TABLE");      // tests: force the in-place Philox exponent path
    return e && e[0] == '1';
}

template <int BMAX>
static ResidentPlan plan_resident2_t(int nb, int max_D, int S, int B)
{
    ResidentPlan p{};
    p.ok = false; p.bmax = BMAX;
    if (max_D > 1024 || (int64_t)S * BMAX > 32768) return p;
    const BeamGeom g = make_geom(max_D);
    p.DPmax = g.DP;
    p.NC = ((S * BMAX + 31) / 32) * 32;
    p.smem = r2_smem_bytes<BMAX>(p.DPmax, p.NC);
    const IrecDevice& dev = irec_device();
    if (p.smem > (size_t)dev.max_smem_optin) return p;
    // sample groups per warp layer: fewest layers with at most 12 warps, then the fewest warps for that
    const int nsg = (S + g.SPW - 1) / g.SPW;
    const int maxw = R2_THREADS / 32;
    const int layers = (nsg + maxw - 1) / maxw;
    int nw = (nsg + layers - 1) / layers;
    if (nw < 8) nw = 8;                           // keep enough threads for the elementwise phases
    if (nw > maxw) nw = maxw;
    p.nthreads = nw * 32;
    if (cudaFuncSetAttribute(k_beam_encode_resident2<BMAX>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)p.smem) != cudaSuccess)
        return p;
    p.grid = std::min(nb, dev.sm_count);
    p.ok = true;
    return p;
}

static int pick_bmax2(int B)
{
    const int opts[] = { 1, 4, 10, 20, 32 };       // fewer instantiations than the first kernel (compile time)
    for (int o : opts) if (B <= o) return o;
    return -1;
}

static ResidentPlan plan_resident2(int nb, int max_D, int S, int B)
{
    switch (pick_bmax2(B)) {
        case 1: return plan_resident2_t<1>(nb, max_D, S, B);
        case 4: return plan_resident2_t<4>(nb, max_D, S, B);
        case 10: return plan_resident2_t<10>(nb, max_D, S, B);
        case 20: return plan_resident2_t<20>(nb, max_D, S, B);
        case 32: return plan_resident2_t<32>(nb, max_D, S, B);
    }
    ResidentPlan p{}; p.ok = false; return p;
}

static void launch_resident2(const ResidentPlan& p, const Resident2Args& a, cudaStream_t s)
{
    switch (p.bmax) {
        case 1: k_beam_encode_resident2<1><<<p.grid, p.nthreads, p.smem, s>>>(a); break;
        case 4: k_beam_encode_resident2<4><<<p.grid, p.nthreads, p.smem, s>>>(a); break;
        case 10: k_beam_encode_resident2<10><<<p.grid, p.nthreads, p.smem, s>>>(a); break;
        case 20: k_beam_encode_resident2<20><<<p.grid, p.nthreads, p.smem, s>>>(a); break;
        case 32: k_beam_encode_resident2<32><<<p.grid, p.nthreads, p.smem, s>>>(a); break;
    }
    irec_count_launch();
}

// ---- general path helpers ----
#define GP_THREADS 256
static size_t gp_score_smem(int cand_cap)
{
    return sizeof(float) * (10008 + (size_t)cand_cap + 256 + 32) + sizeof(int32_t) * ((size_t)cand_cap + 32 + TOPK_CAP + 4 + 2) + 16;
}
static int gp_cand_cap(int D, int bmax)
{
    const BeamGeom g = make_geom(D);
    return (GP_THREADS / 32) * g.SPW * bmax + 64;
}
static int gp_score_grid(int D, int64_t n_samples)
{
    const BeamGeom g = make_geom(D);
    const int64_t nsg = (n_samples + g.SPW - 1) / g.SPW;
    const int64_t per_cta_min = (GP_THREADS / 32);           // at least one round of work per CTA
    int64_t grid = (nsg + per_cta_min - 1) / per_cta_min;
    const int cap = irec_device().sm_count * 4;
    if (grid > cap) grid = cap;
    if (grid < 1) grid = 1;
    return (int)grid;
}

static size_t gp2_score_smem(int cand_cap)
{
    return sizeof(float) * ((size_t)IREC_T2_LEN + (size_t)cand_cap + GP2_THREADS + 32) +
           sizeof(int32_t) * ((size_t)cand_cap + 32 + TOPK_CAP + 4 + 4 + 32) + sizeof(uint16_t) * GP_DL4_LEN + 16;
}
static int gp2_cand_cap(int D, int bmax)
{
    const BeamGeom g = make_geom(D);
    return (GP2_THREADS / 32) * GP2_NS * g.SPW * bmax + GP2_SLACK;
}
static int gp2_score_grid(int D, int64_t n_samples)
{
    const BeamGeom g = make_geom(D);
    const int64_t nsg = (n_samples + g.SPW - 1) / g.SPW;
    int64_t grid = (nsg + (GP2_THREADS / 32) * GP2_NS - 1) / ((GP2_THREADS / 32) * GP2_NS);       // at least one round of work per CTA
    const int cap = irec_device().sm_count;                                   // one CTA per SM (120 KB table + 512 threads)
    if (grid > cap) grid = cap;
    if (grid < 1) grid = 1;
    return (int)grid;
}
static bool gp_use_v2(int D, int bmax)
{
    const char* e = getenv("IREC_GP_V1");
    if (e && e[0] == '1') return false;
    return D <= 1024 && gp2_score_smem(gp2_cand_cap(D, bmax)) <= (size_t)irec_device().max_smem_optin;
}
template <int BMAX>
static int launch_gp2_score_t(const Score2Args& a, int grid, size_t smem, cudaStream_t s)
{
    static bool attr_done = false;
    if (!attr_done) {
        if (cudaFuncSetAttribute(k_gp_score_topb2<BMAX>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 (int)irec_device().max_smem_optin) != cudaSuccess)
            return IREC_E_CUDA;
        attr_done = true;
    }
    k_gp_score_topb2<BMAX><<<grid, GP2_THREADS, smem, s>>>(a);
    irec_count_launch();
    return IREC_OK;
}
static int launch_gp2_score(int bmax, const Score2Args& a, int grid, size_t smem, cudaStream_t s)
{
    switch (bmax) {
        case 1: return launch_gp2_score_t<1>(a, grid, smem, s);
        case 2: return launch_gp2_score_t<2>(a, grid, smem, s);
        case 4: return launch_gp2_score_t<4>(a, grid, smem, s);
        case 8: return launch_gp2_score_t<8>(a, grid, smem, s);
        case 10: return launch_gp2_score_t<10>(a, grid, smem, s);
        case 16: return launch_gp2_score_t<16>(a, grid, smem, s);
        case 20: return launch_gp2_score_t<20>(a, grid, smem, s);
        case 32: return launch_gp2_score_t<32>(a, grid, smem, s);
    }
    return IREC_E_INVALID;
}

template <int BMAX>
static int launch_gp_score_t(const ScoreArgs& a, int grid, size_t smem, cudaStream_t s)
{
    static bool attr_done = false;
    if (!attr_done) {
        if (cudaFuncSetAttribute(k_gp_score_topb<BMAX>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 (int)irec_device().max_smem_optin) != cudaSuccess)
            return IREC_E_CUDA;
        attr_done = true;
    }
    k_gp_score_topb<BMAX><<<grid, GP_THREADS, smem, s>>>(a);
    irec_count_launch();
    return IREC_OK;
}
static int launch_gp_score(int bmax, const ScoreArgs& a, int grid, size_t smem, cudaStream_t s)
{
    switch (bmax) {
        case 1: return launch_gp_score_t<1>(a, grid, smem, s);
        case 2: return launch_gp_score_t<2>(a, grid, smem, s);
        case 4: return launch_gp_score_t<4>(a, grid, smem, s);
        case 8: return launch_gp_score_t<8>(a, grid, smem, s);
        case 10: return launch_gp_score_t<10>(a, grid, smem, s);
        case 16: return launch_gp_score_t<16>(a, grid, smem, s);
        case 20: return launch_gp_score_t<20>(a, grid, smem, s);
        case 32: return launch_gp_score_t<32>(a, grid, smem, s);
    }
    return IREC_E_INVALID;
}

static size_t fused_smem(int DP, int bmax, int cap)
{
    return sizeof(float) * ((size_t)IREC_T2_LEN + 8 * (size_t)DP + 2 * (size_t)bmax * DP + cap + GP2_THREADS + 32 + 1) +
           sizeof(int32_t) * ((size_t)cap + 32 + TOPK_CAP + 4 + 1 + 2 + 32 + 64) + sizeof(irec_record_t) * 32 +
           sizeof(uint16_t) * GP_DL4_LEN + 64;
}
template <int BMAX>
static int launch_fused_t(const FusedArgs& a, int grid, size_t smem, cudaStream_t s)
{
    static bool attr_done = false;
    if (!attr_done) {
        if (cudaFuncSetAttribute(k_gp_fused<BMAX>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)irec_device().max_smem_optin) != cudaSuccess)
            return IREC_E_CUDA;
        attr_done = true;
    }
    void* args[] = { const_cast<FusedArgs*>(&a) };
    if (cudaLaunchCooperativeKernel(reinterpret_cast<void*>(k_gp_fused<BMAX>), dim3(grid), dim3(GP2_THREADS), args, smem, s) != cudaSuccess)
        return IREC_E_CUDA;
    irec_count_launch();
    return IREC_OK;
}
extern "C" {

int irec_beam_uniform_int(int64_t q, int64_t start, int64_t n, int32_t* out, void* stream)
{
    IREC_ENSURE_INIT();
    if (n <= 0) return IREC_OK;
    k_beam_uniform_int<<<(int)std::min<int64_t>((n + 255) / 256, 1024), 256, 0, (cudaStream_t)stream>>>(q, start, n, out);
    irec_count_launch();
    return irec_check_launch("k_beam_uniform_int");
}

int irec_uniform_int_stream(int64_t global_seed, int64_t op_seed, int32_t lo, int32_t hi, int64_t start, int64_t n, int32_t* out,
                            void* stream)
{
    IREC_ENSURE_INIT();
    if (hi <= lo) return irec_fail(IREC_E_INVALID, "uniform_int_stream: need lo < hi");
    if (n <= 0) return IREC_OK;
    k_uniform_int_stream<<<(int)std::min<int64_t>((n + 255) / 256, 1024), 256, 0, (cudaStream_t)stream>>>(
        tf_stream_seeded(global_seed, op_seed), (uint32_t)lo, (uint32_t)(hi - lo), start, n, out);
    irec_count_launch();
    return irec_check_launch("k_uniform_int_stream");
}

int irec_kl_naux(const float* t_loc, const float* t_scale, const float* p_loc, const float* p_scale,
                 const int64_t* gather_idx, const int64_t* block_offsets, int nb, float omega,
                 float* out_kl, int32_t* out_n_aux, void* stream)
{
    IREC_ENSURE_INIT();
    if (nb <= 0) return IREC_OK;
    k_kl_naux<<<std::min(nb, 4 * irec_device().sm_count), 256, 0, (cudaStream_t)stream>>>(
        t_loc, t_scale, p_loc, p_scale, gather_idx, block_offsets, nb, omega, out_kl, out_n_aux);
    irec_count_launch();
    return irec_check_launch("k_kl_naux");
}

/* KL, n_aux and the per-partition per-dim schedule of one block of D contiguous dims (beam_search_coder.py:57-77;
 * coder.py:141-154): out_sa / out_A / out_E / out_M are [max_aux x D], rows t < n_aux are written (none when n_aux is
 * invalid: <= 0, > max_aux or beyond the ratio table -- check out_n_aux). */
int irec_schedule(const float* t_loc, const float* t_scale, const float* p_loc, const float* p_scale, int D, float omega,
                  int max_aux, float* out_kl, int32_t* out_n_aux, float* out_sa, float* out_A, float* out_E, float* out_M,
                  void* stream)
{
    IREC_ENSURE_INIT();
    cudaStream_t s = (cudaStream_t)stream;
    if (D <= 0 || max_aux <= 0 || !out_n_aux) return irec_fail(IREC_E_INVALID, "irec_schedule: bad arguments");
    if (((D + 31) >> 5) > KL_MAX_CHUNKS) return irec_fail(IREC_E_CAPACITY, "irec_schedule: D too large (max 131072 per block)");
    if (!(omega > 0.f)) return irec_fail(IREC_E_INVALID, "irec_schedule: kl_per_partition must be > 0");
    int64_t h_offs[2] = { 0, D };
    int64_t* d_offs = nullptr;      // two offsets: small pooled allocation, stream-ordered
    if (cudaMallocAsync(&d_offs, sizeof(h_offs), s) != cudaSuccess) return irec_fail(IREC_E_CUDA, "irec_schedule: allocation failed");
    cudaMemcpyAsync(d_offs, h_offs, sizeof(h_offs), cudaMemcpyHostToDevice, s);
    k_kl_naux<<<1, 256, 0, s>>>(t_loc, t_scale, p_loc, p_scale, nullptr, d_offs, 1, omega, out_kl, out_n_aux);
    irec_count_launch();
    k_schedule<<<std::max(1, std::min((D + 127) / 128, 1024)), 128, 0, s>>>(t_loc, t_scale, p_loc, p_scale, D, max_aux, out_n_aux,
                                                                             irec_ratio_tab(), irec_ratio_len(), out_sa, out_A, out_E, out_M);
    irec_count_launch();
    cudaFreeAsync(d_offs, s);
    return irec_check_launch("irec_schedule");
}

int irec_beam_decode(const float* p_loc, const float* p_scale, const int64_t* gather_idx,
                     const int64_t* block_offsets, int nb, int S, int64_t seed,
                     const int32_t* indices, int max_aux, const int32_t* n_aux,
                     float* out_sample, int32_t* out_status, void* stream)
{
    IREC_ENSURE_INIT();
    (void)S;
    if (nb <= 0) return IREC_OK;
    if (max_aux <= 0) return irec_fail(IREC_E_INVALID, "beam_decode: max_aux must be > 0");
    k_beam_decode<<<std::min(nb, 8 * irec_device().sm_count), 256, 0, (cudaStream_t)stream>>>(
        p_loc, p_scale, gather_idx, block_offsets, nb, seed, indices, max_aux, n_aux, irec_device().d_T,
        irec_ratio_tab(), irec_ratio_len(), out_sample, out_status);
    irec_count_launch();
    return irec_check_launch("k_beam_decode");
}

// ---------------- opaque per-block state API (general path; multi-GPU candidate sharding) ------
size_t irec_beam_state_bytes(int D, int B, int max_aux) { return state_bytes(D, B, max_aux); }

int irec_beam_state_init(void* state, const float* t_loc, const float* t_scale, const float* p_loc,
                         const float* p_scale, const int64_t* gather_idx, int64_t offset, int D, float omega, int S,
                         int B, int max_aux, int64_t seed, void* stream)
{
    IREC_ENSURE_INIT();
    if (D <= 0 || B <= 0 || B > 32 || S <= 0 || max_aux <= 0) return irec_fail(IREC_E_INVALID, "beam_state_init: bad sizes");
    if (((D + 31) >> 5) > KL_MAX_CHUNKS) return irec_fail(IREC_E_CAPACITY, "beam_state_init: D too large (max 131072 per block)");
    if ((int64_t)S * B >= (1LL << 31)) return irec_fail(IREC_E_CAPACITY, "beam_state_init: S*B must be < 2^31");
    k_gp_init<<<1, 256, 0, (cudaStream_t)stream>>>(state, t_loc, t_scale, p_loc, p_scale, gather_idx, offset, D, B, S,
                                                   max_aux, irec_ratio_len(), omega, seed);
    irec_count_launch();
    return irec_check_launch("k_gp_init");
}

/* reads n_aux/status/kl back (synchronises the stream) */
int irec_beam_state_query(const void* state, int32_t* n_aux, int32_t* status, float* kl, void* stream)
{
    BeamStateHdr h;
    if (cudaMemcpyAsync(&h, state, sizeof(h), cudaMemcpyDeviceToHost, (cudaStream_t)stream) != cudaSuccess ||
        cudaStreamSynchronize((cudaStream_t)stream) != cudaSuccess)
        return irec_fail(IREC_E_CUDA, "beam_state_query: copy failed");
    if (n_aux) *n_aux = h.n_aux;
    if (status) *status = h.status;
    if (kl) *kl = h.kl;
    return IREC_OK;
}

size_t irec_beam_step_workspace_bytes(int D, int B)
{
    if (irec_init() != IREC_OK) return 0;     // the size depends on the device's SM count
    const int grid_max = irec_device().sm_count * 4;
    return sizeof(irec_record_t) * ((size_t)grid_max * 32 + 32) + sizeof(int32_t) * ((size_t)grid_max + 8) +
           (sizeof(float) + sizeof(int32_t)) * ((size_t)grid_max * 32 + 64) + 256;
}

/* schedule of partition t + local top-B over candidates [s_begin, s_end):
 * out_records [B] (sorted best first), out_count [1] */
int irec_beam_step_score(void* state, int D, int B, int t, int64_t s_begin, int64_t s_end, int do_params,
                         irec_record_t* out_records, int32_t* out_count, void* workspace, size_t workspace_bytes,
                         void* stream)
{
    IREC_ENSURE_INIT();
    cudaStream_t s = (cudaStream_t)stream;
    const int bmax = pick_bmax(B);
    if (bmax < 0) return irec_fail(IREC_E_INVALID, "beam_step_score: n_beams must be <= 32");
    if (workspace_bytes < irec_beam_step_workspace_bytes(D, B)) return irec_fail(IREC_E_CAPACITY, "beam_step_score: workspace too small");
    const BeamGeom g = make_geom(D);
    if (do_params) {
        k_gp_params<<<std::max(1, std::min((g.DP + 255) / 256, 64)), 256, 0, s>>>(state, t, irec_ratio_tab());
        irec_count_launch();
    }
    const int grid_max = irec_device().sm_count * 4;
    unsigned char* w = reinterpret_cast<unsigned char*>(workspace);
    irec_record_t* rec = reinterpret_cast<irec_record_t*>(w);
    int32_t* cnt = reinterpret_cast<int32_t*>(rec + (size_t)grid_max * 32 + 32);
    float* g_sc = reinterpret_cast<float*>(cnt + grid_max + 8);
    int32_t* g_id = reinterpret_cast<int32_t*>(g_sc + (size_t)grid_max * 32 + 64);

    const int64_t ns = s_end > s_begin ? s_end - s_begin : 0;
    int grid, rc;
    if (gp_use_v2(D, bmax)) {
        grid = gp2_score_grid(D, ns);
        const int cap = gp2_cand_cap(D, bmax);
        Score2Args a;
        a.state = state; a.t = t; a.T2 = irec_device().d_T2; a.dl4 = irec_device().d_dl4; a.s_begin = s_begin; a.s_end = s_end;
        a.out_rec = rec; a.out_cnt = cnt; a.cand_cap = cap;
        rc = launch_gp2_score(bmax, a, grid, gp2_score_smem(cap), s);
    } else {
        grid = gp_score_grid(D, ns);
        const int cap = gp_cand_cap(D, bmax);
        ScoreArgs a;
        a.state = state; a.t = t; a.T = irec_device().d_T; a.s_begin = s_begin; a.s_end = s_end;
        a.out_rec = rec; a.out_cnt = cnt; a.cand_cap = cap;
        rc = launch_gp_score(bmax, a, grid, gp_score_smem(cap), s);
    }
    if (rc != IREC_OK) return irec_fail(rc, "beam_step_score: launch failed");
    // Bcur is only known on the device: the merge kernel reads ids as s*Bcur+b with the header's Bcur
    k_topb_merge<<<1, 256, 0, s>>>(rec, cnt, grid, B, &reinterpret_cast<BeamStateHdr*>(state)->Bcur, 0, B, out_records,
                                   out_count, g_sc, g_id);
    irec_count_launch();
    return irec_check_launch("beam_step_score");
}

/* merge n_lists x B gathered records (counts[i] valid in list i; counts may be NULL = all B valid)
 * and commit the winners: new beams, hash sums, history; advances the state to partition t+1 */
int irec_beam_step_commit(void* state, int D, int B, int t, const irec_record_t* records, const int32_t* counts,
                          int n_lists, void* workspace, size_t workspace_bytes, void* stream)
{
    IREC_ENSURE_INIT();
    cudaStream_t s = (cudaStream_t)stream;
    if (workspace_bytes < irec_beam_step_workspace_bytes(D, B)) return irec_fail(IREC_E_CAPACITY, "beam_step_commit: workspace too small");
    if (n_lists > irec_device().sm_count * 4) return irec_fail(IREC_E_CAPACITY, "beam_step_commit: too many lists");
    const int grid_max = irec_device().sm_count * 4;
    unsigned char* w = reinterpret_cast<unsigned char*>(workspace);
    irec_record_t* rec = reinterpret_cast<irec_record_t*>(w);
    int32_t* cnt = reinterpret_cast<int32_t*>(rec + (size_t)grid_max * 32 + 32);
    float* g_sc = reinterpret_cast<float*>(cnt + grid_max + 8);
    int32_t* g_id = reinterpret_cast<int32_t*>(g_sc + (size_t)grid_max * 32 + 64);
    irec_record_t* win = rec + (size_t)grid_max * 32;       // last 32 records of the workspace
    int32_t* nwin = cnt + grid_max;
    k_topb_merge<<<1, 256, 0, s>>>(records, counts, n_lists, B, &reinterpret_cast<BeamStateHdr*>(state)->Bcur, 0, B, win, nwin,
                                   g_sc, g_id);
    irec_count_launch();
    const BeamGeom g = make_geom(D);
    const int tasks = (g.DP >> 2) * B;
    k_gp_commit<<<std::max(1, std::min((tasks + 255) / 256, irec_device().sm_count * 2)), 256, 0, s>>>(
        state, t, win, nwin, irec_device().d_T);
    irec_count_launch();
    k_gp_flip<<<1, 1, 0, s>>>(state, t, nwin);
    irec_count_launch();
    return irec_check_launch("beam_step_commit");
}

int irec_beam_state_finish(void* state, int D, const int64_t* gather_idx, int64_t offset, int32_t* out_indices,
                           int32_t* out_n_aux, int32_t* out_status, float* out_sample, void* stream)
{
    IREC_ENSURE_INIT();
    k_gp_finish<<<std::max(1, std::min((D + 255) / 256, 64)), 256, 0, (cudaStream_t)stream>>>(
        state, gather_idx, offset, out_indices, out_n_aux, out_status, out_sample);
    irec_count_launch();
    return irec_check_launch("k_gp_finish");
}

int irec_topb_merge(const irec_record_t* records, int n_records, int Bcur, int B, irec_record_t* out_records,
                    int32_t* out_count, void* workspace, size_t workspace_bytes, void* stream)
{
    IREC_ENSURE_INIT();
    if (n_records <= 0 || B <= 0 || B > 32) return irec_fail(IREC_E_INVALID, "topb_merge: bad sizes");
    if (workspace_bytes < (sizeof(float) + sizeof(int32_t)) * (size_t)n_records)
        return irec_fail(IREC_E_CAPACITY, "topb_merge: workspace too small");
    float* g_sc = reinterpret_cast<float*>(workspace);
    int32_t* g_id = reinterpret_cast<int32_t*>(g_sc + n_records);
    k_topb_merge<<<1, 256, 0, (cudaStream_t)stream>>>(records, nullptr, 1, n_records, nullptr, Bcur, B, out_records, out_count,
                                                      g_sc, g_id);
    irec_count_launch();
    return irec_check_launch("k_topb_merge");
}

// ---------------- record exchange over NVLink peer memory --------------------------------------------
// The candidate-range sharded coder exchanges (B + 1) 16-byte records per rank and auxiliary variable: 336 bytes.  At
// that size an NCCL all-gather is pure latency (launch + protocol, ~20-40 us on 8 GPUs) and sits between two small
// kernels.  Here every rank STORES its records straight into every peer's exchange buffer (peer-mapped device memory,
// NVLink / NVSwitch) and raises a per-sender sequence flag with a system-scope release; it then spins on its own flags
// (system-scope acquire) until all peers' records of this step have landed, and hands them to the merge.
// Exchange buffer of one rank (int32 words; allocated symmetric on all ranks, zero-initialised):
//   [0]               steps completed by the owner (device-side sequence counter: CUDA-graph replays stay in step)
//   [64 .. 64+world)  flag[sender] = last sequence number `sender` has pushed here
//   [128 ..)          slot[parity][sender][(B + 1) * 4]   (parity = sequence & 1: a fast sender's next step never lands
//                     in the slot its peers are still reading)
__global__ void __launch_bounds__(256) k_p2p_exchange(int32_t* const* __restrict__ peer_bufs, int rank, int world, int B,
                                                       const int32_t* __restrict__ local, int32_t* __restrict__ out_records,
                                                       int32_t* __restrict__ out_counts)
{
    __shared__ int s_seq;
    int32_t* mine = peer_bufs[rank];
    if (threadIdx.x == 0) s_seq = mine[0] + 1;
    __syncthreads();
    const int seq = s_seq, W = (B + 1) * 4, parity = seq & 1;
    // push: my records into slot[parity][rank] of every rank (my own included)
    for (int i = threadIdx.x; i < world * W; i += blockDim.x) {
        const int r = i / W, w = i - r * W;
        peer_bufs[r][P2P_SLOT_OFF + (parity * world + rank) * W + w] = local[w];
    }
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x < world) {
        int32_t* flag = peer_bufs[threadIdx.x] + P2P_FLAG_OFF + rank;
        asm volatile("st.release.sys.global.s32 [%0], %1;" ::"l"(flag), "r"(seq) : "memory");
        // wait: the records of every sender for this step
        const int32_t* want = mine + P2P_FLAG_OFF + threadIdx.x;
        int got;
        do {
            asm volatile("ld.acquire.sys.global.s32 %0, [%1];" : "=r"(got) : "l"(want) : "memory");
        } while (got - seq < 0);
    }
    __syncthreads();
    const volatile int32_t* slots = mine + P2P_SLOT_OFF + parity * world * W;
    for (int i = threadIdx.x; i < world * B * 4; i += blockDim.x) {
        const int r = i / (B * 4), w = i - r * (B * 4);
        out_records[i] = slots[r * W + w];
    }
    for (int r = threadIdx.x; r < world; r += blockDim.x) out_counts[r] = slots[r * W + B * 4];
    __syncthreads();
    if (threadIdx.x == 0) mine[0] = seq;
}

size_t irec_p2p_exchange_bytes(int B, int world) { return sizeof(int32_t) * (P2P_SLOT_OFF + 2 * (size_t)world * (size_t)(B + 1) * 4); }

int irec_p2p_exchange(void* const* peer_bufs, int rank, int world, int B, const irec_record_t* local_records_and_count,
                      irec_record_t* out_records, int32_t* out_counts, void* stream)
{
    IREC_ENSURE_INIT();
    if (world < 1 || world > 64 || rank < 0 || rank >= world || B <= 0 || B > 32)
        return irec_fail(IREC_E_INVALID, "p2p_exchange: bad rank / world / B");
    k_p2p_exchange<<<1, 256, 0, (cudaStream_t)stream>>>(reinterpret_cast<int32_t* const*>(peer_bufs), rank, world, B,
                                                        reinterpret_cast<const int32_t*>(local_records_and_count),
                                                        reinterpret_cast<int32_t*>(out_records), out_counts);
    irec_count_launch();
    return irec_check_launch("k_p2p_exchange");
}

// ---------------- fused per-block launch (all auxiliary variables, grid + ranks) ------------------------
int irec_beam_fused_fits(int D, int B)
{
    if (irec_init() != IREC_OK) return 0;
    const int bmax = pick_bmax(B);
    if (bmax < 0 || D <= 0 || D > 1024) return 0;
    return fused_smem(make_geom(D).DP, bmax, gp2_cand_cap(D, bmax)) <= (size_t)irec_device().max_smem_optin ? 1 : 0;
}
size_t irec_beam_fused_workspace_bytes(int B, int world)
{
    if (irec_init() != IREC_OK) return 0;
    const size_t grid = (size_t)irec_device().sm_count;
    return 256 + sizeof(irec_record_t) * (grid * 32 + 64) + sizeof(int32_t) * (grid + 8) +
           (sizeof(float) + sizeof(int32_t)) * ((grid + (size_t)std::max(world, 1)) * 32 + 64) + 64 + 0 * (size_t)B;
}
/* All auxiliary variables of the coder-block in `state` (after irec_beam_state_init) in one cooperative launch: this rank
 * scores candidates [s_begin, s_end), the ranks' top-B lists are exchanged through peer_bufs (as irec_p2p_exchange; NULL
 * and world = 1 on a single GPU).  Then irec_beam_state_finish.  IREC_E_CAPACITY: the block does not fit the fused kernel
 * (D > 256 at 20 beams) -- use irec_beam_step_score / irec_beam_step_commit. */
int irec_beam_encode_fused(void* state, int D, int B, int64_t s_begin, int64_t s_end, void* const* peer_bufs, int rank, int world,
                           void* workspace, size_t workspace_bytes, void* stream)
{
    IREC_ENSURE_INIT();
    cudaStream_t s = (cudaStream_t)stream;
    const int bmax = pick_bmax(B);
    if (bmax < 0 || D <= 0) return irec_fail(IREC_E_INVALID, "beam_encode_fused: bad sizes");
    if (world < 1 || world > 64 || rank < 0 || rank >= world || (world > 1 && !peer_bufs))
        return irec_fail(IREC_E_INVALID, "beam_encode_fused: bad rank / world / peer buffers");
    if (workspace_bytes < irec_beam_fused_workspace_bytes(B, world)) return irec_fail(IREC_E_CAPACITY, "beam_encode_fused: workspace too small");
    const BeamGeom g = make_geom(D);
    if (D > 1024) return irec_fail(IREC_E_CAPACITY, "beam_encode_fused: D > 1024");
    const int cap = gp2_cand_cap(D, bmax);
    const size_t smem = fused_smem(g.DP, bmax, cap);
    if (smem > (size_t)irec_device().max_smem_optin) return irec_fail(IREC_E_CAPACITY, "beam_encode_fused: block state does not fit shared memory");
    const int64_t ns = s_end > s_begin ? s_end - s_begin : 0;
    const int grid = gp2_score_grid(D, ns);
    const size_t gsz = (size_t)irec_device().sm_count;
    unsigned char* w = reinterpret_cast<unsigned char*>(workspace);
    FusedArgs a;
    a.state = state; a.T2 = irec_device().d_T2; a.dl4 = irec_device().d_dl4; a.ratio_tab = irec_ratio_tab();
    a.s_begin = s_begin; a.s_end = s_end;
    a.sync = reinterpret_cast<int32_t*>(w);
    a.lists = reinterpret_cast<irec_record_t*>(w + 256);
    a.win = a.lists + gsz * 32;
    a.list_cnt = reinterpret_cast<int32_t*>(a.win + 64);
    a.nwin = a.list_cnt + gsz;
    a.g_sc = reinterpret_cast<float*>(a.nwin + 8);
    a.g_id = reinterpret_cast<int32_t*>(a.g_sc + (gsz + (size_t)world) * 32 + 64);
    a.peer_bufs = reinterpret_cast<int32_t* const*>(peer_bufs); a.rank = rank; a.world = world; a.cand_cap = cap;
    if (cudaMemsetAsync(a.sync, 0, 256, s) != cudaSuccess) return irec_fail(IREC_E_CUDA, "beam_encode_fused: memset failed");
    a.prof = nullptr;
    static long long* d_prof = nullptr;            // diagnostics only (IREC_GP_PROFILE=1): phase cycle counters, dumped to stderr
    const char* pe = getenv("IREC_GP_PROFILE");
    const bool prof_on = pe && pe[0] == '1';
    if (prof_on) {
        if (!d_prof) cudaMalloc(&d_prof, sizeof(long long) * 8 * 1024);
        cudaMemsetAsync(d_prof, 0, sizeof(long long) * 8 * 1024, s);
        a.prof = d_prof;
    }
    int rc = IREC_E_INVALID;
    switch (bmax) {
        case 1: rc = launch_fused_t<1>(a, grid, smem, s); break;
        case 2: rc = launch_fused_t<2>(a, grid, smem, s); break;
        case 4: rc = launch_fused_t<4>(a, grid, smem, s); break;
        case 8: rc = launch_fused_t<8>(a, grid, smem, s); break;
        case 10: rc = launch_fused_t<10>(a, grid, smem, s); break;
        case 16: rc = launch_fused_t<16>(a, grid, smem, s); break;
        case 20: rc = launch_fused_t<20>(a, grid, smem, s); break;
        case 32: rc = launch_fused_t<32>(a, grid, smem, s); break;
    }
    if (rc != IREC_OK) { cudaGetLastError(); return irec_fail(rc, "beam_encode_fused: cooperative launch failed"); }
    if (prof_on) {
        std::vector<long long> h((size_t)8 * grid);
        cudaStreamSynchronize(s);
        cudaMemcpy(h.data(), d_prof, sizeof(long long) * h.size(), cudaMemcpyDeviceToHost);
        double tot[8] = { 0 };
        for (int i = 0; i < grid; ++i)
            for (int k = 0; k < 8; ++k) tot[k] += (double)h[(size_t)i * 8 + k];
        fprintf(stderr, "[gp profile] rank %d grid %d, cycles per CTA over the launch: schedule %.0f scoring %.0f own-topB+arrive %.0f "
                        "waiting %.0f commit %.0f | last arriver (sum over variables): merge+exchange+publish %.0f\n", rank, grid,
                tot[0] / grid, tot[1] / grid, tot[2] / grid, tot[4] / grid, tot[5] / grid, tot[3]);
    }
    return irec_check_launch("k_gp_fused");
}

// ---------------- the block-batched entry point ------------------------------------------------
size_t irec_beam_encode_workspace_bytes(int nb, int64_t max_block_dim, int S, int B, int max_aux)
{
    if (irec_init() != IREC_OK) return 0;
    if (B > 32) return max_block_dim <= 1024 ? irec_wide_workspace_bytes(nb, (int)max_block_dim, S, B, max_aux) : 0;
    const int bmax = B <= 1 ? 1 : (B <= 10 ? 10 : (B <= 20 ? 20 : 32));     // the largest beam capacity any persistent kernel picks for B
    const size_t resident = r2_ws_hist_bytes(bmax, max_aux) + r2_ws_sched_bytes() + r2_ws_order_bytes(nb) +
                            r2_table_bytes((int)max_block_dim, S, max_aux);
    const size_t general = (state_bytes((int)max_block_dim, B, max_aux) + 255) / 256 * 256 + sizeof(irec_record_t) * 32 + 256 +
                           irec_beam_step_workspace_bytes((int)max_block_dim, B);
    const size_t cluster = 512 + irec_cluster_hist_bytes(nb, max_aux) + r2_ws_order_bytes(nb) + r2_table_bytes((int)max_block_dim, S, max_aux);
    return std::max(std::max(resident, general), cluster);
}

int irec_beam_encode_path(int nb, int64_t max_block_dim, int S, int B)
{
    IREC_ENSURE_INIT();
    if (nb <= 0 || S <= 0 || B <= 0 || max_block_dim <= 0) return 0;
    if (B > 32) return (max_block_dim <= 1024 && irec_wide_supported(nb, (int)max_block_dim, S, B)) ? 4 : 0;
    const int rchoice = resident_choice();
    if (irec_force_general()) return 0;
    if (rchoice == 0) {
        const int G = irec_cluster_choice(nb, (int)max_block_dim, S, B);
        if (G > 0) return 100 + G;
    }
    if ((rchoice == 0 || rchoice == 3) && irec_tmem_plan(nb, (int)max_block_dim, S, B, nullptr)) return 3;
    if (rchoice != 1 && rchoice != 3 && plan_resident2(nb, (int)max_block_dim, S, B).ok) return 2;
    if (rchoice != 2 && rchoice != 3 && plan_resident(nb, (int)max_block_dim, S, B).ok) return 1;
    return 0;
}

int irec_beam_encode(const float* t_loc, const float* t_scale, const float* p_loc, const float* p_scale,
                     const int64_t* gather_idx, const int64_t* block_offsets, int nb, int64_t max_block_dim,
                     float omega, int S, int B, int64_t seed,
                     int32_t* out_indices, int max_aux, int32_t* out_n_aux, int32_t* out_status,
                     float* out_sample, void* workspace, size_t workspace_bytes, void* stream)
{
    IREC_ENSURE_INIT();
    cudaStream_t s = (cudaStream_t)stream;
    if (nb <= 0) return IREC_OK;
    if (S <= 0 || B <= 0 || max_aux <= 0 || max_block_dim <= 0) return irec_fail(IREC_E_INVALID, "beam_encode: bad sizes");
    if ((int64_t)S * B >= (1LL << 31)) return irec_fail(IREC_E_CAPACITY, "beam_encode: S*B must be < 2^31");
    if (!(omega > 0.f)) return irec_fail(IREC_E_INVALID, "beam_encode: kl_per_partition must be > 0");
    if (B > 32) {
        // wide beams (irec_wide.cu): the reference has no limit on n_beams (beam_search_coder.py:15-30)
        if (max_block_dim > 1024 || !irec_wide_supported(nb, (int)max_block_dim, S, B))
            return irec_fail(IREC_E_CAPACITY, "beam_encode: n_beams > 32 needs n_beams <= 1024, block dims <= 1024 and S * n_beams < 2^31");
        if (workspace_bytes < irec_wide_workspace_bytes(nb, (int)max_block_dim, S, B, max_aux))
            return irec_fail(IREC_E_CAPACITY, "beam_encode: workspace too small");
        return irec_launch_wide(t_loc, t_scale, p_loc, p_scale, gather_idx, block_offsets, nb, (int)max_block_dim, omega, S, B, seed,
                                out_indices, max_aux, out_n_aux, out_status, out_sample, workspace, s);
    }
    if (workspace_bytes < irec_beam_encode_workspace_bytes(nb, max_block_dim, S, B, max_aux))
        return irec_fail(IREC_E_CAPACITY, "beam_encode: workspace too small");

    const int rchoice = resident_choice();
    const int cluster_G = (irec_force_general() || rchoice != 0) ? 0 : irec_cluster_choice(nb, (int)max_block_dim, S, B);
    if (cluster_G > 0) {
        // few coder-blocks: one thread-block cluster per block (irec_cluster.cu)
        unsigned char* w = reinterpret_cast<unsigned char*>(workspace);
        R2Plan* dplan = reinterpret_cast<R2Plan*>(w + 256);
        int2* hist = reinterpret_cast<int2*>(w + 512);
        int32_t* order = reinterpret_cast<int32_t*>(w + 512 + irec_cluster_hist_bytes(nb, max_aux));
        if (cudaMemsetAsync(dplan, 0, sizeof(R2Plan), s) != cudaSuccess) return irec_fail(IREC_E_CUDA, "beam_encode: memset failed");
        k_r2_plan<<<1, 1024, 0, s>>>(block_offsets, nb, dplan, order);
        irec_count_launch();
        const size_t tab_bytes = r2_table_bytes((int)max_block_dim, S, max_aux);
        const uint2* tab = nullptr;
        if (tab_bytes && !r2_no_table()) {
            uint2* t = reinterpret_cast<uint2*>(w + 512 + irec_cluster_hist_bytes(nb, max_aux) + r2_ws_order_bytes(nb));
            const int row_stride = make_geom((int)max_block_dim).DP >> 2;
            const R2TabUse u = r2_tab_acquire(t, seed, S, max_aux, row_stride, s);
            r2_build_table(u, dplan, seed, S, max_aux, row_stride, s);
            const int rc = irec_launch_cluster(cluster_G, t_loc, t_scale, p_loc, p_scale, gather_idx, block_offsets, nb, (int)max_block_dim,
                                               omega, S, B, seed, out_indices, max_aux, out_n_aux, out_status, out_sample, hist, order,
                                               dplan, u.tab, u.tab_aux, u.priv, s);
            r2_tab_release(u, s);
            return rc;
        }
        return irec_launch_cluster(cluster_G, t_loc, t_scale, p_loc, p_scale, gather_idx, block_offsets, nb, (int)max_block_dim, omega,
                                   S, B, seed, out_indices, max_aux, out_n_aux, out_status, out_sample, hist, order,
                                   nullptr, tab, max_aux, nullptr, s);
    }
    if (!irec_force_general() && (rchoice == 0 || rchoice == 3)) {
        TmemPlan tp;
        if (irec_tmem_plan(nb, (int)max_block_dim, S, B, &tp)) {
            // tensor-memory kernel (irec_tmem.cu); same workspace layout as resident2
            int* counter = reinterpret_cast<int*>(workspace);
            if (cudaMemsetAsync(counter, 0, 256, s) != cudaSuccess) return irec_fail(IREC_E_CUDA, "beam_encode: memset failed");
            unsigned char* w = reinterpret_cast<unsigned char*>(workspace);
            int2* hist = reinterpret_cast<int2*>(w + 512);
            float* sched = reinterpret_cast<float*>(w + r2_ws_hist_bytes(tp.bmax, max_aux));
            R2Plan* dplan = reinterpret_cast<R2Plan*>(w + 256);
            int32_t* order = reinterpret_cast<int32_t*>(w + r2_ws_hist_bytes(tp.bmax, max_aux) + r2_ws_sched_bytes());
            if (cudaMemsetAsync(dplan, 0, sizeof(R2Plan), s) != cudaSuccess) return irec_fail(IREC_E_CUDA, "beam_encode: memset failed");
            k_r2_plan<<<1, 1024, 0, s>>>(block_offsets, nb, dplan, order);
            irec_count_launch();
            const size_t tab_bytes = r2_table_bytes((int)max_block_dim, S, max_aux);
            if (tab_bytes && !r2_no_table()) {
                uint2* tab = reinterpret_cast<uint2*>(w + r2_ws_hist_bytes(tp.bmax, max_aux) + r2_ws_sched_bytes() + r2_ws_order_bytes(nb));
                const R2TabUse u = r2_tab_acquire(tab, seed, S, max_aux, tp.DPmax >> 2, s);
                r2_build_table(u, dplan, seed, S, max_aux, tp.DPmax >> 2, s);
                const int rc = irec_launch_tmem(tp, t_loc, t_scale, p_loc, p_scale, gather_idx, block_offsets, nb, omega, S, B, seed,
                                                out_indices, max_aux, out_n_aux, out_status, out_sample, hist, sched, counter, order,
                                                dplan, u.tab, u.tab_aux, u.priv, s);
                r2_tab_release(u, s);
                return rc;
            }
            return irec_launch_tmem(tp, t_loc, t_scale, p_loc, p_scale, gather_idx, block_offsets, nb, omega, S, B, seed, out_indices,
                                    max_aux, out_n_aux, out_status, out_sample, hist, sched, counter, order, nullptr, nullptr, max_aux, nullptr, s);
        }
        if (rchoice == 3) return irec_fail(IREC_E_CAPACITY, "beam_encode: IREC_RESIDENT=3 but the sizes do not fit the tensor-memory kernel");
    }
    if (!irec_force_general() && rchoice != 1 && rchoice != 3) {
        const ResidentPlan plan2 = plan_resident2(nb, (int)max_block_dim, S, B);
        if (plan2.ok) {
            int* counter = reinterpret_cast<int*>(workspace);
            if (cudaMemsetAsync(counter, 0, 256, s) != cudaSuccess) return irec_fail(IREC_E_CUDA, "beam_encode: memset failed");
            Resident2Args a;
            a.t_loc = t_loc; a.t_scale = t_scale; a.p_loc = p_loc; a.p_scale = p_scale;
            a.gidx = gather_idx; a.offs = block_offsets; a.nb = nb; a.omega = omega; a.S = S; a.B = B; a.seed = seed;
            a.out_indices = out_indices; a.max_aux = max_aux; a.out_n_aux = out_n_aux; a.out_status = out_status;
            a.out_sample = out_sample; a.T2 = irec_device().d_T2; a.dl4 = irec_device().d_dl4;
            a.ratio_tab = irec_ratio_tab(); a.ratio_len = irec_ratio_len();
            unsigned char* w = reinterpret_cast<unsigned char*>(workspace);
            a.hist = reinterpret_cast<int2*>(w + 512);
            a.sched = reinterpret_cast<float*>(w + r2_ws_hist_bytes(plan2.bmax, max_aux));
            a.work_counter = counter; a.DPmax = plan2.DPmax; a.NC = plan2.NC;
            a.plan = nullptr; a.tab = nullptr; a.tab_aux = max_aux; a.tab_priv = nullptr;
            // distinct block sizes + the queue order (largest blocks first)
            R2Plan* dplan = reinterpret_cast<R2Plan*>(w + 256);
            int32_t* order = reinterpret_cast<int32_t*>(w + r2_ws_hist_bytes(plan2.bmax, max_aux) + r2_ws_sched_bytes());
            if (cudaMemsetAsync(dplan, 0, sizeof(R2Plan), s) != cudaSuccess) return irec_fail(IREC_E_CUDA, "beam_encode: memset failed");
            k_r2_plan<<<1, 1024, 0, s>>>(block_offsets, nb, dplan, order);
            irec_count_launch();
            a.order = order;
            const size_t tab_bytes = r2_table_bytes((int)max_block_dim, S, max_aux);
            if (tab_bytes && !r2_no_table()) {
                // exponent table of this launch (the candidate stream is the same for every coder-block)
                uint2* tab = reinterpret_cast<uint2*>(w + r2_ws_hist_bytes(plan2.bmax, max_aux) + r2_ws_sched_bytes() + r2_ws_order_bytes(nb));
                const R2TabUse u = r2_tab_acquire(tab, seed, S, max_aux, plan2.DPmax >> 2, s);
                r2_build_table(u, dplan, seed, S, max_aux, plan2.DPmax >> 2, s);
                a.plan = dplan; a.tab = u.tab; a.tab_aux = u.tab_aux; a.tab_priv = u.priv;
                launch_resident2(plan2, a, s);
                r2_tab_release(u, s);
                return irec_check_launch("k_beam_encode_resident2");
            }
            launch_resident2(plan2, a, s);
            return irec_check_launch("k_beam_encode_resident2");
        }
        if (rchoice == 2) return irec_fail(IREC_E_CAPACITY, "beam_encode: IREC_RESIDENT=2 but the sizes do not fit the resident2 kernel");
    }
    const ResidentPlan plan = (irec_force_general() || rchoice == 3) ? ResidentPlan{} : plan_resident(nb, (int)max_block_dim, S, B);
    if (plan.ok) {
        int* counter = reinterpret_cast<int*>(workspace);
        if (cudaMemsetAsync(counter, 0, 256, s) != cudaSuccess) return irec_fail(IREC_E_CUDA, "beam_encode: memset failed");
        ResidentArgs a;
        a.t_loc = t_loc; a.t_scale = t_scale; a.p_loc = p_loc; a.p_scale = p_scale;
        a.gidx = gather_idx; a.offs = block_offsets; a.nb = nb; a.omega = omega; a.S = S; a.B = B; a.seed = seed;
        a.out_indices = out_indices; a.max_aux = max_aux; a.out_n_aux = out_n_aux; a.out_status = out_status;
        a.out_sample = out_sample; a.T = irec_device().d_T; a.ratio_tab = irec_ratio_tab();
        a.ratio_len = irec_ratio_len();
        a.hist = reinterpret_cast<int2*>(reinterpret_cast<unsigned char*>(workspace) + 256);
        a.work_counter = counter; a.DPmax = plan.DPmax; a.NC = plan.NC;
        launch_resident(plan, a, s);
        return irec_check_launch("k_beam_encode_resident");
    }

    // general path: block by block, partition by partition (host needs n_aux: one sync per block)
    std::vector<int64_t> offs(nb + 1);
    if (cudaMemcpyAsync(offs.data(), block_offsets, sizeof(int64_t) * (nb + 1), cudaMemcpyDeviceToHost, s) != cudaSuccess ||
        cudaStreamSynchronize(s) != cudaSuccess)
        return irec_fail(IREC_E_CUDA, "beam_encode: reading block offsets failed");
    unsigned char* w = reinterpret_cast<unsigned char*>(workspace);
    void* state = w;
    const size_t sb = (state_bytes((int)max_block_dim, B, max_aux) + 255) / 256 * 256;
    irec_record_t* recs = reinterpret_cast<irec_record_t*>(w + sb);
    int32_t* rcnt = reinterpret_cast<int32_t*>(recs + 32);
    void* step_ws = w + sb + sizeof(irec_record_t) * 32 + 256;
    const size_t step_ws_bytes = irec_beam_step_workspace_bytes((int)max_block_dim, B);
    for (int blk = 0; blk < nb; ++blk) {
        const int D = (int)(offs[blk + 1] - offs[blk]);
        if (D <= 0 || D > max_block_dim) return irec_fail(IREC_E_INVALID, "beam_encode: block size out of range");
        int rc = irec_beam_state_init(state, t_loc, t_scale, p_loc, p_scale, gather_idx, offs[blk], D, omega, S, B, max_aux, seed, s);
        if (rc != IREC_OK) return rc;
        int32_t n_aux = 0, status = 0;
        rc = irec_beam_state_query(state, &n_aux, &status, nullptr, s);
        if (rc != IREC_OK) return rc;
        if (status == IREC_BLK_OK) {
            for (int t = 0; t < n_aux; ++t) {
                rc = irec_beam_step_score(state, D, B, t, 0, S, 1, recs, rcnt, step_ws, step_ws_bytes, s);
                if (rc != IREC_OK) return rc;
                rc = irec_beam_step_commit(state, D, B, t, recs, rcnt, 1, step_ws, step_ws_bytes, s);
                if (rc != IREC_OK) return rc;
            }
        }
        rc = irec_beam_state_finish(state, D, gather_idx, offs[blk], out_indices + (size_t)blk * max_aux, out_n_aux + blk,
                                    out_status + blk, out_sample, s);
        if (rc != IREC_OK) return rc;
    }
    return IREC_OK;
}

}  // extern "C"
