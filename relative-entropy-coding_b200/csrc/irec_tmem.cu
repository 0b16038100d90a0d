// irec_tmem.cu -- K1a third generation: k_beam_encode_tmem, the persistent beam-search encoder with the beams and the
// per-dim coefficients in TENSOR MEMORY and two coder-blocks in flight per SM.
//
// Reference loop being replaced: rec/coding/beam_search_coder.py:53-122 (one context = one coder-block, all of its auxiliary
// variables); results are bit-identical to k_beam_encode_resident2, k_beam_encode_resident and the oracle.
//
// Why (profiles/r1_resident2_e_bench_launch_ncu.md, profiles/r2_tmem_probe.log): resident2 is bound by the shared-memory
// data pipe -- 2.15 wavefronts per quantile gather plus one LDS.128 per beam quad and four per coefficient quad (18 % of
// the wavefronts) -- and leaves the SM idle in the serial phases between scoring passes (top-B, re-materialisation, next
// schedule: 22 % of the time); a second coder-block per SM would hide those, but two 80 KB beam matrices do not fit beside
// the 120 KB quantile table.  Both problems have one cause: data that is LANE-PRIVATE and indexed WARP-UNIFORMLY (lane l
// owns the 32-dim chunk l; every lane of a warp reads "beam b, quad iq" of its own chunk) sits in shared memory.  That
// access pattern is exactly what tensor memory offers through tcgen05.ld/st.32x32b: thread i of warp w addresses TMEM lane
// 32 (w % 4) + i, columns are addressed warp-uniformly, 512 columns x 128 lanes x 4 B = 256 KB per SM that this path never
// used -- and tcgen05.ld runs beside a saturated shared-memory pipe at no cost (measured: 1653 cycles per 480 conflicted
// gathers with or without 14 tcgen05.ld.x4 per 40 gathers).
//
// Layout.  16 warps, two coder-block CONTEXTS per CTA, warp-specialised:
//   * warp w sits on scheduler / TMEM lane quarter q = w & 3 (hardware: a warp only reaches TMEM lanes 32 (w % 4) .. +31);
//   * warps 0..11 (three per quarter) are SCORING warps: they score the candidates of context 0, then of context 1, then of
//     context 0 again, ... with the whole shared-memory pipe;
//   * warps 12..15 (one per quarter) are SERIAL warps: while the scoring warps work on one context they run the other
//     context's top-B, history, re-materialisation and next schedule (and, between coder-blocks, emit / queue / load / KL).
//   Scores-ready and state-ready hand-offs are named barriers used as producer/consumer pairs (bar.arrive on one side,
//   bar.sync on the other).  Two earlier versions let each context own its warps: bound to quarters, a context owned two
//   schedulers and nothing overlapped; spread over all schedulers, the two contexts fell into lock-step (the one that lags
//   gets the whole pipe while the other is in a serial phase and catches up), so their serial phases coincided -- 108 k
//   cycles of joint scoring + 40 k of joint serial work per variable pair, no faster than one context.
// The beams are split over the quarters in parts of HB = 5: with NQB = BMAX / 5 beam parts, quarter q holds part
// bp = q % NQB of BOTH contexts and its warps score sample part sp = q / NQB (BMAX = 20: four beam parts, every warp scores
// all samples against 5 beams; BMAX = 10: two beam parts x two sample halves, the parts replicated in quarters q and q + 2).
// Quarter q, lane (row r, chunk l):
//   columns [160 c, 160 c + 160)        the 5 beams of part bp of context c: beam lb, dim 32 l + i  at column 160 c + 32 lb + i
//   columns [320 + 96 c, 320 + 96 c + 96) sigma_aux, A, E of chunk l for context c, 32 columns each
// = all 512 columns; the fourth coefficient array (M) stays in shared memory (one LDS.128 per 60 gathers).  Shared memory
// also keeps the quantile table (shared by both contexts), per context the scores / top-B scratch and one 40 KB staging
// buffer through which re-materialised beams and new coefficients reach the quarters that hold them.
#define IREC_R2_DEVICE_ONLY
#include "irec_resident2.cuh"
#include "irec_host.h"
#include <stdio.h>

#define TM_THREADS 512
#define TM_SCORE_WARPS 12          // warps 0..11: three per quarter
#define TM_SERIAL_THREADS 128      // warps 12..15: one per quarter
#define TM_BAR_SERIAL 1            // named barriers: serial warps among themselves
#define TM_BAR_SCORES0 2           //   + c: scores of context c ready  (scoring warps arrive, serial warps wait)
#define TM_BAR_STATE0 4            //   + c: state of context c ready   (serial warps arrive, scoring warps wait)
#define TM_HB 5                    // beams per quarter and context
#define TM_COLS 512
#ifndef TM_NSBIG
#define TM_NSBIG 3                 // sample groups per warp and round in the main scoring rounds
#endif
#define TM_COEF_COL0 320           // first coefficient column (after 2 contexts x 5 beams x 32 dims)

// ---------------------------------------------------------------------------------------------
// tcgen05 wrappers (PTX ISA: tcgen05.alloc/dealloc/ld/st/wait/fence).  All .sync.aligned: every lane of the warp executes.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t tm_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void tm_alloc(uint32_t* slot)
{
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tm_smem_u32(slot)), "r"(TM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tm_dealloc(uint32_t taddr)
{
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(TM_COLS) : "memory");
}
__device__ __forceinline__ void tm_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tm_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tm_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

__device__ __forceinline__ void tm_ld4(uint32_t taddr, float (&r)[4])
{
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0, %1, %2, %3}, [%4];"
                 : "=f"(r[0]), "=f"(r[1]), "=f"(r[2]), "=f"(r[3]) : "r"(taddr));
}
__device__ __forceinline__ void tm_ld16(uint32_t taddr, float (&r)[16])
{
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
                 : "=f"(r[0]), "=f"(r[1]), "=f"(r[2]), "=f"(r[3]), "=f"(r[4]), "=f"(r[5]), "=f"(r[6]), "=f"(r[7]),
                   "=f"(r[8]), "=f"(r[9]), "=f"(r[10]), "=f"(r[11]), "=f"(r[12]), "=f"(r[13]), "=f"(r[14]), "=f"(r[15])
                 : "r"(taddr));
}
__device__ __forceinline__ void tm_st4(uint32_t taddr, const float4 v)
{
    asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1, %2, %3, %4};" ::"r"(taddr), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
__device__ __forceinline__ void tm_st16(uint32_t taddr, const float4 a, const float4 b, const float4 c, const float4 d)
{
    asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
                 ::"r"(taddr), "f"(a.x), "f"(a.y), "f"(a.z), "f"(a.w), "f"(b.x), "f"(b.y), "f"(b.z), "f"(b.w),
                   "f"(c.x), "f"(c.y), "f"(c.z), "f"(c.w), "f"(d.x), "f"(d.y), "f"(d.z), "f"(d.w) : "memory");
}
// tcgen05.wait::ld with the loaded registers as in/out operands: nothing that uses them can be scheduled above the wait
__device__ __forceinline__ void tm_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tm_tie4(float (&r)[4]) { asm volatile("" : "+f"(r[0]), "+f"(r[1]), "+f"(r[2]), "+f"(r[3])); }
__device__ __forceinline__ void tm_tie16(float (&r)[16])
{
    asm volatile("" : "+f"(r[0]), "+f"(r[1]), "+f"(r[2]), "+f"(r[3]), "+f"(r[4]), "+f"(r[5]), "+f"(r[6]), "+f"(r[7]),
                      "+f"(r[8]), "+f"(r[9]), "+f"(r[10]), "+f"(r[11]), "+f"(r[12]), "+f"(r[13]), "+f"(r[14]), "+f"(r[15]));
}

// ---------------------------------------------------------------------------------------------
// Chunk sums of NS candidate samples against the HB beams of this warp's half.  Same float32 operation order per
// candidate-dim as r2_score_chunk / the oracle (beam_score):  x = beam + T2[a + c_b] * sigma_aux;  d = x - M;
// acc = fma(fma(A, d, E), d, acc).   tm_beams / tm_coef: TMEM addresses (quarter lane base + first column).
// ---------------------------------------------------------------------------------------------
//   TAB: exponents from the launch table (row[k] = uint2 index of sample k's row + chunk); otherwise from Philox + dl4 in place
//   (j_base[k] = s_k * D + first dim of the chunk) with the one-shot bank spreading of the general path.
template <int HB, int NS, bool TAB>
__device__ __forceinline__ void tm_score_chunk(const char* __restrict__ T2b, uint32_t tm_beams, uint32_t tm_coef,
                                               const float4* __restrict__ M4, int lg,
                                               const uint32_t (&cb)[HB], int P, const uint2* __restrict__ tab_t,
                                               const uint32_t (&row)[NS], const uint16_t* __restrict__ dl4, const TfStream& st,
                                               const uint64_t (&j_base)[NS], float (&acc)[NS][HB])
{
    constexpr int G = R2Group<HB>::G;
    uint2 nxt[NS];
    if (TAB) {
#pragma unroll
        for (int k = 0; k < NS; ++k) nxt[k] = __ldg(tab_t + row[k]);
    }
#pragma unroll 1
    for (int iq = 0; iq < 8; ++iq) {
        uint32_t ad[NS][4];
        if (TAB) {
#pragma unroll
            for (int k = 0; k < NS; ++k) r2_unpack(nxt[k], ad[k][0], ad[k][1], ad[k][2], ad[k][3]);
            if (iq < 7) {
#pragma unroll
                for (int k = 0; k < NS; ++k) nxt[k] = __ldg(tab_t + row[k] + (iq + 1) * P);
            }
        } else {
#pragma unroll
            for (int k = 0; k < NS; ++k) {
                const uint64_t j = j_base[k] + 4 * iq;
                const uint4 u = ((j & 3) == 0) ? tf_stream_group(st, j >> 2) : tf_stream_quad_at(st, j);
                ad[k][0] = r2_spread_banks(r2_exp4(dl4, u.x)); ad[k][1] = r2_spread_banks(r2_exp4(dl4, u.y));
                ad[k][2] = r2_spread_banks(r2_exp4(dl4, u.z)); ad[k][3] = r2_spread_banks(r2_exp4(dl4, u.w));
            }
        }
        float sa[4], A[4], E[4];
        tm_ld4(tm_coef + 4 * iq, sa);
        tm_ld4(tm_coef + 32 + 4 * iq, A);
        tm_ld4(tm_coef + 64 + 4 * iq, E);
        float bm[HB][4];
#pragma unroll
        for (int b = 0; b < HB; ++b) tm_ld4(tm_beams + 32 * b + 4 * iq, bm[b]);
        const float4 nMq = M4[iq * P + lg];           // shared memory holds -M (x - M == x + (-M) bit for bit)
        tm_wait_ld();
        tm_tie4(sa); tm_tie4(A); tm_tie4(E);
#pragma unroll
        for (int b = 0; b < HB; ++b) tm_tie4(bm[b]);
        // dims (0,1) and (2,3) of the quad as packed pairs: v = T*sa, x = beam + v, d = x - M, t = fma(A, d, E) are four packed
        // instructions per pair; the accumulation a = fma(t, d, a) stays scalar and sequential in d (the canonical order)
        const f32x2_t sa2[2] = { f2_pack(sa[0], sa[1]), f2_pack(sa[2], sa[3]) };
        const f32x2_t A2[2] = { f2_pack(A[0], A[1]), f2_pack(A[2], A[3]) };
        const f32x2_t E2[2] = { f2_pack(E[0], E[1]), f2_pack(E[2], E[3]) };
        const f32x2_t nM2[2] = { f2_pack(nMq.x, nMq.y), f2_pack(nMq.z, nMq.w) };
#pragma unroll
        for (int b0 = 0; b0 < HB; b0 += G) {
#pragma unroll
            for (int k = 0; k < NS; ++k) {
                float tv[G][4];
#pragma unroll
                for (int g = 0; g < G; ++g) {
#pragma unroll
                    for (int e = 0; e < 4; ++e)
                        tv[g][e] = *reinterpret_cast<const float*>(T2b + (ad[k][e] + cb[b0 + g]));
                }
#pragma unroll
                for (int g = 0; g < G; ++g) {
                    float a = acc[k][b0 + g];
#pragma unroll
                    for (int h2 = 0; h2 < 2; ++h2) {
                        const f32x2_t x = f2_add(f2_pack(bm[b0 + g][2 * h2], bm[b0 + g][2 * h2 + 1]),
                                                 f2_mul(f2_pack(tv[g][2 * h2], tv[g][2 * h2 + 1]), sa2[h2]));
                        const f32x2_t d = f2_add(x, nM2[h2]);
                        const f32x2_t t = f2_fma(A2[h2], d, E2[h2]);
                        float d0, d1, t0, t1;
                        f2_unpack(d, d0, d1);
                        f2_unpack(t, t0, t1);
                        a = __fmaf_rn(t0, d0, a);
                        a = __fmaf_rn(t1, d1, a);
                    }
                    acc[k][b0 + g] = a;
                }
            }
        }
    }
}

// one round of a warp: NS sample groups against the HB beams [boff, boff + HB) of its half -> s_scores
struct TmSrc {             // where the exponents of a partition come from
    const uint2* tab_t;     // launch table of this block size and auxiliary variable, or nullptr
    int row_stride;
    const uint16_t* dl4;
    TfStream st;
};

template <int HB, int NS, bool TAB>
__device__ __forceinline__ void tm_score_round(const char* T2b, uint32_t tm_beams, uint32_t tm_coef, const float4* M4, const uint32_t* s_cb,
                                               const BeamGeom& g, int lane, const TmSrc& src, int sg_first,
                                               int sg_stride, int S, int Bcur, int boff, float* s_scores)
{
    const int lg = lane & (g.P - 1);
    uint32_t row[NS];
    uint64_t jb[NS];
#pragma unroll
    for (int k = 0; k < NS; ++k) {
        const int sk = (sg_first + k * sg_stride) * g.SPW + lane / g.P;
        const int sc = min(sk, S - 1);
        row[k] = (uint32_t)(sc * src.row_stride + lg);
        jb[k] = TAB ? 0ull : (uint64_t)sc * (uint64_t)g.D + (uint64_t)(32 * lg);
    }
    uint32_t cb[HB];
#pragma unroll
    for (int b = 0; b < HB; ++b) cb[b] = s_cb[boff + b];
    float acc[NS][HB];
#pragma unroll
    for (int k = 0; k < NS; ++k)
#pragma unroll
        for (int b = 0; b < HB; ++b) acc[k][b] = 0.f;
    tm_score_chunk<HB, NS, TAB>(T2b, tm_beams, tm_coef, M4, lg, cb, g.P, src.tab_t, row, src.dl4, src.st, jb, acc);
    float v[NS * HB];
#pragma unroll
    for (int k = 0; k < NS; ++k)
#pragma unroll
        for (int b = 0; b < HB; ++b) v[k * HB + b] = acc[k][b];
    const R2LocalSink sink{ s_scores, S, Bcur, boff };
    r2_tree_store<NS * HB, NS * HB, HB, 0, R2LocalSink>(v, g.P, lane, (sg_first * g.SPW + lane / g.P), sg_stride * g.SPW, sink);
}

// all candidates of one partition that belong to this warp: sample groups u, u + NW, ... (u = this warp's index among the NW
// warps of its context that score the same beam part) against the HB beams of its part
template <int HB, bool TAB>
__device__ __forceinline__ void tm_score_partition(const char* T2b, uint32_t tm_beams, uint32_t tm_coef, const float4* M4,
                                                   const uint32_t* s_cb, const BeamGeom& g, int lane, int u, int NW, const TmSrc& src,
                                                   int S, int Bcur, int boff, float* s_scores)
{
    const int nsg = (S + g.SPW - 1) / g.SPW;
    int sg = 0;
    // rounds of NSBIG sample groups per warp while that many are left (more accumulators per gathered operand: fewer TMEM /
    // exponent loads per gather and fewer gathers in flight per warp), then rounds of 3, 2, 1 (all warps take the same branch)
    if (HB > 1) {
        while (nsg - sg >= TM_NSBIG * NW) {
            tm_score_round<HB, TM_NSBIG, TAB>(T2b, tm_beams, tm_coef, M4, s_cb, g, lane, src, sg + u, NW, S, Bcur, boff, s_scores);
            sg += TM_NSBIG * NW;
        }
    }
    while (nsg - sg > 2 * NW) {
        if (sg + u < nsg)   // groups beyond nsg are clamped inside (scores not stored)
            tm_score_round<HB, 3, TAB>(T2b, tm_beams, tm_coef, M4, s_cb, g, lane, src, sg + u, NW, S, Bcur, boff, s_scores);
        sg += 3 * NW;
    }
    if (nsg - sg > NW) {
        if (sg + u < nsg)
            tm_score_round<HB, 2, TAB>(T2b, tm_beams, tm_coef, M4, s_cb, g, lane, src, sg + u, NW, S, Bcur, boff, s_scores);
        sg += 2 * NW;
    } else if (nsg - sg > 0) {
        if (sg + u < nsg)
            tm_score_round<HB, 1, TAB>(T2b, tm_beams, tm_coef, M4, s_cb, g, lane, src, sg + u, NW, S, Bcur, boff, s_scores);
        sg += NW;
    }
}

struct TmemArgs {
    const float* t_loc; const float* t_scale; const float* p_loc; const float* p_scale;
    const int64_t* gidx; const int64_t* offs; int nb;
    float omega; int S; int B; int64_t seed;
    int32_t* out_indices; int max_aux; int32_t* out_n_aux; int32_t* out_status; float* out_sample;
    const float* T2; const uint16_t* dl4; const float* ratio_tab; int ratio_len;
    int2* hist;            // [2 * gridDim.x][max_aux][BMAX]
    int* work_counter;     // dynamic block queue
    int DPmax;             // padded dims capacity (multiple of 32, <= 1024)
    int NC;                // capacity of a context's score array (>= S * BMAX)
    float* sched;          // [2 * gridDim.x][4][DPmax] per-context scratch: sigma_p^2, sigma_t^2, delta mu, cumulative variance
    const int32_t* order;  // queue position -> coder-block (largest blocks first), or nullptr
    const R2Plan* plan;    // distinct block sizes with an exponent table (nullptr: no table)
    const uint2* tab;      // [R2_MAX_SIZES][tab_aux][S][DPmax / 4]
    int tab_aux;
    const uint2* tab_priv; // the launch's own table (workspace), used when plan->use_private
    int score_lock;        // unused (kept for A/B builds)
    long long* prof;       // nullptr, or [2 * gridDim.x][8] cycle counters per context (IREC_TM_PROFILE=1; diagnostics)
};

// shared memory: quantile table | scratch of the serial warps (one context at a time) | per context: scores, M, next coefficients
template <int BMAX>
__host__ __device__ constexpr size_t tm_stage_floats(int DPmax)
{
    return (size_t)BMAX * (DPmax / 2);             // the parents' 16 dims per chunk of one re-materialisation round
}
template <int BMAX>
__host__ __device__ constexpr size_t tm_serial_bytes(int DPmax)
{
    return 32 * sizeof(double) + sizeof(float) * (256 + 32 + tm_stage_floats<BMAX>(DPmax)) + sizeof(int32_t) * (32 + R2_TOPK_CAP + 4 + 4);
}
__host__ __device__ constexpr size_t tm_ctx_bytes(int DPmax, int NC)
{
    return sizeof(float) * ((size_t)NC + (size_t)DPmax + 4 * (size_t)DPmax) + sizeof(int32_t) * (64 + 32);
}
template <int BMAX>
__host__ __device__ constexpr size_t tm_smem_bytes(int DPmax, int NC)
{
    return sizeof(float) * (size_t)IREC_T2_LEN + tm_serial_bytes<BMAX>(DPmax) + 2 * tm_ctx_bytes(DPmax, NC) + 16;
}

__device__ __forceinline__ void tm_bar_arrive(int id) { asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(TM_THREADS) : "memory"); }
__device__ __forceinline__ void tm_bar_wait(int id) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(TM_THREADS) : "memory"); }

// what the serial warps tell the scoring warps about a context (shared memory; valid after the state-ready barrier)
struct TmCtl {
    int finished;          // 1: no more coder-blocks for this context (the scoring warps stop visiting it)
    int D;                 // dims of the current coder-block
    int Bcur;              // beams alive
    int t;                 // auxiliary variable to score
    unsigned long long tab_blk;      // exponent table of this block size (const uint2*), 0 = in-place exponents
};

// KL, n_aux and status of every coder-block (coder.py:499-501), ahead of the persistent kernel: one CTA per block over the
// whole GPU instead of 128 serial threads per SM evaluating float64 log / expm1 between coder-blocks
__global__ void __launch_bounds__(256) k_tm_kl_status(const float* __restrict__ t_loc, const float* __restrict__ t_scale,
                                                      const float* __restrict__ p_loc, const float* __restrict__ p_scale,
                                                      const int64_t* __restrict__ gidx, const int64_t* __restrict__ offs, int nb,
                                                      float omega, int max_aux, int ratio_len, int32_t* __restrict__ out_n_aux,
                                                      int32_t* __restrict__ out_status)
{
    __shared__ double cs[32];
    for (int blk = blockIdx.x; blk < nb; blk += gridDim.x) {
        const int64_t off = offs[blk];
        const int D = (int)(offs[blk + 1] - off);
        const int nch = (D + 31) >> 5;               // <= 32 (D <= 1024)
        // canonical order: every 32-dim chunk is summed in ascending d by one thread, the chunk sums by the pairwise tree
        for (int ch = threadIdx.x; ch < nch; ch += blockDim.x) {
            double acc = 0.0;
            const int hi = min(D, 32 * ch + 32);
            for (int d = 32 * ch; d < hi; ++d) {
                const int64_t gi = gidx ? gidx[off + d] : off + d;
                acc = __dadd_rn(acc, kl_dim(t_loc[gi], t_scale[gi], p_loc[gi], p_scale[gi]));
            }
            cs[ch] = acc;
        }
        const double kl = block_tree_sum_f64(cs, nch);
        if (threadIdx.x == 0) {
            const int n_aux = n_aux_from_kl((float)kl, omega);
            int status = IREC_BLK_OK;
            if (n_aux <= 0) status = IREC_BLK_BAD_KL;
            else if (n_aux > max_aux || n_aux > ratio_len) status = IREC_BLK_TOO_LONG;
            out_n_aux[blk] = n_aux;
            out_status[blk] = status;
        }
        __syncthreads();
    }
}

// Queue order: longest coder-blocks first (work ~ auxiliary variables x 32-dim chunks), so that the blocks handed out last --
// the ones that decide when the launch ends -- are the short ones; but only in TM_ORDER_CLASSES coarse work classes, and in
// the caller's order inside a class (stable counting sort, single CTA).  The caller's order keeps the coder-blocks of one
// tensor together: Coder.split scatters a tensor's dims over its blocks, so every block touches most 32-byte sectors of the
// tensor's four parameter arrays and of its sample; coded at about the same time by neighbouring CTAs they share those sectors
// in L2, coded far apart (the former order by exact work) every block fetched them from DRAM again -- 1.0 GB read per
// configs[3]-size launch against 0.17 GB algorithmic (profiles/r2_tmem_f_ncu.md).
#define TM_ORDER_CLASSES 8
#define TM_ORDER_THREADS 512
__global__ void __launch_bounds__(TM_ORDER_THREADS) k_tm_order(const int64_t* __restrict__ offs, const int32_t* __restrict__ n_aux,
                                                               const int32_t* __restrict__ status, int nb, int32_t* __restrict__ order)
{
    __shared__ int s_cnt[TM_ORDER_CLASSES][TM_ORDER_THREADS];     // blocks of class c in thread t's contiguous chunk -> start position
    __shared__ int s_max;
    auto work = [&](int b) {
        if (status[b] != IREC_BLK_OK) return 0;
        const int64_t D = offs[b + 1] - offs[b];
        return (int)min((int64_t)0x3fffff, (int64_t)n_aux[b] * ((D + 31) >> 5));
    };
    const int tid = threadIdx.x;
    if (tid == 0) s_max = 1;
    __syncthreads();
    int mx = 1;
    for (int b = tid; b < nb; b += TM_ORDER_THREADS) mx = max(mx, work(b));
    atomicMax(&s_max, mx);
    __syncthreads();
    const int wmax = s_max;
    auto cls = [&](int w) { return (TM_ORDER_CLASSES - 1) - (int)(((int64_t)w * (TM_ORDER_CLASSES - 1)) / wmax); };   // 0 = most work
    const int per = (nb + TM_ORDER_THREADS - 1) / TM_ORDER_THREADS;
    const int b0 = min(nb, tid * per), b1 = min(nb, b0 + per);
    int mine[TM_ORDER_CLASSES];
#pragma unroll
    for (int c = 0; c < TM_ORDER_CLASSES; ++c) mine[c] = 0;
    for (int b = b0; b < b1; ++b) {
        const int c = cls(work(b));
#pragma unroll
        for (int k = 0; k < TM_ORDER_CLASSES; ++k) mine[k] += (k == c);
    }
#pragma unroll
    for (int c = 0; c < TM_ORDER_CLASSES; ++c) s_cnt[c][tid] = mine[c];
    __syncthreads();
    if (tid < TM_ORDER_CLASSES) {                  // per class: exclusive prefix over the threads (chunks in block order)
        int run = 0;
        for (int t = 0; t < TM_ORDER_THREADS; ++t) { const int v = s_cnt[tid][t]; s_cnt[tid][t] = run; run += v; }
        mine[0] = run;                             // class total, parked in a register of thread `tid`
    }
    __shared__ int s_base[TM_ORDER_CLASSES + 1];
    if (tid < TM_ORDER_CLASSES) s_base[tid + 1] = mine[0];
    __syncthreads();
    if (tid == 0) {
        s_base[0] = 0;
        for (int c = 0; c < TM_ORDER_CLASSES; ++c) s_base[c + 1] += s_base[c];
    }
    __syncthreads();
    int pos[TM_ORDER_CLASSES];
#pragma unroll
    for (int c = 0; c < TM_ORDER_CLASSES; ++c) pos[c] = s_base[c] + s_cnt[c][tid];
    for (int b = b0; b < b1; ++b) {
        const int c = cls(work(b));
        int p = 0;
#pragma unroll
        for (int k = 0; k < TM_ORDER_CLASSES; ++k)
            if (k == c) { p = pos[k]; pos[k] = p + 1; }
        order[p] = b;
    }
}

template <int BMAX>
__global__ void __launch_bounds__(TM_THREADS, 1) k_beam_encode_tmem(const TmemArgs a)
{
    static_assert(BMAX == 5 || BMAX == 10 || BMAX == 20, "beam parts of 5 over 1, 2 or 4 TMEM lane quarters");
    constexpr int HB = TM_HB;
    constexpr int NQB = BMAX / HB;                 // beam parts (quarters that hold distinct beams)
    constexpr int NSP = 4 / NQB;                   // sample parts (replicas of a beam part)
    constexpr int NW = 3 * NSP;                    // scoring warps that score the same beam part
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ uint32_t s_tmem_base;
    __shared__ TmCtl s_ctl2[2];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int q = warp & 3, wi = warp >> 2;        // scheduler / TMEM lane quarter, warp of the quarter (3 = the serial warp)
    const int bp = q % NQB, sp = q / NQB;          // beam part held by this quarter, sample part scored by it
    const int DPm = a.DPmax;
    const int row_stride = DPm >> 2;

    // ---- shared memory carve-up ----
    float* s_T2 = reinterpret_cast<float*>(smem_raw);                                    // [IREC_T2_LEN], both contexts
    unsigned char* ser_base = smem_raw + sizeof(float) * (size_t)IREC_T2_LEN;            // scratch of the serial warps
    auto ctx_base = [&](int c) { return ser_base + tm_serial_bytes<BMAX>(DPm) + (size_t)c * tm_ctx_bytes(DPm, a.NC); };
    auto ctx_scores = [&](int c) { return reinterpret_cast<float*>(ctx_base(c)); };      // [NC]
    auto ctx_M = [&](int c) { return ctx_scores(c) + a.NC; };                            // [DPm] NEGATED auxiliary-target means (-M), CI layout
    auto ctx_next = [&](int c) { return ctx_M(c) + DPm; };                               // [4][DPm] coefficients of the NEXT variable
    auto ctx_hsum = [&](int c) { return reinterpret_cast<int32_t*>(ctx_next(c) + 4 * DPm); };   // [2][32]
    auto ctx_cb = [&](int c) { return reinterpret_cast<uint32_t*>(ctx_hsum(c) + 64); };  // [32] 4 * dlog(h_b)

    {
        const float4* src = reinterpret_cast<const float4*>(a.T2);
        float4* dst = reinterpret_cast<float4*>(s_T2);
        for (int i = threadIdx.x; i < IREC_T2_LEN / 4; i += blockDim.x) dst[i] = src[i];
    }
    if (warp == 0) tm_alloc(&s_tmem_base);
    if (threadIdx.x < 2) { s_ctl2[threadIdx.x].finished = 0; }
    tm_fence_before();
    __syncthreads();
    tm_fence_after();
    const uint32_t tm_q = s_tmem_base + (((uint32_t)q * 32u) << 16);                     // this warp's lane quarter
    const char* T2b = reinterpret_cast<const char*>(s_T2);

    if (warp < TM_SCORE_WARPS) {
        // ======================================= scoring warps =======================================
        const int u = sp * 3 + wi;                 // index among the NW warps that score beam part bp
        bool done[2] = { false, false };
        for (int turn = 0; !(done[0] && done[1]); ++turn) {
            const int c = turn & 1;
            if (done[c]) continue;
            tm_bar_wait(TM_BAR_STATE0 + c);        // the serial warps have prepared (or retired) context c
            tm_fence_after();
            const TmCtl ctl = s_ctl2[c];
            if (ctl.finished) { done[c] = true; continue; }
            const BeamGeom g = make_geom(ctl.D);
            const uint32_t tm_beams = tm_q + 160u * c;
            const uint32_t tm_coef = tm_q + TM_COEF_COL0 + 96u * c;
            const float4* M4 = reinterpret_cast<const float4*>(ctx_M(c));
            const uint32_t* s_cb = ctx_cb(c);
            float* s_scores = ctx_scores(c);
            const uint2* tab_blk = reinterpret_cast<const uint2*>(ctl.tab_blk);
            TmSrc src;
            src.tab_t = tab_blk ? tab_blk + (size_t)ctl.t * a.S * row_stride : nullptr;
            src.row_stride = row_stride; src.dl4 = a.dl4; src.st = tf_stream_seeded(a.seed + ctl.t, a.seed + ctl.t);
            const int Bcur = ctl.Bcur;
            // ---- score all S * Bcur candidates (beam_search_coder.py:79-84,97-102): this warp's beam part, its share of the samples ----
            if (src.tab_t) {
                if (Bcur == 1) {
                    if (bp == 0) tm_score_partition<1, true>(T2b, tm_beams, tm_coef, M4, s_cb, g, lane, u, NW, src, a.S, 1, 0, s_scores);
                } else if (bp * HB < Bcur) {
                    tm_score_partition<HB, true>(T2b, tm_beams, tm_coef, M4, s_cb, g, lane, u, NW, src, a.S, Bcur, bp * HB, s_scores);
                }
            } else {
                if (Bcur == 1) {
                    if (bp == 0) tm_score_partition<1, false>(T2b, tm_beams, tm_coef, M4, s_cb, g, lane, u, NW, src, a.S, 1, 0, s_scores);
                } else if (bp * HB < Bcur) {
                    tm_score_partition<HB, false>(T2b, tm_beams, tm_coef, M4, s_cb, g, lane, u, NW, src, a.S, Bcur, bp * HB, s_scores);
                }
            }
            tm_fence_before();
            tm_bar_arrive(TM_BAR_SCORES0 + c);
        }
    } else {
        // ======================================= serial warps =======================================
        const WarpGroup grp{ q * 32 + lane, TM_SERIAL_THREADS, TM_BAR_SERIAL };
        const int tid = grp.tid();
        constexpr int nt = TM_SERIAL_THREADS;
        auto ser_sync = [&]() { tm_fence_before(); grp.sync(); tm_fence_after(); };
        // scratch shared by both contexts (the serial warps work on one context at a time)
        double* s_kl = reinterpret_cast<double*>(ser_base);                                  // [32] (unused since k_tm_kl_status; keeps the layout aligned)
        float* s_gmax = reinterpret_cast<float*>(s_kl + 32);                                 // [256]
        float* s_wsc = s_gmax + 256;                                                         // [32] winners' scores
        float* s_stage = s_wsc + 32;                                                         // parents of a round: [BMAX][4][P] float4
        int32_t* s_wid = reinterpret_cast<int32_t*>(s_stage + tm_stage_floats<BMAX>(DPm));   // [32] winners' flat ids
        int32_t* s_list = s_wid + 32;                                                        // [R2_TOPK_CAP]
        int32_t* s_ctl = s_list + R2_TOPK_CAP;                                               // [4]
        int32_t* s_misc = s_ctl + 4;                                                         // [4]
        float4* stage4 = reinterpret_cast<float4*>(s_stage);

        // per-context state that lives in this thread's registers across the turns
        struct Blk { int active, blk, D, n_aux, t, Bcur, hb; int64_t off; const uint2* tab_blk; };
        Blk st[2];
        st[0].active = 0; st[1].active = 0;
        bool retired[2] = { false, false };
        bool first = true;                         // context 0 starts with queue position blockIdx.x (see below)
        long long pt = 0;
        if (a.prof && tid == 0) pt = clock64();

        for (int turn = 0; !(retired[0] && retired[1]); ++turn) {
            const int c = turn & 1;
            if (retired[c]) continue;
            Blk& B_ = st[c];
            // ---- this context's shared memory, global scratch and TMEM columns ----
            float* s_scores = ctx_scores(c);
            float* s_M = ctx_M(c);
            float* s_next = ctx_next(c);
            int32_t* s_hsum = ctx_hsum(c);
            uint32_t* s_cb = ctx_cb(c);
            const uint32_t tm_beams = tm_q + 160u * c;
            const uint32_t tm_coef = tm_q + TM_COEF_COL0 + 96u * c;
            const int slot = 2 * blockIdx.x + c;
            int2* hist = a.hist + (size_t)slot * a.max_aux * BMAX;
            float* g_cv = a.sched + (size_t)slot * 4 * DPm;                                  // global scratch (this context only), CI layout
            float* g_tv = g_cv + DPm; float* g_dmu = g_tv + DPm; float* g_cum = g_dmu + DPm;
            auto lap = [&](int phase) {
                if (a.prof && tid == 0) {
                    const long long now = clock64();
                    a.prof[(size_t)slot * 8 + phase] += now - pt;
                    pt = now;
                }
            };
            // Coefficients of an auxiliary variable (beam_search_coder.py:64-77) into the context's NEXT buffer (CI layout, zeros
            // in the padding).  They do not depend on the winners, so they are computed while the scoring warps still score the
            // context's current variable.
            auto compute_next = [&](const BeamGeom& g, float ratio) {
                // all global loads of the thread's (at most 8) dims first: the L2 round trips overlap instead of adding up
                float in[8][4];
#pragma unroll
                for (int r = 0; r < 8; ++r) {
                    const int i = tid + r * nt;
                    const bool on = i < g.DP;
                    in[r][0] = on ? g_cv[i] : 0.f; in[r][1] = on ? g_tv[i] : 0.f;
                    in[r][2] = on ? g_dmu[i] : 0.f; in[r][3] = on ? g_cum[i] : 0.f;
                }
#pragma unroll
                for (int r = 0; r < 8; ++r) {
                    const int i = tid + r * nt;
                    if (i < g.DP) {
                        SchedOut o;
                        o.sa = 0.f; o.A = 0.f; o.E = 0.f; o.M = 0.f; o.cum_next = 0.f;
                        if (in[r][0] != 0.f) {
                            o = beam_sched_dim(in[r][0], in[r][1], in[r][2], in[r][3], ratio);
                            g_cum[i] = o.cum_next;
                        }
                        s_next[i] = o.sa; s_next[DPm + i] = o.A; s_next[2 * DPm + i] = o.E; s_next[3 * DPm + i] = o.M;
                    }
                }
            };
            // NEXT buffer -> the scoring warps' operands: sigma_aux / A / E into the context's TMEM columns of every quarter, M into
            // shared memory, table offsets c_b of the new beams.  Callers synchronise before (NEXT complete) and hand over after.
            auto install_next = [&](const BeamGeom& g, const int32_t* hs_new, int Knew) {
                const int lg = lane & (g.P - 1);
                const float4* next4 = reinterpret_cast<const float4*>(s_next);
                for (int pr = 0; pr < 24; ++pr) {                              // (array, quad) pairs
                    const int arr = pr >> 3, iq = pr & 7;
                    tm_st4(tm_coef + 32 * arr + 4 * iq, next4[arr * (DPm >> 2) + iq * g.P + lg]);
                }
                for (int i = tid; i < g.DP; i += nt) s_M[i] = -s_next[3 * DPm + i];       // the scoring loop adds -M
                if (tid < 32) s_cb[tid] = tid < Knew ? (uint32_t)__ldg(a.dl4 + (hash_from_sum(hs_new[tid]) - 1)) : 0u;
                tm_wait_st();
                // every serial warp has read NEXT: the context's next compute_next may overwrite it (it follows at once when the
                // other context has retired -- found by compute-sanitizer racecheck, profiles/r2b_sanitizer_racecheck.log)
                ser_sync();
            };

            if (B_.active) {
                const BeamGeom g = make_geom(B_.D);
                const int lg = lane & (g.P - 1);
                const int D = B_.D, t = B_.t, Bcur = B_.Bcur;
                // =========== while the scoring warps score variable t of this context: coefficients of variable t + 1 ===========
                if (t + 1 < B_.n_aux) compute_next(g, a.ratio_tab[B_.n_aux - 2 - t]);
                lap(4);
                // =========== scores of variable t ready: select, re-materialise, install t + 1 ===========
                tm_bar_wait(TM_BAR_SCORES0 + c);
                tm_fence_after();
                lap(5);
                const int32_t* hs = s_hsum + 32 * B_.hb;
                TmSrc src;
                src.tab_t = B_.tab_blk ? B_.tab_blk + (size_t)t * a.S * row_stride : nullptr;
                src.row_stride = row_stride; src.dl4 = a.dl4; src.st = tf_stream_seeded(a.seed + t, a.seed + t);

                // ---- top-B (beam_search_coder.py:86-89,104-106) ----
                const int Kout = block_topk(s_scores, nullptr, a.S * Bcur, a.B, s_wsc, s_wid, s_gmax, s_list, R2_TOPK_CAP, s_ctl, grp);

                // ---- history + hash sums of the new beams (:92-95); (s_j, b_j) for the re-materialisation ----
                int32_t* hs_new = s_hsum + 32 * (B_.hb ^ 1);
                if (tid < Kout) {
                    const int f = s_wid[tid];
                    const int sj = f / Bcur, bj = f - sj * Bcur;
                    hist[(size_t)t * BMAX + tid] = make_int2(sj, bj);
                    hs_new[tid] = hsum_extend(hs[bj], sj, t);
                    s_list[tid] = sj; s_list[32 + tid] = bj;
                }
                ser_sync();
                lap(2);

                // ---- re-materialise the winners: beam_j <- beam_{b_j} + a(s_j, b_j)  (:92-93), 16 dims of every chunk per round:
                //      (1) every quarter copies the 16 dims of its (old) beams into the staging buffer -- parents readable by all;
                //      (2) every quarter computes ITS OWN five slots from the staged parents (lane = chunk) and stores them straight
                //          into its TMEM columns: five winners per serial warp whatever the parents are, no write before all reads.
                for (int rd = 0; rd < 2; ++rd) {
                    uint2 ex[HB][4];                           // exponent rows of this quarter's winners, requested first (L2 latency)
#pragma unroll
                    for (int lb = 0; lb < HB; ++lb) {
                        const int j = bp * HB + lb;
                        const int sj = j < Kout ? s_list[j] : 0;
#pragma unroll
                        for (int qq = 0; qq < 4; ++qq) {
                            ex[lb][qq] = make_uint2(0u, 0u);
                            if (j >= Kout) continue;
                            if (src.tab_t) {
                                ex[lb][qq] = __ldg(src.tab_t + (size_t)sj * row_stride + (4 * rd + qq) * g.P + lg);
                            } else {                                      // 4 x uint16 word offsets, as the table stores them
                                const int d0 = 32 * lg + 4 * (4 * rd + qq);
                                if (d0 < D) {
                                    const uint4 uu = tf_stream_quad_at(src.st, (uint64_t)sj * (uint64_t)D + (uint64_t)d0);
                                    const uint32_t w0 = r2_exp4(a.dl4, uu.x) >> 2, w1 = d0 + 1 < D ? r2_exp4(a.dl4, uu.y) >> 2 : 0u;
                                    const uint32_t w2 = d0 + 2 < D ? r2_exp4(a.dl4, uu.z) >> 2 : 0u, w3 = d0 + 3 < D ? r2_exp4(a.dl4, uu.w) >> 2 : 0u;
                                    ex[lb][qq] = make_uint2(w0 | (w1 << 16), w2 | (w3 << 16));
                                }
                            }
                        }
                    }
                    if (sp == 0) {                             // one replica of every beam part stages its beams
#pragma unroll
                        for (int lb = 0; lb < HB; ++lb) {
                            const int b = bp * HB + lb;
                            if (b < Bcur) {
                                float par[16];
                                tm_ld16(tm_beams + 32 * lb + 16 * rd, par);
                                tm_wait_ld();
                                tm_tie16(par);
                                if (lane < g.P) {
#pragma unroll
                                    for (int qq = 0; qq < 4; ++qq)
                                        stage4[(b * 4 + qq) * g.P + lg] = make_float4(par[4 * qq], par[4 * qq + 1], par[4 * qq + 2], par[4 * qq + 3]);
                                }
                            }
                        }
                    }
                    float sa[16];
                    tm_ld16(tm_coef + 16 * rd, sa);
                    tm_wait_ld();
                    tm_tie16(sa);
                    ser_sync();
#pragma unroll
                    for (int lb = 0; lb < HB; ++lb) {
                        const int j = bp * HB + lb;
                        if (j < Kout) {
                            const int bj = s_list[32 + j];
                            const uint32_t cb = s_cb[bj];
                            float4 o[4];
#pragma unroll
                            for (int qq = 0; qq < 4; ++qq) {
                                uint32_t e0, e1, e2, e3;
                                r2_unpack(ex[lb][qq], e0, e1, e2, e3);
                                const float4 par = stage4[(bj * 4 + qq) * g.P + lg];
                                o[qq].x = __fadd_rn(par.x, __fmul_rn(*reinterpret_cast<const float*>(T2b + e0 + cb), sa[4 * qq + 0]));
                                o[qq].y = __fadd_rn(par.y, __fmul_rn(*reinterpret_cast<const float*>(T2b + e1 + cb), sa[4 * qq + 1]));
                                o[qq].z = __fadd_rn(par.z, __fmul_rn(*reinterpret_cast<const float*>(T2b + e2 + cb), sa[4 * qq + 2]));
                                o[qq].w = __fadd_rn(par.w, __fmul_rn(*reinterpret_cast<const float*>(T2b + e3 + cb), sa[4 * qq + 3]));
                                // padding dims: sigma_aux = 0 and the parent is 0 there, so the padding stays zero
                            }
                            tm_st16(tm_beams + 32 * lb + 16 * rd, o[0], o[1], o[2], o[3]);
                        }
                    }
                    tm_wait_st();
                    ser_sync();                                // staging is reused by the next round / the next turn
                }
                lap(3);
                B_.Bcur = Kout;
                B_.hb ^= 1;
                B_.t = t + 1;
                if (a.prof && tid == 0) a.prof[(size_t)slot * 8 + 6] += 1;
                if (B_.t < B_.n_aux) {
                    // ---- hand the context back with the coefficients and table offsets c_b of the next auxiliary variable ----
                    install_next(g, hs_new, Kout);
                    if (tid == 0) { s_ctl2[c].Bcur = Kout; s_ctl2[c].t = B_.t; }
                    lap(1);
                    tm_fence_before();
                    tm_bar_arrive(TM_BAR_STATE0 + c);
                    continue;
                }
                // ---- coder-block finished: indices of the best beam (trace the back-pointers) and its sample (:118-122) ----
                if (tid == 0) {
                    int j = 0;
                    int32_t* oi = a.out_indices + (size_t)B_.blk * a.max_aux;
                    for (int tt = B_.n_aux - 1; tt >= 0; --tt) {
                        const int2 e = hist[(size_t)tt * BMAX + j];
                        oi[tt] = e.x;
                        j = e.y;
                    }
                }
                if (q == 0) {                          // beam 0 lives in beam part 0 (quarter 0 holds a copy for every BMAX)
                    for (int iq = 0; iq < 8; ++iq) {
                        float b0[4];
                        tm_ld4(tm_beams + 4 * iq, b0);
                        tm_wait_ld();
                        tm_tie4(b0);
                        if (lane < g.P) {
#pragma unroll
                            for (int e = 0; e < 4; ++e) {
                                const int d = 32 * lg + 4 * iq + e;
                                if (d < D) {
                                    const int64_t gi = a.gidx ? a.gidx[B_.off + d] : B_.off + d;
                                    a.out_sample[gi] = __fadd_rn(b0[e], a.p_loc[gi]);
                                }
                            }
                        }
                    }
                }
                if (a.prof && tid == 0) a.prof[(size_t)slot * 8 + 7] += 1;
                B_.active = 0;
            }

            // =========== next coder-block of this context (queue, load, first schedule) ===========
            // Block queue: context 0 of every CTA starts with queue position blockIdx.x, everything else is drawn dynamically
            // from gridDim.x on -- no CTA holds two coder-blocks before every CTA holds one.
            for (;;) {
                ser_sync();
                if (tid == 0) s_misc[0] = (first && c == 0) ? (int)blockIdx.x : (int)gridDim.x + atomicAdd(a.work_counter, 1);
                ser_sync();
                if (c == 0) first = false;
                const int pos = s_misc[0];
                if (pos >= a.nb) {
                    retired[c] = true;
                    if (a.prof && tid == 0) {           // when this context ran out of work (ns, global timer)
                        unsigned long long now;
                        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
                        a.prof[8 * 2 * 1024 + slot] = (long long)now;
                    }
                    if (tid == 0) s_ctl2[c].finished = 1;
                    tm_fence_before();
                    tm_bar_arrive(TM_BAR_STATE0 + c);     // the scoring warps learn that this context is finished
                    break;
                }
                const int blk = a.order ? a.order[pos] : pos;
                if (a.out_status[blk] != IREC_BLK_OK) continue;       // KL not finite / too many auxiliary variables (k_tm_kl_status)
                const int n_aux = a.out_n_aux[blk];
                const int64_t off = a.offs[blk];
                const int D = (int)(a.offs[blk + 1] - off);
                const BeamGeom g = make_geom(D);
                // exponent table of this block size (if it has one)
                const uint2* tab_blk = r2_tab_of_size(a.plan, a.tab, a.tab_aux, a.tab_priv, a.max_aux, a.S, row_stride, D);
                // ---- load (CI layout, zero padding) ----
                for (int i = tid; i < g.DP; i += nt) {
                    const int l = (i >> 2) & (g.P - 1), iq = i / (4 * g.P), e = i & 3;       // CI(P): i = (iq * P + l) * 4 + e
                    const int d = 32 * l + 4 * iq + e;
                    float cv = 0.f, tv = 0.f, dmu = 0.f;
                    if (d < D) {
                        const int64_t gi = a.gidx ? a.gidx[off + d] : off + d;
                        const float tl = a.t_loc[gi], ts = a.t_scale[gi], pl = a.p_loc[gi], ps = a.p_scale[gi];
                        cv = __fmul_rn(ps, ps); tv = __fmul_rn(ts, ts); dmu = __fadd_rn(tl, -pl);
                    }
                    g_cv[i] = cv; g_tv[i] = tv; g_dmu[i] = dmu; g_cum[i] = 0.f;
                }
                {   // the beams of this context start at zero (t = 0 scores the single empty beam; slots >= Bcur are never read back)
                    const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
                    for (int lb = 0; lb < HB; ++lb) {
                        tm_st16(tm_beams + 32 * lb, z, z, z, z);
                        tm_st16(tm_beams + 32 * lb + 16, z, z, z, z);
                    }
                }
                if (tid < 64) s_hsum[tid] = 0;
                compute_next(g, a.ratio_tab[n_aux - 1]);    // same thread, same dims as the load above: no barrier needed in between
                ser_sync();
                install_next(g, s_hsum, 1);                 // empty index row: hash sum 0
                B_.active = 1; B_.blk = blk; B_.D = D; B_.n_aux = n_aux; B_.t = 0; B_.Bcur = 1; B_.hb = 0; B_.off = off; B_.tab_blk = tab_blk;
                if (tid == 0) {
                    s_ctl2[c].D = D; s_ctl2[c].Bcur = 1; s_ctl2[c].t = 0;
                    s_ctl2[c].tab_blk = reinterpret_cast<unsigned long long>(tab_blk);
                }
                lap(0);
                tm_fence_before();
                tm_bar_arrive(TM_BAR_STATE0 + c);
                break;
            }
        }
    }
    tm_fence_before();
    __syncthreads();
    tm_fence_after();
    if (warp == 0) tm_dealloc(s_tmem_base);
}

// =============================================================================================
// host side
// =============================================================================================
static int tm_pick_bmax(int B)
{
    if (B > 1 && B <= 5) return 5;
    if (B > 5 && B <= 10) return 10;
    if (B > 10 && B <= 20) return 20;
    return -1;                                     // B = 1 and B > 20: the resident2 / resident kernels
}

template <int BMAX>
static bool tm_plan_t(int max_D, int S, TmemPlan& p)
{
    if (max_D > 1024 || max_D < 1 || (int64_t)S * BMAX > 32768 || (int64_t)S * BMAX > R2_TOPK_CAP * 16) return false;
    const BeamGeom g = make_geom(max_D);
    p.bmax = BMAX; p.DPmax = g.DP; p.NC = ((S * BMAX + 31) / 32) * 32;
    p.smem = tm_smem_bytes<BMAX>(p.DPmax, p.NC);
    if (p.smem > (size_t)irec_device().max_smem_optin) return false;
    return cudaFuncSetAttribute(k_beam_encode_tmem<BMAX>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)p.smem) == cudaSuccess;
}

bool irec_tmem_plan(int nb, int max_D, int S, int B, TmemPlan* out)
{
    TmemPlan p{};
    bool ok = false;
    switch (tm_pick_bmax(B)) {
        case 5: ok = tm_plan_t<5>(max_D, S, p); break;
        case 10: ok = tm_plan_t<10>(max_D, S, p); break;
        case 20: ok = tm_plan_t<20>(max_D, S, p); break;
        default: ok = false;
    }
    if (!ok) return false;
    // A CTA of this kernel holds the whole register file of its SM (512 threads x 128 registers): nothing else runs there until
    // it retires.  A caller that pipelines sub-batches on several streams leaves a few SMs free (irec_set_thread_reserved_sms)
    // so that the other streams' small kernels -- the callers' networks between two coder launches -- are not starved.
    const int reserve = std::min(irec_reserved_sms(), irec_device().sm_count - 1);
    p.grid = std::max(1, std::min(nb, irec_device().sm_count - reserve));     // two contexts per CTA; see the queue policy in the kernel
    if (out) *out = p;
    return true;
}

int irec_launch_tmem(const TmemPlan& p, const float* t_loc, const float* t_scale, const float* p_loc, const float* p_scale,
                     const int64_t* gidx, const int64_t* offs, int nb, float omega, int S, int B, int64_t seed,
                     int32_t* out_indices, int max_aux, int32_t* out_n_aux, int32_t* out_status, float* out_sample,
                     int2* hist, float* sched, int* work_counter, const int32_t* order, const void* plan, const void* tab, int tab_aux,
                     const void* tab_priv, cudaStream_t s)
{
    TmemArgs a;
    a.t_loc = t_loc; a.t_scale = t_scale; a.p_loc = p_loc; a.p_scale = p_scale;
    a.gidx = gidx; a.offs = offs; a.nb = nb; a.omega = omega; a.S = S; a.B = B; a.seed = seed;
    a.out_indices = out_indices; a.max_aux = max_aux; a.out_n_aux = out_n_aux; a.out_status = out_status;
    a.out_sample = out_sample; a.T2 = irec_device().d_T2; a.dl4 = irec_device().d_dl4;
    a.ratio_tab = irec_ratio_tab(); a.ratio_len = irec_ratio_len();
    a.hist = hist; a.work_counter = work_counter; a.DPmax = p.DPmax; a.NC = p.NC; a.sched = sched; a.order = order;
    a.plan = reinterpret_cast<const R2Plan*>(plan); a.tab = reinterpret_cast<const uint2*>(tab); a.tab_aux = tab_aux;
    a.tab_priv = reinterpret_cast<const uint2*>(tab_priv);
    a.score_lock = 0;
    k_tm_kl_status<<<std::min(nb, 8 * irec_device().sm_count), 256, 0, s>>>(t_loc, t_scale, p_loc, p_scale, gidx, offs, nb, omega, max_aux,
                                                                           a.ratio_len, out_n_aux, out_status);
    irec_count_launch();
    if (order) {
        k_tm_order<<<1, TM_ORDER_THREADS, 0, s>>>(offs, out_n_aux, out_status, nb, const_cast<int32_t*>(order));
        irec_count_launch();
    }
    a.prof = nullptr;
    static long long* d_prof = nullptr;            // diagnostics only (IREC_TM_PROFILE=1): phase cycle counters, dumped to stderr
    const char* pe = getenv("IREC_TM_PROFILE");
    const bool prof_on = pe && pe[0] == '1';
    if (prof_on) {
        if (!d_prof) cudaMalloc(&d_prof, sizeof(long long) * 9 * 2 * 1024);
        cudaMemsetAsync(d_prof, 0, sizeof(long long) * 9 * 2 * 1024, s);
        a.prof = d_prof;
    }
    switch (p.bmax) {
        case 5: k_beam_encode_tmem<5><<<p.grid, TM_THREADS, p.smem, s>>>(a); break;
        case 10: k_beam_encode_tmem<10><<<p.grid, TM_THREADS, p.smem, s>>>(a); break;
        case 20: k_beam_encode_tmem<20><<<p.grid, TM_THREADS, p.smem, s>>>(a); break;
        default: return irec_fail(IREC_E_INVALID, "irec_launch_tmem: unsupported beam capacity");
    }
    irec_count_launch();
    if (prof_on) {
        std::vector<long long> h(8 * 2 * p.grid);
        cudaStreamSynchronize(s);
        cudaMemcpy(h.data(), d_prof, sizeof(long long) * h.size(), cudaMemcpyDeviceToHost);
        std::vector<long long> fin(2 * p.grid);
        cudaMemcpy(fin.data(), d_prof + 8 * 2 * 1024, sizeof(long long) * fin.size(), cudaMemcpyDeviceToHost);
        std::sort(fin.begin(), fin.end());
        if (!fin.empty())
            fprintf(stderr, "[tmem profile] contexts retire (ms before the last one): median %.2f, 10%% %.2f, first %.2f\n",
                    (fin.back() - fin[fin.size() / 2]) * 1e-6, (fin.back() - fin[fin.size() / 10]) * 1e-6, (fin.back() - fin[0]) * 1e-6);
        double tot[8] = { 0 };
        for (int i = 0; i < 2 * p.grid; ++i)
            for (int k = 0; k < 8; ++k) tot[k] += (double)h[(size_t)i * 8 + k];
        const double vars = tot[6] > 0 ? tot[6] : 1;
        fprintf(stderr, "[tmem profile] contexts %d blocks %.0f variables %.0f | serial warps, cycles per variable: next coefficients %.0f waiting for scores %.0f topk %.0f remat %.0f install %.0f | "
                        "per block: emit + queue + load + first schedule %.0f\n", 2 * p.grid, tot[7], tot[6], tot[4] / vars, tot[5] / vars, tot[2] / vars,
                tot[3] / vars, tot[1] / vars, tot[0] / (tot[7] > 0 ? tot[7] : 1));
    }
    return irec_check_launch("k_beam_encode_tmem");
}
