// irec_resident2.cuh -- K1a second generation: the persistent per-coder-block beam-search encoder.
//
// Reference loop being replaced: rec/coding/beam_search_coder.py:53-122 (one CTA = one coder-block, all
// of its auxiliary variables); results are bit-identical to k_beam_encode_resident and to the oracle.
//
// What changed against the first kernel (profiles/r1_resident_v1_ncu.md: shared-memory wavefronts were
// the top limiter -- 3.6-way conflicted quantile gathers + one beam float4 per 4 candidate-dims -- with
// three IMADs per candidate-dim competing for the FMA pipe):
//   * discrete-log addressing: 10007 is prime, g = 5 generates Z_10007^*, so with a = dlog(r) (per
//     candidate sample and dim, shared by all beams) and c_b = dlog(h_b) (per beam, CTA-uniform)
//         T[(r * h_b) mod 10007] = T2[a + c_b],   T2[e] = T[g^e]   (same float32 bit patterns)
//     the modular product of beam_search_coder.py:45-47 is ONE integer add per candidate-dim;
//   * every lane scores NS candidate samples of its 32-dim chunk at once, so a beam float4 read from
//     shared memory serves 4 * NS candidate-dims instead of 4.
//   * two-choice bank assignment: T2 is stored three times in shared memory, so the value of exponent a
//     lives at word offsets a and a + 10006 (banks differ by 22).  The per-launch exponent table picks,
//     for the 32 lanes of every gather instruction, the copy that keeps the bank multiplicity at <= 2
//     (random banks: 3.5 wavefronts per gather; two choices: ~2.2).  The choice is the same for every
//     beam because all lanes add the same c_b.  To make room for the 120 KB table the schedule inputs
//     (sigma_p^2, sigma_t^2, delta mu, cumulative variance) live in a per-CTA global scratch and the
//     discrete-log table is read from global memory (it is only needed for c_b and the no-table path).
// The float32 operation order of the canonical score (oracle/irec_oracle.c beam_score) is unchanged.
#pragma once
#include "irec_beam.cuh"

#define R2_THREADS 384
#define R2_TOPK_CAP 1024

// byte offset (into T2) of stream word u:  4 * dlog(1 + u mod 10006)   (beam_search_coder.py:39-43)
// DLS: `dl4` is a copy of the table in SHARED memory.  A warp's 32 lookups are random 2-byte reads of a 20 KB table: from
// global memory they touch ~27 different 128-byte lines (as many L1 wavefronts on the data pipe the quantile gathers need),
// from shared memory ~3.5 (random banks).  Used by the general path, whose exponents are computed in place.
template <bool DLS = false>
__device__ __forceinline__ uint32_t r2_exp4(const uint16_t* __restrict__ dl4, uint32_t u)
{
    if (DLS) {
        uint16_t v;
        asm("ld.shared.u16 %0, [%1];" : "=h"(v) : "r"((uint32_t)__cvta_generic_to_shared(dl4 + u % IREC_ORD)));
        return (uint32_t)v;
    }
    return (uint32_t)__ldg(dl4 + u % IREC_ORD);
}

// In-place version of the two-choice bank assignment (no precomputed table): the 32 lanes of a gather instruction find
// out with ONE match.any which of them share a bank (all beams add the same c_b, so the pattern holds for every beam of
// the instruction); the odd-ranked lanes of every group switch to the second copy of T2 (+10006 words = 22 banks on).
// Random banks cost 3.5 wavefronts per gather, this one-shot rule 2.7 (the offline greedy of k_r2_exps: 2.2).
// All 32 lanes of the warp must call it.  Same values either way (T2[a] == T2[a + 10006]).
__device__ __forceinline__ uint32_t r2_spread_banks(uint32_t ad)
{
    const uint32_t peers = __match_any_sync(0xffffffffu, (ad >> 2) & 31u);
    const uint32_t rank = __popc(peers & ((1u << (threadIdx.x & 31)) - 1u));
    return ad + ((rank & 1u) ? 4u * IREC_ORD : 0u);
}

// exponent-table entry: 4 x uint16 word offsets a' (a or a + 10006) -> byte offsets into T2
__device__ __forceinline__ void r2_unpack(const uint2 c, uint32_t& e0, uint32_t& e1, uint32_t& e2, uint32_t& e3)
{
    e0 = (c.x << 2) & 0x3fffcu; e1 = (c.x >> 14) & 0x3fffcu;
    e2 = (c.y << 2) & 0x3fffcu; e3 = (c.y >> 14) & 0x3fffcu;
}

template <int BMAX>
struct R2Group {   // beams per inner batch: G * 4 independent gathers in flight per candidate sample
    static constexpr int G = (BMAX <= 5) ? BMAX : ((BMAX % 5 == 0) ? 5 : 4);
};

// Chunk sums of NS candidate samples against all BMAX beam slots.
//   T2b   : s_T2 as bytes;  ad = 4 a;  cb4[b] = 4 c_b  (0 for unused slots; their sums are ignored)
//   j_base[k] = s_k * D + first dim of the chunk;  q0 = float4 index of the chunk's first quad
//   row[k]: sample s_k's row of the precomputed exponent table (4 x uint16 per quad, CI layout) or nullptr
//           (then the exponents come from Philox + dl4 in place)
// in-place exponents of quad iq for the NS samples of a lane (Philox -> uniform int -> discrete log [-> bank assignment])
template <int NS, bool SPREAD, bool DLS>
__device__ __forceinline__ void r2_exps_in_place(const uint16_t* __restrict__ dl4, const TfStream& st, const uint64_t (&j_base)[NS],
                                                 int iq, uint32_t (&ad)[NS][4])
{
#pragma unroll
    for (int k = 0; k < NS; ++k) {
        const uint64_t j = j_base[k] + 4 * iq;
        const uint4 u = ((j & 3) == 0) ? tf_stream_group(st, j >> 2) : tf_stream_quad_at(st, j);
        ad[k][0] = r2_exp4<DLS>(dl4, u.x); ad[k][1] = r2_exp4<DLS>(dl4, u.y);
        ad[k][2] = r2_exp4<DLS>(dl4, u.z); ad[k][3] = r2_exp4<DLS>(dl4, u.w);
        if (SPREAD) {
#pragma unroll
            for (int e = 0; e < 4; ++e) ad[k][e] = r2_spread_banks(ad[k][e]);
        }
    }
}

// SPREAD: in-place two-choice bank assignment (TAB == false only); PACK: FP32x2 arithmetic; DLS: dl4 in shared memory
template <int BMAX, int NS, bool TAB, bool SPREAD = false, bool PACK = false, bool DLS = false>
__device__ __forceinline__ void r2_score_chunk(const char* __restrict__ T2b, const uint16_t* __restrict__ dl4,
                                               const uint32_t* __restrict__ cb4,
                                               const float4* __restrict__ sa4, const float4* __restrict__ A4,
                                               const float4* __restrict__ E4, const float4* __restrict__ M4,
                                               const float4* __restrict__ beams4, int beam_stride4, int P, int q0,
                                               const TfStream& st, const uint64_t (&j_base)[NS],
                                               const uint2* __restrict__ tab_t, const uint32_t (&row)[NS],
                                               float (&acc)[NS][BMAX])
{
    constexpr int G = R2Group<BMAX>::G;
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
            for (int k = 0; k < NS; ++k) {
                r2_unpack(nxt[k], ad[k][0], ad[k][1], ad[k][2], ad[k][3]);
            }
            if (iq < 7) {
#pragma unroll
                for (int k = 0; k < NS; ++k) nxt[k] = __ldg(tab_t + row[k] + (iq + 1) * P);
            }
        } else {
            r2_exps_in_place<NS, SPREAD, DLS>(dl4, st, j_base, iq, ad);
        }
        const int qi = q0 + iq * P;
        const float4 sa = sa4[qi], A = A4[qi], E = E4[qi], M = M4[qi];
        // PACK: dims (0,1) and (2,3) of the quad as packed FP32x2 pairs (FMUL2 / FADD2 / FFMA2, irec_common.cuh): v = T*sa,
        // x = beam + v, d = x - M, t = fma(A, d, E) are four packed instructions per pair -- the same roundings as the scalar
        // sequence; the accumulation a = fma(t, d, a) stays scalar and sequential in d (the canonical order).  Used where it
        // measured faster (cluster kernel: 12.8 -> 11.5 ms for configs[2]); the general path got slower with it (166 registers)
        const f32x2_t sa2[2] = { f2_pack(sa.x, sa.y), f2_pack(sa.z, sa.w) };
        const f32x2_t A2[2] = { f2_pack(A.x, A.y), f2_pack(A.z, A.w) };
        const f32x2_t E2[2] = { f2_pack(E.x, E.y), f2_pack(E.z, E.w) };
        const f32x2_t nM2[2] = { f2_pack(-M.x, -M.y), f2_pack(-M.z, -M.w) };
#pragma unroll
        for (int b0 = 0; b0 < BMAX; b0 += G) {
            float4 bm[G];
            uint32_t cb[G];
#pragma unroll
            for (int g = 0; g < G; ++g) {
                bm[g] = beams4[(b0 + g) * beam_stride4 + qi];
                cb[g] = cb4[b0 + g];
            }
#pragma unroll
            for (int k = 0; k < NS; ++k) {
                float tv[G][4];
#pragma unroll
                for (int g = 0; g < G; ++g) {
#pragma unroll
                    for (int e = 0; e < 4; ++e)
                        tv[g][e] = *reinterpret_cast<const float*>(T2b + (ad[k][e] + cb[g]));
                }
#pragma unroll
                for (int g = 0; g < G; ++g) {
                    float a = acc[k][b0 + g];
                    if (PACK) {
                        const f32x2_t bm2[2] = { f2_pack(bm[g].x, bm[g].y), f2_pack(bm[g].z, bm[g].w) };
#pragma unroll
                        for (int h2 = 0; h2 < 2; ++h2) {
                            const f32x2_t x = f2_add(bm2[h2], f2_mul(f2_pack(tv[g][2 * h2], tv[g][2 * h2 + 1]), sa2[h2]));
                            const f32x2_t d = f2_add(x, nM2[h2]);
                            const f32x2_t t = f2_fma(A2[h2], d, E2[h2]);
                            float d0, d1, t0, t1;
                            f2_unpack(d, d0, d1);
                            f2_unpack(t, t0, t1);
                            a = __fmaf_rn(t0, d0, a);
                            a = __fmaf_rn(t1, d1, a);
                        }
                    } else {
                        float x, d, t;
                        x = __fadd_rn(bm[g].x, __fmul_rn(tv[g][0], sa.x));
                        d = __fadd_rn(x, -M.x); t = __fmaf_rn(A.x, d, E.x); a = __fmaf_rn(t, d, a);
                        x = __fadd_rn(bm[g].y, __fmul_rn(tv[g][1], sa.y));
                        d = __fadd_rn(x, -M.y); t = __fmaf_rn(A.y, d, E.y); a = __fmaf_rn(t, d, a);
                        x = __fadd_rn(bm[g].z, __fmul_rn(tv[g][2], sa.z));
                        d = __fadd_rn(x, -M.z); t = __fmaf_rn(A.z, d, E.z); a = __fmaf_rn(t, d, a);
                        x = __fadd_rn(bm[g].w, __fmul_rn(tv[g][3], sa.w));
                        d = __fadd_rn(x, -M.w); t = __fmaf_rn(A.w, d, E.w); a = __fmaf_rn(t, d, a);
                    }
                    acc[k][b0 + g] = a;
                }
            }
        }
    }
}

// Canonical pairwise tree over the P chunk sums of a sample group (same totals, bit for bit, as the xor butterfly of
// group_tree_sum: level `lev` adds the sums of lanes l and l ^ 2^lev, and a + b == b + a), done by recursive halving:
// instead of both partners computing every total, the lane with bit `lev` clear keeps the even slots and its partner
// the odd ones -- N/2 shuffles per level instead of N.  After log2(P) levels every lane owns ~N/P totals and stores
// them itself.  N0 = NS * BMAX accumulators, slot a = k * BMAX + b  (sample k of the round, beam b).
template <int N0, int LEV>
struct R2LevelSize { static constexpr int value = (R2LevelSize<N0, LEV - 1>::value + 1) / 2; };   // slots per lane after LEV halvings
template <int N0>
struct R2LevelSize<N0, 0> { static constexpr int value = N0; };

// slot a of a lane after LEV halving levels -> original slot (and whether another lane holds the same total)
template <int N0, int LEV>
__device__ __forceinline__ int r2_unwind(int a, int lane, bool& dup)
{
    if constexpr (LEV == 0) {
        return a;
    } else {
        constexpr int n_prev = R2LevelSize<N0, LEV - 1>::value;
        const int bit = (lane >> (LEV - 1)) & 1;
        const bool tail = (n_prev & 1) && a == n_prev / 2;                    // odd tail slot: replicated in both partners
        dup = dup || (tail && bit);
        return r2_unwind<N0, LEV - 1>(tail ? n_prev - 1 : 2 * a + bit, lane, dup);
    }
}

// Where a finished total goes.  sk = candidate sample, b = beam slot, dup = another lane holds the same total (the
// replicated odd tail of a halving level; only sinks with expensive stores need to look at it).
struct R2LocalSink {
    float* s_scores; int S, Bcur, boff;      // boff: first beam slot of the pass (beams are scored in passes of R2Pass::HB)
    __device__ __forceinline__ void operator()(int sk, int b, float x, bool) const
    {
        b += boff;
        if (b < Bcur && sk < S) s_scores[sk * Bcur + b] = (x == x) ? x : __int_as_float(0xff800000);
    }
};

// Beam slots per scoring pass.  One pass keeps NS * HB accumulators in registers; with all BMAX = 20 slots in one pass
// (60 accumulators + the gather operands) ptxas sits at the 168-register cap of a 384-thread CTA and, depending on
// unrelated code, serialises the quantile gathers (load -> immediate use).  Two passes of 10 re-read the per-dim
// coefficients and the exponents once more (+2.5 % shared-memory wavefronts) and leave ~40 registers for loads in flight.
template <int BMAX>
struct R2Pass {
    static constexpr int N = (BMAX > 10) ? 2 : 1;
    static constexpr int HB = BMAX / N;
    static_assert(HB * N == BMAX, "BMAX must split evenly");
};

template <int N0, int N, int BMAX, int LEVEL, class Sink>
__device__ __forceinline__ void r2_tree_store(float (&v)[N], int P, int lane, int s_first, int s_step, const Sink& sink)
{
    constexpr int stride = 1 << LEVEL;
    if constexpr (LEVEL < 5) {
        if (stride < P) {
            constexpr int M = (N + 1) / 2;
            float o[M];
            const bool hi = (lane & stride) != 0;
#pragma unroll
            for (int i = 0; i < N / 2; ++i) {
                const float send = hi ? v[2 * i] : v[2 * i + 1];
                const float keep = hi ? v[2 * i + 1] : v[2 * i];
                o[i] = __fadd_rn(keep, __shfl_xor_sync(0xffffffffu, send, stride));
            }
            if constexpr (N & 1) o[M - 1] = __fadd_rn(v[N - 1], __shfl_xor_sync(0xffffffffu, v[N - 1], stride));
            r2_tree_store<N0, M, BMAX, LEVEL + 1, Sink>(o, P, lane, s_first, s_step, sink);
            return;
        }
    }
    // LEVEL levels done: slot i of this lane is the total of the original slot found by unwinding the halvings
#pragma unroll
    for (int i = 0; i < N; ++i) {
        bool dup = false;
        const int a = r2_unwind<N0, LEVEL>(i, lane, dup);
        const int k = a / BMAX, b = a - k * BMAX;
        sink(s_first + k * s_step, b, v[i], dup);
    }
}

// one round of a warp: NS sample groups (group = the 32/P samples of one warp row), scores -> s_scores
// tab_t: exponent table of this block size and partition ([S][row_stride] uint2) or nullptr
template <int BMAX, int NS, bool TAB>
__device__ __forceinline__ void r2_score_round(const char* T2b, const uint16_t* dl4, const uint32_t* cb4,
                                               const float4* sa4, const float4* A4, const float4* E4, const float4* M4,
                                               const float4* beams4, const BeamGeom& g, int lane, const TfStream& st,
                                               const uint2* tab_t, int row_stride,
                                               int sg_first, int sg_stride, int S, int Bcur, float* s_scores)
{
    const int lg = lane & (g.P - 1);
    constexpr int HB = R2Pass<BMAX>::HB;
    uint64_t jb[NS];
    uint32_t row[NS];                       // uint2 index of (sample row, first quad of the chunk) in tab_t
#pragma unroll
    for (int k = 0; k < NS; ++k) {
        const int sk = (sg_first + k * sg_stride) * g.SPW + lane / g.P;
        const int sc = min(sk, S - 1);
        jb[k] = TAB ? 0ull : (uint64_t)sc * (uint64_t)g.D + (uint64_t)(32 * lg);
        row[k] = (uint32_t)(sc * row_stride + lg);
    }
#pragma unroll 1
    for (int boff = 0; boff < BMAX; boff += HB) {
        if (boff >= Bcur) break;            // unused slots (warp-uniform)
        float acc[NS][HB];
#pragma unroll
        for (int k = 0; k < NS; ++k)
#pragma unroll
            for (int b = 0; b < HB; ++b) acc[k][b] = 0.f;
        r2_score_chunk<HB, NS, TAB>(T2b, dl4, cb4 + boff, sa4, A4, E4, M4, beams4 + boff * (g.DP >> 2), g.DP >> 2, g.P, lg, st, jb,
                                    tab_t, row, acc);
        float v[NS * HB];
#pragma unroll
        for (int k = 0; k < NS; ++k)
#pragma unroll
            for (int b = 0; b < HB; ++b) v[k * HB + b] = acc[k][b];
        const R2LocalSink sink{ s_scores, S, Bcur, boff };
        r2_tree_store<NS * HB, NS * HB, HB, 0, R2LocalSink>(v, g.P, lane, (sg_first * g.SPW + lane / g.P), sg_stride * g.SPW, sink);
    }
}

// all candidates of one partition: S samples x Bcur beams
template <int BMAX, bool TAB>
__device__ __forceinline__ void r2_score_partition(const char* T2b, const uint16_t* dl4, const uint32_t* cb4,
                                                   const float4* sa4, const float4* A4, const float4* E4, const float4* M4,
                                                   const float4* beams4, const BeamGeom& g, const TfStream& st,
                                                   const uint2* tab_t, int row_stride, int S, int Bcur, float* s_scores)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    const int nsg = (S + g.SPW - 1) / g.SPW;
    int sg = 0;
    // rounds of 3, then 2, then 1 sample groups per warp (all warps take the same branch)
    while (nsg - sg > 2 * nwarps) {
        if (sg + warp < nsg)   // groups beyond nsg are clamped inside (scores not stored)
            r2_score_round<BMAX, 3, TAB>(T2b, dl4, cb4, sa4, A4, E4, M4, beams4, g, lane, st, tab_t, row_stride, sg + warp, nwarps, S, Bcur,
                                      s_scores);
        sg += 3 * nwarps;
    }
    if (nsg - sg > nwarps) {
        if (sg + warp < nsg)
            r2_score_round<BMAX, 2, TAB>(T2b, dl4, cb4, sa4, A4, E4, M4, beams4, g, lane, st, tab_t, row_stride, sg + warp, nwarps, S, Bcur,
                                      s_scores);
        sg += 2 * nwarps;
    } else if (nsg - sg > 0) {
        if (sg + warp < nsg)
            r2_score_round<BMAX, 1, TAB>(T2b, dl4, cb4, sa4, A4, E4, M4, beams4, g, lane, st, tab_t, row_stride, sg + warp, nwarps, S, Bcur,
                                      s_scores);
        sg += nwarps;
    }
}

// Exponent table: per distinct block size D (at most R2_MAX_SIZES per launch), per auxiliary variable t < max_aux,
// per candidate sample s: the byte offsets 4*dlog(r[s,d]) of all dims in CI layout, 4 x uint16 per quad.  r depends
// on (seed + t, s, d, D) only (beam_search_coder.py:38-43: the same seed for every coder-block, coder.py:448), so the
// Philox stream and its discrete logs are computed ONCE per launch instead of once per coder-block.
#define R2_MAX_SIZES 2
struct R2Plan {
    int32_t n_sizes;               // distinct block sizes found (may exceed R2_MAX_SIZES: the rest uses in-place Philox)
    int32_t D[R2_MAX_SIZES];
    int32_t use_private;           // set by k_r2_key_commit: this launch's table is the one in its own workspace, not the shared cache
};
// Device-side description of what a (library-owned, reused) exponent table currently holds.  The table depends on
// (seed, S, row stride, the block sizes, t) only -- a model codes all of its latent tensors with the same seed
// (resnet_vae.py:824), so after the first launch the later ones find their rows already there.  The host invalidates
// it when seed / S / stride / capacity change; the block sizes are known on the device only and are checked there.
struct R2TabKey {
    int32_t valid;
    int32_t D[R2_MAX_SIZES];
    int32_t filled;                // rows t < filled are present for both sizes
};

// table base of block size D for the main kernels (nullptr: this size has no table)
__device__ __forceinline__ const uint2* r2_tab_of_size(const R2Plan* plan, const uint2* tab, int tab_aux, const uint2* tab_priv,
                                                       int max_aux, int S, int row_stride, int D)
{
    if (!tab) return nullptr;
    const bool priv = tab_priv != nullptr && plan->use_private != 0;
    const uint2* base = priv ? tab_priv : tab;
    const int rows = priv ? max_aux : tab_aux;
    const uint2* r = nullptr;
#pragma unroll
    for (int k = 0; k < R2_MAX_SIZES; ++k)
        if (plan->D[k] == D) r = base + (size_t)k * rows * S * row_stride;
    return r;
}
#ifndef IREC_R2_DEVICE_ONLY      // the __global__ kernels below belong to irec_beam.cu only
// Also writes `order`: the coder-blocks sorted by decreasing size (counting sort), the sequence in which the
// persistent CTAs draw them from the queue -- the short blocks of a tensor (its last, partial block) go last and
// fill the tail of the launch instead of leaving SMs idle behind a full-size block.
__global__ void __launch_bounds__(1024) k_r2_plan(const int64_t* __restrict__ offs, int nb, R2Plan* plan, int32_t* __restrict__ order)
{
    // single CTA; plan was zeroed by the host (cudaMemsetAsync)
    __shared__ int s_cnt[1025];
    for (int i = threadIdx.x; i < 1025; i += blockDim.x) s_cnt[i] = 0;
    __syncthreads();
    for (int b = threadIdx.x; b < nb; b += blockDim.x) {
        const int64_t D64 = offs[b + 1] - offs[b];
        const int D = D64 < 0 ? 0 : (D64 > 1024 ? 1024 : (int)D64);
        atomicAdd(&s_cnt[1024 - D], 1);            // bucket 0 = largest
        if (D <= 0) continue;
        for (int k = 0; k < R2_MAX_SIZES; ++k) {
            const int old = atomicCAS(&plan->D[k], 0, D);
            if (old == 0 || old == D) break;
        }
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        int n = 0;
        for (int k = 0; k < R2_MAX_SIZES; ++k) n += plan->D[k] != 0;
        plan->n_sizes = n;
        for (int i = 1; i < R2_MAX_SIZES; ++i)             // canonical (descending) order: the CAS race above must not
            for (int j = i; j > 0 && plan->D[j] > plan->D[j - 1]; --j) {      // decide which table slot a size gets
                const int tmp = plan->D[j]; plan->D[j] = plan->D[j - 1]; plan->D[j - 1] = tmp;
            }
        int run = 0;
        for (int i = 0; i < 1025; ++i) { const int c = s_cnt[i]; s_cnt[i] = run; run += c; }   // exclusive prefix
    }
    __syncthreads();
    for (int b = threadIdx.x; b < nb; b += blockDim.x) {
        const int64_t D64 = offs[b + 1] - offs[b];
        const int D = D64 < 0 ? 0 : (D64 > 1024 ? 1024 : (int)D64);
        order[atomicAdd(&s_cnt[1024 - D], 1)] = b;
    }
}

// One thread per gather family (size k, auxiliary variable t, sample group sg, quad-in-chunk iqd): the 32 lanes
// (sample row r = lane / P, chunk l = lane % P) of the four gather instructions e = 0..3 of that quad.  For each
// instruction the lanes' word offsets a are assigned to copy 0 (a) or copy 1 (a + 10006) of T2 so that no bank
// serves more than `cap` lanes: greedy "less loaded of the two banks", one-hop relocation when both are full,
// cap raised from 2 only if that fails.  Any assignment yields the same values (T2[a] == T2[a + 10006]).
#define R2_BANK_SHIFT (IREC_ORD & 31u)
// What a launch does with the shared (library-owned) table, decided on the device from the key -- identically by
// k_r2_exps and k_r2_key_commit of the same launch (the key cannot change its block sizes in between: a re-key needs every
// other stream's launches to have finished, r2_tab_acquire):
//   R2_TAB_APPEND : the table holds these block sizes; rows >= key->filled are added (rows below are never rewritten, so
//                   launches of other streams may be reading them)
//   R2_TAB_REKEY  : other sizes, and the host has established that no other stream can be using the table: rebuilt in place
//   R2_TAB_PRIVATE: other sizes while another stream may be reading: the launch builds its own table in its workspace
enum { R2_TAB_APPEND = 0, R2_TAB_REKEY = 1, R2_TAB_PRIVATE = 2 };
__device__ __forceinline__ int r2_tab_mode(const R2Plan* plan, const R2TabKey* key, int allow_rekey)
{
    bool same = key->valid != 0;
#pragma unroll
    for (int k = 0; k < R2_MAX_SIZES; ++k) same = same && key->D[k] == plan->D[k];
    return same ? R2_TAB_APPEND : (allow_rekey ? R2_TAB_REKEY : R2_TAB_PRIVATE);
}
__global__ void k_r2_key_commit(R2Plan* __restrict__ plan, R2TabKey* key, int max_aux, int allow_rekey)
{
    const int mode = r2_tab_mode(plan, key, allow_rekey);
    if (mode == R2_TAB_APPEND) {
        atomicMax(&key->filled, max_aux);
    } else if (mode == R2_TAB_REKEY) {
        for (int k = 0; k < R2_MAX_SIZES; ++k) key->D[k] = plan->D[k];
        key->filled = max_aux;
        key->valid = 1;
    } else {
        plan->use_private = 1;
    }
}

// tab_aux: rows per size the table is laid out for (>= max_aux); key: nullptr, or the cache key -- rows it already
// covers for these block sizes are skipped (k_r2_key_commit, launched right after, records the new state)
__global__ void __launch_bounds__(128) k_r2_exps(const R2Plan* __restrict__ plan, const uint16_t* __restrict__ dl4,
                                                 int64_t seed, int S, int max_aux, int tab_aux, int row_stride,
                                                 uint2* __restrict__ tab, const R2TabKey* __restrict__ key,
                                                 uint2* __restrict__ tab_priv, int allow_rekey)
{
    int t_first = 0;
    if (key) {
        const int mode = r2_tab_mode(plan, key, allow_rekey);
        if (mode == R2_TAB_APPEND) {
            t_first = key->filled;                 // may grow under us (another stream appending): rows are then written twice, same values
            if (t_first >= max_aux) return;
        } else if (mode == R2_TAB_PRIVATE) {
            tab = tab_priv; tab_aux = max_aux;     // layout of the workspace table: [R2_MAX_SIZES][max_aux][S][row_stride]
        }
    }
    const int64_t per_size = (int64_t)max_aux * S * 8;
    const int64_t total = per_size * R2_MAX_SIZES;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int k = (int)(i / per_size);
        const int D = plan->D[k];
        if (D == 0) continue;
        int64_t r = i - (int64_t)k * per_size;
        const int iqd = (int)(r & 7); r >>= 3;
        const int sg = (int)(r % S);
        const int t = (int)(r / S);
        if (t < t_first) continue;
        const BeamGeom g = make_geom(D);
        if (sg * g.SPW >= S) continue;
        const TfStream st = tf_stream_seeded(seed + t, seed + t);
        uint16_t av[4][32];
        uint32_t live[4] = { 0u, 0u, 0u, 0u };     // lanes whose dim exists (the others read offset 0: one broadcast address)
        for (int lane = 0; lane < 32; ++lane) {
            const int row = lane / g.P, l = lane & (g.P - 1);
            const int s = sg * g.SPW + row, d0 = 32 * l + 4 * iqd;
            uint32_t e[4] = { 0u, 0u, 0u, 0u };
            if (s < S && d0 < D) {
                const uint4 u = tf_stream_quad_at(st, (uint64_t)s * (uint64_t)D + (uint64_t)d0);
                e[0] = __ldg(dl4 + u.x % IREC_ORD) >> 2; e[1] = __ldg(dl4 + u.y % IREC_ORD) >> 2;
                e[2] = __ldg(dl4 + u.z % IREC_ORD) >> 2; e[3] = __ldg(dl4 + u.w % IREC_ORD) >> 2;
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    if (d0 + q < D) live[q] |= 1u << lane;
                    else e[q] = 0u;
                }
            }
#pragma unroll
            for (int q = 0; q < 4; ++q) av[q][lane] = (uint16_t)e[q];
        }
        uint32_t pick[4];                          // bit lane: use copy 1
#pragma unroll 1
        for (int q = 0; q < 4; ++q) {
            uint32_t ch = 0u;
            for (int cap = 2; cap <= 32; ++cap) {
                unsigned char load[32];
                for (int b = 0; b < 32; ++b) load[b] = 0;
                ch = 0u;
                bool ok = true;
                for (int lane = 0; lane < 32 && ok; ++lane) {
                    if (!((live[q] >> lane) & 1u)) continue;
                    const uint32_t b0 = av[q][lane] & 31u, b1 = (b0 + R2_BANK_SHIFT) & 31u;
                    const uint32_t c = load[b1] < load[b0] ? 1u : 0u;
                    const uint32_t bank = c ? b1 : b0;
                    if (load[bank] < cap) { load[bank]++; ch |= c << lane; continue; }
                    // both banks full: shortest augmenting path (BFS over banks; an edge b -> b' is a placed lane that sits in
                    // b and may sit in b').  Exact: a cap is given up only if NO assignment with that cap exists (greedy with
                    // a one-hop relocation: 2.21 wavefronts per gather on average, this: 2.15; the table is built once per
                    // (seed, S, block sizes) and reused, so the search is free).
                    bool placed = false;
                    {
                        uint32_t visited = 0u;
                        signed char par_bank[32], par_lane[32];
                        unsigned char queue[32];
                        int qh = 0, qt = 0;
                        visited |= 1u << b0; par_bank[b0] = -1; par_lane[b0] = -1; queue[qt++] = (unsigned char)b0;
                        if (b1 != b0) { visited |= 1u << b1; par_bank[b1] = -1; par_lane[b1] = -1; queue[qt++] = (unsigned char)b1; }
                        int found = -1;
                        while (qh < qt && found < 0) {
                            const uint32_t bk = queue[qh++];
                            for (int j = 0; j < lane; ++j) {
                                if (!((live[q] >> j) & 1u)) continue;
                                const uint32_t j0 = av[q][j] & 31u, j1 = (j0 + R2_BANK_SHIFT) & 31u;
                                const uint32_t cj = (ch >> j) & 1u;
                                if ((cj ? j1 : j0) != bk) continue;
                                const uint32_t alt = cj ? j0 : j1;
                                if ((visited >> alt) & 1u) continue;
                                visited |= 1u << alt; par_bank[alt] = (signed char)bk; par_lane[alt] = (signed char)j;
                                if (load[alt] < cap) { found = (int)alt; break; }
                                queue[qt++] = (unsigned char)alt;
                            }
                        }
                        if (found >= 0) {
                            int f = found;
                            load[f]++;                          // the path shifts one lane per bank; only its end gains a lane
                            while (par_bank[f] >= 0) {
                                ch ^= 1u << par_lane[f];        // lane par_lane[f] moves from par_bank[f] to f
                                f = par_bank[f];
                            }
                            ch |= (f == (int)b1 && b1 != b0 ? 1u : 0u) << lane;     // the new lane takes the freed slot of the start bank
                            placed = true;
                        }
                    }
                    ok = placed;
                }
                if (ok) break;
            }
            pick[q] = ch;
        }
        uint2* out = tab + (size_t)k * tab_aux * S * row_stride + (size_t)t * S * row_stride;
        for (int lane = 0; lane < 32; ++lane) {
            const int row = lane / g.P, l = lane & (g.P - 1);
            const int s = sg * g.SPW + row;
            if (s >= S) continue;
            uint32_t w[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) w[q] = (uint32_t)av[q][lane] + (((pick[q] >> lane) & 1u) ? IREC_ORD : 0u);
            out[(size_t)s * row_stride + iqd * g.P + l] = make_uint2(w[0] | (w[1] << 16), w[2] | (w[3] << 16));
        }
    }
}

#endif  // IREC_R2_DEVICE_ONLY

// Winners' new beams, in place (beam_search_coder.py:92-93): beam_j <- beam_{b_j} + a(s_j, b_j), fused with the
// schedule of the NEXT auxiliary variable (beam_search_coder.py:64-77), which does not depend on the winners.
// One WARP per winner j (s_list[j] = s_j, s_list[32 + j] = b_j), lanes stride over the quad columns: the winner's
// exponent row is one coalesced 8-byte load per lane and quad, all issued up front; the parent quad is a conflict-free
// LDS.128.  Every new value is held in registers until a CTA barrier, then stored -- a parent may be another winner's
// destination.  The next schedule (float64 divisions, global scratch loads) is computed into registers between the
// winners' loads and the barrier, so its latency overlaps the table round trips, and is stored after the barrier
// (the winners still read this variable's sigma_aux before it).  hs_new: hash sums of the new beams -> the next
// table offsets c_b.  Kept out of line so that it does not disturb the register allocation of the scoring loop.
struct R2NextSched {
    int on;                        // 0: last auxiliary variable, nothing to prepare
    float ratio;                   // power-law ratio of the next auxiliary variable
    const float* g_cv; const float* g_tv; const float* g_dmu; float* g_cum;     // per-CTA global scratch
    int off_A, off_E, off_M;       // byte offsets of the shared arrays
    int off_hs_new;                // int32[32] hash sums of the new beams
};

template <int BMAX, int WPW, bool TAB>
__device__ __noinline__ void r2_rematerialise(int off_T2, const uint16_t* __restrict__ dl4, int off_cb, int off_list,
                                              int off_sa, int off_beams, int DP, int P,
                                              int D, int Kout, const TfStream st, const uint2* __restrict__ tab_t, int row_stride,
                                              const R2NextSched ns)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];      // byte offsets keep the accesses LDS/STS (not generic)
    const char* T2b = reinterpret_cast<const char*>(smem_raw + off_T2);
    uint32_t* s_cb = reinterpret_cast<uint32_t*>(smem_raw + off_cb);
    const int32_t* s_list = reinterpret_cast<const int32_t*>(smem_raw + off_list);
    const float4* sa4 = reinterpret_cast<const float4*>(smem_raw + off_sa);
    float4* beams4 = reinterpret_cast<float4*>(smem_raw + off_beams);
    const int tid = threadIdx.x, nt = blockDim.x;
    const int lane = tid & 31, warp = tid >> 5, nwarps = nt >> 5;
    const int nq = DP >> 2, lgP = __ffs(P) - 1;                     // P is a power of two; nq <= 256
    // 1. every global load up front: the winners' exponent rows and the schedule inputs (L2 round trips overlap)
    uint2 ex[WPW][8];
    int sjw[WPW], bjw[WPW];
#pragma unroll
    for (int w = 0; w < WPW; ++w) {
        const int j = warp + w * nwarps;
        const bool on = j < Kout;
        sjw[w] = on ? s_list[j] : 0;
        bjw[w] = on ? s_list[32 + j] : 0;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const int qq = lane + 32 * i;
            ex[w][i] = (TAB && on && qq < nq) ? __ldg(tab_t + (size_t)sjw[w] * row_stride + qq) : make_uint2(0u, 0u);
        }
    }
    float g_in[4][4];
    uint32_t cb_next = 0u, live = 0u;
    if (ns.on) {
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            const int i = tid + r * nt;
            const bool on = i < DP;
            g_in[r][0] = on ? ns.g_cv[i] : 0.f;
            g_in[r][1] = on ? ns.g_tv[i] : 0.f;
            g_in[r][2] = on ? ns.g_dmu[i] : 0.f;
            g_in[r][3] = on ? ns.g_cum[i] : 0.f;
        }
        if (tid < 32) {
            const int32_t* hs_new = reinterpret_cast<const int32_t*>(smem_raw + ns.off_hs_new);
            cb_next = tid < Kout ? (uint32_t)__ldg(dl4 + (hash_from_sum(hs_new[tid]) - 1)) : 0u;
        }
    }
    // 2. next schedule into registers: thread owns dims tid, tid + nt, ... (at most 4: DP <= 1024, nt >= 256)
    SchedOut so[4];
#pragma unroll
    for (int r = 0; r < 4; ++r) {
        so[r].sa = 0.f; so[r].A = 0.f; so[r].E = 0.f; so[r].M = 0.f; so[r].cum_next = 0.f;
        if (ns.on && g_in[r][0] != 0.f) {            // padding (cv == 0) stays zero: nothing to store
            so[r] = beam_sched_dim(g_in[r][0], g_in[r][1], g_in[r][2], g_in[r][3], ns.ratio);
            live |= 1u << r;
        }
    }
    // 3. the winners' new values
    float4 nv[WPW][8];
#pragma unroll
    for (int w = 0; w < WPW; ++w)
#pragma unroll
        for (int i = 0; i < 8; ++i) nv[w][i] = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int w = 0; w < WPW; ++w) {
        const int j = warp + w * nwarps;
        if (j < Kout) {
            const int sj = sjw[w], bj = bjw[w];
            const uint32_t cb = s_cb[bj];
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const int qq = lane + 32 * i;
                const int iqd = qq >> lgP, l = qq & (P - 1);        // physical quad -> first dim
                const int d0 = 32 * l + 4 * iqd;
                float4 o = make_float4(0.f, 0.f, 0.f, 0.f);
                if (qq < nq && d0 < D) {
                    uint32_t e0, e1, e2, e3;
                    if (TAB) {
                        r2_unpack(ex[w][i], e0, e1, e2, e3);
                    } else {
                        const uint4 u = tf_stream_quad_at(st, (uint64_t)sj * (uint64_t)D + (uint64_t)d0);
                        e0 = r2_exp4(dl4, u.x); e1 = r2_exp4(dl4, u.y); e2 = r2_exp4(dl4, u.z); e3 = r2_exp4(dl4, u.w);
                    }
                    const float4 sa = sa4[qq];
                    const float4 ob = beams4[bj * nq + qq];
                    o.x = __fadd_rn(ob.x, __fmul_rn(*reinterpret_cast<const float*>(T2b + e0 + cb), sa.x));
                    o.y = __fadd_rn(ob.y, __fmul_rn(*reinterpret_cast<const float*>(T2b + e1 + cb), sa.y));
                    o.z = __fadd_rn(ob.z, __fmul_rn(*reinterpret_cast<const float*>(T2b + e2 + cb), sa.z));
                    o.w = __fadd_rn(ob.w, __fmul_rn(*reinterpret_cast<const float*>(T2b + e3 + cb), sa.w));
                    // dims beyond D inside the last quad: sa = 0 and the parent is 0 there, so the padding stays zero
                }
                nv[w][i] = o;
            }
        }
    }
    __syncthreads();
#pragma unroll
    for (int w = 0; w < WPW; ++w) {
        const int j = warp + w * nwarps;
        if (j < Kout) {
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const int qq = lane + 32 * i;
                if (qq < nq) beams4[j * nq + qq] = nv[w][i];      // pure padding columns receive zeros (they were zero)
            }
        }
    }
    if (ns.on) {
        float* s_sa = reinterpret_cast<float*>(smem_raw + off_sa);
        float* s_A = reinterpret_cast<float*>(smem_raw + ns.off_A);
        float* s_E = reinterpret_cast<float*>(smem_raw + ns.off_E);
        float* s_M = reinterpret_cast<float*>(smem_raw + ns.off_M);
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            const int i = tid + r * nt;
            if ((live >> r) & 1u) {
                s_sa[i] = so[r].sa; s_A[i] = so[r].A; s_E[i] = so[r].E; s_M[i] = so[r].M; ns.g_cum[i] = so[r].cum_next;
            }
        }
        if (tid < 32) s_cb[tid] = cb_next;
    }
}

#ifndef IREC_R2_DEVICE_ONLY
struct Resident2Args {
    const float* t_loc; const float* t_scale; const float* p_loc; const float* p_scale;
    const int64_t* gidx; const int64_t* offs; int nb;
    float omega; int S; int B; int64_t seed;
    int32_t* out_indices; int max_aux; int32_t* out_n_aux; int32_t* out_status; float* out_sample;
    const float* T2; const uint16_t* dl4; const float* ratio_tab; int ratio_len;
    int2* hist;            // [gridDim.x][max_aux][BMAX]
    int* work_counter;     // dynamic block queue
    int DPmax;             // padded dims capacity of the shared arrays (multiple of 32)
    int NC;                // capacity of the score array (>= S * BMAX)
    float* sched;          // [gridDim.x][4][DPmax] per-CTA scratch: sigma_p^2, sigma_t^2, delta mu, cumulative variance
    const int32_t* order;  // queue position -> coder-block (largest blocks first), or nullptr
    const R2Plan* plan;    // distinct block sizes with an exponent table (nullptr: no table)
    const uint2* tab;      // [R2_MAX_SIZES][tab_aux][S][DPmax / 4]
    int tab_aux;           // rows per size in the table layout (>= max_aux)
    const uint2* tab_priv; // [R2_MAX_SIZES][max_aux][S][DPmax / 4] in the launch's workspace, used when plan->use_private (or nullptr)
};

template <int BMAX>
__host__ __device__ constexpr size_t r2_smem_bytes(int DPmax, int NC)
{
    return 32 * sizeof(double) + sizeof(float) * ((size_t)IREC_T2_LEN + (size_t)BMAX * DPmax + (size_t)4 * DPmax + NC + 512 + 32) +
           sizeof(int32_t) * (32 + R2_TOPK_CAP + 4 + 64 + 4 + 32) + 16;
}

template <int BMAX>
__global__ void __launch_bounds__(R2_THREADS, 1) k_beam_encode_resident2(const Resident2Args a)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int tid = threadIdx.x, nt = blockDim.x;
    const int DPm = a.DPmax;

    // ---- shared memory carve-up (every array 16-byte aligned) ----
    double* s_kl = reinterpret_cast<double*>(smem_raw);                  // [32]
    float* s_T2 = reinterpret_cast<float*>(s_kl + 32);                   // [IREC_T2_LEN]
    float* s_beams = s_T2 + IREC_T2_LEN;                                 // [BMAX][DPm]
    float* s_sa = s_beams + (size_t)BMAX * DPm;                          // [DPm] x 4
    float* s_A = s_sa + DPm; float* s_E = s_A + DPm; float* s_M = s_E + DPm;
    float* s_scores = s_M + DPm;                                         // [NC]
    float* g_cv = a.sched + (size_t)blockIdx.x * 4 * DPm;                // global scratch (this CTA only)
    float* g_tv = g_cv + DPm; float* g_dmu = g_tv + DPm; float* g_cum = g_dmu + DPm;
    float* s_gmax = s_scores + a.NC;                                     // [512]
    float* s_wsc = s_gmax + 512;                                         // [32] winners' scores
    int32_t* s_wid = reinterpret_cast<int32_t*>(s_wsc + 32);             // [32] winners' flat ids
    int32_t* s_list = s_wid + 32;                                        // [R2_TOPK_CAP]
    int32_t* s_ctl = s_list + R2_TOPK_CAP;                               // [4]
    int32_t* s_hsum = s_ctl + 4;                                         // [2][32]
    int32_t* s_misc = s_hsum + 64;                                       // [4]
    uint32_t* s_cb = reinterpret_cast<uint32_t*>(s_misc + 4);            // [32] 4 * dlog(h_b)

    {
        const float4* src = reinterpret_cast<const float4*>(a.T2);
        float4* dst = reinterpret_cast<float4*>(s_T2);
        for (int i = tid; i < IREC_T2_LEN / 4; i += nt) dst[i] = src[i];
    }
    const char* T2b = reinterpret_cast<const char*>(s_T2);
    int2* hist = a.hist + (size_t)blockIdx.x * a.max_aux * BMAX;

    for (;;) {
        __syncthreads();
        if (tid == 0) s_misc[0] = atomicAdd(a.work_counter, 1);
        __syncthreads();
        if (s_misc[0] >= a.nb) break;
        const int blk = a.order ? a.order[s_misc[0]] : s_misc[0];
        const int64_t off = a.offs[blk];
        const int D = (int)(a.offs[blk + 1] - off);
        const BeamGeom g = make_geom(D);
        const int row_stride = DPm >> 2;
        // exponent table of this block size (if it has one)
        const uint2* tab_blk = r2_tab_of_size(a.plan, a.tab, a.tab_aux, a.tab_priv, a.max_aux, a.S, row_stride, D);

        // ---- load + KL (coder.py:499-501) ----
        for (int i = tid; i < g.DP; i += nt) {
            g_cv[i] = 0.f; g_tv[i] = 0.f; g_dmu[i] = 0.f; g_cum[i] = 0.f;
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
                g_cv[ci] = __fmul_rn(ps, ps);
                g_tv[ci] = __fmul_rn(ts, ts);
                g_dmu[ci] = __fadd_rn(tl, -pl);
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

        // ---- schedule of the first auxiliary variable (beam_search_coder.py:64-77); the later ones are prepared inside
        //      r2_rematerialise while the winners of the previous variable are re-materialised ----
        {
            const float ratio = a.ratio_tab[n_aux - 1];
            for (int i = tid; i < g.DP; i += nt) {
                const float cv = g_cv[i];
                if (cv != 0.f) {                   // padding stays zero
                    const SchedOut o = beam_sched_dim(cv, g_tv[i], g_dmu[i], g_cum[i], ratio);
                    s_sa[i] = o.sa; s_A[i] = o.A; s_E[i] = o.E; s_M[i] = o.M; g_cum[i] = o.cum_next;
                }
            }
            if (tid < 32) s_cb[tid] = tid < 1 ? (uint32_t)__ldg(a.dl4 + (hash_from_sum(0) - 1)) : 0u;   // empty index row
            __syncthreads();
        }

        for (int t = 0; t < n_aux; ++t) {
            const int32_t* hs = s_hsum + 32 * hb;

            // ---- score all S * Bcur candidates (beam_search_coder.py:79-84,97-102) ----
            const TfStream st = tf_stream_seeded(a.seed + t, a.seed + t);
            const uint2* tab_t = tab_blk ? tab_blk + (size_t)t * a.S * row_stride : nullptr;
            if (tab_t) {
                if (Bcur == 1)
                    r2_score_partition<1, true>(T2b, a.dl4, s_cb, sa4, A4, E4, M4, beams4, g, st, tab_t, row_stride, a.S, 1, s_scores);
                else
                    r2_score_partition<BMAX, true>(T2b, a.dl4, s_cb, sa4, A4, E4, M4, beams4, g, st, tab_t, row_stride, a.S, Bcur, s_scores);
            } else {
                if (Bcur == 1)
                    r2_score_partition<1, false>(T2b, a.dl4, s_cb, sa4, A4, E4, M4, beams4, g, st, tab_t, row_stride, a.S, 1, s_scores);
                else
                    r2_score_partition<BMAX, false>(T2b, a.dl4, s_cb, sa4, A4, E4, M4, beams4, g, st, tab_t, row_stride, a.S, Bcur, s_scores);
            }
            __syncthreads();

            // ---- top-B (beam_search_coder.py:86-89,104-106) ----
            const int Kout = block_topk(s_scores, nullptr, a.S * Bcur, a.B, s_wsc, s_wid, s_gmax, s_list, R2_TOPK_CAP, s_ctl);

            // ---- history + hash sums of the new beams (beam_search_coder.py:92-95); (s_j, b_j) for the re-materialisation
            //      (s_list is free after block_topk) ----
            int32_t* hs_new = s_hsum + 32 * (hb ^ 1);
            if (tid < Kout) {
                const int f = s_wid[tid];
                const int sj = f / Bcur, bj = f - sj * Bcur;
                hist[(size_t)t * BMAX + tid] = make_int2(sj, bj);
                hs_new[tid] = hsum_extend(hs[bj], sj, t);
                s_list[tid] = sj; s_list[32 + tid] = bj;
            }
            __syncthreads();

            // ---- re-materialise the winners: beam_j <- beam_{b_j} + a(s_j, b_j)  (:92-93), and prepare the next variable ----
            {
                const unsigned char* b0 = smem_raw;
                R2NextSched ns;
                ns.on = (t + 1 < n_aux) ? 1 : 0;
                ns.ratio = ns.on ? a.ratio_tab[n_aux - 2 - t] : 0.f;
                ns.g_cv = g_cv; ns.g_tv = g_tv; ns.g_dmu = g_dmu; ns.g_cum = g_cum;
                ns.off_A = (int)(reinterpret_cast<const unsigned char*>(s_A) - b0);
                ns.off_E = (int)(reinterpret_cast<const unsigned char*>(s_E) - b0);
                ns.off_M = (int)(reinterpret_cast<const unsigned char*>(s_M) - b0);
                ns.off_hs_new = (int)(reinterpret_cast<const unsigned char*>(hs_new) - b0);
                const int o_T2 = (int)(reinterpret_cast<const unsigned char*>(s_T2) - b0);
                const int o_cb = (int)(reinterpret_cast<const unsigned char*>(s_cb) - b0);
                const int o_list = (int)(reinterpret_cast<const unsigned char*>(s_list) - b0);
                const int o_sa = (int)(reinterpret_cast<const unsigned char*>(s_sa) - b0);
                const int o_beams = (int)(reinterpret_cast<const unsigned char*>(s_beams) - b0);
                constexpr int W12 = (BMAX + 11) / 12, W8 = (BMAX + 7) / 8;      // winners per warp with 12 / at least 8 warps
                if (!tab_t)
                    r2_rematerialise<BMAX, W8, false>(o_T2, a.dl4, o_cb, o_list, o_sa, o_beams, g.DP, g.P, D, Kout, st, tab_t, row_stride, ns);
                else if (W12 < W8 && nt >= 384)
                    r2_rematerialise<BMAX, W12, true>(o_T2, a.dl4, o_cb, o_list, o_sa, o_beams, g.DP, g.P, D, Kout, st, tab_t, row_stride, ns);
                else
                    r2_rematerialise<BMAX, W8, true>(o_T2, a.dl4, o_cb, o_list, o_sa, o_beams, g.DP, g.P, D, Kout, st, tab_t, row_stride, ns);
            }
            __syncthreads();
            Bcur = Kout;
            hb ^= 1;
        }

        // ---- emit: indices of the best beam (trace the back-pointers) and its sample (:118-122) ----
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
#endif  // IREC_R2_DEVICE_ONLY
