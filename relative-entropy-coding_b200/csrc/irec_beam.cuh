// irec_beam.cuh -- device building blocks of the beam-search coder kernels.
//
//  * score_chunk     : the hot loop.  One lane scores one 32-dim chunk of one candidate sample
//                      against all current beams (Philox -> uniform int -> hash mix -> quantile LUT
//                      -> beam add -> centred quadratic log-ratio), accumulators in registers.
//  * group_tree_sum  : canonical pairwise tree over the chunk partial sums (warp shuffles).
//  * block_topk      : exact, deterministic top-K of n (score, flat index) pairs by
//                      (score desc, index asc), CTA-wide.
//
// Shared-memory layout "CI(P)" of every per-dim array (P = power of two >= #chunks, <= 32):
//   dim d = 32*l + i  ->  ((i>>2)*P + l)*4 + (i&3)        (length 32*P, zero padded)
// so lane l reads the float4 of dims 32l+4q..32l+4q+3 at float4 index q*P + l: consecutive lanes,
// consecutive 16-byte words, no bank conflicts.  For D > 1024 the array is a sequence of
// 1024-dim slots, each in CI(32).
#pragma once
#include "irec_common.cuh"

__host__ __device__ __forceinline__ int ci_index(int d, int P)
{
    const int slot = d / (32 * P);          // only > 0 when P == 32
    const int rem = d - slot * 32 * P;
    const int l = rem >> 5, i = rem & 31;
    return slot * 32 * P + (((i >> 2) * P + l) << 2) + (i & 3);
}

struct BeamGeom {
    int D;        // dims of the block
    int nch;      // ceil(D/32)
    int P;        // lanes per sample = min(32, next_pow2(nch))
    int nslots;   // ceil(nch/32)  (1 unless D > 1024)
    int DP;       // padded dims = 32*P*nslots
    int SPW;      // samples per warp = 32/P
};

__host__ __device__ __forceinline__ BeamGeom make_geom(int D)
{
    BeamGeom g;
    g.D = D;
    g.nch = (D + 31) >> 5;
    const int p2 = next_pow2_int(g.nch);
    g.P = p2 < 32 ? p2 : 32;
    g.nslots = (g.nch + 31) >> 5;
    g.DP = 32 * g.P * g.nslots;
    g.SPW = 32 / g.P;
    return g;
}

// ---------------------------------------------------------------------------------------------
// score_chunk: accumulate the canonical chunk sums of candidate sample `s` (stream element base
// j_base = s*D + first dim of the chunk) against ALL BMAX beam slots (slots >= Bcur carry hq = 0 and
// are ignored by the caller; there is no branch in the loop).
//   per dim:  k = (r*h_b) mod 10007; z = T[k]*sa; x = beam_b + z; d = x - m;
//             acc_b = fma(fma(A, d, E), d, acc_b)
// Exact modular product in three IMADs: with r' = floor(r * 2^32 / P) (once per element) the quotient
// q = umulhi(h, r') equals floor(r*h/P) exactly (P prime, so r*h/P is never within 2^-18 of an
// integer), hence 4k = r*(4h) - q*(4P) is the byte offset into the table.
// q0 = float4 index of the chunk's first quad inside the CI arrays (slot*8*P + l).
// ---------------------------------------------------------------------------------------------
struct BeamHash {
    uint32_t h4;   // 4 * h_b  (0 for unused slots)
    uint32_t h;    // h_b
};

__device__ __forceinline__ uint32_t shoup_rprime(uint32_t r)
{
    return (uint32_t)(((uint64_t)r << 32) / IREC_PRIME);
}

__device__ __forceinline__ float table_at(const float* __restrict__ T, uint32_t r, uint32_t rp, const BeamHash& bh)
{
    const uint32_t q = __umulhi(bh.h, rp);
    const uint32_t k4 = r * bh.h4 - q * (4u * IREC_PRIME);
    return *reinterpret_cast<const float*>(reinterpret_cast<const char*>(T) + k4);
}

template <int BMAX>
struct BeamGroup {   // beams are processed in groups of G so that G*4 independent chains are in flight
    static constexpr int G = (BMAX <= 5) ? BMAX : ((BMAX % 5 == 0) ? 5 : 4);
};

template <int BMAX>
__device__ __forceinline__ void score_chunk(const float* __restrict__ T,
                                            const float4* __restrict__ sa4, const float4* __restrict__ A4,
                                            const float4* __restrict__ E4, const float4* __restrict__ M4,
                                            const float4* __restrict__ beams4, int beam_stride4, int P, int q0,
                                            const TfStream& st, uint64_t j_base, const BeamHash (&h)[BMAX],
                                            float (&acc)[BMAX])
{
    constexpr int G = BeamGroup<BMAX>::G;
    const bool aligned = (j_base & 3) == 0;
#pragma unroll 1
    for (int iq = 0; iq < 8; ++iq) {
        const uint4 u = aligned ? tf_stream_group(st, (j_base >> 2) + iq) : tf_stream_quad_at(st, j_base + 4 * iq);
        const uint32_t r0 = beam_r_from_u32(u.x), r1 = beam_r_from_u32(u.y);
        const uint32_t r2 = beam_r_from_u32(u.z), r3 = beam_r_from_u32(u.w);
        const uint32_t p0 = shoup_rprime(r0), p1 = shoup_rprime(r1), p2 = shoup_rprime(r2), p3 = shoup_rprime(r3);
        const int qi = q0 + iq * P;
        const float4 sa = sa4[qi], A = A4[qi], E = E4[qi], M = M4[qi];
#pragma unroll
        for (int b0 = 0; b0 < BMAX; b0 += G) {
            float tv[G][4];
            float4 bm[G];
#pragma unroll
            for (int g = 0; g < G; ++g) {
                tv[g][0] = table_at(T, r0, p0, h[b0 + g]);
                tv[g][1] = table_at(T, r1, p1, h[b0 + g]);
                tv[g][2] = table_at(T, r2, p2, h[b0 + g]);
                tv[g][3] = table_at(T, r3, p3, h[b0 + g]);
                bm[g] = beams4[(b0 + g) * beam_stride4 + qi];
            }
#pragma unroll
            for (int g = 0; g < G; ++g) {
                float x, d, t, a = acc[b0 + g];
                x = __fadd_rn(bm[g].x, __fmul_rn(tv[g][0], sa.x));
                d = __fadd_rn(x, -M.x); t = __fmaf_rn(A.x, d, E.x); a = __fmaf_rn(t, d, a);
                x = __fadd_rn(bm[g].y, __fmul_rn(tv[g][1], sa.y));
                d = __fadd_rn(x, -M.y); t = __fmaf_rn(A.y, d, E.y); a = __fmaf_rn(t, d, a);
                x = __fadd_rn(bm[g].z, __fmul_rn(tv[g][2], sa.z));
                d = __fadd_rn(x, -M.z); t = __fmaf_rn(A.z, d, E.z); a = __fmaf_rn(t, d, a);
                x = __fadd_rn(bm[g].w, __fmul_rn(tv[g][3], sa.w));
                d = __fadd_rn(x, -M.w); t = __fmaf_rn(A.w, d, E.w); a = __fmaf_rn(t, d, a);
                acc[b0 + g] = a;
            }
        }
    }
}

// canonical pairwise tree over the P chunk sums held by the P lanes of a sample group
// (xor butterfly: stride 1, 2, 4, ...; a + b == b + a bitwise, so every lane ends with the total)
template <int BMAX>
__device__ __forceinline__ void group_tree_sum(float (&acc)[BMAX], int P)
{
    for (int stride = 1; stride < P; stride <<= 1) {
#pragma unroll
        for (int b = 0; b < BMAX; ++b) acc[b] = __fadd_rn(acc[b], __shfl_xor_sync(0xffffffffu, acc[b], stride));
    }
}

// Scores of sample s (for the P lanes of its group) against all beams; handles D > 1024 through
// a binary-counter pairwise combine of the 1024-dim slot totals (same tree as the oracle's).
template <int BMAX, bool MULTISLOT>
__device__ __forceinline__ void score_sample(const float* __restrict__ T, const float4* sa4, const float4* A4,
                                             const float4* E4, const float4* M4, const float4* beams4,
                                             const BeamGeom& g, int lg, const TfStream& st, uint64_t s,
                                             const BeamHash (&h)[BMAX], float (&acc)[BMAX])
{
    const int bstride4 = g.DP >> 2;
    if (!MULTISLOT) {
#pragma unroll
        for (int b = 0; b < BMAX; ++b) acc[b] = 0.f;
        // lanes with lg >= nch only see zero padding (A = E = sa = M = beam = 0): contributes +0
        score_chunk<BMAX>(T, sa4, A4, E4, M4, beams4, bstride4, g.P, lg, st,
                          s * (uint64_t)g.D + (uint64_t)(32 * lg), h, acc);
        group_tree_sum<BMAX>(acc, g.P);
    } else {
        float stack[12][BMAX];              // local memory; only for D > 1024 (up to 2^22 dims)
        for (int m = 0; m < g.nslots; ++m) {
#pragma unroll
            for (int b = 0; b < BMAX; ++b) acc[b] = 0.f;
            score_chunk<BMAX>(T, sa4, A4, E4, M4, beams4, bstride4, 32, m * 256 + lg, st,
                              s * (uint64_t)g.D + (uint64_t)(1024 * m + 32 * lg), h, acc);
            group_tree_sum<BMAX>(acc, 32);
            int lvl = 0;
            while ((m >> lvl) & 1) {
#pragma unroll
                for (int b = 0; b < BMAX; ++b) acc[b] = __fadd_rn(stack[lvl][b], acc[b]);
                ++lvl;
            }
#pragma unroll
            for (int b = 0; b < BMAX; ++b) stack[lvl][b] = acc[b];
        }
        // fold the remaining partial sub-trees (missing right halves are zeros: v + 0 = v)
        bool have = false;
        for (int lvl = 0; lvl < 12; ++lvl) {
            if ((g.nslots >> lvl) & 1) {
#pragma unroll
                for (int b = 0; b < BMAX; ++b) acc[b] = have ? __fadd_rn(stack[lvl][b], acc[b]) : stack[lvl][b];
                have = true;
            }
        }
    }
}

// =============================================================================================
// per-dim schedule of partition t (beam_search_coder.py:64-77,108-109; coder.py:141-154)
// float32 ops exactly in the reference's order; coefficients of the centred quadratic in float64.
// =============================================================================================
struct SchedOut {
    float sa, A, E, M, cum_next;
};
__device__ __forceinline__ SchedOut beam_sched_dim(float cv, float tv, float dmu, float cum, float ratio)
{
    SchedOut o;
    const float v = __fmul_rn(ratio, __fadd_rn(cv, -cum));
    const float tot = __fadd_rn(v, cum);
    const float m = __fdiv_rn(__fmul_rn(dmu, tot), cv);
    const float s2 = __fadd_rn(__fdiv_rn(__fmul_rn(tv, __fmul_rn(tot, tot)), __fmul_rn(cv, cv)),
                               __fdiv_rn(__fmul_rn(tot, __fadd_rn(cv, -tot)), cv));
    o.sa = __fsqrt_rn(v);
    o.A = (float)__dmul_rn(0.5, __dsub_rn(__ddiv_rn(1.0, (double)tot), __ddiv_rn(1.0, (double)s2)));
    o.E = (float)__ddiv_rn((double)m, (double)tot);
    o.M = m;
    o.cum_next = __fadd_rn(cum, v);
    return o;
}

// per-dim KL term in float64 (TFP kl_normal_normal)
__device__ __forceinline__ double kl_dim(float tl, float ts, float pl, float ps)
{
    const double sp = (double)ps;
    const double dl = __dsub_rn(log((double)ts), log(sp));
    const double dm = __dsub_rn(__ddiv_rn((double)tl, sp), __ddiv_rn((double)pl, sp));
    return __dsub_rn(__dadd_rn(__dmul_rn(0.5, __dmul_rn(dm, dm)), __dmul_rn(0.5, expm1(__dmul_rn(2.0, dl)))), dl);
}

__device__ __forceinline__ int n_aux_from_kl(float kl, float omega)
{
    const float q = __fdiv_rn(kl, omega);
    if (!(q == q) || isinf(q)) return -1;
    return (int)ceilf(q);
}

// Who takes part in a CTA-level helper: the whole CTA (default), or a group of whole warps that synchronises on its own
// named barrier (the two coder-block contexts of k_beam_encode_tmem, irec_tmem.cu).
struct CtaGroup {
    __device__ __forceinline__ int tid() const { return threadIdx.x; }
    __device__ __forceinline__ int nt() const { return blockDim.x; }
    __device__ __forceinline__ void sync() const { __syncthreads(); }
};
struct WarpGroup {
    int t, n, bar;                 // thread index inside the group, group size (multiple of 32), barrier id (1..15)
    __device__ __forceinline__ int tid() const { return t; }
    __device__ __forceinline__ int nt() const { return n; }
    __device__ __forceinline__ void sync() const { asm volatile("bar.sync %0, %1;" ::"r"(bar), "r"(n) : "memory"); }
};

// canonical tree over nch chunk sums in shared memory (all threads of the group call)
template <class Group>
__device__ __forceinline__ double block_tree_sum_f64(double* cs, int nch, const Group& grp)
{
    const int P = next_pow2_int(nch);
    grp.sync();
    for (int stride = 1; stride < P; stride <<= 1) {
        for (int i = grp.tid() * 2 * stride; i + stride < nch; i += grp.nt() * 2 * stride)
            cs[i] = __dadd_rn(cs[i], cs[i + stride]);
        grp.sync();
    }
    return cs[0];
}
__device__ __forceinline__ double block_tree_sum_f64(double* cs, int nch) { return block_tree_sum_f64(cs, nch, CtaGroup{}); }

// Slot in a shared-memory list for the lanes of a warp that have something to append: ONE atomicAdd per warp (same-address
// shared atomics serialise: the first scoring round of a variable appends every candidate, 7680 atomics -- 12 % of the warp
// samples of k_gp_fused in profiles/r2_gp_fused_b_ncu.md).  All 32 lanes must call it; returns -1 for lanes with pass == false.
__device__ __forceinline__ int warp_append_pos(int32_t* cnt, bool pass)
{
    const unsigned m = __ballot_sync(0xffffffffu, pass);
    if (m == 0u) return -1;
    const int lane = threadIdx.x & 31, leader = __ffs(m) - 1;
    int base = 0;
    if (lane == leader) base = atomicAdd(cnt, __popc(m));
    base = __shfl_sync(0xffffffffu, base, leader);
    return pass ? base + __popc(m & ((1u << lane) - 1u)) : -1;
}
// ---------------------------------------------------------------------------------------------
// block_topk: out[0..Kout) = the Kout = min(K, n) best of n candidates, best first, by
// (score desc, id asc).  sc/id may live in shared or global memory; id == nullptr means id = i.
// Scratch (shared): s_gmax [>= blockDim.x floats], s_list [cap ints], s_ctl [4 ints/floats].
// All threads of the CTA must call it.  NaN scores must have been mapped to -inf by the caller.
// ---------------------------------------------------------------------------------------------
template <class Group>
__device__ __forceinline__ int block_topk(const float* sc, const int32_t* id, int n, int K, float* out_sc,
                                          int32_t* out_id, float* s_gmax, int32_t* s_list, int cap, int32_t* s_ctl, const Group& grp)
{
    const int tid = grp.tid(), nt = grp.nt();
    const int Kout = K < n ? K : n;
    if (Kout <= 0) return 0;
    const float NEG_INF = __int_as_float(0xff800000);

    // 1. group maxima -> threshold tau = Kout-th largest group max (a lower bound of the answer's
    //    Kout-th score because distinct groups hold distinct candidates)
    int gsz = 32;
    while (gsz > 1 && nt / gsz < Kout) gsz >>= 1;
    const bool prune = (nt / gsz >= Kout) && (n > cap / 2 || n > 4 * Kout);
    if (tid == 0) { s_ctl[0] = 0; s_ctl[1] = __float_as_int(NEG_INF); }
    if (prune) {
        float lmax = NEG_INF;
        for (int i = tid; i < n; i += nt) lmax = fmaxf(lmax, sc[i]);
        for (int stride = gsz >> 1; stride >= 1; stride >>= 1)
            lmax = fmaxf(lmax, __shfl_xor_sync(0xffffffffu, lmax, stride));
        if ((tid & (gsz - 1)) == 0) s_gmax[tid / gsz] = lmax;
        grp.sync();
        const int G = nt / gsz;
        if (tid < G) {
            const float mine = s_gmax[tid];
            int cnt = 0;
            for (int j = 0; j < G; ++j) {
                const float o = s_gmax[j];
                cnt += (o > mine) || (o == mine && j < tid);
            }
            if (cnt == Kout - 1) s_ctl[1] = __float_as_int(mine);
        }
    }
    grp.sync();
    const float tau = __int_as_float(s_ctl[1]);

    // 2. survivors (one shared atomic per warp: whole warps stay in the loop)
    for (int i0 = 0; i0 < n; i0 += nt) {
        const int i = i0 + tid;
        const int pos = warp_append_pos(&s_ctl[0], i < n && sc[i] >= tau);
        if (pos >= 0 && pos < cap) s_list[pos] = i;
    }
    grp.sync();
    const int ns = s_ctl[0];

    if (ns <= cap) {
        // 3. exact rank among survivors
        for (int e = tid; e < ns; e += nt) {
            const int i = s_list[e];
            const float v = sc[i];
            const int32_t f = id ? id[i] : i;
            int rank = 0;
            for (int j = 0; j < ns; ++j) {
                const int i2 = s_list[j];
                const float v2 = sc[i2];
                const int32_t f2 = id ? id[i2] : i2;
                rank += (v2 > v) || (v2 == v && f2 < f);
            }
            if (rank < Kout) { out_sc[rank] = v; out_id[rank] = f; }
        }
        grp.sync();
    } else {
        // 4. fallback (massive ties): Kout rounds of CTA-wide arg-best with exclusion of the
        //    already selected (everything strictly better than the previous pick is selected)
        float prev_v = 0.f;
        int32_t prev_f = -1;
        for (int k = 0; k < Kout; ++k) {
            float bv = NEG_INF;
            int32_t bf = 0x7fffffff;
            for (int i = tid; i < n; i += nt) {
                const float v = sc[i];
                const int32_t f = id ? id[i] : i;
                const bool after_prev = (k == 0) || (v < prev_v) || (v == prev_v && f > prev_f);
                if (after_prev && ((v > bv) || (v == bv && f < bf))) { bv = v; bf = f; }
            }
            for (int stride = 16; stride >= 1; stride >>= 1) {
                const float ov = __shfl_xor_sync(0xffffffffu, bv, stride);
                const int32_t of = __shfl_xor_sync(0xffffffffu, bf, stride);
                if ((ov > bv) || (ov == bv && of < bf)) { bv = ov; bf = of; }
            }
            if ((tid & 31) == 0) { s_gmax[tid >> 5] = bv; s_list[tid >> 5] = bf; }
            grp.sync();
            if (tid == 0) {
                for (int w = 1; w < (nt >> 5); ++w) {
                    const float ov = s_gmax[w];
                    const int32_t of = s_list[w];
                    if ((ov > bv) || (ov == bv && of < bf)) { bv = ov; bf = of; }
                }
                out_sc[k] = bv; out_id[k] = bf;
            }
            grp.sync();
            prev_v = out_sc[k]; prev_f = out_id[k];
            grp.sync();
        }
    }
    return Kout;
}
__device__ __forceinline__ int block_topk(const float* sc, const int32_t* id, int n, int K, float* out_sc,
                                          int32_t* out_id, float* s_gmax, int32_t* s_list, int cap, int32_t* s_ctl)
{
    return block_topk(sc, id, n, K, out_sc, out_id, s_gmax, s_list, cap, s_ctl, CtaGroup{});
}
