// irec_cluster.cu -- K1c: the beam-search encoder for FEW coder-blocks per launch (single-image latency).
//
// Reference loop being replaced: rec/coding/beam_search_coder.py:53-122, called once per latent tensor by
// GaussianCoder.encode (rec/coding/coder.py:435-452).  One image of the reference's models hands the coder 9 (resnet_vae,
// [16,16,32] latents at block_size 1000) or 13 (level 2 of large_level_2_vae) coder-blocks at a time, and the levels are
// sequential (resnet_vae.py:821-826): the persistent one-CTA-per-block kernel (irec_resident2.cuh) would keep 9 of the
// 148 SMs busy.  Here one thread-block CLUSTER of G CTAs owns one coder-block:
//   * beam b lives in CTA b % G (slot b / G); a CTA scores all S candidate samples against ITS beams only, with the
//     same lane/chunk layout, exponent table and two-choice bank assignment as resident2 (the bank assignment does not
//     depend on the beam), and the same canonical float32 operation order;
//   * the scores are all-gathered through distributed shared memory: every finished total is stored straight from the
//     reduction tree into the score array of all G CTAs (st.shared::cluster), double buffered, ONE cluster barrier per
//     auxiliary variable;
//   * every CTA runs the same deterministic top-B on the full score array (identical winners everywhere, no second
//     exchange), keeps the hash sums of all beams, and re-materialises the winners it owns: the parent beam is read
//     from its owner's shared memory (ld.shared::cluster), the new beam is written locally into the other half of a
//     double-buffered beam store -- no barrier between the read and the write;
//   * KL, n_aux and the per-dim schedule are computed redundantly by every CTA (bit-identical, O(D) per variable).
// Results are bit-identical to k_beam_encode_resident2 and to the oracle (tests/test_gpu_parity.py).
#define IREC_R2_DEVICE_ONLY
#include "irec_resident2.cuh"
#include "irec_host.h"

#define RC_THREADS 384

__device__ __forceinline__ uint32_t rc_cta_rank()
{
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ uint32_t rc_cluster_id()
{
    uint32_t r;
    asm volatile("mov.u32 %0, %%clusterid.x;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void rc_cluster_sync()
{
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// shared::cluster address of the same shared-memory location in CTA `rank` of this cluster
__device__ __forceinline__ uint32_t rc_map(const void* local_smem, uint32_t rank)
{
    const uint32_t a = (uint32_t)__cvta_generic_to_shared(local_smem);
    uint32_t o;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(o) : "r"(a), "r"(rank));
    return o;
}
__device__ __forceinline__ void rc_st_f32(uint32_t addr, float v)
{
    asm volatile("st.shared::cluster.f32 [%0], %1;" ::"r"(addr), "f"(v) : "memory");
}
__device__ __forceinline__ float4 rc_ld_f4(uint32_t addr)
{
    float4 v;
    asm volatile("ld.shared::cluster.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr) : "memory");
    return v;
}

// totals of this CTA's candidates -> the score array (current half) of every CTA of the cluster
struct RcClusterSink {
    uint32_t scores_addr;      // shared::cta address of the current score buffer (same offset in every CTA)
    int S, Bcur, G, rank, nloc;
    __device__ __forceinline__ void operator()(int sk, int slot, float x, bool dup) const
    {
        if (dup || slot >= nloc || sk >= S) return;
        const float y = (x == x) ? x : __int_as_float(0xff800000);
        const uint32_t a = scores_addr + 4u * (uint32_t)(sk * Bcur + slot * G + rank);
        for (int rk = 0; rk < G; ++rk) {
            uint32_t o;
            asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(o) : "r"(a), "r"(rk));
            rc_st_f32(o, y);
        }
    }
};

template <int BSL, int NS, bool TAB>
__device__ __forceinline__ void rc_score_round(const char* T2b, const uint16_t* dl4, const uint32_t* cb4,
                                               const float4* sa4, const float4* A4, const float4* E4, const float4* M4,
                                               const float4* beams4, const BeamGeom& g, int lane, const TfStream& st,
                                               const uint2* tab_t, int row_stride, int sg_first, int sg_stride, int S,
                                               const RcClusterSink& sink)
{
    const int lg = lane & (g.P - 1);
    uint64_t jb[NS];
    uint32_t row[NS];
    float acc[NS][BSL];
#pragma unroll
    for (int k = 0; k < NS; ++k) {
        const int s = (sg_first + k * sg_stride) * g.SPW + lane / g.P;
        const int sc = min(s, S - 1);
        jb[k] = TAB ? 0ull : (uint64_t)sc * (uint64_t)g.D + (uint64_t)(32 * lg);
        row[k] = (uint32_t)(sc * row_stride + lg);
#pragma unroll
        for (int b = 0; b < BSL; ++b) acc[k][b] = 0.f;
    }
    r2_score_chunk<BSL, NS, TAB, false, true>(T2b, dl4, cb4, sa4, A4, E4, M4, beams4, g.DP >> 2, g.P, lg, st, jb, tab_t, row, acc);
    float v[NS * BSL];
#pragma unroll
    for (int k = 0; k < NS; ++k)
#pragma unroll
        for (int b = 0; b < BSL; ++b) v[k * BSL + b] = acc[k][b];
    r2_tree_store<NS * BSL, NS * BSL, BSL, 0, RcClusterSink>(v, g.P, lane, (sg_first * g.SPW + lane / g.P), sg_stride * g.SPW, sink);
}

// all S candidate samples against the BSL local beam slots
template <int BSL, bool TAB>
__device__ __forceinline__ void rc_score_partition(const char* T2b, const uint16_t* dl4, const uint32_t* cb4,
                                                   const float4* sa4, const float4* A4, const float4* E4, const float4* M4,
                                                   const float4* beams4, const BeamGeom& g, const TfStream& st,
                                                   const uint2* tab_t, int row_stride, int S, const RcClusterSink& sink)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    const int nsg = (S + g.SPW - 1) / g.SPW;
    int sg = 0;
    while (nsg - sg > 2 * nwarps) {
        if (sg + warp < nsg)
            rc_score_round<BSL, 3, TAB>(T2b, dl4, cb4, sa4, A4, E4, M4, beams4, g, lane, st, tab_t, row_stride, sg + warp, nwarps, S, sink);
        sg += 3 * nwarps;
    }
    if (nsg - sg > nwarps) {
        if (sg + warp < nsg)
            rc_score_round<BSL, 2, TAB>(T2b, dl4, cb4, sa4, A4, E4, M4, beams4, g, lane, st, tab_t, row_stride, sg + warp, nwarps, S, sink);
    } else if (nsg - sg > 0) {
        if (sg + warp < nsg)
            rc_score_round<BSL, 1, TAB>(T2b, dl4, cb4, sa4, A4, E4, M4, beams4, g, lane, st, tab_t, row_stride, sg + warp, nwarps, S, sink);
    }
}

struct ClusterArgs {
    const float* t_loc; const float* t_scale; const float* p_loc; const float* p_scale;
    const int64_t* gidx; const int64_t* offs; int nb;
    float omega; int S; int B; int64_t seed;
    int32_t* out_indices; int max_aux; int32_t* out_n_aux; int32_t* out_status; float* out_sample;
    const float* T2; const uint16_t* dl4; const float* ratio_tab; int ratio_len;
    int2* hist;            // [nb][max_aux][32]
    int DPmax;             // padded dims capacity of the shared arrays (multiple of 32)
    int NC;                // capacity of one score buffer (>= S * B)
    const int32_t* order;  // cluster -> coder-block (largest blocks first), or nullptr
    const R2Plan* plan;    // distinct block sizes with an exponent table (nullptr: no table)
    const uint2* tab;      // [R2_MAX_SIZES][tab_aux][S][DPmax / 4]
    int tab_aux;
    const uint2* tab_priv; // the launch's own table (workspace), used when plan->use_private
    int G;                 // CTAs per cluster
};

template <int BPC>
__host__ __device__ constexpr size_t rc_smem_bytes(int DPmax, int NC)
{
    return 32 * sizeof(double) +
           sizeof(float) * ((size_t)IREC_T2_LEN + (size_t)2 * BPC * DPmax + (size_t)9 * DPmax + (size_t)2 * NC + 512 + 32) +
           sizeof(int32_t) * (32 + R2_TOPK_CAP + 4 + 64 + 4 + 32 + 8) + 16;
}

// Winners owned by this CTA (j % G == rank): beam_j <- beam_{b_j} + a(s_j, b_j)  (beam_search_coder.py:92-93).
// The parent is read from its owner's current beam buffer through distributed shared memory, the new beam goes to the
// other buffer of this CTA.  s_list[j] = s_j, s_list[32 + j] = b_j.
template <int BPC>
__device__ __noinline__ void rc_rematerialise(int off_T2, const uint16_t* __restrict__ dl4, int off_cb, int off_list,
                                              int off_sa, int off_cur, int off_nxt, int DP, int P, int D, int Kout,
                                              int G, int rank, const TfStream st, const uint2* __restrict__ tab_t, int row_stride)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const char* T2b = reinterpret_cast<const char*>(smem_raw + off_T2);
    const uint32_t* s_cb = reinterpret_cast<const uint32_t*>(smem_raw + off_cb);
    const int32_t* s_list = reinterpret_cast<const int32_t*>(smem_raw + off_list);
    const float4* sa4 = reinterpret_cast<const float4*>(smem_raw + off_sa);
    float4* nxt4 = reinterpret_cast<float4*>(smem_raw + off_nxt);
    const int nq = DP >> 2, lgP = __ffs(P) - 1, lgnq = lgP + 3;
    for (int idx = threadIdx.x; idx < BPC * nq; idx += blockDim.x) {
        const int slot = idx >> lgnq, qq = idx & (nq - 1);
        const int j = slot * G + rank;
        if (j >= Kout) continue;
        const int sj = s_list[j], bj = s_list[32 + j];
        const int iqd = qq >> lgP, l = qq & (P - 1);
        const int d0 = 32 * l + 4 * iqd;
        float4 o = make_float4(0.f, 0.f, 0.f, 0.f);
        if (d0 < D) {
            uint32_t e0, e1, e2, e3;
            if (tab_t) {
                r2_unpack(__ldg(tab_t + (size_t)sj * row_stride + qq), e0, e1, e2, e3);
            } else {
                const uint4 u = tf_stream_quad_at(st, (uint64_t)sj * (uint64_t)D + (uint64_t)d0);
                e0 = r2_exp4(dl4, u.x); e1 = r2_exp4(dl4, u.y); e2 = r2_exp4(dl4, u.z); e3 = r2_exp4(dl4, u.w);
            }
            const uint32_t cb = s_cb[bj];
            const float4 sa = sa4[qq];
            const int pslot = bj / G, prank = bj - pslot * G;
            const float4 ob = rc_ld_f4(rc_map(smem_raw + off_cur + ((size_t)pslot * nq + qq) * sizeof(float4), (uint32_t)prank));
            o.x = __fadd_rn(ob.x, __fmul_rn(*reinterpret_cast<const float*>(T2b + e0 + cb), sa.x));
            o.y = __fadd_rn(ob.y, __fmul_rn(*reinterpret_cast<const float*>(T2b + e1 + cb), sa.y));
            o.z = __fadd_rn(ob.z, __fmul_rn(*reinterpret_cast<const float*>(T2b + e2 + cb), sa.z));
            o.w = __fadd_rn(ob.w, __fmul_rn(*reinterpret_cast<const float*>(T2b + e3 + cb), sa.w));
            // dims beyond D inside the last quad: sa = 0 and the parent is 0 there, so the padding stays zero
        }
        nxt4[slot * nq + qq] = o;
    }
}

template <int BPC>
__global__ void __launch_bounds__(RC_THREADS, 1) k_beam_encode_cluster(const ClusterArgs a)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int tid = threadIdx.x, nt = blockDim.x;
    const int DPm = a.DPmax;
    const int G = a.G, rank = (int)rc_cta_rank();

    // ---- shared memory carve-up (every array 16-byte aligned) ----
    double* s_kl = reinterpret_cast<double*>(smem_raw);                  // [32]
    float* s_T2 = reinterpret_cast<float*>(s_kl + 32);                   // [IREC_T2_LEN]
    float* s_beams = s_T2 + IREC_T2_LEN;                                 // [2][BPC][DPm]
    float* s_sa = s_beams + (size_t)2 * BPC * DPm;                       // [DPm] x 4: schedule of the current variable
    float* s_A = s_sa + DPm; float* s_E = s_A + DPm; float* s_M = s_E + DPm;
    float* s_cv = s_M + DPm;                                             // [DPm] x 4: sigma_p^2, sigma_t^2, delta mu, cumulative variance
    float* s_tv = s_cv + DPm; float* s_dmu = s_tv + DPm; float* s_cum = s_dmu + DPm;
    float* s_ploc = s_cum + DPm;                                         // [DPm] (rank 0: prior means for the emit)
    float* s_scores = s_ploc + DPm;                                      // [2][NC]
    float* s_gmax = s_scores + (size_t)2 * a.NC;                         // [512]
    float* s_wsc = s_gmax + 512;                                         // [32] winners' scores
    int32_t* s_wid = reinterpret_cast<int32_t*>(s_wsc + 32);             // [32] winners' flat ids
    int32_t* s_list = s_wid + 32;                                        // [R2_TOPK_CAP]
    int32_t* s_ctl = s_list + R2_TOPK_CAP;                               // [4]
    int32_t* s_hsum = s_ctl + 4;                                         // [2][32]
    int32_t* s_misc = s_hsum + 64;                                       // [4]
    uint32_t* s_cb = reinterpret_cast<uint32_t*>(s_misc + 4);            // [32] 4 * dlog(h_b), all beams
    uint32_t* s_cbl = s_cb + 32;                                         // [8]  the same for this CTA's slots (0 = unused)

    {
        const float4* src = reinterpret_cast<const float4*>(a.T2);
        float4* dst = reinterpret_cast<float4*>(s_T2);
        for (int i = tid; i < IREC_T2_LEN / 4; i += nt) dst[i] = src[i];
    }
    const char* T2b = reinterpret_cast<const char*>(s_T2);
    const int cid = (int)rc_cluster_id();
    const int blk = a.order ? a.order[cid] : cid;
    int2* hist = a.hist + (size_t)blk * a.max_aux * 32;
    const int64_t off = a.offs[blk];
    const int D = (int)(a.offs[blk + 1] - off);
    const BeamGeom g = make_geom(D);
    const int row_stride = DPm >> 2;
    const uint2* tab_blk = r2_tab_of_size(a.plan, a.tab, a.tab_aux, a.tab_priv, a.max_aux, a.S, row_stride, D);

    // ---- load + KL (coder.py:499-501), redundantly in every CTA ----
    for (int i = tid; i < g.DP; i += nt) {
        s_cv[i] = 0.f; s_tv[i] = 0.f; s_dmu[i] = 0.f; s_cum[i] = 0.f;
        s_sa[i] = 0.f; s_A[i] = 0.f; s_E[i] = 0.f; s_M[i] = 0.f;
    }
    for (int i = tid; i < 2 * BPC * g.DP; i += nt) s_beams[i] = 0.f;     // rows use stride g.DP
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
            s_ploc[ci] = pl;
        }
        s_kl[c] = acc;
    }
    const double kld = block_tree_sum_f64(s_kl, g.nch);
    const int n_aux = n_aux_from_kl((float)kld, a.omega);
    int status = IREC_BLK_OK;
    if (n_aux <= 0) status = IREC_BLK_BAD_KL;
    else if (n_aux > a.max_aux || n_aux > a.ratio_len) status = IREC_BLK_TOO_LONG;
    if (tid == 0 && rank == 0) { a.out_n_aux[blk] = n_aux; a.out_status[blk] = status; }
    if (tid < 64) s_hsum[tid] = 0;
    rc_cluster_sync();                             // every CTA of the cluster has initialised its shared memory
    if (status != IREC_BLK_OK) return;             // the same decision in every CTA of the cluster

    int Bcur = 1, hb = 0, cur = 0;                 // hb: current half of s_hsum; cur: current beam buffer
    const float4* sa4 = reinterpret_cast<const float4*>(s_sa);
    const float4* A4 = reinterpret_cast<const float4*>(s_A);
    const float4* E4 = reinterpret_cast<const float4*>(s_E);
    const float4* M4 = reinterpret_cast<const float4*>(s_M);

    for (int t = 0; t < n_aux; ++t) {
        // ---- schedule (beam_search_coder.py:64-77) + the beams' table offsets c_b = dlog(simple_hash) ----
        const float ratio = a.ratio_tab[n_aux - 1 - t];
        for (int i = tid; i < g.DP; i += nt) {
            const float cv = s_cv[i];
            if (cv != 0.f) {                       // padding stays zero
                const SchedOut o = beam_sched_dim(cv, s_tv[i], s_dmu[i], s_cum[i], ratio);
                s_sa[i] = o.sa; s_A[i] = o.A; s_E[i] = o.E; s_M[i] = o.M; s_cum[i] = o.cum_next;
            }
        }
        const int32_t* hs = s_hsum + 32 * hb;
        if (tid < 32) s_cb[tid] = tid < Bcur ? (uint32_t)__ldg(a.dl4 + (hash_from_sum(hs[tid]) - 1)) : 0u;
        if (tid >= 32 && tid < 40) {
            const int b = (tid - 32) * G + rank;
            s_cbl[tid - 32] = ((tid - 32) < BPC && b < Bcur) ? (uint32_t)__ldg(a.dl4 + (hash_from_sum(hs[b]) - 1)) : 0u;
        }
        __syncthreads();

        // ---- score the S candidates of this CTA's beams (beam_search_coder.py:79-84,97-102) -> all CTAs ----
        const TfStream st = tf_stream_seeded(a.seed + t, a.seed + t);
        const uint2* tab_t = tab_blk ? tab_blk + (size_t)t * a.S * row_stride : nullptr;
        const int nloc = rank < Bcur ? (Bcur - rank + G - 1) / G : 0;        // beams b < Bcur with b % G == rank
        float* sc_buf = s_scores + (size_t)(t & 1) * a.NC;
        const float4* beams4 = reinterpret_cast<const float4*>(s_beams + (size_t)cur * BPC * g.DP);
        if (nloc > 0) {
            const RcClusterSink sink{ (uint32_t)__cvta_generic_to_shared(sc_buf), a.S, Bcur, G, rank, nloc };
            if (tab_t) {
                if (nloc == 1)
                    rc_score_partition<1, true>(T2b, a.dl4, s_cbl, sa4, A4, E4, M4, beams4, g, st, tab_t, row_stride, a.S, sink);
                else
                    rc_score_partition<BPC, true>(T2b, a.dl4, s_cbl, sa4, A4, E4, M4, beams4, g, st, tab_t, row_stride, a.S, sink);
            } else {
                if (nloc == 1)
                    rc_score_partition<1, false>(T2b, a.dl4, s_cbl, sa4, A4, E4, M4, beams4, g, st, tab_t, row_stride, a.S, sink);
                else
                    rc_score_partition<BPC, false>(T2b, a.dl4, s_cbl, sa4, A4, E4, M4, beams4, g, st, tab_t, row_stride, a.S, sink);
            }
        }
        rc_cluster_sync();                         // all S * Bcur scores have arrived in this CTA's buffer

        // ---- top-B (beam_search_coder.py:86-89,104-106): the same winners in every CTA ----
        const int Kout = block_topk(sc_buf, nullptr, a.S * Bcur, a.B, s_wsc, s_wid, s_gmax, s_list, R2_TOPK_CAP, s_ctl);

        // ---- history + hash sums of the new beams (beam_search_coder.py:92-95) ----
        int32_t* hs_new = s_hsum + 32 * (hb ^ 1);
        if (tid < Kout) {
            const int f = s_wid[tid];
            const int sj = f / Bcur, bj = f - sj * Bcur;
            if (rank == 0) hist[(size_t)t * 32 + tid] = make_int2(sj, bj);
            hs_new[tid] = hsum_extend(hs[bj], sj, t);
            s_list[tid] = sj; s_list[32 + tid] = bj;                          // s_list is free after block_topk
        }
        __syncthreads();

        // ---- re-materialise the winners this CTA owns (:92-93) into the other beam buffer ----
        {
            const unsigned char* b0 = smem_raw;
            const float* curp = s_beams + (size_t)cur * BPC * g.DP;
            const float* nxtp = s_beams + (size_t)(cur ^ 1) * BPC * g.DP;
            rc_rematerialise<BPC>((int)(reinterpret_cast<const unsigned char*>(s_T2) - b0), a.dl4,
                                  (int)(reinterpret_cast<const unsigned char*>(s_cb) - b0),
                                  (int)(reinterpret_cast<const unsigned char*>(s_list) - b0),
                                  (int)(reinterpret_cast<const unsigned char*>(s_sa) - b0),
                                  (int)(reinterpret_cast<const unsigned char*>(curp) - b0),
                                  (int)(reinterpret_cast<const unsigned char*>(nxtp) - b0),
                                  g.DP, g.P, D, Kout, G, rank, st, tab_t, row_stride);
        }
        __syncthreads();
        Bcur = Kout;
        hb ^= 1;
        cur ^= 1;
    }

    // nobody leaves while its beams may still be read by the others
    rc_cluster_sync();

    // ---- emit (rank 0 owns beam 0): indices of the best beam (back-pointers) and its sample (:118-122) ----
    if (rank != 0) return;
    if (tid == 0) {
        int j = 0;
        int32_t* oi = a.out_indices + (size_t)blk * a.max_aux;
        for (int t = n_aux - 1; t >= 0; --t) {
            const int2 e = hist[(size_t)t * 32 + j];
            oi[t] = e.x;
            j = e.y;
        }
    }
    const float* best = s_beams + (size_t)cur * BPC * g.DP;
    for (int d = tid; d < D; d += nt) {
        const int64_t gi = a.gidx ? a.gidx[off + d] : off + d;
        const int ci = ci_index(d, g.P);
        a.out_sample[gi] = __fadd_rn(best[ci], s_ploc[ci]);
    }
}

// =============================================================================================
// host side
// =============================================================================================
size_t irec_cluster_hist_bytes(int nb, int max_aux)
{
    return (sizeof(int2) * (size_t)nb * (size_t)max_aux * 32 + 255) / 256 * 256;
}

// cluster size for this launch: 0 = do not use the cluster kernel
int irec_cluster_choice(int nb, int max_D, int S, int B)
{
    const char* e = getenv("IREC_CLUSTER");        // 0 = never, 4/8 = force that cluster size (tests / A-B runs)
    int G = -1;
    if (e && e[0]) G = atoi(e);
    if (G == 0) return 0;
    if (max_D > 1024 || B > 32 || (int64_t)S * B > 16384) return 0;
    const int sms = irec_device().sm_count;
    if (G < 0) {
        if (B < 4 || max_D <= 256) return 0;       // nothing to split / too little work per variable
        if (nb * 8 <= sms) G = 8;
        else if (nb * 4 <= sms) G = 4;
        else return 0;
    }
    if (G != 4 && G != 8) return 0;
    return G;
}

template <int BPC>
static int launch_cluster_t(const ClusterArgs& a, cudaStream_t s)
{
    const size_t smem = rc_smem_bytes<BPC>(a.DPmax, a.NC);
    if (smem > (size_t)irec_device().max_smem_optin) return irec_fail(IREC_E_CAPACITY, "beam_encode: cluster kernel does not fit shared memory");
    if (cudaFuncSetAttribute(k_beam_encode_cluster<BPC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess)
        return irec_fail(IREC_E_CUDA, "beam_encode: cudaFuncSetAttribute(cluster) failed");
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)(a.nb * a.G), 1, 1);
    cfg.blockDim = dim3(RC_THREADS, 1, 1);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = s;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = (unsigned)a.G; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    if (cudaLaunchKernelEx(&cfg, k_beam_encode_cluster<BPC>, a) != cudaSuccess)
        return irec_fail(IREC_E_CUDA, "beam_encode: cluster launch failed");
    irec_count_launch();
    return irec_check_launch("k_beam_encode_cluster");
}

int irec_launch_cluster(int G, const float* t_loc, const float* t_scale, const float* p_loc, const float* p_scale,
                        const int64_t* gidx, const int64_t* offs, int nb, int max_D, float omega, int S, int B, int64_t seed,
                        int32_t* out_indices, int max_aux, int32_t* out_n_aux, int32_t* out_status, float* out_sample,
                        int2* hist, const int32_t* order, const void* plan, const void* tab, int tab_aux, const void* tab_priv, cudaStream_t s)
{
    ClusterArgs a;
    a.t_loc = t_loc; a.t_scale = t_scale; a.p_loc = p_loc; a.p_scale = p_scale;
    a.gidx = gidx; a.offs = offs; a.nb = nb; a.omega = omega; a.S = S; a.B = B; a.seed = seed;
    a.out_indices = out_indices; a.max_aux = max_aux; a.out_n_aux = out_n_aux; a.out_status = out_status;
    a.out_sample = out_sample; a.T2 = irec_device().d_T2; a.dl4 = irec_device().d_dl4;
    a.ratio_tab = irec_ratio_tab(); a.ratio_len = irec_ratio_len();
    a.hist = hist; a.DPmax = make_geom(max_D).DP; a.NC = ((S * B + 31) / 32) * 32;
    a.order = order; a.plan = reinterpret_cast<const R2Plan*>(plan); a.tab = reinterpret_cast<const uint2*>(tab); a.tab_aux = tab_aux; a.tab_priv = reinterpret_cast<const uint2*>(tab_priv);
    a.G = G;
    const int bpc = (B + G - 1) / G;
    switch (bpc) {
        case 1: return launch_cluster_t<1>(a, s);
        case 2: return launch_cluster_t<2>(a, s);
        case 3: return launch_cluster_t<3>(a, s);
        case 4: return launch_cluster_t<4>(a, s);
        case 5: return launch_cluster_t<5>(a, s);
        default: return launch_cluster_t<8>(a, s);       // B <= 32, G >= 4
    }
}
