/*
 * irec.h -- C ABI of libirec.so: the B200 (sm_100a) implementation of the iREC encode/decode
 * inner loop of gergely-flamich/relative-entropy-coding (rec/coding).
 *
 * Plain pointers and sizes only.  Unless marked HOST, every pointer is a DEVICE pointer valid on
 * the current CUDA device; `stream` is a cudaStream_t passed as void*.  All entry points return 0
 * on success or a negative IREC_E_* code (text via irec_last_error_string()).  Calls are
 * asynchronous on `stream` unless stated otherwise.  Citations are relative to the reference
 * repository root.
 *
 * Layout of a call: a flat float32 tensor of N elements is cut into `nb` coder-blocks; block b
 * covers positions [block_offsets[b], block_offsets[b+1]) of the PERMUTED order, and position e
 * maps to flat element  gather_idx ? gather_idx[e] : e  (this is Coder.split/merge,
 * rec/coding/coder.py:38-122, fused into the kernels' loads and stores).
 */
#ifndef IREC_H_
#define IREC_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define IREC_OK 0
#define IREC_E_INVALID (-1)   /* bad argument                                                    */
#define IREC_E_CUDA (-2)      /* CUDA runtime error                                              */
#define IREC_E_CAPACITY (-3)  /* n_aux > max_aux, workspace too small, unsupported size          */
#define IREC_E_CODING (-4)    /* the reference's CodingError conditions (KL not finite, n_aux=0) */

/* per-block status written to out_status[] by the encode kernels */
#define IREC_BLK_OK 0
#define IREC_BLK_BAD_KL 1      /* KL is NaN/inf or n_aux <= 0 (beam coder: reference crashes on an unbound name) */
#define IREC_BLK_TOO_LONG 2    /* n_aux > max_aux */

int irec_version(void);
/* sha256 (hex) of the sources, headers and compiler flags this binary was built from; __graft_entry__.build() compares
 * it with the tree and rebuilds on mismatch */
const char* irec_build_hash(void);
const char* irec_last_error_string(void);

/* One-time per-device tables: the 10006-entry float32 normal-quantile table
 * T[k] = ndtri_f32(float32(k)/float32(10007)) (replaces tfd.Normal.quantile in
 * rec/coding/beam_search_coder.py:48-49) and the power-law ratio table
 * ratio(i) = float32(pow(i+1, -0.7864636765648174)) (rec/coding/coder.py:16,218-220).
 * Called lazily by every other entry point; synchronous on first use. */
int irec_init(void);
/* copies the device quantile table (10007 floats, entry 0 unused) to HOST memory (tests) */
int irec_get_ndtri_table(float* host_out);
/* HOST: ratio(i) as the library computes it */
float irec_aux_ratio(int i);
/* Learned auxiliary variance ratios (GaussianCoder(extrapolate_auxiliary_ratios=False),
 * rec/coding/coder.py:197-231: `aux_variable_variance_ratios[index]` instead of the power law).
 * Installs a DEVICE table of n float32 ratios for every later call made by the CALLING HOST THREAD
 * (encode, decode, the sharded state functions) until it is cleared with (NULL, 0).  The table is
 * borrowed: it must stay alive until the work enqueued under it has finished.  A coder-block that needs
 * more auxiliary variables than n reports IREC_BLK_TOO_LONG (the reference raises CodingError there,
 * coder.py:226-231). */
int irec_set_thread_aux_ratios(const float* dev_ratios, int n);
/* Streams: SMs that the persistent batch kernel of irec_beam_encode leaves free, for every later call made by the CALLING
 * HOST THREAD (k < 0: back to the default, 0 or the environment variable IREC_RESERVE_SMS).  A CTA of that kernel occupies
 * its SM completely; a caller that runs independent sub-batches on several streams (a level's launch of sub-batch A beside
 * the small network kernels that prepare sub-batch B's next level) reserves a few SMs so that those kernels can start while
 * a launch is running.  Results do not depend on it. */
int irec_set_thread_reserved_sms(int k);
/* HOST: number of auxiliary variables the ratio table in force for the calling thread covers (the learned table's
 * length, else the power-law table's 65536): an index list longer than this cannot be decoded (coder.py:226-231) */
int irec_aux_ratio_len(void);

/* HOST helpers restating TensorFlow's seed plumbing (python/framework/random_seed.py,
 * python/eager/context.py): op seed of the first unseeded random op after tf.random.set_seed(seed),
 * i.e. random.Random(seed).randint(0, 2**31-1). */
int64_t irec_tf_op_seed(int64_t seed);

/* HOST: the permutation of Coder.split (rec/coding/coder.py:60-67): tf.random.set_seed(seed);
 * tf.random.shuffle(range(n)).  perm_out is HOST memory, n int64 entries. */
int irec_split_permutation(int64_t n, int64_t seed, int64_t* perm_out);

/* raw candidate streams (tests / diagnostics), device output:
 *   irec_beam_uniform_int : r[j] = 1 + Philox(q)[start+j] % 10006  (beam_search_coder.py:39-43)
 *   irec_is_normal_stream : z[j] of Normal(0,1).sample after set_seed(seed) (importance_sampling.py:38,54) */
int irec_beam_uniform_int(int64_t q, int64_t start, int64_t n, int32_t* out, void* stream);
int irec_is_normal_stream(int64_t seed, int64_t start, int64_t n, float* out, void* stream);
/*   irec_normal_stream_seeded : z[j] of tf.random.normal(..., seed=op_seed) after set_seed(global_seed), i.e.
 *   tfd.Normal.sample(n, seed=op_seed) before scale/shift -- the candidate buffers of the rejection sampler's
 *   NaiveSampleGenerator (rec/coding/sample_generator.py:53-66) */
int irec_normal_stream_seeded(int64_t global_seed, int64_t op_seed, int64_t start, int64_t n, float* out, void* stream);
/*   irec_uniform_int_stream : tf.random.uniform(shape, lo, hi, dtype=int32, seed=op_seed) after set_seed(global_seed):
 *   lo + u32 % (hi - lo) over the same Philox stream -- the group / sample assignments of the rejection sampler's
 *   PseudoSampleGenerator (rec/coding/sample_generator.py:84-93); irec_beam_uniform_int is the (1, 10007) case */
int irec_uniform_int_stream(int64_t global_seed, int64_t op_seed, int32_t lo, int32_t hi, int64_t start, int64_t n, int32_t* out,
                            void* stream);
/* HOST (tests): the table-driven float64 log / sincos behind the Box-Muller candidates (csrc/irec_boxmuller.cuh),
 * evaluated on the host with the operation sequence of the device, for m[i] = the 23 mantissa bits of a Philox word:
 * logf_out = float(log(max(Uint32ToFloat(m), 1e-7f))), (sin_out, cos_out) = float(sin/cos(float(2 pi Uint32ToFloat(m)))).
 * Any output pointer may be NULL.  Used to check ALL 2^23 arguments against the float64-libm definition. */
int irec_bm_components_host(const uint32_t* m, int64_t n, float* logf_out, float* sin_out, float* cos_out);

/* KL(target || coder) summed per block and n_aux = ceil(KL / omega)
 * (rec/coding/coder.py:499-501, beam_search_coder.py:57-59).  out_kl/out_n_aux: [nb]. */
int irec_kl_naux(const float* t_loc, const float* t_scale, const float* p_loc, const float* p_scale,
                 const int64_t* gather_idx, const int64_t* block_offsets, int nb, float omega,
                 float* out_kl, int32_t* out_n_aux, void* stream);

/* KL, n_aux and the per-partition per-dim coefficients of ONE block of D contiguous dims -- the schedule both the encode
 * and the decode kernels evaluate on the fly (rec/coding/beam_search_coder.py:57-77, coder.py:141-154): for auxiliary
 * variable t and dim d, sa = sqrt(v_t) (scale of the candidates), M = mean of the auxiliary target, and the centred
 * quadratic log-weight coefficients A = (1/tot - 1/s2) / 2, E = M / tot.  Outputs [max_aux x D] row-major, rows t < n_aux
 * written; nothing is written when n_aux is invalid (check out_n_aux: <= 0, > max_aux, beyond the ratio table). */
int irec_schedule(const float* t_loc, const float* t_scale, const float* p_loc, const float* p_scale, int D, float omega,
                  int max_aux, float* out_kl, int32_t* out_n_aux, float* out_sa, float* out_A, float* out_E, float* out_M,
                  void* stream);

/* ---- beam-search coder (rec/coding/beam_search_coder.py) ------------------------------------ */

/* bytes of device workspace irec_beam_encode needs for these sizes */
size_t irec_beam_encode_workspace_bytes(int nb, int64_t max_block_dim, int S, int B, int max_aux);

/* BeamSearchCoder.encode_block over nb blocks (beam_search_coder.py:53-122; called from
 * GaussianCoder.encode, coder.py:412-457).  S = int(exp(kl_per_partition*extra_samples)) candidates
 * per partition, B = n_beams, omega = kl_per_partition (nats), coding seed `seed` (partition t of
 * every block uses seed+t).  Outputs: out_indices [nb x max_aux] (partition order, first n_aux[b]
 * valid), out_n_aux [nb], out_status [nb] (IREC_BLK_*), out_sample [N] (= beam[0] + p_loc,
 * scattered through gather_idx).  1 <= B <= 1024 (the reference has no limit; B > 32 takes k_beam_encode_wide and needs
 * max_block_dim <= 1024), S * B < 2^31; IREC_E_CAPACITY otherwise. */
int irec_beam_encode(const float* t_loc, const float* t_scale, const float* p_loc, const float* p_scale,
                     const int64_t* gather_idx, const int64_t* block_offsets, int nb, int64_t max_block_dim,
                     float omega, int S, int B, int64_t seed,
                     int32_t* out_indices, int max_aux, int32_t* out_n_aux, int32_t* out_status,
                     float* out_sample, void* workspace, size_t workspace_bytes, void* stream);

/* Which kernel irec_beam_encode runs for these sizes on the current device (diagnostics, tests, bench reports):
 * 0 = one partition per launch (general path), 1 = k_beam_encode_resident, 2 = k_beam_encode_resident2 (persistent CTA per
 * coder-block), 3 = k_beam_encode_tmem (beams and coefficients in tensor memory, two coder-blocks per SM), 4 = k_beam_encode_wide
 * (n_beams > 32), 100 + G = k_beam_encode_cluster with G CTAs per coder-block (few blocks per launch: one image).  For B > 32
 * a return of 0 means "not supported". */
int irec_beam_encode_path(int nb, int64_t max_block_dim, int S, int B);

/* BeamSearchCoder.decode_block over nb blocks (beam_search_coder.py:124-148).  indices
 * [nb x max_aux] in partition order (the order encode returns), n_aux [nb].  out_status [nb] (may be
 * NULL): IREC_BLK_TOO_LONG for a block whose n_aux is negative, exceeds max_aux or exceeds the
 * auxiliary-ratio table in force (the reference raises CodingError, coder.py:226-231); such a block
 * decodes to NaN. */
int irec_beam_decode(const float* p_loc, const float* p_scale, const int64_t* gather_idx,
                     const int64_t* block_offsets, int nb, int S, int64_t seed,
                     const int32_t* indices, int max_aux, const int32_t* n_aux,
                     float* out_sample, int32_t* out_status, void* stream);

/* ---- one partition at a time, candidate range [s_begin, s_end) -- the multi-GPU path ---------
 * The candidate index space of ONE coder-block is sharded: every rank holds a replica of the
 * block's state (an opaque device buffer laid out by the library), scores its own contiguous range
 * of candidate samples s, and the per-rank top-B records are exchanged (ncclAllGather of
 * B * sizeof(irec_record_t) bytes) and merged identically everywhere; winners are re-materialised
 * locally from the counter-based RNG, so no sample data crosses NVLink.
 * The same calls with the full range [0, S) are the single-GPU general path for blocks that do not
 * fit the shared-memory-resident kernel (D > 1024 or very large S).
 *   per block:      irec_beam_state_init -> irec_beam_state_query (n_aux; syncs)
 *   per partition:  irec_beam_step_score (local top-B)  -> [all-gather] -> irec_beam_step_commit
 *   at the end:     irec_beam_state_finish (indices, sample)                                      */
typedef struct {
    float score;     /* canonical log-weight without the per-partition constant */
    int32_t s;       /* candidate sample index  (beam_search_coder.py:89 best_ind_aux)  */
    int32_t b;       /* parent beam index       (beam_search_coder.py:88 best_ind_beam) */
    int32_t pad;
} irec_record_t;

size_t irec_beam_state_bytes(int D, int B, int max_aux);
int irec_beam_state_init(void* state, const float* t_loc, const float* t_scale, const float* p_loc,
                         const float* p_scale, const int64_t* gather_idx, int64_t offset, int D, float omega, int S,
                         int B, int max_aux, int64_t seed, void* stream);
/* HOST outputs; synchronises `stream` */
int irec_beam_state_query(const void* state, int32_t* n_aux, int32_t* status, float* kl, void* stream);
size_t irec_beam_step_workspace_bytes(int D, int B);
/* out_records [B] sorted best first, out_count [1] (device) */
int irec_beam_step_score(void* state, int D, int B, int t, int64_t s_begin, int64_t s_end, int do_params,
                         irec_record_t* out_records, int32_t* out_count, void* workspace, size_t workspace_bytes,
                         void* stream);
/* records: n_lists lists of B records each; counts[i] valid entries in list i (NULL: all B) */
int irec_beam_step_commit(void* state, int D, int B, int t, const irec_record_t* records, const int32_t* counts,
                          int n_lists, void* workspace, size_t workspace_bytes, void* stream);
int irec_beam_state_finish(void* state, int D, const int64_t* gather_idx, int64_t offset, int32_t* out_indices,
                           int32_t* out_n_aux, int32_t* out_status, float* out_sample, void* stream);
/* stand-alone merge of n_records records into the best B by (score desc, s*Bcur+b asc);
 * workspace >= 8 * n_records bytes */
int irec_topb_merge(const irec_record_t* records, int n_records, int Bcur, int B, irec_record_t* out_records,
                    int32_t* out_count, void* workspace, size_t workspace_bytes, void* stream);

/* Exchange step of the candidate-range sharded coder over NVLink peer memory instead of an NCCL all-gather
 * (the argsort at beam_search_coder.py:86 split across GPUs).  peer_bufs: DEVICE array of `world`
 * pointers to the ranks' exchange buffers (symmetric allocations of irec_p2p_exchange_bytes(B, world) bytes,
 * zero-initialised, peer-mapped into this process, entry `rank` = the local one).  Pushes the B records + count of
 * this rank (layout of irec_beam_step_score's out_records / out_count, contiguous) to every rank, waits for every
 * rank's records of the same step, writes them as out_records [world][B] / out_counts [world] for
 * irec_beam_step_commit.  Every rank must call it the same number of times. */
size_t irec_p2p_exchange_bytes(int B, int world);
int irec_p2p_exchange(void* const* peer_bufs, int rank, int world, int B, const irec_record_t* local_records_and_count,
                      irec_record_t* out_records, int32_t* out_counts, void* stream);

/* The same loop as ONE cooperative launch per coder-block (all auxiliary variables; beam_search_coder.py:66-109): every
 * CTA keeps a replica of the block state in shared memory, scores its share of this rank's candidates [s_begin, s_end),
 * and per auxiliary variable the last CTA to arrive merges the grid's top-B lists, exchanges the rank list with the peers
 * through peer_bufs (layout and protocol of irec_p2p_exchange; NULL with world = 1) and publishes the winners.  Replaces
 * the seven launches per variable of step_score / p2p_exchange / step_commit where the block state fits shared memory
 * (D <= 256 at 20 beams; IREC_E_CAPACITY otherwise).  Call between irec_beam_state_init and irec_beam_state_finish;
 * every rank must call it for the same block. */
int irec_beam_fused_fits(int D, int B);          /* 1: the fused launch covers these sizes on the current device */
size_t irec_beam_fused_workspace_bytes(int B, int world);
int irec_beam_encode_fused(void* state, int D, int B, int64_t s_begin, int64_t s_end, void* const* peer_bufs, int rank,
                           int world, void* workspace, size_t workspace_bytes, void* stream);

/* ---- importance sampler (rec/coding/importance_sampling.py, samplers.py:61-101) -------------- */

/* encode_gaussian_importance_sample with alpha = inf (importance_sampling.py:9-79): one partition,
 * S candidates, D dims (contiguous arrays).  out_index: device int64[1]; out_sample: [D]. */
size_t irec_is_workspace_bytes(int D);
int irec_is_coded_sample(const float* t_loc, const float* t_scale, const float* p_loc, const float* p_scale,
                         int D, int64_t S, int64_t seed, int64_t* out_index, float* out_sample,
                         void* workspace, size_t workspace_bytes, void* stream);
/* decode_gaussian_importance_sample (importance_sampling.py:82-103); index is a device int64[1] */
int irec_is_decode_sample(const float* p_loc, const float* p_scale, int D, const int64_t* index, int64_t seed,
                          float* out_sample, void* stream);

/* GaussianCoder.encode_block / decode_block with an ImportanceSampler over nb blocks
 * (coder.py:493-584): auxiliary-variable loop with on-device conditioning (coder.py:141-171).
 * out_indices [nb x max_aux] int64 in the order the reference appends them (i = n_aux-1..1, final);
 * out_n_idx [nb] = max(n_aux, 1).  Encode: blocks of at most 4096 dims.  Decode: out_status [nb] (may be
 * NULL) flags blocks whose index count is < 1, > max_aux or > the ratio table in force with
 * IREC_BLK_TOO_LONG (the reference raises IndexError / CodingError); they decode to NaN.  Both calls are
 * asynchronous on `stream`. */
size_t irec_is_block_workspace_bytes(int max_aux);            /* decode; minimum for encode */
/* encode with room for the launch's candidate table: the S x D standard normals of every partition are the same for
 * all coder-blocks of one size (the coding seed does not depend on the block, coder.py:523,538), so they are evaluated
 * once per launch and streamed from L2; with only irec_is_block_workspace_bytes the kernel regenerates them per block */
size_t irec_is_encode_workspace_bytes(int nb, int64_t max_block_dim, int64_t S, int max_aux);
int irec_is_encode(const float* t_loc, const float* t_scale, const float* p_loc, const float* p_scale,
                   const int64_t* gather_idx, const int64_t* block_offsets, int nb, int64_t max_block_dim,
                   float omega, int64_t S, int64_t seed,
                   int64_t* out_indices, int max_aux, int32_t* out_n_idx, int32_t* out_status,
                   float* out_sample, void* workspace, size_t workspace_bytes, void* stream);
int irec_is_decode(const float* p_loc, const float* p_scale, const int64_t* gather_idx,
                   const int64_t* block_offsets, int nb, int64_t max_block_dim, int64_t seed,
                   const int64_t* indices, int max_aux, const int32_t* n_idx,
                   float* out_sample, int32_t* out_status, void* workspace, size_t workspace_bytes, void* stream);

/* number of kernels this library has launched in the calling process (bench.py "gpu_launches") */
int64_t irec_launch_count(void);

#ifdef __cplusplus
}
#endif
#endif /* IREC_H_ */
