/*
 * cogstream_b200 -- C ABI of the B200-native Visual Stream Compression hot path.
 *
 * The reference (LiamZhao326/CogStream) is pure Python + torch and has NO FFI of its own;
 * this header is the boundary a maintainer binds (ctypes) to replace the
 * torch op sequences of
 *     model/kmeans_with_time.py:4-137        (temporal-semantic k-means)
 *     model/cogreasoner_chat.py:50-64        (select_additional_frames)
 *     model/cogreasoner_chat.py:278-333      (select_events_based_on_summary, non-LLM parts)
 *     model/cogreasoner_chat.py:383-432      (_get_compression_mask)
 *     model/cogreasoner_chat.py:434-447      (compress_unimportant_events)
 *     model/cogreasoner_chat.py:449-476      (_compress_visual_tokens)
 * Each entry point cites the reference lines it replaces.
 *
 * Conventions
 *   - plain pointers and sizes only; every pointer is DEVICE memory owned by the caller unless
 *     it says "host"; nothing is allocated, freed or retained by the library;
 *   - all work is enqueued on the given CUDA stream (a cudaStream_t passed as void*); no entry
 *     point synchronises with the host;
 *   - return 0 on success, a negative csb_status otherwise; csb_last_error() gives the text of
 *     the last failure on the calling thread;
 *   - sm_100a only: every launcher fails with CSB_ERR_ARCH on another device; there is no CPU
 *     fallback;
 *   - "bf16" buffers are raw 16-bit bfloat16; rows are contiguous and 16-byte aligned
 *     (D % 8 == 0), which the reference already needs for its .view() calls
 *     (kmeans_with_time.py:35, cogreasoner_chat.py:279).
 */
#ifndef COGSTREAM_B200_H_
#define COGSTREAM_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CSB_ABI_VERSION 2

#if defined(__GNUC__)
#define CSB_API __attribute__((visibility("default")))
#else
#define CSB_API
#endif

typedef enum csb_status {
  CSB_OK = 0,
  CSB_ERR_ARG = -1,       /* bad shape / null pointer / misalignment */
  CSB_ERR_ARCH = -2,      /* device is not sm_100 */
  CSB_ERR_CUDA = -3,      /* launch or runtime failure (text in csb_last_error) */
  CSB_ERR_WORKSPACE = -4, /* workspace too small */
  CSB_ERR_UNSUPPORTED = -5
} csb_status;

typedef enum csb_dtype { CSB_F32 = 0, CSB_F64 = 1, CSB_BF16 = 2 } csb_dtype;

typedef void* csb_stream_t; /* cudaStream_t */

CSB_API int csb_abi_version(void);
CSB_API const char* csb_last_error(void);
/* 0 if the current device can run the kernels (compute capability 10.0 exactly: the cubins are sm_100a), else
 * CSB_ERR_ARCH. */
CSB_API int csb_check_device(void);
/* Kernels this library has launched in this process so far (bench.py reports the count per timed region). */
CSB_API uint64_t csb_kernel_launches(void);

/* ---------------------------------------------------------------------------------------------
 * (a) frame-token pooling + squared L2 norm.                     cogreasoner_chat.py:444-446 (pool),
 *     x [T,P,D] bf16 -> pooled_bf16 [T,D] (fp32 accumulate, one RN)  kmeans_with_time.py:73 (the ||x||^2
 *                       pooled_f32  [T,D], sqnorm [T] fp32           term of cdist); any output may be NULL.
 * ------------------------------------------------------------------------------------------- */
CSB_API int csb_pool_norm(const void* x_bf16, int64_t T, int64_t P, int64_t D, void* pooled_bf16, float* pooled_f32,
                  float* sqnorm, csb_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * (b) k-means.
 * csb_gram: G[i][j] = <x_i, x_j> over the flattened [P*D] frame vectors, fp64 [T,T] (both
 *     triangles).  bf16 products are exact in fp32; the tcgen05 kernel accumulates fp32 per
 *     K-split in TMEM and the splits are summed in fp64 in a fixed order (deterministic).
 *     Every feature distance the reference takes from torch.cdist (kmeans_with_time.py:48, :73;
 *     cogreasoner_chat.py:61) is ||x_i||^2 + ||c||^2 - 2<x_i,c> with c a frame or a mean of
 *     frames, i.e. a function of G.
 *     impl: 0 = tcgen05/TMA kernel (default), 1 = CUDA-core reference kernel (validation only); higher bits are
 *     validation / profiling switches (512: force the partial-tile + reduce path).
 * ------------------------------------------------------------------------------------------- */
CSB_API size_t csb_gram_workspace_bytes(int64_t T, int64_t K);
CSB_API int csb_gram(const void* x_bf16, int64_t T, int64_t K, double* G, void* workspace, size_t workspace_bytes,
             int impl, csb_stream_t stream);

/* Incremental form for an append-only stream (the reference re-clusters the whole history for every new
 * question, evaluate/answer_generate.py:120-135): frames [0,T_old) are already in G (leading dimension ldg >= T); only the
 * tiles that touch frames [T_old,T) are computed -- work ~ (T - T_old) * T * K instead of T^2 * K.  x holds all T frames. */
CSB_API int csb_gram_extend(const void* x_bf16, int64_t T_old, int64_t T, int64_t K, double* G, int64_t ldg, void* workspace,
                            size_t workspace_bytes, int impl, csb_stream_t stream);

typedef struct csb_kmeans_params {
  int64_t T;              /* frames */
  int64_t k;              /* clusters, T > k >= 1 (T <= k is the caller's early return, :30-32) */
  int32_t ts_dtype;       /* CSB_F32 or CSB_F64: dtype of timestamps / centroid_times */
  int32_t max_iteration;  /* :4  default 30; < 0: the k-means++ init only (the frame-sharded Lloyd phases follow) */
  double alpha;           /* :4  default 2 */
  double tol;             /* :4  default 1e-4 */
  int64_t n_init_given;   /* how many entries of init_indices are prescribed (>=1: the first centre is
                             always drawn by the host, :41); the remaining k - n are sampled k-means++ */
  int64_t n_reseed_pool;  /* entries available in reseed_pool (host random.randint(0, T-1) draws) */
  int64_t ldg;            /* leading dimension of G (elements); 0 means T */
  void* rng_state;        /* device uint64[2] {seed, offset} or NULL.  With it the solver draws its own randomness (Philox4x32-10):
                             the Exp(1) noise of every sampling round (exp_noise may then be NULL), the random.randint draws for
                             degenerate rounds and empty clusters (reseed_pool NULL: they cannot run out), and -- with
                             n_init_given == 0 -- the first centre.  offset is advanced by one per solve (CUDA-graph replays draw
                             fresh numbers).  Same distributions as the reference's generators, not the same streams: the
                             drop-in function keeps torch's / random's own streams through exp_noise + reseed_pool. */
  int64_t frames_per_warp; /* assignment pass: sequential frame passes of a warp per Lloyd iteration.  A pass covers 32 / L frames
                             side by side (L lanes per frame: 8 / 16 / 32 for k <= 32 / 64 / more), and the cooperative launch
                             uses T / (8 * (32 / L) * frames_per_warp) CTAs.  0 = default (1 pass when L < 32, else 2): 38 CTAs
                             at T = 600, k = 40.  Fewer passes shorten an iteration, more leave SMs to concurrent streams. */
} csb_kmeans_params;

/* Device-side result block, written by csb_kmeans_solve (read it after a stream sync). */
typedef struct csb_kmeans_status {
  int32_t iterations;        /* Lloyd iterations executed */
  int32_t converged;         /* 1 if centre shift <= tol (:130) */
  int32_t n_reseed_used;     /* entries of reseed_pool consumed: degenerate init rounds (:57) first, then
                                empty-cluster re-seeds (:118), in the order the reference calls random.randint */
  int32_t degenerate_rounds; /* init rounds that hit probs_sum == 0 (:55) */
  int32_t reseed_overflow;   /* 1 if reseed_pool ran out (result invalid; re-run with a larger pool) */
  int32_t n_noise_used;      /* rows of exp_noise consumed (= multinomial calls the reference would make) */
  float min_margin;          /* smallest (2nd best - best) final distance over all rows/iterations */
  float last_shift;          /* centre shift of the last iteration */
} csb_kmeans_status;

CSB_API size_t csb_kmeans_workspace_bytes(int64_t T, int64_t k);
/*
 * k-means++ init (kmeans_with_time.py:41-62) + Lloyd loop (:71-131) in ONE cooperative launch,
 * no host round trips.  Sampling round m computes probs exactly as :46-59 and picks
 * argmax(probs / exp_noise[m-1]) -- what torch.multinomial(probs, 1) evaluates internally with
 * q ~ Exp(1) (the host draws exp_noise from the torch generator the reference would have used).
 * Noise rows are consumed one per sampling round; a round with probs_sum == 0 (:55-57) takes the next
 * reseed_pool entry instead, so both host RNG streams are consumed exactly like the reference does.
 * Streams of more than 1024 frames take two launches: the k-means++ rounds run on one thread-block cluster (<= 16 CTAs; the round
 * records travel through distributed shared memory; with params->rng_state further clusters of the same launch pre-generate
 * the noise rows), then the cooperative Lloyd kernel.  min_margin is exact whenever it is below ~1e-6 (k > 128: clusters
 * whose fp32 distance estimate is provably not the smallest are not evaluated exactly); everything else is bit-identical.
 *   G             [T,T] fp64 (csb_gram)
 *   timestamps    [T]   ts_dtype
 *   init_indices  [k]   int64; first n_init_given entries are inputs (0 allowed with params->rng_state), the rest are outputs
 *   exp_noise     [k-1,T] fp32 or NULL when n_init_given == k
 *   reseed_pool   [n_reseed_pool] int64 frame indices for empty clusters, consumed in order
 *   assignments   [T]   int64  out (:104, lowest index wins ties)
 *   centroid_times[k]   ts_dtype out (:115/:120)
 *   member_offsets[k+1] int32 out, members [T+k] int32 out: CSR of the frames that DEFINE each returned
 *                  centre (its assigned frames ascending; a re-seeded empty cluster lists its seed frame)
 *   event_order   [T] int32 out or NULL, event_offsets [k+1] int32 out or NULL: frames sorted by (event, frame) and
 *                  the per-event ranges -- the chronological per-event gather order of cogreasoner_chat.py:291-296
 *                  (same as csb_event_order on the returned assignments; a re-seeded empty cluster has an empty range)
 *   status        device csb_kmeans_status
 */
CSB_API int csb_kmeans_solve(const double* G, const void* timestamps, const csb_kmeans_params* params /* host */,
                     int64_t* init_indices, const float* exp_noise, const int64_t* reseed_pool,
                     int64_t* assignments, void* centroid_times, int32_t* member_offsets, int32_t* members,
                     int32_t* event_order, int32_t* event_offsets, csb_kmeans_status* status, void* workspace,
                     size_t workspace_bytes, csb_stream_t stream);

/* ---- frame-sharded Lloyd loop: ONE long stream over the GPUs of a box (first half) --------------------------
 * Every rank holds G (the per-stream cache keeps it current) and runs the k-means++ init redundantly -- csb_kmeans_solve with
 * max_iteration < 0 = init only (k - 1 dependent sampling rounds: nothing to shard).  The Lloyd iteration is split by frames:
 * rank r owns frames [r * per_rank, min(T, (r + 1) * per_rank)).  The caller drives the phases and issues the collectives
 * between them (NCCL all-gather over NVLink; the buffers are laid out rank-major so the gathers are in place):
 *     CSB_LLOYD_INIT    S[0] rows of the own frames, state reset                 (after the init)
 *   repeat until converged / max_iteration:
 *     CSB_LLOYD_ASSIGN  distances + argmin of the own frames -> assign32[f_begin:f_end)        (kmeans_with_time.py:73-104)
 *         all-gather assign32  [world * per_rank] int32
 *     CSB_LLOYD_UPDATE  (identical on every rank) membership, centre times, re-seeds (:107-120); S rows of the own frames for
 *                       the changed clusters; exchange[rank] = {S_nxt[i][a_i], S_cur[i][a_i] of the own frames, seed entries}
 *         all-gather exchange  [world * (2 * per_rank + k + 1)] fp64
 *     CSB_LLOYD_SHIFT   (identical on every rank) ||c_j||^2, centre shift (:123-125), convergence flag -> state
 *         read csb_kmeans_lloyd_converged (one small D2H per iteration, like the reference's `if shift <= tol`)
 *     CSB_LLOYD_OUTPUT  the outputs of csb_kmeans_solve (assignments, centroid_times, CSR, event order, status)
 * Same device functions and summation orders as csb_kmeans_solve: assignments, iterations, re-seeds, centre times bit-identical.
 * All ranks must pass the same G, timestamps, init_indices, reseed_pool / rng_state contents. */
enum { CSB_LLOYD_INIT = 1, CSB_LLOYD_ASSIGN = 2, CSB_LLOYD_UPDATE = 3, CSB_LLOYD_SHIFT = 4, CSB_LLOYD_OUTPUT = 5 };
typedef struct csb_lloyd_shard {
  int32_t phase;
  int32_t reserved;
  int64_t f_begin, f_end;  /* own frames */
  int64_t per_rank;        /* frames per rank = ceil(T / world) (or any value with world * per_rank >= T) */
  void* exchange;          /* device fp64 [world * (2 * per_rank + k + 1)] */
  void* state;             /* device, csb_kmeans_lloyd_state_bytes(T, k); int32 word 3 of its last 256 bytes = converged flag */
  size_t state_bytes;
} csb_lloyd_shard;
CSB_API size_t csb_kmeans_lloyd_state_bytes(int64_t T, int64_t k);
CSB_API int csb_kmeans_lloyd_phase(const double* G, const void* timestamps, const csb_kmeans_params* params /* host */,
                           const csb_lloyd_shard* shard /* host */, int64_t* init_indices, const int64_t* reseed_pool,
                           int32_t* assign32 /* [world * per_rank] */, int64_t* assignments, void* centroid_times,
                           int32_t* member_offsets, int32_t* members, int32_t* event_order, int32_t* event_offsets,
                           csb_kmeans_status* status, void* workspace /* csb_kmeans_workspace_bytes */, size_t workspace_bytes,
                           csb_stream_t stream);

/* Centroid features from the defining members: C[j] = (sum of member rows, ascending, fp32) / n_j.
 * kmeans_with_time.py:110-120, one pass over the member rows; atomic-free and deterministic. */
CSB_API int csb_centroid_update(const void* x_bf16, int64_t T, int64_t K, const int32_t* member_offsets,
                        const int32_t* members, int64_t k, float* centroids /* [k,K] */, csb_stream_t stream);

/* csb_centroid_update fused with the chronological per-event token gather (cogreasoner_chat.py:291-296): one pass
 * over x produces the centroids AND the event-major packed copy packed[event_offsets[j] + r] = x[members of j][r].
 * Members of a centre beyond its event range (the seed frame of a re-seeded empty cluster) only feed the centroid.
 * centroids or packed may be NULL. */
CSB_API int csb_centroid_gather(const void* x_bf16, int64_t T, int64_t K, const int32_t* member_offsets,
                                const int32_t* members, const int32_t* event_offsets, int64_t k, float* centroids,
                                void* packed_bf16, csb_stream_t stream);

/* dist[i] = || x_i - C[assign[i]] ||_2 (fp32), one pass over x.  cogreasoner_chat.py:57-61. */
CSB_API int csb_dist_to_assigned(const void* x_bf16, int64_t T, int64_t K, const float* centroids, int64_t k,
                         const int64_t* assignments, float* dist, csb_stream_t stream);
/* Same distances from the Gram matrix and the event CSR of csb_event_order (no pass over x).
 *   rowsum_ws [T] fp64 scratch. */
CSB_API int csb_dist_to_assigned_gram(const double* G, int64_t ldg, int64_t T, int64_t k, const int64_t* assignments,
                              const int32_t* event_offsets, const int32_t* order, float* dist,
                              double* rowsum_ws, csb_stream_t stream);
/* Per event the n_rep members with the smallest dist (ascending dist, lowest index on ties); events with
 * <= n_rep members list all members ascending.  cogreasoner_chat.py:55-63.
 *   reps [k,n_rep] int64 (-1 padded), rep_counts [k] int32, rep_flags [T] uint8 (1 = representative). */
CSB_API int csb_select_representatives(const float* dist, const int64_t* assignments, int64_t T, int64_t k, int64_t n_rep,
                               int64_t* reps, int32_t* rep_counts, uint8_t* rep_flags, csb_stream_t stream);

/* Both of the above from the Gram matrix in ONE launch (a warp per event; meant for events of up to a few hundred frames --
 * the work per event is quadratic in its size): dist [T] is filled as by csb_dist_to_assigned_gram. */
CSB_API int csb_representatives_gram(const double* G, int64_t ldg, int64_t T, int64_t k, const int32_t* event_offsets,
                                     const int32_t* order, int64_t n_rep, float* dist, int64_t* reps,
                                     int32_t* rep_counts, uint8_t* rep_flags, csb_stream_t stream);

/* Chronological per-event order (cogreasoner_chat.py:291-296): order[] lists frame indices sorted by
 * (event, frame); event_offsets [k+1].  Then csb_gather_rows(order) packs the event tokens. */
CSB_API int csb_event_order(const int64_t* assignments, int64_t T, int64_t k, int32_t* order, int32_t* event_offsets,
                    csb_stream_t stream);
/* dst row r = src row index[r] for r < n (n read from n_dev when non-NULL, else n_host). */
CSB_API int csb_gather_rows(const void* src, int64_t row_bytes, const int32_t* index, int64_t n_host,
                    const int64_t* n_dev, int64_t n_max, void* dst, csb_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * (c) question relevance.  cogreasoner_chat.py:325-330: cosine_similarity(q, events) evaluated with
 *     the dtype's op-by-op rounding (norm, clamp(1e-8), divide, multiply, sum), cos < threshold with the
 *     threshold rounded to the dtype.  dtype CSB_BF16 or CSB_F32.
 *       cos_out [k] (dtype), low_flags [k] uint8
 * ------------------------------------------------------------------------------------------- */
CSB_API int csb_cosine_flags(const void* question, const void* events, int64_t k, int64_t Dl, int32_t dtype,
                     double threshold, void* cos_out, uint8_t* low_flags, csb_stream_t stream);
/* minor[i] = low_flags[assign[i]] && !rep_flags[i]   (cogreasoner_chat.py:331-332)
 *   minor_flags [T] uint8, minor_list [T] int32 ascending, n_minor device int64 */
CSB_API int csb_minor_frames(const int64_t* assignments, const uint8_t* low_flags, const uint8_t* rep_flags, int64_t T,
                     uint8_t* minor_flags, int32_t* minor_list, int64_t* n_minor, csb_stream_t stream);

/* csb_cosine_flags + csb_minor_frames in one launch (the last cosine CTA to finish builds the list).
 *   ticket: a zero-initialised uint32 the library resets after use (one per concurrently running stream). */
CSB_API int csb_relevance_minor(const void* question, const void* events, int64_t k, int64_t Dl, int32_t dtype,
                                double threshold, const int64_t* assignments, const uint8_t* rep_flags, int64_t T,
                                void* cos_out, uint8_t* low_flags, uint8_t* minor_flags, int32_t* minor_list,
                                int64_t* n_minor, uint32_t* ticket, csb_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * (d) token mask, pooling of minor frames, chronological compaction.
 * csb_pixel_mask: one video item [t, P, W] (W = merge^2 * C) -> mask [t*P] uint8.
 *     cogreasoner_chat.py:407-422: keep iff mean|x_f - x_{f-1}| * 255 > threshold in the pixel dtype
 *     (bf16: subtract, mean (fp32 accumulate, one RN), *255 and the threshold each round to bf16);
 *     first frame compares threshold+1; frames with < min_tokens kept keep their first min_tokens;
 *     frames flagged in minor_flags[frame0 + f] keep only token 0.  dtype CSB_BF16 or CSB_F32.
 * ------------------------------------------------------------------------------------------- */
/* The two halves of csb_pixel_mask: the raw difference mask does not depend on the clustering or the question, so a
 * caller can run it early (e.g. on a side stream while the k-means solver runs) and apply the per-frame rules later. */
CSB_API int csb_pixel_diff(const void* pixels, int64_t t, int64_t P, int64_t W, int32_t dtype, double threshold,
                           uint8_t* mask, csb_stream_t stream);
/* Same with two scheduling knobs for a caller that runs the kernel NEXT TO a latency-bound one of the same stream (the k-means
 * solver): max_ctas > 0 caps the grid (64 CTAs stream the 508 MB of config 2 in ~160 us, hidden under the ~190 us solver, and
 * leave the solver's L2 latency alone; 0 = the full grid), flags & CSB_PIXEL_EVICT_FIRST loads the pixels with an L2 evict-first
 * policy (they are read exactly once; the solver's 2.9 MB Gram matrix then stays L2-resident).  csb_pixel_diff = (0, EVICT_FIRST). */
#define CSB_PIXEL_EVICT_FIRST 1
CSB_API int csb_pixel_diff_ex(const void* pixels, int64_t t, int64_t P, int64_t W, int32_t dtype, double threshold,
                              uint8_t* mask, int64_t max_ctas, int32_t flags, csb_stream_t stream);
CSB_API int csb_mask_finalize(uint8_t* mask, int64_t t, int64_t P, int64_t min_tokens, const uint8_t* minor_flags,
                              int64_t frame0, csb_stream_t stream);
CSB_API int csb_pixel_mask(const void* pixels, int64_t t, int64_t P, int64_t W, int32_t dtype, double threshold,
                   int64_t min_tokens, const uint8_t* minor_flags, int64_t frame0, uint8_t* mask,
                   csb_stream_t stream);

/* In-place token-0 pooling of the listed frames (cogreasoner_chat.py:444-446) on a caller-made clone. */
CSB_API int csb_pool_frames_inplace(void* x_bf16, int64_t T, int64_t P, int64_t D, const int32_t* frame_list,
                            int64_t n_host, const int64_t* n_dev, csb_stream_t stream);

CSB_API size_t csb_compact_workspace_bytes(int64_t n_tokens, int64_t n_ids);
/* cogreasoner_chat.py:458 + :444-446 fused: out = x[mask] in chronological order, where token 0 of a
 * frame flagged in pool_flags is emitted as that frame's pooled mean (so no 249 MB clone is made).
 *   kept_index [n_tokens] int32 (token index of every kept row), n_kept device int64. */
CSB_API int csb_compact_tokens(const void* x_bf16, int64_t T, int64_t P, int64_t D, const uint8_t* mask,
                       const uint8_t* pool_flags, void* out_bf16, int32_t* kept_index, int64_t* n_kept,
                       void* workspace, size_t workspace_bytes, csb_stream_t stream);
/* cogreasoner_chat.py:459-474: keep every text id and the <image> ids whose mask bit is set (the j-th image
 * id pairs with mask[j]); optional attention_mask / labels / position_ids (int64) are filtered alike and
 * position_ids are rebuilt as 0..len-1 per sequence (a sequence starts where the filtered id is 0).
 *   n_out device int64[2]: {ids kept, <image> ids seen}.  The caller compares n_out[1] with n_mask: the reference's
 *   `text_masks[image_selected] = compression_mask` (:463) raises when they differ, this entry point does not (image ids past
 *   n_mask are dropped).  Any optional in/out pair may be NULL. */
CSB_API int csb_filter_ids(const int64_t* input_ids, int64_t n_ids, int64_t image_token_index, const uint8_t* mask,
                   int64_t n_mask, const int64_t* attention_mask, const int64_t* labels,
                   const int64_t* position_ids, int64_t* out_ids, int64_t* out_attention_mask,
                   int64_t* out_labels, int64_t* out_position_ids, int64_t* n_out, void* workspace,
                   size_t workspace_bytes, csb_stream_t stream);

/* The phase-2 tail as ONE entry point (cogreasoner_chat.py:413-422, :444-446, :458-474): per-frame mask rules (csb_mask_finalize),
 * token scan + compaction with the pooled token of the minor frames (csb_compact_tokens) and the id filter (csb_filter_ids without
 * position ids).  Two launches: one cooperative kernel whose token chunks are whole frames (so the per-frame rules are CTA-local)
 * and whose barrier word is carried from launch to launch (no memset), then the pooling + row-gather kernel.  Shapes the fused
 * kernel cannot take (P > 1024, more chunks than co-resident CTAs) run the three separate entry points -- same results.
 *   mask        [T*P] uint8, in: the raw csb_pixel_diff mask (or a final mask with finalize = 0), out: the final mask
 *   minor_flags [T] uint8 or NULL; pool_minor != 0 emits token 0 of a flagged frame as the frame's pooled mean
 *   frame_kind  [T] uint8 or NULL: multi-item inputs (:392-430), see the field
 *   input_ids   [n_ids] int64 or NULL (then no id filtering); n_out int64[2] = {ids kept, <image> ids seen}
 *   workspace   csb_compact_stream_workspace_bytes(T*P, n_ids) bytes, ZERO-FILLED ONCE by the caller before its first use and
 *               private to one CUDA stream at a time (it holds the carried barrier state). */
typedef struct csb_compact_args {
  const void* x_bf16;      /* [T*P, D] */
  int64_t T, P, D;
  uint8_t* mask;
  const uint8_t* minor_flags;
  const uint8_t* frame_kind; /* [T] uint8 or NULL: 1 = frame of an image item / single-frame video (cogreasoner_chat.py:401-403:
                                all tokens kept, no per-frame rule; only looked at with finalize != 0), 0 = video frame */
  int64_t min_tokens;
  int32_t finalize;
  int32_t pool_minor;
  const int64_t* input_ids;
  int64_t n_ids;
  int64_t image_token_index;
  const int64_t* attention_mask;
  const int64_t* labels;
  void* out_bf16;          /* [T*P, D], first n_kept rows valid */
  int32_t* kept_index;     /* [T*P] */
  int64_t* n_kept;         /* device int64 */
  int64_t* out_ids;
  int64_t* out_attention_mask;
  int64_t* out_labels;
  int64_t* n_out;          /* device int64[2] */
  void* workspace;
  size_t workspace_bytes;
} csb_compact_args;
CSB_API size_t csb_compact_stream_workspace_bytes(int64_t n_tokens, int64_t n_ids);
CSB_API int csb_compact_stream(const csb_compact_args* args /* host */, csb_stream_t stream);

/* The selection step of phase 2 as ONE launch (cogreasoner_chat.py:284 select_additional_frames, :317-318 event mean pool,
 * :325-333 cosine / threshold / minor list): k CTAs score the events against the question, ceil(k/8) CTAs pick the
 * representatives from the Gram matrix (csb_representatives_gram), the last CTA to finish builds the minor-frame list.
 * Event embeddings come either ready (`events` [k,Dl]) or as the packed last hidden states of the k summary forwards
 * (`hidden` [cu_seqlens[k], Dl] + cu_seqlens): then a split partial-sum pass runs first and every cosine CTA finishes the mean
 * of its own event in shared memory (the bits of csb_seq_mean_pool; the pooled rows are also stored to events_out if given).
 * Meant for events of up to a few hundred frames (T <= 256 k), like csb_representatives_gram. */
typedef struct csb_relevance_args {
  const void* question;        /* [Dl] */
  const void* events;          /* [k, Dl] or NULL */
  const void* hidden;          /* [cu_seqlens[k], Dl] or NULL */
  const int32_t* cu_seqlens;   /* [k + 1] with hidden */
  int64_t k, Dl;
  int32_t dtype;               /* CSB_BF16 / CSB_F32 */
  int32_t reserved;
  double threshold;
  const double* gram;          /* [T, T], leading dimension ldg */
  int64_t ldg;
  const int32_t* event_offsets;
  const int32_t* event_order;
  int64_t n_rep;
  const int64_t* assignments;  /* [T] */
  int64_t T;
  void* events_out;            /* [k, Dl] or NULL (with hidden) */
  void* cos_out;               /* [k] dtype */
  uint8_t* low_flags;          /* [k] */
  float* dist;                 /* [T] */
  int64_t* reps;               /* [k, n_rep] */
  int32_t* rep_counts;         /* [k] */
  uint8_t* rep_flags;          /* [T] */
  uint8_t* minor_flags;        /* [T] */
  int32_t* minor_list;         /* [T] */
  int64_t* n_minor;            /* device int64 */
  uint32_t* ticket;            /* zero-initialised, reset by the library (one per concurrently running stream) */
  void* workspace;             /* csb_relevance_select_workspace_bytes(k, Dl) with hidden, else may be NULL */
  size_t workspace_bytes;
} csb_relevance_args;
CSB_API size_t csb_relevance_select_workspace_bytes(int64_t k, int64_t Dl);
CSB_API int csb_relevance_select(const csb_relevance_args* args /* host */, csb_stream_t stream);

/* ---- between the two phases, and the last step of the call site ------------- */

CSB_API size_t csb_embed_scatter_workspace_bytes(int64_t n_ids);
/* cogreasoner_chat.py:303-305 (event-summary prompts) and :567-572 (final prompt), fused into one pass:
 *     inputs_embeds = embed_tokens(input_ids).clone()
 *     inputs_embeds[input_ids == image_token_index] = inputs_embeds[...] * 0.0 + features
 * out[i] = embed_table[input_ids[i]] for a text id, (embed_table[image_token_index] * 0.0 + features[j]) for the
 * j-th <image> id (both ops rounded in `dtype`, so NaN / signed-zero behaviour is the reference's).
 * input_ids may be several prompts packed back to back (the k event prompts): features are consumed in order,
 * which is the event-major `packed` buffer of csb_centroid_gather.
 *   embed_table [vocab, D], features [n_features, D], out [n_ids, D], all `dtype` (CSB_BF16 / CSB_F32), 16-byte
 *   aligned rows; status device int32[2]: {number of <image> ids, number of ids outside [0, vocab)} -- the caller
 *   checks status[0] == n_features (the reference raises a shape error otherwise) ; rows without a source are 0. */
CSB_API int csb_embed_scatter(const int64_t* input_ids, int64_t n_ids, int64_t image_token_index, const void* embed_table,
                      int64_t vocab, const void* features, int64_t n_features, int64_t D, int dtype, void* out,
                      int32_t* status, void* workspace, size_t workspace_bytes, csb_stream_t stream);
/* Same, reading the feature rows THROUGH the chronological per-event frame order (cogreasoner_chat.py:286-305 in one pass, with
 * no event-major copy of the features in between): the j-th <image> id takes feature row
 *     frame_order[j / tokens_per_frame] * tokens_per_frame + j % tokens_per_frame,
 * frame_order [n_frames] int32 being csb_kmeans_solve's event_order.  features [n_frames * tokens_per_frame, D]. */
CSB_API int csb_embed_scatter_ordered(const int64_t* input_ids, int64_t n_ids, int64_t image_token_index, const void* embed_table,
                              int64_t vocab, const void* features, const int32_t* frame_order, int64_t n_frames,
                              int64_t tokens_per_frame, int64_t D, int dtype, void* out, int32_t* status, void* workspace,
                              size_t workspace_bytes, csb_stream_t stream);

CSB_API size_t csb_seq_mean_pool_workspace_bytes(int64_t n_seq, int64_t Dl);
/* cogreasoner_chat.py:317-318 and :322: torch.mean(last_hidden_state, dim=1) of every sequence of a packed
 * var-len batch in one launch: out[j] = mean(hidden[cu_seqlens[j] : cu_seqlens[j+1]]) with fp32 accumulation in
 * row order, one division, one rounding to `dtype`.
 *   hidden [cu_seqlens[n_seq], Dl] `dtype`, cu_seqlens device int32 [n_seq + 1], out [n_seq, Dl] `dtype`. */
CSB_API int csb_seq_mean_pool(const void* hidden, const int32_t* cu_seqlens, int64_t n_seq, int64_t Dl, int dtype, void* out,
                      void* workspace, size_t workspace_bytes, csb_stream_t stream);

/* ---- upstream of the path: the vision tower's tail ------------------------------------------------------
 * model/modeling_videollama3_encoder.py:479 (post_layernorm) + :485-502 (view / permute / reshape / permute /
 * F.interpolate(size = grid / merge, mode='bilinear') / permute / view) in ONE pass: the encoder's last hidden states arrive
 * merge-block-contiguous (the merge^2 patches of a merged token are consecutive rows), so the chain is "LayerNorm the rows of a
 * block, blend them" -- for even merge the four central patches with weights 1/4, for odd merge the central patch alone (what
 * align_corners=False bilinear sampling at scale = merge evaluates).  ATen's op order: LayerNorm in fp32 with biased variance,
 * rounded to bf16; blend in fp32, rounded to bf16.  Item-independent (valid across items that share `merge`).
 *   hidden [n_rows, C] bf16, gamma / beta [C] bf16, out [n_rows / merge^2, C] bf16; C % 8 == 0, C <= 2048. */
CSB_API int csb_ln_merge(const void* hidden_bf16, int64_t n_rows, int64_t C, int64_t merge, const void* gamma_bf16,
                         const void* beta_bf16, double eps, void* out_bf16, csb_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* COGSTREAM_B200_H_ */
