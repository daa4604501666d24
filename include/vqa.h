/*
 * vqa.h -- C ABI of the B200-native dense-retrieval engine (libvqa_b200.so).
 *
 * This is the drop-in boundary for the ONE hot path of
 * vTuanpham/Vietnamese_QA_System that this repository accelerates: scoring
 * query embeddings against the document-embedding index and returning the top-k
 * passages (reference call sites: inference_pipeline/db_utils/heavy_ranker.py:78-101).
 *
 * The reference is 100 % Python and has NO native interface of its own; the
 * arithmetic it calls lives in third-party txtai -> faiss-cpu.  Each entry point
 * below therefore cites the reference line (or the txtai backend method called
 * from that line) whose work it replaces.  A Python maintainer binds these with
 * ctypes (INTEGRATION.md shows the stub); nothing in the signatures is a torch
 * type -- plain pointers, sizes and a CUDA stream handle passed as void*.
 *
 * Conventions
 *   - every function returns a vqa_status (0 = OK, negative = error) and never
 *     throws or aborts; vqa_last_error() returns a thread-local message.
 *   - all *_dev pointers are device pointers on the index's device; the engine
 *     BORROWS them (the caller -- PyTorch in the shipped host layer -- owns all
 *     device memory).  All device work is stream-ordered on `stream`.
 *   - vqa_search performs no allocation and no host synchronisation: the caller
 *     passes a workspace of vqa_workspace_bytes() bytes, so a search is CUDA-graph
 *     capturable.
 *   - ordering everywhere: score descending, ties -> lower doc id
 *     (BASELINE.json north_star; txtai NumPy backend's stable sort).
 *   - there is NO CPU fallback: without a CUDA device every compute entry point
 *     returns VQA_E_CUDA.
 *   - threading: an index handle is immutable between vqa_index_bind /
 *     vqa_index_set_tuning calls, and any number of threads may call vqa_search*
 *     on it concurrently PROVIDED each in-flight call has its own workspace (or
 *     staging buffer) and output buffers -- two searches that share a workspace
 *     race on its candidate lists, whatever streams they run on.  bind /
 *     set_tuning / destroy must not overlap any other call on the same handle.
 *   - the search path reads NO environment variables: kernel-selection knobs are
 *     a vqa_tuning_t stored in the handle (library defaults, overridden by VQA_*
 *     variables read ONCE in vqa_index_create, or set explicitly with
 *     vqa_index_set_tuning).
 */
#ifndef VQA_H_
#define VQA_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(_WIN32)
#define VQA_API __declspec(dllexport)
#else
#define VQA_API __attribute__((visibility("default")))
#endif

#define VQA_VERSION 121 /* 111: + vqa_plan_describe, vqa_search_host_async; 120: + vqa_tuning_t (no getenv on the search path); 121: + vqa_merge_segments (k <= 1024) */

typedef enum vqa_status {
    VQA_OK = 0,
    VQA_E_INVALID = -1,     /* bad argument (shape, dtype, alignment, k)        */
    VQA_E_CUDA = -2,        /* CUDA runtime / driver error, or no device        */
    VQA_E_UNSUPPORTED = -3, /* valid request the engine does not implement yet  */
    VQA_E_NOMEM = -4        /* workspace too small / host allocation failure    */
} vqa_status;

typedef enum vqa_dtype {
    VQA_F32 = 0,
    VQA_BF16 = 1,
    VQA_F16 = 2,
    VQA_I64 = 3, /* masks only */
    VQA_I32 = 4, /* masks only */
    VQA_U8 = 5   /* masks only (bool) */
} vqa_dtype;

typedef enum vqa_mode {
    /* fp32 arithmetic in the canonical FMA / butterfly order:
     * ids bit-identical to the CPU oracle, works for fp32 / bf16 / fp16 rows. */
    VQA_MODE_VERIFY = 0,
    /* free summation order; the engine picks the kernel family by row type, batch and k: for 16-bit rows with
     * dim % 64 == 0 one of the tcgen05 kernels (queries in shared memory up to 32, in tensor memory beyond, CTA pairs
     * beyond 128), otherwise the HBM-streaming CUDA-core kernel (vqa_search_plan). */
    VQA_MODE_FAST = 1,
    /* FAST, but force one kernel family (benchmarks / tests). */
    VQA_MODE_FAST_STREAM = 2,
    VQA_MODE_FAST_TENSOR = 3,
    /* tcgen05 with the query block resident in tensor memory (large batches, dim <= 1024) */
    VQA_MODE_FAST_TS = 4,
    /* the same with CTA pairs: tcgen05.mma.cta_group::2, 256 queries per pair (tensor-bound regime, B > 128) */
    VQA_MODE_FAST_PAIR = 5
} vqa_mode;

typedef struct vqa_index vqa_index_t; /* opaque */

/* Library version (VQA_VERSION). */
VQA_API int vqa_version(void);

/* Thread-local message for the last non-OK status returned on this thread. */
VQA_API const char *vqa_last_error(void);

/* Number of visible CUDA devices (0 when there is no driver / GPU).  Never fails. */
VQA_API int vqa_device_count(void);

/*
 * Create an index descriptor for one row shard.
 * Replaces: txtai ANN backend construction under Embeddings.index()/load()
 *           (heavy_ranker.py:86,88,91-94).
 *   n_rows           rows (documents) in THIS shard
 *   dim              embedding dimension; dim * sizeof(dtype) must be a multiple of 16
 *   dtype            storage type of the rows: VQA_F32 | VQA_BF16 | VQA_F16
 *   device           CUDA device ordinal holding the rows
 *   first_global_id  position of this shard's row 0 in the whole index
 *                    (rank r of G holds rows [r*ceil(N/G), ...))
 */
VQA_API int vqa_index_create(vqa_index_t **out, int64_t n_rows, int32_t dim, int32_t dtype,
                             int32_t device, int64_t first_global_id);

/*
 * Bind (borrow) the caller's device buffer holding the shard, row-major,
 * `row_stride_bytes` between rows (>= dim*sizeof(dtype), multiple of 16, base
 * 16-byte aligned).  Builds the TMA tensor map used by the tensor-core kernel.
 * Replaces: faiss add_with_ids under Embeddings.index() (heavy_ranker.py:86,88).
 */
VQA_API int vqa_index_bind(vqa_index_t *h, const void *rows_dev, int64_t n_rows,
                           int64_t row_stride_bytes);

VQA_API int vqa_index_destroy(vqa_index_t *h);

/*
 * Diagnostic: while `stamps_dev` is set (NULL switches it off), every CTA of the smem-resident tcgen05 scan
 * kernel writes 32 uint64 stamps -- %globaltimer (ns) in slots 0..15, clock64 in 16..31 -- of: 0 entry,
 * 1 first / 2 last TMA issue, 3 first MMA / 4 last commit, 5 queries staged, 6..12 tiles 0, 1, 3, 7, 15, 31, 63
 * leaving the epilogue, 13 last tile, 14 exit.  The buffer must hold 256 bytes per SM.  This is how
 * profiles/r2_timeline_*.json (where the per-launch fixed cost goes) were produced; no effect on results.
 */
VQA_API int vqa_debug_timeline(vqa_index_t *h, void *stamps_dev, size_t bytes);

/*
 * Kernel-selection knobs of one index handle.  Every field has a measured default (vqa_tuning_default);
 * they exist for benchmarks and tests -- a drop-in user never touches them.  -1 / 0 = "auto" where noted.
 * Replaces: nothing in the reference (txtai exposes faiss' `nprobe`/`components` strings the same way,
 * through Embeddings(**cfg), heavy_ranker.py:78-83).
 */
typedef struct vqa_tuning {
    int32_t size;          /* sizeof(vqa_tuning_t) as the caller compiled it (ABI growth check)                  */
    int32_t ts_extra;      /* spare candidate ranks kept by the screen-then-rescore scans, 0..96 (6)              */
    int32_t ss_screen;     /* smem-resident tcgen05 kernel: screen mode instead of hi/lo columns, 0|1, -1 auto =
                              only for scans of >= 12 GB with more than 16 queries (-1)                             */
    int32_t mma_kps;       /* 64-column blocks per TMA ring stage, 0 = auto, else 1..16                           */
    int32_t mma_stages;    /* cap on ring stages, 0 = auto                                                        */
    int32_t mma_groups;    /* smem-resident kernel: query chunks side by side per launch, 1..4 (4)                */
    int32_t mma_multicast; /* those chunks as a cluster with TMA multicast, 0|1 (1)                               */
    int32_t ts_qs;         /* TMEM-resident-query kernel: QS variant (part of the query block in smem), 0|1 (1)   */
    int32_t ts_ks;         /* QS: 64-column query blocks kept in shared memory, -1 = auto, else 0..16             */
    int32_t ts_split;      /* TS kernel: hi+lo rows (1) or storage-precision screen (0), -1 = auto                */
    int32_t ts_groups;     /* TS kernel: chunks of 128 queries per launch (cluster size), 1..4 (2)                */
    int32_t reduce_select; /* radix-select candidate reduce for k > 32 and for re-scoring reduces, 0|1 (1)        */
    int32_t reduce_early;  /* early exit in the k <= 32 warp reduce over sorted internal lists, 0|1 (1)           */
    int32_t pdl_chain;     /* 2nd+ scan launch of one search overlaps the previous reduce, 0|1                    */
    int32_t tma_l2promo;   /* CUtensorMapL2promotion of the document tensor map, 0..3 (3 = 256 B)                 */
    int32_t tma_hint;      /* L2 policy of the document stream: 0 normal, 1 evict-first, 2 evict-last (1)         */
    int32_t stream_max_b;  /* FAST: batches up to this size take the CUDA-core streaming kernel, 0..8 (0) ...     */
    int32_t stream_min_mb; /* ... when the shard is at least this many MB (its fixed cost is ~80 us higher), (8000)*/
    int32_t pair;          /* FAST: batches of > 128 queries take the cta_group::2 CTA-pair kernel, 0|1 (1)       */
    int32_t dyn_tiles;     /* smem-resident kernel without clusters: CTAs take tiles from a shared counter, 0|1 (1)*/
    int32_t seed;          /* smem-resident kernel, register lists: warm-up bound from the first tiles' maxima, 0|1 (1) */
    int32_t wide;          /* FAST: 33..128 queries (dim <= 768) on single-CTA 128-document tiles (pair.cuh), 0|1 (1) */
    int32_t ts_m64;        /* TS kernel, <= 64 queries, screen mode: M = 64 instructions, 0|1 (1)                   */
    int32_t smem_reserve_kb; /* shared memory per SM the TS / 128-document-tile scans leave free so that the re-scoring
                                reduce of the previous batch can run NEXT to them in a pipelined loop, 0..64 (0)    */
} vqa_tuning_t;

/* Library defaults (no environment). */
VQA_API int vqa_tuning_default(vqa_tuning_t *t);
/* Defaults overridden by the VQA_* environment variables of the same names (VQA_TS_EXTRA, VQA_TS_QS ...);
 * a value that does not parse or is out of range is VQA_E_INVALID.  vqa_index_create calls this once. */
VQA_API int vqa_tuning_from_env(vqa_tuning_t *t);
/* Validate and store `t` in the handle (rebuilding the tensor maps if tma_l2promo changed). */
VQA_API int vqa_index_set_tuning(vqa_index_t *h, const vqa_tuning_t *t);
VQA_API int vqa_index_get_tuning(const vqa_index_t *h, vqa_tuning_t *t);

/* Bytes of device workspace vqa_search needs for a batch of `n_queries`, top `k`.
 * Zero the workspace once after allocating it (cudaMemset); the library keeps it consistent from then on.  A
 * workspace that was never zeroed still gives correct results -- stale words are recognised by their epoch tags --
 * but its first search pays a one-off ~0.3 ms repair of the tile counter (a CAS that 148 CTAs contend for). */
VQA_API int vqa_workspace_bytes(const vqa_index_t *h, int32_t n_queries, int32_t k, int32_t mode,
                                size_t *bytes);

/*
 * Score `n_queries` query embeddings against the shard and select the top k.
 * Replaces: txtai ann.search(queries, limit) -> faiss index.search under
 *           Embeddings.search()/batchsearch() (heavy_ranker.py:98,100).
 *   queries_dev   float32 [n_queries, dim] row-major, already L2-normalised
 *                 (stride q_stride elements between rows)
 *   k             1..128
 *   out_scores_dev float32 [n_queries, k]  descending
 *   out_ids_dev    int64   [n_queries, k]  first_global_id + row position;
 *                  slots beyond the shard's row count: score -inf, id -1
 * The [n_queries, n_rows] score matrix is never written to memory.
 */
VQA_API int vqa_search(const vqa_index_t *h, const float *queries_dev, int64_t q_stride,
                       int32_t n_queries, int32_t k, int32_t mode, float *out_scores_dev,
                       int64_t *out_ids_dev, void *workspace_dev, size_t workspace_bytes,
                       void *stream);

/*
 * The same search on TWO streams: the scan on `scan_stream`, the candidate reduce (the kernel that writes
 * out_scores_dev / out_ids_dev) on `reduce_stream` behind an event the library records after the scan.  The scan
 * stream is then free for the NEXT search's scan the moment this scan ends -- in a loop over independent batches
 * the reduce (and whatever the caller orders behind it on `reduce_stream`: the cross-GPU exchange, the merge, the
 * D2H copy) overlaps the next scan.  Consecutive searches in flight must use distinct workspaces and outputs.
 * Replaces: nothing in the reference; it is how ShardedFlat.search_pipelined keeps the scans back to back.
 */
VQA_API int vqa_search_2s(const vqa_index_t *h, const float *queries_dev, int64_t q_stride,
                          int32_t n_queries, int32_t k, int32_t mode, float *out_scores_dev,
                          int64_t *out_ids_dev, void *workspace_dev, size_t workspace_bytes,
                          void *scan_stream, void *reduce_stream);

/*
 * Same search with HOST buffers: H2D copy of the queries, search, D2H copy of
 * the results, stream-synchronised before returning.  `staging_dev` must hold
 * vqa_search_host_staging_bytes() bytes (it contains the search workspace: zero it once after allocation).  This is the call the reference-facing
 * plugin makes for numpy inputs (heavy_ranker.py:98-101 receives host lists).
 */
VQA_API int vqa_search_host_staging_bytes(const vqa_index_t *h, int32_t n_queries, int32_t k,
                                          int32_t mode, size_t *bytes);
VQA_API int vqa_search_host(const vqa_index_t *h, const float *queries_host, int32_t n_queries,
                            int32_t k, int32_t mode, float *out_scores_host,
                            int64_t *out_ids_host, void *staging_dev, size_t staging_bytes,
                            void *stream);
/* The same three copies and launches WITHOUT the final synchronisation: the call returns as soon as the
 * work is enqueued, and the host buffers (which must be pinned for the copies to be asynchronous) hold the
 * result once `stream` -- or an event recorded on it after the call -- has completed.  Lets a serving loop
 * overlap its own post-processing (id -> passage fetch) and the next batch's enqueue with the search; each
 * in-flight call needs its own staging and output buffers. */
VQA_API int vqa_search_host_async(const vqa_index_t *h, const float *queries_host, int32_t n_queries,
                                  int32_t k, int32_t mode, float *out_scores_host,
                                  int64_t *out_ids_host, void *staging_dev, size_t staging_bytes,
                                  void *stream);

/*
 * Merge `n_lists` candidate lists per query (the row shards' results after the
 * all-gather) into the global top k_out.  Padding entries have id < 0.
 * Replaces: nothing in the reference (single process); it is the exchange step
 *           of the row-sharded engine.
 *   cand_*_dev layout [n_lists, n_queries, k_in]
 */
VQA_API int vqa_merge_topk(const float *cand_scores_dev, const int64_t *cand_ids_dev,
                           int32_t n_lists, int32_t n_queries, int32_t k_in, int32_t k_out,
                           float *out_scores_dev, int64_t *out_ids_dev, int32_t device,
                           void *stream);

/*
 * Same merge over a strided layout: list l's scores start at cand_scores_dev + l*list_stride_scores
 * (float elements) and its ids at cand_ids_dev + l*list_stride_ids (int64 elements).  This is the
 * layout one all-gather of each rank's packed [scores | ids] result block produces, so the exchange
 * step needs no repacking kernel.
 */
VQA_API int vqa_merge_topk_strided(const float *cand_scores_dev, const int64_t *cand_ids_dev,
                                   int64_t list_stride_scores, int64_t list_stride_ids, int32_t n_lists,
                                   int32_t n_queries, int32_t k_in, int32_t k_out, float *out_scores_dev,
                                   int64_t *out_ids_dev, int32_t device, void *stream);

/*
 * Exchange step over NVLink / NVSwitch PEER MEMORY instead of NCCL (optional; needs buffers mapped on
 * every rank, e.g. torch symmetric memory).  vqa_exchange_push copies this rank's packed result block
 * into its slot of every peer's gather buffer with peer stores and then publishes `epoch` in that
 * peer's flag (release, system scope); vqa_merge_topk_wait is vqa_merge_topk_strided whose kernel
 * first acquires all n_lists flags (>= epoch).  peer_*_ptrs are HOST arrays of `world` device
 * pointers (peer r's slot for this rank / peer r's flag for this rank); flags_dev is this rank's own
 * flag array, one per rank.
 */
VQA_API int vqa_exchange_push(const void *local_block_dev, size_t block_bytes, void *const *peer_slot_ptrs,
                              uint64_t *const *peer_flag_ptrs, int32_t world, uint64_t epoch, int32_t device,
                              void *stream);
VQA_API int vqa_merge_topk_wait(const float *cand_scores_dev, const int64_t *cand_ids_dev,
                                int64_t list_stride_scores, int64_t list_stride_ids, int32_t n_lists,
                                int32_t n_queries, int32_t k_in, int32_t k_out, float *out_scores_dev,
                                int64_t *out_ids_dev, const uint64_t *flags_dev, uint64_t epoch,
                                int32_t device, void *stream);

/*
 * Top k for k beyond one search call's limit (128 < k <= 1024; the hybrid search of txtai fetches 10 x limit
 * dense candidates, heavy_ranker.py:98,100 with limit > 12).  The caller cuts the shard into row segments in
 * ascending row order (one vqa_index handle per segment over a slice of the same rows), runs vqa_search on each
 * for its own top k_seg, and this call sorts the n_segments x k_seg survivors of every query and writes the best
 * k_out: score descending (-0 == +0), ties -> the lower position.
 *   seg_scores_dev float32 / seg_ids_dev int64 [n_segments, n_queries, k_seg]   each list as vqa_search wrote it
 *                 (unused slots -inf / -1); n_segments * k_seg <= 8192
 *   saturated_dev int32 [n_segments]   out: non-zero when the segment's LAST kept candidate is inside some query's
 *                 answer, i.e. the segment may hold more of that query's top k_out than k_seg.  The answer is exact
 *                 when no segment with more than k_seg rows is saturated; otherwise the caller halves the saturated
 *                 segments, searches the halves and merges again (vietnamese_qa_system_b200/ops.py does).
 * vqa_merge_segments_limits reports the largest k_out and n_segments * k_seg.
 * Replaces: nothing in the reference (faiss answers any k in one call); it is how this engine composes k > 128.
 */
VQA_API int vqa_merge_segments_limits(int32_t *max_k_out, int32_t *max_candidates);
VQA_API int vqa_merge_segments(const float *seg_scores_dev, const int64_t *seg_ids_dev, int32_t n_segments,
                               int32_t n_queries, int32_t k_seg, int32_t k_out, float *out_scores_dev,
                               int64_t *out_ids_dev, int32_t *saturated_dev, int32_t device, void *stream);

/*
 * Fused masked mean-pool (+ optional L2 normalise) over encoder hidden states:
 *   e[b,:] = sum_s h[b,s,:]*m[b,s] / max(sum_s m[b,s], 1e-9);  e /= ||e||_2
 * Replaces: txtai MeanPooling.forward + normalize under Embeddings.search()/
 *           index() (heavy_ranker.py:86,88,98,100; in-tree twin src/test.py:97-99).
 *   hidden_dev [B,S,D] h_dtype (F32|BF16|F16), contiguous;  mask_dev [B,S]
 *   m_dtype (I64|I32|U8|F32);  out_dev [B,D] float32.
 */
VQA_API int vqa_pool_normalize(const void *hidden_dev, int32_t h_dtype, const void *mask_dev,
                               int32_t m_dtype, int32_t batch, int32_t seq, int32_t dim,
                               int32_t normalize, float *out_dev, int32_t device, void *stream);

/*
 * Row-wise L2 normalise float32 [n,dim] (in place allowed) and optionally cast
 * into a second buffer of `cast_dtype` (the index storage type).  Zero rows stay 0.
 * Replaces: txtai normalize (numpy, in place) under Embeddings.index()/search().
 *   cast_out_dev may be NULL; row strides in elements.
 */
VQA_API int vqa_normalize_rows(const float *in_dev, int64_t in_stride, int64_t n_rows, int32_t dim,
                               float *out_dev, int64_t out_stride, void *cast_out_dev,
                               int32_t cast_dtype, int64_t cast_stride, int32_t device,
                               void *stream);

/*
 * Two-index agreement rule, batched (heavy_ranker.py:110):
 *   accept[i] = ids_a[i] == ids_b[i] && scores_a[i] + scores_b[i] > threshold
 * combined[i] = scores_a[i] + scores_b[i].  All device pointers, length n.
 */
VQA_API int vqa_agree(const int64_t *ids_a_dev, const float *scores_a_dev,
                      const int64_t *ids_b_dev, const float *scores_b_dev, int64_t n,
                      double threshold, uint8_t *accept_dev, float *combined_dev, int32_t device,
                      void *stream);

/* Introspection used by bench/tests: which kernel family a FAST search would use
 * (VQA_MODE_FAST_STREAM or VQA_MODE_FAST_TENSOR), how many kernels one search launches. */
VQA_API int vqa_search_plan(const vqa_index_t *h, int32_t n_queries, int32_t k, int32_t mode,
                            int32_t *family, int32_t *n_launches);

/* The same planning WITHOUT a device or a bound index (host arithmetic only): what vqa_search would do
 * for an index of n_rows x dim `dtype` rows on a GPU with `sm_count` SMs and `max_smem` bytes of opt-in
 * dynamic shared memory per CTA (B200: 148, 232448).  Lets the CPU test suite check the routing and that
 * every planned launch fits shared and tensor memory.  out[16] (int32): 0 family, 1 queries per pass,
 * 2 passes, 3 side-by-side groups, 4 ring stages, 5 k-blocks per stage, 6 MMA N (tensor family),
 * 7 hi/lo split (tensor: ss_split, TS: ts_split), 8 QS variant, 9 query blocks in shared memory,
 * 10 list length inside the scan, 11 candidates kept by the reduce (k_out), 12 reduce re-scores (0/1),
 * 13 TMEM columns used, 14 M = 64 variant (0/1), 15 reserved; *smem_bytes = dynamic shared memory of the scan kernel. */
VQA_API int vqa_plan_describe(int64_t n_rows, int32_t dim, int32_t dtype, int32_t n_queries, int32_t k,
                              int32_t mode, int32_t sm_count, int32_t max_smem, int32_t *out,
                              size_t *smem_bytes);
/* The same with explicit knobs (tuning == NULL: vqa_tuning_from_env, as vqa_plan_describe). */
VQA_API int vqa_plan_describe_tuned(int64_t n_rows, int32_t dim, int32_t dtype, int32_t n_queries, int32_t k,
                                    int32_t mode, int32_t sm_count, int32_t max_smem,
                                    const vqa_tuning_t *tuning, int32_t *out, size_t *smem_bytes);

/* ------------------------------------------------------------------------------------------------
 * Sparse (BM25) leg and hybrid fusion -- SURVEY.md 8(f) rank 3.
 * The reference builds its indexes with txtai.Embeddings(hybrid=True, ...)
 * (inference_pipeline/db_utils/heavy_ranker.py:78-83): next to the dense index txtai keeps a BM25
 * term index (scoring = {"method": "bm25", "terms": True, "normalize": True}), asks each leg for
 * 10 x limit candidates at :98,100 and adds the scores per id with weights [0.5, 0.5].
 * ---------------------------------------------------------------------------------------------- */

typedef struct vqa_sparse vqa_sparse_t; /* opaque */

/* Compile-time limits: distinct known terms per query, candidates kept per query. */
VQA_API int vqa_sparse_limits(int32_t *max_query_terms, int32_t *max_candidates);

/*
 * Term-index descriptor over document positions [0, n_docs): CSR postings
 *   offsets int64[n_terms + 1], docs int32[n_postings] (ascending within a term),
 *   weights float32[n_postings] (BM25 weight of the (term, document) pair).
 * Replaces: txtai scoring.Terms (the sqlite-backed term index built under Embeddings.index(),
 *           heavy_ranker.py:86,88).  The buffers are BORROWED.
 */
VQA_API int vqa_sparse_create(vqa_sparse_t **out, int64_t n_docs, int64_t n_terms, int64_t n_postings,
                              int32_t device);
VQA_API int vqa_sparse_bind(vqa_sparse_t *h, const int64_t *offsets_dev, const int32_t *docs_dev,
                            const float *weights_dev);
VQA_API int vqa_sparse_destroy(vqa_sparse_t *h);

/*
 * Build side: BM25 weight of every posting, evaluated in float64 in txtai's operation order
 *   k = k1 * ((1 - b) + b * doc_len / avgdl);  w = idf * (freq * (k1 + 1)) / (freq + k)
 * and rounded once to float32.  Replaces: txtai BM25.score under Terms.weights.
 *   idf_dev float64[n_terms], doc_len_dev int32[n_docs], freqs_dev int32[n_postings].
 */
VQA_API int vqa_bm25_weights(const int64_t *offsets_dev, int64_t n_terms, const int32_t *docs_dev,
                             const int32_t *freqs_dev, int64_t n_postings, const double *idf_dev,
                             const int32_t *doc_len_dev, double k1, double b, double avgdl,
                             float *weights_out_dev, int32_t device, void *stream);

VQA_API int vqa_sparse_workspace_bytes(const vqa_sparse_t *h, int32_t n_queries, int32_t k_cand_max,
                                       size_t *bytes);

/*
 * Score n_queries tokenised queries against the term index and select the best `limit` documents.
 * Replaces: txtai Terms.search + TFIDF.search (score normalisation) under Embeddings.search()
 *           (heavy_ranker.py:98,100).
 *   q_terms_dev  int32 [n_queries, max_terms]  term ids; per query first the n_rare terms that are
 *                accumulated over all their documents (in the query's first-occurrence order), then
 *                the n_common terms (document frequency > cutoff * n_docs) that are merged only
 *                into the surviving candidates.  Ids outside [0, n_terms) count as unknown terms.
 *   q_freqs_dev  float32 [n_queries, max_terms] occurrences of the term in the query
 *   q_meta_dev   int32 [n_queries, 4]           n_rare, n_common, k_cand (candidates kept before
 *                the common-term merge: limit, or 5 x limit when n_common > 0), unused
 *   out_scores_dev float64 [n_queries, limit]   descending; normalize != 0 applies
 *                min(score / min(top + avgscore, 6 * avgscore), 1); documents with score 0 are
 *                never returned: unused slots hold score -inf, id -1
 *   out_ids_dev  int64 [n_queries, limit]       document positions
 * Accumulation is fp32 multiply-then-add in term order (no fused multiply-add), so scores are
 * bit-identical to the CPU accumulation they replace.  No allocation, no host synchronisation.
 */
VQA_API int vqa_sparse_search(const vqa_sparse_t *h, const int32_t *q_terms_dev, const float *q_freqs_dev,
                              const int32_t *q_meta_dev, int32_t max_terms, int32_t n_queries,
                              int32_t k_cand_max, int32_t limit, int32_t normalize, double avgscore,
                              double *out_scores_dev, int64_t *out_ids_dev, void *workspace_dev,
                              size_t workspace_bytes, void *stream);

/*
 * Hybrid fusion of the two legs' candidate lists (ids < 0 are padding):
 *   fused[id] = 0.0 + dense * w_dense (+ sparse * w_sparse), float64 like the Python floats it replaces;
 *   order = fused descending, ties keep insertion order (dense candidates first, then sparse-only).
 * Replaces: txtai Search.search's union/sort under Embeddings.search() with hybrid=True.
 *   out_scores_dev float64 [n_queries, limit], out_ids_dev int64 [n_queries, limit]; unused slots -inf / -1.
 */
VQA_API int vqa_hybrid_fuse(const float *dense_scores_dev, const int64_t *dense_ids_dev, int32_t k_dense,
                            const double *sparse_scores_dev, const int64_t *sparse_ids_dev, int32_t k_sparse,
                            int32_t n_queries, double w_dense, double w_sparse, int32_t limit,
                            double *out_scores_dev, int64_t *out_ids_dev, int32_t device, void *stream);

/*
 * Reciprocal-rank fusion of the same two candidate lists: fused[id] = 0.0 + (1.0 / (rank_dense + 1)) * w_dense
 * (+ (1.0 / (rank_sparse + 1)) * w_sparse), rank = position in the leg's own list.  What txtai's Search does
 * instead of the weighted score sum when the scoring index is not normalised (raw BM25 scores are unbounded).
 * In both fusions a leg whose weight is <= 0 is ignored (txtai: `scores if weights[v] > 0 else []`).
 */
VQA_API int vqa_hybrid_fuse_rrf(const int64_t *dense_ids_dev, int32_t k_dense, const int64_t *sparse_ids_dev,
                                int32_t k_sparse, int32_t n_queries, double w_dense, double w_sparse,
                                int32_t limit, double *out_scores_dev, int64_t *out_ids_dev, int32_t device,
                                void *stream);

/* The agreement rule (heavy_ranker.py:110) on float64 scores -- what hybrid indexes return. */
VQA_API int vqa_agree_f64(const int64_t *ids_a_dev, const double *scores_a_dev, const int64_t *ids_b_dev,
                          const double *scores_b_dev, int64_t n, double threshold, uint8_t *accept_dev,
                          double *combined_dev, int32_t device, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* VQA_H_ */
