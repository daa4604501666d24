// launch.h -- host-side launch interface between api.cu and the kernel
// translation units (one per storage type so they compile in parallel).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "consts.h"

namespace vqa {

struct ScanLaunch {
    int dtype;  // vqa_dtype of the rows
    int bt;     // queries per pass: 1, 2, 4, 8
    int grid;
    const void *rows;
    long long n_rows;
    long long row_stride_bytes;
    int dim;
    const float *q;
    long long q_stride;
    int nq;
    int k;
    float *cand_s;
    uint32_t *cand_i;
    long long cand_stride;
};

struct MmaLaunch {
    const CUtensorMap *tmap;
    bool bf16;
    int ncol;
    int split;   // 1: hi/lo column pairs; 0: screen mode (one storage-precision column per query)
    int stages;  // ring stages, each kps x 16 KB
    int kps;
    int grid;
    int n_groups;  // query chunks of ncol/2 handled side by side in one launch (grid % n_groups == 0)
    int multicast; // 1: launch as clusters of n_groups CTAs with TMA multicast (tmap box = 128/n_groups rows)
    const float *q;
    long long q_stride;
    int nq;
    int k;
    long long n_rows;
    int dim;
    float *cand_s;
    uint32_t *cand_i;
    long long cand_stride;
    unsigned long long *tau_g;  // [nq] shared thresholds for this pass (or nullptr)
    uint32_t epoch;
    int pdl = 0;  // opt-in: launch with programmatic stream serialization (2nd+ scan of one search, see mma_launch.cu)
    int tma_hint = 1;  // L2 policy of the document stream: 0 normal, 1 evict-first, 2 evict-last
    unsigned long long *timeline = nullptr;  // diagnostic per-CTA stamps (vqa_debug_timeline), normally nullptr
    unsigned long long *tile_ctr = nullptr;  // dynamic tile schedule (launches without clusters), see MmaParams
    unsigned long long *slot_g = nullptr;    // warm-up seed slots [nq][32] (register-list path), see MmaParams
};

struct TsLaunch {
    const CUtensorMap *tmap;  // box = 64 columns x (64 / cluster) rows
    bool bf16;
    int split;    // 1: 64 queries per CTA as hi + lo rows; 0: 128 queries per CTA, storage-precision queries
    int a_fp16;   // queries as fp16 against bf16 documents
    int stages;
    int kps;
    int grid;
    int n_groups;
    int multicast;
    const float *q;
    long long q_stride;
    int nq;
    int k;
    long long n_rows;
    int dim;
    float *cand_s;
    uint32_t *cand_i;
    long long cand_stride;
    unsigned long long *tau_g;
    uint32_t epoch;
    int pdl = 0;  // opt-in: launch with programmatic stream serialization (2nd+ scan of one search)
    int qs = 0;   // 1: the QS kernel variant (part of the query block in shared memory); opt-in, see ts.cuh
    int ks = 0;   // QS: 64-column blocks of the query block kept in shared memory
    int m64 = 0;  // QS, <= 64 queries, no hi/lo rows: M = 64 instructions (half the tensor-memory A read per MMA)
    unsigned long long *timeline = nullptr;  // diagnostic per-CTA counters (vqa_debug_timeline), normally nullptr
};

// CTA-pair kernel (pair.cuh): cta_group::2 MMAs, 256 queries per pair, screen mode only
struct PairLaunch {
    const CUtensorMap *tmap;  // box = 64 columns x 64 rows
    bool bf16;
    int stages;
    int kps;
    int grid;                 // 2 x pairs (pair = 1) or CTAs (pair = 0)
    // stages x kps = 64-column blocks of a TILE in the ring (8 KB per CTA each with pair = 1, 16 KB with pair = 0)
    const float *q;
    long long q_stride;
    int nq;                   // <= 256
    int k;                    // list length inside the scan (<= 32)
    long long n_rows;
    int dim;
    float *cand_s;
    uint32_t *cand_i;
    long long cand_stride;
    unsigned long long *tau_g;
    uint32_t epoch;
    int ks;                   // 64-column query blocks kept in shared memory
    int pair = 1;             // 1: CTA pairs (grid = 2 x pairs, <= 256 queries); 0: one CTA per tile stream (<= 128 queries)
    unsigned long long *timeline = nullptr;
};
cudaError_t launch_pair(const PairLaunch &a, cudaStream_t st);
size_t pair_smem_bytes(int boxes, int ks);

cudaError_t launch_ts(const TsLaunch &a, cudaStream_t st);
size_t ts_smem_bytes(int k, int boxes, int split, int ks = 0, int nq = 1 << 30, int qs = 0, int m64 = 0);
cudaError_t launch_scan(const ScanLaunch &a, cudaStream_t st);
cudaError_t launch_scan_f32(const ScanLaunch &a, cudaStream_t st);
cudaError_t launch_scan_bf16(const ScanLaunch &a, cudaStream_t st);
cudaError_t launch_scan_f16(const ScanLaunch &a, cudaStream_t st);
cudaError_t launch_mma(const MmaLaunch &a, cudaStream_t st);
// clusters of `cluster` CTAs of the tensor-core kernel that can be co-resident (0 if the query fails)
int mma_max_active_clusters(bool bf16, int ncol, int split, int cluster, size_t smem_bytes);

// optional exact re-scoring stage of the candidate reduce (see ReduceParams in scan.cuh)
struct Rescore {
    const void *rows = nullptr;
    long long stride = 0;
    int dim = 0;
    int bf16 = 1;
    const float *q = nullptr;
    long long q_stride = 0;
    int k_final = 0;
};

// kernel-selection knobs of the candidate reduce (from the handle's vqa_tuning_t; never from the environment)
struct ReduceOpts {
    int select = 0;         // radix-select kernel for k_out > 32 and for re-scoring reduces
    int early = 0;          // early exit over sorted internal lists (k_out <= 32 warp kernel)
    int trigger_early = 0;  // release PDL dependents at once (the next scan of the same search does not read our output)
    int no_pdl = 0;         // plain launch (the reduce runs on another stream than its scan, behind an event)
};

cudaError_t launch_reduce_u32(const float *cand_s, const uint32_t *cand_i, long long list_stride,
                              long long query_stride, int n_lists, int k_in, int k_out, long long id_base,
                              float *out_s, long long *out_i, int n_queries, unsigned long long *tau_g_reset,
                              int list_mod, int queries_per_group, cudaStream_t st, const ReduceOpts &opts,
                              const Rescore *rs = nullptr, unsigned long long *slot_reset = nullptr);
struct WaitFlags {
    const unsigned long long *flags = nullptr;
    int n = 0;
    unsigned long long epoch = 0;
};
cudaError_t launch_exchange_push(const void *local, size_t bytes, void *const *peer_slots,
                                 unsigned long long *const *peer_flags, int world, unsigned long long epoch,
                                 cudaStream_t st);
cudaError_t launch_reduce_i64(const float *cand_s, const long long *cand_i, long long list_stride,
                              long long list_stride_i, long long query_stride, int n_lists, int k_in, int k_out, long long id_base,
                              float *out_s, long long *out_i, int n_queries, cudaStream_t st,
                              const WaitFlags *wf = nullptr);
// segment merge (scan.cuh): [n_seg][n_queries][k_seg] sorted lists -> top-k_out per query, k_out <= segmerge_max_k()
int segmerge_max_k();
int segmerge_max_cand();
cudaError_t launch_merge_segments(const float *seg_s, const long long *seg_i, int n_seg, int n_queries, int k_seg,
                                  int k_out, float *out_s, long long *out_i, int *saturated, cudaStream_t st);
cudaError_t launch_pool(const void *hidden, int h_dtype, const void *mask, int m_dtype, int batch, int seq,
                        int dim, int normalize, float *out, cudaStream_t st);
cudaError_t launch_normalize(const float *in, long long in_stride, long long n_rows, int dim, float *out,
                             long long out_stride, void *cast_out, int cast_kind, long long cast_stride,
                             cudaStream_t st);
cudaError_t launch_agree(const long long *ids_a, const float *sa, const long long *ids_b, const float *sb,
                         long long n, double threshold, unsigned char *accept, float *combined, cudaStream_t st);


// ---- sparse (BM25) leg + hybrid fusion (sparse.cuh) ----
struct SparseLaunch {
    const long long *offsets;
    const int *docs;
    const float *weights;
    long long n_docs;
    long long n_terms;
    const int *q_terms;
    const float *q_freqs;
    const int *q_meta;
    int max_terms;
    int n_queries;
    int kcap;
    int ctas_per_query;
    int tiles_per_cta;
    unsigned long long *cand;
    int limit;
    int normalize;
    double avgscore;
    double *out_s;
    long long *out_i;
};
int sparse_max_terms();
int sparse_max_cand();
void sparse_plan(long long n_docs, int n_queries, int sm_count, int *ctas_per_query, int *tiles_per_cta);
cudaError_t launch_sparse_search(const SparseLaunch &a, cudaStream_t st);
cudaError_t launch_bm25_weights(const long long *offsets, long long n_terms, const int *docs, const int *freqs,
                                long long n_postings, const double *idf, const int *doc_len, double k1, double b,
                                double avgdl, float *weights, cudaStream_t st);
cudaError_t launch_hybrid_fuse(const float *dense_s, const long long *dense_i, int kd, const double *sparse_s,
                               const long long *sparse_i, int ks, int n_queries, double w_dense, double w_sparse,
                               int limit, int rrf, double *out_s, long long *out_i, cudaStream_t st);

cudaError_t launch_agree_f64(const long long *ids_a, const double *sa, const long long *ids_b, const double *sb,
                             long long n, double threshold, unsigned char *accept, double *combined, cudaStream_t st);

}  // namespace vqa
