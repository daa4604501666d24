// sparse_launch.cu -- host launchers of the sparse (BM25) leg and the hybrid fusion.
#include "../../include/vqa.h"
#include "launch.h"
#include "sparse.cuh"

namespace vqa {

int sparse_max_terms() { return kSpMaxTerms; }
int sparse_max_cand() { return kSpMaxCand; }

// CTAs per query and tiles per CTA: about two CTAs per SM over the whole batch, contiguous tile runs
void sparse_plan(long long n_docs, int n_queries, int sm_count, int *ctas_per_query, int *tiles_per_cta) {
    const long long n_tiles = n_docs > 0 ? (n_docs + kSpTile - 1) / kSpTile : 1;
    long long cpq = (2LL * sm_count + n_queries - 1) / n_queries;
    if (cpq < 1) cpq = 1;
    if (cpq > n_tiles) cpq = n_tiles;
    const long long tpc = (n_tiles + cpq - 1) / cpq;
    cpq = (n_tiles + tpc - 1) / tpc;
    *ctas_per_query = (int)cpq;
    *tiles_per_cta = (int)tpc;
}

static size_t scan_smem_bytes() {
    return (size_t)kSpTile * sizeof(float) + (size_t)kSpSort * sizeof(unsigned long long) +
           (size_t)kSpBoundSlots * sizeof(int) + (size_t)2 * kSpMaxTerms * sizeof(long long);
}

cudaError_t launch_sparse_search(const SparseLaunch &a, cudaStream_t st) {
    SparseParams p;
    p.offsets = a.offsets;
    p.docs = a.docs;
    p.weights = a.weights;
    p.n_docs = a.n_docs;
    p.n_terms = a.n_terms;
    p.q_terms = a.q_terms;
    p.q_freqs = a.q_freqs;
    p.q_meta = a.q_meta;
    p.max_terms = a.max_terms;
    p.kcap = a.kcap;
    p.ctas_per_query = a.ctas_per_query;
    p.tiles_per_cta = a.tiles_per_cta;
    p.cand = a.cand;
    p.limit = a.limit;
    p.normalize = a.normalize;
    p.avgscore = a.avgscore;
    p.out_s = a.out_s;
    p.out_i = a.out_i;
    const size_t smem = scan_smem_bytes();
    static bool attr_set[64] = {};
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    if (dev < 0 || dev >= 64 || !attr_set[dev]) {
        e = cudaFuncSetAttribute(sparse_scan_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        if (dev >= 0 && dev < 64) attr_set[dev] = true;
    }
    sparse_scan_kernel<<<dim3(a.ctas_per_query, a.n_queries), kSpThreads, smem, st>>>(p);
    e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    sparse_finish_kernel<<<a.n_queries, kSpThreads, 0, st>>>(p);
    return cudaGetLastError();
}

cudaError_t launch_bm25_weights(const long long *offsets, long long n_terms, const int *docs, const int *freqs,
                                long long n_postings, const double *idf, const int *doc_len, double k1, double b,
                                double avgdl, float *weights, cudaStream_t st) {
    if (n_postings == 0) return cudaSuccess;
    long long blocks = (n_postings + 255) / 256;
    if (blocks > 148 * 16) blocks = 148 * 16;
    // (k1 + 1) and (1 - b) are formed once in double, as the Python expression does
    bm25_weights_kernel<<<(unsigned)blocks, 256, 0, st>>>(offsets, n_terms, docs, freqs, n_postings, idf, doc_len, k1,
                                                         k1 + 1.0, 1.0 - b, b, avgdl, weights);
    return cudaGetLastError();
}

cudaError_t launch_hybrid_fuse(const float *dense_s, const long long *dense_i, int kd, const double *sparse_s,
                               const long long *sparse_i, int ks, int n_queries, double w_dense, double w_sparse,
                               int limit, int rrf, double *out_s, long long *out_i, cudaStream_t st) {
    FuseParams p;
    p.rrf = rrf;
    p.dense_s = dense_s;
    p.dense_i = dense_i;
    p.kd = kd;
    p.sparse_s = sparse_s;
    p.sparse_i = sparse_i;
    p.ks = ks;
    p.w_dense = w_dense;
    p.w_sparse = w_sparse;
    p.limit = limit;
    p.out_s = out_s;
    p.out_i = out_i;
    const size_t smem = (size_t)(kd + ks) * (sizeof(double) + sizeof(long long) + sizeof(int));
    hybrid_fuse_kernel<<<n_queries, 256, smem, st>>>(p);
    return cudaGetLastError();
}

cudaError_t launch_agree_f64(const long long *ids_a, const double *sa, const long long *ids_b, const double *sb,
                             long long n, double threshold, unsigned char *accept, double *combined, cudaStream_t st) {
    const long long blocks = (n + 255) / 256;
    agree_f64_kernel<<<(unsigned)blocks, 256, 0, st>>>(ids_a, sa, ids_b, sb, n, threshold, accept, combined);
    return cudaGetLastError();
}

}  // namespace vqa
