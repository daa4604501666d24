// sparse.cuh -- the sparse (BM25) leg of hybrid search and the dense/sparse fusion.
//
// Reference: the saved indexes are built with txtai.Embeddings(hybrid=True, ...)
// (inference_pipeline/db_utils/heavy_ranker.py:78-83); behind that flag txtai keeps a BM25 term
// index next to the dense one, asks each for 10 x limit candidates and adds the scores per id
// (SURVEY.md 8(f) rank 3, Appendix A).  txtai's Terms.search fills a dense fp32 score array of
// N entries per query on the CPU and argpartitions it.
//
// Here the postings live in HBM as CSR (offsets int64[T+1], docs int32[P] ascending per term,
// weights fp32[P] = the BM25 weight of that (term, document) pair) and the score array never
// exists in memory: a CTA owns a contiguous document range, finds each query term's posting
// sub-range for it by binary search and either (few postings) sorts them by position in shared
// memory and sums each position's run, or (many) accumulates 16 K-document score tiles in SHARED
// memory -- in both cases in the query's term order with fp32 multiply then add, no contraction:
// bit-identical to the CPU accumulation -- and selects straight away.  HBM traffic = the query
// terms' postings, read once, coalesced (8 bytes per posting).
//
// Selection: candidates are 64-bit keys (order-preserving score bits << 32 | ~position), so one
// unsigned compare implements "score descending, ties -> lower position".  A CTA keeps its best
// k_cand keys plus an append buffer in shared memory and bitonic-sorts the 2048 slots when the
// buffer could overflow; after the first tiles the running threshold rejects almost everything.
#pragma once

#include "common.cuh"

namespace vqa {

constexpr int kSpThreads = 512;
constexpr int kSpTile = 16384;                    // documents per shared-memory score tile (64 KB)
constexpr int kSpSort = 2048;                     // key slots per CTA (16 KB)
constexpr int kSpMaxCand = 1024;                  // most candidates a query may keep (k_cand)
constexpr int kSpMaxTerms = 64;                   // distinct known terms per query
constexpr int kSpBoundSlots = 4096;               // tile-edge posting offsets held in shared memory (16 KB)
constexpr int kSpFastMax = 2048;                  // postings in a CTA's document range merged by sorting instead of tiles
constexpr int kSpMetaStride = 4;                  // q_meta row: n_rare, n_common, k_cand, unused

struct SparseParams {
    const long long *offsets;
    const int *docs;
    const float *weights;
    long long n_docs;
    long long n_terms;
    const int *q_terms;    // [B, max_terms]: the n_rare accumulate-everywhere terms in query order, then the n_common
    const float *q_freqs;  // [B, max_terms]: occurrences of the term in the query
    const int *q_meta;     // [B, kSpMetaStride]
    int max_terms;
    int kcap;              // key slots per (query, CTA) in `cand` (>= every k_cand)
    int ctas_per_query;
    int tiles_per_cta;
    unsigned long long *cand;  // [B, ctas_per_query, kcap]
    // finish stage
    int limit;
    int normalize;
    double avgscore;
    double *out_s;     // [B, limit]
    long long *out_i;  // [B, limit]
};

__device__ __forceinline__ unsigned long long sp_make_key(float s, uint32_t doc) {
    uint32_t u = __float_as_uint(s);
    u = (u & 0x80000000u) ? ~u : (u | 0x80000000u);
    return ((unsigned long long)u << 32) | (unsigned long long)(~doc);
}
__device__ __forceinline__ float sp_key_score(unsigned long long key) {
    uint32_t u = (uint32_t)(key >> 32);
    u = (u & 0x80000000u) ? (u & 0x7fffffffu) : ~u;
    return __uint_as_float(u);
}
__device__ __forceinline__ uint32_t sp_key_doc(unsigned long long key) { return ~(uint32_t)key; }

// first index in [lo, hi) whose document is >= v
__device__ __forceinline__ long long sp_lower_bound(const int *docs, long long lo, long long hi, long long v) {
    while (lo < hi) {
        const long long mid = lo + ((hi - lo) >> 1);
        if ((long long)__ldg(docs + mid) < v) lo = mid + 1;
        else hi = mid;
    }
    return lo;
}

// posting range of a query term; ids outside the vocabulary (the caller's arrays are not validated on
// the host -- they live in device memory) read as an empty list instead of out of bounds
__device__ __forceinline__ void sp_term_range(const SparseParams &p, int term, long long &lo, long long &hi) {
    lo = hi = 0;
    if (term >= 0 && term < p.n_terms) {
        lo = p.offsets[term];
        hi = p.offsets[term + 1];
    }
}

struct SpQuery {
    int n_rare, n_common, k_cand;
};
__device__ __forceinline__ SpQuery sp_query(const SparseParams &p, int b) {
    const int *meta = p.q_meta + (size_t)b * kSpMetaStride;
    SpQuery q;
    q.n_rare = max(0, min(meta[0], p.max_terms));
    q.n_common = max(0, min(meta[1], p.max_terms - q.n_rare));
    q.k_cand = max(1, min(meta[2], p.kcap));
    return q;
}

// Bitonic sort of the occupied keys, descending (empty slots are 0 and stay at the end), then keep the
// best k_cand.  `count` and `thr` are CTA-uniform registers; the threshold never drops below `floor_key`.
// Ends with a barrier.
__device__ __forceinline__ void sp_flush(unsigned long long *keys, int &count, int k_cand, unsigned long long &thr,
                                         int *s_count, unsigned long long floor_key = 0ull) {
    const int tid = threadIdx.x;
    // slots >= count are always 0, so sorting the power-of-two prefix that covers `count` is enough
    int n = 2;
    while (n < count) n <<= 1;
    __syncthreads();
    for (int k = 2; k <= n; k <<= 1) {
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int i = tid; i < n; i += kSpThreads) {
                const int x = i ^ j;
                if (x > i) {
                    const unsigned long long a = keys[i], b = keys[x];
                    const bool desc = (i & k) == 0;
                    if (desc ? (a < b) : (a > b)) {
                        keys[i] = b;
                        keys[x] = a;
                    }
                }
            }
            __syncthreads();
        }
    }
    if (count > k_cand) {
        for (int i = k_cand + tid; i < count; i += kSpThreads) keys[i] = 0ull;
        count = k_cand;
    }
    thr = (count >= k_cand) ? keys[k_cand - 1] : 0ull;
    if (thr < floor_key) thr = floor_key;
    if (tid == 0) *s_count = count;
    __syncthreads();
}

// Bitonic sort of a[0..n) ascending, n a power of two <= kSpSort.  Starts and ends with a barrier.
__device__ __forceinline__ void sp_sort_asc(unsigned long long *a, int n) {
    const int tid = threadIdx.x;
    __syncthreads();
    for (int k = 2; k <= n; k <<= 1) {
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int i = tid; i < n; i += kSpThreads) {
                const int x = i ^ j;
                if (x > i) {
                    const unsigned long long u = a[i], v = a[x];
                    const bool asc = (i & k) == 0;
                    if (asc ? (u > v) : (u < v)) {
                        a[i] = v;
                        a[x] = u;
                    }
                }
            }
            __syncthreads();
        }
    }
}

// One CTA-wide offer round: each thread proposes at most one key.  Returns with a barrier; `count`
// stays uniform because the barrier itself counts the appends.
__device__ __forceinline__ void sp_offer(unsigned long long *keys, int &count, unsigned long long thr, int *s_count,
                                         bool have, unsigned long long key) {
    const bool pass = have && key > thr;
    if (pass) keys[atomicAdd(s_count, 1)] = key;
    count += __syncthreads_count(pass);
}

// ---------------------------------------------------------------------------------------------
// Stage 1: accumulate + select.  grid = (ctas_per_query, B), 512 threads, 97 KB smem.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kSpThreads) sparse_scan_kernel(SparseParams p) {
    extern __shared__ __align__(16) unsigned char sp_smem[];
    float *acc = reinterpret_cast<float *>(sp_smem);
    unsigned long long *keys = reinterpret_cast<unsigned long long *>(sp_smem + (size_t)kSpTile * sizeof(float));
    // bound: [n_rare][group + 1] <= kSpBoundSlots posting offsets at tile edges, relative to the term's first posting
    int *bound = reinterpret_cast<int *>(keys + kSpSort);
    // crange: [2 * kSpMaxTerms] posting range of each term inside this CTA's document range
    long long *crange = reinterpret_cast<long long *>(bound + kSpBoundSlots);
    __shared__ int s_count;

    const int tid = threadIdx.x;
    const int b = blockIdx.y;
    const SpQuery qm = sp_query(p, b);
    const int n_rare = qm.n_rare, k_cand = qm.k_cand;
    const int *terms = p.q_terms + (size_t)b * p.max_terms;
    const float *freqs = p.q_freqs + (size_t)b * p.max_terms;

    const long long n_tiles = (p.n_docs + kSpTile - 1) / kSpTile;
    const long long first_tile = (long long)blockIdx.x * p.tiles_per_cta;
    const long long last_tile = min(n_tiles, first_tile + p.tiles_per_cta);

    for (int i = tid; i < kSpSort; i += kSpThreads) keys[i] = 0ull;
    if (tid == 0) s_count = 0;
    int count = 0;
    // Zero-score documents never reach the result (the finish stage drops score 0) except as the
    // zero-fill of a candidate list that a deferred common term may still lift -- and then only the
    // first k_cand positions of the whole index can be among the k_cand best.  So the threshold
    // starts at (0, k_cand) [common terms present] or (0, 0) [none]: only positive scores, or those
    // first positions, ever pass, and posting-free tiles are skipped from the start.
    const unsigned long long floor_key = sp_make_key(0.0f, qm.n_common > 0 ? (uint32_t)k_cand : 0u);
    unsigned long long thr = floor_key;
    __syncthreads();

    // posting sub-range of every term inside this CTA's document range
    const long long doc_lo = min(p.n_docs, first_tile * (long long)kSpTile);
    const long long doc_hi = min(p.n_docs, max(first_tile, last_tile) * (long long)kSpTile);
    for (int w = tid; w < 2 * n_rare; w += kSpThreads) {
        long long lo, hi;
        sp_term_range(p, terms[w >> 1], lo, hi);
        crange[w] = sp_lower_bound(p.docs, lo, hi, (w & 1) ? doc_hi : doc_lo);
    }
    __syncthreads();
    long long p_cta = 0;
    for (int j = 0; j < n_rare; ++j) p_cta += crange[2 * j + 1] - crange[2 * j];

    if (p_cta <= kSpFastMax) {
        // ---- few postings in this range (the usual case for rare terms): no score tile at all.  Stage the
        // postings as (position, term slot, staging index) keys, sort, and sum each position's run in term
        // order -- work proportional to the postings, not to the documents.
        unsigned long long *sk = reinterpret_cast<unsigned long long *>(acc);  // [kSpFastMax], aliases the tile
        float *wbuf = acc + 2 * kSpFastMax;                                    // [kSpFastMax]
        const int P = (int)p_cta;
        int n = 2;
        while (n < P) n <<= 1;
        int prefix = 0;
        for (int j = 0; j < n_rare; ++j) {
            const long long a = crange[2 * j], z = crange[2 * j + 1];
            for (long long pp = a + tid; pp < z; pp += kSpThreads) {
                const int i = prefix + (int)(pp - a);
                sk[i] = ((unsigned long long)(uint32_t)__ldg(p.docs + pp) << 32) | ((unsigned long long)j << 16) |
                        (unsigned long long)i;
                wbuf[i] = __ldg(p.weights + pp);
            }
            prefix += (int)(z - a);
        }
        for (int i = P + tid; i < n; i += kSpThreads) sk[i] = ~0ull;
        sp_sort_asc(sk, n);
        for (int i0 = 0; i0 < P; i0 += kSpThreads) {
            if (count + kSpThreads > kSpSort) sp_flush(keys, count, k_cand, thr, &s_count, floor_key);
            const int i = i0 + tid;
            bool head = false;
            unsigned long long key = 0ull;
            if (i < P) {
                const uint32_t doc = (uint32_t)(sk[i] >> 32);
                head = i == 0 || (uint32_t)(sk[i - 1] >> 32) != doc;
                if (head) {
                    float sum = 0.0f;
                    for (int t = i; t < P && (uint32_t)(sk[t] >> 32) == doc; ++t) {
                        const unsigned long long e = sk[t];
                        sum = __fadd_rn(sum, __fmul_rn(freqs[(int)((e >> 16) & 0xffffu)], wbuf[(int)(e & 0xffffu)]));
                    }
                    key = sp_make_key(sum, doc);
                }
            }
            sp_offer(keys, count, thr, &s_count, head, key);
        }
        if (qm.n_common > 0) {
            // zero-fill positions (see floor_key): documents below k_cand without a posting here
            const long long zhi = min(doc_hi, (long long)k_cand);
            for (long long p0 = doc_lo; p0 < zhi; p0 += kSpThreads) {
                if (count + kSpThreads > kSpSort) sp_flush(keys, count, k_cand, thr, &s_count, floor_key);
                const long long d = p0 + tid;
                bool have = false;
                if (d < zhi) {
                    int lo = 0, hi = P;
                    while (lo < hi) {
                        const int mid = (lo + hi) >> 1;
                        if ((uint32_t)(sk[mid] >> 32) < (uint32_t)d) lo = mid + 1;
                        else hi = mid;
                    }
                    have = !(lo < P && (uint32_t)(sk[lo] >> 32) == (uint32_t)d);
                }
                sp_offer(keys, count, thr, &s_count, have, sp_make_key(0.0f, (uint32_t)d));
            }
        }
    } else {
        // ---- many postings: dense score tiles.  `group` = tile edges whose posting offsets fit the
        // shared-memory table at once
        const int group = max(1, min(p.tiles_per_cta, kSpBoundSlots / max(n_rare, 1) - 1));
        for (long long g0 = first_tile; g0 < last_tile; g0 += group) {
            const int ng = (int)min((long long)group, last_tile - g0);
            // posting boundaries of every term at the ng + 1 tile edges of this group (independent binary
            // searches, one per thread, all in flight together)
            for (int w = tid; w < n_rare * (ng + 1); w += kSpThreads) {
                const int j = w / (ng + 1), t = w - j * (ng + 1);
                long long lo, hi;
                sp_term_range(p, terms[j], lo, hi);
                const long long edge = (g0 + t) * (long long)kSpTile;
                bound[j * (group + 1) + t] = (int)(sp_lower_bound(p.docs, crange[2 * j], crange[2 * j + 1], edge) - lo);
            }
            __syncthreads();
            for (int t = 0; t < ng; ++t) {
                const long long base = (g0 + t) * (long long)kSpTile;
                const int lim = (int)min((long long)kSpTile, p.n_docs - base);
                bool any = false;
                for (int j = 0; j < n_rare; ++j) any |= bound[j * (group + 1) + t + 1] > bound[j * (group + 1) + t];
                // a tile without postings holds only zero scores: none can enter once the threshold is at or
                // above (0, first position of the tile) (uniform decision)
                if (!any && sp_make_key(0.0f, (uint32_t)base) <= thr) continue;
                // lim <= kSpTile, so the 16-byte stores stay inside the tile buffer
                for (int e = tid * 4; e < lim; e += kSpThreads * 4)
                    *reinterpret_cast<float4 *>(acc + e) = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
                __syncthreads();
                for (int j = 0; j < n_rare; ++j) {
                    const int a = bound[j * (group + 1) + t], z = bound[j * (group + 1) + t + 1];
                    if (z > a) {  // uniform
                        const long long off = p.offsets[terms[j]];  // z > a implies a valid term
                        const float qf = freqs[j];
                        constexpr int U = 4;  // postings per thread whose loads are in flight together
                        for (int p0 = a; p0 < z; p0 += U * kSpThreads) {
                            int dd[U];
                            float ww[U];
#pragma unroll
                            for (int u = 0; u < U; ++u) {
                                const int pp = p0 + u * kSpThreads + tid;
                                dd[u] = pp < z ? __ldg(p.docs + off + pp) - (int)base : -1;
                                ww[u] = pp < z ? __ldg(p.weights + off + pp) : 0.0f;
                            }
#pragma unroll
                            for (int u = 0; u < U; ++u)
                                // a document appears once per term: no two threads touch the same slot
                                if (dd[u] >= 0) acc[dd[u]] = __fadd_rn(acc[dd[u]], __fmul_rn(qf, ww[u]));
                        }
                        __syncthreads();
                    }
                }
                // Common case: nothing in the tile beats the threshold -> one barrier.  Screened on the score
                // alone (a float compare per element, 16-byte loads): in a tile that lies wholly after the
                // threshold's position an equal score loses the tie, so the compare is strict there.
                const float thr_s = sp_key_score(thr);
                const bool strict = (uint32_t)base >= sp_key_doc(thr);
                bool mine = false;
                for (int e = tid * 4; e < lim; e += kSpThreads * 4) {
                    if (e + 3 < lim) {
                        const float4 v = *reinterpret_cast<const float4 *>(acc + e);
                        const float m = fmaxf(fmaxf(v.x, v.y), fmaxf(v.z, v.w));
                        mine |= strict ? (m > thr_s) : (m >= thr_s);
                    } else {
                        for (int r = e; r < lim; ++r) mine |= strict ? (acc[r] > thr_s) : (acc[r] >= thr_s);
                    }
                }
                if (!__syncthreads_or(mine)) continue;
                for (int c0 = 0; c0 < lim; c0 += kSpThreads) {
                    if (count + kSpThreads > kSpSort) sp_flush(keys, count, k_cand, thr, &s_count, floor_key);
                    const int e = c0 + tid;
                    const bool have = e < lim;
                    sp_offer(keys, count, thr, &s_count, have, have ? sp_make_key(acc[e], (uint32_t)(base + e)) : 0ull);
                }
            }
            __syncthreads();  // bound[] is rewritten by the next group
        }
    }
    sp_flush(keys, count, k_cand, thr, &s_count, floor_key);
    unsigned long long *out = p.cand + ((size_t)b * p.ctas_per_query + blockIdx.x) * p.kcap;
    for (int i = tid; i < p.kcap; i += kSpThreads) out[i] = i < count ? keys[i] : 0ull;
}

// ---------------------------------------------------------------------------------------------
// Stage 2: one CTA per query.  Merge the CTA lists, add the deferred common terms to the surviving
// candidates (binary search per candidate and term, in query order), re-sort, drop zero scores,
// normalise (txtai: min(score / min(top + avgscore, 6 * avgscore), 1)) and emit `limit` results.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kSpThreads) sparse_finish_kernel(SparseParams p) {
    __shared__ unsigned long long keys[kSpSort];
    __shared__ int s_count;
    const int tid = threadIdx.x;
    const int b = blockIdx.x;
    const SpQuery qm = sp_query(p, b);
    const int n_rare = qm.n_rare, n_common = qm.n_common, k_cand = qm.k_cand;
    const int *terms = p.q_terms + (size_t)b * p.max_terms;
    const float *freqs = p.q_freqs + (size_t)b * p.max_terms;

    for (int i = tid; i < kSpSort; i += kSpThreads) keys[i] = 0ull;
    if (tid == 0) s_count = 0;
    int count = 0;
    unsigned long long thr = 0ull;
    __syncthreads();

    const long long total = (long long)p.ctas_per_query * p.kcap;
    const unsigned long long *cand = p.cand + (size_t)b * total;
    constexpr int U = 4;  // keys loaded per thread before the offer rounds (independent loads in flight)
    for (long long c0 = 0; c0 < total; c0 += (long long)U * kSpThreads) {
        unsigned long long mine[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const long long e = c0 + (long long)u * kSpThreads + tid;
            mine[u] = e < total ? cand[e] : 0ull;  // 0 = empty slot, never above the threshold
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            if (count + kSpThreads > kSpSort) sp_flush(keys, count, k_cand, thr, &s_count);
            sp_offer(keys, count, thr, &s_count, true, mine[u]);
        }
    }
    sp_flush(keys, count, k_cand, thr, &s_count);

    if (n_common > 0) {
        for (int i = tid; i < count; i += kSpThreads) {
            const unsigned long long key = keys[i];
            float s = sp_key_score(key);
            const uint32_t doc = sp_key_doc(key);
            for (int j = 0; j < n_common; ++j) {
                long long lo, hi;
                sp_term_range(p, terms[n_rare + j], lo, hi);
                const long long at = sp_lower_bound(p.docs, lo, hi, (long long)doc);
                if (at < hi && (uint32_t)__ldg(p.docs + at) == doc)
                    s = __fadd_rn(s, __fmul_rn(freqs[n_rare + j], __ldg(p.weights + at)));
            }
            keys[i] = sp_make_key(s, doc);
        }
        sp_flush(keys, count, k_cand, thr, &s_count);
    }

    const float top = count > 0 ? sp_key_score(keys[0]) : 0.0f;
    double maxscore = 1.0;
    if (p.normalize) maxscore = fmin(__dadd_rn((double)top, p.avgscore), __dmul_rn(6.0, p.avgscore));
    for (int r = tid; r < p.limit; r += kSpThreads) {
        double s = __longlong_as_double(0xfff0000000000000LL);  // -inf
        long long id = -1;
        if (r < count) {
            const float f = sp_key_score(keys[r]);
            if (f > 0.0f) {
                s = p.normalize ? fmin(__ddiv_rn((double)f, maxscore), 1.0) : (double)f;
                id = (long long)sp_key_doc(keys[r]);
            }
        }
        p.out_s[(size_t)b * p.limit + r] = s;
        p.out_i[(size_t)b * p.limit + r] = id;
    }
}

// ---------------------------------------------------------------------------------------------
// Build side: BM25 weight of every posting, in the double-precision operation order of txtai's
// BM25.score, rounded once to fp32 (Terms.weights' astype(float32)).
//   k = k1 * ((1 - b) + b * len / avgdl);  w = idf * (f * (k1 + 1)) / (f + k)
// ---------------------------------------------------------------------------------------------
__global__ void bm25_weights_kernel(const long long *offsets, long long n_terms, const int *docs, const int *freqs,
                                    long long n_postings, const double *idf, const int *doc_len, double k1,
                                    double k1_plus_1, double one_minus_b, double bb, double avgdl, float *weights) {
    for (long long pidx = (long long)blockIdx.x * blockDim.x + threadIdx.x; pidx < n_postings;
         pidx += (long long)gridDim.x * blockDim.x) {
        // term of this posting: last t with offsets[t] <= pidx
        long long lo = 0, hi = n_terms;
        while (lo < hi) {
            const long long mid = lo + ((hi - lo) >> 1);
            if (offsets[mid + 1] <= pidx) lo = mid + 1;
            else hi = mid;
        }
        const double f = (double)freqs[pidx];
        const double len = (double)doc_len[docs[pidx]];
        const double k = __dmul_rn(k1, __dadd_rn(one_minus_b, __ddiv_rn(__dmul_rn(bb, len), avgdl)));
        const double w = __ddiv_rn(__dmul_rn(idf[lo], __dmul_rn(f, k1_plus_1)), __dadd_rn(f, k));
        weights[pidx] = __double2float_rn(w);
    }
}

// ---------------------------------------------------------------------------------------------
// Hybrid fusion (txtai Search): per query, fused[id] = 0.0 + dense * w_dense (+ sparse * w_sparse), in
// doubles like the Python floats it replaces; order = fused descending, ties keep insertion order
// (dense candidates first, then sparse-only ones) as Python's stable sort does.  One CTA per query.
// rrf: the leg's contribution is (1.0 / (rank + 1)) * w instead of score * w (rank = position in the leg's
// own candidate list) -- what txtai does when the sparse scores are unbounded raw BM25.  A leg whose weight
// is <= 0 is skipped altogether, as txtai's `scores if weights[v] > 0 else []` does.
// ---------------------------------------------------------------------------------------------
struct FuseParams {
    const float *dense_s;
    const long long *dense_i;
    int kd;
    const double *sparse_s;
    const long long *sparse_i;
    int ks;
    double w_dense, w_sparse;
    int limit;
    int rrf;  // 1: reciprocal-rank fusion (txtai when the scoring index is NOT normalised): 1/(rank+1) * w per leg
    double *out_s;
    long long *out_i;
};

__global__ void __launch_bounds__(256) hybrid_fuse_kernel(FuseParams p) {
    extern __shared__ __align__(16) unsigned char fz_smem[];
    const int n = p.kd + p.ks;
    double *fused = reinterpret_cast<double *>(fz_smem);
    long long *ids = reinterpret_cast<long long *>(fused + n);
    int *valid = reinterpret_cast<int *>(ids + n);
    const int b = blockIdx.x, tid = threadIdx.x;
    const float *ds = p.dense_s + (size_t)b * p.kd;
    const long long *di = p.dense_i + (size_t)b * p.kd;
    const double *ss = p.sparse_s + (size_t)b * p.ks;
    const long long *si = p.sparse_i + (size_t)b * p.ks;

    const bool use_d = p.w_dense > 0.0, use_s = p.w_sparse > 0.0;
    for (int e = tid; e < n; e += blockDim.x) {
        const long long id = e < p.kd ? di[e] : si[e - p.kd];
        ids[e] = (e < p.kd ? use_d : use_s) ? id : -1;  // a skipped leg is all padding
    }
    for (int r = tid; r < p.limit; r += blockDim.x) {
        p.out_s[(size_t)b * p.limit + r] = __longlong_as_double(0xfff0000000000000LL);
        p.out_i[(size_t)b * p.limit + r] = -1;
    }
    __syncthreads();
    for (int e = tid; e < n; e += blockDim.x) {
        const long long id = ids[e];
        int ok = id >= 0;
        double f = 0.0;
        if (ok && e < p.kd) {
            const double c = p.rrf ? __ddiv_rn(1.0, (double)(e + 1)) : (double)ds[e];
            f = __dadd_rn(0.0, __dmul_rn(c, p.w_dense));
            for (int j = 0; j < p.ks; ++j)
                if (ids[p.kd + j] == id) {
                    const double cs = p.rrf ? __ddiv_rn(1.0, (double)(j + 1)) : ss[j];
                    f = __dadd_rn(f, __dmul_rn(cs, p.w_sparse));
                    break;
                }
        } else if (ok) {
            const double cs = p.rrf ? __ddiv_rn(1.0, (double)(e - p.kd + 1)) : ss[e - p.kd];
            f = __dadd_rn(0.0, __dmul_rn(cs, p.w_sparse));
            for (int j = 0; j < p.kd; ++j)
                if (ids[j] == id) {  // already fused into the dense entry
                    ok = 0;
                    break;
                }
        }
        fused[e] = f;
        valid[e] = ok;
    }
    __syncthreads();
    for (int e = tid; e < n; e += blockDim.x) {
        if (!valid[e]) continue;
        const double f = fused[e];
        int rank = 0;
        for (int o = 0; o < n; ++o)
            rank += valid[o] && (fused[o] > f || (fused[o] == f && o < e));
        if (rank < p.limit) {
            p.out_s[(size_t)b * p.limit + rank] = f;
            p.out_i[(size_t)b * p.limit + rank] = ids[e];
        }
    }
}

// heavy_ranker.py:110 with hybrid indexes: the two scores are already Python-float (double) sums
__global__ void agree_f64_kernel(const long long *ids_a, const double *sa, const long long *ids_b, const double *sb,
                                 long long n, double threshold, unsigned char *accept, double *combined) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double sum = __dadd_rn(sa[i], sb[i]);
    if (accept) accept[i] = (ids_a[i] == ids_b[i] && sum > threshold) ? 1 : 0;
    if (combined) combined[i] = sum;
}

}  // namespace vqa
