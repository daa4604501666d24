/*
 * oracle.c -- CPU restatement of the reference's dense-retrieval hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing in the product package
 * (vietnamese_qa_system_b200/) may import, link or execute this file; only
 * tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs use it, and there only as the checker / the CPU baseline.
 *
 * PARITY UNPINNED: the reference (vTuanpham/Vietnamese_QA_System) holds no
 * retrieval arithmetic, tests, golden vectors or fixtures of its own.  Its
 * retriever call (inference_pipeline/db_utils/heavy_ranker.py:78-101) delegates
 * to the third-party package `txtai` (requirements.txt:74, unpinned, not
 * vendored, not installable here) which in turn calls faiss-cpu.  This file
 * restates txtai's published exact-search semantics (6.x line, contemporary
 * with transformers==4.33.1 / sentence-transformers==2.2.2):
 *
 *   - mean pooling:   sum_s(h[b,s,:]*m[b,s]) / max(sum_s m[b,s], 1e-9)
 *                     (txtai models/pooling MeanPooling; in-tree twin
 *                      src/test.py:97-99 via SentenceTransformer.encode)
 *   - normalise:      x /= ||x||_2 in fp32, rows of documents and queries
 *                     (under heavy_ranker.py:86,88 and :98,100)
 *   - score:          inner product of unit vectors = cosine
 *                     (faiss METRIC_INNER_PRODUCT; src/test.py:104)
 *   - top-k:          k largest, descending score; ties -> lower position
 *                     (txtai NumPy backend: stable sorted(..., reverse=True);
 *                      BASELINE.json north_star: "ties broken by lower doc_id")
 *   - position -> id: ids[position]  (heavy_ranker.py:74-76, setup_db.py:14)
 *   - agreement rule: uid_a == uid_b and score_a + score_b > 0.4
 *                     (heavy_ranker.py:110)
 *
 * Two score flavours are provided:
 *   semantic  -- fp64 accumulation, the mathematical meaning of the result;
 *   canonical -- fp32 with a fixed FMA/reduction order:
 *                32 "lanes"; lane l owns the 16-byte chunks c = l, l+32, ...
 *                of a row (E = 4 fp32 or 8 sixteen-bit elements per chunk),
 *                accumulates acc = fmaf(q[d], x[d], acc) over its elements in
 *                increasing d from +0.0f, and the 32 partials are combined by
 *                the XOR butterfly m = 16, 8, 4, 2, 1.  The CUDA verify kernel
 *                follows the same order, so ids AND score bits can be compared.
 *
 * Build: gcc -O2 -ffp-contract=off -shared -fPIC (see oracle/Makefile).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define ORACLE_API __attribute__((visibility("default")))

/* (score desc, id asc): returns 1 when a ranks strictly before b. */
static inline int before(float as, int64_t ai, float bs, int64_t bi) {
    return (as > bs) || (as == bs && ai < bi);
}

/* ---- a3: L2 normalise rows (fp64 accumulate, fp32 result).  Zero-norm rows
 * stay all-zero (numpy would give NaN; SURVEY.md 8(a) a3 defines this away). */
ORACLE_API void oracle_normalize_rows(const float *x, int64_t n, int d, float *out) {
    for (int64_t r = 0; r < n; ++r) {
        const float *xr = x + r * (int64_t)d;
        double ss = 0.0;
        for (int j = 0; j < d; ++j) ss += (double)xr[j] * (double)xr[j];
        double nrm = sqrt(ss);
        for (int j = 0; j < d; ++j)
            out[r * (int64_t)d + j] = nrm > 0.0 ? (float)((double)xr[j] / nrm) : 0.0f;
    }
}

/* ---- a2: masked mean pool (+ optional normalise).  mask is float weights
 * (the reference casts attention_mask to float).  fp64 accumulate. */
ORACLE_API void oracle_mean_pool(const float *h, const float *mask, int b, int s, int d,
                                 int normalize, float *out) {
    double *acc = (double *)malloc(sizeof(double) * (size_t)d);
    for (int i = 0; i < b; ++i) {
        memset(acc, 0, sizeof(double) * (size_t)d);
        double cnt = 0.0;
        for (int t = 0; t < s; ++t) {
            double m = (double)mask[(int64_t)i * s + t];
            cnt += m;
            if (m == 0.0) continue;
            const float *row = h + ((int64_t)i * s + t) * d;
            for (int j = 0; j < d; ++j) acc[j] += (double)row[j] * m;
        }
        double den = cnt < 1e-9 ? 1e-9 : cnt;
        double ss = 0.0;
        for (int j = 0; j < d; ++j) {
            /* the reference materialises the pooled vector in fp32 before normalising */
            float p = (float)(acc[j] / den);
            out[(int64_t)i * d + j] = p;
            ss += (double)p * (double)p;
        }
        if (normalize) {
            double nrm = sqrt(ss);
            for (int j = 0; j < d; ++j) {
                float p = out[(int64_t)i * d + j];
                out[(int64_t)i * d + j] = nrm > 0.0 ? (float)((double)p / nrm) : 0.0f;
            }
        }
    }
}

/* ---- a4 (canonical): one fp32 dot product in the fixed order. */
ORACLE_API float oracle_dot_canonical(const float *q, const float *x, int d, int elems_per_chunk) {
    float part[32], nxt[32];
    int nchunks = (d + elems_per_chunk - 1) / elems_per_chunk;
    for (int lane = 0; lane < 32; ++lane) {
        float acc = 0.0f;
        for (int c = lane; c < nchunks; c += 32) {
            for (int e = 0; e < elems_per_chunk; ++e) {
                int j = c * elems_per_chunk + e;
                if (j < d) acc = fmaf(q[j], x[j], acc);
            }
        }
        part[lane] = acc;
    }
    for (int m = 16; m >= 1; m >>= 1) {
        for (int lane = 0; lane < 32; ++lane) nxt[lane] = part[lane] + part[lane ^ m];
        memcpy(part, nxt, sizeof(part));
    }
    return part[0];
}

/* ---- a4 (semantic): fp64 accumulate. */
ORACLE_API double oracle_dot_semantic(const float *q, const float *x, int d) {
    double acc = 0.0;
    for (int j = 0; j < d; ++j) acc += (double)q[j] * (double)x[j];
    return acc;
}

/* sorted insert of (s, id) into a descending list of length *len (capacity k). */
static void list_insert(float *ls, int64_t *li, int *len, int k, float s, int64_t id) {
    int n = *len;
    if (n == k && !before(s, id, ls[k - 1], li[k - 1])) return;
    int p = n < k ? n : k - 1;
    while (p > 0 && before(s, id, ls[p - 1], li[p - 1])) {
        ls[p] = ls[p - 1];
        li[p] = li[p - 1];
        --p;
    }
    ls[p] = s;
    li[p] = id;
    if (n < k) *len = n + 1;
}

/* ---- a4 + a5: exact flat search.  docs [n,d] fp32 (16-bit storage is passed
 * upcast, with elems_per_chunk = 8), queries [b,d] fp32.  mode 0 = canonical
 * fp32 order, mode 1 = semantic (fp64 accumulate, rounded to fp32 for ranking).
 * ids are first_id + position.  Unfilled slots (k > n): score -inf, id -1. */
ORACLE_API void oracle_search(const float *docs, int64_t n, int d, const float *queries, int b,
                              int k, int mode, int elems_per_chunk, int64_t first_id,
                              float *out_scores, int64_t *out_ids) {
    for (int i = 0; i < b; ++i) {
        float *ls = out_scores + (int64_t)i * k;
        int64_t *li = out_ids + (int64_t)i * k;
        int len = 0;
        const float *q = queries + (int64_t)i * d;
        for (int64_t r = 0; r < n; ++r) {
            const float *x = docs + r * (int64_t)d;
            float s = mode == 0 ? oracle_dot_canonical(q, x, d, elems_per_chunk)
                                : (float)oracle_dot_semantic(q, x, d);
            list_insert(ls, li, &len, k, s, first_id + r);
        }
        for (int j = len; j < k; ++j) {
            ls[j] = -INFINITY;
            li[j] = -1;
        }
    }
}

/* ---- scores only (for tolerance checks): out [b,n]. */
ORACLE_API void oracle_scores(const float *docs, int64_t n, int d, const float *queries, int b,
                              int mode, int elems_per_chunk, float *out) {
    for (int i = 0; i < b; ++i)
        for (int64_t r = 0; r < n; ++r)
            out[(int64_t)i * n + r] =
                mode == 0 ? oracle_dot_canonical(queries + (int64_t)i * d, docs + r * (int64_t)d, d,
                                                 elems_per_chunk)
                          : (float)oracle_dot_semantic(queries + (int64_t)i * d,
                                                       docs + r * (int64_t)d, d);
}

/* ---- K4: merge `lists` candidate lists per query, each [b,k] sorted or not;
 * entries with id < 0 are padding.  cand_* layout: [lists][b][k]. */
ORACLE_API void oracle_merge_topk(const float *cand_scores, const int64_t *cand_ids, int lists,
                                  int b, int k_in, int k_out, float *out_scores,
                                  int64_t *out_ids) {
    for (int i = 0; i < b; ++i) {
        float *ls = out_scores + (int64_t)i * k_out;
        int64_t *li = out_ids + (int64_t)i * k_out;
        int len = 0;
        for (int l = 0; l < lists; ++l)
            for (int j = 0; j < k_in; ++j) {
                int64_t off = ((int64_t)l * b + i) * k_in + j;
                if (cand_ids[off] < 0) continue;
                list_insert(ls, li, &len, k_out, cand_scores[off], cand_ids[off]);
            }
        for (int j = len; j < k_out; ++j) {
            ls[j] = -INFINITY;
            li[j] = -1;
        }
    }
}

/* ---- a8: two-index agreement rule (heavy_ranker.py:110). */
ORACLE_API int oracle_agree(int64_t uid_a, float score_a, int64_t uid_b, float score_b,
                            double threshold) {
    /* Python adds the two scores as floats (fp64) and compares with the literal 0.4 */
    return uid_a == uid_b && ((double)score_a + (double)score_b) > threshold;
}
