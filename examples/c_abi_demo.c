/*
 * c_abi_demo.c -- libvqa_b200.so driven from plain C: no Python, no PyTorch.
 *
 * Builds a small L2-normalised index on the host, uploads it with the CUDA runtime, searches it
 * through the C ABI of include/vqa.h with HOST query/result buffers (vqa_search_host) in the
 * bit-exact fp32 verify mode, and checks ids against a brute-force loop in this file.
 *
 *   gcc -O2 -I../include -I/usr/local/cuda/include c_abi_demo.c -o c_abi_demo \
 *       -L../vietnamese_qa_system_b200 -lvqa_b200 -L/usr/local/cuda/lib64 -lcudart -lm \
 *       -Wl,-rpath,'$ORIGIN/../vietnamese_qa_system_b200'
 */
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#include "vqa.h"

#define CHECK_VQA(x)                                                          \
    do {                                                                      \
        int rc_ = (x);                                                        \
        if (rc_ != VQA_OK) {                                                  \
            fprintf(stderr, "%s -> %d: %s\n", #x, rc_, vqa_last_error());      \
            return 1;                                                         \
        }                                                                     \
    } while (0)
#define CHECK_CUDA(x)                                                         \
    do {                                                                      \
        cudaError_t e_ = (x);                                                 \
        if (e_ != cudaSuccess) {                                              \
            fprintf(stderr, "%s -> %s\n", #x, cudaGetErrorString(e_));         \
            return 1;                                                         \
        }                                                                     \
    } while (0)

static float frand(uint64_t *s) { /* xorshift, uniform in (-1, 1) */
    *s ^= *s << 13;
    *s ^= *s >> 7;
    *s ^= *s << 17;
    return (float)((*s >> 11) * (1.0 / 9007199254740992.0)) * 2.0f - 1.0f;
}

static void normalize(float *x, int d) {
    double ss = 0;
    for (int j = 0; j < d; ++j) ss += (double)x[j] * x[j];
    float inv = (float)(1.0 / sqrt(ss));
    for (int j = 0; j < d; ++j) x[j] *= inv;
}

int main(void) {
    const int64_t n = 20000;
    const int d = 768, b = 4, k = 5;
    if (vqa_device_count() == 0) {
        fprintf(stderr, "no CUDA device: %s\n", "this engine has no CPU fallback");
        return 2;
    }
    printf("libvqa_b200 version %d\n", vqa_version());
    uint64_t seed = 88172645463325252ull;
    float *docs = (float *)malloc(sizeof(float) * n * d), *q = (float *)malloc(sizeof(float) * b * d);
    for (int64_t r = 0; r < n; ++r) {
        for (int j = 0; j < d; ++j) docs[r * d + j] = frand(&seed);
        normalize(docs + r * d, d);
    }
    for (int i = 0; i < b; ++i) { /* queries: noisy copies of known rows, so the right answer is obvious */
        for (int j = 0; j < d; ++j) q[i * d + j] = docs[(int64_t)(1000 + 37 * i) * d + j] + 0.05f * frand(&seed);
        normalize(q + i * d, d);
    }

    float *docs_dev = NULL;
    CHECK_CUDA(cudaMalloc((void **)&docs_dev, sizeof(float) * n * d));
    CHECK_CUDA(cudaMemcpy(docs_dev, docs, sizeof(float) * n * d, cudaMemcpyHostToDevice));

    vqa_index_t *index = NULL;
    CHECK_VQA(vqa_index_create(&index, n, d, VQA_F32, 0, /*first_global_id=*/0));
    CHECK_VQA(vqa_index_bind(index, docs_dev, n, (int64_t)d * sizeof(float)));

    size_t staging_bytes = 0;
    CHECK_VQA(vqa_search_host_staging_bytes(index, b, k, VQA_MODE_VERIFY, &staging_bytes));
    void *staging = NULL;
    CHECK_CUDA(cudaMalloc(&staging, staging_bytes));
    CHECK_CUDA(cudaMemset(staging, 0, staging_bytes)); /* zeroed once after allocation (include/vqa.h) */
    float scores[4 * 5];
    int64_t ids[4 * 5];
    CHECK_VQA(vqa_search_host(index, q, b, k, VQA_MODE_VERIFY, scores, ids, staging, staging_bytes, /*stream=*/NULL));

    int ok = 1;
    for (int i = 0; i < b; ++i) {
        int64_t best = -1;
        double bs = -2;
        for (int64_t r = 0; r < n; ++r) {
            double s = 0;
            for (int j = 0; j < d; ++j) s += (double)q[i * d + j] * docs[r * d + j];
            if (s > bs) {
                bs = s;
                best = r;
            }
        }
        printf("query %d: top-%d ids", i, k);
        for (int j = 0; j < k; ++j) printf(" %lld(%.4f)", (long long)ids[i * k + j], scores[i * k + j]);
        printf("   brute-force top-1 %lld(%.4f)\n", (long long)best, bs);
        ok = ok && ids[i * k] == best && ids[i * k] == 1000 + 37 * i && fabs(scores[i * k] - bs) < 1e-5;
        for (int j = 1; j < k; ++j) ok = ok && scores[i * k + j] <= scores[i * k + j - 1];
    }
    /* error path: status code + message, never abort */
    int rc = vqa_search_host(index, q, b, 1000, VQA_MODE_VERIFY, scores, ids, staging, staging_bytes, NULL);
    ok = ok && rc == VQA_E_INVALID;
    printf("k=1000 -> status %d (%s)\n", rc, vqa_last_error());

    vqa_index_destroy(index);
    cudaFree(staging);
    cudaFree(docs_dev);
    free(docs);
    free(q);
    puts(ok ? "OK" : "MISMATCH");
    return ok ? 0 : 1;
}
