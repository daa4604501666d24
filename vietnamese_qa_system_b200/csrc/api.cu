// api.cu -- the C ABI of libvqa_b200.so (include/vqa.h): argument validation,
// kernel-family dispatch, launches.  No torch types, no hidden allocations in
// the search path, no CPU fallback.
#include <cuda.h>
#include <cuda_runtime.h>

#include <atomic>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <new>
#include <string>

#include "../../include/vqa.h"
#include "launch.h"

namespace {

thread_local std::string g_err;

int fail(int code, const char *fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    g_err = buf;
    return code;
}

#define CUDA_TRY(expr)                                                                          \
    do {                                                                                        \
        cudaError_t e_ = (expr);                                                                \
        if (e_ != cudaSuccess)                                                                  \
            return fail(VQA_E_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e_),     \
                        __FILE__, __LINE__);                                                    \
    } while (0)

struct DeviceGuard {
    int prev = -1;
    bool switched = false;
    cudaError_t err = cudaSuccess;
    explicit DeviceGuard(int dev) {
        err = cudaGetDevice(&prev);
        if (err == cudaSuccess && prev != dev) {
            err = cudaSetDevice(dev);
            switched = err == cudaSuccess;
        }
    }
    ~DeviceGuard() {
        if (switched) cudaSetDevice(prev);
    }
};

int elem_size(int dtype) {
    switch (dtype) {
        case VQA_F32: return 4;
        case VQA_BF16: return 2;
        case VQA_F16: return 2;
        default: return 0;
    }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                  const cuuint64_t *, const cuuint32_t *, const cuuint32_t *,
                                  CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                  CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode_fn() {
    static EncodeTiledFn fn = nullptr;
    static bool tried = false;
    if (!tried) {
        tried = true;
        void *p = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
            qres == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(p);
        else
            (void)cudaGetLastError();
    }
    return fn;
}

}  // namespace

struct vqa_index {
    int64_t n_rows = 0;
    int dim = 0;
    int dtype = 0;
    int device = 0;
    int64_t first_id = 0;
    const void *rows = nullptr;
    int64_t stride = 0;
    int sm_count = 0;
    int max_smem = 0;
    bool tmap_ok = false;
    // tmap[j]: TMA box of 64 columns x (128 >> j) rows -- j = log2(cluster size) of the multicast launch
    alignas(64) CUtensorMap tmap[4];
    vqa_tuning_t tune;  // kernel-selection knobs: resolved once (create / set_tuning), never read from the environment later
    // caches filled by the (const) search path; guarded by `mu` so that concurrent searches on one handle are safe
    mutable std::mutex mu;
    mutable int max_clusters[4][5] = {};  // cudaOccupancyMaxActiveClusters by [log2 cluster][ncol slot]
    struct PlanSlot {
        int nq = 0, k = 0, mode = -1;
        int plan[16] = {};
    };
    mutable PlanSlot plan_cache[8];
    mutable unsigned plan_next = 0;
    unsigned long long *timeline = nullptr;  // vqa_debug_timeline
    size_t timeline_bytes = 0;
    // vqa_search_2s: scan stream -> reduce stream hand-over.  A stream wait captures the event's record at enqueue
    // time, so ONE event serves every call (guarded by `mu`); created on first use, destroyed with the handle.
    mutable cudaEvent_t handoff = nullptr;
    ~vqa_index() {
        if (handoff) cudaEventDestroy(handoff);
    }
};

// sparse (BM25) term index: CSR postings borrowed from the caller
struct vqa_sparse {
    int64_t n_docs = 0;
    int64_t n_terms = 0;
    int64_t n_postings = 0;
    int device = 0;
    int sm_count = 0;
    const long long *offsets = nullptr;
    const int *docs = nullptr;
    const float *weights = nullptr;
};

namespace {

// ---- planning -----------------------------------------------------------------
struct Plan {
    int family;   // VQA_MODE_FAST_STREAM (also verify) or VQA_MODE_FAST_TENSOR
    int pass_nq;  // queries per pass
    int ncol;     // tensor: MMA N
    int stages;   // tensor: smem ring depth (stages of kps x 16 KB)
    int kps;      // tensor: k-blocks per stage
    int passes;
    int groups;   // tensor: query chunks handled side by side per launch (documents shared through L2)
    int ss_split; // tensor family: 1 hi/lo column pairs, 0 screen mode (+ exact re-scoring in the reduce)
    int ts_split; // TS family: hi+lo rows (64 queries per CTA) or storage-precision queries (128 per CTA)
    int ts_afp16;
    int ts_qs;    // TS family: the opt-in QS kernel variant (part of the query block in shared memory)
    int ts_ks;    // QS: 64-column blocks of the query block kept in shared memory
    int ts_m64;   // QS, <= 64 queries, screen mode: M = 64 instructions
    int grid;
};

// ---- knobs ---------------------------------------------------------------------
struct KnobSpec {
    const char *env;
    int32_t vqa_tuning_t::*field;
    int lo, hi, dflt;
};
// defaults = what the B200 measurements of rounds 1-2 support (profiles/r2_*): QS variant of the TMEM-resident-query
// kernel and the radix-select / re-scoring reduce on, batches of <= 2 on the CUDA-core streaming kernel
const KnobSpec kKnobs[] = {
    {"VQA_TS_EXTRA", &vqa_tuning_t::ts_extra, 0, 96, 6},
    {"VQA_SS_SCREEN", &vqa_tuning_t::ss_screen, -1, 1, -1},
    {"VQA_MMA_KPS", &vqa_tuning_t::mma_kps, 0, 16, 0},
    {"VQA_MMA_STAGES", &vqa_tuning_t::mma_stages, 0, vqa::kMaxStages, 0},
    {"VQA_MMA_GROUPS", &vqa_tuning_t::mma_groups, 1, 4, 4},
    {"VQA_MMA_MULTICAST", &vqa_tuning_t::mma_multicast, 0, 1, 1},
    {"VQA_TS_QS", &vqa_tuning_t::ts_qs, 0, 1, 1},
    {"VQA_TS_KS", &vqa_tuning_t::ts_ks, -1, 16, -1},
    {"VQA_TS_SPLIT", &vqa_tuning_t::ts_split, -1, 1, -1},
    {"VQA_TS_GROUPS", &vqa_tuning_t::ts_groups, 1, 4, 2},
    {"VQA_REDUCE_SELECT", &vqa_tuning_t::reduce_select, 0, 1, 1},
    {"VQA_REDUCE_EARLY", &vqa_tuning_t::reduce_early, 0, 1, 1},
    {"VQA_PDL_CHAIN", &vqa_tuning_t::pdl_chain, 0, 1, 0},
    {"VQA_TMA_L2PROMO", &vqa_tuning_t::tma_l2promo, 0, 3, 3},
    {"VQA_TMA_HINT", &vqa_tuning_t::tma_hint, 0, 2, 1},
    {"VQA_STREAM_MAX_B", &vqa_tuning_t::stream_max_b, 0, 8, 0},
    {"VQA_STREAM_MIN_MB", &vqa_tuning_t::stream_min_mb, 0, 1 << 30, 8000},
    {"VQA_PAIR", &vqa_tuning_t::pair, 0, 1, 1},
    {"VQA_DYN_TILES", &vqa_tuning_t::dyn_tiles, 0, 1, 1},
    {"VQA_SEED", &vqa_tuning_t::seed, 0, 1, 1},
    {"VQA_WIDE", &vqa_tuning_t::wide, 0, 1, 1},
    {"VQA_TS_M64", &vqa_tuning_t::ts_m64, 0, 1, 1},
    {"VQA_SMEM_RESERVE_KB", &vqa_tuning_t::smem_reserve_kb, 0, 64, 0},
};

void tuning_defaults(vqa_tuning_t *t) {
    std::memset(t, 0, sizeof(*t));
    t->size = (int32_t)sizeof(vqa_tuning_t);
    for (const KnobSpec &kn : kKnobs) t->*(kn.field) = kn.dflt;
}

int tuning_validate(const vqa_tuning_t *t) {
    if (!t) return fail(VQA_E_INVALID, "tuning is null");
    if (t->size != (int32_t)sizeof(vqa_tuning_t))
        return fail(VQA_E_INVALID, "vqa_tuning_t size mismatch: caller %d, library %d (start from vqa_tuning_default)",
                    (int)t->size, (int)sizeof(vqa_tuning_t));
    for (const KnobSpec &kn : kKnobs) {
        const int v = t->*(kn.field);
        if (v < kn.lo || v > kn.hi)
            return fail(VQA_E_INVALID, "tuning knob %s = %d out of range [%d, %d]", kn.env + 4, v, kn.lo, kn.hi);
    }
    return VQA_OK;
}

int tuning_from_env(vqa_tuning_t *t) {
    tuning_defaults(t);
    for (const KnobSpec &kn : kKnobs) {
        const char *e = std::getenv(kn.env);
        if (!e || !*e) continue;
        char *end = nullptr;
        const long v = std::strtol(e, &end, 10);
        if (end == e || *end != '\0' || v < kn.lo || v > kn.hi)
            return fail(VQA_E_INVALID, "environment knob %s=\"%s\" is not an integer in [%d, %d]", kn.env, e, kn.lo, kn.hi);
        t->*(kn.field) = (int32_t)v;
    }
    return VQA_OK;
}

bool tensor_eligible(const vqa_index *h) {
    return h->tmap_ok && (h->dtype == VQA_BF16 || h->dtype == VQA_F16) && h->dim % 64 == 0 && h->dim >= 64;
}

int spare_ranks(const vqa_index *h) { return h->tune.ts_extra; }

// shared memory a scan CTA may take: the device limit minus what the knob keeps free for a kernel that should run
// NEXT to the scan on the same SM (the re-scoring reduce of the previous batch in a pipelined loop needs ~21 KB)
size_t smem_budget(const vqa_index *h, bool big_reduce) {
    const size_t keep = big_reduce ? (size_t)h->tune.smem_reserve_kb * 1024 : 0;
    return (size_t)h->max_smem > keep ? (size_t)h->max_smem - keep : 0;
}

// pick the widest MMA N (<= what the batch needs) whose smem ring still has >= 4 boxes.
// Screen mode (k + spare <= 32): one storage-precision column per query, up to 32 queries per CTA,
// the k + spare best re-scored exactly in the reduce.  Otherwise hi/lo column pairs.
bool plan_tensor(const vqa_index *h, int nq, int k, Plan *pl) {
    const vqa_tuning_t &tu = h->tune;
    // Measured (profiles/r1_tune_screen.log): for <= 32 queries per CTA the hi/lo kernel already runs at
    // the HBM roofline and the re-scoring stage costs ~60 us per search, so screen mode is off for short scans.
    // A LONG scan with more than 16 queries is another matter: sustained, the hi/lo kernel's tensor work runs the
    // GPU into its power cap (1.5 GHz), and halving it buys more than the re-scoring costs -- A/B/A at 10 M x 768,
    // B = 32: 2.124 / 2.129 ms against 2.066 ms (profiles/r2_call35.log).  ss_screen = -1 (default) therefore
    // screens when one scan streams >= 12 GB (break-even: 60 us against ~3 % of the scan) and nq > 16.
    const long long scan_bytes = (long long)h->n_rows * h->dim * elem_size(h->dtype);
    const bool want_screen = tu.ss_screen < 0 ? (nq > 16 && scan_bytes >= 12000000000LL) : tu.ss_screen != 0;
    const bool screen = want_screen && k + spare_ranks(h) <= 32;
    const int kk = screen ? k + spare_ranks(h) : k;
    const int cands[4] = {128, 64, 32, 16};
    const int first = screen ? 2 : 0;  // screen mode keeps its lists in registers: <= 32 queries per CTA
    int want = screen ? nq : nq * 2;
    for (int ci = first; ci < 4; ++ci) {
        int ncol = cands[ci];
        if (ci < 3 && cands[ci + 1] >= want) continue;  // a narrower tile still covers the batch
        size_t fixed = vqa::mma_smem_bytes_rt(ncol, h->dim, kk, 0, screen ? 0 : 1);
        if (fixed >= (size_t)h->max_smem) continue;
        int blocks = (int)(((size_t)h->max_smem - fixed) / vqa::kStageBytes);  // 16 KB boxes that fit
        if (blocks < 4) continue;
        // two adjacent 128-byte column blocks of the same rows per stage: the pair is requested
        // together, so each 256-byte DRAM/L2 granule is touched once (measured: +24% bandwidth)
        const int kb = h->dim / vqa::kBlockK;
        int kps = tu.mma_kps > 0 ? tu.mma_kps : (kb % 2 == 0 ? 2 : (kb % 3 == 0 ? 3 : 1));
        if (kps < 1 || kb % kps != 0) kps = 1;
        int stages = blocks / kps;
        if (stages > vqa::kMaxStages) stages = vqa::kMaxStages;
        if (tu.mma_stages > 0 && tu.mma_stages < stages) stages = tu.mma_stages;
        if (stages < 2) continue;
        pl->family = VQA_MODE_FAST_TENSOR;
        pl->ncol = ncol;
        pl->ss_split = screen ? 0 : 1;
        pl->pass_nq = screen ? ncol : ncol / 2;
        pl->stages = stages;
        pl->kps = kps;
        pl->passes = (nq + pl->pass_nq - 1) / pl->pass_nq;
        pl->groups = pl->passes < tu.mma_groups ? pl->passes : tu.mma_groups;
        long long tiles = (h->n_rows + vqa::kTileRows - 1) / vqa::kTileRows;
        pl->grid = (int)(tiles < h->sm_count ? (tiles > 0 ? tiles : 1) : h->sm_count);
        return true;
    }
    return false;
}

// The TMEM-resident-query kernel holds at most 12 of the query block's 64-column blocks in tensor memory
// (384 columns + two accumulator stages); the QS variant keeps the rest in shared memory (dim <= 1024).
bool ts_eligible(const vqa_index *h) {
    return tensor_eligible(h) && (h->dim <= 768 || (h->tune.ts_qs != 0 && h->dim <= 1024));
}

// TMEM-resident queries: shared memory holds only the document ring and the per-row lists
bool plan_ts(const vqa_index *h, int nq, int k, Plan *pl) {
    const vqa_tuning_t &tu = h->tune;
    // k <= 16: screen with storage-precision queries (128 per CTA), keep 32 candidates per query and
    // re-score them exactly in the reduce.  Larger k: hi + lo rows (64 queries per CTA).
    const int kb = h->dim / vqa::kBlockK;
    // QS variant: ks of the kb query blocks in shared memory.  At most 12 fit tensor memory beside two accumulator
    // stages; 10 leave THREE.  Measured (profiles/r2_call1.log): at dim 768 ks = 2 beats 0 / 4 / 6 at every batch
    // size on the 10 M-row index (B = 64: 2.28 vs 2.59 ms at ks = 4) and ties at the 1.25 M-row shard; at dim 1024
    // ks = 6 ties 4 and 8 loses 40 % (the ring shrinks to 80 KB).  So auto = kb - 10.
    const int qs = tu.ts_qs != 0 ? 1 : 0;
    int ks = 0;
    if (qs) {
        const int ks_min = kb > 12 ? kb - 12 : 0;
        ks = tu.ts_ks >= 0 ? tu.ts_ks : (kb > 10 ? kb - 10 : 0);
        if (ks < ks_min) ks = ks_min;
        if (ks > kb) ks = kb;
    } else if (h->dim > 768) {
        return false;
    }
    // Screen mode beyond the register lists (k + spare > 32) re-scores through the radix-select reduce, the only
    // big-k reduce with a re-scoring stage: fp16 rows only (11-bit queries; bf16 queries would need ~28 spare
    // ranks at top-100).  Anything else with k + spare > 32: hi/lo rows.
    const int spare = spare_ranks(h);
    const bool big_screen_ok = qs && tu.reduce_select != 0 && k + spare <= vqa::kMaxK;
    int split = tu.ts_split >= 0 ? tu.ts_split : ((k + spare <= 32 || (big_screen_ok && h->dtype == VQA_F16)) ? 0 : 1);
    if (!split && k + spare > 32 && !big_screen_ok) split = 1;
    const int kscan = split ? k : k + spare;
    const int m64 = (qs && !split && nq <= 64 && tu.ts_m64) ? 1 : 0;
    const size_t fixed = vqa::ts_smem_bytes(kscan, 0, split, ks, nq, qs, m64);
    const size_t budget = smem_budget(h, !split && kscan <= 32);   // (screen mode with register lists: select-kernel reduce)
    if (fixed >= budget) return false;
    int boxes = (int)((budget - fixed) / (vqa::kStageBytes / 2));  // 8 KB boxes
    // 8 KB boxes: four column blocks per ring stage halve the per-byte handshakes (measured 2.74 -> 2.57 ms
    // at B = 128, 5.17 -> 4.38 ms at B = 256 on 10M x 768)
    int kps = tu.mma_kps > 0 ? tu.mma_kps : (kb % 4 == 0 ? 4 : (kb % 3 == 0 ? 3 : (kb % 2 == 0 ? 2 : 1)));
    if (kps < 1 || kb % kps != 0) kps = 1;
    // QS: the query blocks shrink the ring; keep at least three stages in flight before widening them
    while (qs && kps > 1 && boxes / kps < 3) kps = (kps % 2 == 0) ? kps / 2 : 1;
    int stages = boxes / kps;
    if (stages > vqa::kMaxStages) stages = vqa::kMaxStages;
    if (stages < 2) return false;
    pl->family = VQA_MODE_FAST_TS;
    pl->ts_split = split;
    pl->ts_qs = qs;
    pl->ts_ks = ks;
    pl->ts_afp16 = 0;  // (fp16 queries against bf16 rows would halve the rounding, but the MMA rejects mixed operands)
    pl->ts_m64 = m64;
    pl->pass_nq = (split || m64) ? 64 : 128;
    pl->ncol = 0;
    pl->stages = stages;
    pl->kps = kps;
    pl->passes = (nq + pl->pass_nq - 1) / pl->pass_nq;
    pl->groups = pl->passes < tu.ts_groups ? pl->passes : tu.ts_groups;
    pl->grid = h->sm_count;
    return true;
}

// CTA-pair kernel (pair.cuh): screen mode only; the tensor-memory part of the query block must leave two
// 128-column accumulator stages, so at most 8 of its 64-column blocks stay there
bool pair_eligible(const vqa_index *h, int k) {
    return tensor_eligible(h) && h->dim <= 1024 && k + spare_ranks(h) <= 32;
}

// ring geometry of the 128-document-tile kernel: `single` = one CTA fetches both 64-row halves of a tile (16 KB per
// 64-column block), else each CTA of a pair fetches its own half (8 KB)
bool pair_geometry(const vqa_index *h, int ks, bool single, int *stages, int *kps_out) {
    const vqa_tuning_t &tu = h->tune;
    const int kb = h->dim / vqa::kBlockK;
    const size_t fixed = vqa::pair_smem_bytes(0, ks);
    const size_t budget = smem_budget(h, true);
    if (fixed >= budget) return false;
    const size_t unit = single ? vqa::kStageBytes : vqa::kStageBytes / 2;
    const int slots = (int)((budget - fixed) / unit);   // 64-column blocks of a tile that fit the ring
    // measured (profiles/r2_call5.log, pairs at dim 768): 4 blocks per stage 3.41 ms, 3 or 6: 3.94, 2: 5.17, 1: 9.57
    int kps = tu.mma_kps > 0 ? tu.mma_kps : (kb % 4 == 0 ? 4 : (kb % 3 == 0 ? 3 : (kb % 2 == 0 ? 2 : 1)));
    if (kps < 1 || kb % kps != 0) kps = 1;
    while (kps > 1 && slots / kps < 2) kps = (kps % 2 == 0) ? kps / 2 : 1;
    int st = slots / kps;
    if (st > vqa::kMaxStages) st = vqa::kMaxStages;
    if (st < 2) return false;
    *stages = st;
    *kps_out = kps;
    return true;
}

bool plan_pair(const vqa_index *h, int nq, int k, Plan *pl) {
    if (!pair_eligible(h, k)) return false;
    const vqa_tuning_t &tu = h->tune;
    const int kb = h->dim / vqa::kBlockK;
    const int ks_min = kb > 8 ? kb - 8 : 0;
    int ks = tu.ts_ks >= 0 ? tu.ts_ks : ks_min;
    if (ks < ks_min) ks = ks_min;
    if (ks > kb) ks = kb;
    int stages = 0, kps = 0, s1 = 0, k1 = 0;
    if (!pair_geometry(h, ks, nq <= 128, &stages, &kps)) return false;
    if (nq > 128 && nq % 256 != 0 && nq % 256 <= 128 && !pair_geometry(h, ks, true, &s1, &k1)) return false;  // the tail launch
    pl->family = VQA_MODE_FAST_PAIR;
    pl->ts_split = 0;
    pl->ts_qs = 1;
    pl->ts_ks = ks;
    pl->pass_nq = nq <= 128 ? 128 : 256;
    pl->stages = stages;
    pl->kps = kps;
    pl->passes = nq <= 128 ? 1 : (nq + 255) / 256;
    pl->groups = 1;
    pl->grid = nq <= 128 ? h->sm_count : (h->sm_count & ~1);
    return true;
}

void plan_stream(const vqa_index *h, int nq, Plan *pl) {
    pl->family = VQA_MODE_FAST_STREAM;
    pl->pass_nq = nq >= 5 ? 8 : (nq >= 3 ? 4 : (nq == 2 ? 2 : 1));
    pl->passes = (nq + pl->pass_nq - 1) / pl->pass_nq;
    pl->ncol = 0;
    pl->stages = 0;
    pl->kps = 0;
    pl->groups = 1;
    pl->ss_split = 1;
    pl->ts_split = 1;
    pl->ts_afp16 = 0;
    pl->ts_qs = 0;
    pl->ts_ks = 0;
    pl->grid = h->sm_count;
}

int make_plan_uncached(const vqa_index *h, int nq, int k, int mode, Plan *pl) {
    std::memset(pl, 0, sizeof(*pl));
    if (mode == VQA_MODE_VERIFY || mode == VQA_MODE_FAST_STREAM) {
        plan_stream(h, nq, pl);
        return VQA_OK;
    }
    if (mode == VQA_MODE_FAST_TENSOR) {
        if (!tensor_eligible(h) || !plan_tensor(h, nq, k, pl))
            return fail(VQA_E_UNSUPPORTED,
                        "tensor-core path needs bf16/fp16 rows, dim %% 64 == 0 and a bound index (dim=%d dtype=%d)",
                        h->dim, h->dtype);
        return VQA_OK;
    }
    if (mode == VQA_MODE_FAST_TS) {
        if (!ts_eligible(h) || !plan_ts(h, nq, k, pl))
            return fail(VQA_E_UNSUPPORTED, "TMEM-resident-query path needs bf16/fp16 rows and dim %% 64 == 0, dim <= 1024 "
                                           "(dim <= 768 with the QS variant switched off)");
        return VQA_OK;
    }
    if (mode == VQA_MODE_FAST_PAIR) {
        if (!plan_pair(h, nq, k, pl))
            return fail(VQA_E_UNSUPPORTED, "CTA-pair path needs bf16/fp16 rows, dim %% 64 == 0, dim <= 1024 and "
                                           "k + spare ranks <= 32 (dim=%d k=%d)", h->dim, k);
        return VQA_OK;
    }
    if (mode == VQA_MODE_FAST) {
        // Measured on B200 (profiles/r2_call*.log):
        //  * up to 32 queries, k <= 32: the TMA-fed tcgen05 kernel with the queries resident in shared memory (hi/lo
        //    columns) -- CUDA cores cannot keep up with HBM beyond ~4 queries per streamed element.  Since the warm-up
        //    seed and the dynamic tile schedule it also wins at B = 1, 2 (10 M x 768: 2.07 ms against 2.19 ms for the
        //    CUDA-core streaming kernel; 0.274 against 0.363 ms on the 8-GPU shard), so stream_max_b defaults to 0:
        //    the streaming kernel serves verify mode, fp32 rows and dims that are not multiples of 64;
        //  * beyond that, and k > 32: the TMEM-resident-query kernel serves 128 queries per CTA from one HBM pass
        //    (screen with storage-precision queries, exact re-scoring in the reduce; hi/lo rows + heaps for big k).
        const long long shard_mb = (long long)h->n_rows * h->dim * elem_size(h->dtype) / 1000000;
        const bool sixteen = h->dtype == VQA_BF16 || h->dtype == VQA_F16;
        if (nq <= h->tune.stream_max_b && k <= 32 && (!sixteen || shard_mb >= h->tune.stream_min_mb)) {
            plan_stream(h, nq, pl);
            return VQA_OK;
        }
        //  * more than 128 queries (tune.pair): CTA pairs, cta_group::2 MMAs -- the tensor-bound regime.
        //    tune.wide: the same 128-document tiles on single CTAs for 33..128 queries -- measured (profiles/r2_call7.log)
        //    at dim 768: B = 64 / 128 2.31 / 2.46 -> 2.23 / 2.28 ms on 10 M rows, 0.341 / 0.361 -> 0.335 / 0.348 ms on
        //    the 8-GPU shard; at dim 1024 (8 of 16 query blocks become shared-memory operands) it loses 2-5 %: dim <= 768
        if (((h->tune.pair && nq > 128) || (h->tune.wide && nq > 32 && nq <= 128 && h->dim <= 768)) &&
            plan_pair(h, nq, k, pl))
            return VQA_OK;
        if ((nq > 32 || k > 32) && ts_eligible(h) && plan_ts(h, nq, k, pl)) return VQA_OK;
        if (tensor_eligible(h) && plan_tensor(h, nq, k, pl)) return VQA_OK;
        plan_stream(h, nq, pl);
        return VQA_OK;
    }
    return fail(VQA_E_INVALID, "unknown mode %d", mode);
}

static_assert(sizeof(Plan) <= 16 * sizeof(int), "Plan must fit a plan-cache slot");

// plans are pure functions of (handle shape, tuning, nq, k, mode): cached per handle, invalidated by bind / set_tuning
int make_plan(const vqa_index *h, int nq, int k, int mode, Plan *pl) {
    {
        std::lock_guard<std::mutex> lk(h->mu);
        for (const auto &sl : h->plan_cache)
            if (sl.mode == mode && sl.nq == nq && sl.k == k) {
                std::memcpy(pl, sl.plan, sizeof(Plan));
                return VQA_OK;
            }
    }
    int rc = make_plan_uncached(h, nq, k, mode, pl);
    if (rc) return rc;
    std::lock_guard<std::mutex> lk(h->mu);
    auto &sl = h->plan_cache[h->plan_next++ % 8];
    sl.nq = nq;
    sl.k = k;
    sl.mode = mode;
    std::memcpy(sl.plan, pl, sizeof(Plan));
    return VQA_OK;
}

void invalidate_plans(vqa_index *h) {
    std::lock_guard<std::mutex> lk(h->mu);
    for (auto &sl : h->plan_cache) sl.mode = -1;
    std::memset(h->max_clusters, 0, sizeof(h->max_clusters));
}

int check_search_args(const vqa_index *h, int nq, int k) {
    if (!h) return fail(VQA_E_INVALID, "null index handle");
    if (nq < 1) return fail(VQA_E_INVALID, "n_queries must be >= 1 (got %d)", nq);
    if (k < 1 || k > vqa::kMaxK) return fail(VQA_E_INVALID, "k must be in [1, %d] (got %d)", vqa::kMaxK, k);
    return VQA_OK;
}

// candidate lists are kept at least 32 wide so that the screen-then-rescore path can over-fetch; beyond the
// register lists the big-k screen mode keeps k + spare candidates per list
size_t cand_elems(const vqa_index *h, int nq, int k) {
    const int spare = spare_ranks(h);
    int w = k + spare <= 32 ? 32 : k + spare;
    if (w > vqa::kMaxK) w = k < 32 ? 32 : k;
    return (size_t)h->sm_count * nq * w;
}

int build_tmaps(vqa_index *h) {
    h->tmap_ok = false;
    const int es = elem_size(h->dtype);
    if (h->n_rows > 0 && h->rows && es == 2 && h->dim % 64 == 0) {
        EncodeTiledFn enc = get_encode_fn();
        if (enc) {
            bool ok = true;
            for (int j = 0; j < 4 && ok; ++j) {
                cuuint64_t gdim[2] = {(cuuint64_t)h->dim, (cuuint64_t)h->n_rows};
                cuuint64_t gstride[1] = {(cuuint64_t)h->stride};
                cuuint32_t box[2] = {(cuuint32_t)vqa::kBlockK, (cuuint32_t)(vqa::kTileRows >> j)};
                cuuint32_t estr[2] = {1, 1};
                CUresult r = enc(&h->tmap[j], h->dtype == VQA_BF16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16
                                                                   : CU_TENSOR_MAP_DATA_TYPE_FLOAT16,
                                 2, const_cast<void *>(h->rows), gdim, gstride, box, estr,
                                 CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                                 (CUtensorMapL2promotion)h->tune.tma_l2promo, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
                ok = (r == CUDA_SUCCESS);
            }
            h->tmap_ok = ok;
        }
    }
    return VQA_OK;
}

}  // namespace

// =================================================================================
extern "C" {

int vqa_version(void) { return VQA_VERSION; }

const char *vqa_last_error(void) { return g_err.c_str(); }

int vqa_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) {
        (void)cudaGetLastError();
        return 0;
    }
    return n;
}

int vqa_index_create(vqa_index_t **out, int64_t n_rows, int32_t dim, int32_t dtype, int32_t device,
                     int64_t first_global_id) {
    if (!out) return fail(VQA_E_INVALID, "out is null");
    *out = nullptr;
    const int es = elem_size(dtype);
    if (!es) return fail(VQA_E_INVALID, "row dtype must be F32, BF16 or F16 (got %d)", dtype);
    if (n_rows < 0 || n_rows > 0x7fffffffLL)
        return fail(VQA_E_INVALID, "n_rows per shard must be in [0, 2^31) (got %lld)", (long long)n_rows);
    if (dim < 1 || ((int64_t)dim * es) % 16 != 0)
        return fail(VQA_E_INVALID, "dim*sizeof(dtype) must be a positive multiple of 16 bytes (dim=%d)", dim);
    if (dim > 8192) return fail(VQA_E_UNSUPPORTED, "dim > 8192 not supported (dim=%d)", dim);
    int ndev = vqa_device_count();
    if (ndev == 0) return fail(VQA_E_CUDA, "no CUDA device available (this engine has no CPU fallback)");
    if (device < 0 || device >= ndev) return fail(VQA_E_INVALID, "device %d out of range [0,%d)", device, ndev);
    vqa_index *h = new (std::nothrow) vqa_index();
    if (!h) return fail(VQA_E_NOMEM, "host allocation failed");
    h->n_rows = n_rows;
    h->dim = dim;
    h->dtype = dtype;
    h->device = device;
    h->first_id = first_global_id;
    cudaDeviceProp prop;
    cudaError_t e = cudaGetDeviceProperties(&prop, device);
    if (e != cudaSuccess) {
        delete h;
        return fail(VQA_E_CUDA, "cudaGetDeviceProperties: %s", cudaGetErrorString(e));
    }
    if (prop.major < 10) {
        delete h;
        return fail(VQA_E_UNSUPPORTED, "device %d is sm_%d%d; this library is built for sm_100a only", device,
                    prop.major, prop.minor);
    }
    h->sm_count = prop.multiProcessorCount;
    h->max_smem = (int)prop.sharedMemPerBlockOptin;
    // the ONLY place the search path's knobs meet the environment: once per handle, validated
    int rc = tuning_from_env(&h->tune);
    if (rc) {
        delete h;
        return rc;
    }
    *out = h;
    return VQA_OK;
}

int vqa_tuning_default(vqa_tuning_t *t) {
    if (!t) return fail(VQA_E_INVALID, "tuning is null");
    tuning_defaults(t);
    return VQA_OK;
}

int vqa_tuning_from_env(vqa_tuning_t *t) {
    if (!t) return fail(VQA_E_INVALID, "tuning is null");
    return tuning_from_env(t);
}

int vqa_index_set_tuning(vqa_index_t *h, const vqa_tuning_t *t) {
    if (!h) return fail(VQA_E_INVALID, "null index handle");
    int rc = tuning_validate(t);
    if (rc) return rc;
    const bool remap = t->tma_l2promo != h->tune.tma_l2promo;
    h->tune = *t;
    invalidate_plans(h);
    if (remap && h->rows) build_tmaps(h);
    return VQA_OK;
}

int vqa_index_get_tuning(const vqa_index_t *h, vqa_tuning_t *t) {
    if (!h || !t) return fail(VQA_E_INVALID, "null argument");
    *t = h->tune;
    return VQA_OK;
}

int vqa_index_bind(vqa_index_t *h, const void *rows_dev, int64_t n_rows, int64_t row_stride_bytes) {
    if (!h) return fail(VQA_E_INVALID, "null index handle");
    const int es = elem_size(h->dtype);
    if (n_rows != h->n_rows)
        return fail(VQA_E_INVALID, "n_rows mismatch: created with %lld, bound %lld", (long long)h->n_rows,
                    (long long)n_rows);
    if (n_rows > 0 && !rows_dev) return fail(VQA_E_INVALID, "rows_dev is null");
    if (row_stride_bytes < (int64_t)h->dim * es || row_stride_bytes % 16 != 0)
        return fail(VQA_E_INVALID, "row_stride_bytes must be >= dim*sizeof and a multiple of 16 (got %lld)",
                    (long long)row_stride_bytes);
    if (reinterpret_cast<uintptr_t>(rows_dev) % 16 != 0) return fail(VQA_E_INVALID, "rows_dev must be 16-byte aligned");
    h->rows = rows_dev;
    h->stride = row_stride_bytes;
    invalidate_plans(h);
    return build_tmaps(h);
}

int vqa_debug_timeline(vqa_index_t *h, void *stamps_dev, size_t bytes) {
    if (!h) return fail(VQA_E_INVALID, "null index handle");
    if (stamps_dev && (bytes < 256 || reinterpret_cast<uintptr_t>(stamps_dev) % 8 != 0))
        return fail(VQA_E_INVALID, "timeline buffer must be 8-byte aligned and hold >= 256 bytes per CTA");
    h->timeline = static_cast<unsigned long long *>(stamps_dev);
    h->timeline_bytes = stamps_dev ? bytes : 0;
    return VQA_OK;
}

int vqa_index_destroy(vqa_index_t *h) {
    delete h;
    return VQA_OK;
}

int vqa_search_plan(const vqa_index_t *h, int32_t n_queries, int32_t k, int32_t mode, int32_t *family,
                    int32_t *n_launches) {
    int rc = check_search_args(h, n_queries, k);
    if (rc) return rc;
    Plan pl;
    rc = make_plan(h, n_queries, k, mode, &pl);
    if (rc) return rc;
    if (family) *family = pl.family;
    if (n_launches)
        *n_launches = (pl.family == VQA_MODE_FAST_TENSOR || pl.family == VQA_MODE_FAST_TS ||
                       pl.family == VQA_MODE_FAST_PAIR)
                          ? 2 * ((pl.passes + pl.groups - 1) / pl.groups)
                          : 2;  // stream family: one scan launch (grid.y = passes) + one reduce
    return VQA_OK;
}

int vqa_plan_describe_tuned(int64_t n_rows, int32_t dim, int32_t dtype, int32_t n_queries, int32_t k, int32_t mode,
                            int32_t sm_count, int32_t max_smem, const vqa_tuning_t *tuning, int32_t *out,
                            size_t *smem_bytes) {
    if (!out || !smem_bytes) return fail(VQA_E_INVALID, "null output pointer");
    const int es = elem_size(dtype);
    if (!es) return fail(VQA_E_INVALID, "row dtype must be F32, BF16 or F16 (got %d)", dtype);
    if (n_rows < 0 || n_rows > 0x7fffffffLL || dim < 1 || ((int64_t)dim * es) % 16 != 0 || dim > 8192)
        return fail(VQA_E_INVALID, "bad index shape");
    if (sm_count < 1 || max_smem < 1) return fail(VQA_E_INVALID, "bad device description");
    vqa_index fake;
    fake.n_rows = n_rows;
    fake.dim = dim;
    fake.dtype = dtype;
    fake.sm_count = sm_count;
    fake.max_smem = max_smem;
    fake.tmap_ok = n_rows > 0 && es == 2 && dim % 64 == 0;  // what vqa_index_bind would have built
    int rc = tuning ? tuning_validate(tuning) : tuning_from_env(&fake.tune);
    if (rc) return rc;
    if (tuning) fake.tune = *tuning;
    rc = check_search_args(&fake, n_queries, k);
    if (rc) return rc;
    Plan pl;
    rc = make_plan_uncached(&fake, n_queries, k, mode, &pl);
    if (rc) return rc;
    for (int i = 0; i < 16; ++i) out[i] = 0;
    out[0] = pl.family;
    out[1] = pl.pass_nq;
    out[2] = pl.passes;
    out[3] = pl.groups;
    out[4] = pl.stages;
    out[5] = pl.kps;
    out[6] = pl.ncol;
    *smem_bytes = 0;
    if (pl.family == VQA_MODE_FAST_TS) {
        const int kscan = pl.ts_split ? k : k + spare_ranks(&fake);
        const int nq_launch = n_queries < pl.groups * pl.pass_nq ? n_queries : pl.groups * pl.pass_nq;
        out[7] = pl.ts_split;
        out[8] = pl.ts_qs;
        out[9] = pl.ts_ks;
        out[10] = kscan;
        out[11] = pl.ts_split ? kscan : (kscan > 32 ? vqa::kMaxK : 32);
        out[12] = pl.ts_split ? 0 : 1;
        out[13] = (dim / vqa::kBlockK - pl.ts_ks) * (vqa::kBlockK / 2);  // query block; the rest are accumulators
        out[14] = pl.ts_m64;
        *smem_bytes = vqa::ts_smem_bytes(kscan, pl.stages * pl.kps, pl.ts_split, pl.ts_ks, nq_launch, pl.ts_qs, pl.ts_m64);
    } else if (pl.family == VQA_MODE_FAST_PAIR) {
        const int kscan = k + spare_ranks(&fake);
        out[7] = 0;
        out[8] = 1;
        out[9] = pl.ts_ks;
        out[10] = kscan;
        out[11] = 32;
        out[12] = 1;
        out[13] = (dim / vqa::kBlockK - pl.ts_ks) * (vqa::kBlockK / 2);  // query block; the rest: 128-column accumulators
        *smem_bytes = vqa::pair_smem_bytes(pl.stages * pl.kps * (n_queries <= 128 ? 2 : 1), pl.ts_ks);
    } else if (pl.family == VQA_MODE_FAST_TENSOR) {
        const int kscan = pl.ss_split ? k : k + spare_ranks(&fake);
        out[7] = pl.ss_split;
        out[10] = kscan;
        out[11] = pl.ss_split ? kscan : 32;
        out[12] = pl.ss_split ? 0 : 1;
        *smem_bytes = vqa::mma_smem_bytes_rt(pl.ncol, dim, kscan, pl.stages * pl.kps, pl.ss_split);
    } else {
        out[10] = k;
        out[11] = k;
    }
    return VQA_OK;
}

int vqa_plan_describe(int64_t n_rows, int32_t dim, int32_t dtype, int32_t n_queries, int32_t k, int32_t mode,
                      int32_t sm_count, int32_t max_smem, int32_t *out, size_t *smem_bytes) {
    return vqa_plan_describe_tuned(n_rows, dim, dtype, n_queries, k, mode, sm_count, max_smem, nullptr, out, smem_bytes);
}

int vqa_workspace_bytes(const vqa_index_t *h, int32_t n_queries, int32_t k, int32_t mode, size_t *bytes) {
    int rc = check_search_args(h, n_queries, k);
    if (rc) return rc;
    if (!bytes) return fail(VQA_E_INVALID, "bytes is null");
    (void)mode;
    // candidates: per CTA, per query, k entries of (float score, u32 row); + one shared-threshold slot per query
    // + 64 tile counters of the dynamic tile schedule (one per scan launch of a search), 32 warm-up seed slots per
    // query + alignment slack
    *bytes = cand_elems(h, n_queries, k) * 8 + (size_t)n_queries * 8 + 64 * 8 + (size_t)n_queries * 32 * 8 + 1024;
    return VQA_OK;
}

static int search_impl(const vqa_index_t *h, const float *queries_dev, int64_t q_stride, int32_t n_queries, int32_t k,
                       int32_t mode, float *out_scores_dev, int64_t *out_ids_dev, void *workspace_dev,
                       size_t workspace_bytes, void *stream, void *reduce_stream);

int vqa_search(const vqa_index_t *h, const float *queries_dev, int64_t q_stride, int32_t n_queries, int32_t k,
               int32_t mode, float *out_scores_dev, int64_t *out_ids_dev, void *workspace_dev,
               size_t workspace_bytes, void *stream) {
    return search_impl(h, queries_dev, q_stride, n_queries, k, mode, out_scores_dev, out_ids_dev, workspace_dev,
                       workspace_bytes, stream, nullptr);
}

int vqa_search_2s(const vqa_index_t *h, const float *queries_dev, int64_t q_stride, int32_t n_queries, int32_t k,
                  int32_t mode, float *out_scores_dev, int64_t *out_ids_dev, void *workspace_dev,
                  size_t workspace_bytes, void *scan_stream, void *reduce_stream) {
    if (!reduce_stream || reduce_stream == scan_stream)
        return fail(VQA_E_INVALID, "vqa_search_2s needs a reduce stream different from the scan stream");
    return search_impl(h, queries_dev, q_stride, n_queries, k, mode, out_scores_dev, out_ids_dev, workspace_dev,
                       workspace_bytes, scan_stream, reduce_stream);
}

static int search_impl(const vqa_index_t *h, const float *queries_dev, int64_t q_stride, int32_t n_queries, int32_t k,
                       int32_t mode, float *out_scores_dev, int64_t *out_ids_dev, void *workspace_dev,
                       size_t workspace_bytes, void *stream, void *reduce_stream) {
    int rc = check_search_args(h, n_queries, k);
    if (rc) return rc;
    if (!queries_dev || !out_scores_dev || !out_ids_dev || !workspace_dev)
        return fail(VQA_E_INVALID, "null device pointer argument");
    if (h->n_rows > 0 && !h->rows) return fail(VQA_E_INVALID, "index has no bound rows (call vqa_index_bind)");
    if (q_stride < h->dim || q_stride % 4 != 0 || reinterpret_cast<uintptr_t>(queries_dev) % 16 != 0)
        return fail(VQA_E_INVALID, "queries must be 16-byte aligned with q_stride >= dim and q_stride %% 4 == 0");
    size_t need = 0;
    vqa_workspace_bytes(h, n_queries, k, mode, &need);
    if (workspace_bytes < need)
        return fail(VQA_E_NOMEM, "workspace too small: need %zu bytes, got %zu", need, workspace_bytes);
    Plan pl;
    rc = make_plan(h, n_queries, k, mode, &pl);
    if (rc) return rc;

    DeviceGuard guard(h->device);
    if (guard.err != cudaSuccess) return fail(VQA_E_CUDA, "cudaSetDevice: %s", cudaGetErrorString(guard.err));
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    cudaStream_t rst = reinterpret_cast<cudaStream_t>(reduce_stream);
    // Two-stream form: every candidate reduce is enqueued on `rst` behind an event recorded after its scan, so the
    // scan stream is free for the NEXT search's scan the moment this one ends (the caller gives consecutive searches
    // distinct workspaces and outputs).  Returns the stream the reduce must be launched on.
    cudaError_t handoff_err = cudaSuccess;
    auto reduce_on = [&]() -> cudaStream_t {
        if (!rst) return st;
        std::lock_guard<std::mutex> lk(h->mu);
        if (!h->handoff) handoff_err = cudaEventCreateWithFlags(&h->handoff, cudaEventDisableTiming);
        if (handoff_err == cudaSuccess) handoff_err = cudaEventRecord(h->handoff, st);
        if (handoff_err == cudaSuccess) handoff_err = cudaStreamWaitEvent(rst, h->handoff, 0);
        return rst;
    };

    uintptr_t ws = (reinterpret_cast<uintptr_t>(workspace_dev) + 255) & ~(uintptr_t)255;
    float *cand_s = reinterpret_cast<float *>(ws);
    uint32_t *cand_i = reinterpret_cast<uint32_t *>(cand_s + cand_elems(h, n_queries, k));
    unsigned long long *tau_g = reinterpret_cast<unsigned long long *>(
        (reinterpret_cast<uintptr_t>(cand_i + cand_elems(h, n_queries, k)) + 255) & ~(uintptr_t)255);
    unsigned long long *tile_ctr = reinterpret_cast<unsigned long long *>(
        (reinterpret_cast<uintptr_t>(tau_g + n_queries) + 255) & ~(uintptr_t)255);  // [64], one per scan launch
    unsigned long long *slot_g = tile_ctr + 64;  // [n_queries][32]
    const long long cand_stride = (long long)n_queries * k;
    vqa::ReduceOpts ropts;
    ropts.no_pdl = rst ? 1 : 0;   // (programmatic launch relaxes the order against the previous KERNEL of a stream only)
    ropts.select = h->tune.reduce_select;
    ropts.early = h->tune.reduce_early;
    ropts.trigger_early = h->tune.pdl_chain;
    static std::atomic<uint32_t> g_epoch{1};
    const uint32_t epoch = g_epoch.fetch_add(2, std::memory_order_relaxed);  // odd, unique, never 0 (0 = cleared slot)

    // TMEM-resident-query kernel over the queries [qb, qe) of this search (the whole batch, or the tail a CTA-pair
    // launch leaves over)
    auto run_ts = [&](const Plan &pl, int qb, int qe) -> int {
        // list length inside the scan: with screen-then-rescore a few spare ranks absorb the reordering
        // that the queries' storage rounding can cause (score error ~5e-5 against rank gaps of ~7e-4)
        const int kscan = pl.ts_split ? k : k + spare_ranks(h);
        const long long cstride = (long long)n_queries * kscan;
        const int per_launch = pl.groups * pl.pass_nq;
        const long long tiles = (h->n_rows + 63) / 64;
        for (int l0 = qb; l0 < qe; l0 += per_launch) {
            const int nq = qe - l0 < per_launch ? qe - l0 : per_launch;
            const int chunks = (nq + pl.pass_nq - 1) / pl.pass_nq;
            int g = 1, lg = 0;
            while (g < chunks) {
                g <<= 1;
                ++lg;
            }
            const bool mc = g > 1;
            long long streams = h->sm_count / g;
            if (mc && g == 4 && streams > 32) streams = 32;  // clusters of 4 do not tile every GPC
            if (streams > tiles) streams = tiles;
            if (streams < 1) streams = 1;
            vqa::TsLaunch a;
            a.tmap = &h->tmap[1 + lg];
            a.bf16 = h->dtype == VQA_BF16;
            a.split = pl.ts_split;
            a.a_fp16 = pl.ts_afp16;
            a.qs = pl.ts_qs;
            a.ks = pl.ts_ks;
            a.m64 = pl.ts_m64;
            a.pdl = (l0 > qb && h->tune.pdl_chain) ? 1 : 0;
            a.stages = pl.stages;
            a.kps = pl.kps;
            a.grid = (int)streams * g;
            a.n_groups = g;
            a.multicast = mc ? 1 : 0;
            a.q = queries_dev + (long long)l0 * q_stride;
            a.q_stride = q_stride;
            a.nq = nq;
            a.k = kscan;
            a.n_rows = h->n_rows;
            a.dim = h->dim;
            a.cand_s = cand_s + (long long)l0 * kscan;
            a.cand_i = cand_i + (long long)l0 * kscan;
            a.cand_stride = cstride;
            a.tau_g = tau_g + l0;
            a.epoch = epoch;
            a.timeline = (h->timeline && h->timeline_bytes >= (size_t)a.grid * 32 * 8) ? h->timeline : nullptr;
            cudaError_t e = vqa::launch_ts(a, st);
            if (e != cudaSuccess) return fail(VQA_E_CUDA, "TS scan launch failed: %s", cudaGetErrorString(e));
            vqa::Rescore rs;
            rs.rows = h->rows;
            rs.stride = h->stride;
            rs.dim = h->dim;
            rs.bf16 = h->dtype == VQA_BF16;
            rs.q = queries_dev + (long long)l0 * q_stride;
            rs.q_stride = q_stride;
            rs.k_final = k;
            e = vqa::launch_reduce_u32(cand_s + (long long)l0 * kscan, cand_i + (long long)l0 * kscan, cstride, kscan, a.grid,
                                       kscan, pl.ts_split ? kscan : (kscan > 32 ? vqa::kMaxK : 32), h->first_id,
                                       out_scores_dev + (long long)l0 * k,
                                       (long long *)out_ids_dev + (long long)l0 * k, nq, tau_g + l0, g, pl.pass_nq, reduce_on(),
                                       ropts, pl.ts_split ? nullptr : &rs);
            if (e != cudaSuccess || handoff_err != cudaSuccess)
                return fail(VQA_E_CUDA, "reduce launch failed: %s", cudaGetErrorString(e != cudaSuccess ? e : handoff_err));
        }
        return VQA_OK;
    };
    if (pl.family == VQA_MODE_FAST_TS && h->n_rows > 0) return run_ts(pl, 0, n_queries);

    if (pl.family == VQA_MODE_FAST_PAIR && h->n_rows > 0) {
        // 128-document tiles: launches of up to 256 queries on CTA pairs (cta_group::2); a launch of <= 128 queries
        // (a small batch, or the tail of a large one) runs the same tiles on single CTAs
        const int kscan = k + spare_ranks(h);
        const long long cstride = (long long)n_queries * kscan;
        const long long tiles = (h->n_rows + vqa::kTileRows - 1) / vqa::kTileRows;
        for (int l0 = 0; l0 < n_queries; l0 += 256) {
            const int nq = n_queries - l0 < 256 ? n_queries - l0 : 256;
            const bool single = nq <= 128;
            int stages = pl.stages, kps = pl.kps;
            if (single != (n_queries <= 128) && !pair_geometry(h, pl.ts_ks, single, &stages, &kps))
                return fail(VQA_E_UNSUPPORTED, "no ring geometry for the tail launch of a CTA-pair search");
            long long streams = single ? h->sm_count : h->sm_count / 2;
            if (streams > tiles) streams = tiles;
            if (streams < 1) streams = 1;
            vqa::PairLaunch a;
            a.tmap = &h->tmap[1];
            a.bf16 = h->dtype == VQA_BF16;
            a.stages = stages;
            a.kps = kps;
            a.pair = single ? 0 : 1;
            a.grid = (int)streams * (single ? 1 : 2);
            a.q = queries_dev + (long long)l0 * q_stride;
            a.q_stride = q_stride;
            a.nq = nq;
            a.k = kscan;
            a.n_rows = h->n_rows;
            a.dim = h->dim;
            a.cand_s = cand_s + (long long)l0 * kscan;
            a.cand_i = cand_i + (long long)l0 * kscan;
            a.cand_stride = cstride;
            a.tau_g = tau_g + l0;
            a.epoch = epoch;
            a.ks = pl.ts_ks;
            a.timeline = (h->timeline && h->timeline_bytes >= (size_t)a.grid * 32 * 8) ? h->timeline : nullptr;
            cudaError_t e = vqa::launch_pair(a, st);
            if (e != cudaSuccess) return fail(VQA_E_CUDA, "128-document-tile scan launch failed: %s", cudaGetErrorString(e));
            vqa::Rescore rs;
            rs.rows = h->rows;
            rs.stride = h->stride;
            rs.dim = h->dim;
            rs.bf16 = h->dtype == VQA_BF16;
            rs.q = queries_dev + (long long)l0 * q_stride;
            rs.q_stride = q_stride;
            rs.k_final = k;
            e = vqa::launch_reduce_u32(cand_s + (long long)l0 * kscan, cand_i + (long long)l0 * kscan, cstride, kscan, a.grid,
                                       kscan, 32, h->first_id, out_scores_dev + (long long)l0 * k,
                                       (long long *)out_ids_dev + (long long)l0 * k, nq, tau_g + l0, single ? 1 : 2, 128, reduce_on(),
                                       ropts, &rs);
            if (e != cudaSuccess || handoff_err != cudaSuccess)
                return fail(VQA_E_CUDA, "reduce launch failed: %s", cudaGetErrorString(e != cudaSuccess ? e : handoff_err));
        }
        return VQA_OK;
    }

    if (pl.family == VQA_MODE_FAST_TENSOR && h->n_rows > 0) {
        // Each launch covers up to groups * pass_nq queries: CTA c scans tile stream c / g for query chunk
        // c % g, so one pass over HBM serves the whole launch; its candidate lists are reduced right away.
        const int kscan = pl.ss_split ? k : k + spare_ranks(h);  // list length inside the scan
        const long long cstride = (long long)n_queries * kscan;
        const int per_launch = pl.groups * pl.pass_nq;
        const long long tiles = (h->n_rows + vqa::kTileRows - 1) / vqa::kTileRows;
        const bool use_mc = h->tune.mma_multicast != 0;
        for (int l0 = 0; l0 < n_queries; l0 += per_launch) {
            const int nq = n_queries - l0 < per_launch ? n_queries - l0 : per_launch;
            const int chunks = (nq + pl.pass_nq - 1) / pl.pass_nq;
            int g = chunks, lg = 0;
            if (use_mc && chunks > 1) {  // cluster sizes are powers of two; a short launch gets empty chunks
                while ((1 << lg) < chunks) ++lg;
                g = 1 << lg;
            }
            const bool mc = use_mc && g > 1;
            long long streams = h->sm_count / g;
            if (mc) {
                const int slot = pl.ncol == 16 ? 0 : (pl.ncol == 32 ? 1 : (pl.ncol == 64 ? 2 : 3));
                int cached;
                {
                    std::lock_guard<std::mutex> lk(h->mu);
                    cached = h->max_clusters[lg][slot];
                }
                if (cached == 0) {
                    cached = vqa::mma_max_active_clusters(
                        h->dtype == VQA_BF16, pl.ncol, pl.ss_split, g,
                        vqa::mma_smem_bytes_rt(pl.ncol, h->dim, kscan, pl.stages * pl.kps, pl.ss_split));
                    if (cached <= 0) cached = -1;
                    std::lock_guard<std::mutex> lk(h->mu);
                    h->max_clusters[lg][slot] = cached;
                }
                if (cached > 0 && cached < streams) streams = cached;
            }
            if (streams > tiles) streams = tiles;
            if (streams < 1) streams = 1;
            vqa::MmaLaunch a;
            a.tmap = &h->tmap[mc ? lg : 0];
            a.bf16 = h->dtype == VQA_BF16;
            a.ncol = pl.ncol;
            a.split = pl.ss_split;
            a.stages = pl.stages;
            a.kps = pl.kps;
            a.grid = (int)streams * g;
            a.n_groups = g;
            a.multicast = mc ? 1 : 0;
            a.q = queries_dev + (long long)l0 * q_stride;
            a.q_stride = q_stride;
            a.nq = nq;
            a.k = kscan;
            a.n_rows = h->n_rows;
            a.dim = h->dim;
            a.cand_s = cand_s + (long long)l0 * kscan;
            a.cand_i = cand_i + (long long)l0 * kscan;
            a.cand_stride = cstride;
            a.tau_g = tau_g + l0;
            a.epoch = epoch;
            a.pdl = (l0 > 0 && h->tune.pdl_chain) ? 1 : 0;
            a.tma_hint = h->tune.tma_hint;
            a.tile_ctr = h->tune.dyn_tiles ? tile_ctr + (l0 / per_launch) % 64 : nullptr;
            a.slot_g = (h->tune.seed && pl.pass_nq <= 32 && kscan <= 32) ? slot_g + (long long)l0 * 32 : nullptr;
            a.timeline = (h->timeline && h->timeline_bytes >= (size_t)a.grid * 32 * 8) ? h->timeline : nullptr;
            cudaError_t e = vqa::launch_mma(a, st);
            if (e != cudaSuccess) return fail(VQA_E_CUDA, "tensor scan launch failed: %s", cudaGetErrorString(e));
            vqa::Rescore rs;
            rs.rows = h->rows;
            rs.stride = h->stride;
            rs.dim = h->dim;
            rs.bf16 = h->dtype == VQA_BF16;
            rs.q = queries_dev + (long long)l0 * q_stride;
            rs.q_stride = q_stride;
            rs.k_final = k;
            e = vqa::launch_reduce_u32(cand_s + (long long)l0 * kscan, cand_i + (long long)l0 * kscan, cstride, kscan, a.grid,
                                       kscan, pl.ss_split ? kscan : 32, h->first_id, out_scores_dev + (long long)l0 * k,
                                       (long long *)out_ids_dev + (long long)l0 * k, nq, tau_g + l0, g, pl.pass_nq, reduce_on(),
                                       ropts, pl.ss_split ? nullptr : &rs, a.slot_g);
            if (e != cudaSuccess || handoff_err != cudaSuccess)
                return fail(VQA_E_CUDA, "reduce launch failed: %s", cudaGetErrorString(e != cudaSuccess ? e : handoff_err));
        }
        return VQA_OK;
    }

    int n_lists = 0;
    if (h->n_rows > 0) {
        // All passes of <= pass_nq queries go out as ONE launch (grid.y = passes).  The SMs are divided
        // among the passes so that they run side by side in one wave -- on a small index no SM sits through
        // a CTA's fixed cost for nothing, on a large one the passes stream the same rows together and
        // share them through L2 -- and no CTA gets fewer than ~512 rows.
        const int passes_total = (n_queries + pl.pass_nq - 1) / pl.pass_nq;
        long long gx = (h->n_rows + 511) / 512;
        if (gx > pl.grid) gx = pl.grid;
        if (passes_total > 1 && gx > pl.grid / passes_total) gx = pl.grid / passes_total;
        if (gx < 1) gx = 1;
        pl.grid = (int)gx;
        n_lists = pl.grid;
        const int per_launch = pl.pass_nq * 32768;
        for (int p0 = 0; p0 < n_queries; p0 += per_launch) {
            const int nq = n_queries - p0 < per_launch ? n_queries - p0 : per_launch;
            vqa::ScanLaunch a;
            a.dtype = h->dtype;
            a.bt = pl.pass_nq;
            a.grid = pl.grid;
            a.rows = h->rows;
            a.n_rows = h->n_rows;
            a.row_stride_bytes = h->stride;
            a.dim = h->dim;
            a.q = queries_dev + (long long)p0 * q_stride;
            a.q_stride = q_stride;
            a.nq = nq;
            a.k = k;
            a.cand_s = cand_s + (long long)p0 * k;
            a.cand_i = cand_i + (long long)p0 * k;
            a.cand_stride = cand_stride;
            cudaError_t e = vqa::launch_scan(a, st);
            if (e != cudaSuccess) return fail(VQA_E_CUDA, "scan launch failed: %s", cudaGetErrorString(e));
        }
    }
    cudaError_t e = vqa::launch_reduce_u32(cand_s, cand_i, cand_stride, k, n_lists, k, k, h->first_id, out_scores_dev,
                                           (long long *)out_ids_dev, n_queries, nullptr, 1, 1, reduce_on(), ropts);
    if (e != cudaSuccess || handoff_err != cudaSuccess)
                return fail(VQA_E_CUDA, "reduce launch failed: %s", cudaGetErrorString(e != cudaSuccess ? e : handoff_err));
    return VQA_OK;
}

int vqa_search_host_staging_bytes(const vqa_index_t *h, int32_t n_queries, int32_t k, int32_t mode,
                                  size_t *bytes) {
    size_t ws = 0;
    int rc = vqa_workspace_bytes(h, n_queries, k, mode, &ws);
    if (rc) return rc;
    size_t q = ((size_t)n_queries * h->dim * 4 + 255) & ~(size_t)255;
    size_t os = ((size_t)n_queries * k * 4 + 255) & ~(size_t)255;
    size_t oi = ((size_t)n_queries * k * 8 + 255) & ~(size_t)255;
    *bytes = q + os + oi + ws + 256;
    return VQA_OK;
}

static int search_host_impl(const vqa_index_t *h, const float *queries_host, int32_t n_queries, int32_t k, int32_t mode,
                            float *out_scores_host, int64_t *out_ids_host, void *staging_dev, size_t staging_bytes,
                            void *stream, bool sync);

int vqa_search_host(const vqa_index_t *h, const float *queries_host, int32_t n_queries, int32_t k, int32_t mode,
                    float *out_scores_host, int64_t *out_ids_host, void *staging_dev, size_t staging_bytes,
                    void *stream) {
    return search_host_impl(h, queries_host, n_queries, k, mode, out_scores_host, out_ids_host, staging_dev,
                            staging_bytes, stream, true);
}

int vqa_search_host_async(const vqa_index_t *h, const float *queries_host, int32_t n_queries, int32_t k, int32_t mode,
                          float *out_scores_host, int64_t *out_ids_host, void *staging_dev, size_t staging_bytes,
                          void *stream) {
    return search_host_impl(h, queries_host, n_queries, k, mode, out_scores_host, out_ids_host, staging_dev,
                            staging_bytes, stream, false);
}

static int search_host_impl(const vqa_index_t *h, const float *queries_host, int32_t n_queries, int32_t k, int32_t mode,
                            float *out_scores_host, int64_t *out_ids_host, void *staging_dev, size_t staging_bytes,
                            void *stream, bool sync) {
    size_t need = 0;
    int rc = vqa_search_host_staging_bytes(h, n_queries, k, mode, &need);
    if (rc) return rc;
    if (!queries_host || !out_scores_host || !out_ids_host || !staging_dev)
        return fail(VQA_E_INVALID, "null pointer argument");
    if (staging_bytes < need)
        return fail(VQA_E_NOMEM, "staging too small: need %zu bytes, got %zu", need, staging_bytes);
    DeviceGuard guard(h->device);
    if (guard.err != cudaSuccess) return fail(VQA_E_CUDA, "cudaSetDevice: %s", cudaGetErrorString(guard.err));
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    uintptr_t base = (reinterpret_cast<uintptr_t>(staging_dev) + 255) & ~(uintptr_t)255;
    size_t qb = ((size_t)n_queries * h->dim * 4 + 255) & ~(size_t)255;
    size_t osb = ((size_t)n_queries * k * 4 + 255) & ~(size_t)255;
    size_t oib = ((size_t)n_queries * k * 8 + 255) & ~(size_t)255;
    float *q_dev = reinterpret_cast<float *>(base);
    float *os_dev = reinterpret_cast<float *>(base + qb);
    int64_t *oi_dev = reinterpret_cast<int64_t *>(base + qb + osb);
    void *ws = reinterpret_cast<void *>(base + qb + osb + oib);
    size_t ws_bytes = staging_bytes - (size_t)(reinterpret_cast<uintptr_t>(ws) - reinterpret_cast<uintptr_t>(staging_dev));
    CUDA_TRY(cudaMemcpyAsync(q_dev, queries_host, (size_t)n_queries * h->dim * 4, cudaMemcpyHostToDevice, st));
    rc = vqa_search(h, q_dev, h->dim, n_queries, k, mode, os_dev, oi_dev, ws, ws_bytes, stream);
    if (rc) return rc;
    CUDA_TRY(cudaMemcpyAsync(out_scores_host, os_dev, (size_t)n_queries * k * 4, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaMemcpyAsync(out_ids_host, oi_dev, (size_t)n_queries * k * 8, cudaMemcpyDeviceToHost, st));
    if (sync) CUDA_TRY(cudaStreamSynchronize(st));
    return VQA_OK;
}

int vqa_merge_topk_strided(const float *cand_scores_dev, const int64_t *cand_ids_dev, int64_t list_stride_scores,
                           int64_t list_stride_ids, int32_t n_lists, int32_t n_queries, int32_t k_in, int32_t k_out,
                           float *out_scores_dev, int64_t *out_ids_dev, int32_t device, void *stream) {
    if (!cand_scores_dev || !cand_ids_dev || !out_scores_dev || !out_ids_dev)
        return fail(VQA_E_INVALID, "null device pointer argument");
    if (n_lists < 1 || n_queries < 1 || k_in < 1) return fail(VQA_E_INVALID, "n_lists, n_queries, k_in must be >= 1");
    if (k_out < 1 || k_out > vqa::kMaxK) return fail(VQA_E_INVALID, "k_out must be in [1, %d]", vqa::kMaxK);
    if (list_stride_scores < (int64_t)n_queries * k_in || list_stride_ids < (int64_t)n_queries * k_in)
        return fail(VQA_E_INVALID, "list strides must be >= n_queries * k_in elements");
    if (vqa_device_count() == 0) return fail(VQA_E_CUDA, "no CUDA device available (no CPU fallback)");
    DeviceGuard guard(device);
    if (guard.err != cudaSuccess) return fail(VQA_E_CUDA, "cudaSetDevice: %s", cudaGetErrorString(guard.err));
    cudaError_t e = vqa::launch_reduce_i64(cand_scores_dev, (const long long *)cand_ids_dev, list_stride_scores,
                                           list_stride_ids, k_in, n_lists, k_in, k_out, 0, out_scores_dev,
                                           (long long *)out_ids_dev, n_queries,
                                           reinterpret_cast<cudaStream_t>(stream));
    if (e != cudaSuccess) return fail(VQA_E_CUDA, "merge launch failed: %s", cudaGetErrorString(e));
    return VQA_OK;
}

int vqa_exchange_push(const void *local_block_dev, size_t block_bytes, void *const *peer_slot_ptrs,
                      uint64_t *const *peer_flag_ptrs, int32_t world, uint64_t epoch, int32_t device, void *stream) {
    if (!local_block_dev || !peer_slot_ptrs || !peer_flag_ptrs) return fail(VQA_E_INVALID, "null pointer argument");
    if (world < 1 || world > 16) return fail(VQA_E_INVALID, "world must be in [1, 16] (got %d)", world);
    if (block_bytes == 0 || block_bytes % 16 != 0 || reinterpret_cast<uintptr_t>(local_block_dev) % 16 != 0)
        return fail(VQA_E_INVALID, "block must be 16-byte aligned and a multiple of 16 bytes");
    for (int r = 0; r < world; ++r)
        if (!peer_slot_ptrs[r] || !peer_flag_ptrs[r] || reinterpret_cast<uintptr_t>(peer_slot_ptrs[r]) % 16 != 0)
            return fail(VQA_E_INVALID, "peer pointer %d is null or misaligned", r);
    if (vqa_device_count() == 0) return fail(VQA_E_CUDA, "no CUDA device available (no CPU fallback)");
    DeviceGuard guard(device);
    if (guard.err != cudaSuccess) return fail(VQA_E_CUDA, "cudaSetDevice: %s", cudaGetErrorString(guard.err));
    cudaError_t e = vqa::launch_exchange_push(local_block_dev, block_bytes, peer_slot_ptrs,
                                              reinterpret_cast<unsigned long long *const *>(peer_flag_ptrs), world,
                                              (unsigned long long)epoch, reinterpret_cast<cudaStream_t>(stream));
    if (e != cudaSuccess) return fail(VQA_E_CUDA, "exchange push launch failed: %s", cudaGetErrorString(e));
    return VQA_OK;
}

int vqa_merge_topk_wait(const float *cand_scores_dev, const int64_t *cand_ids_dev, int64_t list_stride_scores,
                        int64_t list_stride_ids, int32_t n_lists, int32_t n_queries, int32_t k_in, int32_t k_out,
                        float *out_scores_dev, int64_t *out_ids_dev, const uint64_t *flags_dev, uint64_t epoch,
                        int32_t device, void *stream) {
    if (!cand_scores_dev || !cand_ids_dev || !out_scores_dev || !out_ids_dev || !flags_dev)
        return fail(VQA_E_INVALID, "null device pointer argument");
    if (n_lists < 1 || n_lists > 32 || n_queries < 1 || k_in < 1)
        return fail(VQA_E_INVALID, "n_lists must be in [1, 32]; n_queries, k_in >= 1");
    if (k_out < 1 || k_out > vqa::kMaxK) return fail(VQA_E_INVALID, "k_out must be in [1, %d]", vqa::kMaxK);
    if (list_stride_scores < (int64_t)n_queries * k_in || list_stride_ids < (int64_t)n_queries * k_in)
        return fail(VQA_E_INVALID, "list strides must be >= n_queries * k_in elements");
    if (vqa_device_count() == 0) return fail(VQA_E_CUDA, "no CUDA device available (no CPU fallback)");
    DeviceGuard guard(device);
    if (guard.err != cudaSuccess) return fail(VQA_E_CUDA, "cudaSetDevice: %s", cudaGetErrorString(guard.err));
    vqa::WaitFlags wf;
    wf.flags = reinterpret_cast<const unsigned long long *>(flags_dev);
    wf.n = n_lists;
    wf.epoch = epoch;
    cudaError_t e = vqa::launch_reduce_i64(cand_scores_dev, (const long long *)cand_ids_dev, list_stride_scores,
                                           list_stride_ids, k_in, n_lists, k_in, k_out, 0, out_scores_dev,
                                           (long long *)out_ids_dev, n_queries, reinterpret_cast<cudaStream_t>(stream),
                                           &wf);
    if (e != cudaSuccess) return fail(VQA_E_CUDA, "merge launch failed: %s", cudaGetErrorString(e));
    return VQA_OK;
}

int vqa_merge_topk(const float *cand_scores_dev, const int64_t *cand_ids_dev, int32_t n_lists, int32_t n_queries,
                   int32_t k_in, int32_t k_out, float *out_scores_dev, int64_t *out_ids_dev, int32_t device,
                   void *stream) {
    return vqa_merge_topk_strided(cand_scores_dev, cand_ids_dev, (int64_t)n_queries * k_in, (int64_t)n_queries * k_in,
                                  n_lists, n_queries, k_in, k_out, out_scores_dev, out_ids_dev, device, stream);
}

int vqa_merge_segments_limits(int32_t *max_k_out, int32_t *max_candidates) {
    if (max_k_out) *max_k_out = vqa::segmerge_max_k();
    if (max_candidates) *max_candidates = vqa::segmerge_max_cand();
    return VQA_OK;
}

int vqa_merge_segments(const float *seg_scores_dev, const int64_t *seg_ids_dev, int32_t n_segments, int32_t n_queries,
                       int32_t k_seg, int32_t k_out, float *out_scores_dev, int64_t *out_ids_dev,
                       int32_t *saturated_dev, int32_t device, void *stream) {
    if (!seg_scores_dev || !seg_ids_dev || !out_scores_dev || !out_ids_dev || !saturated_dev)
        return fail(VQA_E_INVALID, "null device pointer argument");
    if (n_segments < 1 || n_queries < 1 || k_seg < 1)
        return fail(VQA_E_INVALID, "n_segments, n_queries, k_seg must be >= 1");
    if ((int64_t)n_segments * k_seg > vqa::segmerge_max_cand())
        return fail(VQA_E_INVALID, "n_segments * k_seg must be <= %d (got %lld)", vqa::segmerge_max_cand(),
                    (long long)n_segments * k_seg);
    if (k_out < 1 || k_out > vqa::segmerge_max_k())
        return fail(VQA_E_INVALID, "k_out must be in [1, %d] (got %d)", vqa::segmerge_max_k(), k_out);
    if (vqa_device_count() == 0) return fail(VQA_E_CUDA, "no CUDA device available (no CPU fallback)");
    DeviceGuard guard(device);
    if (guard.err != cudaSuccess) return fail(VQA_E_CUDA, "cudaSetDevice: %s", cudaGetErrorString(guard.err));
    cudaError_t e = vqa::launch_merge_segments(seg_scores_dev, (const long long *)seg_ids_dev, n_segments, n_queries,
                                               k_seg, k_out, out_scores_dev, (long long *)out_ids_dev, saturated_dev,
                                               reinterpret_cast<cudaStream_t>(stream));
    if (e != cudaSuccess) return fail(VQA_E_CUDA, "segment merge launch failed: %s", cudaGetErrorString(e));
    return VQA_OK;
}

int vqa_pool_normalize(const void *hidden_dev, int32_t h_dtype, const void *mask_dev, int32_t m_dtype,
                       int32_t batch, int32_t seq, int32_t dim, int32_t normalize, float *out_dev, int32_t device,
                       void *stream) {
    if (!hidden_dev || !mask_dev || !out_dev) return fail(VQA_E_INVALID, "null device pointer argument");
    const int es = elem_size(h_dtype);
    if (!es) return fail(VQA_E_INVALID, "hidden dtype must be F32, BF16 or F16 (got %d)", h_dtype);
    if (m_dtype != VQA_I64 && m_dtype != VQA_I32 && m_dtype != VQA_U8 && m_dtype != VQA_F32)
        return fail(VQA_E_INVALID, "mask dtype must be I64, I32, U8 or F32 (got %d)", m_dtype);
    if (batch < 1 || seq < 1) return fail(VQA_E_INVALID, "batch and seq must be >= 1");
    if (dim < 1 || (dim * es) % 16 != 0 || dim > 8192)
        return fail(VQA_E_INVALID, "dim*sizeof(dtype) must be a multiple of 16 bytes and dim <= 8192 (dim=%d)", dim);
    if (reinterpret_cast<uintptr_t>(hidden_dev) % 16 != 0) return fail(VQA_E_INVALID, "hidden_dev must be 16-byte aligned");
    if (vqa_device_count() == 0) return fail(VQA_E_CUDA, "no CUDA device available (no CPU fallback)");
    DeviceGuard guard(device);
    if (guard.err != cudaSuccess) return fail(VQA_E_CUDA, "cudaSetDevice: %s", cudaGetErrorString(guard.err));
    cudaError_t e = vqa::launch_pool(hidden_dev, h_dtype, mask_dev, m_dtype, batch, seq, dim, normalize, out_dev,
                                     reinterpret_cast<cudaStream_t>(stream));
    if (e != cudaSuccess) return fail(VQA_E_CUDA, "pool launch failed: %s", cudaGetErrorString(e));
    return VQA_OK;
}

int vqa_normalize_rows(const float *in_dev, int64_t in_stride, int64_t n_rows, int32_t dim, float *out_dev,
                       int64_t out_stride, void *cast_out_dev, int32_t cast_dtype, int64_t cast_stride,
                       int32_t device, void *stream) {
    if (!in_dev) return fail(VQA_E_INVALID, "in_dev is null");
    if (!out_dev && !cast_out_dev) return fail(VQA_E_INVALID, "no output buffer given");
    if (n_rows < 0) return fail(VQA_E_INVALID, "n_rows must be >= 0");
    if (dim < 4 || dim % 4 != 0) return fail(VQA_E_INVALID, "dim must be a positive multiple of 4 (dim=%d)", dim);
    if (in_stride < dim || in_stride % 4 != 0 || (out_dev && (out_stride < dim || out_stride % 4 != 0)))
        return fail(VQA_E_INVALID, "row strides must be >= dim and multiples of 4 elements");
    if (reinterpret_cast<uintptr_t>(in_dev) % 16 != 0 || reinterpret_cast<uintptr_t>(out_dev) % 16 != 0)
        return fail(VQA_E_INVALID, "buffers must be 16-byte aligned");
    int cast_kind = 0;
    if (cast_out_dev) {
        if (cast_dtype == VQA_BF16) cast_kind = 1;
        else if (cast_dtype == VQA_F16) cast_kind = 2;
        else return fail(VQA_E_INVALID, "cast dtype must be BF16 or F16 (got %d)", cast_dtype);
        if (cast_stride < dim || cast_stride % 4 != 0 || reinterpret_cast<uintptr_t>(cast_out_dev) % 8 != 0)
            return fail(VQA_E_INVALID, "cast buffer must be 8-byte aligned with stride >= dim, multiple of 4");
    }
    if (vqa_device_count() == 0) return fail(VQA_E_CUDA, "no CUDA device available (no CPU fallback)");
    if (n_rows == 0) return VQA_OK;
    DeviceGuard guard(device);
    if (guard.err != cudaSuccess) return fail(VQA_E_CUDA, "cudaSetDevice: %s", cudaGetErrorString(guard.err));
    cudaError_t e = vqa::launch_normalize(in_dev, in_stride, n_rows, dim, out_dev, out_stride, cast_out_dev,
                                          cast_kind, cast_stride, reinterpret_cast<cudaStream_t>(stream));
    if (e != cudaSuccess) return fail(VQA_E_CUDA, "normalize launch failed: %s", cudaGetErrorString(e));
    return VQA_OK;
}

int vqa_agree(const int64_t *ids_a_dev, const float *scores_a_dev, const int64_t *ids_b_dev,
              const float *scores_b_dev, int64_t n, double threshold, uint8_t *accept_dev, float *combined_dev,
              int32_t device, void *stream) {
    if (!ids_a_dev || !scores_a_dev || !ids_b_dev || !scores_b_dev)
        return fail(VQA_E_INVALID, "null device pointer argument");
    if (n < 0) return fail(VQA_E_INVALID, "n must be >= 0");
    if (vqa_device_count() == 0) return fail(VQA_E_CUDA, "no CUDA device available (no CPU fallback)");
    if (n == 0) return VQA_OK;
    DeviceGuard guard(device);
    if (guard.err != cudaSuccess) return fail(VQA_E_CUDA, "cudaSetDevice: %s", cudaGetErrorString(guard.err));
    cudaError_t e = vqa::launch_agree((const long long *)ids_a_dev, scores_a_dev, (const long long *)ids_b_dev,
                                      scores_b_dev, n, threshold, accept_dev, combined_dev,
                                      reinterpret_cast<cudaStream_t>(stream));
    if (e != cudaSuccess) return fail(VQA_E_CUDA, "agree launch failed: %s", cudaGetErrorString(e));
    return VQA_OK;
}

// ---- sparse (BM25) leg + hybrid fusion ----------------------------------------------------------

int vqa_sparse_limits(int32_t *max_query_terms, int32_t *max_candidates) {
    if (max_query_terms) *max_query_terms = vqa::sparse_max_terms();
    if (max_candidates) *max_candidates = vqa::sparse_max_cand();
    return VQA_OK;
}

int vqa_sparse_create(vqa_sparse_t **out, int64_t n_docs, int64_t n_terms, int64_t n_postings, int32_t device) {
    if (!out) return fail(VQA_E_INVALID, "out is null");
    *out = nullptr;
    if (n_docs < 0 || n_docs > 0x7fffffffLL)
        return fail(VQA_E_INVALID, "n_docs must be in [0, 2^31) (got %lld)", (long long)n_docs);
    if (n_terms < 0 || n_terms > 0x7fffffffLL)
        return fail(VQA_E_INVALID, "n_terms must be in [0, 2^31) (got %lld)", (long long)n_terms);
    if (n_postings < 0) return fail(VQA_E_INVALID, "n_postings must be >= 0");
    int ndev = vqa_device_count();
    if (ndev == 0) return fail(VQA_E_CUDA, "no CUDA device available (this engine has no CPU fallback)");
    if (device < 0 || device >= ndev) return fail(VQA_E_INVALID, "device %d out of range [0,%d)", device, ndev);
    cudaDeviceProp prop;
    cudaError_t e = cudaGetDeviceProperties(&prop, device);
    if (e != cudaSuccess) return fail(VQA_E_CUDA, "cudaGetDeviceProperties: %s", cudaGetErrorString(e));
    if (prop.major < 10)
        return fail(VQA_E_UNSUPPORTED, "device %d is sm_%d%d; this library is built for sm_100a only", device,
                    prop.major, prop.minor);
    vqa_sparse *h = new (std::nothrow) vqa_sparse();
    if (!h) return fail(VQA_E_NOMEM, "host allocation failed");
    h->n_docs = n_docs;
    h->n_terms = n_terms;
    h->n_postings = n_postings;
    h->device = device;
    h->sm_count = prop.multiProcessorCount;
    *out = h;
    return VQA_OK;
}

int vqa_sparse_bind(vqa_sparse_t *h, const int64_t *offsets_dev, const int32_t *docs_dev, const float *weights_dev) {
    if (!h) return fail(VQA_E_INVALID, "null sparse handle");
    if (!offsets_dev) return fail(VQA_E_INVALID, "offsets_dev is null");
    if (h->n_postings > 0 && (!docs_dev || !weights_dev)) return fail(VQA_E_INVALID, "null postings pointer");
    h->offsets = reinterpret_cast<const long long *>(offsets_dev);
    h->docs = docs_dev;
    h->weights = weights_dev;
    return VQA_OK;
}

int vqa_sparse_destroy(vqa_sparse_t *h) {
    delete h;
    return VQA_OK;
}

int vqa_bm25_weights(const int64_t *offsets_dev, int64_t n_terms, const int32_t *docs_dev, const int32_t *freqs_dev,
                     int64_t n_postings, const double *idf_dev, const int32_t *doc_len_dev, double k1, double b,
                     double avgdl, float *weights_out_dev, int32_t device, void *stream) {
    if (n_terms < 0 || n_postings < 0) return fail(VQA_E_INVALID, "n_terms and n_postings must be >= 0");
    if (n_postings > 0 && (!offsets_dev || !docs_dev || !freqs_dev || !idf_dev || !doc_len_dev || !weights_out_dev))
        return fail(VQA_E_INVALID, "null device pointer argument");
    if (n_postings > 0 && n_terms < 1) return fail(VQA_E_INVALID, "postings without terms");
    if (!(avgdl > 0.0) && n_postings > 0) return fail(VQA_E_INVALID, "avgdl must be > 0");
    if (vqa_device_count() == 0) return fail(VQA_E_CUDA, "no CUDA device available (no CPU fallback)");
    if (n_postings == 0) return VQA_OK;
    DeviceGuard guard(device);
    if (guard.err != cudaSuccess) return fail(VQA_E_CUDA, "cudaSetDevice: %s", cudaGetErrorString(guard.err));
    cudaError_t e = vqa::launch_bm25_weights(reinterpret_cast<const long long *>(offsets_dev), n_terms, docs_dev,
                                             freqs_dev, n_postings, idf_dev, doc_len_dev, k1, b, avgdl,
                                             weights_out_dev, reinterpret_cast<cudaStream_t>(stream));
    if (e != cudaSuccess) return fail(VQA_E_CUDA, "bm25 weights launch failed: %s", cudaGetErrorString(e));
    return VQA_OK;
}

static int sparse_check_shape(const vqa_sparse_t *h, int32_t n_queries, int32_t max_terms, int32_t k_cand_max) {
    if (!h) return fail(VQA_E_INVALID, "null sparse handle");
    if (n_queries < 1 || n_queries > 65535) return fail(VQA_E_INVALID, "n_queries must be in [1, 65535]");
    if (max_terms < 1 || max_terms > vqa::sparse_max_terms())
        return fail(VQA_E_INVALID, "max_terms must be in [1, %d] (got %d)", vqa::sparse_max_terms(), max_terms);
    if (k_cand_max < 1 || k_cand_max > vqa::sparse_max_cand())
        return fail(VQA_E_INVALID, "k_cand_max must be in [1, %d] (got %d)", vqa::sparse_max_cand(), k_cand_max);
    return VQA_OK;
}

int vqa_sparse_workspace_bytes(const vqa_sparse_t *h, int32_t n_queries, int32_t k_cand_max, size_t *bytes) {
    if (!bytes) return fail(VQA_E_INVALID, "bytes is null");
    int rc = sparse_check_shape(h, n_queries, 1, k_cand_max);
    if (rc != VQA_OK) return rc;
    int cpq = 1, tpc = 1;
    vqa::sparse_plan(h->n_docs, n_queries, h->sm_count, &cpq, &tpc);
    *bytes = (size_t)n_queries * cpq * k_cand_max * sizeof(unsigned long long);
    return VQA_OK;
}

int vqa_sparse_search(const vqa_sparse_t *h, const int32_t *q_terms_dev, const float *q_freqs_dev,
                      const int32_t *q_meta_dev, int32_t max_terms, int32_t n_queries, int32_t k_cand_max,
                      int32_t limit, int32_t normalize, double avgscore, double *out_scores_dev,
                      int64_t *out_ids_dev, void *workspace_dev, size_t workspace_bytes, void *stream) {
    int rc = sparse_check_shape(h, n_queries, max_terms, k_cand_max);
    if (rc != VQA_OK) return rc;
    if (!h->offsets) return fail(VQA_E_INVALID, "sparse index has no postings bound (call vqa_sparse_bind)");
    if (!q_terms_dev || !q_freqs_dev || !q_meta_dev || !out_scores_dev || !out_ids_dev || !workspace_dev)
        return fail(VQA_E_INVALID, "null device pointer argument");
    if (limit < 1 || limit > k_cand_max) return fail(VQA_E_INVALID, "limit must be in [1, k_cand_max]");
    if (normalize && !(avgscore > 0.0)) return fail(VQA_E_INVALID, "normalize needs avgscore > 0");
    size_t need = 0;
    rc = vqa_sparse_workspace_bytes(h, n_queries, k_cand_max, &need);
    if (rc != VQA_OK) return rc;
    if (workspace_bytes < need)
        return fail(VQA_E_NOMEM, "workspace too small: %zu bytes given, %zu needed", workspace_bytes, need);
    if (reinterpret_cast<uintptr_t>(workspace_dev) % 8 != 0) return fail(VQA_E_INVALID, "workspace must be 8-byte aligned");
    DeviceGuard guard(h->device);
    if (guard.err != cudaSuccess) return fail(VQA_E_CUDA, "cudaSetDevice: %s", cudaGetErrorString(guard.err));
    vqa::SparseLaunch a;
    a.offsets = h->offsets;
    a.docs = h->docs;
    a.weights = h->weights;
    a.n_docs = h->n_docs;
    a.n_terms = h->n_terms;
    a.q_terms = q_terms_dev;
    a.q_freqs = q_freqs_dev;
    a.q_meta = q_meta_dev;
    a.max_terms = max_terms;
    a.n_queries = n_queries;
    a.kcap = k_cand_max;
    vqa::sparse_plan(h->n_docs, n_queries, h->sm_count, &a.ctas_per_query, &a.tiles_per_cta);
    a.cand = static_cast<unsigned long long *>(workspace_dev);
    a.limit = limit;
    a.normalize = normalize ? 1 : 0;
    a.avgscore = avgscore;
    a.out_s = out_scores_dev;
    a.out_i = reinterpret_cast<long long *>(out_ids_dev);
    cudaError_t e = vqa::launch_sparse_search(a, reinterpret_cast<cudaStream_t>(stream));
    if (e != cudaSuccess) return fail(VQA_E_CUDA, "sparse search launch failed: %s", cudaGetErrorString(e));
    return VQA_OK;
}

static int hybrid_fuse_impl(const float *dense_scores_dev, const int64_t *dense_ids_dev, int32_t k_dense,
                            const double *sparse_scores_dev, const int64_t *sparse_ids_dev, int32_t k_sparse,
                            int32_t n_queries, double w_dense, double w_sparse, int32_t limit, int rrf,
                            double *out_scores_dev, int64_t *out_ids_dev, int32_t device, void *stream) {
    if (!dense_ids_dev || !sparse_ids_dev || !out_scores_dev || !out_ids_dev)
        return fail(VQA_E_INVALID, "null device pointer argument");
    if (!rrf && (!dense_scores_dev || !sparse_scores_dev)) return fail(VQA_E_INVALID, "null score pointer argument");
    if (n_queries < 1) return fail(VQA_E_INVALID, "n_queries must be >= 1");
    if (k_dense < 1 || k_sparse < 1 || k_dense + k_sparse > 2048)
        return fail(VQA_E_INVALID, "k_dense, k_sparse must be >= 1 and sum to <= 2048");
    if (limit < 1 || limit > k_dense + k_sparse) return fail(VQA_E_INVALID, "limit must be in [1, k_dense + k_sparse]");
    if (vqa_device_count() == 0) return fail(VQA_E_CUDA, "no CUDA device available (no CPU fallback)");
    DeviceGuard guard(device);
    if (guard.err != cudaSuccess) return fail(VQA_E_CUDA, "cudaSetDevice: %s", cudaGetErrorString(guard.err));
    cudaError_t e = vqa::launch_hybrid_fuse(dense_scores_dev, reinterpret_cast<const long long *>(dense_ids_dev), k_dense,
                                            sparse_scores_dev, reinterpret_cast<const long long *>(sparse_ids_dev),
                                            k_sparse, n_queries, w_dense, w_sparse, limit, rrf, out_scores_dev,
                                            reinterpret_cast<long long *>(out_ids_dev),
                                            reinterpret_cast<cudaStream_t>(stream));
    if (e != cudaSuccess) return fail(VQA_E_CUDA, "hybrid fuse launch failed: %s", cudaGetErrorString(e));
    return VQA_OK;
}

int vqa_hybrid_fuse(const float *dense_scores_dev, const int64_t *dense_ids_dev, int32_t k_dense,
                    const double *sparse_scores_dev, const int64_t *sparse_ids_dev, int32_t k_sparse,
                    int32_t n_queries, double w_dense, double w_sparse, int32_t limit, double *out_scores_dev,
                    int64_t *out_ids_dev, int32_t device, void *stream) {
    return hybrid_fuse_impl(dense_scores_dev, dense_ids_dev, k_dense, sparse_scores_dev, sparse_ids_dev, k_sparse,
                            n_queries, w_dense, w_sparse, limit, 0, out_scores_dev, out_ids_dev, device, stream);
}

int vqa_hybrid_fuse_rrf(const int64_t *dense_ids_dev, int32_t k_dense, const int64_t *sparse_ids_dev, int32_t k_sparse,
                        int32_t n_queries, double w_dense, double w_sparse, int32_t limit, double *out_scores_dev,
                        int64_t *out_ids_dev, int32_t device, void *stream) {
    return hybrid_fuse_impl(nullptr, dense_ids_dev, k_dense, nullptr, sparse_ids_dev, k_sparse, n_queries, w_dense,
                            w_sparse, limit, 1, out_scores_dev, out_ids_dev, device, stream);
}

int vqa_agree_f64(const int64_t *ids_a_dev, const double *scores_a_dev, const int64_t *ids_b_dev,
                  const double *scores_b_dev, int64_t n, double threshold, uint8_t *accept_dev, double *combined_dev,
                  int32_t device, void *stream) {
    if (!ids_a_dev || !scores_a_dev || !ids_b_dev || !scores_b_dev)
        return fail(VQA_E_INVALID, "null device pointer argument");
    if (n < 0) return fail(VQA_E_INVALID, "n must be >= 0");
    if (vqa_device_count() == 0) return fail(VQA_E_CUDA, "no CUDA device available (no CPU fallback)");
    if (n == 0) return VQA_OK;
    DeviceGuard guard(device);
    if (guard.err != cudaSuccess) return fail(VQA_E_CUDA, "cudaSetDevice: %s", cudaGetErrorString(guard.err));
    cudaError_t e = vqa::launch_agree_f64((const long long *)ids_a_dev, scores_a_dev, (const long long *)ids_b_dev,
                                          scores_b_dev, n, threshold, accept_dev, combined_dev,
                                          reinterpret_cast<cudaStream_t>(stream));
    if (e != cudaSuccess) return fail(VQA_E_CUDA, "agree launch failed: %s", cudaGetErrorString(e));
    return VQA_OK;
}

}  // extern "C"
