// emu_kernels.cpp -- TEST INFRASTRUCTURE: host entry points that run the product's CUDA-core kernels
// (transformed copies under _build/gen/, see build.py) on the fiber emulator.  All pointers are HOST
// pointers.  Loaded by tests/test_emu_kernels.py through ctypes; never part of the product.
#include "cuda_emu.h"

// the CUDA runtime calls the launchers make: no device here, launches go to the emulator
#define cudaGetDevice(p) (*(p) = 0, cudaSuccess)
#define cudaFuncSetAttribute(...) cudaSuccess
#define cudaGetLastError() cudaSuccess
#define cudaMemsetAsync(p, v, n, st) (std::memset((p), (v), (n)), cudaSuccess)

// cudaLaunchKernelEx (PDL launch of the reduce kernels, cluster launch of the tensor kernels): same grid / block,
// the cluster dimension attribute is honoured, the others are ignored
template <typename K, typename... P>
static cudaError_t emu_launch_kernel_ex(const cudaLaunchConfig_t *cfg, K kernel, P... p) {
    int cluster = 1;
    for (unsigned i = 0; i < cfg->numAttrs; ++i)
        if (cfg->attrs[i].id == cudaLaunchAttributeClusterDimension) cluster = (int)cfg->attrs[i].val.clusterDim.x;
    emu::launch(cfg->gridDim, cfg->blockDim.x, cfg->dynamicSmemBytes, [&]() { kernel(p...); }, cluster);
    return cudaSuccess;
}
#define cudaLaunchKernelEx emu_launch_kernel_ex
#define cudaOccupancyMaxActiveClusters(...) cudaErrorNotSupported

#include "sparse_launch.cu"  // includes launch.h, sparse.cuh -> common.cuh (generated copies)
#include "scan_f32.cu"       // launch_scan_f32 / _bf16 / _f16 (scan_launch.cuh -> scan.cuh)
#include "scan_bf16.cu"
#include "scan_f16.cu"
#include "misc_launch.cu"    // launch_scan, launch_reduce_*, launch_pool, launch_normalize, launch_agree
#include "mma_launch.cu"     // launch_mma (mma.cuh over the host models of ptx.cuh: tests/emu/ptx_emu.cuh)
#include "ts_launch.cu"      // launch_ts  (ts.cuh: queries resident in tensor memory)

namespace {
thread_local std::string g_emu_err;
template <typename F>
int guarded(F f) {
    try {
        f();
        return 0;
    } catch (const std::exception &e) {
        g_emu_err = e.what();
        return -1;
    }
}
// The product passes the reduce's knobs from the handle's vqa_tuning_t; this TEST harness takes them from the
// environment of the test (monkeypatch.setenv), so one emulator build covers every variant.
vqa::ReduceOpts emu_opts() {
    auto on = [](const char *name) {
        const char *e = std::getenv(name);
        return e != nullptr && std::atoi(e) != 0;
    };
    vqa::ReduceOpts o;
    o.select = on("VQA_REDUCE_SELECT") ? 1 : 0;
    o.early = on("VQA_REDUCE_EARLY") ? 1 : 0;
    o.trigger_early = on("VQA_PDL_CHAIN") ? 1 : 0;
    return o;
}
}  // namespace

extern "C" {

const char *emu_last_error() { return g_emu_err.c_str(); }

// list-update events of the register-list tensor-core epilogue since the last call (see build.py)
long long emu_events_read_reset() {
    const long long c = emu_event_counter();
    emu_event_counter() = 0;
    return c;
}

// fiber order inside a scheduling pass: 0 thread order, 1 reverse, 2 random permutation per pass (seeded)
void emu_set_schedule(int mode, unsigned long long seed) {
    emu::S().schedule = mode;
    emu::S().rng = seed ? seed : 0x9E3779B97F4A7C15ull;
}

// vqa_sparse_search with the same planning (vqa::sparse_plan) and launcher (vqa::launch_sparse_search)
int emu_sparse_search(const long long *offsets, const int *docs, const float *weights, long long n_docs,
                      long long n_terms, const int *q_terms, const float *q_freqs, const int *q_meta, int max_terms,
                      int n_queries, int k_cand_max, int limit, int normalize, double avgscore, int sm_count,
                      double *out_s, long long *out_i) {
    return guarded([&] {
        vqa::SparseLaunch a;
        a.offsets = offsets;
        a.docs = docs;
        a.weights = weights;
        a.n_docs = n_docs;
        a.n_terms = n_terms;
        a.q_terms = q_terms;
        a.q_freqs = q_freqs;
        a.q_meta = q_meta;
        a.max_terms = max_terms;
        a.n_queries = n_queries;
        a.kcap = k_cand_max;
        vqa::sparse_plan(n_docs, n_queries, sm_count, &a.ctas_per_query, &a.tiles_per_cta);
        std::vector<unsigned long long> ws((size_t)n_queries * a.ctas_per_query * k_cand_max);
        a.cand = ws.data();
        a.limit = limit;
        a.normalize = normalize;
        a.avgscore = avgscore;
        a.out_s = out_s;
        a.out_i = out_i;
        if (vqa::launch_sparse_search(a, nullptr) != cudaSuccess) throw std::runtime_error("launch failed");
    });
}

int emu_bm25_weights(const long long *offsets, long long n_terms, const int *docs, const int *freqs,
                     long long n_postings, const double *idf, const int *doc_len, double k1, double b, double avgdl,
                     float *weights) {
    return guarded([&] {
        if (vqa::launch_bm25_weights(offsets, n_terms, docs, freqs, n_postings, idf, doc_len, k1, b, avgdl, weights,
                                     nullptr) != cudaSuccess)
            throw std::runtime_error("launch failed");
    });
}

int emu_hybrid_fuse(const float *ds, const long long *di, int kd, const double *ss, const long long *si, int ks,
                    int n_queries, double wd, double wsp, int limit, int rrf, double *out_s, long long *out_i) {
    return guarded([&] {
        if (vqa::launch_hybrid_fuse(ds, di, kd, ss, si, ks, n_queries, wd, wsp, limit, rrf, out_s, out_i, nullptr) !=
            cudaSuccess)
            throw std::runtime_error("launch failed");
    });
}

// K1 through the product's dispatch (launch_pool: warp-per-token kernel by row length, generic kernel otherwise)
int emu_pool_normalize(const void *hidden, int h_dtype, const void *mask, int m_dtype, int batch, int seq, int dim,
                       int normalize, float *out) {
    return guarded([&] {
        if (vqa::launch_pool(hidden, h_dtype, mask, m_dtype, batch, seq, dim, normalize, out, nullptr) != cudaSuccess)
            throw std::runtime_error("launch failed");
    });
}
int emu_normalize_rows(const float *in, long long n_rows, int dim, float *out, void *cast_out, int cast_kind) {
    return guarded([&] {
        if (vqa::launch_normalize(in, dim, n_rows, dim, out, dim, cast_out, cast_kind, dim, nullptr) != cudaSuccess)
            throw std::runtime_error("launch failed");
    });
}

int emu_agree(const long long *ids_a, const float *sa, const long long *ids_b, const float *sb, long long n,
              double threshold, unsigned char *accept, float *combined) {
    return guarded([&] {
        if (vqa::launch_agree(ids_a, sa, ids_b, sb, n, threshold, accept, combined, nullptr) != cudaSuccess)
            throw std::runtime_error("launch failed");
    });
}

int emu_merge_topk(const float *cand_s, const long long *cand_i, int n_lists, int n_queries, int k_in, int k_out,
                   float *out_s, long long *out_i) {
    return guarded([&] {
        const long long stride = (long long)n_queries * k_in;
        if (vqa::launch_reduce_i64(cand_s, cand_i, stride, stride, k_in, n_lists, k_in, k_out, 0, out_s, out_i,
                                   n_queries, nullptr) != cudaSuccess)
            throw std::runtime_error("launch failed");
    });
}

// vqa_merge_segments: [n_seg][n_queries][k_seg] sorted lists -> top k_out + the per-segment saturation flags
int emu_merge_segments(const float *seg_s, const long long *seg_i, int n_seg, int n_queries, int k_seg, int k_out,
                       float *out_s, long long *out_i, int *saturated) {
    return guarded([&] {
        if (vqa::launch_merge_segments(seg_s, seg_i, n_seg, n_queries, k_seg, k_out, out_s, out_i, saturated,
                                       nullptr) != cudaSuccess)
            throw std::runtime_error("launch failed");
    });
}

// Peer-memory exchange (vqa_exchange_push / vqa_merge_topk_wait) with host buffers standing in for the peers'
// symmetric memory: push this rank's packed block into every peer slot + publish the epoch flag ...
int emu_exchange_push(const void *local, size_t bytes, void *const *peer_slots, unsigned long long *const *peer_flags,
                      int world, unsigned long long epoch) {
    return guarded([&] {
        if (vqa::launch_exchange_push(local, bytes, peer_slots, peer_flags, world, epoch, nullptr) != cudaSuccess)
            throw std::runtime_error("push launch failed");
    });
}
// ... and the merge whose kernel first acquires all flags
int emu_merge_topk_wait(const float *cand_s, const long long *cand_i, long long stride_s, long long stride_i,
                        int n_lists, int n_queries, int k, float *out_s, long long *out_i,
                        const unsigned long long *flags, unsigned long long epoch) {
    return guarded([&] {
        vqa::WaitFlags wf;
        wf.flags = flags;
        wf.n = n_lists;
        wf.epoch = epoch;
        if (vqa::launch_reduce_i64(cand_s, cand_i, stride_s, stride_i, k, n_lists, k, k, 0, out_s, out_i, n_queries,
                                   nullptr, &wf) != cudaSuccess)
            throw std::runtime_error("merge launch failed");
    });
}

// The candidate reduce with the exact re-scoring stage (what follows the TMEM-resident tensor-core scan in
// screen mode): cand_* [n_lists][n_queries][k_in] with u32 row ids, rows in 16-bit storage.
int emu_reduce_rescore(const float *cand_s, const uint32_t *cand_i, int n_lists, int n_queries, int k_in, int k_out,
                       int k_final, long long id_base, const void *rows, int dim, int bf16, const float *q,
                       float *out_s, long long *out_i) {
    return guarded([&] {
        vqa::Rescore rs;
        rs.rows = rows;
        rs.stride = (long long)dim * 2;
        rs.dim = dim;
        rs.bf16 = bf16;
        rs.q = q;
        rs.q_stride = dim;
        rs.k_final = k_final;
        const long long stride = (long long)n_queries * k_in;
        if (vqa::launch_reduce_u32(cand_s, cand_i, stride, k_in, n_lists, k_in, k_out, id_base, out_s, out_i, n_queries,
                                   nullptr, 1, 1, nullptr, emu_opts(), &rs) != cudaSuccess)
            throw std::runtime_error("reduce launch failed");
    });
}

// The cross-CTA candidate reduce alone (u32 row ids, no re-scoring): cand_* [n_lists][n_queries][k_in]; with
// list_mod > 1 query j only appears in the lists l with l % list_mod == j / queries_per_group (api.cu's
// side-by-side query chunks).  k_out > 32 takes reduce_topk_kernel or, under VQA_REDUCE_SELECT=1, the radix select.
int emu_reduce_u32(const float *cand_s, const uint32_t *cand_i, int n_lists, int n_queries, int k_in, int k_out,
                   long long id_base, int list_mod, int queries_per_group, unsigned long long *tau_g, float *out_s,
                   long long *out_i) {
    return guarded([&] {
        const long long stride = (long long)n_queries * k_in;
        if (vqa::launch_reduce_u32(cand_s, cand_i, stride, k_in, n_lists, k_in, k_out, id_base, out_s, out_i, n_queries,
                                   tau_g, list_mod, queries_per_group, nullptr, emu_opts(), nullptr) != cudaSuccess)
            throw std::runtime_error("reduce launch failed");
    });
}

// vqa_search in FAST_TENSOR mode (hi/lo column pairs, no clusters): mma_topk_kernel + the candidate reduce, wired
// as api.cu does it.  rows: 16-bit storage [n_rows][dim]; ncol in {16, 32, 64, 128}; the batch is cut into chunks
// of ncol / 2 queries handled side by side (n_groups = chunks).
int emu_search_tensor(const void *rows, int bf16, long long n_rows, int dim, const float *q, int n_queries, int k,
                      long long first_id, int sm_count, int ncol, int stages, int kps, int multicast,
                      float *out_s, long long *out_i) {
    return guarded([&] {
        const int pass_nq = ncol / 2;
        int g = (n_queries + pass_nq - 1) / pass_nq;
        if (multicast) {  // cluster sizes are powers of two; a short launch gets empty chunks (api.cu)
            int lg = 0;
            while ((1 << lg) < g) ++lg;
            g = 1 << lg;
        }
        const bool mc = multicast && g > 1;
        const long long tiles = (n_rows + vqa::kTileRows - 1) / vqa::kTileRows;
        long long streams = sm_count / g;
        if (streams > tiles) streams = tiles;
        if (streams < 1) streams = 1;
        CUtensorMap tmap;
        std::memset(&tmap, 0, sizeof(tmap));
        emu::EmuTmap m;
        m.base = static_cast<const unsigned char *>(rows);
        m.dim0 = (unsigned long long)dim;
        m.dim1 = (unsigned long long)n_rows;
        m.stride1_bytes = (unsigned long long)dim * 2;
        m.box0 = 64;
        m.box1 = vqa::kTileRows / (mc ? g : 1);  // multicast: each CTA fetches a slice of the box for everybody
        m.elem_bytes = 2;
        m.magic = emu::kTmapMagic;
        std::memcpy(&tmap, &m, sizeof(m));
        const int grid = (int)streams * g;
        const long long cstride = (long long)n_queries * k;
        std::vector<float> cand_s((size_t)grid * cstride);
        std::vector<uint32_t> cand_i((size_t)grid * cstride);
        std::vector<unsigned long long> tau_g((size_t)n_queries, 0ull);
        vqa::MmaLaunch a;
        a.tmap = &tmap;
        a.bf16 = bf16 != 0;
        a.ncol = ncol;
        a.split = 1;
        a.stages = stages;
        a.kps = kps;
        a.grid = grid;
        a.n_groups = g;
        a.multicast = mc ? 1 : 0;
        a.q = q;
        a.q_stride = dim;
        a.nq = n_queries;
        a.k = k;
        a.n_rows = n_rows;
        a.dim = dim;
        a.cand_s = cand_s.data();
        a.cand_i = cand_i.data();
        a.cand_stride = cstride;
        a.tau_g = tau_g.data();
        a.epoch = 1;
        // dynamic tile schedule (api.cu: launches without clusters) unless the test switches it off; the counter
        // starts as garbage of another epoch, as a recycled workspace would hold
        static unsigned long long tile_ctr = 0xdeadbeef00000007ull;
        const char *dyn = std::getenv("VQA_DYN_TILES");
        a.tile_ctr = (dyn == nullptr || std::atoi(dyn) != 0) ? &tile_ctr : nullptr;
        // warm-up seed slots (api.cu: register-list path): stale values of another epoch, as a recycled workspace holds
        std::vector<unsigned long long> slots((size_t)n_queries * 32, 0x0000000700000000ull | 0xffffffffull);
        const char *seed = std::getenv("VQA_SEED");
        a.slot_g = ((seed == nullptr || std::atoi(seed) != 0) && pass_nq <= 32 && k <= 32) ? slots.data() : nullptr;
        if (vqa::launch_mma(a, nullptr) != cudaSuccess) throw std::runtime_error("tensor scan launch failed");
        if (a.tile_ctr != nullptr && g == 1 && tile_ctr != 0) throw std::runtime_error("tile counter not reset by the last CTA");
        if (vqa::launch_reduce_u32(cand_s.data(), cand_i.data(), cstride, k, grid, k, k, first_id, out_s, out_i,
                                   n_queries, tau_g.data(), g, pass_nq, nullptr, emu_opts(), nullptr, a.slot_g) != cudaSuccess)
            throw std::runtime_error("reduce launch failed");
        if (a.slot_g != nullptr)
            for (unsigned long long v : slots)
                if (v != 0) throw std::runtime_error("seed slots not cleared by the reduce");
    });
}

// vqa_search in FAST_TS mode (queries resident in tensor memory, no clusters): ts_topk_kernel + the reduce, wired
// as api.cu does it.  split = 0: storage-precision screen with k + spare candidates per query, the 32 best
// re-scored exactly by the reduce; split = 1: hi + lo query rows (64 queries per CTA), no re-scoring.
// qs = 1: the QS kernel variant with the last ks 64-column blocks of the query block in shared memory; with
// k + spare > 32 and split = 0 the radix-select reduce (VQA_REDUCE_SELECT=1) re-scores the 128 best.
int emu_search_ts(const void *rows, int bf16, long long n_rows, int dim, const float *q, int n_queries, int k,
                  long long first_id, int sm_count, int split, int spare, int stages, int kps, int multicast,
                  int qs, int ks, float *out_s, long long *out_i) {
    return guarded([&] {
        const int pass_nq = split ? 64 : 128;
        const int kscan = split ? k : k + spare;
        int g = (n_queries + pass_nq - 1) / pass_nq;
        if (multicast) {
            int lg = 0;
            while ((1 << lg) < g) ++lg;
            g = 1 << lg;
        }
        const bool mc = multicast && g > 1;
        const long long tiles = (n_rows + 63) / 64;
        long long streams = sm_count / g;
        if (streams > tiles) streams = tiles;
        if (streams < 1) streams = 1;
        CUtensorMap tmap;
        std::memset(&tmap, 0, sizeof(tmap));
        emu::EmuTmap m;
        m.base = static_cast<const unsigned char *>(rows);
        m.dim0 = (unsigned long long)dim;
        m.dim1 = (unsigned long long)n_rows;
        m.stride1_bytes = (unsigned long long)dim * 2;
        m.box0 = 64;
        m.box1 = 64 / (mc ? g : 1);
        m.elem_bytes = 2;
        m.magic = emu::kTmapMagic;
        std::memcpy(&tmap, &m, sizeof(m));
        const int grid = (int)streams * g;
        const long long cstride = (long long)n_queries * kscan;
        std::vector<float> cand_s((size_t)grid * cstride);
        std::vector<uint32_t> cand_i((size_t)grid * cstride);
        std::vector<unsigned long long> tau_g((size_t)n_queries, 0ull);
        vqa::TsLaunch a;
        a.tmap = &tmap;
        a.bf16 = bf16 != 0;
        a.split = split;
        a.a_fp16 = 0;
        a.stages = stages;
        a.kps = kps;
        a.grid = grid;
        a.n_groups = g;
        a.multicast = mc ? 1 : 0;
        a.q = q;
        a.q_stride = dim;
        a.nq = n_queries;
        a.k = kscan;
        a.n_rows = n_rows;
        a.dim = dim;
        a.cand_s = cand_s.data();
        a.cand_i = cand_i.data();
        a.cand_stride = cstride;
        a.tau_g = tau_g.data();
        a.epoch = 1;
        a.qs = qs;
        a.ks = ks;
        if (vqa::ts_smem_bytes(kscan, stages * kps, split, qs ? ks : 0, n_queries, qs) > 227 * 1024)
            throw std::runtime_error("TS plan does not fit shared memory");
        if (vqa::launch_ts(a, nullptr) != cudaSuccess) throw std::runtime_error("TS scan launch failed");
        vqa::Rescore rs;
        rs.rows = rows;
        rs.stride = (long long)dim * 2;
        rs.dim = dim;
        rs.bf16 = bf16;
        rs.q = q;
        rs.q_stride = dim;
        rs.k_final = k;
        if (vqa::launch_reduce_u32(cand_s.data(), cand_i.data(), cstride, kscan, grid, kscan,
                                   split ? kscan : (kscan > 32 ? vqa::kMaxK : 32), first_id,
                                   out_s, out_i, n_queries, tau_g.data(), g, pass_nq, nullptr, emu_opts(),
                                   split ? nullptr : &rs) != cudaSuccess)
            throw std::runtime_error("reduce launch failed");
    });
}

// vqa_search in VERIFY / FAST_STREAM mode: the CUDA-core scan kernel + the candidate reduce, planned exactly as
// api.cu does it (plan_stream + the grid rule of the stream family).
int emu_search_stream(const void *rows, int dtype, long long n_rows, int dim, const float *q, int n_queries, int k,
                      long long first_id, int sm_count, float *out_s, long long *out_i) {
    return guarded([&] {
        const int es = dtype == VQA_F32 ? 4 : 2;
        const int pass_nq = n_queries >= 5 ? 8 : (n_queries >= 3 ? 4 : (n_queries == 2 ? 2 : 1));
        const int passes_total = (n_queries + pass_nq - 1) / pass_nq;
        long long gx = (n_rows + 511) / 512;
        if (gx > sm_count) gx = sm_count;
        if (passes_total > 1 && gx > sm_count / passes_total) gx = sm_count / passes_total;
        if (gx < 1) gx = 1;
        const long long cand_stride = (long long)n_queries * k;
        std::vector<float> cand_s((size_t)gx * cand_stride);
        std::vector<uint32_t> cand_i((size_t)gx * cand_stride);
        int n_lists = 0;
        if (n_rows > 0) {
            n_lists = (int)gx;
            vqa::ScanLaunch a;
            a.dtype = dtype;
            a.bt = pass_nq;
            a.grid = (int)gx;
            a.rows = rows;
            a.n_rows = n_rows;
            a.row_stride_bytes = (long long)dim * es;
            a.dim = dim;
            a.q = q;
            a.q_stride = dim;
            a.nq = n_queries;
            a.k = k;
            a.cand_s = cand_s.data();
            a.cand_i = cand_i.data();
            a.cand_stride = cand_stride;
            if (vqa::launch_scan(a, nullptr) != cudaSuccess) throw std::runtime_error("scan launch failed");
        }
        if (vqa::launch_reduce_u32(cand_s.data(), cand_i.data(), cand_stride, k, n_lists, k, k, first_id, out_s, out_i,
                                   n_queries, nullptr, 1, 1, nullptr, emu_opts()) != cudaSuccess)
            throw std::runtime_error("reduce launch failed");
    });
}

}  // extern "C"
