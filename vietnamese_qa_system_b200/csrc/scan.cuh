// scan.cuh -- K2a + K3: HBM-streaming similarity scan with fused top-k.
//
// Replaces faiss IndexFlatIP search (sgemm + heap) under txtai ann.search, i.e.
// the work behind Embeddings.search at heavy_ranker.py:98,100, for SMALL query
// batches where the pass over the document matrix is purely HBM-bound.
//
// One persistent CTA per SM.  Each warp streams R whole rows per step with
// 128-bit no-allocate loads (lane l owns 16-byte chunks l, l+32, ... of a row),
// multiplies against BT fp32 queries held in shared memory, reduces the 32
// partials with the XOR butterfly (transposed, so R*BT values cost R*BT-1
// shuffles instead of 5*R*BT), and offers a score to the CTA-shared top-k list
// only when it beats that query's published threshold.  The [B,N] score matrix
// never exists.  Arithmetic order is the canonical order of SURVEY.md App. C, so
// in fp32 the scores are bit-identical to oracle.c's oracle_dot_canonical.
//
// Roofline: HBM.  Algorithmic bytes per launch = n_rows * dim * sizeof(T).
#pragma once

#include "common.cuh"

namespace vqa {

struct ScanParams {
    const unsigned char *rows;
    long long n_rows;
    long long row_stride_bytes;
    int dim;
    const float *q;      // first query of this pass
    long long q_stride;  // elements
    int nq;              // queries in this LAUNCH; blockIdx.y = pass p handles queries [p*BT, min(nq, (p+1)*BT))
    int k;
    float *cand_s;       // [grid][cand_stride] ; this pass writes [.., nq*k) at its offset
    uint32_t *cand_i;
    long long cand_stride;  // elements between consecutive CTAs' blocks
};

constexpr int kScanThreads = 512;
constexpr int kScanWarps = kScanThreads / 32;

// Transposed XOR-butterfly reduction of NV per-lane partials across the warp.
// Order of masks is 16,8,4,2,1 (canonical).  On return v[0] of lane l holds the
// full sum of value index (l >> (5 - log2(NV))).
template <int NV, int CNT, int M>
struct TReduce {
    __device__ __forceinline__ static void run(float (&v)[NV], int lane) {
        if constexpr (CNT > 1) {
            constexpr int half = CNT / 2;
            const bool upper = (lane & M) != 0;
#pragma unroll
            for (int i = 0; i < half; ++i) {
                const float keep = upper ? v[i + half] : v[i];
                const float send = upper ? v[i] : v[i + half];
                v[i] = __fadd_rn(keep, __shfl_xor_sync(kFullMask, send, M));
            }
            if constexpr (M > 1) TReduce<NV, half, M / 2>::run(v, lane);
        } else {
            v[0] = __fadd_rn(v[0], __shfl_xor_sync(kFullMask, v[0], M));
            if constexpr (M > 1) TReduce<NV, 1, M / 2>::run(v, lane);
        }
    }
};

constexpr int ilog2(int x) { return x <= 1 ? 0 : 1 + ilog2(x / 2); }

// shared-memory layout of one query, permuted so that lane l's float4 loads are
// conflict free: float4 index = ((it * H + h) * 32 + lane), H = E/4.
template <typename T>
__host__ __device__ inline int scan_q_floats(int dim) {
    constexpr int E = 16 / (int)sizeof(T);
    int nchunks = dim / E;
    int iters = (nchunks + 31) / 32;
    return iters * 32 * E;
}

template <typename T, int BT>
__host__ __device__ inline size_t scan_smem_bytes(int dim, int k) {
    return (size_t)BT * scan_q_floats<T>(dim) * sizeof(float) + list_smem_bytes<uint32_t>(BT, k);
}

// ITERS > 0: dim*sizeof(T) == ITERS*512 exactly (fully unrolled, all loads of a
// step in flight at once).  ITERS == 0: any dim with dim*sizeof(T) % 16 == 0.
template <typename T, int BT, int R, int ITERS>
__global__ void __launch_bounds__(kScanThreads, 1) scan_topk_kernel(const ScanParams p) {
    constexpr int E = Elem<T>::E;
    constexpr int H = E / 4;
    constexpr int NV = R * BT;
    static_assert(NV <= 32 && (NV & (NV - 1)) == 0, "R*BT must be a power of two <= 32");
    constexpr int SH = 5 - ilog2(NV);  // lanes per value after the reduction

    extern __shared__ __align__(16) unsigned char smem[];
    const int tid = threadIdx.x;
    const int lane = tid & 31;
    const int warp = tid >> 5;
    const int nchunks = p.dim / E;
    const int iters = ITERS > 0 ? ITERS : (nchunks + 31) / 32;
    const int qfl = iters * 32 * E;

    // one launch covers every pass of <= BT queries (grid.y): no launch gaps on small indexes
    const int q0 = blockIdx.y * BT;
    const int nq = p.nq - q0 < BT ? p.nq - q0 : BT;
    const float *qsrc = p.q + (long long)q0 * p.q_stride;
    float4 *qs = reinterpret_cast<float4 *>(smem);
    ListView<uint32_t> L = list_carve<uint32_t>(smem + (size_t)BT * qfl * sizeof(float), BT, p.k);
    list_init(L, BT, tid, kScanThreads);
    grid_launch_dependents();  // lets the (PDL) candidate-reduce grid be scheduled as SMs drain

    // stage the queries (zero-fill beyond nq / beyond dim)
    for (int idx = tid; idx < BT * iters * H * 32; idx += kScanThreads) {
        const int ln = idx & 31;
        const int h = (idx >> 5) % H;
        const int it = (idx / (32 * H)) % iters;
        const int b = idx / (32 * H * iters);
        const int chunk = it * 32 + ln;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (b < nq && chunk < nchunks)
            v = *reinterpret_cast<const float4 *>(qsrc + (long long)b * p.q_stride + chunk * E + h * 4);
        qs[idx] = v;
    }
    __syncthreads();

    const long long gw = (long long)blockIdx.x * kScanWarps + warp;
    const long long nwarps = (long long)gridDim.x * kScanWarps;
    const long long ngroups = (p.n_rows + R - 1) / R;

    for (long long g = gw; g < ngroups; g += nwarps) {
        const long long row0 = g * R;
        float acc[NV];
#pragma unroll
        for (int i = 0; i < NV; ++i) acc[i] = 0.f;

        if constexpr (ITERS > 0) {
            uint4 w[R][ITERS];
#pragma unroll
            for (int r = 0; r < R; ++r) {
                // rows past the end re-read the last row (their scores are discarded)
                const long long row = row0 + r < p.n_rows ? row0 + r : p.n_rows - 1;
                const unsigned char *rp = p.rows + row * p.row_stride_bytes + lane * 16;
#pragma unroll
                for (int it = 0; it < ITERS; ++it) w[r][it] = ldg_stream(rp + it * 512);
            }
#pragma unroll
            for (int it = 0; it < ITERS; ++it) {
                float x[R][E];
#pragma unroll
                for (int r = 0; r < R; ++r) Elem<T>::unpack(w[r][it], x[r]);
#pragma unroll
                for (int h = 0; h < H; ++h) {
#pragma unroll
                    for (int b = 0; b < BT; ++b) {
                        const float4 qv = qs[((b * ITERS + it) * H + h) * 32 + lane];
#pragma unroll
                        for (int r = 0; r < R; ++r) {
                            float a = acc[r * BT + b];
                            a = __fmaf_rn(qv.x, x[r][h * 4 + 0], a);
                            a = __fmaf_rn(qv.y, x[r][h * 4 + 1], a);
                            a = __fmaf_rn(qv.z, x[r][h * 4 + 2], a);
                            a = __fmaf_rn(qv.w, x[r][h * 4 + 3], a);
                            acc[r * BT + b] = a;
                        }
                    }
                }
            }
        } else {
            for (int it = 0; it < iters; ++it) {
                const int chunk = it * 32 + lane;
                if (chunk < nchunks) {
                    uint4 w[R];
#pragma unroll
                    for (int r = 0; r < R; ++r) {
                        const long long row = row0 + r < p.n_rows ? row0 + r : p.n_rows - 1;
                        w[r] = ldg_stream(p.rows + row * p.row_stride_bytes + (long long)chunk * 16);
                    }
                    float x[R][E];
#pragma unroll
                    for (int r = 0; r < R; ++r) Elem<T>::unpack(w[r], x[r]);
#pragma unroll
                    for (int h = 0; h < H; ++h) {
#pragma unroll
                        for (int b = 0; b < BT; ++b) {
                            const float4 qv = qs[((b * iters + it) * H + h) * 32 + lane];
#pragma unroll
                            for (int r = 0; r < R; ++r) {
                                float a = acc[r * BT + b];
                                a = __fmaf_rn(qv.x, x[r][h * 4 + 0], a);
                                a = __fmaf_rn(qv.y, x[r][h * 4 + 1], a);
                                a = __fmaf_rn(qv.z, x[r][h * 4 + 2], a);
                                a = __fmaf_rn(qv.w, x[r][h * 4 + 3], a);
                                acc[r * BT + b] = a;
                            }
                        }
                    }
                }
            }
        }

        TReduce<NV, NV, 16>::run(acc, lane);
        const float sc = acc[0];
        const int vi = lane >> SH;  // value index = r*BT + b
        const int r = vi / BT;
        const int b = vi % BT;
        const long long row = row0 + r;
        const bool rep = (lane & ((1 << SH) - 1)) == 0;
        const float thr = *(volatile float *)(L.tau + b);
        const bool pass = rep && row < p.n_rows && b < nq && sc >= thr;
        unsigned m = __ballot_sync(kFullMask, pass);
        while (m) {
            const int src = __ffs(m) - 1;
            m &= m - 1;
            const float cs = __shfl_sync(kFullMask, sc, src);
            const int cvi = src >> SH;
            list_insert<uint32_t>(L, cvi % BT, cs, (uint32_t)(row0 + cvi / BT));
        }
    }

    __syncthreads();
    // publish this CTA's lists: cand[cta][b][k]
    float *cs = p.cand_s + (long long)blockIdx.x * p.cand_stride + (long long)q0 * p.k;
    uint32_t *ci = p.cand_i + (long long)blockIdx.x * p.cand_stride + (long long)q0 * p.k;
    for (int idx = tid; idx < nq * p.k; idx += kScanThreads) {
        const int b = idx / p.k, e = idx % p.k;
        cs[idx] = L.s[b * L.kcap + e];
        ci[idx] = L.i[b * L.kcap + e];
    }
}

// ---------------------------------------------------------------------------
// Cross-CTA reduce: one CTA per query merges n_lists candidate lists of length
// k_in (u32 local row index, or i64 global id for the cross-rank merge K4) into
// the final top k_out, writes float32 scores and int64 ids (id_base + index).
// ---------------------------------------------------------------------------
constexpr int kReduceWarpsPerCta = 4;

template <typename IdT>
struct ReduceParams {
    const float *cand_s;
    const IdT *cand_i;
    long long list_stride;    // score elements between lists
    long long list_stride_i;  // id elements between lists
    long long query_stride;   // elements between queries inside a list
    int n_lists;
    int list_mod;           // > 1: query q only appears in lists l with l % list_mod == q / queries_per_group
    int queries_per_group;
    int n_queries;
    int k_in;
    int k_out;
    long long id_base;
    float *out_s;     // [nq][k_out]
    long long *out_i;
    unsigned long long *tau_g_reset;  // [n_queries] shared-threshold slots to clear for the next search, or nullptr
    unsigned long long *slot_reset;   // [n_queries][32] warm-up seed slots of the tcgen05 scan to clear (a graph replay
                                      // reuses the epoch, so stale maxima of other queries would pass as bounds), or nullptr
    // Opt-in (VQA_REDUCE_EARLY=1), k_out <= 32 kernel: every candidate list is sorted best-first, and the lists are
    // visited entry-major, so once ceil(n_lists / 32) consecutive chunks -- a window that holds one entry of EVERY
    // list -- offered nothing that beats the running k-th best, no later entry of any list can: stop reading.
    int early_exit;
    // Opt-in (VQA_PDL_CHAIN=1): release programmatic dependents at once.  The only dependent launched that way
    // without its own griddepcontrol.wait is the NEXT scan of the same search, which reads nothing this kernel
    // writes -- it may then overlap this reduce instead of waiting for it.
    int trigger_early;
    // Exact re-scoring of the merged candidates (screen-then-rescore): the scan ranked documents with
    // storage-precision queries; the k_out survivors get their exact fp32 dot product (fp32 query x
    // stored row) here, are re-sorted, and the best k_final are written.  rs_rows == nullptr: off.
    const unsigned char *rs_rows;
    long long rs_stride;   // bytes between rows
    int rs_dim;
    int rs_bf16;           // 1 bf16 rows, 0 fp16 rows
    const float *rs_q;     // fp32 queries of this reduce launch
    long long rs_q_stride;
    int k_final;           // entries written per query (<= k_out); out arrays are [n_queries][k_final]
    // Peer-memory exchange: the candidate lists were pushed into this rank's gather buffer by its peers'
    // push kernels; wait (acquire, system scope) until every one of the wait_n flags has reached wait_epoch.
    const unsigned long long *wait_flags;
    int wait_n;
    unsigned long long wait_epoch;
};

__device__ __forceinline__ unsigned long long ld_acquire_sys_u64(const unsigned long long *p) {
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_sys_u64(unsigned long long *p, unsigned long long v) {
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

// Exchange step over NVLink / NVSwitch peer memory (no NCCL on the data path): CTA r copies this rank's
// packed [scores | ids] block into slot `rank` of peer r's gather buffer with 16-byte peer stores, then
// publishes `epoch` in that peer's flag slot with release semantics at system scope.
struct PushParams {
    const uint4 *local;         // this rank's packed block
    unsigned long long n16;     // its size in 16-byte units
    uint4 *peer_slot[16];       // peer r's gather buffer, already offset to (parity, slot = my rank)
    unsigned long long *peer_flag[16];  // peer r's flag for (parity, my rank)
    unsigned long long epoch;
};

static __global__ void __launch_bounds__(256) exchange_push_kernel(const PushParams p) {
    const int r = blockIdx.x;
    uint4 *dst = p.peer_slot[r];
    for (unsigned long long i = threadIdx.x; i < p.n16; i += blockDim.x) dst[i] = p.local[i];
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence_system();
        st_release_sys_u64(p.peer_flag[r], p.epoch);
    }
}

__device__ __forceinline__ float shfl_xor_any(float v, int m) { return __shfl_xor_sync(kFullMask, v, m); }
__device__ __forceinline__ uint32_t shfl_xor_any(uint32_t v, int m) { return __shfl_xor_sync(kFullMask, v, m); }
__device__ __forceinline__ long long shfl_xor_any(long long v, int m) { return __shfl_xor_sync(kFullMask, v, m); }

template <typename IdT>
__device__ __forceinline__ void bitonic_step_t(float &s, IdT &i, int stride, bool keep_before) {
    const float ps = shfl_xor_any(s, stride);
    const IdT pi = shfl_xor_any(i, stride);
    const bool mine_before = ranks_before<IdT>(s, i, ps, pi);
    if (mine_before != keep_before) {
        s = ps;
        i = pi;
    }
}

// k_out <= 32: ONE WARP per query, the running top-32 lives in registers (lane e = rank e).
// Candidates are visited entry-major (every list's best first), 32 at a time; a chunk with no
// candidate beating the current k-th is skipped by one ballot, otherwise it is bitonic-sorted and
// bitonic-merged into the list.  No shared memory, no locks.
template <typename IdT>
__global__ void __launch_bounds__(kReduceWarpsPerCta * 32) reduce_topk_warp_kernel(const ReduceParams<IdT> p) {
    const int lane = threadIdx.x & 31;
    const int q = blockIdx.x * kReduceWarpsPerCta + (threadIdx.x >> 5);
    if (p.trigger_early) grid_launch_dependents();
    grid_dependency_wait();
    if (q >= p.n_queries) return;
    if (p.wait_flags != nullptr) {
        if (lane < p.wait_n) {
            unsigned long long spins = 0;
            while (ld_acquire_sys_u64(p.wait_flags + lane) < p.wait_epoch) {
                if (++spins > (1ull << 31)) __trap();  // a peer died: fail the launch instead of hanging
            }
        }
        __syncwarp();
    }
    // NL partial lists, each fed by every NL-th chunk: the sort + merge of a chunk is a chain of ~20 dependent
    // shuffle steps (~1.2 us for one warp alone), and chunks of one list depend on each other -- four independent
    // lists advance four chains in lock step instead (straight-line code, same instructions).  A candidate is offered
    // only if it beats the best k-th place any partial list has reached (no candidate below it can be in the union's
    // top k); the lists are merged at the end.
    constexpr int NL = 4;
    float ls[NL];
    IdT li[NL];
#pragma unroll
    for (int l = 0; l < NL; ++l) {
        ls[l] = neg_inf();
        li[l] = invalid_id<IdT>();
    }
    float tau = neg_inf();
    const int lmod = p.list_mod > 1 ? p.list_mod : 1;
    const int lgrp = p.list_mod > 1 ? q / p.queries_per_group : 0;
    const int n_eff = p.n_lists / lmod;  // lists that hold this query
    const long long total = (long long)n_eff * p.k_in;
    const float *qs = p.cand_s + (long long)q * p.query_stride;
    const IdT *qi = p.cand_i + (long long)q * p.query_stride;
    constexpr int PF = 8;  // chunks whose (independent) loads are issued together: hides the L2 latency
    const int quiet_need = (n_eff + 31) / 32;
    int quiet = 0;       // consecutive chunks without a passing candidate (warp-uniform)
    bool stop = false;
    for (long long base0 = 0; base0 < total && !stop; base0 += 32 * PF) {
      float sv[PF];
      IdT iv[PF];
#pragma unroll
      for (int u = 0; u < PF; ++u) {
          const long long idx = base0 + u * 32 + lane;
          sv[u] = neg_inf();
          iv[u] = invalid_id<IdT>();
          if (idx < total) {
              const int e = (int)(idx / n_eff);
              const int l = lgrp + (int)(idx - (long long)e * n_eff) * lmod;
              sv[u] = qs[(long long)l * p.list_stride + e];
              iv[u] = qi[(long long)l * p.list_stride_i + e];
          }
      }
#pragma unroll
      for (int u0 = 0; u0 < PF; u0 += NL) {
        if (base0 + u0 * 32 >= total) break;  // warp-uniform
        float cs[NL];
        IdT ci[NL];
        unsigned any = 0;
#pragma unroll
        for (int l = 0; l < NL; ++l) {
            const float s = sv[u0 + l];
            const IdT id = iv[u0 + l];
            bool valid = id != invalid_id<IdT>();
            if constexpr (sizeof(IdT) == 8) valid = valid && id >= 0;
            const bool pass = valid && s >= tau;
            const unsigned m = __ballot_sync(kFullMask, pass);
            // early exit counts chunks in order: a chunk without a pass extends the quiet run, one with a pass ends it
            if (base0 + (u0 + l) * 32 < total) {
                if (m == 0) ++quiet;
                else quiet = 0;
            }
            any |= m;
            cs[l] = pass ? s : neg_inf();
            ci[l] = pass ? id : invalid_id<IdT>();
        }
        if (any == 0) {
            if (p.early_exit && quiet >= quiet_need) {
                stop = true;
                break;
            }
            continue;
        }
#pragma unroll
        for (int size = 2; size <= 32; size <<= 1) {
#pragma unroll
            for (int stride = size >> 1; stride > 0; stride >>= 1) {
                const bool keep = ((lane & stride) == 0) == ((lane & size) == 0 || size == 32);
#pragma unroll
                for (int l = 0; l < NL; ++l) bitonic_step_t<IdT>(cs[l], ci[l], stride, keep);
            }
        }
#pragma unroll
        for (int l = 0; l < NL; ++l) {
            const float rs = __shfl_sync(kFullMask, ls[l], 31 - lane);
            const IdT ri = shfl_any(li[l], 31 - lane);
            if (ranks_before<IdT>(rs, ri, cs[l], ci[l])) {
                cs[l] = rs;
                ci[l] = ri;
            }
        }
#pragma unroll
        for (int stride = 16; stride > 0; stride >>= 1) {
#pragma unroll
            for (int l = 0; l < NL; ++l) bitonic_step_t<IdT>(cs[l], ci[l], stride, (lane & stride) == 0);
        }
#pragma unroll
        for (int l = 0; l < NL; ++l) {
            ls[l] = cs[l];
            li[l] = ci[l];
            const IdT last = shfl_any(li[l], p.k_out - 1);
            const float t = __shfl_sync(kFullMask, ls[l], p.k_out - 1);
            if (last != invalid_id<IdT>()) tau = fmaxf(tau, t);
        }
        if (p.early_exit && quiet >= quiet_need) {
            stop = true;
            break;
        }
      }
    }
    // union of the partial lists: three merges of sorted lists (five compare-exchange steps each)
#pragma unroll
    for (int l = 1; l < NL; ++l) {
        float cs = ls[l];
        IdT ci = li[l];
        const float rs = __shfl_sync(kFullMask, ls[0], 31 - lane);
        const IdT ri = shfl_any(li[0], 31 - lane);
        if (ranks_before<IdT>(rs, ri, cs, ci)) {
            cs = rs;
            ci = ri;
        }
#pragma unroll
        for (int stride = 16; stride > 0; stride >>= 1) bitonic_step_t<IdT>(cs, ci, stride, (lane & stride) == 0);
        ls[0] = cs;
        li[0] = ci;
    }
    if constexpr (sizeof(IdT) == 4) {
        if (p.rs_rows != nullptr) {
            // lane e re-scores candidate e exactly: 4 independent fp32 FMA chains over the row
            float exact = neg_inf();
            if (lane < p.k_out && li[0] != invalid_id<IdT>()) {
                const unsigned char *row = p.rs_rows + (long long)li[0] * p.rs_stride;
                const float *qv = p.rs_q + (long long)q * p.rs_q_stride;
                float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
                const int nch = p.rs_dim / 8;
                constexpr int RB = 12;  // 16-byte row chunks fetched together (the row is a cold DRAM read)
                for (int c0 = 0; c0 < nch; c0 += RB) {
                    uint4 w[RB];
#pragma unroll
                    for (int u = 0; u < RB; ++u)
                        w[u] = c0 + u < nch ? ldg_stream(row + (c0 + u) * 16) : make_uint4(0, 0, 0, 0);
#pragma unroll
                    for (int u = 0; u < RB; ++u) {
                        if (c0 + u < nch) {
                            const float4 q0 = *reinterpret_cast<const float4 *>(qv + (c0 + u) * 8);
                            const float4 q1 = *reinterpret_cast<const float4 *>(qv + (c0 + u) * 8 + 4);
                            float x[8];
                            if (p.rs_bf16) Elem<__nv_bfloat16>::unpack(w[u], x);
                            else Elem<__half>::unpack(w[u], x);
                            a0 = fmaf(q0.x, x[0], a0);
                            a1 = fmaf(q0.y, x[1], a1);
                            a2 = fmaf(q0.z, x[2], a2);
                            a3 = fmaf(q0.w, x[3], a3);
                            a0 = fmaf(q1.x, x[4], a0);
                            a1 = fmaf(q1.y, x[5], a1);
                            a2 = fmaf(q1.z, x[6], a2);
                            a3 = fmaf(q1.w, x[7], a3);
                        }
                    }
                }
                exact = (a0 + a1) + (a2 + a3);
            } else {
                li[0] = invalid_id<IdT>();
            }
            ls[0] = exact;
#pragma unroll
            for (int size = 2; size <= 32; size <<= 1) {
#pragma unroll
                for (int stride = size >> 1; stride > 0; stride >>= 1) {
                    const bool desc = (lane & size) == 0 || size == 32;
                    bitonic_step_t<IdT>(ls[0], li[0], stride, ((lane & stride) == 0) == desc);
                }
            }
        }
    }
    const int kf = p.k_final > 0 ? p.k_final : p.k_out;
    if (lane < kf) {
        const bool ok = li[0] != invalid_id<IdT>();
        p.out_s[(long long)q * kf + lane] = ok ? ls[0] : neg_inf();
        p.out_i[(long long)q * kf + lane] = ok ? (long long)li[0] + p.id_base : -1LL;
    }
    if (p.tau_g_reset != nullptr && lane == 0) p.tau_g_reset[q] = 0ull;  // a graph replay reuses the epoch
    if (p.slot_reset != nullptr) p.slot_reset[(long long)q * 32 + lane] = 0ull;
}

// k_out in (32, 128]: one CTA of 8 warps per query.  Phase 1: warp w folds the candidate lists
// l = w, w+8, ... (all of a list's entries loaded up front) into ITS OWN sorted shared-memory list
// (uncontended lock).  Phase 2: warp 0 folds the other seven lists into its own and writes the result.
// The input lists need not be sorted.
constexpr int kReduceBigWarps = 8;

template <typename IdT>
__global__ void __launch_bounds__(kReduceBigWarps * 32) reduce_topk_kernel(const ReduceParams<IdT> p) {
    extern __shared__ __align__(16) unsigned char smem[];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int q = blockIdx.x;
    ListView<IdT> L = list_carve<IdT>(smem, kReduceBigWarps, p.k_out);
    list_init(L, kReduceBigWarps, tid, kReduceBigWarps * 32);
    if (p.trigger_early) grid_launch_dependents();
    __syncthreads();
    grid_dependency_wait();
    if (p.wait_flags != nullptr) {   // peer-memory exchange: every rank's block must have landed (acquire, system scope)
        if (tid < p.wait_n) {
            unsigned long long spins = 0;
            while (ld_acquire_sys_u64(p.wait_flags + tid) < p.wait_epoch) {
                if (++spins > (1ull << 31)) __trap();  // a peer died: fail the launch instead of hanging
            }
        }
        __syncthreads();
    }
    const int lmod = p.list_mod > 1 ? p.list_mod : 1;
    const int lgrp = p.list_mod > 1 ? q / p.queries_per_group : 0;
    const int n_eff = p.n_lists / lmod;
    constexpr int CH = kMaxK / 32;  // chunks of 32 entries per list (k_in <= 128)
    for (int j = warp; j < n_eff; j += kReduceBigWarps) {
        const int l = lgrp + j * lmod;
        const float *ls = p.cand_s + (long long)l * p.list_stride + (long long)q * p.query_stride;
        const IdT *li = p.cand_i + (long long)l * p.list_stride_i + (long long)q * p.query_stride;
        float sv[CH];
        IdT iv[CH];
#pragma unroll
        for (int c = 0; c < CH; ++c) {
            const int e = c * 32 + lane;
            sv[c] = neg_inf();
            iv[c] = invalid_id<IdT>();
            if (e < p.k_in) {
                sv[c] = ls[e];
                iv[c] = li[e];
            }
        }
#pragma unroll
        for (int c = 0; c < CH; ++c) {
            if (c * 32 >= p.k_in) break;  // warp-uniform
            bool valid = iv[c] != invalid_id<IdT>();
            if constexpr (sizeof(IdT) == 8) valid = valid && iv[c] >= 0;
            const float thr = *(volatile float *)(L.tau + warp);
            unsigned m = __ballot_sync(kFullMask, valid && sv[c] >= thr);
            while (m) {
                const int src = __ffs(m) - 1;
                m &= m - 1;
                list_insert<IdT>(L, warp, __shfl_sync(kFullMask, sv[c], src), shfl_any(iv[c], src));
            }
        }
    }
    __syncthreads();
    if (warp == 0) {
        for (int w = 1; w < kReduceBigWarps; ++w) {
            for (int e0 = 0; e0 < L.kcap; e0 += 32) {  // list w is sorted: stop at the first chunk with no taker
                const float s = L.s[w * L.kcap + e0 + lane];
                const IdT id = L.i[w * L.kcap + e0 + lane];
                const float thr = *(volatile float *)L.tau;
                unsigned m = __ballot_sync(kFullMask, id != invalid_id<IdT>() && s >= thr);
                if (m == 0) break;
                while (m) {
                    const int src = __ffs(m) - 1;
                    m &= m - 1;
                    list_insert<IdT>(L, 0, __shfl_sync(kFullMask, s, src), shfl_any(id, src));
                }
            }
        }
        __syncwarp();
        for (int e = lane; e < p.k_out; e += 32) {
            const IdT id = L.i[e];
            const bool ok = id != invalid_id<IdT>();
            p.out_s[(long long)q * p.k_out + e] = ok ? L.s[e] : neg_inf();
            p.out_i[(long long)q * p.k_out + e] = ok ? (long long)id + p.id_base : -1LL;
        }
        if (p.tau_g_reset != nullptr && lane == 0) p.tau_g_reset[q] = 0ull;
    }
}

// k_out in (32, 128], u32 row ids: RADIX SELECT, one CTA of 256 threads per query (opt-in: VQA_REDUCE_SELECT=1).
// reduce_topk_kernel above inserts the ~n_lists * k_in candidates of a query one by one into lock-guarded
// sorted lists (0.5 ms at top-100 over 148 lists).  Here every candidate becomes one 64-bit key
//     ordered(score) << 32 | (0x7fffffff - row) << 1 | (score was -0.0)
// whose unsigned order IS the ranking (score descending, ties -> lower row; rows are < 2^31, distinct across
// lists, so the last bit never decides).  The keys are staged in shared memory, an MSB-first 8-bit radix
// select finds the k-th largest (warp-aggregated histogram updates; stops as soon as the chosen bin holds
// exactly the keys still wanted), the survivors are compacted, optionally re-scored exactly (screen mode),
// bitonic-sorted and written.  Same total order as ranks_before, so the result is identical to the other
// reduce kernels' -- including the sign of a zero score.
constexpr int kSelThreads = 256;  // = histogram bins: thread t clears bin t
constexpr int kSelWarps = kSelThreads / 32;
static_assert(kSelThreads == 256 && kSelThreads >= kMaxK, "one thread per radix bin and per output slot");

__device__ __forceinline__ uint32_t ord_f32(float f) {
    const uint32_t u = __float_as_uint(f);
    return (u >> 31) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float unord_f32(uint32_t o) {
    return __uint_as_float((o >> 31) ? (o & 0x7fffffffu) : ~o);
}
__device__ __forceinline__ unsigned long long select_key(float s, uint32_t row) {
    const bool negzero = __float_as_uint(s) == 0x80000000u;
    return ((unsigned long long)ord_f32(negzero ? 0.f : s) << 32) | ((unsigned long long)(0x7fffffffu - row) << 1) |
           (negzero ? 1ull : 0ull);
}

// [keys: n_cand x 8 B][selected: kMaxK x 8 B][histogram: 256 x 4 B][control words: 16 x 4 B]
inline size_t select_smem_bytes(long long n_cand) { return (size_t)n_cand * 8 + (size_t)kMaxK * 8 + 256 * 4 + 64; }

static __global__ void __launch_bounds__(kSelThreads) reduce_select_kernel(const ReduceParams<uint32_t> p) {
    extern __shared__ __align__(16) unsigned char smem[];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int q = blockIdx.x;
    const int lmod = p.list_mod > 1 ? p.list_mod : 1;
    const int lgrp = p.list_mod > 1 ? q / p.queries_per_group : 0;
    const int n_eff = p.n_lists / lmod;
    const int n_cand = n_eff * p.k_in;
    unsigned long long *keys = reinterpret_cast<unsigned long long *>(smem);
    unsigned long long *sel = keys + n_cand;
    uint32_t *hist = reinterpret_cast<uint32_t *>(sel + kMaxK);
    uint32_t *ctl = hist + 256;  // 0: valid candidates, 1: digit, 2: still wanted, 3: done, 4: selected
    if (tid < 16) ctl[tid] = 0;
    if (p.trigger_early) grid_launch_dependents();
    __syncthreads();
    grid_dependency_wait();

    // 1. candidates -> keys (0 = empty slot)
    int nv = 0;
    for (int i = tid; i < n_cand; i += kSelThreads) {
        const int j = i / p.k_in, e = i - j * p.k_in;
        const int l = lgrp + j * lmod;
        const float s = p.cand_s[(long long)l * p.list_stride + (long long)q * p.query_stride + e];
        const uint32_t id = p.cand_i[(long long)l * p.list_stride_i + (long long)q * p.query_stride + e];
        unsigned long long key = 0;
        if (id != invalid_id<uint32_t>()) {
            key = select_key(s, id);
            ++nv;
        }
        keys[i] = key;
    }
#pragma unroll
    for (int m = 16; m >= 1; m >>= 1) nv += __shfl_xor_sync(kFullMask, nv, m);
    if (lane == 0 && nv > 0) atomicAdd(&ctl[0], (uint32_t)nv);
    __syncthreads();
    const int n_valid = (int)ctl[0];
    const int k_eff = n_valid < p.k_out ? n_valid : p.k_out;
    const int kf = p.k_final > 0 ? p.k_final : p.k_out;

    // 2. radix select of the k_eff-th largest key, most significant byte first
    unsigned long long prefix = 0, mask = 0;
    uint32_t wanted = (uint32_t)k_eff;
    bool done = k_eff == 0;
    for (int shift = 56; shift >= 0 && !done; shift -= 8) {
        hist[tid] = 0;
        __syncthreads();
        for (int i0 = 0; i0 < n_cand; i0 += kSelThreads) {  // uniform trip count: the warp votes below
            const int i = i0 + tid;
            uint32_t d = 256;  // not a member
            if (i < n_cand) {
                const unsigned long long key = keys[i];
                if (key != 0 && (key & mask) == prefix) d = (uint32_t)(key >> shift) & 255u;
            }
            const unsigned peers = __match_any_sync(kFullMask, d);
            if (d < 256 && lane == __ffs((int)peers) - 1) atomicAdd(&hist[d], (uint32_t)__popc(peers));
        }
        __syncthreads();
        if (warp == 0) {
            // lane L owns bins 255-8L ... 248-8L (descending); `above` = members in higher bins than its own
            uint32_t c[8], tot = 0;
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                c[j] = hist[255 - (lane * 8 + j)];
                tot += c[j];
            }
            uint32_t incl = tot;
#pragma unroll
            for (int off = 1; off < 32; off <<= 1) {
                const uint32_t t = __shfl_up_sync(kFullMask, incl, off);
                if (lane >= off) incl += t;
            }
            uint32_t above = incl - tot;
            if (above < wanted && wanted <= above + tot) {  // exactly one lane
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    if (wanted > above && wanted <= above + c[j]) {
                        ctl[1] = (uint32_t)(255 - (lane * 8 + j));
                        ctl[2] = wanted - above;
                        ctl[3] = c[j] == wanted - above ? 1u : 0u;  // the whole bin is wanted: stop here
                    }
                    above += c[j];
                }
            }
        }
        __syncthreads();
        prefix |= (unsigned long long)ctl[1] << shift;
        mask |= 255ull << shift;
        wanted = ctl[2];
        done = ctl[3] != 0;
    }

    // 3. survivors: every key whose selected leading bytes are >= the prefix (exactly k_eff of them)
    for (int i0 = 0; i0 < n_cand; i0 += kSelThreads) {
        const int i = i0 + tid;
        unsigned long long key = 0;
        if (i < n_cand && k_eff > 0) key = keys[i];
        const bool in = key != 0 && (key & mask) >= prefix;
        const unsigned m = __ballot_sync(kFullMask, in);
        uint32_t base = 0;
        if (lane == 0 && m != 0) base = atomicAdd(&ctl[4], (uint32_t)__popc(m));
        base = __shfl_sync(kFullMask, base, 0);
        const uint32_t pos = base + (uint32_t)__popc(m & ((1u << lane) - 1u));
        if (in && pos < (uint32_t)kMaxK) sel[pos] = key;
    }
    __syncthreads();
    const int n_sel = ctl[4] < (uint32_t)kMaxK ? (int)ctl[4] : kMaxK;
    if (tid < kMaxK && tid >= n_sel) sel[tid] = 0;
    __syncthreads();

    // 4. screen mode: the survivors' exact fp32 scores (fp32 query x stored row): one warp per candidate, lanes
    //    stride over the row's 16-byte chunks (coalesced), four candidates per round so that their cold row
    //    reads are in flight together
    if (p.rs_rows != nullptr) {
        const float *qv = p.rs_q + (long long)q * p.rs_q_stride;
        const int nch = p.rs_dim / 8;
        constexpr int RC = 4;
        for (int e0 = warp; e0 < n_sel; e0 += kSelWarps * RC) {
            const unsigned char *rp[RC];
            uint32_t rowid[RC];
            bool on[RC];
            float a0[RC], a1[RC];
#pragma unroll
            for (int u = 0; u < RC; ++u) {
                const int e = e0 + u * kSelWarps;
                on[u] = e < n_sel;
                const unsigned long long key = on[u] ? sel[e] : 0ull;
                rowid[u] = 0x7fffffffu - (uint32_t)((key >> 1) & 0x7fffffffull);
                rp[u] = p.rs_rows + (long long)(on[u] ? rowid[u] : 0u) * p.rs_stride;
                a0[u] = a1[u] = 0.f;
            }
            for (int c = lane; c < nch; c += 32) {
                uint4 w[RC];
#pragma unroll
                for (int u = 0; u < RC; ++u) w[u] = on[u] ? ldg_stream(rp[u] + c * 16) : make_uint4(0, 0, 0, 0);
                const float4 q0 = *reinterpret_cast<const float4 *>(qv + c * 8);
                const float4 q1 = *reinterpret_cast<const float4 *>(qv + c * 8 + 4);
#pragma unroll
                for (int u = 0; u < RC; ++u) {
                    float x[8];
                    if (p.rs_bf16) Elem<__nv_bfloat16>::unpack(w[u], x);
                    else Elem<__half>::unpack(w[u], x);
                    a0[u] = fmaf(q0.x, x[0], a0[u]);
                    a1[u] = fmaf(q0.y, x[1], a1[u]);
                    a0[u] = fmaf(q0.z, x[2], a0[u]);
                    a1[u] = fmaf(q0.w, x[3], a1[u]);
                    a0[u] = fmaf(q1.x, x[4], a0[u]);
                    a1[u] = fmaf(q1.y, x[5], a1[u]);
                    a0[u] = fmaf(q1.z, x[6], a0[u]);
                    a1[u] = fmaf(q1.w, x[7], a1[u]);
                }
            }
#pragma unroll
            for (int u = 0; u < RC; ++u) {
                float exact = a0[u] + a1[u];
#pragma unroll
                for (int m = 16; m >= 1; m >>= 1) exact += __shfl_xor_sync(kFullMask, exact, m);
                if (lane == 0 && on[u]) sel[e0 + u * kSelWarps] = select_key(exact, rowid[u]);
            }
        }
        __syncthreads();
    }

    // 5. bitonic sort of the kMaxK slots, best first (empty slots = 0 sink to the end)
    for (int size = 2; size <= kMaxK; size <<= 1) {
        for (int stride = size >> 1; stride > 0; stride >>= 1) {
            if (tid < kMaxK) {
                const int partner = tid ^ stride;
                if (partner > tid) {
                    const unsigned long long a = sel[tid], b = sel[partner];
                    const bool desc = (tid & size) == 0;
                    if ((a < b) == desc) {
                        sel[tid] = b;
                        sel[partner] = a;
                    }
                }
            }
            __syncthreads();
        }
    }
    if (tid < kf) {
        const unsigned long long key = sel[tid];
        const bool ok = key != 0;
        const float s = (key & 1ull) ? __uint_as_float(0x80000000u) : unord_f32((uint32_t)(key >> 32));
        const uint32_t row = 0x7fffffffu - (uint32_t)((key >> 1) & 0x7fffffffull);
        p.out_s[(long long)q * kf + tid] = ok ? s : neg_inf();
        p.out_i[(long long)q * kf + tid] = ok ? (long long)row + p.id_base : -1LL;
    }
    if (p.tau_g_reset != nullptr && tid == 0) p.tau_g_reset[q] = 0ull;
}


// ---- segment merge: top-k for k beyond the per-list capacity (kMaxK) -----------------------------------------
// A search for 128 < k <= kWideK is composed on the host (ops.FlatShard, vqa_merge_segments in include/vqa.h): the
// shard is cut into row segments, each searched for its own top-k_seg with the kernels above, and this kernel
// sorts the n_seg x k_seg survivors of one query (one CTA each) and writes the best k_out.  The answer is exact
// unless a segment holds more than k_seg of the true top-k_out; that can only be the case when the segment's LAST
// kept candidate is itself inside the answer, which is what `saturated[seg]` reports -- the host halves those
// segments and merges again.  Order: score descending (-0 == +0), ties -> the lower position; segments are in
// ascending row order and every list is sorted that way, so the candidate slot number is the tie-break.
constexpr int kWideK = 1024;        // largest k_out
constexpr int kSegMaxCand = 8192;   // n_seg * k_seg, padded to a power of two: 64 KB of keys
constexpr int kSegThreads = 1024;

struct SegMergeParams {
    const float *seg_s;        // [n_seg][n_queries][k_seg]
    const long long *seg_i;    // same shape; -1 = empty slot
    int n_seg, n_queries, k_seg, k_out;
    float *out_s;              // [n_queries][k_out]
    long long *out_i;
    int *saturated;            // [n_seg], OR over the queries; zeroed by the launcher
};

static __global__ void __launch_bounds__(kSegThreads) merge_segments_kernel(const SegMergeParams p) {
    extern __shared__ __align__(16) unsigned char smem[];
    unsigned long long *keys = reinterpret_cast<unsigned long long *>(smem);
    const int tid = threadIdx.x, q = blockIdx.x;
    const int n_cand = p.n_seg * p.k_seg;
    int n_pad = 2;
    while (n_pad < n_cand) n_pad <<= 1;
    const size_t qbase = (size_t)q * p.k_seg, sstride = (size_t)p.n_queries * p.k_seg;
    for (int c = tid; c < n_pad; c += kSegThreads) {
        unsigned long long key = 0ull;  // empty slots sort last
        if (c < n_cand) {
            const int s = c / p.k_seg, j = c - s * p.k_seg;
            const size_t at = (size_t)s * sstride + qbase + j;
            if (p.seg_i[at] >= 0)
                key = ((unsigned long long)ord_f32(p.seg_s[at] + 0.f) << 32) | (unsigned long long)(0xffffffffu - (uint32_t)c);
        }
        keys[c] = key;
    }
    __syncthreads();
    for (int size = 2; size <= n_pad; size <<= 1) {
        for (int stride = size >> 1; stride > 0; stride >>= 1) {
            for (int t = tid; t < (n_pad >> 1); t += kSegThreads) {
                const int lo = 2 * t - (t & (stride - 1)), hi = lo + stride;
                const unsigned long long a = keys[lo], b = keys[hi];
                const bool desc = (lo & size) == 0;
                if ((a < b) == desc) {
                    keys[lo] = b;
                    keys[hi] = a;
                }
            }
            __syncthreads();
        }
    }
    for (int r = tid; r < p.k_out; r += kSegThreads) {
        const unsigned long long key = r < n_pad ? keys[r] : 0ull;
        float s = neg_inf();
        long long id = -1LL;
        if (key != 0ull) {
            const int c = (int)(0xffffffffu - (uint32_t)key);
            const int sg = c / p.k_seg, j = c - sg * p.k_seg;
            const size_t at = (size_t)sg * sstride + qbase + j;
            s = p.seg_s[at];
            id = p.seg_i[at];
        }
        p.out_s[(size_t)q * p.k_out + r] = s;
        p.out_i[(size_t)q * p.k_out + r] = id;
    }
    const unsigned long long kth = p.k_out <= n_pad ? keys[p.k_out - 1] : 0ull;
    for (int sg = tid; sg < p.n_seg; sg += kSegThreads) {
        const size_t at = (size_t)sg * sstride + qbase + (p.k_seg - 1);
        if (p.seg_i[at] >= 0) {
            const int c = sg * p.k_seg + p.k_seg - 1;
            const unsigned long long key =
                ((unsigned long long)ord_f32(p.seg_s[at] + 0.f) << 32) | (unsigned long long)(0xffffffffu - (uint32_t)c);
            if (key >= kth) atomicOr(&p.saturated[sg], 1);
        }
    }
}


}  // namespace vqa
