// ptx_emu.cuh -- TEST INFRASTRUCTURE: functional host models of the inline-PTX wrappers in
// vietnamese_qa_system_b200/csrc/ptx.cuh (same namespace, same names, same signatures), so that the
// tensor-core kernels can run on the fiber emulator (cuda_emu.h).  tests/emu/build.py installs this file as
// gen/ptx.cuh and checks that every wrapper of the real header has a counterpart here.
//
// What is modelled (enough for the kernels' control flow and arithmetic, not for timing):
//   * shared-memory addresses   -- offsets into the emulator's dynamic shared-memory arena
//   * mbarrier                  -- pending-arrival count + transaction bytes + phase parity in the 8 bytes
//   * TMA 2-D tiled load        -- synchronous copy of a (64 elements x rows) box with the 128-byte swizzle
//                                  (16-byte chunk index XOR row-in-group, on absolute shared addresses),
//                                  zero fill outside the tensor, complete_tx on the mbarrier
//   * tensor memory             -- 128 lanes x 512 columns of 32 bits per CTA
//   * tcgen05.mma kind::f16     -- D[m][n] (+)= sum_k A[m][k] * B[n][k] over one K = 16 step, A and B read
//                                  through K-major SWIZZLE_128B shared-memory descriptors, executed at issue;
//                                  tcgen05.commit therefore arrives at once
//   * tcgen05.ld 32x32b.x16     -- thread t of the warp reads lane (taddr.lane + t), 16 columns
//   * thread-block clusters     -- up to 4 CTAs run together, each with its own shared and tensor memory;
//                                  TMA multicast and multicast commits address the same offset in every CTA
// Not modelled: cache policies, memory proxies, timing.
#pragma once

#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace emu {

// the "tensor map" the emulated TMA reads: stored in the first bytes of a CUtensorMap blob
struct EmuTmap {
    const unsigned char *base;
    unsigned long long dim0, dim1;      // elements per row, rows
    unsigned long long stride1_bytes;   // bytes between rows
    unsigned box0, box1;                // box: elements x rows
    unsigned elem_bytes;
    unsigned magic;
};
constexpr unsigned kTmapMagic = 0x454D5554u;  // "EMUT"

struct MBar {  // the 8 bytes of an mbarrier
    int32_t tx;
    uint16_t pending;
    uint8_t expected;
    uint8_t phase;
};
static_assert(sizeof(MBar) == 8, "mbarrier is 8 bytes");

struct TensorState {  // per CTA
    uint32_t tmem[128][512];
    int nb_arrived[16] = {}, nb_gen[16] = {};
};
inline TensorState &T() {
    static TensorState t[kMaxCluster];
    return t[cur_cta()];
}

inline unsigned char *smem_base() { return cur_smem(); }
inline unsigned char *smem_base_of(int cta) { return S().dyn_smem[cta]; }
// offset of a shared-memory pointer inside its CTA's window (the same offset names the same object in a peer CTA)
inline uint32_t smem_off(const void *p) {
    const unsigned char *c = static_cast<const unsigned char *>(p);
    for (int k = 0; k < kMaxCluster; ++k) {
        const long long d = c - S().dyn_smem[k];
        if (d >= 0 && d < (long long)sizeof(S().dyn_smem[k])) return (uint32_t)d;
    }
    throw std::runtime_error("emu: pointer outside shared memory");
}
// SWIZZLE_128B: bits [4,7) of the address XOR bits [7,10)
inline uint32_t swz128(uint32_t a) { return a ^ (((a >> 7) & 7u) << 4); }

inline void mbar_check(MBar *b) {
    if (b->pending == 0 && b->tx == 0) {
        b->phase ^= 1;
        b->pending = b->expected;
        ++S().progress;
    }
}

inline void named_bar_sync(int id, int count) {
    TensorState &t = T();
    const int gen = t.nb_gen[id];
    if (++t.nb_arrived[id] == count) {
        t.nb_arrived[id] = 0;
        ++t.nb_gen[id];
        ++S().progress;
    } else
        while (t.nb_gen[id] == gen) yield();
}
inline void named_bar_arrive(int id, int count) {
    TensorState &t = T();
    if (++t.nb_arrived[id] == count) {
        t.nb_arrived[id] = 0;
        ++t.nb_gen[id];
        ++S().progress;
    }
}

}  // namespace emu

namespace vqa {
namespace ptx {

inline uint32_t smem_u32(const void *p) { return emu::smem_off(p); }

// elect.sync is a convergence point of the warp: lanes cannot drift more than one loop iteration apart
inline bool elect_one() {
    emu::warp_barrier();
    return (threadIdx.x & 31u) == 0;
}

// ---- mbarrier -------------------------------------------------------------
inline void mbar_init(uint64_t *bar, uint32_t count) {
    emu::MBar *b = reinterpret_cast<emu::MBar *>(bar);
    b->tx = 0;
    b->pending = (uint16_t)count;
    b->expected = (uint8_t)count;
    b->phase = 0;
}
inline void fence_mbar_init() {}
inline void mbar_arrive(uint64_t *bar) {
    emu::MBar *b = reinterpret_cast<emu::MBar *>(bar);
    if (b->pending == 0) throw std::runtime_error("emu: mbarrier arrive overflow");
    --b->pending;
    emu::mbar_check(b);
}
inline void mbar_arrive_expect_tx(uint64_t *bar, uint32_t bytes) {
    emu::MBar *b = reinterpret_cast<emu::MBar *>(bar);
    b->tx += (int32_t)bytes;
    if (b->pending == 0) throw std::runtime_error("emu: mbarrier arrive overflow");
    --b->pending;
    emu::mbar_check(b);
}
inline bool mbar_try_wait(uint64_t *bar, uint32_t parity) {
    emu::MBar *b = reinterpret_cast<emu::MBar *>(bar);
    const bool ok = b->phase != (uint8_t)(parity & 1u);  // the phase with this parity has completed
    if (!ok) emu::yield();
    return ok;
}
#ifndef VQA_SPIN_LIMIT
#define VQA_SPIN_LIMIT (1u << 26)
#endif
inline void mbar_wait(uint64_t *bar, uint32_t parity) {
    while (!mbar_try_wait(bar, parity)) {
    }  // a wait that can never complete is reported by the scheduler as a deadlock
}

// ---- proxies / fences -------------------------------------------------------
inline void fence_proxy_async_smem() {}
inline void tc_fence_before_sync() {}
inline void tc_fence_after_sync() {}

// ---- TMA --------------------------------------------------------------------
constexpr uint64_t kEvictFirst = 0x12F0000000000000ull;
constexpr uint64_t kEvictLast = 0x14F0000000000000ull;

inline uint64_t globaltimer_ns() { return 0; }  // diagnostic stamps: no clock in the emulator
inline uint64_t sm_clock() { return 0; }
inline void prefetch_tmap(const void *) {}
// copy one box into CTA `cta`'s shared memory at offset `dst` and complete its bytes on that CTA's mbarrier
inline void emu_tma_box(int cta, uint32_t dst, const emu::EmuTmap *m, int32_t c0, int32_t c1, uint32_t bar_off) {
    const uint32_t row_bytes = m->box0 * m->elem_bytes;
    if (row_bytes != 128) throw std::runtime_error("emu: only 128-byte box rows (SWIZZLE_128B) are modelled");
    unsigned char *base = emu::smem_base_of(cta);
    for (uint32_t r = 0; r < m->box1; ++r) {
        const long long row = (long long)c1 + r;
        for (uint32_t ch = 0; ch < 8; ++ch) {
            unsigned char *out = base + emu::swz128(dst + r * 128 + ch * 16);
            const long long col = (long long)c0 + (long long)ch * (16 / m->elem_bytes);
            const bool inside = row >= 0 && row < (long long)m->dim1 && col >= 0 &&
                                col + (16 / m->elem_bytes) <= (long long)m->dim0;
            if (inside) std::memcpy(out, m->base + row * m->stride1_bytes + col * m->elem_bytes, 16);
            else std::memset(out, 0, 16);
        }
    }
    emu::MBar *b = reinterpret_cast<emu::MBar *>(base + bar_off);
    b->tx -= (int32_t)(m->box1 * row_bytes);
    emu::mbar_check(b);
}
inline const emu::EmuTmap *emu_tmap(const void *tmap) {
    const emu::EmuTmap *m = static_cast<const emu::EmuTmap *>(tmap);
    if (m->magic != emu::kTmapMagic) throw std::runtime_error("emu: not an emulated tensor map");
    return m;
}
inline void tma_load_2d(void *smem_dst, const void *tmap, int32_t c0, int32_t c1, uint64_t *bar, uint64_t) {
    emu_tma_box(emu::cur_cta(), emu::smem_off(smem_dst), emu_tmap(tmap), c0, c1, emu::smem_off(bar));
}
// multicast: the box lands at the same offset in every CTA of the mask and completes on the mbarrier at the
// same offset in each of them
inline void tma_load_2d_multicast(void *smem_dst, const void *tmap, int32_t c0, int32_t c1, uint64_t *bar,
                                  uint16_t cta_mask, uint64_t) {
    for (int c = 0; c < emu::S().cluster; ++c)
        if (cta_mask & (1u << c)) emu_tma_box(c, emu::smem_off(smem_dst), emu_tmap(tmap), c0, c1, emu::smem_off(bar));
}

// ---- thread-block clusters ------------------------------------------------------------
inline uint32_t cluster_ctarank() { return (uint32_t)emu::cur_cta(); }
inline void cluster_sync_all() { emu::cluster_barrier(); }

// ---- tcgen05: TMEM allocation -----------------------------------------------
inline void tmem_alloc(uint32_t *smem_dst, uint32_t) { *smem_dst = 0; }  // lane 0, column 0
inline void tmem_relinquish() {}
inline void tmem_dealloc(uint32_t, uint32_t) {}

// ---- tcgen05: MMA -----------------------------------------------------------
inline uint64_t umma_desc_k_sw128(uint32_t smem_addr) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr & 0x3FFFFu) >> 4);
    d |= (uint64_t)1 << 16;
    d |= (uint64_t)(1024 >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}
constexpr uint32_t umma_idesc_f16(int m, int n, bool bf16) {
    return (1u << 4) | ((bf16 ? 1u : 0u) << 7) | ((bf16 ? 1u : 0u) << 10) | ((uint32_t)(n >> 3) << 17) |
           ((uint32_t)(m >> 4) << 24);
}
inline float emu_elem16(const unsigned char *p, bool bf16) {
    uint16_t h;
    std::memcpy(&h, p, 2);
    if (bf16) {
        const uint32_t u = (uint32_t)h << 16;
        float f;
        std::memcpy(&f, &u, 4);
        return f;
    }
    __half_raw r;
    r.x = h;
    return __half2float(__half(r));
}
// element (row, kk) of a K-major SWIZZLE_128B operand tile described by `desc` (one K = 16 step)
inline float emu_operand(uint64_t desc, int row, int kk, bool bf16) {
    const uint32_t start = (uint32_t)(desc & 0x3FFFu) << 4;
    const uint32_t sbo = (uint32_t)((desc >> 32) & 0x3FFFu) << 4;
    if (((desc >> 61) & 7u) != 2u) throw std::runtime_error("emu: only SWIZZLE_128B descriptors are modelled");
    const uint32_t logical = start + (uint32_t)(row >> 3) * sbo + (uint32_t)(row & 7) * 128 + (uint32_t)kk * 2;
    return emu_elem16(emu::smem_base() + emu::swz128(logical), bf16);
}
inline void umma_f16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
    const int n = (int)((idesc >> 17) & 0x3Fu) << 3, m = (int)((idesc >> 24) & 0x1Fu) << 4;
    const bool bf16 = ((idesc >> 7) & 7u) == 1u;
    const uint32_t lane0 = tmem_d >> 16, col0 = tmem_d & 0xFFFFu;
    if (lane0 + (uint32_t)m > 128 || col0 + (uint32_t)n > 512) throw std::runtime_error("emu: MMA outside tensor memory");
    float b[256][16];
    for (int j = 0; j < n; ++j)
        for (int kk = 0; kk < 16; ++kk) b[j][kk] = emu_operand(desc_b, j, kk, bf16);
    for (int i = 0; i < m; ++i) {
        float a[16];
        for (int kk = 0; kk < 16; ++kk) a[kk] = emu_operand(desc_a, i, kk, bf16);
        for (int j = 0; j < n; ++j) {
            float acc = 0.0f;
            for (int kk = 0; kk < 16; ++kk) acc = std::fmaf(a[kk], b[j][kk], acc);
            uint32_t &cell = emu::T().tmem[lane0 + i][col0 + j];
            float d;
            std::memcpy(&d, &cell, 4);
            d = accumulate ? d + acc : acc;
            std::memcpy(&cell, &d, 4);
        }
    }
}
inline void umma_commit(uint64_t *bar) { mbar_arrive(bar); }  // MMAs execute at issue: nothing is in flight
// arrive on the mbarrier at this offset in every CTA of the mask
inline void umma_commit_multicast(uint64_t *bar, uint16_t cta_mask) {
    const uint32_t off = emu::smem_off(bar);
    for (int c = 0; c < emu::S().cluster; ++c)
        if (cta_mask & (1u << c)) mbar_arrive(reinterpret_cast<uint64_t *>(emu::smem_base_of(c) + off));
}

// ---- tcgen05: TMEM -> registers ----------------------------------------------
inline void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
    const uint32_t lane = (taddr >> 16) + (threadIdx.x & 31u), col = taddr & 0xFFFFu;
    if (lane >= 128 || col + 16 > 512) throw std::runtime_error("emu: tcgen05.ld outside tensor memory");
    for (int j = 0; j < 16; ++j) r[j] = emu::T().tmem[lane][col + j];
}
inline void tmem_ld_wait() {}

// ---- models of the three inline-PTX helpers that live in ts.cuh (build.py routes them here) ----------------
// tcgen05.st 32x32b.x16: thread t of the warp writes lane (taddr.lane + t), 16 columns
inline void model_tmem_st16(uint32_t taddr, const uint32_t (&r)[16]) {
    const uint32_t lane = (taddr >> 16) + (threadIdx.x & 31u), col = taddr & 0xFFFFu;
    if (lane >= 128 || col + 16 > 512) throw std::runtime_error("emu: tcgen05.st outside tensor memory");
    for (int j = 0; j < 16; ++j) emu::T().tmem[lane][col + j] = r[j];
}
// tcgen05.mma with the A operand in tensor memory: row m of A is lane m, one K = 16 step is 8 consecutive
// 32-bit columns holding the 16 sixteen-bit elements in order (element 2c in the low half of column c)
inline void model_umma_f16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
    const int n = (int)((idesc >> 17) & 0x3Fu) << 3, m = (int)((idesc >> 24) & 0x1Fu) << 4;
    const bool a_bf16 = ((idesc >> 7) & 7u) == 1u, b_bf16 = ((idesc >> 10) & 7u) == 1u;
    const uint32_t lane0 = tmem_d >> 16, col0 = tmem_d & 0xFFFFu;
    const uint32_t alane0 = tmem_a >> 16, acol0 = tmem_a & 0xFFFFu;
    if (lane0 + (uint32_t)m > 128 || col0 + (uint32_t)n > 512 || alane0 + (uint32_t)m > 128 || acol0 + 8 > 512)
        throw std::runtime_error("emu: MMA outside tensor memory");
    float b[256][16];
    for (int j = 0; j < n; ++j)
        for (int kk = 0; kk < 16; ++kk) b[j][kk] = emu_operand(desc_b, j, kk, b_bf16);
    for (int i = 0; i < m; ++i) {
        float a[16];
        for (int kk = 0; kk < 16; ++kk) {
            const uint32_t cell = emu::T().tmem[alane0 + i][acol0 + kk / 2];
            const uint16_t h = (uint16_t)(kk & 1 ? cell >> 16 : cell & 0xFFFFu);
            a[kk] = emu_elem16(reinterpret_cast<const unsigned char *>(&h), a_bf16);
        }
        for (int j = 0; j < n; ++j) {
            float acc = 0.0f;
            for (int kk = 0; kk < 16; ++kk) acc = std::fmaf(a[kk], b[j][kk], acc);
            uint32_t &cell = emu::T().tmem[lane0 + i][col0 + j];
            float d;
            std::memcpy(&d, &cell, 4);
            d = accumulate ? d + acc : acc;
            std::memcpy(&cell, &d, 4);
        }
    }
}

}  // namespace ptx
}  // namespace vqa
