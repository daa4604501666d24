// cta2_probe.cu -- round-2 hardware probe (NOT part of the product): what exactly does
// `tcgen05.mma.cta_group::2` compute, and with which allocation / commit protocol?
//
// Why: the large-batch kernel (ts.cuh) is capped near 0.58 of the tensor peak because with M = 128 query rows per
// CTA every document byte is written to shared memory once (TMA) and read once (MMA B operand): 2 x 64 B/cycle at
// the MMA floor = the whole 128 B/cycle of the SM.  A CTA pair (M = 256, each SM storing and reading HALF of every
// document tile) halves that.  The CPU emulator has no model of cta_group::2 and the guides at hand describe it only
// in outline, so this probe pins the semantics on the real part before any kernel is built on them:
//
//   hypothesis H1 (CUTLASS / DeepGEMM usage):  cluster of 2 CTAs, ranks 0 (leader) and 1
//     A (M = 256 x K): rows   0..127 from CTA 0, rows 128..255 from CTA 1   (same smem / TMEM offset in both)
//     B (N x K)      : rows   0..N/2-1 from CTA 0's shared memory, rows N/2..N-1 from CTA 1's (same offset)
//     D (256 x N)    : rows   0..127 in CTA 0's TMEM lanes, rows 128..255 in CTA 1's, N columns each
//     issued once by the leader; tcgen05.commit.cta_group::2 ... multicast::cluster signals both CTAs.
//
// It runs the SS form (A from shared memory) and the TS form (A from tensor memory, what ts.cuh needs) on small
// exactly-representable integers and prints, per form, whether D matches H1 (and if not, the first mismatches and
// whether a swapped-half layout matches instead).  Every wait is bounded (trap instead of hang).
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o gpurun_out/cta2_probe tools/cta2_probe.cu && gpurun_out/cta2_probe
//   (flags: --alloc-leader-only : only the leader CTA executes tcgen05.alloc.cta_group::2)
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "../vietnamese_qa_system_b200/csrc/ptx.cuh"

using namespace vqa;

constexpr int kN = 64;        // MMA N (documents per pair tile): 32 rows of B per CTA
constexpr int kK = 64;        // one 128-byte swizzle row: 4 MMAs of K = 16
constexpr int kRows = 128;    // A rows per CTA

__device__ __forceinline__ void tmem_alloc2(uint32_t *smem_dst, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(ptx::smem_u32(smem_dst)), "r"(ncols)
                 : "memory");
}
__device__ __forceinline__ void tmem_relinquish2() {
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc2(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void umma2_ss(uint32_t d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t acc) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(d),
        "l"(da), "l"(db), "r"(idesc), "r"(acc)
        : "memory");
}
__device__ __forceinline__ void umma2_ts(uint32_t d, uint32_t a_tmem, uint64_t db, uint32_t idesc, uint32_t acc) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], [%1], %2, %3, p;\n\t}\n" ::"r"(d),
        "r"(a_tmem), "l"(db), "r"(idesc), "r"(acc)
        : "memory");
}
__device__ __forceinline__ void umma2_commit_mc(uint64_t *bar, uint16_t mask) {
    asm volatile(
        "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
            ptx::smem_u32(bar)),
        "h"(mask)
        : "memory");
}
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t (&r)[16]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
        "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};" ::"r"(taddr),
        "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
        "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
        : "memory");
}

// small exactly representable test matrices (global row index, k)
__host__ __device__ inline float a_val(int row, int k) { return (float)(((row * 3 + k * 5) % 7) - 3); }
__host__ __device__ inline float b_val(int row, int k) { return (float)(((row * 5 + k * 2) % 5) - 2); }

__device__ __forceinline__ uint32_t pack2(float a, float b) {
    const __nv_bfloat162 v = __floats2bfloat162_rn(a, b);
    return *reinterpret_cast<const uint32_t *>(&v);
}

// form 0: SS (A from shared memory); form 1: TS (A from tensor memory)
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128, 1)
cta2_probe_kernel(float *out /* [2 ranks][128 lanes][kN] */, int form, int alloc_leader_only) {
    extern __shared__ unsigned char smem_raw[];
    unsigned char *smem = reinterpret_cast<unsigned char *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    unsigned char *a_smem = smem;                     // 128 rows x 128 B (16 KB), K-major, 128B swizzle
    unsigned char *b_smem = smem + 16384;             // 32 rows x 128 B (4 KB): this CTA's half of B
    uint64_t *bar = reinterpret_cast<uint64_t *>(smem + 16384 + 4096);
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(bar + 2);
    const int tid = threadIdx.x, warp = tid >> 5;
    const uint32_t rank = ptx::cluster_ctarank();

    // operands of THIS CTA: A rows rank*128 + r, B rows rank*(kN/2) + j
    {
        const int r = tid;  // one A row per thread
        for (int c = 0; c < 8; ++c) {
            uint32_t w[4];
            for (int e = 0; e < 4; ++e) w[e] = pack2(a_val(rank * kRows + r, c * 8 + 2 * e), a_val(rank * kRows + r, c * 8 + 2 * e + 1));
            *reinterpret_cast<uint4 *>(a_smem + r * 128 + ((c ^ (r & 7)) << 4)) = make_uint4(w[0], w[1], w[2], w[3]);
        }
        if (tid < kN / 2) {
            const int j = tid;
            for (int c = 0; c < 8; ++c) {
                uint32_t w[4];
                for (int e = 0; e < 4; ++e)
                    w[e] = pack2(b_val(rank * (kN / 2) + j, c * 8 + 2 * e), b_val(rank * (kN / 2) + j, c * 8 + 2 * e + 1));
                *reinterpret_cast<uint4 *>(b_smem + j * 128 + ((c ^ (j & 7)) << 4)) = make_uint4(w[0], w[1], w[2], w[3]);
            }
        }
    }
    if (tid == 0) {
        ptx::mbar_init(bar, 1);
        ptx::fence_mbar_init();
    }
    ptx::fence_proxy_async_smem();
    __syncthreads();
    ptx::cluster_sync_all();
    if (warp == 0 && (!alloc_leader_only || rank == 0)) {
        tmem_alloc2(tmem_slot, 128);
        tmem_relinquish2();
    }
    ptx::tc_fence_before_sync();
    __syncthreads();
    ptx::cluster_sync_all();
    ptx::tc_fence_after_sync();
    uint32_t tmem_base = *reinterpret_cast<volatile uint32_t *>(tmem_slot);
    if (alloc_leader_only && rank != 0) tmem_base = 0;  // assumption under test: the pair shares the column range

    // TS form: this CTA's 128 A rows -> TMEM columns 64..95 (kK / 2 = 32 columns), D in columns 0..63
    if (form == 1) {
        const int r = tid;
        const uint32_t trow = tmem_base + ((uint32_t)(warp * 32) << 16) + 64;
        for (int c0 = 0; c0 < kK / 2; c0 += 16) {
            uint32_t w[16];
            for (int j = 0; j < 16; ++j) w[j] = pack2(a_val(rank * kRows + r, (c0 + j) * 2), a_val(rank * kRows + r, (c0 + j) * 2 + 1));
            tmem_st16(trow + c0, w);
        }
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    }
    ptx::tc_fence_before_sync();
    __syncthreads();
    ptx::cluster_sync_all();
    ptx::tc_fence_after_sync();

    // the leader issues 4 MMAs of K = 16 for the whole pair, then one commit that signals both CTAs
    if (rank == 0 && warp == 0) {
        if (ptx::elect_one()) {
            const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(kN >> 3) << 17) | ((uint32_t)(256 >> 4) << 24);
            const uint64_t da0 = ptx::umma_desc_k_sw128(ptx::smem_u32(a_smem));
            const uint64_t db0 = ptx::umma_desc_k_sw128(ptx::smem_u32(b_smem));
            for (int k4 = 0; k4 < kK / 16; ++k4) {
                if (form == 0) umma2_ss(tmem_base, da0 + (uint64_t)(k4 * 2), db0 + (uint64_t)(k4 * 2), idesc, k4 != 0);
                else umma2_ts(tmem_base, tmem_base + 64 + (uint32_t)(k4 * 8), db0 + (uint64_t)(k4 * 2), idesc, k4 != 0);
            }
            umma2_commit_mc(bar, (uint16_t)0x3);
        }
        __syncwarp();
    }
    ptx::mbar_wait(bar, 0);  // bounded spin: traps instead of hanging if the commit never arrives here
    ptx::tc_fence_after_sync();

    // D: lane = row within this CTA, kN columns
    const uint32_t taddr = tmem_base + ((uint32_t)(warp * 32) << 16);
    for (int c0 = 0; c0 < kN; c0 += 16) {
        uint32_t acc[16];
        ptx::tmem_ld16(taddr + c0, acc);
        ptx::tmem_ld_wait();
        for (int j = 0; j < 16; ++j) out[((size_t)rank * kRows + tid) * kN + c0 + j] = __uint_as_float(acc[j]);
    }
    ptx::tc_fence_before_sync();
    __syncthreads();
    ptx::cluster_sync_all();
    if (warp == 0 && (!alloc_leader_only || rank == 0)) {
        ptx::tc_fence_after_sync();
        tmem_dealloc2(tmem_base, 128);
    }
}

static int check(const std::vector<float> &out, int a_swap, int b_swap, bool verbose) {
    int bad = 0;
    for (int r = 0; r < 2; ++r)
        for (int i = 0; i < kRows; ++i)
            for (int j = 0; j < kN; ++j) {
                const int arow = (a_swap ? 1 - r : r) * kRows + i;
                const int half = j / (kN / 2), jj = j % (kN / 2);
                const int brow = (b_swap ? 1 - half : half) * (kN / 2) + jj;
                float want = 0.f;
                for (int k = 0; k < kK; ++k) want += a_val(arow, k) * b_val(brow, k);
                const float got = out[((size_t)r * kRows + i) * kN + j];
                if (got != want) {
                    if (verbose && bad < 6) std::printf("    cta %d row %d col %d: got %g want %g\n", r, i, j, got, want);
                    ++bad;
                }
            }
    return bad;
}

int main(int argc, char **argv) {
    int alloc_leader_only = 0;
    for (int i = 1; i < argc; ++i)
        if (!std::strcmp(argv[i], "--alloc-leader-only")) alloc_leader_only = 1;
    float *out_d = nullptr;
    const size_t n = 2 * (size_t)kRows * kN;
    if (cudaMalloc(&out_d, n * sizeof(float)) != cudaSuccess) {
        std::printf("cudaMalloc failed\n");
        return 1;
    }
    const size_t smem = 1024 + 16384 + 4096 + 64;
    cudaFuncSetAttribute(cta2_probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    for (int form = 0; form < 2; ++form) {
        cudaMemset(out_d, 0xff, n * sizeof(float));
        cta2_probe_kernel<<<2, 128, smem>>>(out_d, form, alloc_leader_only);
        cudaError_t e = cudaDeviceSynchronize();
        std::printf("== %s form, alloc by %s: %s\n", form == 0 ? "SS (A in shared memory)" : "TS (A in tensor memory)",
                    alloc_leader_only ? "the leader only" : "both CTAs", cudaGetErrorString(e));
        if (e != cudaSuccess) return 2;  // a trap poisons the context: stop here
        std::vector<float> out(n);
        cudaMemcpy(out.data(), out_d, n * sizeof(float), cudaMemcpyDeviceToHost);
        const int h1 = check(out, 0, 0, false);
        std::printf("   H1 (A rows and B rows split by CTA rank, D rows split by CTA rank): %s (%d mismatches)\n",
                    h1 == 0 ? "MATCH" : "no", h1);
        if (h1 != 0) {
            check(out, 0, 0, true);
            std::printf("   A halves swapped: %d mismatches; B halves swapped: %d; both: %d\n", check(out, 1, 0, false),
                        check(out, 0, 1, false), check(out, 1, 1, false));
        }
    }
    cudaFree(out_d);
    return 0;
}
