// m64_probe.cu -- hardware probe (NOT part of the product): which tensor-memory lanes does an M = 64
// tcgen05.mma.cta_group::1 use for A (TS form) and D?  BASELINE configs[3] runs 64 queries per CTA: an M = 64
// instruction reads half the A bytes from tensor memory (the read port is what paces ts.cuh's M = 128 x N = 64 tiles).
// A[r][0] = r + 1 (other K = 0), B[j][0] = 1  =>  D[r][j] = r + 1: reading column 0 of all 128 lanes shows where row
// r lands.  TS form: lane t of tensor memory holds t + 1, so D also shows which A lanes are read as which rows.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o gpurun_out/m64_probe tools/m64_probe.cu && gpurun_out/m64_probe
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <cstdio>
#include <vector>

#include "../vietnamese_qa_system_b200/csrc/ptx.cuh"

using namespace vqa;

constexpr int kN = 64;

__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t (&r)[16]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
        "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};" ::"r"(taddr),
        "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
        "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
        : "memory");
}
__device__ __forceinline__ void umma_ts(uint32_t d, uint32_t a_tmem, uint64_t db, uint32_t idesc, uint32_t acc) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}\n" ::"r"(d),
        "r"(a_tmem), "l"(db), "r"(idesc), "r"(acc)
        : "memory");
}
__device__ __forceinline__ uint32_t pack2(float a, float b) {
    const __nv_bfloat162 v = __floats2bfloat162_rn(a, b);
    return *reinterpret_cast<const uint32_t *>(&v);
}

// form 0: SS (A = 64 rows in shared memory); form 1: TS (A in tensor memory, every lane t holds t + 1)
__global__ void __launch_bounds__(128, 1) m64_probe_kernel(float *out /* [128 lanes][2 cols] */, int form, int m) {
    extern __shared__ unsigned char smem_raw[];
    unsigned char *smem = reinterpret_cast<unsigned char *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    unsigned char *a_smem = smem;             // 128 rows x 128 B
    unsigned char *b_smem = smem + 16384;     // 64 rows x 128 B
    uint64_t *bar = reinterpret_cast<uint64_t *>(smem + 16384 + 8192);
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(bar + 2);
    const int tid = threadIdx.x, warp = tid >> 5;
    for (int c = 0; c < 8; ++c) {
        const uint4 z = make_uint4(c == 0 ? pack2((float)(tid + 1), 0.f) : 0u, 0u, 0u, 0u);
        *reinterpret_cast<uint4 *>(a_smem + tid * 128 + ((c ^ (tid & 7)) << 4)) = z;
        if (tid < kN) {
            const uint4 o = make_uint4(c == 0 ? pack2(1.f, 0.f) : 0u, 0u, 0u, 0u);
            *reinterpret_cast<uint4 *>(b_smem + tid * 128 + ((c ^ (tid & 7)) << 4)) = o;
        }
    }
    if (tid == 0) {
        ptx::mbar_init(bar, 1);
        ptx::fence_mbar_init();
    }
    ptx::fence_proxy_async_smem();
    __syncthreads();
    if (warp == 0) {
        ptx::tmem_alloc(tmem_slot, 256);
        ptx::tmem_relinquish();
    }
    ptx::tc_fence_before_sync();
    __syncthreads();
    ptx::tc_fence_after_sync();
    const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t *>(tmem_slot);
    const uint32_t trow = tmem_base + ((uint32_t)(warp * 32) << 16);
    {   // clear D (columns 0..63) and write A (columns 128..135: 16 K-elements as 8 columns of bf16 pairs)
        uint32_t z[16];
        for (int j = 0; j < 16; ++j) z[j] = __float_as_uint(-7.f);
        for (int c0 = 0; c0 < kN; c0 += 16) tmem_st16(trow + c0, z);
        uint32_t w[16];
        for (int j = 0; j < 16; ++j) w[j] = 0u;
        w[0] = pack2((float)(tid + 1), 0.f);
        tmem_st16(trow + 128, w);
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    }
    ptx::tc_fence_before_sync();
    __syncthreads();
    ptx::tc_fence_after_sync();
    if (warp == 0) {
        if (ptx::elect_one()) {
            const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(kN >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
            const uint64_t da = ptx::umma_desc_k_sw128(ptx::smem_u32(a_smem));
            const uint64_t db = ptx::umma_desc_k_sw128(ptx::smem_u32(b_smem));
            if (form == 0) ptx::umma_f16(tmem_base, da, db, idesc, 0u);
            else umma_ts(tmem_base, tmem_base + 128, db, idesc, 0u);
            ptx::umma_commit(bar);
        }
        __syncwarp();
    }
    ptx::mbar_wait(bar, 0);
    ptx::tc_fence_after_sync();
    uint32_t acc[16];
    ptx::tmem_ld16(trow, acc);
    ptx::tmem_ld_wait();
    out[tid * 2 + 0] = __uint_as_float(acc[0]);
    out[tid * 2 + 1] = __uint_as_float(acc[5]);
    ptx::tc_fence_before_sync();
    __syncthreads();
    if (warp == 0) {
        ptx::tc_fence_after_sync();
        ptx::tmem_dealloc(tmem_base, 256);
    }
}

int main() {
    float *out_d = nullptr;
    cudaMalloc(&out_d, 256 * sizeof(float));
    const size_t smem = 1024 + 16384 + 8192 + 64;
    cudaFuncSetAttribute(m64_probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    for (int m : {128, 64}) {
        for (int form = 0; form < 2; ++form) {
            cudaMemset(out_d, 0, 256 * sizeof(float));
            m64_probe_kernel<<<1, 128, smem>>>(out_d, form, m);
            cudaError_t e = cudaDeviceSynchronize();
            std::printf("== M = %d, %s form: %s\n", m, form == 0 ? "SS" : "TS", cudaGetErrorString(e));
            if (e != cudaSuccess) return 2;
            std::vector<float> out(256);
            cudaMemcpy(out.data(), out_d, 256 * sizeof(float), cudaMemcpyDeviceToHost);
            std::printf("   D[lane][col 0] (row + 1 of A that landed in this lane; -7 = untouched):\n");
            for (int l = 0; l < 128; ++l) std::printf("%s%4g", l % 32 == 0 ? "   " : "", out[l * 2]), (l % 32 == 31 ? std::printf("\n") : 0);
            bool same = true;
            for (int l = 0; l < 128; ++l) same = same && out[l * 2] == out[l * 2 + 1];
            std::printf("   column 5 equals column 0 in every lane: %s\n", same ? "yes" : "NO");
        }
    }
    cudaFree(out_d);
    return 0;
}
