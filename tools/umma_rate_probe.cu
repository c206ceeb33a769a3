// Probe: issue-to-completion rate of back-to-back tcgen05.mma (kind::f16, M=128, K=16, SS mode, cta_group::1)
// as a function of N, of whether consecutive MMAs accumulate into the same TMEM tile, and of whether the
// A descriptor is the aligned canonical tile (SBO 1024) or a shifted halo window (start + 11 rows, SBO 1280).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o tools/umma_rate_probe tools/umma_rate_probe.cu
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t sbo) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr & 0x3FFFFu) >> 4);
    d |= (uint64_t)1 << 16;
    d |= (uint64_t)(sbo >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}

// mode bit0: alternate two accumulators; bit1: shifted A window; bit2: rotate over 4 A tiles / 4 B tiles (no operand reuse)
__global__ void __launch_bounds__(128, 1) probe(int N, int mode, int iters, long long *out) {
    extern __shared__ __align__(1024) uint8_t raw[];
    const uint32_t base = (smem_u32(raw) + 1023u) & ~1023u;
    __shared__ uint32_t tmem_slot;
    __shared__ __align__(8) uint64_t bar;
    const int t = threadIdx.x, warp = t >> 5;
    for (int i = t; i < 160 * 1024 / 4; i += 128) reinterpret_cast<uint32_t *>(raw + (base - smem_u32(raw)))[i] = 0x3c003c00u;  // 1.0h
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    if (t == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&tmem_slot)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = tmem_slot;
    if (t == 0) {
        const uint32_t idesc = (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
        const bool alt = mode & 1, shifted = mode & 2, rot = mode & 4;
        const uint32_t a_base = base + (shifted ? 11 * 128 : 0), sbo = shifted ? 1280 : 1024;
        const uint32_t b_base = base + 96 * 1024;
        const long long t0 = clock64();
        for (int i = 0; i < iters; ++i) {
            const uint32_t ai = rot ? (uint32_t)((i >> 2) & 3) * 24576u : 0u;   // 4 A tiles (halo-sized slots)
            const uint32_t bi = rot ? (uint32_t)((i >> 2) & 1) * 32768u : 0u;   // 2 B tiles
            const uint64_t da = make_desc(a_base + ai, sbo) + (uint64_t)(2 * (i & 3));
            const uint64_t db = make_desc(b_base + bi, 1024) + (uint64_t)(2 * (i & 3));
            const uint32_t d = tmem + (alt ? (uint32_t)((i & 1) * N) : 0u);
            asm volatile(
                "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                ::"r"(d), "l"(da), "l"(db), "r"(idesc), "r"(i > 1 ? 1u : 0u) : "memory");
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
        uint32_t ok;
        do {
            asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                         : "=r"(ok) : "r"(smem_u32(&bar)) : "memory");
        } while (!ok);
        out[0] = clock64() - t0;
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem) : "memory");
}

int main() {
    long long *d;
    cudaMalloc(&d, 8 * 148);
    const int smem = 162 * 1024 + 1024;
    cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    const int iters = 4096;
    printf("cycles per tcgen05.mma (M=128, K=16, f16, SS), %d back-to-back, one CTA\n", iters);
    printf("%-6s %-28s %10s %12s\n", "N", "variant", "cyc/mma", "ideal N/2");
    for (int N : {32, 64, 128, 256}) {
        for (int mode = 0; mode < 8; ++mode) {
            if ((mode & 1) && N == 256 && false) continue;
            long long best = 1ll << 60;
            for (int rep = 0; rep < 3; ++rep) {
                probe<<<1, 128, smem>>>(N, mode, iters, d);
                cudaError_t e = cudaDeviceSynchronize();
                if (e != cudaSuccess) { printf("CUDA error %s\n", cudaGetErrorString(e)); return 1; }
                long long c;
                cudaMemcpy(&c, d, 8, cudaMemcpyDeviceToHost);
                if (c < best) best = c;
            }
            char name[64];
            snprintf(name, sizeof(name), "%s acc, %s A%s", (mode & 1) ? "alt2" : "same", (mode & 2) ? "shifted" : "aligned",
                     (mode & 4) ? ", rotating tiles" : "");
            printf("%-6d %-28s %10.1f %12d\n", N, name, (double)best / iters, N / 2);
        }
    }
    // all SMs busy: does the per-MMA cost change when 148 CTAs run (power / clocks)?
    for (int N : {32, 128, 256}) {
        probe<<<148, 128, smem>>>(N, 0, iters, d);
        cudaDeviceSynchronize();
        long long c;
        cudaMemcpy(&c, d, 8, cudaMemcpyDeviceToHost);
        printf("%-6d %-28s %10.1f %12d\n", N, "same acc, aligned, 148 CTAs", (double)c / iters, N / 2);
    }
    return 0;
}
