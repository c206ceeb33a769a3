// Probe: how does tcgen05.mma resolve the 128B swizzle of a K-major A operand whose descriptor start
// address is NOT 1024-byte aligned (shifted by whole 128-byte rows) and whose 8-row groups are SBO
// bytes apart with SBO not a multiple of 1024?  The answer decides whether a convolution can issue
// one MMA per filter tap straight out of a single halo tile in shared memory (shifted windows).
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -o tools/umma_shift_probe tools/umma_shift_probe.cu
//
// The halo region holds HROWS rows x 64 fp16, stored the way TMA SWIZZLE_128B stores a dense box at
// a 1024-aligned base: 16-byte chunk j of row R lives at chunk (j ^ (R & 7)).
// Hypotheses for the element the MMA reads as A[m][16B-chunk j], with g = m / 8, i = m % 8,
// R = k0 + g * (SBO / 128) + i (physical row), bo = descriptor base_offset:
//   ABS : swizzle from absolute address bits  -> logical chunk j of row R        (what we want)
//   REL : swizzle phase = (i + bo) & 7        -> reads physical chunk j ^ ((i + bo) & 7) of row R
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>

constexpr int HROWS = 384, N = 32, K = 64;

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t sbo, uint32_t bo) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr & 0x3FFFFu) >> 4);
    d |= (uint64_t)1 << 16;
    d |= (uint64_t)(sbo >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)(bo & 7) << 49;
    d |= (uint64_t)2 << 61;
    return d;
}

__global__ void __launch_bounds__(128, 1)
probe(const __half *Ag, const __half *Bg, float *out, int k0, int sbo, int bo) {
    extern __shared__ __align__(1024) uint8_t raw[];
    const uint32_t base = (smem_u32(raw) + 1023u) & ~1023u;
    uint8_t *gen = raw + (base - smem_u32(raw));
    uint8_t *a = gen, *b = gen + HROWS * 128;
    __shared__ uint32_t tmem_slot;
    __shared__ __align__(8) uint64_t bar;
    const int t = threadIdx.x, warp = t >> 5;
    for (int idx = t; idx < HROWS * 8; idx += 128) {
        const int r = idx >> 3, j = idx & 7;
        *reinterpret_cast<uint4 *>(a + r * 128 + ((j ^ (r & 7)) << 4)) = *reinterpret_cast<const uint4 *>(Ag + r * K + j * 8);
    }
    for (int idx = t; idx < N * 8; idx += 128) {
        const int r = idx >> 3, j = idx & 7;
        *reinterpret_cast<uint4 *>(b + r * 128 + ((j ^ (r & 7)) << 4)) = *reinterpret_cast<const uint4 *>(Bg + r * K + j * 8);
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    if (t == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 32;" ::"r"(smem_u32(&tmem_slot)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = tmem_slot;
    if (t == 0) {
        const uint32_t idesc = (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
        const uint64_t da = make_desc(base + (uint32_t)k0 * 128u, (uint32_t)sbo, (uint32_t)bo);
        const uint64_t db = make_desc(base + HROWS * 128, 1024, 0);
        for (int k = 0; k < 4; ++k) {
            uint32_t acc = k > 0;
            asm volatile(
                "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                ::"r"(tmem), "l"(da + 2 * k), "l"(db + 2 * k), "r"(idesc), "r"(acc) : "memory");
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
    }
    {
        uint32_t ok;
        do {
            asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                         : "=r"(ok) : "r"(smem_u32(&bar)) : "memory");
        } while (!ok);
    }
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    uint32_t r[32];
    const uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16);
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
          "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
          "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr) : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    for (int j = 0; j < 32; ++j) out[t * N + j] = __uint_as_float(r[j]);
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 32;" ::"r"(tmem) : "memory");
}

int main() {
    std::vector<__half> A(HROWS * K), B(N * K);
    std::vector<float> Af(HROWS * K), Bf(N * K);
    srand(1);
    for (int i = 0; i < HROWS * K; ++i) { Af[i] = (float)(rand() % 15 - 7); A[i] = __float2half(Af[i]); }
    for (int i = 0; i < N * K; ++i) { Bf[i] = (float)(rand() % 7 - 3); B[i] = __float2half(Bf[i]); }
    __half *dA, *dB;
    float *dO;
    cudaMalloc(&dA, A.size() * 2); cudaMalloc(&dB, B.size() * 2); cudaMalloc(&dO, 128 * N * 4);
    cudaMemcpy(dA, A.data(), A.size() * 2, cudaMemcpyHostToDevice);
    cudaMemcpy(dB, B.data(), B.size() * 2, cudaMemcpyHostToDevice);
    const int smem = HROWS * 128 + N * 128 + 2048;
    cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    const int k0s[] = {0, 1, 2, 3, 5, 8, 10, 13};
    const int sbos[] = {1024, 1280, 1536, 2048, 2304};
    std::vector<float> O(128 * N);
    int n_abs = 0, n_cfg = 0;
    for (int sbo : sbos) for (int k0 : k0s) for (int bom = 0; bom < 2; ++bom) {
        const int bo = bom ? (k0 & 7) : 0;
        if (bom && bo == 0) continue;
        cudaMemset(dO, 0, 128 * N * 4);
        probe<<<1, 128, smem>>>(dA, dB, dO, k0, sbo, bo);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("sbo=%d k0=%d bo=%d CUDA error %s\n", sbo, k0, bo, cudaGetErrorString(e)); return 1; }
        cudaMemcpy(O.data(), dO, 128 * N * 4, cudaMemcpyDeviceToHost);
        // hypotheses
        const char *names[4] = {"ABS", "REL(i+bo)", "REL(i)", "REL(R0-phase)"};
        int ok[4] = {1, 1, 1, 1};
        for (int h = 0; h < 4; ++h) {
            for (int m = 0; m < 128 && ok[h]; ++m) {
                const int g = m / 8, i = m % 8;
                const int R = k0 + g * (sbo / 128) + i;
                for (int n = 0; n < N; ++n) {
                    float s = 0;
                    for (int j = 0; j < 8; ++j) {
                        int lj;  // logical chunk of row R that the MMA reads as K-chunk j
                        if (h == 0) lj = j;
                        else if (h == 1) lj = (j ^ ((i + bo) & 7)) ^ (R & 7);
                        else if (h == 2) lj = (j ^ (i & 7)) ^ (R & 7);
                        else lj = (j ^ ((i + k0) & 7)) ^ (R & 7);
                        for (int c = 0; c < 8; ++c) s += Af[R * K + lj * 8 + c] * Bf[n * K + j * 8 + c];
                    }
                    if (s != O[m * N + n]) { ok[h] = 0; break; }
                }
            }
        }
        printf("sbo=%4d k0=%2d bo=%d :", sbo, k0, bo);
        for (int h = 0; h < 4; ++h) if (ok[h]) printf(" %s", names[h]);
        if (!(ok[0] | ok[1] | ok[2] | ok[3])) printf(" none");
        printf("\n");
        n_abs += ok[0]; ++n_cfg;
    }
    printf("ABS holds in %d of %d configurations\n", n_abs, n_cfg);
    return 0;
}
