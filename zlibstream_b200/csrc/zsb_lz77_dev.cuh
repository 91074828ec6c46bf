// zsb_lz77_dev.cuh -- device helpers shared by the LZ77 kernels (zsb_lz77.cu, zsb_lz77_lists.cu): bulk async copy
// global -> shared (TMA engine, cp.async.bulk + mbarrier) and the Adler-32 partial sums taken from the staged bytes.
#pragma once

#include "zsb_common.cuh"

namespace zsb {
namespace {

// ---- bulk async copy (TMA engine) global -> shared -------------------------------------------------

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, int count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t phase)
{
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE_%=;\n\t"
        "bra WAIT_%=;\n\t"
        "DONE_%=:\n\t}" ::"r"(smem_u32(bar)),
        "r"(phase)
        : "memory");
}

// Cooperative copy of n bytes global -> shared; uses the bulk-copy engine for the 16-byte aligned body.
template <int NT>
__device__ void stage_input(uint8_t* dst, const uint8_t* src, int n, uint64_t* bar, uint32_t& phase)
{
    const int tid = threadIdx.x;
    const bool aligned = ((reinterpret_cast<uintptr_t>(src) | reinterpret_cast<uintptr_t>(dst)) & 15) == 0;
    int body = aligned ? (n & ~15) : 0;
    if (body > 0) {
        if (tid == 0) {
            mbar_expect_tx(bar, (uint32_t)body);
            bulk_g2s(dst, src, (uint32_t)body, bar);
        }
    }
    for (int i = body + tid; i < n; i += NT) dst[i] = src[i];
    if (body > 0) {
        mbar_wait(bar, phase);
        phase ^= 1;
    }
}

// Partial Adler sums of n staged bytes: s1 = sum b, s2 = sum (n - j) * b_j   (Adler32.cs:270-326 unrolled)
template <int NT>
__device__ void adler_chunk(const uint8_t* p, int n, uint32_t* red, uint32_t& a, uint32_t& b)
{
    const int tid = threadIdx.x;
    unsigned long long s1 = 0, s2 = 0;
    for (int i = tid * 4; i < n; i += NT * 4) {
        if (i + 4 <= n && ((reinterpret_cast<uintptr_t>(p + i) & 3) == 0)) {
            uint32_t v = *reinterpret_cast<const uint32_t*>(p + i);
            uint32_t b0 = v & 0xFF, b1 = (v >> 8) & 0xFF, b2 = (v >> 16) & 0xFF, b3 = v >> 24;
            uint32_t t1 = b0 + b1 + b2 + b3;
            uint32_t w = (uint32_t)(n - i);
            s1 += t1;
            s2 += (unsigned long long)w * t1 - (b1 + 2 * b2 + 3 * b3);
        } else {
            for (int j = i; j < n && j < i + 4; j++) {
                s1 += p[j];
                s2 += (unsigned long long)(n - j) * p[j];
            }
        }
    }
    // warp + block reduce
    for (int off = 16; off > 0; off >>= 1) {
        s1 += __shfl_down_sync(0xffffffffu, s1, off);
        s2 += __shfl_down_sync(0xffffffffu, s2, off);
    }
    unsigned long long* r64 = reinterpret_cast<unsigned long long*>(red);
    const int warp = tid >> 5, lane = tid & 31;
    if (lane == 0) {
        r64[warp * 2] = s1;
        r64[warp * 2 + 1] = s2;
    }
    __syncthreads();
    if (tid == 0) {
        unsigned long long t1 = 0, t2 = 0;
        for (int w = 0; w < NT / 32; w++) {
            t1 += r64[w * 2];
            t2 += r64[w * 2 + 1];
        }
        // running (a, b) over the segment: b += n * a_old + s2, a += s1  (mod 65521)
        unsigned long long nb = (unsigned long long)b + (unsigned long long)(n % ADLER_BASE) * a + t2 % ADLER_BASE;
        b = (uint32_t)(nb % ADLER_BASE);
        a = (uint32_t)((a + t1) % ADLER_BASE);
    }
    __syncthreads();
}


}  // namespace
}  // namespace zsb
