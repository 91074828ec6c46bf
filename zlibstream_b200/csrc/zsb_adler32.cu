// zsb_adler32.cu -- Adler-32 on the device: per-chunk partial sums + an associative fold.
//
// Replaces Adler32.Calculate (Adler32.cs:51-326: scalar / SSSE3 / AVX2 all compute s1 = 1 + sum b,
// s2 = sum of running s1, both mod 65521) and adds the adler32_combine the reference lacks
// (SURVEY section 7.1).  HBM-bound: 1 byte read per input byte, nothing written.
#include "zsb_kernels.cuh"

namespace zsb {

namespace {

constexpr int AD_THREADS = 256;

// (s1, s2) of `len2` bytes appended after a prefix whose running checksum is (a1, b1); the appended
// part was summed with seed 1.  zlib's adler32_combine.
__device__ __forceinline__ uint2 adler_combine(uint2 left, uint2 right, uint64_t len2)
{
    const uint32_t rem = (uint32_t)(len2 % ADLER_BASE);
    uint64_t a = (uint64_t)left.x + right.x + ADLER_BASE - 1;
    uint64_t b = (uint64_t)left.y + right.y + (uint64_t)rem * ((left.x + ADLER_BASE - 1) % ADLER_BASE);
    return make_uint2((uint32_t)(a % ADLER_BASE), (uint32_t)(b % ADLER_BASE));
}

}  // namespace

// One CTA per kAdlerChunk bytes: (s1, s2) with seed 1.
__global__ void __launch_bounds__(AD_THREADS) adler_chunks_kernel(const uint8_t* __restrict__ p, size_t n, uint2* out)
{
    const size_t c0 = (size_t)blockIdx.x * kAdlerChunk;
    const uint32_t len = (uint32_t)((n - c0) < kAdlerChunk ? (n - c0) : kAdlerChunk);
    const uint8_t* q = p + c0;
    const int tid = threadIdx.x;
    unsigned long long s1 = 0, s2 = 0;
    const bool aligned = (reinterpret_cast<uintptr_t>(q) & 15) == 0;
    const uint32_t body = aligned ? (len & ~15u) : 0u;
    for (uint32_t i = tid * 16; i < body; i += AD_THREADS * 16) {
        const uint4 v = __ldg(reinterpret_cast<const uint4*>(q + i));
        const uint32_t w[4] = {v.x, v.y, v.z, v.w};
        uint32_t t1 = 0, t2 = 0;  // t2 = sum (16 - j) * b_j
#pragma unroll
        for (int k = 0; k < 4; k++) {
            const uint32_t b0 = w[k] & 0xFF, b1 = (w[k] >> 8) & 0xFF, b2 = (w[k] >> 16) & 0xFF, b3 = w[k] >> 24;
            t1 += b0 + b1 + b2 + b3;
            t2 += (16 - 4 * k) * b0 + (15 - 4 * k) * b1 + (14 - 4 * k) * b2 + (13 - 4 * k) * b3;
        }
        s1 += t1;
        s2 += (unsigned long long)(len - i - 16) * t1 + t2;
    }
    for (uint32_t i = body + tid; i < len; i += AD_THREADS) {
        const uint32_t b = q[i];
        s1 += b;
        s2 += (unsigned long long)(len - i) * b;
    }
    for (int off = 16; off > 0; off >>= 1) {
        s1 += __shfl_down_sync(0xffffffffu, s1, off);
        s2 += __shfl_down_sync(0xffffffffu, s2, off);
    }
    __shared__ unsigned long long r1[AD_THREADS / 32], r2[AD_THREADS / 32];
    if ((tid & 31) == 0) { r1[tid >> 5] = s1; r2[tid >> 5] = s2; }
    __syncthreads();
    if (tid == 0) {
        unsigned long long t1 = 0, t2 = 0;
        for (int w = 0; w < AD_THREADS / 32; w++) { t1 += r1[w]; t2 += r2[w]; }
        // seed 1: s1 = 1 + sum b ; s2 = len * 1 + sum (len - i) b_i
        out[blockIdx.x] = make_uint2((uint32_t)((1 + t1) % ADLER_BASE), (uint32_t)((len + t2) % ADLER_BASE));
    }
}

// Single CTA: result = seed folded with every part in order.  Each thread folds a contiguous run of
// parts, then the per-thread results are folded in order by thread 0 (nseg / 1024 + 1024 combines).
__global__ void __launch_bounds__(1024) adler_fold_kernel(const uint2* parts, uint32_t nparts, uint32_t part_size, size_t n,
                                                          uint32_t seed, uint32_t* result)
{
    __shared__ uint2 acc_s[1024];
    __shared__ unsigned long long len_s[1024];
    const uint32_t tid = threadIdx.x;
    const uint32_t per = (nparts + 1023) / 1024;
    const uint32_t lo = tid * per, hi = (lo + per < nparts) ? lo + per : nparts;
    uint2 acc = make_uint2(1, 0);
    unsigned long long len = 0;
    for (uint32_t k = lo; k < hi; k++) {
        const size_t off = (size_t)k * part_size;
        const uint64_t l = (n - off) < part_size ? (n - off) : part_size;
        acc = adler_combine(acc, parts[k], l);
        len += l;
    }
    acc_s[tid] = acc;
    len_s[tid] = len;
    __syncthreads();
    if (tid == 0) {
        uint2 r = make_uint2(seed & 0xFFFF, seed >> 16);
        for (uint32_t t = 0; t < 1024; t++)
            if (len_s[t]) r = adler_combine(r, acc_s[t], len_s[t]);
        *result = (r.y << 16) | r.x;
    }
}

cudaError_t launch_adler_fold(const uint2* seg_adler, uint32_t nseg, uint32_t seg_size, size_t n, uint32_t seed,
                              uint32_t* result, cudaStream_t stream)
{
    adler_fold_kernel<<<1, 1024, 0, stream>>>(seg_adler, nseg, seg_size, n, seed, result);
    return cudaGetLastError();
}

cudaError_t launch_adler32(const uint8_t* p, size_t n, uint32_t seed, uint2* scratch, uint32_t scratch_len,
                           uint32_t* result, cudaStream_t stream)
{
    const size_t chunks = (n + kAdlerChunk - 1) / kAdlerChunk;
    if (chunks > scratch_len) return cudaErrorInvalidValue;
    if (chunks) {
        adler_chunks_kernel<<<(unsigned)chunks, AD_THREADS, 0, stream>>>(p, n, scratch);
        cudaError_t e = cudaGetLastError();
        if (e != cudaSuccess) return e;
    }
    return launch_adler_fold(scratch, (uint32_t)chunks, kAdlerChunk, n, seed, result, stream);
}

}  // namespace zsb
