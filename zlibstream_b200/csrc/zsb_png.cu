// zsb_png.cu -- PNG front end on the device: scanline filtering and the CRC-32 of the IDAT chunks, so that
// pixels never leave the GPU uncompressed.  The caller this replaces is the PNG encoder that sits on top of the reference in
// ImageSharp (/root/reference/readme.md:22); the filters are those of the PNG specification (section 9: None, Sub, Up,
// Average, Paeth), "adaptive" is the minimum-sum-of-absolute-differences heuristic of section 12.8 that ImageSharp and libpng
// use.  The filtered scanlines then go through zsb_deflate_segments unchanged (zsb_api.cu: zsb_png_encode).
#include "zsb_kernels.cuh"

namespace zsb {

namespace {

__device__ __forceinline__ int paeth(int a, int b, int c)
{
    const int p = a + b - c, pa = abs(p - a), pb = abs(p - b), pc = abs(p - c);
    return (pa <= pb && pa <= pc) ? a : (pb <= pc ? b : c);
}

__device__ __forceinline__ uint8_t filter_byte(int f, int raw, int a, int b, int c)
{
    switch (f) {
        case 1: return (uint8_t)(raw - a);
        case 2: return (uint8_t)(raw - b);
        case 3: return (uint8_t)(raw - ((a + b) >> 1));
        case 4: return (uint8_t)(raw - paeth(a, b, c));
        default: return (uint8_t)raw;
    }
}

constexpr int PF_THREADS = 256;

}  // namespace

// One CTA per scanline.  Filtering reads RAW neighbours (left, up, upper left), so every byte is independent work.
__global__ void __launch_bounds__(PF_THREADS) png_filter_kernel(PngFilterParams P)
{
    __shared__ unsigned long long red[5][PF_THREADS / 32];
    __shared__ int chosen_s;
    const uint32_t row = blockIdx.x;
    const uint8_t* cur = P.px + (size_t)row * P.rowbytes;
    const uint8_t* up = row ? cur - P.rowbytes : nullptr;
    uint8_t* dst = P.out + (size_t)row * (P.rowbytes + 1);
    const int tid = threadIdx.x, bpp = (int)P.bpp;
    int f = P.filter;
    if (f == 5) {
        // sum of |signed filtered byte| per filter type; the smallest wins, ties go to the lower type (None, Sub, Up, ...)
        unsigned long long s[5] = {0, 0, 0, 0, 0};
        for (uint32_t x = tid; x < P.rowbytes; x += PF_THREADS) {
            const int raw = cur[x], a = x >= (uint32_t)bpp ? cur[x - bpp] : 0, b = up ? up[x] : 0, c = (up && x >= (uint32_t)bpp) ? up[x - bpp] : 0;
#pragma unroll
            for (int k = 0; k < 5; k++) {
                const int v = (int8_t)filter_byte(k, raw, a, b, c);
                s[k] += (unsigned)(v < 0 ? -v : v);
            }
        }
#pragma unroll
        for (int k = 0; k < 5; k++) {
            for (int off = 16; off > 0; off >>= 1) s[k] += __shfl_down_sync(0xffffffffu, s[k], off);
            if ((tid & 31) == 0) red[k][tid >> 5] = s[k];
        }
        __syncthreads();
        if (tid == 0) {
            unsigned long long best = ~0ull;
            int bi = 0;
            for (int k = 0; k < 5; k++) {
                unsigned long long t = 0;
                for (int w = 0; w < PF_THREADS / 32; w++) t += red[k][w];
                if (t < best) { best = t; bi = k; }
            }
            chosen_s = bi;
        }
        __syncthreads();
        f = chosen_s;
    }
    if (tid == 0) dst[0] = (uint8_t)f;
    for (uint32_t x = tid; x < P.rowbytes; x += PF_THREADS) {
        const int raw = cur[x], a = x >= (uint32_t)bpp ? cur[x - bpp] : 0, b = up ? up[x] : 0, c = (up && x >= (uint32_t)bpp) ? up[x - bpp] : 0;
        dst[1 + x] = filter_byte(f, raw, a, b, c);
    }
}

cudaError_t launch_png_filter(const PngFilterParams& P, cudaStream_t stream)
{
    if (P.height == 0) return cudaSuccess;
    png_filter_kernel<<<P.height, PF_THREADS, 0, stream>>>(P);
    return cudaGetLastError();
}

// ---- CRC-32 (IEEE 802.3, the PNG chunk checksum) ------------------------------------------------------------------------------
// crc(A || B) = x^(8 |B|) * crc(A) + crc(B) in GF(2)[x] / p(x) (bit-reflected, the way zlib's crc32_combine states it), so
// the bytes are cut into pieces, every piece gets its own CRC and the pieces fold in any grouping.
namespace {

constexpr uint32_t CRC_POLY = 0xEDB88320u;
constexpr int CRC_THREADS = 256;
constexpr uint32_t CRC_PIECE = kCrcChunk / CRC_THREADS;   // 256 bytes per thread

__host__ __device__ inline uint32_t crc_multmodp(uint32_t a, uint32_t b)
{
    uint32_t m = 1u << 31, p = 0;
    for (;;) {
        if (a & m) {
            p ^= b;
            if ((a & (m - 1)) == 0) break;
        }
        m >>= 1;
        b = (b & 1) ? (b >> 1) ^ CRC_POLY : b >> 1;
    }
    return p;
}

__device__ __forceinline__ uint32_t crc_bytes(const uint32_t* tab, const uint8_t* p, uint32_t n)
{
    uint32_t c = 0xFFFFFFFFu;
    for (uint32_t i = 0; i < n; i++) c = tab[(c ^ p[i]) & 0xFF] ^ (c >> 8);
    return c ^ 0xFFFFFFFFu;
}

}  // namespace

// x^(n * 2^k) mod p(x)
uint32_t crc32_x2nmodp(uint64_t n, unsigned k)
{
    static uint32_t x2n[32];
    static bool ready = false;
    if (!ready) {
        uint32_t p = 1u << 30;   // x^1
        x2n[0] = p;
        for (int i = 1; i < 32; i++) x2n[i] = p = crc_multmodp(p, p);
        ready = true;
    }
    uint32_t p = 1u << 31;       // x^0 == 1
    while (n) {
        if (n & 1) p = crc_multmodp(x2n[k & 31], p);
        n >>= 1;
        k++;
    }
    return p;
}

uint32_t crc32_combine_host(uint32_t crc1, uint32_t crc2, uint64_t len2) { return crc_multmodp(crc32_x2nmodp(len2, 3), crc1) ^ crc2; }

uint32_t crc32_host(const uint8_t* p, size_t n)
{
    uint32_t c = 0xFFFFFFFFu;
    for (size_t i = 0; i < n; i++) {
        c ^= p[i];
        for (int k = 0; k < 8; k++) c = (c & 1) ? (c >> 1) ^ CRC_POLY : c >> 1;
    }
    return c ^ 0xFFFFFFFFu;
}

// One CTA per kCrcChunk bytes: 256 pieces of 256 bytes, folded pairwise (level l joins spans of 256 << l bytes; P.level_mult[l]
// = x^(8 * (256 << l))).  The last, shorter chunk is one thread's.
__global__ void __launch_bounds__(CRC_THREADS) crc32_chunks_kernel(const uint8_t* __restrict__ p, size_t n, uint32_t* out, CrcMults M)
{
    __shared__ uint32_t tab[256];
    __shared__ uint32_t part[CRC_THREADS];
    const int tid = threadIdx.x;
    {
        uint32_t c = (uint32_t)tid;
        for (int k = 0; k < 8; k++) c = (c & 1) ? (c >> 1) ^ CRC_POLY : c >> 1;
        tab[tid] = c;
    }
    __syncthreads();
    const size_t c0 = (size_t)blockIdx.x * kCrcChunk;
    const size_t left = n - c0;
    if (left < kCrcChunk) {
        if (tid == 0) out[blockIdx.x] = crc_bytes(tab, p + c0, (uint32_t)left);
        return;
    }
    part[tid] = crc_bytes(tab, p + c0 + (size_t)tid * CRC_PIECE, CRC_PIECE);
    __syncthreads();
    for (int l = 0, stride = 1; stride < CRC_THREADS; l++, stride <<= 1) {
        if ((tid & (2 * stride - 1)) == 0) part[tid] = crc_multmodp(M.level[l], part[tid]) ^ part[tid + stride];
        __syncthreads();
    }
    if (tid == 0) out[blockIdx.x] = part[0];
}

// One thread per IDAT chunk: crc("IDAT" || data) from the per-64-KiB CRCs of the zlib stream; IDAT k holds stream bytes
// [k * idat_bytes, min(n, (k + 1) * idat_bytes)), idat_bytes a multiple of kCrcChunk.
__global__ void png_idat_crc_kernel(const uint32_t* chunk_crc, size_t n, size_t idat_bytes, uint32_t nidat, uint32_t crc_type,
                                    uint32_t* out, CrcMults M)
{
    const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= nidat) return;
    const size_t b = (size_t)k * idat_bytes, e = (b + idat_bytes < n) ? b + idat_bytes : n;
    uint32_t crc = crc_type;
    for (size_t o = b; o < e; o += kCrcChunk) {
        const size_t len = (e - o) < kCrcChunk ? (e - o) : kCrcChunk;
        uint32_t mult = M.chunk;
        if (len < kCrcChunk) {   // x^(8 len) by square-and-multiply from x^8, x^16, ...
            mult = 1u << 31;
            uint32_t sq = M.x8;
            for (size_t r = len; r; r >>= 1) {
                if (r & 1) mult = crc_multmodp(sq, mult);
                sq = crc_multmodp(sq, sq);
            }
        }
        crc = crc_multmodp(mult, crc) ^ chunk_crc[o / kCrcChunk];
    }
    out[k] = crc;
}

CrcMults crc32_mults()
{
    CrcMults M;
    for (int l = 0; l < 8; l++) M.level[l] = crc32_x2nmodp((uint64_t)CRC_PIECE << l, 3);
    M.chunk = crc32_x2nmodp(kCrcChunk, 3);
    M.x8 = crc32_x2nmodp(1, 3);
    return M;
}

cudaError_t launch_crc32_chunks(const uint8_t* p, size_t n, uint32_t* chunk_crc, cudaStream_t stream)
{
    const size_t chunks = (n + kCrcChunk - 1) / kCrcChunk;
    if (!chunks) return cudaSuccess;
    crc32_chunks_kernel<<<(unsigned)chunks, CRC_THREADS, 0, stream>>>(p, n, chunk_crc, crc32_mults());
    return cudaGetLastError();
}

cudaError_t launch_png_idat_crc(const uint32_t* chunk_crc, size_t n, size_t idat_bytes, uint32_t nidat, uint32_t* out, cudaStream_t stream)
{
    if (!nidat) return cudaSuccess;
    static const uint8_t kIdat[4] = {'I', 'D', 'A', 'T'};
    png_idat_crc_kernel<<<(nidat + 63) / 64, 64, 0, stream>>>(chunk_crc, n, idat_bytes, nidat, crc32_host(kIdat, 4), out, crc32_mults());
    return cudaGetLastError();
}

}  // namespace zsb
