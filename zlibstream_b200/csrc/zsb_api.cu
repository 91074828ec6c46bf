// zsb_api.cu -- the C ABI of libzsb200 (include/zsb200.h): context, scratch memory, kernel orchestration.
//
// Host-side counterpart of ZLibStream (ZlibStream.cs) + Deflate.Compress / Inflate.Decompress control flow,
// reduced to what a full-flush-segmented stream needs.  No CPU fallback: every entry point needs a device.
#include "../../include/zsb200.h"
#include "zsb_kernels.cuh"

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

using namespace zsb;

namespace {

struct DevBuf {
    void* p = nullptr;
    size_t cap = 0;
    cudaError_t reserve(size_t bytes)
    {
        if (bytes <= cap) return cudaSuccess;
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
        cudaError_t e = cudaMalloc(&p, bytes);
        if (e == cudaSuccess) cap = bytes;
        return e;
    }
    void release()
    {
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
    }
    template <typename T> T* as() const { return reinterpret_cast<T*>(p); }
};

// Deflate.cs:80-98
const LevelCfg kLevels[10] = {{0, 0, 0, 0, FN_STORED},    {4, 4, 8, 4, FN_FAST},        {4, 5, 16, 8, FN_FAST},
                              {4, 6, 32, 32, FN_FAST},    {4, 4, 16, 16, FN_SLOW},      {8, 16, 32, 32, FN_SLOW},
                              {8, 16, 128, 128, FN_SLOW}, {8, 32, 128, 256, FN_SLOW},   {32, 128, 258, 1024, FN_SLOW},
                              {32, 258, 258, 4096, FN_SLOW}};

const char* const kInflateMsg[] = {  // Inflate.cs / InflateBlocks.cs / InfTree.cs / InfCodes.cs wording
    "",
    "invalid block type",
    "invalid stored block lengths",
    "too many length or distance symbols",
    "invalid bit length repeat",
    "oversubscribed dynamic bit lengths tree",
    "incomplete dynamic bit lengths tree",
    "oversubscribed literal/length tree",
    "incomplete literal/length tree",
    "oversubscribed distance tree",
    "incomplete distance tree",
    "empty distance tree with lengths",
    "invalid literal/length code",
    "invalid distance code",
    "invalid distance too far back",
    "buffer error",
    "buffer error",
    "segment does not decode to the segment size"};

}  // namespace

struct zsb_ctx {
    int device = 0;
    int num_sms = 0;
    cudaStream_t own_stream = nullptr;
    cudaStream_t copy_stream = nullptr;     // host <-> device copies of one batch overlap the kernels of its neighbours
    static constexpr int kMaxPipe = 16;          // pieces of a piped host-buffer call (events, totals)
    cudaEvent_t ev_in[kMaxPipe] = {}, ev_done[kMaxPipe] = {};
    uint64_t* h_totals = nullptr;           // pinned: output offset after each batch
    cudaStream_t stream = nullptr;
    std::string err;
    uint64_t launches = 0;
    uint32_t last_first_carry = 0;          // Q14: carried prefix of the first segment of the last inflate call
    InflateParams last_inflate{};           // that call's parameters (device output only): zsb_inflate_fill_carry
    size_t last_inflate_produced = 0;
    bool last_inflate_valid = false;
    // scratch
    DevBuf lists, in_stage, out_stage, sym, blocks, nblocks, codes, layout, seg_adler, seg_end, small, halo_stage, inf_len, inf_status, inf_scratch, inf_carry, idx_cand, png_raw, png_filt, png_z, png_crc, png_out;
    uint16_t* crc_tab = nullptr;
    unsigned long long* lz_stats = nullptr;  // 32 counters, only with ZSB_LZ_STATS=1
    // timing of the last call
    std::vector<cudaEvent_t> ev;  // pairs
    std::vector<int> ev_kind;     // 0 lz77, 1 huffman, 2 inflate
    std::vector<cudaEvent_t> ev_mid;   // between the two level-1 kernels of a batch (pairs with ev[2i], ev[2i+1] of kind 0)
    std::vector<int> ev_mid_of;        // index into ev of the pair's first event
    float ms[5] = {0, 0, 0, 0, 0};     // 0 lz77 stage, 1 huffman stage, 2 inflate, 3 lz77_lists_kernel, 4 lz77_resolve_kernel
    bool ms_valid = true;
};

namespace {

int fail(zsb_ctx* c, int code, const std::string& msg)
{
    if (c) c->err = msg;
    return code;
}
int cuda_fail(zsb_ctx* c, cudaError_t e, const char* what)
{
    return fail(c, e == cudaErrorMemoryAllocation ? ZSB_MEM_ERROR : ZSB_ERRNO, std::string(what) + ": " + cudaGetErrorString(e));
}
// An error return must not leave copies in flight against the caller's host buffers (the piped deflate path queues them on
// copy_stream): both streams are drained before the error goes back.
int cuda_fail_sync(zsb_ctx* c, cudaError_t e, const char* what)
{
    if (c) {
        cudaStreamSynchronize(c->stream);
        if (c->copy_stream) cudaStreamSynchronize(c->copy_stream);
    }
    return cuda_fail(c, e, what);
}
#define CU(call)                                                       \
    do {                                                               \
        cudaError_t e__ = (call);                                      \
        if (e__ != cudaSuccess) return cuda_fail_sync(ctx, e__, #call); \
    } while (0)

void crc32c_slices(uint16_t* out /* 4*256 */)
{
    uint32_t t[4][256];
    for (uint32_t i = 0; i < 256; i++) {
        uint32_t c = i;
        for (int k = 0; k < 8; k++) c = (c >> 1) ^ (0x82F63B78u & (0u - (c & 1u)));
        t[0][i] = c;
    }
    for (int s = 1; s < 4; s++)
        for (uint32_t i = 0; i < 256; i++) t[s][i] = (t[s - 1][i] >> 8) ^ t[0][t[s - 1][i] & 0xFF];
    // crc32c(0, b0 | b1<<8 | b2<<16 | b3<<24) = t[3][b0] ^ t[2][b1] ^ t[1][b2] ^ t[0][b3]
    for (uint32_t i = 0; i < 256; i++) {
        out[i] = (uint16_t)t[3][i];
        out[256 + i] = (uint16_t)t[2][i];
        out[512 + i] = (uint16_t)t[1][i];
        out[768 + i] = (uint16_t)t[0][i];
    }
}

void timing_begin(zsb_ctx* c)
{
    for (auto e : c->ev) cudaEventDestroy(e);
    c->ev.clear();
    c->ev_kind.clear();
    for (auto e : c->ev_mid) cudaEventDestroy(e);
    c->ev_mid.clear();
    c->ev_mid_of.clear();
    for (float& m : c->ms) m = 0;
    c->ms_valid = false;
}
void timing_mark(zsb_ctx* c, int kind, bool start)
{
    cudaEvent_t e;
    cudaEventCreate(&e);
    cudaEventRecord(e, c->stream);
    c->ev.push_back(e);
    if (start) c->ev_kind.push_back(kind);
}
void timing_collect(zsb_ctx* c)
{
    if (c->ms_valid) return;
    for (size_t i = 0; i + 1 < c->ev.size(); i += 2) {
        float m = 0;
        if (cudaEventElapsedTime(&m, c->ev[i], c->ev[i + 1]) == cudaSuccess) c->ms[c->ev_kind[i / 2]] += m;
    }
    for (size_t k = 0; k < c->ev_mid.size(); k++) {
        const size_t i = (size_t)c->ev_mid_of[k];
        float a = 0, b = 0;
        if (i + 1 < c->ev.size() && cudaEventElapsedTime(&a, c->ev[i], c->ev_mid[k]) == cudaSuccess &&
            cudaEventElapsedTime(&b, c->ev_mid[k], c->ev[i + 1]) == cudaSuccess) {
            c->ms[3] += a;
            c->ms[4] += b;
        }
    }
    if (c->ms[0] == 0 && c->ms[3] + c->ms[4] > 0) c->ms[0] = c->ms[3] + c->ms[4];
    else if (c->ev_mid.empty() && c->ms[3] + c->ms[4] > 0) c->ms[0] += c->ms[3] + c->ms[4];   // + a short last segment's kernel
    c->ms_valid = true;
}

__global__ void finalize_stream_kernel(uint8_t* out, size_t out_cap, uint64_t* total_out, const uint32_t* adler,
                                       int write_header, uint32_t header, int write_trailer, int* status, const uint64_t* total_bits)
{
    // zlib header (Deflate.cs:464-493) and, for Finish, the final empty static block `03 00` followed by the
    // big-endian Adler-32 (Deflate.cs:627-635)
    if (*status != 0) return;
    if (write_header) {
        out[0] = (uint8_t)(header >> 8);
        out[1] = (uint8_t)header;
    }
    if (write_trailer == 4) {
        // FlushMode.PartialFlush: the stream stands in the middle of a byte.  The final empty static block (BFINAL 1, type 01,
        // END_BLOCK = 7 zero bits: the 10-bit value 3) goes in at the bit position, Bi_windup pads, then the Adler-32.
        const uint64_t bits = *total_bits;
        const uint64_t end = (bits + 10 + 7) >> 3;
        if (end + 4 > out_cap) { *status = ZSB_BUF_ERROR; *total_out = end + 4; return; }
        const uint32_t v = 3u << (bits & 7);
        out[bits >> 3] |= (uint8_t)v;
        out[(bits >> 3) + 1] |= (uint8_t)(v >> 8);
        const uint32_t a = *adler;
        out[end] = (uint8_t)(a >> 24);
        out[end + 1] = (uint8_t)(a >> 16);
        out[end + 2] = (uint8_t)(a >> 8);
        out[end + 3] = (uint8_t)a;
        *total_out = end + 4;
        return;
    }
    if (write_trailer) {   // 1: `03 00` + Adler-32; 2: the final block is already in the stream (one-piece stream), Adler-32 only;
                           // 3: level 0, the final block is an empty STORED block `01 00 00 FF FF` (Deflate.Stored.cs:80 with eof)
        uint64_t o = *total_out;
        const uint64_t need = write_trailer == 2 ? 4 : (write_trailer == 3 ? 9 : 6);
        if (o + need > out_cap) { *status = ZSB_BUF_ERROR; *total_out = o + need; return; }
        const uint32_t a = *adler;
        if (write_trailer == 1) {
            out[o++] = 0x03;
            out[o++] = 0x00;
        } else if (write_trailer == 3) {
            out[o++] = 0x01;
            out[o++] = 0x00;
            out[o++] = 0x00;
            out[o++] = 0xFF;
            out[o++] = 0xFF;
        }
        out[o] = (uint8_t)(a >> 24);
        out[o + 1] = (uint8_t)(a >> 16);
        out[o + 2] = (uint8_t)(a >> 8);
        out[o + 3] = (uint8_t)a;
        *total_out = o + 4;
    }
}

size_t env_size(const char* name, size_t dflt)
{
    const char* v = getenv(name);
    if (!v || !*v) return dflt;
    return (size_t)strtoull(v, nullptr, 10);
}

}  // namespace

namespace {

// Segment index of a stream that comes without a side-car (the reference's own resynchronisation is
// InflateSync, Inflate.cs:386-448).  Every 00 00 FF FF in the body is a candidate segment START (the byte after it); all
// candidates are decoded side by side in count-only mode, each until the first empty stored block it meets, which gives
// the piece's true end.  Following those ends from the first byte of the body visits exactly the real flush points --
// chance occurrences of the pattern (inside stored or Huffman data) are never reached.  Pieces are then grouped into
// segments of seg_size bytes; zero-length pieces (the repeated markers of Q10) stay with the segment they follow.
int build_segment_index(zsb_ctx* ctx, const uint8_t* d_in, size_t n, uint64_t first, uint64_t body_end, size_t seg_size,
                        std::vector<uint64_t>& ends, uint64_t* total_bytes = nullptr)
{
    ends.clear();
    if (total_bytes) *total_bytes = 0;
    if (body_end <= first) return ZSB_OK;
    cudaStream_t st = ctx->stream;
    size_t cap = std::max<size_t>(1u << 20, (body_end - first) / 64 + 1024);
    CU(ctx->idx_cand.reserve(cap * sizeof(uint64_t)));
    uint8_t* small = ctx->small.as<uint8_t>();
    uint32_t* d_ncand = reinterpret_cast<uint32_t*>(small + 24);
    unsigned int* d_work = reinterpret_cast<unsigned int*>(small + 16);
    uint32_t ncand = 0;
    for (int attempt = 0; attempt < 2; attempt++) {
        CU(cudaMemsetAsync(d_ncand, 0, sizeof(uint32_t), st));
        CU(launch_index_markers(d_in, first, body_end, ctx->idx_cand.as<uint64_t>(), (uint32_t)std::min<size_t>(cap, 0xFFFFFFFFu), d_ncand, st));
        ctx->launches++;
        CU(cudaMemcpyAsync(&ncand, d_ncand, sizeof ncand, cudaMemcpyDeviceToHost, st));
        CU(cudaStreamSynchronize(st));
        if (ncand <= cap) break;
        // more occurrences of 00 00 FF FF than the first guess holds (level-0 streams, data that contains the pattern): the
        // kernel has counted them all, so grow the list to that count and scan again
        cap = ncand;
        CU(ctx->idx_cand.reserve(cap * sizeof(uint64_t)));
    }
    if (ncand > cap) return fail(ctx, ZSB_DATA_ERROR, "too many full-flush marker candidates");
    std::vector<uint64_t> cand(ncand);
    if (ncand) CU(cudaMemcpyAsync(cand.data(), ctx->idx_cand.p, ncand * sizeof(uint64_t), cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    std::sort(cand.begin(), cand.end());
    if (cand.empty() || cand.back() != body_end) return fail(ctx, ZSB_DATA_ERROR, "stream does not end with a full-flush marker");
    const size_t np = cand.size();   // piece k starts at cand[k-1] (k = 0: first); nothing starts at cand[np-1] = body_end
    CU(ctx->inf_scratch.reserve(inflate_lanes_scratch_bytes(ctx->num_sms)));
    CU(ctx->seg_end.reserve(np * sizeof(uint64_t)));
    CU(ctx->inf_len.reserve(np * sizeof(uint32_t)));
    CU(ctx->inf_status.reserve(np * sizeof(int)));
    CU(ctx->idx_cand.reserve(np * sizeof(uint64_t)));
    CU(cudaMemcpyAsync(ctx->seg_end.p, cand.data(), np * sizeof(uint64_t), cudaMemcpyHostToDevice, st));
    CU(cudaMemsetAsync(d_work, 0, sizeof(unsigned int), st));
    InflateParams I;
    I.in = d_in;
    I.n = n;
    I.seg_end = ctx->seg_end.as<uint64_t>();
    I.first_offset = first;
    I.nseg = (uint32_t)np;
    I.seg_size = (uint32_t)seg_size;
    I.out = nullptr;
    I.out_cap = 0;
    I.seg_out_len = ctx->inf_len.as<uint32_t>();
    I.seg_adler = nullptr;
    I.seg_status = ctx->inf_status.as<int>();
    I.work_counter = d_work;
    I.count_only = 1;
    I.count_body_end = body_end;
    I.seg_found_end = ctx->idx_cand.as<uint64_t>();   // the unsorted candidate list is no longer needed
    CU(launch_inflate_lanes(I, ctx->num_sms, ctx->inf_scratch.as<uint8_t>(), ctx->inf_scratch.cap, st));
    ctx->launches++;
    std::vector<int> status(np);
    std::vector<uint32_t> lens(np);
    std::vector<uint64_t> found(np);
    CU(cudaMemcpyAsync(status.data(), ctx->inf_status.p, np * sizeof(int), cudaMemcpyDeviceToHost, st));
    CU(cudaMemcpyAsync(lens.data(), ctx->inf_len.p, np * sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
    CU(cudaMemcpyAsync(found.data(), ctx->idx_cand.p, np * sizeof(uint64_t), cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    uint64_t pos = first, acc = 0;
    while (pos < body_end) {
        size_t k = 0;
        if (pos != first) {
            auto it = std::lower_bound(cand.begin(), cand.end(), pos);
            if (it == cand.end() || *it != pos) return fail(ctx, ZSB_DATA_ERROR, "invalid stored block lengths");
            k = (size_t)(it - cand.begin()) + 1;
        }
        if (k >= np) return fail(ctx, ZSB_DATA_ERROR, "invalid stored block lengths");
        if (status[k] != IE_OK) {
            int e = status[k];
            if (e < 0 || e > IE_SEG_SIZE) e = IE_TRUNCATED;
            if (e == IE_OUTPUT_OVERFLOW) e = IE_SEG_SIZE;   // more than seg_size bytes before the next flush point
            return fail(ctx, e == IE_TRUNCATED ? ZSB_BUF_ERROR : ZSB_DATA_ERROR, kInflateMsg[e]);
        }
        const uint64_t e = found[k];
        if (e <= pos || e > body_end) return fail(ctx, ZSB_DATA_ERROR, "invalid stored block lengths");
        if (total_bytes) *total_bytes += lens[k];
        if (lens[k] == 0 && acc == 0 && !ends.empty()) ends.back() = e;   // Q10: a repeated marker belongs to the segment before it
        else {
            acc += lens[k];
            if (acc > seg_size) return fail(ctx, ZSB_DATA_ERROR, kInflateMsg[IE_SEG_SIZE]);
            if (acc == seg_size) {
                ends.push_back(e);
                acc = 0;
            }
        }
        pos = e;
    }
    if (acc) ends.push_back(body_end);                     // a short last segment
    else if (ends.empty()) ends.push_back(body_end);       // nothing but markers: one empty segment
    return ZSB_OK;
}

}  // namespace

extern "C" {

const char* zsb_version(void) { return "zsb200 0.1.0 sm_100a"; }

int zsb_ctx_create(int device, zsb_ctx** out)
{
    if (!out) return ZSB_STREAM_ERROR;
    *out = nullptr;
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess || count == 0 || device < 0 || device >= count) return ZSB_ERRNO;
    zsb_ctx* ctx = new zsb_ctx();
    ctx->device = device;
    if (cudaSetDevice(device) != cudaSuccess) { delete ctx; return ZSB_ERRNO; }
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) { delete ctx; return ZSB_ERRNO; }
    if (prop.major < 10) { delete ctx; return ZSB_VERSION_ERROR; }  // built for sm_100a only
    ctx->num_sms = prop.multiProcessorCount;
    if (cudaStreamCreateWithFlags(&ctx->own_stream, cudaStreamNonBlocking) != cudaSuccess) { delete ctx; return ZSB_ERRNO; }
    if (cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking) != cudaSuccess) { zsb_ctx_destroy(ctx); return ZSB_ERRNO; }
    for (int i = 0; i < zsb_ctx::kMaxPipe; i++)
        if (cudaEventCreateWithFlags(&ctx->ev_in[i], cudaEventDisableTiming) != cudaSuccess ||
            cudaEventCreateWithFlags(&ctx->ev_done[i], cudaEventDisableTiming) != cudaSuccess) { zsb_ctx_destroy(ctx); return ZSB_ERRNO; }
    if (cudaHostAlloc(reinterpret_cast<void**>(&ctx->h_totals), zsb_ctx::kMaxPipe * sizeof(uint64_t), cudaHostAllocDefault) != cudaSuccess) { zsb_ctx_destroy(ctx); return ZSB_ERRNO; }
    ctx->stream = ctx->own_stream;
    uint16_t tab[1024];
    crc32c_slices(tab);
    if (cudaMalloc(&ctx->crc_tab, sizeof tab) != cudaSuccess ||
        cudaMemcpy(ctx->crc_tab, tab, sizeof tab, cudaMemcpyHostToDevice) != cudaSuccess ||
        ctx->small.reserve(256) != cudaSuccess) {
        zsb_ctx_destroy(ctx);
        return ZSB_MEM_ERROR;
    }
    if (env_size("ZSB_LZ_STATS", 0)) {
        if (cudaMalloc(&ctx->lz_stats, 32 * sizeof(unsigned long long)) != cudaSuccess) { zsb_ctx_destroy(ctx); return ZSB_MEM_ERROR; }
        cudaMemset(ctx->lz_stats, 0, 32 * sizeof(unsigned long long));
    }
    *out = ctx;
    return ZSB_OK;
}

void zsb_ctx_destroy(zsb_ctx* ctx)
{
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    for (auto e : ctx->ev) cudaEventDestroy(e);
    DevBuf* bufs[] = {&ctx->lists, &ctx->in_stage, &ctx->out_stage, &ctx->sym, &ctx->blocks, &ctx->nblocks, &ctx->codes, &ctx->layout,
                      &ctx->seg_adler, &ctx->seg_end, &ctx->small, &ctx->halo_stage, &ctx->inf_len, &ctx->inf_status, &ctx->inf_scratch, &ctx->inf_carry, &ctx->idx_cand, &ctx->png_raw, &ctx->png_filt, &ctx->png_z, &ctx->png_crc, &ctx->png_out};
    for (DevBuf* b : bufs) b->release();
    if (ctx->crc_tab) cudaFree(ctx->crc_tab);
    if (ctx->lz_stats) cudaFree(ctx->lz_stats);
    if (ctx->own_stream) cudaStreamDestroy(ctx->own_stream);
    if (ctx->copy_stream) cudaStreamDestroy(ctx->copy_stream);
    for (int i = 0; i < zsb_ctx::kMaxPipe; i++) {
        if (ctx->ev_in[i]) cudaEventDestroy(ctx->ev_in[i]);
        if (ctx->ev_done[i]) cudaEventDestroy(ctx->ev_done[i]);
    }
    if (ctx->h_totals) cudaFreeHost(ctx->h_totals);
    delete ctx;
}

const char* zsb_last_error(const zsb_ctx* ctx) { return ctx ? ctx->err.c_str() : "no context"; }

int zsb_ctx_set_stream(zsb_ctx* ctx, void* s)
{
    if (!ctx) return ZSB_STREAM_ERROR;
    ctx->stream = s ? reinterpret_cast<cudaStream_t>(s) : ctx->own_stream;
    return ZSB_OK;
}
void* zsb_ctx_get_stream(const zsb_ctx* ctx) { return ctx ? (void*)ctx->stream : nullptr; }
uint64_t zsb_ctx_launch_count(const zsb_ctx* ctx) { return ctx ? ctx->launches : 0; }

int zsb_ctx_last_kernel_ms(const zsb_ctx* cctx, float* lz, float* hf, float* inf)
{
    zsb_ctx* ctx = const_cast<zsb_ctx*>(cctx);
    if (!ctx) return ZSB_STREAM_ERROR;
    cudaStreamSynchronize(ctx->stream);
    timing_collect(ctx);
    if (lz) *lz = ctx->ms[0];
    if (hf) *hf = ctx->ms[1];
    if (inf) *inf = ctx->ms[2];
    return ZSB_OK;
}

int zsb_ctx_last_inflate_carry(const zsb_ctx* ctx, uint32_t* first_segment_carry)
{
    if (!ctx || !first_segment_carry) return ZSB_STREAM_ERROR;
    *first_segment_carry = ctx->last_first_carry;
    return ZSB_OK;
}

int zsb_inflate_fill_carry(zsb_ctx* ctx, uint8_t prev_byte, uint32_t* adler)
{
    if (!ctx) return ZSB_STREAM_ERROR;
    ctx->err.clear();
    if (!ctx->last_inflate_valid) return fail(ctx, ZSB_STREAM_ERROR, "no inflate call with a device output to complete");
    CU(cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    InflateParams I = ctx->last_inflate;
    I.prev_byte = prev_byte;
    CU(launch_inflate_carry_fixup(I, st));
    const size_t produced = ctx->last_inflate_produced;
    const size_t achunks = (produced + kAdlerChunk - 1) / kAdlerChunk;
    CU(ctx->seg_adler.reserve((achunks ? achunks : 1) * sizeof(uint2)));
    uint32_t* d_adler = reinterpret_cast<uint32_t*>(ctx->small.as<uint8_t>() + 12);
    CU(launch_adler32(I.out, produced, 1u, ctx->seg_adler.as<uint2>(), (uint32_t)(achunks ? achunks : 1), d_adler, st));
    ctx->launches += 3;
    uint32_t got = 1;
    CU(cudaMemcpyAsync(&got, d_adler, sizeof got, cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    if (adler) *adler = got;
    ctx->last_first_carry = 0;
    return ZSB_OK;
}

int zsb_ctx_last_kernel_ms_detail(const zsb_ctx* cctx, float* out, int n)
{
    zsb_ctx* ctx = const_cast<zsb_ctx*>(cctx);
    if (!ctx || !out) return ZSB_STREAM_ERROR;
    cudaStreamSynchronize(ctx->stream);
    timing_collect(ctx);
    for (int i = 0; i < n; i++) out[i] = i < 5 ? ctx->ms[i] : 0.f;
    return ZSB_OK;
}

// ---- PNG front end ---------------------------------------------------------------------------------------------
namespace {
int png_channels(int color_type) { return color_type == 0 ? 1 : color_type == 2 ? 3 : color_type == 4 ? 2 : color_type == 6 ? 4 : 0; }
void put_be32(uint8_t* p, uint32_t v) { p[0] = (uint8_t)(v >> 24); p[1] = (uint8_t)(v >> 16); p[2] = (uint8_t)(v >> 8); p[3] = (uint8_t)v; }
__global__ void png_put_crcs_kernel(uint8_t* out, const uint32_t* crc, uint32_t nidat, size_t first_pos, size_t idat_bytes, size_t zlen)
{
    const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= nidat) return;
    // IDAT k: [len 4]["IDAT" 4][data][crc 4]; data of IDAT j < k is idat_bytes long
    const size_t data_end = ((size_t)(k + 1) * idat_bytes < zlen) ? (size_t)(k + 1) * idat_bytes : zlen;
    const size_t pos = first_pos + (size_t)k * 12 + 8 + data_end;
    const uint32_t v = crc[k];
    out[pos] = (uint8_t)(v >> 24);
    out[pos + 1] = (uint8_t)(v >> 16);
    out[pos + 2] = (uint8_t)(v >> 8);
    out[pos + 3] = (uint8_t)v;
}
size_t png_default_idat(size_t zbound) { return zbound <= 0x7FFF0000ull ? 0x7FFF0000ull : (size_t)1 << 30; }
}  // namespace

size_t zsb_png_bound(uint32_t width, uint32_t height, int color_type, size_t seg_size, size_t idat_bytes)
{
    const int ch = png_channels(color_type);
    if (!ch) return 0;
    if (!seg_size) seg_size = 65536;
    const size_t filt = (size_t)height * ((size_t)width * ch + 1);
    const size_t z = zsb_deflate_bound(filt, seg_size);
    if (!idat_bytes) idat_bytes = png_default_idat(z);
    return 8 + 25 + z + ((z + idat_bytes - 1) / idat_bytes + 1) * 12 + 12;
}

int zsb_png_encode(zsb_ctx* ctx, const void* pixels, uint32_t width, uint32_t height, int color_type, int filter, int level,
                   int strategy, size_t seg_size, size_t idat_bytes, unsigned flags, void* out, size_t out_cap, size_t* out_len)
{
    if (!ctx) return ZSB_STREAM_ERROR;
    ctx->err.clear();
    if (out_len) *out_len = 0;
    const int ch = png_channels(color_type);
    if (!ch) return fail(ctx, ZSB_STREAM_ERROR, "color_type must be 0 (gray), 2 (RGB), 4 (gray + alpha) or 6 (RGBA), 8 bits per sample");
    if (!pixels || !out || width == 0 || height == 0) return fail(ctx, ZSB_STREAM_ERROR, "stream error");
    if (filter < 0 || filter > 5) return fail(ctx, ZSB_STREAM_ERROR, "filter must be 0..4 (PNG filter types) or 5 (adaptive)");
    if (!seg_size) seg_size = 65536;
    const size_t rowbytes = (size_t)width * ch;
    if (rowbytes > 0x7FFFFFF0ull) return fail(ctx, ZSB_STREAM_ERROR, "scanline too long");
    const size_t raw_bytes = rowbytes * height, filt_bytes = (rowbytes + 1) * height;
    const size_t zcap = zsb_deflate_bound(filt_bytes, seg_size);
    if (!idat_bytes) idat_bytes = png_default_idat(zcap);
    if (idat_bytes % kCrcChunk || idat_bytes > 0x7FFF0000ull) return fail(ctx, ZSB_STREAM_ERROR, "idat_bytes must be a multiple of 65536 below 2 GiB");
    CU(cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;

    const uint8_t* d_px = reinterpret_cast<const uint8_t*>(pixels);
    if (!(flags & ZSB_IN_DEVICE)) {
        CU(ctx->png_raw.reserve(raw_bytes));
        CU(cudaMemcpyAsync(ctx->png_raw.p, pixels, raw_bytes, cudaMemcpyHostToDevice, st));
        d_px = ctx->png_raw.as<uint8_t>();
    }
    CU(ctx->png_filt.reserve(filt_bytes));
    CU(ctx->png_z.reserve(zcap + 16));
    PngFilterParams F;
    F.px = d_px;
    F.out = ctx->png_filt.as<uint8_t>();
    F.rowbytes = (uint32_t)rowbytes;
    F.height = height;
    F.bpp = (uint32_t)ch;
    F.filter = filter;
    CU(launch_png_filter(F, st));
    ctx->launches++;

    // the filtered scanlines are the deflater's input: full-flush segments, exactly what the reference would emit for them
    size_t zlen = 0;
    uint32_t adler = 0;
    int rc = zsb_deflate_segments(ctx, ctx->png_filt.p, filt_bytes, seg_size, level, strategy, ZSB_HASH_CRC32C, 0, nullptr,
                                  ZSB_IN_DEVICE | ZSB_OUT_DEVICE, ctx->png_z.p, zcap, &zlen, nullptr, &adler);
    if (rc != ZSB_OK) return rc;

    const size_t nchunks = (zlen + kCrcChunk - 1) / kCrcChunk;
    const uint32_t nidat = (uint32_t)((zlen + idat_bytes - 1) / idat_bytes);
    CU(ctx->png_crc.reserve((nchunks + nidat + 2) * sizeof(uint32_t)));
    uint32_t* d_chunk_crc = ctx->png_crc.as<uint32_t>();
    uint32_t* d_idat_crc = d_chunk_crc + nchunks;
    CU(launch_crc32_chunks(ctx->png_z.as<uint8_t>(), zlen, d_chunk_crc, st));
    CU(launch_png_idat_crc(d_chunk_crc, zlen, idat_bytes, nidat, d_idat_crc, st));
    ctx->launches += 2;

    const size_t total = 8 + 25 + zlen + (size_t)nidat * 12 + 12;
    if (out_len) *out_len = total;
    if (total > out_cap) return fail(ctx, ZSB_BUF_ERROR, "buffer error");
    uint8_t* d_dst = reinterpret_cast<uint8_t*>(out);
    if (!(flags & ZSB_OUT_DEVICE)) {
        CU(ctx->png_out.reserve(total));
        d_dst = ctx->png_out.as<uint8_t>();
    }
    // signature, IHDR
    uint8_t head[33] = {0x89, 'P', 'N', 'G', 0x0D, 0x0A, 0x1A, 0x0A, 0, 0, 0, 13, 'I', 'H', 'D', 'R'};
    put_be32(head + 16, width);
    put_be32(head + 20, height);
    head[24] = 8;                    // bit depth
    head[25] = (uint8_t)color_type;
    head[26] = 0;                    // compression method: deflate
    head[27] = 0;                    // filter method 0
    head[28] = 0;                    // no interlace
    put_be32(head + 29, crc32_host(head + 12, 17));
    CU(cudaMemcpyAsync(d_dst, head, sizeof head, cudaMemcpyHostToDevice, st));
    CU(cudaStreamSynchronize(st));   // `head` is a stack buffer
    std::vector<uint8_t> hdrs((size_t)nidat * 8);
    size_t pos = 33;
    for (uint32_t k = 0; k < nidat; k++) {
        const size_t b = (size_t)k * idat_bytes, e = std::min(zlen, b + idat_bytes);
        uint8_t* h = hdrs.data() + (size_t)k * 8;
        put_be32(h, (uint32_t)(e - b));
        h[4] = 'I'; h[5] = 'D'; h[6] = 'A'; h[7] = 'T';
        CU(cudaMemcpyAsync(d_dst + pos, h, 8, cudaMemcpyHostToDevice, st));
        CU(cudaMemcpyAsync(d_dst + pos + 8, ctx->png_z.as<uint8_t>() + b, e - b, cudaMemcpyDeviceToDevice, st));
        pos += 8 + (e - b) + 4;
    }
    if (nidat) {
        png_put_crcs_kernel<<<(nidat + 63) / 64, 64, 0, st>>>(d_dst, d_idat_crc, nidat, 33, idat_bytes, zlen);
        CU(cudaGetLastError());
        ctx->launches++;
    }
    static const uint8_t kIend[12] = {0, 0, 0, 0, 'I', 'E', 'N', 'D', 0xAE, 0x42, 0x60, 0x82};
    CU(cudaMemcpyAsync(d_dst + pos, kIend, 12, cudaMemcpyHostToDevice, st));
    CU(cudaStreamSynchronize(st));
    if (!(flags & ZSB_OUT_DEVICE)) {
        CU(cudaMemcpyAsync(out, d_dst, total, cudaMemcpyDeviceToHost, st));
        CU(cudaStreamSynchronize(st));
    }
    return ZSB_OK;
}

size_t zsb_deflate_bound(size_t n, size_t seg_size)
{
    const size_t segs = seg_size ? (n + seg_size - 1) / seg_size : 1;
    // worst case: every block stored (5 bytes per <= 16383 symbols = bytes), per segment up to ~16 bytes of
    // empty blocks and markers, header and trailer
    return n + (n / 16383 + segs + 1) * 5 + segs * 24 + 16;
}

int zsb_deflate_segments(zsb_ctx* ctx, const void* in, size_t n, size_t seg_size, int level, int strategy,
                         int hash_variant, uint64_t first_segment_index, const void* halo, unsigned flags, void* out,
                         size_t out_cap, size_t* out_len, uint64_t* seg_end_offsets, uint32_t* adler)
{
    if (!ctx) return ZSB_STREAM_ERROR;
    ctx->err.clear();
    if (out_len) *out_len = 0;
    // argument checks mirror Deflate's constructor (Deflate.cs:241-281)
    if (level == -1) level = 6;
    if (level < 0 || level > 9) return fail(ctx, ZSB_STREAM_ERROR, "level");
    if (strategy < 0 || strategy > ZSB_FIXED) return fail(ctx, ZSB_STREAM_ERROR, "strategy");
    if (hash_variant != ZSB_HASH_CRC32C && hash_variant != ZSB_HASH_MUL) return fail(ctx, ZSB_STREAM_ERROR, "hash_variant");
    if ((n && !in) || !out) return fail(ctx, ZSB_STREAM_ERROR, "stream error");
    // One Write with FlushMode.NoFlush + Dispose (the reference's benchmark call pattern, DeflateCorpusBenchmark.cs:86-99): the
    // whole input is one segment of the stream's first kind whose last block carries BFINAL; no flush marker, no `03 00`.
    const bool single = (flags & ZSB_SINGLE_STREAM) != 0;
    // FlushMode.SyncFlush on every Write (Deflate.cs:583-595): the same marker as a full flush, but the hash table is kept, so
    // a segment may refer to the 32 KiB before it.  Segments are then neither independent work (one CTA parses them in
    // order; HuffmanOnly and Rle, which never read the hash table, keep their parallel path) nor independently decodable.
    // FlushMode.PartialFlush (Deflate.cs:566-581): Tr_align (one or two empty static blocks, Trees.cs:658-680) instead of the
    // stored marker; the hash table is kept as for a sync flush and the stream is not even byte aligned between Writes.
    const bool partial_flush = (flags & ZSB_PARTIAL_FLUSH) != 0;
    const bool sync_flush = (flags & ZSB_SYNC_FLUSH) != 0 || partial_flush;
    if (sync_flush && (single || first_segment_index != 0 || level == 0 || (flags & ZSB_NO_HEADER)))
        return fail(ctx, ZSB_STREAM_ERROR, "ZSB_SYNC_FLUSH / ZSB_PARTIAL_FLUSH need the whole stream in one call (header included), level >= 1");
    if (single) {
        if (n > (64u << 20)) return fail(ctx, ZSB_STREAM_ERROR, "a one-piece stream holds at most 64 MiB on the device path");
        if (first_segment_index != 0) return fail(ctx, ZSB_STREAM_ERROR, "a one-piece stream starts at segment 0");
        seg_size = n <= 65536 ? 65536 : (n + 32767) & ~(size_t)32767;
    }
    if (seg_size < 65536 || (seg_size & 32767) != 0 || seg_size > (64u << 20))
        return fail(ctx, ZSB_STREAM_ERROR, "seg_size must be a multiple of 32768, >= 65536 and <= 64 MiB");
    // Rle overrides the level's function even at level 0 (Q13), where memLevel 7 makes blocks end at 8191 symbols
    // (Deflate.cs:19-20,246-249): that combination is not built.
    if (level == 0 && strategy == ZSB_RLE)
        return fail(ctx, ZSB_STREAM_ERROR, "level 0 with the Rle strategy is not implemented on the device path");
    const bool level0 = level == 0;   // DeflateStored
    CU(cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    timing_begin(ctx);

    // the bit packer stores 32-bit words relative to `out`; host buffers are staged into an aligned device buffer anyway
    if ((flags & ZSB_OUT_DEVICE) && (reinterpret_cast<uintptr_t>(out) & 3u)) return fail(ctx, ZSB_STREAM_ERROR, "device output pointer must be 4-byte aligned");
    if ((flags & ZSB_OFFSETS_DEVICE) && (reinterpret_cast<uintptr_t>(seg_end_offsets) & 7u)) return fail(ctx, ZSB_STREAM_ERROR, "device offsets pointer must be 8-byte aligned");
    const bool write_header = !(flags & ZSB_NO_HEADER);
    const bool write_trailer = !(flags & ZSB_NO_TRAILER);
    const size_t nseg = (n + seg_size - 1) / seg_size;
    if (nseg > 0xFFFFFFF0ull) return fail(ctx, ZSB_STREAM_ERROR, "too many segments");

    // ---- device views of input / output ----
    const uint8_t* d_in = reinterpret_cast<const uint8_t*>(in);
    if (!(flags & ZSB_IN_DEVICE) && n) {
        CU(ctx->in_stage.reserve(n));
        d_in = ctx->in_stage.as<uint8_t>();   // filled batch by batch below
    }
    const bool host_in = !(flags & ZSB_IN_DEVICE) && n, host_out = !(flags & ZSB_OUT_DEVICE);
    const uint8_t* d_halo = reinterpret_cast<const uint8_t*>(halo);
    if (halo && !(flags & ZSB_IN_DEVICE)) {
        CU(ctx->halo_stage.reserve(W_SIZE));
        CU(cudaMemcpyAsync(ctx->halo_stage.p, halo, W_SIZE, cudaMemcpyHostToDevice, st));
        d_halo = ctx->halo_stage.as<uint8_t>();
    }
    if (first_segment_index == 0) d_halo = nullptr;
    uint8_t* d_out = reinterpret_cast<uint8_t*>(out);
    if (!(flags & ZSB_OUT_DEVICE)) {
        CU(ctx->out_stage.reserve(out_cap ? out_cap : 1));
        d_out = ctx->out_stage.as<uint8_t>();
    }
    uint64_t* d_seg_end = nullptr;
    if (seg_end_offsets && nseg) {
        if (flags & ZSB_OFFSETS_DEVICE) d_seg_end = seg_end_offsets;
        else {
            CU(ctx->seg_end.reserve(nseg * sizeof(uint64_t)));
            d_seg_end = ctx->seg_end.as<uint64_t>();
        }
    }

    // small device scalars: [0] total_out (u64), [8] status (int), [12] adler (u32), [16] work counter (u32)
    uint8_t* small = ctx->small.as<uint8_t>();
    uint64_t* d_total = reinterpret_cast<uint64_t*>(small);
    int* d_status = reinterpret_cast<int*>(small + 8);
    uint32_t* d_adler = reinterpret_cast<uint32_t*>(small + 12);
    unsigned int* d_work = reinterpret_cast<unsigned int*>(small + 16);
    struct { uint64_t total; int status; uint32_t adler; uint32_t work[3]; } init = {write_header ? 2u : 0u, 0, 1u, {0u, 0u, 0u}};
    CU(cudaMemcpyAsync(small, &init, sizeof init, cudaMemcpyHostToDevice, st));
    if (out_cap < (size_t)(write_header ? 2 : 0)) return fail(ctx, ZSB_BUF_ERROR, "buffer error");

    CU(ctx->seg_adler.reserve((nseg ? nseg : 1) * sizeof(uint2)));

    // ---- batches of segments: bounded symbol scratch ----
    const size_t sym_bytes_per_seg = seg_size * sizeof(uint32_t);
    const uint32_t max_blocks = (uint32_t)(seg_size / BLOCK_SYMS + 2);
    const LevelCfg cfg = kLevels[level];
    // level 1 runs as two kernels with 32 bytes of candidate records per input byte between them (zsb_lz77_lists.cu)
    bool use_lists = false;
    {
        Lz77Params probe{};
        probe.lists = reinterpret_cast<uint4*>(1);
        probe.seg_size = (uint32_t)seg_size;
        probe.strategy = strategy;
        probe.good = cfg.good; probe.lazy = cfg.lazy; probe.nice = cfg.nice; probe.chain = cfg.chain;
        probe.fn = strategy == ZSB_RLE ? FN_RLE : cfg.fn;
        probe.stats = ctx->lz_stats;
        const char* l1 = getenv("ZSB_L1_KERNEL");
        use_lists = lz77_lists_applies(probe) && !(l1 && strcmp(l1, "batch") == 0) && !single && !sync_flush;
    }
    const size_t list_bytes_per_seg = use_lists ? lz77_lists_bytes_per_segment((uint32_t)seg_size) : 0;
    const size_t per_seg = sym_bytes_per_seg + list_bytes_per_seg + max_blocks * (sizeof(BlockRec) + sizeof(BlockCodesDev)) + sizeof(uint32_t) + sizeof(SegLayout);
    // the resolve kernel parses one segment per warp: a batch should hold a few thousand segments (5 k resident warps on B200)
    size_t budget = env_size("ZSB_SCRATCH_BYTES", use_lists ? (size_t)48 << 30 : (size_t)8 << 30);
    size_t batch = budget / per_seg;
    if (batch < 1) batch = 1;
    if (batch > nseg) batch = nseg;
    if (ctx->sym.cap < batch * sym_bytes_per_seg || ctx->lists.cap < batch * list_bytes_per_seg) {
        // the scratch has to grow: clamp the budget to half of what the device can still give (the query costs up to
        // milliseconds on a busy box, so it is not made when the scratch of an earlier call is already big enough)
        size_t free_b = 0, total_b = 0;
        if (cudaMemGetInfo(&free_b, &total_b) == cudaSuccess) {
            const size_t have = free_b + ctx->sym.cap + ctx->blocks.cap + ctx->codes.cap + ctx->lists.cap;
            if (budget > have / 2) budget = have / 2;
        }
        batch = budget / per_seg;
        if (batch < 1) batch = 1;
        if (batch > nseg) batch = nseg;
    }
    // Host buffers: a few batches, so that the upload of batch b + 1 and the download of batch b - 1's output run beside the
    // kernels of batch b.  Batches are whole multiples of the SM count (the LZ77 kernels hand out one segment per SM at a time).
    if (sync_flush && batch < nseg) return fail(ctx, ZSB_MEM_ERROR, "ZSB_SYNC_FLUSH: the stream does not fit one batch of scratch");
    if (partial_flush) {   // segments share their edge bytes: the packer ORs those into a zeroed buffer
        const size_t zb = std::min(out_cap, zsb_deflate_bound(n, seg_size));
        CU(cudaMemsetAsync(d_out, 0, zb, st));
    }
    bool piped = false;
    const size_t scratch_batch = batch;   // segments the scratch budget allows at once
    size_t first_batch = 0;   // piped: the first batch may be smaller than the others (0 = no)
    if ((host_in || host_out) && nseg >= (size_t)ctx->num_sms * 16 && !sync_flush) {
        // about four batches.  Level 1: the resolve kernel parses one segment per warp, 32 warps per SM, so a batch is a whole
        // number of those waves (4096 segments on 4736 warp slots cost a full wave each: 45 -> 56 ms per GiB); what is left over
        // goes FIRST, so that the kernels start after the shortest upload.  Other levels: a multiple of the SM count.
        // When the scratch holds the whole call (split schedule below: kernel A per piece, ONE resolve over everything), only
        // kernel A runs per piece and it hands out one segment per SM: eight pieces, so that it starts after an eighth of the upload.
        const bool whole_call = use_lists && batch >= nseg;
        const size_t unit = (use_lists && !whole_call) ? (size_t)ctx->num_sms * 32 : (size_t)ctx->num_sms;
        const size_t nb_target = env_size("ZSB_PIPE_BATCHES", whole_call ? (size_t)zsb_ctx::kMaxPipe : 4);
        size_t k = ((nseg + unit - 1) / unit + nb_target - 1) / nb_target;   // => at most nb_target pieces
        if (k < 1) k = 1;
        size_t pb = k * unit;
        if (pb < batch) {
            batch = pb;
            first_batch = nseg % batch;
        }
        piped = (nseg + batch - 1) / batch <= (size_t)zsb_ctx::kMaxPipe;
        if (!piped) first_batch = 0;
    }
    auto batch_len = [&](size_t s0) { return (s0 == 0 && first_batch) ? first_batch : batch; };
    const bool split_ab = piped && use_lists && scratch_batch >= nseg;   // scratch (records, symbols) for every segment of the call
    const size_t scratch_segs = split_ab ? nseg : batch;
    if (host_in) {
        if (piped) {
            int b = 0;
            for (size_t s0 = 0; s0 < nseg; s0 += batch_len(s0), b++) {
                const size_t off = s0 * seg_size, len = std::min(n - off, batch_len(s0) * seg_size);
                CU(cudaMemcpyAsync(ctx->in_stage.as<uint8_t>() + off, reinterpret_cast<const uint8_t*>(in) + off, len, cudaMemcpyHostToDevice, ctx->copy_stream));
                CU(cudaEventRecord(ctx->ev_in[b], ctx->copy_stream));
            }
        } else CU(cudaMemcpyAsync(ctx->in_stage.p, in, n, cudaMemcpyHostToDevice, st));
    }
    const uint64_t hdr_bytes = write_header ? 2u : 0u;
    uint64_t copied_to = hdr_bytes;   // host output bytes [hdr_bytes, copied_to) are already on their way
    if (nseg) {
        CU(ctx->sym.reserve(scratch_segs * sym_bytes_per_seg));
        if (use_lists) CU(ctx->lists.reserve(scratch_segs * list_bytes_per_seg));
        CU(ctx->blocks.reserve(scratch_segs * max_blocks * sizeof(BlockRec)));
        CU(ctx->codes.reserve(scratch_segs * max_blocks * sizeof(BlockCodesDev)));
        CU(ctx->nblocks.reserve(scratch_segs * sizeof(uint32_t)));
        CU(ctx->layout.reserve(scratch_segs * sizeof(SegLayout)));
    }

    // scratch_seg0 = index of segment s0 inside the scratch arrays: 0 when every batch reuses them, s0 when they hold the whole call
    auto lz_params = [&](size_t s0, uint32_t cnt, size_t scratch_seg0) {
        Lz77Params L;
        L.in = d_in + s0 * seg_size;
        L.n = n - s0 * seg_size;
        L.seg_size = (uint32_t)seg_size;
        L.nseg = cnt;
        L.first_seg_index = first_segment_index + s0;
        // the 32768 bytes before the batch's first segment: the caller's halo for the first batch, the input itself afterwards
        // (Rle looks one byte back, Q14; a short last segment sees the previous window, Q7)
        L.halo = s0 == 0 ? d_halo : d_in + s0 * seg_size - W_SIZE;
        L.strategy = strategy;
        L.hash_variant = hash_variant;
        L.good = cfg.good; L.lazy = cfg.lazy; L.nice = cfg.nice; L.chain = cfg.chain;
        L.fn = strategy == ZSB_RLE ? FN_RLE : cfg.fn;
        L.sym = ctx->sym.as<uint32_t>() + scratch_seg0 * seg_size;
        L.sym_stride = seg_size;
        L.blocks = ctx->blocks.as<BlockRec>() + scratch_seg0 * max_blocks;
        L.max_blocks = max_blocks;
        L.nblocks = ctx->nblocks.as<uint32_t>() + scratch_seg0;
        L.seg_adler = ctx->seg_adler.as<uint2>() + s0;
        L.work_counter = d_work;
        L.crc_tab = ctx->crc_tab;
        L.stats = ctx->lz_stats;
        L.lists = use_lists ? ctx->lists.as<uint4>() + scratch_seg0 * seg_size * 2 : nullptr;
        L.chained = sync_flush ? 1 : 0;
        return L;
    };
    int bidx = 0;
    auto huffman_stage = [&](size_t s0, uint32_t cnt, size_t scratch_seg0) -> int {
        const Lz77Params L = lz_params(s0, cnt, scratch_seg0);
        HuffParams H;
        H.in = L.in;
        H.n = L.n;
        H.seg_size = L.seg_size;
        H.nseg = cnt;
        H.first_seg_index = L.first_seg_index;
        H.strategy = strategy;
        H.sym = L.sym;
        H.sym_stride = L.sym_stride;
        H.blocks = L.blocks;
        H.max_blocks = max_blocks;
        H.nblocks = L.nblocks;
        H.codes = ctx->codes.as<BlockCodesDev>() + scratch_seg0 * max_blocks;
        H.layout = ctx->layout.as<SegLayout>() + scratch_seg0;
        H.out_base = 0;  // layout_kernel reads the running total from *total_out
        H.total_out = d_total;
        H.seg_end = d_seg_end ? d_seg_end + s0 : nullptr;
        H.out = d_out;
        H.out_cap = out_cap;
        H.status = d_status;
        H.finish_mode = single ? 1 : 0;
        H.partial_flush = partial_flush ? 1 : 0;
        H.total_bits = reinterpret_cast<uint64_t*>(small + 32);
        H.level0 = level0 ? 1 : 0;
        timing_mark(ctx, 1, true);
        CU(launch_block_trees(H, st));
        CU(launch_layout(H, st));
        CU(launch_pack(H, st));
        timing_mark(ctx, 1, false);
        ctx->launches += 3;
        if (piped && host_out) {
            // the output of batch b is final once it is packed (segments are byte aligned; only the 2-byte header and the
            // trailer are written later): note where it ends, and send the previous batch's bytes to the host meanwhile
            CU(cudaMemcpyAsync(ctx->h_totals + bidx, d_total, sizeof(uint64_t), cudaMemcpyDeviceToHost, st));
            CU(cudaEventRecord(ctx->ev_done[bidx], st));
            if (bidx > 0) {
                CU(cudaEventSynchronize(ctx->ev_done[bidx - 1]));
                const uint64_t upto = std::min<uint64_t>(ctx->h_totals[bidx - 1], out_cap);
                if (upto > copied_to) {
                    CU(cudaMemcpyAsync(reinterpret_cast<uint8_t*>(out) + copied_to, d_out + copied_to, upto - copied_to, cudaMemcpyDeviceToHost, ctx->copy_stream));
                    copied_to = upto;
                }
            }
        }
        return ZSB_OK;
    };

    if (split_ab) {
        // Level 1 with host buffers and scratch for the whole call: the list kernel runs piece by piece behind the uploads, the
        // parse runs ONCE over every segment (cutting it into batches costs a partly filled wave of warps per batch: 45 -> 52
        // ms per GiB), the Huffman stage runs piece by piece in front of the downloads.
        const size_t nfull = std::min(nseg, n / seg_size);   // a short last segment takes the one-kernel path
        for (size_t s0 = 0; s0 < nseg; s0 += batch_len(s0), bidx++) {
            const size_t e0 = std::min(nfull, s0 + batch_len(s0));
            if (host_in) CU(cudaStreamWaitEvent(st, ctx->ev_in[bidx], 0));
            if (e0 <= s0) continue;
            Lz77Params L = lz_params(s0, (uint32_t)(e0 - s0), s0);
            CU(cudaMemsetAsync(d_work, 0, 3 * sizeof(unsigned int), st));
            timing_mark(ctx, 3, true);
            CU(launch_lz77_lists(L, ctx->num_sms, st));
            timing_mark(ctx, 3, false);
            ctx->launches++;
        }
        if (nfull) {
            Lz77Params L = lz_params(0, (uint32_t)nfull, 0);
            CU(cudaMemsetAsync(d_work, 0, 3 * sizeof(unsigned int), st));
            timing_mark(ctx, 4, true);
            CU(launch_lz77_resolve(L, ctx->num_sms, st));
            timing_mark(ctx, 4, false);
            ctx->launches++;
        }
        if (nfull < nseg) {
            Lz77Params T = lz_params(nfull, (uint32_t)(nseg - nfull), nfull);
            T.lists = nullptr;
            CU(cudaMemsetAsync(d_work, 0, 3 * sizeof(unsigned int), st));
            timing_mark(ctx, 0, true);
            CU(launch_lz77(T, ctx->num_sms, st));
            timing_mark(ctx, 0, false);
            ctx->launches++;
        }
        bidx = 0;
        for (size_t s0 = 0; s0 < nseg; s0 += batch_len(s0), bidx++) {
            const uint32_t cnt = (uint32_t)((nseg - s0) < batch_len(s0) ? (nseg - s0) : batch_len(s0));
            int rc = huffman_stage(s0, cnt, s0);
            if (rc != ZSB_OK) return rc;
        }
    } else
    for (size_t s0 = 0; s0 < nseg; s0 += batch_len(s0), bidx++) {
        const uint32_t cnt = (uint32_t)((nseg - s0) < batch_len(s0) ? (nseg - s0) : batch_len(s0));
        if (piped && host_in) CU(cudaStreamWaitEvent(st, ctx->ev_in[bidx], 0));
        Lz77Params L = lz_params(s0, cnt, 0);
        CU(cudaMemsetAsync(d_work, 0, 3 * sizeof(unsigned int), st));
        timing_mark(ctx, 0, true);
        if (use_lists) {   // an event between the two level-1 kernels: their times are reported apart
            cudaEvent_t mid;
            CU(cudaEventCreate(&mid));
            ctx->ev_mid.push_back(mid);
            ctx->ev_mid_of.push_back((int)ctx->ev.size() - 1);
            L.ev_mid = mid;
        }
        CU(launch_lz77(L, ctx->num_sms, st));
        timing_mark(ctx, 0, false);
        ctx->launches++;
        int rc = huffman_stage(s0, cnt, 0);
        if (rc != ZSB_OK) return rc;
    }
    // ---- Adler-32 of the whole input, header, trailer ----
    if (level0) {   // no LZ77 kernel has read the input: checksum it here (per 64 KiB chunk, then the same fold)
        const size_t achunks = (n + kAdlerChunk - 1) / kAdlerChunk;
        CU(ctx->seg_adler.reserve((achunks ? achunks : 1) * sizeof(uint2)));
        CU(launch_adler32(d_in, n, 1u, ctx->seg_adler.as<uint2>(), (uint32_t)(achunks ? achunks : 1), d_adler, st));
        ctx->launches++;
    } else CU(launch_adler_fold(ctx->seg_adler.as<uint2>(), (uint32_t)nseg, (uint32_t)seg_size, n, 1u, d_adler, st));
    ctx->launches++;
    // Deflate.cs:466-480 (Q11): old-zlib FLEVEL, no strategy adjustment
    uint32_t header = (8u + ((15u - 8u) << 4)) << 8;
    uint32_t level_flags = ((uint32_t)(level - 1) & 0xff) >> 1;
    if (level_flags > 3) level_flags = 3;
    header |= level_flags << 6;
    header += 31 - (header % 31);
    finalize_stream_kernel<<<1, 1, 0, st>>>(d_out, out_cap, d_total, d_adler, write_header ? 1 : 0, header,
                                            write_trailer ? ((single && nseg) ? 2 : (level0 ? 3 : ((partial_flush && nseg) ? 4 : 1))) : 0, d_status,
                                            reinterpret_cast<const uint64_t*>(small + 32));
    CU(cudaGetLastError());
    ctx->launches++;

    struct { uint64_t total; int status; uint32_t adler; } fin;
    CU(cudaMemcpyAsync(&fin, small, sizeof fin, cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    if (piped) CU(cudaStreamSynchronize(ctx->copy_stream));   // nothing of this call may still be writing into `out` when it returns
    if (out_len) *out_len = (size_t)fin.total;
    if (adler) *adler = fin.adler;
    if (fin.status != 0 || fin.total > out_cap) return fail(ctx, ZSB_BUF_ERROR, "buffer error");
    if (host_out && fin.total) {
        if (piped && copied_to > hdr_bytes && copied_to <= fin.total) {   // header, then what the batches have not sent yet
            if (hdr_bytes) CU(cudaMemcpyAsync(out, d_out, hdr_bytes, cudaMemcpyDeviceToHost, st));
            if (fin.total > copied_to)
                CU(cudaMemcpyAsync(reinterpret_cast<uint8_t*>(out) + copied_to, d_out + copied_to, fin.total - copied_to, cudaMemcpyDeviceToHost, st));
            CU(cudaStreamSynchronize(ctx->copy_stream));
        } else CU(cudaMemcpyAsync(out, d_out, fin.total, cudaMemcpyDeviceToHost, st));
    }
    if (seg_end_offsets && nseg && !(flags & ZSB_OFFSETS_DEVICE))
        CU(cudaMemcpyAsync(seg_end_offsets, d_seg_end, nseg * sizeof(uint64_t), cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    return ZSB_OK;
}

// The final block Finish appends after the last full flush: the empty static block `03 00` (Q12) or, for a level-0 stream, the
// empty STORED block `01 00 00 FF FF` (Deflate.Stored.cs:80 with eof) -- 6 or 9 bytes with the Adler-32.
static int segmented_trailer_len(zsb_ctx* ctx, const void* in, size_t n, unsigned flags, size_t* len)
{
    *len = 6;
    if (n < 9) return ZSB_OK;
    uint8_t tail9[9];
    if (flags & ZSB_IN_DEVICE) CU(cudaMemcpy(tail9, reinterpret_cast<const uint8_t*>(in) + n - 9, 9, cudaMemcpyDeviceToHost));
    else memcpy(tail9, reinterpret_cast<const uint8_t*>(in) + n - 9, 9);
    static const uint8_t kStoredFinal[5] = {0x01, 0x00, 0x00, 0xFF, 0xFF};
    if (!(tail9[3] == 0x03 && tail9[4] == 0x00) && memcmp(tail9, kStoredFinal, 5) == 0) *len = 9;
    return ZSB_OK;
}

int zsb_inflate_segments(zsb_ctx* ctx, const void* in, size_t n, const uint64_t* seg_end_offsets, size_t nseg,
                         size_t seg_size, unsigned flags, void* out, size_t out_cap, size_t* out_len, uint32_t* adler)
{
    if (!ctx) return ZSB_STREAM_ERROR;
    ctx->err.clear();
    if (out_len) *out_len = 0;
    if (!in || (!out && out_cap)) return fail(ctx, ZSB_STREAM_ERROR, "stream error");
    const bool single = (flags & ZSB_SINGLE_STREAM) != 0;
    const bool has_header = !(flags & ZSB_NO_HEADER), has_trailer = !(flags & ZSB_NO_TRAILER);
    // a one-piece stream (written without full flushes) is one segment that ends with its final block; its trailer is the
    // Adler-32 alone (Inflate.cs:286-342)
    size_t trailer_len = has_trailer ? (single ? 4u : 6u) : 0u;
    if (has_trailer && !single) {
        CU(cudaSetDevice(ctx->device));
        int rc = segmented_trailer_len(ctx, in, n, flags, &trailer_len);
        if (rc != ZSB_OK) return rc;
    }
    uint64_t single_end = 0;
    if (single) {
        if (n < (has_header ? 2u : 0u) + trailer_len) return fail(ctx, ZSB_BUF_ERROR, "buffer error");
        if (out_cap > 0xF0000000ull) return fail(ctx, ZSB_STREAM_ERROR, "a one-piece stream may decode to at most 0xF0000000 bytes");
        single_end = n - trailer_len;
        seg_end_offsets = &single_end;
        nseg = 1;
        seg_size = out_cap ? out_cap : 1;
        flags &= ~ZSB_OFFSETS_DEVICE;
    }
    if (seg_size == 0 || (!single && seg_size > (64u << 20))) return fail(ctx, ZSB_STREAM_ERROR, "seg_size");
    if (seg_end_offsets && nseg > 0xFFFFFFF0ull) return fail(ctx, ZSB_STREAM_ERROR, "too many segments");
    CU(cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    timing_begin(ctx);

    // ---- header checks on the host (Inflate.cs:120-178) need the first two bytes ----
    uint8_t hdr[2] = {0, 0};
    const uint8_t* d_in = reinterpret_cast<const uint8_t*>(in);
    // Host buffers, a side-car and enough segments: the call is cut into up to 8 batches of >= 16384 segments (the lane kernel
    // decodes one batch in the time the download of the previous one takes), the upload of batch b + 1 and the download of batch
    // b - 1 run on the copy stream beside the kernel of batch b.  The download (1 / ratio times the upload) bounds the call.
    const size_t pipe_min = env_size("ZSB_INFLATE_PIPE_SEGMENTS", 65536);   // tests lower it
    const bool pipe_inf = !(flags & (ZSB_IN_DEVICE | ZSB_OUT_DEVICE | ZSB_OFFSETS_DEVICE)) && !single && seg_end_offsets && nseg >= pipe_min && nseg >= 2;
    size_t pipe_per = 0;
    constexpr size_t kInflatePieces = 8;
    if (pipe_inf) {
        pipe_per = std::max<size_t>(std::max<size_t>(1, pipe_min / 4), (nseg + kInflatePieces - 1) / kInflatePieces);
        CU(ctx->in_stage.reserve(n ? n : 1));
        d_in = ctx->in_stage.as<uint8_t>();
        if (n >= 2) memcpy(hdr, in, 2);
        int b = 0;
        for (size_t k0 = 0; k0 < nseg; k0 += pipe_per, b++) {
            const size_t k1 = std::min(nseg, k0 + pipe_per);
            const uint64_t c0 = k0 ? seg_end_offsets[k0 - 1] : 0, c1 = k1 == nseg ? n : seg_end_offsets[k1 - 1];
            if (c1 > n || c0 > c1) return fail(ctx, ZSB_BUF_ERROR, "buffer error");
            CU(cudaMemcpyAsync(ctx->in_stage.as<uint8_t>() + c0, reinterpret_cast<const uint8_t*>(in) + c0, c1 - c0, cudaMemcpyHostToDevice, ctx->copy_stream));
            CU(cudaEventRecord(ctx->ev_in[b], ctx->copy_stream));
        }
    } else if (!(flags & ZSB_IN_DEVICE)) {
        CU(ctx->in_stage.reserve(n ? n : 1));
        CU(cudaMemcpyAsync(ctx->in_stage.p, in, n, cudaMemcpyHostToDevice, st));
        d_in = ctx->in_stage.as<uint8_t>();
        if (n >= 2) memcpy(hdr, in, 2);
    } else if (has_header && n >= 2) {
        CU(cudaMemcpyAsync(hdr, d_in, 2, cudaMemcpyDeviceToHost, st));
        CU(cudaStreamSynchronize(st));
    }
    if (has_header) {
        if (n < 2) return fail(ctx, ZSB_BUF_ERROR, "buffer error");
        if ((hdr[0] & 0xf) != 8) return fail(ctx, ZSB_DATA_ERROR, "unknown compression method");
        if ((hdr[0] >> 4) + 8 > 15) return fail(ctx, ZSB_DATA_ERROR, "invalid window size");
        if ((((unsigned)hdr[0] << 8) + hdr[1]) % 31 != 0) return fail(ctx, ZSB_DATA_ERROR, "incorrect header check");
        if (hdr[1] & 0x20) return fail(ctx, ZSB_NEED_DICT, "need dictionary");
    }
    std::vector<uint64_t> host_ends;
    if (!seg_end_offsets) {
        // no side-car: find the segments by their full-flush markers
        if (n < (has_header ? 2u : 0u) + trailer_len) return fail(ctx, ZSB_BUF_ERROR, "buffer error");
        int rc = build_segment_index(ctx, d_in, n, has_header ? 2 : 0, n - trailer_len, seg_size, host_ends);
        if (rc != ZSB_OK) return rc;
        seg_end_offsets = host_ends.data();
        nseg = host_ends.size();
        flags &= ~ZSB_OFFSETS_DEVICE;
        if (nseg > 0xFFFFFFF0ull) return fail(ctx, ZSB_STREAM_ERROR, "too many segments");
    }
    const uint64_t* d_seg_end = seg_end_offsets;
    if (!(flags & ZSB_OFFSETS_DEVICE)) {
        CU(ctx->seg_end.reserve((nseg ? nseg : 1) * sizeof(uint64_t)));
        CU(cudaMemcpyAsync(ctx->seg_end.p, seg_end_offsets, nseg * sizeof(uint64_t), cudaMemcpyHostToDevice, st));
        d_seg_end = ctx->seg_end.as<uint64_t>();
    }
    uint64_t last_end = has_header ? 2 : 0;
    if (nseg) {
        if (flags & ZSB_OFFSETS_DEVICE) {
            CU(cudaMemcpyAsync(&last_end, seg_end_offsets + (nseg - 1), sizeof(uint64_t), cudaMemcpyDeviceToHost, st));
            CU(cudaStreamSynchronize(st));
        } else last_end = seg_end_offsets[nseg - 1];
    }
    if (last_end + trailer_len > n) return fail(ctx, ZSB_BUF_ERROR, "buffer error");

    uint8_t* d_out = reinterpret_cast<uint8_t*>(out);
    if (!(flags & ZSB_OUT_DEVICE)) {
        CU(ctx->out_stage.reserve(out_cap ? out_cap : 1));
        d_out = ctx->out_stage.as<uint8_t>();
    }
    CU(ctx->inf_len.reserve((nseg ? nseg : 1) * sizeof(uint32_t)));
    CU(ctx->inf_status.reserve((nseg ? nseg : 1) * sizeof(int)));
    const size_t total_out = out_cap < nseg * seg_size ? out_cap : nseg * seg_size;
    const size_t achunks = (total_out + kAdlerChunk - 1) / kAdlerChunk;
    CU(ctx->seg_adler.reserve((achunks ? achunks : 1) * sizeof(uint2)));

    uint8_t* small = ctx->small.as<uint8_t>();
    uint32_t* d_adler = reinterpret_cast<uint32_t*>(small + 12);
    unsigned int* d_work = reinterpret_cast<unsigned int*>(small + 16);
    unsigned int* d_carry_flag = reinterpret_cast<unsigned int*>(small + 20);
    CU(cudaMemsetAsync(d_work, 0, 2 * sizeof(unsigned int), st));
    // Two kernels (profiles/r01i_inflate_sweep.txt, r01n_inflate_sweep.txt): a warp per segment decodes one segment fastest and wins while the
    // segments are few; a lane per segment shares every issued instruction 32 ways and wins once there are enough
    // segments to fill the SMs' resident warps.  ZSB_INFLATE_KERNEL=warp|lanes|lanes_small forces one of them.
    const int forced_kernel = [] {   // read per call: the tests switch kernels at run time
        const char* e = getenv("ZSB_INFLATE_KERNEL");
        return !e ? 0 : strcmp(e, "warp") == 0 ? 1 : strcmp(e, "lanes") == 0 ? 2 : strcmp(e, "lanes_small") == 0 ? 3 : 0;
    }();
    const bool use_lanes = forced_kernel ? forced_kernel >= 2 : nseg >= (size_t)ctx->num_sms * 81;   // crossover measured at ~11 k segments
    if (use_lanes || pipe_inf) CU(ctx->inf_scratch.reserve(inflate_lanes_scratch_bytes(ctx->num_sms)));
    CU(ctx->inf_carry.reserve((nseg ? nseg : 1) * sizeof(uint32_t)));

    InflateParams I;
    I.in = d_in;
    I.n = n;
    I.seg_end = d_seg_end;
    I.first_offset = has_header ? 2 : 0;
    I.nseg = (uint32_t)nseg;
    I.seg_size = (uint32_t)seg_size;
    I.out = d_out;
    I.out_cap = total_out;
    I.seg_out_len = ctx->inf_len.as<uint32_t>();
    I.seg_adler = nullptr;
    I.seg_status = ctx->inf_status.as<int>();
    I.work_counter = d_work;
    I.seg_carry = ctx->inf_carry.as<uint32_t>();
    I.carry_flag = d_carry_flag;
    I.continuation = (flags & ZSB_CONTINUATION) ? 1 : 0;
    ctx->last_inflate_valid = false;
    size_t piped_to = 0;   // pipe_inf: output bytes [0, piped_to) are already on their way to the host
    // beyond ~3 segments per lane of the default kernel's residency the small-table instance (12 warps per SM) is faster
    auto launch_one = [&](const InflateParams& J) -> int {
        const bool lanes = forced_kernel ? forced_kernel >= 2 : J.nseg >= (uint32_t)ctx->num_sms * 81;
        const bool small_tables = forced_kernel ? forced_kernel == 3 : J.nseg >= (uint32_t)ctx->num_sms * 7 * 32 * 3;
        timing_mark(ctx, 2, true);
        if (lanes && small_tables) CU(launch_inflate_lanes_small(J, ctx->num_sms, ctx->inf_scratch.as<uint8_t>(), ctx->inf_scratch.cap, st));
        else if (lanes) CU(launch_inflate_lanes(J, ctx->num_sms, ctx->inf_scratch.as<uint8_t>(), ctx->inf_scratch.cap, st));
        else CU(launch_inflate(J, ctx->num_sms, st));
        timing_mark(ctx, 2, false);
        ctx->launches++;
        return ZSB_OK;
    };
    if (pipe_inf) {
        CU(cudaMemsetAsync(d_carry_flag, 0, sizeof(unsigned int), st));
        int b = 0;
        for (size_t k0 = 0; k0 < nseg; k0 += pipe_per, b++) {
            const size_t k1 = std::min(nseg, k0 + pipe_per);
            InflateParams J = I;
            J.seg_end = d_seg_end + k0;
            J.first_offset = k0 ? seg_end_offsets[k0 - 1] : I.first_offset;
            J.nseg = (uint32_t)(k1 - k0);
            J.out = d_out + k0 * seg_size;
            J.out_cap = total_out > k0 * seg_size ? total_out - k0 * seg_size : 0;
            J.seg_out_len = I.seg_out_len + k0;
            J.seg_status = I.seg_status + k0;
            J.seg_carry = I.seg_carry + k0;
            J.continuation = (k0 > 0 || I.continuation) ? 1 : 0;
            CU(cudaStreamWaitEvent(st, ctx->ev_in[b], 0));
            CU(cudaMemsetAsync(d_work, 0, 2 * sizeof(unsigned int), st));
            int rc = launch_one(J);
            if (rc != ZSB_OK) return rc;
            CU(cudaEventRecord(ctx->ev_done[b], st));
            if (k1 < nseg) {   // every segment but the stream's last holds seg_size bytes (a wrong one fails the call below)
                const size_t upto = std::min(out_cap, k1 * seg_size);
                CU(cudaStreamWaitEvent(ctx->copy_stream, ctx->ev_done[b], 0));
                if (upto > piped_to) {
                    CU(cudaMemcpyAsync(reinterpret_cast<uint8_t*>(out) + piped_to, d_out + piped_to, upto - piped_to, cudaMemcpyDeviceToHost, ctx->copy_stream));
                    piped_to = upto;
                }
            }
        }
    } else {
        int rc = launch_one(I);
        if (rc != ZSB_OK) return rc;
    }
    std::vector<int> status(nseg);
    std::vector<uint32_t> lens(nseg);
    uint32_t got_adler = 1;
    uint8_t trailer[9] = {0};
    unsigned int carry_flag = 0;
    CU(cudaMemcpyAsync(&carry_flag, d_carry_flag, sizeof carry_flag, cudaMemcpyDeviceToHost, st));
    if (nseg) {
        CU(cudaMemcpyAsync(status.data(), ctx->inf_status.p, nseg * sizeof(int), cudaMemcpyDeviceToHost, st));
        CU(cudaMemcpyAsync(lens.data(), ctx->inf_len.p, nseg * sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
    }
    if (has_trailer) CU(cudaMemcpyAsync(trailer, d_in + last_end, trailer_len, cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    size_t produced = 0;
    for (size_t k = 0; k < nseg; k++) {
        if (status[k] != IE_OK) {
            int e = status[k];
            if (e < 0 || e > IE_SEG_SIZE) e = IE_TRUNCATED;
            char buf[160];
            snprintf(buf, sizeof buf, "%s (segment %zu)", kInflateMsg[e], k);
            // the reference reports only the message; the segment index is appended for the batch caller
            if (pipe_inf) cudaStreamSynchronize(ctx->copy_stream);   // no copy into the caller's buffer outlives the call
            return fail(ctx, (e == IE_TRUNCATED || e == IE_OUTPUT_OVERFLOW) ? ZSB_BUF_ERROR : ZSB_DATA_ERROR, buf);
        }
        produced += lens[k];
    }
    if (out_len) *out_len = produced;
    if (produced > out_cap) return fail(ctx, ZSB_BUF_ERROR, "buffer error");   // never reached: the kernels stop at out_cap
    ctx->last_first_carry = 0;
    if (flags & ZSB_OUT_DEVICE) {
        ctx->last_inflate = I;
        ctx->last_inflate_produced = produced;
        ctx->last_inflate_valid = true;
    }
    if (carry_flag) {   // Q14: Rle segments that open with a run of the previous segment's last byte
        if (pipe_inf) {
            CU(cudaStreamSynchronize(ctx->copy_stream));
            piped_to = 0;   // the batches went out with placeholders: the completed output is copied once more below
        }
        CU(launch_inflate_carry_fixup(I, st));
        ctx->launches++;
        // the first segment of the call has nothing before it here: its carried prefix is the caller's to fill (sharded streams)
        CU(cudaMemcpyAsync(&ctx->last_first_carry, ctx->inf_carry.p, sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
        CU(cudaStreamSynchronize(st));
    }
    if (pipe_inf && !carry_flag && produced > piped_to) {   // the last batch's bytes leave while the checksum pass runs
        CU(cudaMemcpyAsync(reinterpret_cast<uint8_t*>(out) + piped_to, d_out + piped_to, produced - piped_to, cudaMemcpyDeviceToHost, ctx->copy_stream));
        piped_to = produced;
    }
    // running Adler-32 of the output (InflateBlocks.cs:755,792) as a separate pass over the decoded bytes
    CU(launch_adler32(d_out, produced, 1u, ctx->seg_adler.as<uint2>(), (uint32_t)(achunks ? achunks : 1), d_adler, st));
    ctx->launches += 2;
    CU(cudaMemcpyAsync(&got_adler, d_adler, sizeof got_adler, cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    if (pipe_inf) CU(cudaStreamSynchronize(ctx->copy_stream));
    if (adler) *adler = got_adler;
    if (has_trailer) {
        // segmented: the final block must be the empty static block `03 00` (Q12); then the big-endian Adler-32
        const uint8_t* tw = trailer + (trailer_len - 4);
        if (!single && trailer_len == 6 && (trailer[0] != 0x03 || trailer[1] != 0x00)) return fail(ctx, ZSB_DATA_ERROR, "invalid block type");
        const uint32_t want = ((uint32_t)tw[0] << 24) | ((uint32_t)tw[1] << 16) | ((uint32_t)tw[2] << 8) | tw[3];
        if (want != got_adler) return fail(ctx, ZSB_DATA_ERROR, "incorrect data check");  // Inflate.cs:335-342
    }
    if (!(flags & ZSB_OUT_DEVICE) && produced > piped_to) {
        CU(cudaMemcpyAsync(reinterpret_cast<uint8_t*>(out) + piped_to, d_out + piped_to, produced - piped_to, cudaMemcpyDeviceToHost, st));
        CU(cudaStreamSynchronize(st));
    }
    return has_trailer ? ZSB_STREAM_END : ZSB_OK;
}

int zsb_build_segment_index(zsb_ctx* ctx, const void* in, size_t n, size_t seg_size, unsigned flags, uint64_t* seg_end_offsets,
                            size_t cap, size_t* nseg, uint64_t* uncompressed_size)
{
    if (!ctx) return ZSB_STREAM_ERROR;
    ctx->err.clear();
    if (nseg) *nseg = 0;
    if (uncompressed_size) *uncompressed_size = 0;
    if (!in || !nseg) return fail(ctx, ZSB_STREAM_ERROR, "stream error");
    if (seg_size == 0 || seg_size > (64u << 20)) return fail(ctx, ZSB_STREAM_ERROR, "seg_size");
    CU(cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    const bool has_header = !(flags & ZSB_NO_HEADER), has_trailer = !(flags & ZSB_NO_TRAILER);
    size_t trailer_len = 0;
    if (has_trailer) {
        int rc = segmented_trailer_len(ctx, in, n, flags, &trailer_len);
        if (rc != ZSB_OK) return rc;
    }
    if (n < (has_header ? 2u : 0u) + trailer_len) return fail(ctx, ZSB_BUF_ERROR, "buffer error");
    const uint8_t* d_in = reinterpret_cast<const uint8_t*>(in);
    if (!(flags & ZSB_IN_DEVICE)) {
        CU(ctx->in_stage.reserve(n ? n : 1));
        CU(cudaMemcpyAsync(ctx->in_stage.p, in, n, cudaMemcpyHostToDevice, st));
        d_in = ctx->in_stage.as<uint8_t>();
    }
    std::vector<uint64_t> ends;
    uint64_t total = 0;
    int rc = build_segment_index(ctx, d_in, n, has_header ? 2 : 0, n - trailer_len, seg_size, ends, &total);
    if (rc != ZSB_OK) return rc;
    *nseg = ends.size();
    if (uncompressed_size) *uncompressed_size = total;
    if (seg_end_offsets) {
        if (ends.size() > cap) return fail(ctx, ZSB_BUF_ERROR, "buffer error");
        if (flags & ZSB_OFFSETS_DEVICE) {
            CU(cudaMemcpyAsync(seg_end_offsets, ends.data(), ends.size() * sizeof(uint64_t), cudaMemcpyHostToDevice, st));
            CU(cudaStreamSynchronize(st));
        } else memcpy(seg_end_offsets, ends.data(), ends.size() * sizeof(uint64_t));
    }
    return ZSB_OK;
}

int zsb_adler32(zsb_ctx* ctx, uint32_t seed, const void* p, size_t n, unsigned flags, uint32_t* result)
{
    if (!ctx || !result) return ZSB_STREAM_ERROR;
    ctx->err.clear();
    if (n == 0) { *result = seed; return ZSB_OK; }  // Adler32.cs:63-66
    if (!p) return fail(ctx, ZSB_STREAM_ERROR, "stream error");
    CU(cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    const uint8_t* d_p = reinterpret_cast<const uint8_t*>(p);
    if (!(flags & ZSB_IN_DEVICE)) {
        CU(ctx->in_stage.reserve(n));
        CU(cudaMemcpyAsync(ctx->in_stage.p, p, n, cudaMemcpyHostToDevice, st));
        d_p = ctx->in_stage.as<uint8_t>();
    }
    const size_t chunks = (n + kAdlerChunk - 1) / kAdlerChunk;
    if (chunks > 0xFFFFFFF0ull) return fail(ctx, ZSB_STREAM_ERROR, "input too large");
    CU(ctx->seg_adler.reserve(chunks * sizeof(uint2)));
    uint32_t* d_adler = reinterpret_cast<uint32_t*>(ctx->small.as<uint8_t>() + 12);
    CU(launch_adler32(d_p, n, seed, ctx->seg_adler.as<uint2>(), (uint32_t)chunks, d_adler, st));
    ctx->launches += 2;
    CU(cudaMemcpyAsync(result, d_adler, sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    return ZSB_OK;
}

uint32_t zsb_adler32_combine(uint32_t a1, uint32_t a2, uint64_t len2)
{
    const uint32_t BASE = 65521u;
    const uint32_t rem = (uint32_t)(len2 % BASE);
    uint64_t s1 = (uint64_t)(a1 & 0xFFFF) + (a2 & 0xFFFF) + BASE - 1;
    uint64_t s2 = (uint64_t)(a1 >> 16) + (a2 >> 16) + (uint64_t)rem * (((a1 & 0xFFFF) + BASE - 1) % BASE);
    return (uint32_t)(s1 % BASE) | ((uint32_t)(s2 % BASE) << 16);
}

}  // extern "C"

// ---- diagnostics (tests only): raw copies of the last deflate batch's intermediate buffers -----------------
extern "C" ZSB_API int zsb_debug_fetch(zsb_ctx* ctx, int what, uint64_t byte_offset, void* dst, size_t bytes)
{
    if (!ctx || !dst) return ZSB_STREAM_ERROR;
    const DevBuf* b = nullptr;
    switch (what) {
        case 0: b = &ctx->sym; break;        // uint32 dist | lc << 16, [seg][seg_size]
        case 1: b = &ctx->blocks; break;     // BlockRec [seg][max_blocks]
        case 2: b = &ctx->nblocks; break;    // uint32 [seg]
        case 3: b = &ctx->layout; break;     // SegLayout [seg]
        case 4: b = &ctx->codes; break;      // BlockCodesDev [seg][max_blocks]
        case 5: b = &ctx->seg_adler; break;  // uint2 [seg]
        case 7: b = &ctx->lists; break;      // level 1 candidate records, 32 bytes per input byte of the last batch
        case 6: {                            // LZ77 diagnostics counters (ZSB_LZ_STATS=1), cumulative
            if (!ctx->lz_stats || byte_offset + bytes > 32 * sizeof(unsigned long long)) return fail(ctx, ZSB_BUF_ERROR, "no stats");
            CU(cudaSetDevice(ctx->device));
            CU(cudaMemcpy(dst, reinterpret_cast<const uint8_t*>(ctx->lz_stats) + byte_offset, bytes, cudaMemcpyDeviceToHost));
            return ZSB_OK;
        }
        default: return ZSB_STREAM_ERROR;
    }
    if (byte_offset + bytes > b->cap) return fail(ctx, ZSB_BUF_ERROR, "debug fetch out of range");
    CU(cudaSetDevice(ctx->device));
    CU(cudaMemcpy(dst, reinterpret_cast<const uint8_t*>(b->p) + byte_offset, bytes, cudaMemcpyDeviceToHost));
    return ZSB_OK;
}
