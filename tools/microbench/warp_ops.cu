// Microbenchmark of the warp-level primitives the LZ77 batch parser leans on (single warp, dependent chains).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o warp_ops warp_ops.cu ; run under gpurun.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__global__ void bench(unsigned long long* out, int distinct)
{
    __shared__ uint32_t tab[4096];
    __shared__ uint16_t t16[8192];
    const int lane = threadIdx.x;
    for (int i = lane; i < 4096; i += 32) tab[i] = (i * 2654435761u) >> 20;
    for (int i = lane; i < 8192; i += 32) t16[i] = (uint16_t)((i * 2654435761u) >> 17);
    __syncwarp();
    const int N = 256;
    uint32_t v = (lane % distinct) * 977u + 13u;
    long long t0, t1;
    uint32_t acc = 0;
    // 1. match_any, dependent
    t0 = clock64();
    for (int i = 0; i < N; i++) { uint32_t m = __match_any_sync(0xffffffffu, v + (acc & 1)); acc += m; }
    t1 = clock64();
    if (lane == 0) out[0] = (t1 - t0) / N;
    // 2. shfl dependent
    t0 = clock64();
    for (int i = 0; i < N; i++) { acc = __shfl_sync(0xffffffffu, acc + lane, (acc + i) & 31); }
    t1 = clock64();
    if (lane == 0) out[1] = (t1 - t0) / N;
    // 3. ballot + ffs dependent
    t0 = clock64();
    for (int i = 0; i < N; i++) { uint32_t b = __ballot_sync(0xffffffffu, ((acc >> lane) & 1) == 0); acc = __ffs(b | 0x80000000u) + i; }
    t1 = clock64();
    if (lane == 0) out[2] = (t1 - t0) / N;
    // 4. LDS u32 dependent chain (random index)
    t0 = clock64();
    for (int i = 0; i < N; i++) { acc = tab[(acc + lane) & 4095]; }
    t1 = clock64();
    if (lane == 0) out[3] = (t1 - t0) / N;
    // 5. LDS u16 dependent chain
    t0 = clock64();
    for (int i = 0; i < N; i++) { acc = t16[(acc + lane * 7) & 8191]; }
    t1 = clock64();
    if (lane == 0) out[4] = (t1 - t0) / N;
    // 6. atomicMax smem + syncwarp + read back
    t0 = clock64();
    for (int i = 0; i < N; i++) { uint32_t s = (acc + lane * 131) & 4095; atomicMax(&tab[s], (uint32_t)(i << 5) | lane); __syncwarp(); acc += tab[s]; }
    t1 = clock64();
    if (lane == 0) out[5] = (t1 - t0) / N;
    // 7. dependent ALU chain (lop/shift/add mix), 8 ops per iteration
    t0 = clock64();
    for (int i = 0; i < N; i++) { acc = ((acc << 1) ^ i) + 3; acc = (acc >> 1) | 5; acc = acc * 3 + 1; acc ^= acc >> 7; }
    t1 = clock64();
    if (lane == 0) out[6] = (t1 - t0) / N;
    // 8. reduce_or
    t0 = clock64();
    for (int i = 0; i < N; i++) { acc = __reduce_or_sync(0xffffffffu, acc + lane) + i; }
    t1 = clock64();
    if (lane == 0) out[7] = (t1 - t0) / N;
    // 9. global store (fire and forget) + dependent ALU
    t0 = clock64();
    for (int i = 0; i < N; i++) { out[64 + lane + 32 * (i & 7)] = acc; acc += i; }
    t1 = clock64();
    if (lane == 0) out[8] = (t1 - t0) / N;
    // 10. STS then dependent LDS (same address, store-to-load)
    t0 = clock64();
    for (int i = 0; i < N; i++) { tab[(lane * 33 + i) & 4095] = acc; __syncwarp(); acc += tab[(lane * 33 + i + 1) & 4095]; }
    t1 = clock64();
    if (lane == 0) out[9] = (t1 - t0) / N;
    // 11. popc + funnel + vote chain similar to resolve iteration
    t0 = clock64();
    for (int i = 0; i < N; i++) { uint32_t ge = ~((1u << (acc & 31)) - 1u); uint32_t st = (0x84210842u ^ i) & ge; int nx = st ? __ffs(st) - 1 : 32; acc = nx + __shfl_sync(0xffffffffu, lane + 3, nx & 31); }
    t1 = clock64();
    if (lane == 0) out[10] = (t1 - t0) / N;
    if (lane == 0) out[63] = acc;
}

int main()
{
    unsigned long long* d;
    cudaMalloc(&d, 4096 * 8);
    const char* names[] = {"match_any", "shfl", "ballot+ffs", "lds32 chain", "lds16 chain", "atomicMax+sync+lds", "alu x~10", "reduce_or", "stg+alu", "sts->lds", "resolve-like iter"};
    for (int distinct : {1, 2, 8, 32}) {
        bench<<<1, 32>>>(d, distinct);
        bench<<<1, 32>>>(d, distinct);
        unsigned long long h[16];
        cudaMemcpy(h, d, sizeof h, cudaMemcpyDeviceToHost);
        printf("distinct=%d:", distinct);
        for (int i = 0; i < 11; i++) printf(" %s=%llu", names[i], h[i]);
        printf("\n");
    }
    printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    return 0;
}
