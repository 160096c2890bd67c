// microbenchmark: slot-claiming atomics + dependent record stores
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); return 1; } } while (0)
__device__ __forceinline__ uint32_t hash(uint32_t x) { x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16; return x; }
// mode bit0: dependent store (slot from atomic), bit1: do atomic, bit2: do store, bit3: RED (ignore result)
template <int NB>
__global__ void k(uint32_t *cnt, uint4 *slots, size_t canvas, size_t kstride, uint32_t n, int mode, uint32_t jitter, uint32_t W) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint32_t home[NB], kk[NB];
    // thread i -> pixel near (i % W, i / W) displaced by a small pseudo-random jitter per frame
    uint32_t tx = (i >> 5) % (W / 8), ty = (i >> 5) / (W / 8);      // 8x4 tiles
    uint32_t x0 = tx * 8 + (i & 7), y0 = ty * 4 + ((i >> 3) & 3);
#pragma unroll
    for (int s = 0; s < NB; ++s) {
        uint32_t h = hash(i * 4 + s);
        uint32_t x = (x0 + (h % (2 * jitter + 1))) % W, y = (y0 + ((h >> 8) % (2 * jitter + 1))) % W;
        home[s] = y * W + x;
        kk[s] = 0;
    }
    if (mode & 2) {
#pragma unroll
        for (int s = 0; s < NB; ++s) {
            if (mode & 8) atomicAdd(&cnt[s * canvas + home[s]], 1u);
            else kk[s] = atomicAdd(&cnt[s * canvas + home[s]], 1u);
        }
    }
    if (mode & 4) {
#pragma unroll
        for (int s = 0; s < NB; ++s) {
            uint32_t slot = (mode & 1) ? min(kk[s], 2u) : (hash(i + s) % 10 < 7 ? 0 : (hash(i + s) % 10 < 9 ? 1 : 2));
            slots[slot * kstride + s * canvas + home[s]] = make_uint4(i, s, kk[s], 0);
        }
    }
}
template <int NB>
int run() {
    const uint32_t W = 1024, n = W * W;
    size_t canvas = (size_t) W * W, kstride = canvas * NB;
    uint32_t *cnt; uint4 *slots;
    CK(cudaMalloc(&cnt, canvas * NB * 4)); CK(cudaMalloc(&slots, kstride * 3 * 16));
    void *flush; CK(cudaMalloc(&flush, 256 << 20));
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (uint32_t jitter : {1u, 4u}) {
        for (int mode : {4, 7}) {
            float best = 1e9;
            for (int rep = 0; rep < 5; ++rep) {
                CK(cudaMemset(cnt, 0, canvas * NB * 4));
                if (rep == 0) CK(cudaMemset(flush, 1, 256 << 20));
                cudaEventRecord(e0);
                k<NB><<<(n + 255) / 256, 256>>>(cnt, slots, canvas, kstride, n, mode, jitter, W);
                cudaEventRecord(e1);
                CK(cudaEventSynchronize(e1));
                float ms; cudaEventElapsedTime(&ms, e0, e1);
                if (rep > 0 && ms < best) best = ms;
            }
            printf("NB %d jitter %3u mode %2d (%s%s%s%s): %.1f us per launch, %.2f us/frame\n", NB, jitter, mode, (mode & 2) ? "atomic " : "", (mode & 8) ? "(red) " : "",
                   (mode & 4) ? "store " : "", (mode & 1) ? "dependent" : "", best * 1000, best * 1000 / NB);
        }
    }
    cudaFree(cnt); cudaFree(slots); cudaFree(flush);
    return 0;
}
int main() { run<1>(); run<2>(); run<4>(); return 0; }
