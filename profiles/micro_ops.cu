// microbenchmark: per-SM throughput of the instructions the render kernels are made of (sm_100a)
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o micro_ops micro_ops.cu && ./micro_ops
// Every test: 148 x C CTAs of 256 threads, each thread runs ITERS iterations of U independent instances of the operation;
// reported: SM cycles per warp-instruction (clock64 span of the slowest CTA x CTAs per SM / warp-instructions per SM).
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); return 1; } } while (0)

__device__ __forceinline__ uint32_t hash(uint32_t x) { x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16; return x; }

enum Op { ATOMS_DISTINCT, ATOMS_RANDOM, ATOMS_SAME, REDS_RANDOM, ATOMS64_RANDOM, RMW32_RANDOM, RMW128_STRIDE5, LDS64_RANDOM, STS16_RANDOM, MATCH_ANY, VOTE6,
          IDP2A, IMAD, IMADHI, IMADWIDE, PRMT, LOP3, IADD3, DADD, DMUL, DFMA, I2F, F2I, MUFU_RCP, SHFL, FFMA, LDS32_SEQ, NOPS };
const char *names[] = {"ATOMS u32 ret, 32 distinct banks", "ATOMS u32 ret, random (1089 homes)", "ATOMS u32 ret, one address", "RED.shared u32 random", "ATOMS u64 ret random",
                       "LDS32+IADD+STS32 random", "LDS128+LDS32+5 IMAD+STS128+STS32, stride 20 B", "LDS.64 random", "STS.U16 random", "match_any", "6 x VOTE+LOP3 (match emulation)",
                       "IDP.2A", "IMAD", "IMAD.HI", "IMAD.WIDE", "PRMT", "LOP3", "IADD3", "DADD", "DMUL", "DFMA", "I2F", "F2I", "MUFU.RCP", "SHFL.IDX", "FFMA", "LDS.32 sequential"};

template <int OP, int U>
__global__ void __launch_bounds__(256) k(uint32_t iters, uint32_t *out, long long *cyc) {
    __shared__ __align__(16) uint32_t sm[8192];
    const uint32_t tid = threadIdx.x, lane = tid & 31u;
    for (uint32_t i = tid; i < 8192u; i += 256u) sm[i] = i;
    __syncthreads();
    uint32_t a[U], acc = 0;
    double d[U];
    float f[U];
#pragma unroll
    for (int u = 0; u < U; ++u) { a[u] = hash(tid * 16u + u + blockIdx.x * 4096u); d[u] = 1.0 + a[u] * 1e-9; f[u] = 1.0f + (a[u] & 1023u); }
    const long long t0 = clock64();
    for (uint32_t it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < U; ++u) {
            if (OP == ATOMS_DISTINCT) { acc += atomicAdd(&sm[((lane * 7u + it + u) & 31u) + 32u * ((a[u] >> 8) & 31u)], 1u); }
            else if (OP == ATOMS_RANDOM) { a[u] = a[u] * 1664525u + 1013904223u; acc += atomicAdd(&sm[a[u] >> 22], 1u); }
            else if (OP == ATOMS_SAME) { acc += atomicAdd(&sm[u], 1u); }
            else if (OP == REDS_RANDOM) { a[u] = a[u] * 1664525u + 1013904223u; atomicAdd(&sm[a[u] >> 22], 1u); }
            else if (OP == ATOMS64_RANDOM) { a[u] = a[u] * 1664525u + 1013904223u; acc += (uint32_t) atomicAdd((unsigned long long *) &sm[2u * (a[u] >> 22)], 1ull); }
            else if (OP == RMW32_RANDOM) { a[u] = a[u] * 1664525u + 1013904223u; uint32_t *p = &sm[a[u] >> 22]; *(volatile uint32_t *) p = *(volatile uint32_t *) p + 1u; }
            else if (OP == RMW128_STRIDE5) {
                a[u] = a[u] * 1664525u + 1013904223u;
                uint32_t *p = &sm[((a[u] >> 27) * 32u + lane) * 5u];     // consecutive pixels per warp, 20-byte accumulators
                uint32_t v0 = ((volatile uint32_t *) p)[0], v1 = ((volatile uint32_t *) p)[1], v2 = ((volatile uint32_t *) p)[2], v3 = ((volatile uint32_t *) p)[3], v4 = ((volatile uint32_t *) p)[4];
                const uint32_t n = a[u] & 0xffffu;
                v0 += (a[u] & 255u) * n; v1 += ((a[u] >> 8) & 255u) * n; v2 += ((a[u] >> 16) & 255u) * n; v3 += (a[u] >> 24) * n; v4 += n;
                ((volatile uint32_t *) p)[0] = v0; ((volatile uint32_t *) p)[1] = v1; ((volatile uint32_t *) p)[2] = v2; ((volatile uint32_t *) p)[3] = v3; ((volatile uint32_t *) p)[4] = v4;
            }
            else if (OP == LDS64_RANDOM) { a[u] = a[u] * 1664525u + 1013904223u; unsigned long long v = *(volatile unsigned long long *) &sm[2u * (a[u] >> 21)]; acc += (uint32_t) v ^ (uint32_t) (v >> 32); }
            else if (OP == STS16_RANDOM) { a[u] = a[u] * 1664525u + 1013904223u; ((volatile uint16_t *) sm)[a[u] >> 20] = (uint16_t) a[u]; }
            else if (OP == MATCH_ANY) { a[u] = a[u] * 1664525u + 1013904223u; acc += __match_any_sync(0xffffffffu, a[u] >> 27); }
            else if (OP == VOTE6) {
                a[u] = a[u] * 1664525u + 1013904223u;
                uint32_t m = 0xffffffffu, key = a[u] >> 26;
#pragma unroll
                for (int b = 0; b < 6; ++b) { const uint32_t v = __ballot_sync(0xffffffffu, (key >> b) & 1u); m &= ((key >> b) & 1u) ? v : ~v; }
                acc += m;
            }
            else if (OP == IDP2A) { a[u] = __dp2a_lo(a[u], 0x01020304u + it, a[u]); }
            else if (OP == IMAD) { a[u] = a[u] * 1664525u + it; }
            else if (OP == IMADHI) { a[u] = __umulhi(a[u], 0x9e3779b9u + it) + 12345u; }
            else if (OP == IMADWIDE) { unsigned long long w = (unsigned long long) a[u] * (it | 1u) + a[u]; a[u] = (uint32_t) (w >> 32) ^ (uint32_t) w; }
            else if (OP == PRMT) { a[u] = __byte_perm(a[u], it, 0x2103); }
            else if (OP == LOP3) { a[u] = (a[u] & it) ^ (a[u] >> 3 | 0x55u); }
            else if (OP == IADD3) { a[u] = a[u] + it + 77u; }
            else if (OP == DADD) { d[u] = d[u] + 1.000000001; }
            else if (OP == DMUL) { d[u] = d[u] * 1.000000001; }
            else if (OP == DFMA) { d[u] = fma(d[u], 1.000000001, 0.5); }
            else if (OP == I2F) { f[u] += __uint2float_rz(a[u] + it); }
            else if (OP == F2I) { a[u] += (uint32_t) __float2uint_rz(f[u] + (float) u) ; f[u] += 1.0f; }
            else if (OP == MUFU_RCP) { f[u] = __frcp_rn(f[u]) + 1.5f; }
            else if (OP == SHFL) { a[u] = __shfl_sync(0xffffffffu, a[u], (int) ((lane + 1u + u) & 31u)); }
            else if (OP == FFMA) { f[u] = f[u] * 1.0001f + 0.5f; }
            else if (OP == LDS32_SEQ) { acc += ((volatile uint32_t *) sm)[(tid + 256u * ((it + u) & 15u))]; }
        }
    }
    const long long t1 = clock64();
#pragma unroll
    for (int u = 0; u < U; ++u) acc += a[u] + (uint32_t) d[u] + (uint32_t) f[u];
    if (acc == 0x12345678u) out[0] = acc;
    if (tid == 0) cyc[blockIdx.x] = t1 - t0;
}

template <int OP>
int run(int ctas_per_sm) {
    const int U = 4;
    const uint32_t iters = 2000;
    uint32_t *out; long long *cyc;
    const int grid = 148 * ctas_per_sm;
    CK(cudaMalloc(&out, 4)); CK(cudaMalloc(&cyc, grid * 8));
    for (int rep = 0; rep < 2; ++rep) { k<OP, U><<<grid, 256>>>(iters, out, cyc); CK(cudaDeviceSynchronize()); }
    static long long h[148 * 8];
    CK(cudaMemcpy(h, cyc, grid * 8, cudaMemcpyDeviceToHost));
    long long mx = 0; for (int i = 0; i < grid; ++i) if (h[i] > mx) mx = h[i];
    const double warp_instr_per_sm = (double) iters * U * 8 * ctas_per_sm;
    printf("%-52s %d CTA/SM: %8.3f SM-cycles per warp-op  (%.2f lane-ops per cycle per SM)\n", names[OP], ctas_per_sm, mx / warp_instr_per_sm, 32.0 * warp_instr_per_sm / mx);
    cudaFree(out); cudaFree(cyc);
    return 0;
}

int main() {
    cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
    printf("%s, %d SMs\n", p.name, p.multiProcessorCount);
    for (int c : {1, 4}) {
        run<ATOMS_DISTINCT>(c); run<ATOMS_RANDOM>(c); run<ATOMS_SAME>(c); run<REDS_RANDOM>(c); run<ATOMS64_RANDOM>(c); run<RMW32_RANDOM>(c); run<RMW128_STRIDE5>(c);
        run<LDS64_RANDOM>(c); run<STS16_RANDOM>(c); run<LDS32_SEQ>(c); run<MATCH_ANY>(c); run<VOTE6>(c); run<SHFL>(c);
        run<IDP2A>(c); run<IMAD>(c); run<IMADHI>(c); run<IMADWIDE>(c); run<PRMT>(c); run<LOP3>(c); run<IADD3>(c); run<FFMA>(c);
        run<DADD>(c); run<DMUL>(c); run<DFMA>(c); run<I2F>(c); run<F2I>(c); run<MUFU_RCP>(c);
    }
    return 0;
}
