// tools/ubench.cu -- integer-pipe microbenchmarks for sm_100a (not part of the product).
// Measures warp-instruction throughput per SM per clock of the operations the bit packer / unpacker is
// built from, so that the kernels can be balanced between the ALU pipe (LOP3 / SHF / IADD3 / SEL / ISETP)
// and the FMA pipe (IMAD / IMAD.WIDE).  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ubench ubench.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

typedef unsigned int u32;
typedef unsigned long long u64;

constexpr int ITERS = 4096;

template <int MODE>
__global__ void __launch_bounds__(256) k(u32* out, u32 seed)
{
    u32 a[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) a[i] = seed + threadIdx.x * 8 + i;
    u32 nb = seed & 31, m = seed | 1;
#pragma unroll 1
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            if (MODE == 0) a[i] = a[i] * m + nb;                                   // IMAD
            if (MODE == 1) { u64 x; asm("mad.wide.u32 %0, %1, %2, %3;" : "=l"(x) : "r"(a[i]), "r"(m), "l"((u64)nb)); a[i] = (u32)x ^ (u32)(x >> 32); }   // IMAD.WIDE + LOP
            if (MODE == 2) a[i] = (a[i] ^ m) | nb;                                 // LOP3
            if (MODE == 3) a[i] = __funnelshift_l(a[i], m, a[i]);                  // SHF
            if (MODE == 4) a[i] = a[i] > m ? a[i] - nb : a[i] + m;                 // ISETP + SEL/IADD
            if (MODE == 5) { a[i] = (a[i] ^ m) | nb; a[(i + 1) & 7] = a[(i + 1) & 7] * m + nb; }   // LOP3 + IMAD mix
            if (MODE == 6) { u64 x; asm("mad.wide.u32 %0, %1, %2, %3;" : "=l"(x) : "r"(a[i]), "r"(m), "l"((u64)nb)); a[i] = (u32)(x >> 32) + (u32)x; }   // IMAD.WIDE + IADD
            if (MODE == 7) a[i] = __shfl_xor_sync(0xffffffffu, a[i], 1) + 1;       // SHFL
            if (MODE == 8) a[i] = (u32)__clz((int)a[i]) + a[i];                    // FLO
            if (MODE == 9) a[i] = __popc(a[i]) + a[i];                             // POPC
        }
    }
    u32 r = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) r ^= a[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}

// put32 variants: append 12 fields of n bits to a per-thread bit stream in shared memory
template <int MODE>
__global__ void __launch_bounds__(256) kput(u32* out, u32 seed, u32 n)
{
    extern __shared__ u32 sm[];
    u32* my = sm + threadIdx.x;                      // column layout: word k at sm[k * 256 + tid]
    u32 lo = 0, nb = threadIdx.x & 31, wp = 0, v = seed + threadIdx.x;
    const u32 mask = (1u << n) - 1;
#pragma unroll 1
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < 12; ++i) {
            const u32 f = (v >> i) & mask;
            if (MODE == 0) {                         // IMAD.WIDE form
                const u32 pw = 1u << nb;
                u64 x;
                asm("mul.wide.u32 %0, %1, %2;" : "=l"(x) : "r"(f), "r"(pw));
                x |= lo;
                const u32 t = nb + n;
                const bool p = t >= 32;
                if (p) my[(wp & 15) * 256] = (u32)x;
                lo = p ? (u32)(x >> 32) : (u32)x;
                wp += p ? 1 : 0;
                nb = t & 31;
            } else {                                 // shift / or form
                const u32 a0 = lo | (f << nb);
                const u32 a1 = __funnelshift_l(f, 0u, nb);
                const u32 t = nb + n;
                const bool p = t >= 32;
                if (p) my[(wp & 15) * 256] = a0;
                lo = p ? a1 : a0;
                wp += p ? 1 : 0;
                nb = t & 31;
            }
        }
        v = v * 1664525u + 1013904223u;
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = lo + wp + my[0];
}

template <typename F>
float time_ms(F f)
{
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    f();
    cudaDeviceSynchronize();
    cudaEventRecord(e0);
    f();
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    return ms;
}

int main()
{
    cudaDeviceProp prop;
    cudaGetDeviceProperties(&prop, 0);
    const int sms = prop.multiProcessorCount;
    int clk_khz = 0;
    cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0);
    u32* out;
    cudaMalloc(&out, sizeof(u32) * sms * 8 * 256);
    const int grid = sms * 8;                        // 8 CTAs x 8 warps = 64 warps per SM
    const char* names[] = {"IMAD", "IMAD.WIDE+LOP", "LOP3", "SHF", "ISETP+SEL", "LOP3+IMAD mix", "IMAD.WIDE+IADD", "SHFL", "FLO+IADD", "POPC+IADD"};
    const double per_iter[] = {8, 16, 8, 8, 24, 16, 16, 16, 16, 16};   // rough instruction counts per inner pass (see SASS)
    printf("device %s, %d SMs, clock %d MHz (nominal)\n", prop.name, sms, clk_khz / 1000);
#define RUN(M) { float ms = time_ms([&] { k<M><<<grid, 256>>>(out, 12345u); }); \
        double ops = (double)grid * 8 /*warps*/ * ITERS * 8; \
        printf("mode %d %-16s %8.3f ms  -> %6.2f source-ops (warp) per SM per clk @1.9GHz\n", M, names[M], ms, ops / sms / (ms * 1e-3 * 1.9e9)); (void)per_iter; }
    RUN(0) RUN(1) RUN(2) RUN(3) RUN(4) RUN(5) RUN(6) RUN(7) RUN(8) RUN(9)
    for (u32 n : {6u, 12u, 24u}) {
        float a = time_ms([&] { kput<0><<<grid, 256, 16 * 256 * 4>>>(out, 777u, n); });
        float b = time_ms([&] { kput<1><<<grid, 256, 16 * 256 * 4>>>(out, 777u, n); });
        double puts = (double)grid * 8 * ITERS * 12;
        printf("put32 n=%2u: IMAD.WIDE form %7.3f ms (%5.2f clk/put/SM), shift form %7.3f ms (%5.2f clk/put/SM)\n", n, a,
               a * 1e-3 * 1.9e9 * sms / puts, b, b * 1e-3 * 1.9e9 * sms / puts);
    }
    return 0;
}
