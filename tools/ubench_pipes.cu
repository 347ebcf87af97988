// Pipe-rate micro-benchmarks that size the attention softmax phase on sm_100a (development tool, not product code).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/ubench_pipes tools/ubench_pipes.cu && tools/ubench_pipes
// Each test: 1 CTA per SM, W warps, every warp runs ITERS iterations of an unrolled body; prints SM cycles per
// warp-instruction group so that rates read as "cycles per warp-wide op per SMSP".
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); return 1; } } while (0)

__device__ __forceinline__ float ex2(float x) { float y; asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ uint64_t fma2(uint64_t a, uint64_t b, uint64_t c) { uint64_t r; asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r; }
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
          "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
          "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t (&r)[16]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
        "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
        ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]),
        "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]) : "memory");
}

// mode 0: LDTM x32 back to back (wait after each)   1: LDTM x32, 4 in flight then wait
// mode 2: MUFU.EX2 x32 independent                  3: FFMA2 x32 independent
// mode 4: F2FP pack x32                             5: max3 x32
// mode 6: STTM x16 back to back                     7: FFMA (scalar, 3-reg) x32
// mode 8: mixed softmax-like chunk (LDTM x32 + 16 FFMA2 + 32 MUFU + 16 FADD2 + 16 F2FP + STTM x16)
// mode 9: mode 8 without MUFU (poly-3 on FMA pipe instead)
template <int MODE>
__global__ void __launch_bounds__(512, 1) k(int iters, long long* out_cycles, float* sink) {
    __shared__ uint32_t slot;
    const int warp = threadIdx.x >> 5;
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&slot)), "r"(512) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tb = slot + (uint32_t((warp & 3) * 32) << 16) + (warp >> 2) * 128;
    float acc[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) acc[i] = threadIdx.x * 1e-3f + i;
    uint32_t v[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = i;
    __syncthreads();
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
        if (MODE == 0) {
            tmem_ld32(tb + (it & 3) * 32, v);
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
            for (int i = 0; i < 32; i += 8) acc[i] += __uint_as_float(v[i]);
        } else if (MODE == 1) {
            uint32_t a[32], b[32], c[32], d[32];
            tmem_ld32(tb, a); tmem_ld32(tb + 32, b); tmem_ld32(tb + 64, c); tmem_ld32(tb + 96, d);
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
            for (int i = 0; i < 32; i += 8) acc[i] += __uint_as_float(a[i] ^ b[i] ^ c[i] ^ d[i]);
        } else if (MODE == 2) {
#pragma unroll
            for (int i = 0; i < 32; ++i) acc[i] = ex2(acc[i]);
        } else if (MODE == 3) {
#pragma unroll
            for (int i = 0; i < 32; i += 2) {
                uint64_t x; asm volatile("mov.b64 %0, {%1, %2};" : "=l"(x) : "f"(acc[i]), "f"(acc[i + 1]));
                x = fma2(x, x, x); x = fma2(x, x, x);
                asm volatile("mov.b64 {%0, %1}, %2;" : "=f"(acc[i]), "=f"(acc[i + 1]) : "l"(x));
            }
        } else if (MODE == 4) {
#pragma unroll
            for (int i = 0; i < 32; ++i) {
                uint32_t p; asm volatile("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(p) : "f"(acc[i]), "f"(acc[(i + 1) & 31]));
                v[i] ^= p;
            }
        } else if (MODE == 5) {
#pragma unroll
            for (int i = 0; i < 32; ++i) asm volatile("max.f32 %0, %0, %1, %2;" : "+f"(acc[i]) : "f"(acc[(i + 1) & 31]), "f"(acc[(i + 2) & 31]));
        } else if (MODE == 6) {
            uint32_t p[16];
#pragma unroll
            for (int i = 0; i < 16; ++i) p[i] = v[i] + it;
            tmem_st16(tb + (it & 7) * 16, p);
            asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
        } else if (MODE == 7) {
#pragma unroll
            for (int i = 0; i < 32; ++i) acc[i] = fmaf(acc[i], acc[(i + 1) & 31], acc[(i + 2) & 31]);
        } else if (MODE == 8 || MODE == 9) {
            uint32_t s[32], p[16];
            tmem_ld32(tb + (it & 1) * 32, s);
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            float l0 = 0.f, l1 = 0.f;
#pragma unroll
            for (int i = 0; i < 16; ++i) {
                float a0 = fmaf(__uint_as_float(s[2 * i]), 0.18f, -acc[0]);
                float a1 = fmaf(__uint_as_float(s[2 * i + 1]), 0.18f, -acc[0]);
                float p0, p1;
                if (MODE == 8) { p0 = ex2(a0); p1 = ex2(a1); }
                else {
                    // Cody-Waite + degree-3 polynomial, scalar form (the packed form is what the kernel would use)
                    float t0f = a0 + 12582912.f, t1f = a1 + 12582912.f;
                    float f0 = a0 - (t0f - 12582912.f), f1 = a1 - (t1f - 12582912.f);
                    float q0 = fmaf(fmaf(fmaf(0.0555f, f0, 0.2402f), f0, 0.6931f), f0, 1.f);
                    float q1 = fmaf(fmaf(fmaf(0.0555f, f1, 0.2402f), f1, 0.6931f), f1, 1.f);
                    p0 = __uint_as_float(__float_as_uint(q0) + (__float_as_uint(t0f) << 23));
                    p1 = __uint_as_float(__float_as_uint(q1) + (__float_as_uint(t1f) << 23));
                }
                l0 += p0; l1 += p1;
                asm volatile("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(p[i]) : "f"(p1), "f"(p0));
            }
            acc[1] += l0 + l1;
            tmem_st16(tb + 64 + (it & 1) * 16, p);
        }
    }
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    long long t1 = clock64();
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 32; ++i) s += acc[i] + __uint_as_float(v[i]);
    if (s == 123.456f) sink[0] = s;
    if ((threadIdx.x & 31) == 0) out_cycles[blockIdx.x * 16 + warp] = t1 - t0;
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(slot), "r"(512) : "memory");
}

template <int MODE>
int run(const char* name, int warps, double ops_per_iter) {
    const int iters = 2000, nsm = 148;
    long long* d; float* sink;
    CK(cudaMalloc(&d, nsm * 16 * sizeof(long long)));
    CK(cudaMalloc(&sink, 4));
    k<MODE><<<nsm, warps * 32>>>(100, d, sink);
    k<MODE><<<nsm, warps * 32>>>(iters, d, sink);
    CK(cudaDeviceSynchronize());
    static long long h[148 * 16];
    CK(cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost));
    long long mx = 0;
    for (int b = 0; b < nsm; ++b) for (int w = 0; w < warps; ++w) if (h[b * 16 + w] > mx) mx = h[b * 16 + w];
    double cyc_per_iter = double(mx) / iters;
    // per SMSP: warps/4 warps share one scheduler
    printf("%-34s warps=%2d  cyc/iter(warp)=%8.1f  cyc per warp-op per SMSP=%6.2f\n", name, warps, cyc_per_iter,
           cyc_per_iter / (ops_per_iter * (warps / 4.0)));
    cudaFree(d); cudaFree(sink);
    return 0;
}

int main() {
    for (int w : {4, 8, 16}) {
        run<0>("LDTM.x32 serial (4KB/warp-op)", w, 1);
        run<1>("LDTM.x32 4-deep", w, 4);
        run<6>("STTM.x16 serial (2KB/warp-op)", w, 1);
        run<2>("MUFU.EX2", w, 32);
        run<3>("FFMA2", w, 32);
        run<7>("FFMA", w, 32);
        run<4>("F2FP.BF16 pack", w, 32);
        run<5>("FMNMX3", w, 32);
        run<8>("softmax chunk (32 elem) MUFU", w, 1);
        run<9>("softmax chunk (32 elem) poly3", w, 1);
    }
    return 0;
}
