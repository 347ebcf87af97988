// Microbenchmark: does the ACCESS SHAPE of the GEMM epilogue's fp32 read-modify-write (+ bf16 copy) limit its HBM
// throughput?  Same tile schedule and concurrency as gemm_tc_kernel's epilogue (148 CTAs x 8 warps, 128 x 256 tiles,
// warp = 32 rows x 128 columns), no MMA, three global access shapes:
//   A  "chunked": per instruction 4 rows x 128 B, the 4 column chunks of a warp one after the other (what the kernel did)
//   B  "rows":    per instruction 1 row x 512 B (whole 128-column slice of a row), rows one after the other
//   C  "rows2":   like B, but loads of 8 rows in flight before the first use
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/ubench_epilogue tools/ubench_epilogue.cu
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>

__device__ __forceinline__ uint32_t pack2(float a, float b) {
    __nv_bfloat162 v = __floats2bfloat162_rn(a, b);
    return *reinterpret_cast<uint32_t*>(&v);
}

// FLAGS (chunked shape only): 1 = also stream the A operand (bf16 [M, N], 16 B per lane, read and folded into the value),
//                             2 = write the bf16 copy, 4 = write per-row partial sums (8 B per row per warp slice)
template <int FLAGS>
__global__ void __launch_bounds__(256, 1) rmw2_kernel(float* __restrict__ x, __nv_bfloat16* __restrict__ o,
                                                      const uint4* __restrict__ A, float2* __restrict__ st, int M, int N) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int q = warp & 3, half = warp >> 2;
    const int ntn = N / 256, tiles = (M / 128) * ntn;
    const int rsub = lane >> 3, c8 = lane & 7;
    for (int tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
        const int mp = tile / ntn, nt = tile - mp * ntn;
        const int row0 = mp * 128 + q * 32, col0 = nt * 256 + half * 128;
        float extra = 0.f;
        if (FLAGS & 1) {  // this warp's share of the A tile: 32 rows x 256 B (half of the row per N tile, quarter per warp)
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const uint4 v = __ldg(A + ((size_t)(row0 + i * 4 + rsub) * N * 2 + (size_t)(nt * 2 + half) * 256) / 16 + c8 + 8 * (i & 1));
                extra += __uint_as_float(v.x & 0xffff0000u) * 1e-30f;
            }
        }
        float4 res[8];
        float s1[8], s2[8];
#pragma unroll
        for (int ps = 0; ps < 8; ++ps) {
            res[ps] = *reinterpret_cast<const float4*>(x + (size_t)(row0 + ps * 4 + rsub) * N + col0 + c8 * 4);
            s1[ps] = s2[ps] = 0.f;
        }
#pragma unroll 1
        for (int chunk = 0; chunk < 4; ++chunk) {
#pragma unroll
            for (int ps = 0; ps < 8; ++ps) {
                const size_t off = (size_t)(row0 + ps * 4 + rsub) * N + col0 + chunk * 32 + c8 * 4;
                float4 a = res[ps];
                a.x += 1.f + extra; a.y += 1.f; a.z += 1.f; a.w += 1.f;
                if (chunk < 3) res[ps] = *reinterpret_cast<const float4*>(x + off + 32);
                *reinterpret_cast<float4*>(x + off) = a;
                if (FLAGS & 2) *reinterpret_cast<uint2*>(o + off) = make_uint2(pack2(a.x, a.y), pack2(a.z, a.w));
                if (FLAGS & 4) {
                    s1[ps] += (a.x + a.y) + (a.z + a.w);
                    s2[ps] = fmaf(a.x, a.x, fmaf(a.y, a.y, fmaf(a.z, a.z, fmaf(a.w, a.w, s2[ps]))));
                }
            }
        }
        if (FLAGS & 4) {
            float w1 = 0.f, w2 = 0.f;
#pragma unroll
            for (int ps = 0; ps < 8; ++ps) {
#pragma unroll
                for (int o2 = 1; o2 < 8; o2 <<= 1) {
                    s1[ps] += __shfl_xor_sync(0xffffffffu, s1[ps], o2);
                    s2[ps] += __shfl_xor_sync(0xffffffffu, s2[ps], o2);
                }
                if (c8 == ps) { w1 = s1[ps]; w2 = s2[ps]; }
            }
            st[(size_t)(row0 + c8 * 4 + rsub) * (N / 128) + col0 / 128] = make_float2(w1, w2);
        }
    }
}

// 16 epilogue warps: warp = 32 rows x 64 columns in 4 chunks of 16 columns; per instruction 8 rows x 64 B
template <int FLAGS>
__global__ void __launch_bounds__(512, 1) rmw16_kernel(float* __restrict__ x, __nv_bfloat16* __restrict__ o,
                                                       const uint4* __restrict__ A, float2* __restrict__ st, int M, int N) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int q = warp & 3, quarter = warp >> 2;
    const int ntn = N / 256, tiles = (M / 128) * ntn;
    const int rsub = lane >> 2, c4 = lane & 3;
    for (int tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
        const int mp = tile / ntn, nt = tile - mp * ntn;
        const int row0 = mp * 128 + q * 32, col0 = nt * 256 + quarter * 64;
        float extra = 0.f;
        if (FLAGS & 1) {
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const uint4 v = __ldg(A + ((size_t)(row0 + i * 8 + rsub) * N * 2 + (size_t)(nt * 4 + quarter) * 128) / 16 + c4 + 4 * (i & 1));
                extra += __uint_as_float(v.x & 0xffff0000u) * 1e-30f;
            }
        }
        float4 res[4];
        float s1[4], s2[4];
#pragma unroll
        for (int ps = 0; ps < 4; ++ps) {
            res[ps] = *reinterpret_cast<const float4*>(x + (size_t)(row0 + ps * 8 + rsub) * N + col0 + c4 * 4);
            s1[ps] = s2[ps] = 0.f;
        }
#pragma unroll 1
        for (int chunk = 0; chunk < 4; ++chunk) {
#pragma unroll
            for (int ps = 0; ps < 4; ++ps) {
                const size_t off = (size_t)(row0 + ps * 8 + rsub) * N + col0 + chunk * 16 + c4 * 4;
                float4 a = res[ps];
                a.x += 1.f + extra; a.y += 1.f; a.z += 1.f; a.w += 1.f;
                if (chunk < 3) res[ps] = *reinterpret_cast<const float4*>(x + off + 16);
                *reinterpret_cast<float4*>(x + off) = a;
                if (FLAGS & 2) *reinterpret_cast<uint2*>(o + off) = make_uint2(pack2(a.x, a.y), pack2(a.z, a.w));
                if (FLAGS & 4) {
                    s1[ps] += (a.x + a.y) + (a.z + a.w);
                    s2[ps] = fmaf(a.x, a.x, fmaf(a.y, a.y, fmaf(a.z, a.z, fmaf(a.w, a.w, s2[ps]))));
                }
            }
        }
        if (FLAGS & 4) {
            float w1 = 0.f, w2 = 0.f;
#pragma unroll
            for (int ps = 0; ps < 4; ++ps) {
#pragma unroll
                for (int o2 = 1; o2 < 4; o2 <<= 1) {
                    s1[ps] += __shfl_xor_sync(0xffffffffu, s1[ps], o2);
                    s2[ps] += __shfl_xor_sync(0xffffffffu, s2[ps], o2);
                }
                if (c4 == ps) { w1 = s1[ps]; w2 = s2[ps]; }
            }
            st[(size_t)(row0 + c4 * 8 + rsub) * (N / 64) + col0 / 64] = make_float2(w1, w2);
        }
    }
}

template <int MODE>
__global__ void __launch_bounds__(256, 1) rmw_kernel(float* __restrict__ x, __nv_bfloat16* __restrict__ o, int M, int N) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int q = warp & 3, half = warp >> 2;
    const int ntn = N / 256, tiles = (M / 128) * ntn;
    for (int tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
        const int mp = tile / ntn, nt = tile - mp * ntn;
        const int row0 = mp * 128 + q * 32, col0 = nt * 256 + half * 128;
        if (MODE == 0) {
            const int rsub = lane >> 3, c8 = lane & 7;
            float4 res[8];
#pragma unroll
            for (int ps = 0; ps < 8; ++ps)
                res[ps] = *reinterpret_cast<const float4*>(x + (size_t)(row0 + ps * 4 + rsub) * N + col0 + c8 * 4);
#pragma unroll 1
            for (int chunk = 0; chunk < 4; ++chunk) {
#pragma unroll
                for (int ps = 0; ps < 8; ++ps) {
                    const size_t off = (size_t)(row0 + ps * 4 + rsub) * N + col0 + chunk * 32 + c8 * 4;
                    float4 a = res[ps];
                    a.x += 1.f; a.y += 1.f; a.z += 1.f; a.w += 1.f;
                    if (chunk < 3) res[ps] = *reinterpret_cast<const float4*>(x + off + 32);
                    *reinterpret_cast<float4*>(x + off) = a;
                    *reinterpret_cast<uint2*>(o + off) = make_uint2(pack2(a.x, a.y), pack2(a.z, a.w));
                }
            }
        } else {
            constexpr int PF = MODE == 1 ? 4 : 8;
            float4 res[PF];
#pragma unroll
            for (int i = 0; i < PF; ++i) res[i] = *reinterpret_cast<const float4*>(x + (size_t)(row0 + i) * N + col0 + lane * 4);
#pragma unroll 1
            for (int r = 0; r < 32; r += PF) {
#pragma unroll
                for (int i = 0; i < PF; ++i) {
                    const size_t off = (size_t)(row0 + r + i) * N + col0 + lane * 4;
                    float4 a = res[i];
                    a.x += 1.f; a.y += 1.f; a.z += 1.f; a.w += 1.f;
                    if (r + PF < 32) res[i] = *reinterpret_cast<const float4*>(x + off + (size_t)PF * N);
                    *reinterpret_cast<float4*>(x + off) = a;
                    *reinterpret_cast<uint2*>(o + off) = make_uint2(pack2(a.x, a.y), pack2(a.z, a.w));
                }
            }
        }
    }
}

int main(int argc, char** argv) {
    const int M = argc > 1 ? atoi(argv[1]) : 302080, N = argc > 2 ? atoi(argv[2]) : 512;
    float* x;
    __nv_bfloat16* o;
    cudaMalloc(&x, (size_t)M * N * 4);
    cudaMalloc(&o, (size_t)M * N * 2);
    cudaMemset(x, 0, (size_t)M * N * 4);
    cudaEvent_t a, b;
    cudaEventCreate(&a);
    cudaEventCreate(&b);
    const double bytes = (double)M * N * 10.0;
    for (int mode = 0; mode < 3; ++mode) {
        float best = 1e9f;
        for (int it = 0; it < 6; ++it) {
            cudaEventRecord(a);
            if (mode == 0) rmw_kernel<0><<<148, 256>>>(x, o, M, N);
            if (mode == 1) rmw_kernel<1><<<148, 256>>>(x, o, M, N);
            if (mode == 2) rmw_kernel<2><<<148, 256>>>(x, o, M, N);
            cudaEventRecord(b);
            cudaEventSynchronize(b);
            float ms;
            cudaEventElapsedTime(&ms, a, b);
            if (it > 0 && ms < best) best = ms;
        }
        printf("{\"mode\": %d, \"ms\": %.4f, \"GBs\": %.1f, \"err\": \"%s\"}\n", mode, best, bytes / best * 1e-6,
               cudaGetErrorString(cudaGetLastError()));
    }
    uint4* A;
    float2* st;
    cudaMalloc(&A, (size_t)M * N * 2);
    cudaMemset(A, 0, (size_t)M * N * 2);
    cudaMalloc(&st, (size_t)M * (N / 128) * 8);
    for (int flags = 0; flags < 8; ++flags) {
        float best = 1e9f;
        for (int it = 0; it < 6; ++it) {
            cudaEventRecord(a);
            switch (flags) {
                case 0: rmw2_kernel<0><<<148, 256>>>(x, o, A, st, M, N); break;
                case 1: rmw2_kernel<1><<<148, 256>>>(x, o, A, st, M, N); break;
                case 2: rmw2_kernel<2><<<148, 256>>>(x, o, A, st, M, N); break;
                case 3: rmw2_kernel<3><<<148, 256>>>(x, o, A, st, M, N); break;
                case 4: rmw2_kernel<4><<<148, 256>>>(x, o, A, st, M, N); break;
                case 5: rmw2_kernel<5><<<148, 256>>>(x, o, A, st, M, N); break;
                case 6: rmw2_kernel<6><<<148, 256>>>(x, o, A, st, M, N); break;
                case 7: rmw2_kernel<7><<<148, 256>>>(x, o, A, st, M, N); break;
            }
            cudaEventRecord(b);
            cudaEventSynchronize(b);
            float ms;
            cudaEventElapsedTime(&ms, a, b);
            if (it > 0 && ms < best) best = ms;
        }
        const double by = (double)M * N * (8.0 + ((flags & 1) ? 2.0 : 0.0) + ((flags & 2) ? 2.0 : 0.0));
        printf("{\"flags(1=A,2=bf16,4=stats)\": %d, \"ms\": %.4f, \"GBs\": %.1f, \"err\": \"%s\"}\n", flags, best, by / best * 1e-6,
               cudaGetErrorString(cudaGetLastError()));
    }
    float2* st16;
    cudaMalloc(&st16, (size_t)M * (N / 64) * 8);
    for (int flags = 3; flags < 8; flags += 4) {
        float best = 1e9f;
        for (int it = 0; it < 6; ++it) {
            cudaEventRecord(a);
            if (flags == 3) rmw16_kernel<3><<<148, 512>>>(x, o, A, st16, M, N);
            else rmw16_kernel<7><<<148, 512>>>(x, o, A, st16, M, N);
            cudaEventRecord(b);
            cudaEventSynchronize(b);
            float ms;
            cudaEventElapsedTime(&ms, a, b);
            if (it > 0 && ms < best) best = ms;
        }
        printf("{\"16 warps, flags\": %d, \"ms\": %.4f, \"GBs\": %.1f, \"err\": \"%s\"}\n", flags, best,
               (double)M * N * 12.0 / best * 1e-6, cudaGetErrorString(cudaGetLastError()));
    }
    return 0;
}
