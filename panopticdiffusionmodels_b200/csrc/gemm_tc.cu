// tcgen05 / TMEM / TMA bf16 GEMM with fused epilogue for sm_100a.
//
//   out = [gelu]( [A1 | A2] . W^T + bias ) [+ resid]      (libs/uvit_t2i.py:69,90,179; libs/timm.py:106-110)
//
// Persistent, warp-specialised, one CTA per SM:
//   warp 0      TMA producer   (A via 3-D map [K, rows, batch], W via 2-D map [K, N]; SWIZZLE_128B)
//   warp 1      TMEM allocator + single-thread tcgen05.mma issuer (UMMA 128 x 256 x 16, fp32 accumulate)
//   warps 2..9  epilogue: tcgen05.ld -> per-warp smem transpose -> coalesced 128-bit global I/O
// Pipelines: STAGES-deep smem ring (full/empty mbarriers) and a 2-deep TMEM accumulator ring
// (2 x 256 columns) so the epilogue of tile i overlaps the MMAs of tile i+1.
// The long-skip concat (uvit_t2i.py:179) is never materialised: the K loop streams A1 then A2 through
// two tensor maps.  Row views (two-stream zero-conv on mx[:, :334]) use the batch coordinate of the map.
#include <map>
#include <mutex>
#include <tuple>

#include "common.cuh"
#include "ptx.cuh"

namespace pdm {
namespace {

constexpr int BM = 128, BN = 256, BK = 64, STAGES = 3;
constexpr int EPI_WARPS = 8;
constexpr int THREADS = 64 + EPI_WARPS * 32;
constexpr int A_BYTES = BM * BK * 2;
constexpr int B_BYTES = BN * BK * 2;
constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
constexpr int SCR_STRIDE = 36;  // floats per scratch row (32 + 4 pad, keeps 16 B alignment, conflict-free)
constexpr int SCR_BYTES = 32 * SCR_STRIDE * 4;
constexpr int SMEM_BYTES = 1024 /*align slack*/ + STAGES * STAGE_BYTES + EPI_WARPS * SCR_BYTES + 256 /*barriers*/;
constexpr uint32_t TMEM_COLS = 512;

struct TcParams {
    int KB1, KB;  // k-blocks taken from A1, total k-blocks
    int N, Lr, tpb, ntn, total_tiles;
    const float* bias;
    const float* resid;
    long long resid_bs;
    float* out32;
    long long out32_bs;
    bf16* out2;
    long long out2_bs;
    int gelu;
};

// GELU(erf) for the bf16 path (libs/timm.py:101 -> nn.GELU()).  x.Phi(x) = 0.5 x (1 + tanh(x (a + b x^2 + c x^4)))
// with (a, b, c) fitted to the exact erf form: max abs deviation 3.1e-5 on [-8, 8] (the textbook 2-term tanh form is
// 4.7e-4); tanh.approx adds <= 2^-11 relative.  Both are far below the bf16 rounding of the stored activation.
// 7 FMA-pipe ops + 1 MUFU per element instead of erff's ~25: the fc1 epilogue stays under the MMA time.
__device__ __forceinline__ float gelu_erf(float x) {
    const float u = fminf(x * x, 64.f);
    const float w = x * fmaf(u, fmaf(u, -3.56580544e-04f, 3.70435562e-02f), 7.97452612e-01f);
    float t;
    asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(w));
    const float hx = 0.5f * x;
    return fmaf(hx, t, hx);
}

__global__ void __launch_bounds__(THREADS, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap tmA1, const __grid_constant__ CUtensorMap tmA2,
               const __grid_constant__ CUtensorMap tmB, const TcParams p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* scr_base = smem + STAGES * STAGE_BYTES;
    uint64_t* bars = reinterpret_cast<uint64_t*>(scr_base + EPI_WARPS * SCR_BYTES);
    uint64_t* full = bars;                     // [STAGES]
    uint64_t* empty = bars + STAGES;           // [STAGES]
    uint64_t* tfull = bars + 2 * STAGES;       // [2]
    uint64_t* tempty = bars + 2 * STAGES + 2;  // [2]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * STAGES + 4);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;

    if (warp == 0 && lane == 0) {
        ptx::prefetch_tmap(&tmA1);
        ptx::prefetch_tmap(&tmA2);
        ptx::prefetch_tmap(&tmB);
        for (int i = 0; i < STAGES; ++i) {
            ptx::mbar_init(&full[i], 1);
            ptx::mbar_init(&empty[i], 1);
        }
        for (int i = 0; i < 2; ++i) {
            ptx::mbar_init(&tfull[i], 1);
            ptx::mbar_init(&tempty[i], EPI_WARPS);
        }
        ptx::fence_mbar_init();
    }
    if (warp == 1) {
        ptx::tmem_alloc(tmem_slot, TMEM_COLS);
        ptx::tmem_relinquish();
    }
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ===================== TMA producer =====================
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
                const int mt = tile / p.ntn, nt = tile - mt * p.ntn;
                const int b = mt / p.tpb, t0 = (mt - b * p.tpb) * BM;
                for (int kb = 0; kb < p.KB; ++kb) {
                    ptx::mbar_wait(&empty[stage], phase ^ 1);
                    uint8_t* sa = smem + stage * STAGE_BYTES;
                    uint8_t* sb = sa + A_BYTES;
                    ptx::mbar_expect_tx(&full[stage], STAGE_BYTES);
                    if (kb < p.KB1)
                        ptx::tma_load_3d(&tmA1, &full[stage], sa, kb * BK, t0, b);
                    else
                        ptx::tma_load_3d(&tmA2, &full[stage], sa, (kb - p.KB1) * BK, t0, b);
                    ptx::tma_load_3d(&tmB, &full[stage], sb, kb * BK, nt * BN, 0);
                    if (++stage == STAGES) {
                        stage = 0;
                        phase ^= 1;
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer (one thread) =====================
        if (lane == 0) {
            constexpr uint32_t idesc = ptx::make_idesc_bf16(BM, BN);
            int stage = 0;
            uint32_t phase = 0;
            int it = 0;
            for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x, ++it) {
                const int as = it & 1;
                const uint32_t aphase = (it >> 1) & 1;
                ptx::mbar_wait(&tempty[as], aphase ^ 1);
                ptx::tc_fence_after();
                const uint32_t d_tmem = tmem_base + as * BN;
                for (int kb = 0; kb < p.KB; ++kb) {
                    ptx::mbar_wait(&full[stage], phase);
                    ptx::tc_fence_after();
                    const uint32_t sa = ptx::smem_u32(smem + stage * STAGE_BYTES);
                    const uint64_t adesc = ptx::make_smem_desc_sw128(sa, 1024);
                    const uint64_t bdesc = ptx::make_smem_desc_sw128(sa + A_BYTES, 1024);
#pragma unroll
                    for (int k = 0; k < BK / 16; ++k) {
                        // advance 16 bf16 = 32 bytes along K inside the 128-byte swizzle row
                        ptx::mma_bf16_ss(d_tmem, adesc + 2 * k, bdesc + 2 * k, idesc, (kb | k) != 0);
                    }
                    ptx::mma_commit(&empty[stage]);  // frees the smem slot when these MMAs retire
                    if (++stage == STAGES) {
                        stage = 0;
                        phase ^= 1;
                    }
                }
                ptx::mma_commit(&tfull[as]);  // accumulator ready for the epilogue
            }
        }
    } else {
        // ===================== epilogue =====================
        const int ew = warp - 2;
        const int q = warp & 3;   // TMEM lane quarter this warp may access
        const int half = ew >> 2; // which 128-column half of the tile
        float* scr = reinterpret_cast<float*>(scr_base + ew * SCR_BYTES);
        const int rsub = lane >> 3, c4 = lane & 7;
        int it = 0;
        for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x, ++it) {
            const int mt = tile / p.ntn, nt = tile - mt * p.ntn;
            const int b = mt / p.tpb, t0 = (mt - b * p.tpb) * BM;
            const int as = it & 1;
            const uint32_t aphase = (it >> 1) & 1;
            ptx::mbar_wait(&tfull[as], aphase);
            ptx::tc_fence_after();
            const int trow0 = t0 + q * 32;  // first token row handled by this warp
#pragma unroll 1
            for (int chunk = 0; chunk < 4; ++chunk) {
                const int col0 = half * 128 + chunk * 32;
                uint32_t v[32];
                ptx::tmem_ld_32x32(tmem_base + (uint32_t(q * 32) << 16) + as * BN + col0, v);
                ptx::tmem_ld_wait();
                if (chunk == 3) {
                    // all TMEM reads of this warp for this accumulator are done: hand it back to the MMA warp
                    ptx::tc_fence_before();
                    __syncwarp();
                    if (lane == 0) ptx::mbar_arrive(&tempty[as]);
                }
                const int col = nt * BN + col0 + c4 * 4;
                const bool col_ok = col < p.N;
                // prefetch the residual rows while the transpose goes through shared memory
                float4 res[8];
                if (p.resid) {
#pragma unroll
                    for (int ps = 0; ps < 8; ++ps) {
                        const int t = trow0 + ps * 4 + rsub;
                        res[ps] = make_float4(0.f, 0.f, 0.f, 0.f);
                        if (col_ok && t < p.Lr)
                            res[ps] = *reinterpret_cast<const float4*>(p.resid + ((long long)b * p.resid_bs + t) * p.N + col);
                    }
                }
                float* srow = scr + lane * SCR_STRIDE;
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    *reinterpret_cast<float4*>(srow + 4 * j) =
                        make_float4(__uint_as_float(v[4 * j]), __uint_as_float(v[4 * j + 1]),
                                    __uint_as_float(v[4 * j + 2]), __uint_as_float(v[4 * j + 3]));
                }
                __syncwarp();
                float4 bias4 = make_float4(0.f, 0.f, 0.f, 0.f);
                if (p.bias && col_ok) bias4 = __ldg(reinterpret_cast<const float4*>(p.bias + col));
#pragma unroll
                for (int ps = 0; ps < 8; ++ps) {
                    const int r = ps * 4 + rsub;
                    const int t = trow0 + r;
                    float4 a = *reinterpret_cast<const float4*>(scr + r * SCR_STRIDE + c4 * 4);
                    a.x += bias4.x;
                    a.y += bias4.y;
                    a.z += bias4.z;
                    a.w += bias4.w;
                    if (p.gelu) {
                        a.x = gelu_erf(a.x);
                        a.y = gelu_erf(a.y);
                        a.z = gelu_erf(a.z);
                        a.w = gelu_erf(a.w);
                    }
                    if (p.resid) {
                        a.x += res[ps].x;
                        a.y += res[ps].y;
                        a.z += res[ps].z;
                        a.w += res[ps].w;
                    }
                    if (col_ok && t < p.Lr) {
                        if (p.out32)
                            *reinterpret_cast<float4*>(p.out32 + ((long long)b * p.out32_bs + t) * p.N + col) = a;
                        if (p.out2) {
                            __nv_bfloat162 lo = __floats2bfloat162_rn(a.x, a.y), hi = __floats2bfloat162_rn(a.z, a.w);
                            uint2 pk;
                            pk.x = *reinterpret_cast<uint32_t*>(&lo);
                            pk.y = *reinterpret_cast<uint32_t*>(&hi);
                            *reinterpret_cast<uint2*>(p.out2 + ((long long)b * p.out2_bs + t) * p.N + col) = pk;
                        }
                    }
                }
                __syncwarp();
            }
        }
    }

    ptx::tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        ptx::tc_fence_after();
        ptx::tmem_dealloc(tmem_base, TMEM_COLS);
    }
}

// ------------------------------------------------------------------------------------------
// host side: tensor-map encoding through the driver entry point (no link dependency on libcuda)
// ------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode() {
    static EncodeTiledFn fn = nullptr;
    static std::once_flag once;
    std::call_once(once, [] {
        void* ptr = nullptr;
        cudaDriverEntryPointQueryResult qres;
        cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres);
        if (e == cudaSuccess && qres == cudaDriverEntryPointSuccess) fn = reinterpret_cast<EncodeTiledFn>(ptr);
    });
    PDM_REQUIRE(fn != nullptr, "cuTensorMapEncodeTiled not available from the CUDA driver");
    return fn;
}

typedef std::tuple<const void*, long long, long long, long long, long long, int, int> MapKey;
std::map<MapKey, CUtensorMap> g_map_cache;
std::mutex g_map_mutex;

}  // namespace

// bf16 [nbatch][rows][K] view with batch stride bs rows; box = [64 (K), box_rows, 1]; SWIZZLE_128B
CUtensorMap make_tmap_bf16_3d(const void* ptr, long long K, long long rows, long long nbatch, long long bs,
                              int box_rows, int box_k) {
    MapKey key(ptr, K, rows, nbatch, bs, box_rows, box_k);
    {
        std::lock_guard<std::mutex> lk(g_map_mutex);
        auto it = g_map_cache.find(key);
        if (it != g_map_cache.end()) return it->second;
    }
    PDM_REQUIRE((reinterpret_cast<uintptr_t>(ptr) & 15) == 0, "tensor map: base must be 16-byte aligned");
    PDM_REQUIRE((K * 2) % 16 == 0, "tensor map: row pitch must be a multiple of 16 bytes");
    CUtensorMap m;
    cuuint64_t dims[3] = {(cuuint64_t)K, (cuuint64_t)rows, (cuuint64_t)nbatch};
    cuuint64_t strides[2] = {(cuuint64_t)(K * 2), (cuuint64_t)(bs * K * 2)};
    cuuint32_t box[3] = {(cuuint32_t)box_k, (cuuint32_t)box_rows, 1};
    cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = get_encode()(&m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(ptr), dims, strides, box, estr,
                              CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                              CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    PDM_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled failed (code " + std::to_string((int)r) + ")");
    std::lock_guard<std::mutex> lk(g_map_mutex);
    g_map_cache[key] = m;
    return m;
}

void clear_tmap_cache() {
    std::lock_guard<std::mutex> lk(g_map_mutex);
    g_map_cache.clear();
}

static int g_num_sms = 0;
static int num_sms() {
    if (g_num_sms == 0) {
        int dev = 0;
        PDM_CHECK_CUDA(cudaGetDevice(&dev));
        PDM_CHECK_CUDA(cudaDeviceGetAttribute(&g_num_sms, cudaDevAttrMultiProcessorCount, dev));
    }
    return g_num_sms;
}

void gemm_tc_bf16(const GemmProblem& g, cudaStream_t s) {
    PDM_REQUIRE(g.A1 && g.W16 && g.Lr > 0 && g.nb > 0, "gemm_tc: bad problem");
    PDM_REQUIRE(g.N % 4 == 0, "gemm_tc: N must be a multiple of 4");
    PDM_REQUIRE(g.K1 % 8 == 0 && (!g.A2 || (g.K1 % BK == 0 && g.K2 % 8 == 0)), "gemm_tc: K alignment");
    const int K2 = g.A2 ? g.K2 : 0;
    const int K = g.K1 + K2;
    TcParams p;
    p.KB1 = ceil_div(g.K1, BK);
    p.KB = p.KB1 + ceil_div(K2, BK);
    p.N = g.N;
    p.Lr = g.Lr;
    p.tpb = ceil_div(g.Lr, BM);
    p.ntn = ceil_div(g.N, BN);
    p.total_tiles = g.nb * p.tpb * p.ntn;
    p.bias = g.bias;
    p.resid = g.resid;
    p.resid_bs = g.resid_bs ? g.resid_bs : g.Lr;
    p.out32 = g.out32;
    p.out32_bs = g.out32_bs ? g.out32_bs : g.Lr;
    p.out2 = (bf16*)g.out2;
    p.out2_bs = g.out2_bs ? g.out2_bs : g.Lr;
    p.gelu = g.gelu ? 1 : 0;
    const CUtensorMap tmA1 = make_tmap_bf16_3d(g.A1, g.K1, g.Lr, g.nb, g.a1_bs ? g.a1_bs : g.Lr, BM, BK);
    const CUtensorMap tmA2 =
        g.A2 ? make_tmap_bf16_3d(g.A2, g.K2, g.Lr, g.nb, g.a2_bs ? g.a2_bs : g.Lr, BM, BK) : tmA1;
    const CUtensorMap tmB = make_tmap_bf16_3d(g.W16, K, g.N, 1, g.N, BN, BK);
    static bool attr_set = false;
    if (!attr_set) {
        PDM_CHECK_CUDA(cudaFuncSetAttribute(gemm_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
        attr_set = true;
    }
    const int grid = std::min(p.total_tiles, num_sms());
    gemm_tc_kernel<<<grid, THREADS, SMEM_BYTES, s>>>(tmA1, tmA2, tmB, p);
    check_launch("gemm_tc");
}

}  // namespace pdm
