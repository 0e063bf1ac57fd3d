// dfb_gl.cu -- GroupedLinearEinsum (DeepFilterNet/df/modules.py:741-780) on the 5th-generation tensor cores.
//
//   Y[m, g*Hg + n] = act( sum_i X[m, g*Ig + i] * W[g][i][n] ) * oscale + ooffset + R[m, g*Hg + n]
//
// at fp32-level accuracy (BF16x3: both operands as BF16 hi/lo planes, products hi*hi + lo*hi + hi*lo, fp32
// accumulation in TMEM).  The block-diagonal weight is NOT expanded to a dense matrix: every group is its own
// small MMA (M = 128 rows of X, N = Hg padded to a multiple of 16, K = Ig in steps of 16), so a K step costs the
// issue time of an N = 16..128 instruction instead of a dense 128 x 128 one.
//
// Persistent CTA = (group slice, row-tile group).  Its weight slice (gpc groups, <= 100 KB as BF16 hi | lo) is fetched
// once with two bulk copies and stays in shared memory as the B operand (K-major, no swizzle, 8 x 16 B core matrices --
// the image is laid out on the host, weights.py gl_bx_image).  The X planes [M][K] (written by the producing kernel's
// epilogue) stream through a TMA ring of [128 rows x 64 k] hi + lo boxes (128-byte swizzle) as the A operand.  A K step
// of 16 never straddles a box or a group (Ig % 16 == 0, slice start % 64 == 0).  Accumulators: gpc * Hgp <= 256 TMEM
// columns, double buffered, so the epilogue of row tile i overlaps the MMAs of tile i + 1.
//   warps 0-7 : epilogue -- tcgen05.ld (lane = row, 16 columns at a time) -> per-warp shared staging tile -> transposed
//               read (4 lanes cover 64 contiguous bytes of one row) -> activation / scale / residual -> fp32 and/or
//               BF16 hi/lo plane stores in full 32-byte sectors
//   warp 8    : TMA producer        warp 9 : MMA issuer (whole warp, elect.sync issue)
#include <cuda.h>
#include <cuda_bf16.h>

#include <map>
#include <mutex>
#include <tuple>

#include "dfb_common.cuh"
#include "dfb_ptx.cuh"

namespace dfb {

constexpr int kGxThreads = 320, kGxMaxStages = 6, kGxBoxBytes = 128 * 128 /* 128 rows x 64 bf16 */;
constexpr int kGxStageRow = 80;                       // bytes per staged row: 16 columns fp32 + 16 B pad
constexpr int kGxStageBytes = 32 * kGxStageRow;       // per epilogue warp

struct GlBxParams {
    const unsigned short *w_img;  // [hi | lo][G][Ig/8][Hgp/8][8][8] BF16
    const float *res; int64_t ldr;
    float *y; int64_t ldy;
    unsigned short *y_hi, *y_lo; int64_t ldp;
    int M, G, Ig, Hg, Hgp, gpc, act, stages;
    float oscale, ooffset;
};

__device__ __forceinline__ float gx_act(float x, int act) {
    if (act == 1) return fmaxf(x, 0.f);
    if (act == 2) return gt_tanh(x);   // 1 - 2 / (1 + e^2x) on the MUFU units (~1e-7 absolute, as in the GRU gates); tanhf made df_out epilogue bound
    return x;
}

__global__ void __launch_bounds__(kGxThreads, 1)
k_gl_bx(const __grid_constant__ CUtensorMap tmXhi, const __grid_constant__ CUtensorMap tmXlo, GlBxParams p) {
    extern __shared__ __align__(1024) unsigned char gx_smem_raw[];
    const uint32_t sb = (smem_u32(gx_smem_raw) + 1023u) & ~1023u;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int slice = blockIdx.x, g0 = slice * p.gpc;
    const int Dc = p.gpc * p.Hgp;                          // accumulator columns of one buffer
    const int nboxes = p.gpc * p.Ig / 64;                  // 64-wide K boxes per row tile
    const int ntiles = (p.M + 127) / 128;
    const uint32_t gbytes = (uint32_t)p.Ig * p.Hgp * 2;    // one group's block in one plane
    const uint32_t wplane = gbytes * p.gpc;                // the slice in one plane
    // shared memory map
    const uint32_t s_x = sb;                                             // [stages][hi | lo][16 KB]
    const uint32_t s_w = s_x + (uint32_t)p.stages * 2 * kGxBoxBytes;     // W hi | lo
    const uint32_t s_stage = (s_w + 2 * wplane + 127u) & ~127u;          // 8 staging tiles
    const uint32_t s_bar = s_stage + 8 * kGxStageBytes;                  // full[6] empty[6] tfull[2] tempty[2] wbar
    const uint32_t b_full = s_bar, b_empty = s_bar + 8 * kGxMaxStages, b_tfull = b_empty + 8 * kGxMaxStages,
                   b_tempty = b_tfull + 16, b_w = b_tempty + 16, s_tmem = b_w + 8;
    if (threadIdx.x == 0) {
        for (int s = 0; s < p.stages; s++) { mbar_init_a(b_full + 8 * s, 1); mbar_init_a(b_empty + 8 * s, 1); }
        for (int i = 0; i < 2; i++) { mbar_init_a(b_tfull + 8 * i, 1); mbar_init_a(b_tempty + 8 * i, 8); }
        mbar_init_a(b_w, 1);
        fence_barrier_init();
        // the weight slice: two bulk copies (hi plane, lo plane)
        mbar_expect_tx_a(b_w, 2 * wplane);
        const unsigned char *src = reinterpret_cast<const unsigned char *>(p.w_img);
        bulk_load(s_w, src + (size_t)g0 * gbytes, wplane, b_w);
        bulk_load(s_w + wplane, src + (size_t)p.G * gbytes + (size_t)g0 * gbytes, wplane, b_w);
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(s_tmem), "r"(512) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = lds32(s_tmem);
    if (warp == 8) {
        // ===== TMA producer
        if (lane == 0) {
            tma_prefetch_desc(&tmXhi); tma_prefetch_desc(&tmXlo);
            const int col0 = g0 * p.Ig;
            int it = 0;
            for (int tile = blockIdx.y; tile < ntiles; tile += gridDim.y)
                for (int b = 0; b < nboxes; b++, it++) {
                    const int s = it % p.stages, n = it / p.stages;
                    if (n > 0) mbar_wait_a(b_empty + 8 * s, (uint32_t)((n - 1) & 1));
                    mbar_expect_tx_a(b_full + 8 * s, 2 * kGxBoxBytes);
                    const uint32_t dst = s_x + (uint32_t)s * 2 * kGxBoxBytes;
                    asm volatile(
                        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                        ::"r"(dst), "l"((uint64_t)&tmXhi), "r"(col0 + b * 64), "r"(tile * 128), "r"(b_full + 8 * s) : "memory");
                    asm volatile(
                        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                        ::"r"(dst + kGxBoxBytes), "l"((uint64_t)&tmXlo), "r"(col0 + b * 64), "r"(tile * 128), "r"(b_full + 8 * s) : "memory");
                }
        }
    } else if (warp == 9) {
        // ===== MMA issuer: D_g[128 rows][Hgp] += X[:, group g's K slice] . W_g
        const uint32_t idesc = umma_idesc_bf16(128, p.Hgp);
        const uint32_t tmem_u = __shfl_sync(0xffffffffu, tmem, 0);
        const uint32_t lbo = (uint32_t)(p.Hgp / 8) * 128u, sbo = 128u;
        mbar_wait_a(b_w, 0);
        fence_proxy_async();
        int it = 0, li = 0;
        for (int tile = blockIdx.y; tile < ntiles; tile += gridDim.y, li++) {
            const int buf = li & 1;
            if (li >= 2) mbar_wait_a(b_tempty + 8 * buf, (uint32_t)(((li >> 1) - 1) & 1));
            tc_fence_after();
            for (int b = 0; b < nboxes; b++, it++) {
                const int s = it % p.stages, n = it / p.stages;
                mbar_wait_a(b_full + 8 * s, (uint32_t)(n & 1));
                tc_fence_after();
                const uint32_t xs = s_x + (uint32_t)s * 2 * kGxBoxBytes;
                const uint64_t xh = umma_desc_sw128(xs), xl = umma_desc_sw128(xs + kGxBoxBytes);
#pragma unroll
                for (int k = 0; k < 4; k++) {  // K step 16 = 32 bytes inside the 128-byte swizzle row of X
                    const int col = b * 64 + k * 16;
                    const int gl = col / p.Ig, kk = (col - gl * p.Ig) >> 4;
                    const uint32_t d = tmem_u + (uint32_t)(buf * Dc + gl * p.Hgp);
                    const uint32_t wa = s_w + (uint32_t)gl * gbytes + (uint32_t)kk * 2u * lbo;
                    const uint64_t wh = umma_desc_interleave(wa, lbo, sbo), wl = umma_desc_interleave(wa + wplane, lbo, sbo);
                    umma_bf16_ss_elect(d, xh + 2 * k, wh, idesc, kk != 0);
                    umma_bf16_ss_elect(d, xl + 2 * k, wh, idesc, 1u);
                    umma_bf16_ss_elect(d, xh + 2 * k, wl, idesc, 1u);
                }
                // frees the ring slot once the MMAs above have read it
                asm volatile("{\n\t.reg .pred e;\n\telect.sync _|e, 0xffffffff;\n\t"
                             "@e tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n\t}\n" ::"r"(b_empty + 8 * s) : "memory");
            }
            asm volatile("{\n\t.reg .pred e;\n\telect.sync _|e, 0xffffffff;\n\t"
                         "@e tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n\t}\n" ::"r"(b_tfull + 8 * buf) : "memory");
        }
    } else {
        // ===== epilogue: warp w owns TMEM lanes [32 (w % 4), +32) = rows, and the column half w / 4 of the buffer
        const int q = warp & 3, half = warp >> 2;
        const uint32_t stg = s_stage + (uint32_t)warp * kGxStageBytes;
        const int rr = lane >> 2, pc = lane & 3;               // transposed read: row rr + 8 i, 16-byte piece pc
        int li = 0;
        for (int tile = blockIdx.y; tile < ntiles; tile += gridDim.y, li++) {
            const int buf = li & 1;
            mbar_wait_a(b_tfull + 8 * buf, (uint32_t)((li >> 1) & 1));
            tc_fence_after();
            const int64_t mbase = (int64_t)tile * 128 + q * 32;
            for (int c = half * (Dc / 2); c < (half + 1) * (Dc / 2); c += 16) {
                float v[16];
                tmem_ld16(tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)(buf * Dc + c), v);
#pragma unroll
                for (int j = 0; j < 4; j++)
                    sts128(stg + lane * kGxStageRow + j * 16, make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]));
                __syncwarp();
                const int gl = c / p.Hgp, n = c - gl * p.Hgp + 4 * pc;   // column inside the group
                const int64_t col = (int64_t)(g0 + gl) * p.Hg + n;
                if (n < p.Hg) {
                    // residual rows first, all four in flight (as load / store pairs per row the compiler has to assume that
                    // res aliases y -- it does for df_out -- and serialises four DRAM round trips per chunk)
                    float4 rv[4];
#pragma unroll
                    for (int i = 0; i < 4; i++) {
                        const int64_t m = mbase + rr + 8 * i;
                        rv[i] = (p.res && m < p.M) ? *reinterpret_cast<const float4 *>(p.res + m * p.ldr + col) : make_float4(0.f, 0.f, 0.f, 0.f);
                    }
#pragma unroll
                    for (int i = 0; i < 4; i++) {
                        const int r = rr + 8 * i;
                        const int64_t m = mbase + r;
                        if (m >= p.M) continue;
                        float4 x = lds128(stg + r * kGxStageRow + pc * 16);
                        x.x = gx_act(x.x, p.act) * p.oscale + p.ooffset + rv[i].x; x.y = gx_act(x.y, p.act) * p.oscale + p.ooffset + rv[i].y;
                        x.z = gx_act(x.z, p.act) * p.oscale + p.ooffset + rv[i].z; x.w = gx_act(x.w, p.act) * p.oscale + p.ooffset + rv[i].w;
                        if (p.y) *reinterpret_cast<float4 *>(p.y + m * p.ldy + col) = x;
                        if (p.y_hi) {
                            uint32_t h0, l0, h1, l1;
                            bf16x2_split(x.x, x.y, h0, l0);
                            bf16x2_split(x.z, x.w, h1, l1);
                            *reinterpret_cast<uint2 *>(p.y_hi + m * p.ldp + col) = make_uint2(h0, h1);
                            *reinterpret_cast<uint2 *>(p.y_lo + m * p.ldp + col) = make_uint2(l0, l1);
                        }
                    }
                }
                __syncwarp();
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive_a(b_tempty + 8 * buf);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem, 512);
}

// fp32 [M][K] (row pitch ldx) -> BF16 hi / lo planes [M][K] (pitch K): fallback producer for inputs whose own
// producer has no plane-writing epilogue (FFMA precision modes, the H = 512 FFMA recurrence)
__global__ void __launch_bounds__(256) k_to_planes(const float *__restrict__ x, int64_t ldx, int64_t M, int K,
                                                   unsigned short *__restrict__ hi, unsigned short *__restrict__ lo) {
    const int64_t n4 = M * (K / 4);
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t m = i / (K / 4);
        const int k = (int)(i - m * (K / 4)) * 4;
        const float4 v = *reinterpret_cast<const float4 *>(x + m * ldx + k);
        uint32_t h0, l0, h1, l1;
        bf16x2_split(v.x, v.y, h0, l0);
        bf16x2_split(v.z, v.w, h1, l1);
        *reinterpret_cast<uint2 *>(hi + m * K + k) = make_uint2(h0, h1);
        *reinterpret_cast<uint2 *>(lo + m * K + k) = make_uint2(l0, l1);
    }
}

// ------------------------------------------------------------------------------- host side ----
typedef CUresult (*PFN_encodeTiled_gl)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                       const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                       CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

// Tensor maps are cached per (device, base, rows, cols, pitch): the arena hands out the same addresses call after call,
// so steady-state launches do not re-encode.
int cached_map_bf16(CUtensorMap *out, const void *base, int64_t rows, int64_t cols, int64_t ld, int box_rows) {
    static std::mutex mu;
    static std::map<std::tuple<int, const void *, int64_t, int64_t, int64_t, int>, CUtensorMap> cache;
    static PFN_encodeTiled_gl enc = nullptr;
    int dev = 0;
    cudaGetDevice(&dev);
    std::lock_guard<std::mutex> g(mu);
    auto key = std::make_tuple(dev, base, rows, cols, ld, box_rows);
    auto it = cache.find(key);
    if (it != cache.end()) { *out = it->second; return DFB_OK; }
    if (!enc) {
        void *fp = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fp, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess)
            return fail(DFB_ERR_CUDA, "cuTensorMapEncodeTiled entry point not available");
        enc = (PFN_encodeTiled_gl)fp;
    }
    cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
    cuuint64_t strides[1] = {(cuuint64_t)ld * 2};
    cuuint32_t box[2] = {64u, (cuuint32_t)box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUtensorMap m;
    CUresult r = enc(&m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, (void *)base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(DFB_ERR_CUDA, "cuTensorMapEncodeTiled (bf16) failed (%d)", (int)r);
    if (cache.size() > 4096) cache.clear();
    cache[key] = m;
    *out = m;
    return DFB_OK;
}

// 2-D fp32 row-major [rows][cols] (row pitch ld floats), box = [box_rows][32 floats = 128 B], 128-byte swizzle
int cached_map_f32_sw128(CUtensorMap *out, const void *base, int64_t rows, int64_t cols, int64_t ld, int box_rows) {
    static std::mutex mu;
    static std::map<std::tuple<int, const void *, int64_t, int64_t, int64_t, int>, CUtensorMap> cache;
    static PFN_encodeTiled_gl enc = nullptr;
    int dev = 0;
    cudaGetDevice(&dev);
    std::lock_guard<std::mutex> g(mu);
    auto key = std::make_tuple(dev, base, rows, cols, ld, box_rows);
    auto it = cache.find(key);
    if (it != cache.end()) { *out = it->second; return DFB_OK; }
    if (!enc) {
        void *fp = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fp, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess)
            return fail(DFB_ERR_CUDA, "cuTensorMapEncodeTiled entry point not available");
        enc = (PFN_encodeTiled_gl)fp;
    }
    cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
    cuuint64_t strides[1] = {(cuuint64_t)ld * 4};
    cuuint32_t box[2] = {32u, (cuuint32_t)box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUtensorMap m;
    CUresult r = enc(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, (void *)base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(DFB_ERR_CUDA, "cuTensorMapEncodeTiled (fp32) failed (%d)", (int)r);
    if (cache.size() > 4096) cache.clear();
    cache[key] = m;
    *out = m;
    return DFB_OK;
}

int launch_to_planes(cudaStream_t s, const float *x, int64_t ldx, int64_t M, int K, unsigned short *hi, unsigned short *lo) {
    if (K % 4 || ldx % 4) return fail(DFB_ERR_UNSUPPORTED, "to_planes: K = %d", K);
    int dev = 0, sms = 0;
    DFB_CUDA(cudaGetDevice(&dev));
    DFB_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    DFB_PROF("k_to_planes", s);
    k_to_planes<<<sms * 8, 256, 0, s>>>(x, ldx, M, K, hi, lo);
    DFB_LAUNCH_CHECK();
    return DFB_OK;
}

// Geometry of the tensor-core grouped linear for (G, Ig, Hg); returns false when the shape is outside the kernel.
bool gl_bx_geometry(int G, int Ig, int Hg, int *gpc_out, int *hgp_out, int *stages_out) {
    if (Ig % 16 || Hg % 4 || G < 1) return false;
    const int Hgp = (Hg + 15) / 16 * 16;
    if (Hgp > 256) return false;
    int best = 0;
    for (int gpc = 1; gpc <= G; gpc++) {
        if (G % gpc) continue;
        if (gpc * Hgp > 256 || (gpc * Hgp) % 32) continue;
        if ((gpc * Ig) % 64) continue;
        if ((size_t)gpc * Ig * Hgp * 4 > 100 * 1024) continue;
        best = gpc;
    }
    if (!best) return false;
    const int w = best * Ig * Hgp * 4;
    int stages = (227 * 1024 - 2048 - w - 8 * kGxStageBytes - 256) / (2 * kGxBoxBytes);
    if (stages > kGxMaxStages) stages = kGxMaxStages;
    if (stages < 2) return false;
    *gpc_out = best; *hgp_out = Hgp; *stages_out = stages;
    return true;
}

// Y = act(GL(X)) ... with X given as BF16 planes [M][K = G * Ig] (pitch ldx elements) and the host-packed weight image.
int launch_gl_bx(cudaStream_t s, const unsigned short *x_hi, const unsigned short *x_lo, int64_t ldx, const float *w_img,
                 const float *res, int64_t ldr, float *y, int64_t ldy, unsigned short *y_hi, unsigned short *y_lo, int64_t ldp,
                 int64_t M, int G, int Ig, int Hg, int act, float oscale, float ooffset) {
    int gpc = 0, Hgp = 0, stages = 0;
    if (!gl_bx_geometry(G, Ig, Hg, &gpc, &Hgp, &stages) || M <= 0 || M > 0x7fffffff || (ldx % 8) || ((uintptr_t)x_hi & 15) ||
        ((uintptr_t)x_lo & 15) || (y && ((ldy % 4) || ((uintptr_t)y & 15))) || (res && ((ldr % 4) || ((uintptr_t)res & 15))) ||
        (y_hi && ((ldp % 4) || ((uintptr_t)y_hi & 7) || ((uintptr_t)y_lo & 7))))
        return DFB_ERR_UNSUPPORTED;
    CUtensorMap mh, ml;
    int rc;
    if ((rc = cached_map_bf16(&mh, x_hi, M, (int64_t)G * Ig, ldx, 128)) || (rc = cached_map_bf16(&ml, x_lo, M, (int64_t)G * Ig, ldx, 128)))
        return rc;
    GlBxParams p{reinterpret_cast<const unsigned short *>(w_img), res, ldr, y, ldy, y_hi, y_lo, ldp,
                 (int)M, G, Ig, Hg, Hgp, gpc, act, stages, oscale, ooffset};
    const int smem = 1024 + stages * 2 * kGxBoxBytes + gpc * Ig * Hgp * 4 + 128 + 8 * kGxStageBytes + 256;
    static PerDeviceOnce attr_once;
    if (auto once_guard = attr_once.first())
        DFB_CUDA(cudaFuncSetAttribute(k_gl_bx, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    int dev = 0, sms = 0;
    DFB_CUDA(cudaGetDevice(&dev));
    DFB_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    const int slices = G / gpc, ntiles = (int)((M + 127) / 128);
    int groups = sms / slices;
    if (groups < 1) groups = 1;
    if (groups > ntiles) groups = ntiles;
    DFB_PROF("k_gl_bx", s);
    k_gl_bx<<<dim3((unsigned)slices, (unsigned)groups), kGxThreads, smem, s>>>(mh, ml, p);
    DFB_LAUNCH_CHECK();
    return DFB_OK;
}

}  // namespace dfb
