// mf_cnn.cu -- Mask R-CNN backbone (ResNet-101 + FPN) as tcgen05 / TMEM tensor-core GEMMs (sm_100a).
//
// Replaces the dense-contraction part of the reference's Keras/TensorFlow sidecar
// (Core/Segmentation/MaskRCNN/MaskRCNN.py.in:55-58,101-111 -> matterport mrcnn `resnet_graph` +
// FPN top-down path; network source un-vendored, build.sh:278).  Every convolution is an
// implicit-GEMM  D[M x Cout] = A[M x K] * W[Cout x K]^T  with M = N*Hout*Wout output pixels,
// K = kh*kw*Cin, activations NHWC bf16, frozen BatchNorm + conv bias folded into the weights
// and a per-channel bias, residual add and ReLU fused into the epilogue.
//
// GEMM kernel (one 128 x BN output tile per CTA, 192 threads):
//   warp 0     TMA producer: cp.async.bulk.tensor.2d (SWIZZLE_128B) A/B k-blocks of 64 into a
//              4-stage shared-memory ring, mbarrier expect_tx / complete_tx
//   warp 1     TMEM allocator + single-thread tcgen05.mma.cta_group::1.kind::f16 issuer
//              (M=128, N=BN, K=16 x4 per k-block, fp32 accumulators in TMEM), tcgen05.commit
//              releases ring slots and finally signals the epilogue
//   warps 2-5  epilogue: tcgen05.ld 32x32b.x32 (each warp its 32-lane TMEM quarter) ->
//              + bias, + residual, ReLU -> bf16 -> 64-byte row segments to HBM
// 1x1 stride-1 convolutions feed the activation tensor straight to TMA (no im2col); 3x3, 7x7
// and strided 1x1 go through a bf16 im2col buffer (round 1; TMA im2col descriptors = next).
#include "mf_common.cuh"
#include "mf_kernels.h"
#include <cuda.h>
#include <cuda_bf16.h>
#include <vector>
#include <string>
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <map>
#include <array>

namespace mfb {

// ------------------------------------------------------------------------------------------
// PTX wrappers
// ------------------------------------------------------------------------------------------
MF_D uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
MF_D void mbar_init(uint64_t* bar, uint32_t count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count)); }
MF_D void mbar_expect_tx(uint64_t* bar, uint32_t bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory"); }
MF_D void mbar_wait(uint64_t* bar, uint32_t parity)
{
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "WAIT_LOOP:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE;\n\t"
        "bra WAIT_LOOP;\n\t"
        "DONE:\n\t}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
MF_D void tma_load_2d(const CUtensorMap* map, uint64_t* bar, void* dst, int c0, int c1)
{
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                 ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1) : "memory");
}
MF_D void tma_load_3d(const CUtensorMap* map, uint64_t* bar, void* dst, int c0, int c1, int c2)
{
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
                 ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
MF_D void tcgen05_commit(uint64_t* bar) { asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory"); }
MF_D void tcgen05_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
MF_D void tcgen05_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
MF_D void umma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate)
{
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
// shared-memory matrix descriptor: K-major operand, SWIZZLE_128B, rows of 128 B, 8-row groups 1024 B apart
MF_D uint64_t make_smem_desc(uint32_t saddr)
{
    uint64_t d = 0;
    d |= (uint64_t)((saddr & 0x3FFFF) >> 4);          // start address, 16-byte units          bits [0,14)
    d |= (uint64_t)1 << 16;                            // leading byte offset (unused for SW128 K-major) = 1
    d |= (uint64_t)(1024 >> 4) << 32;                  // stride byte offset: 8 rows x 128 B      bits [32,46)
    d |= (uint64_t)1 << 46;                            // descriptor version (Blackwell)
    d |= (uint64_t)2 << 61;                            // layout type: SWIZZLE_128B
    return d;
}
MF_D void tmem_ld32(uint32_t taddr, uint32_t* r)
{
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 "
                 "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
                   "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),
                   "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
                   "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
                 : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// ------------------------------------------------------------------------------------------
// GEMM: out[M x N] (bf16) = relu?( A[M x K] * B[N x K]^T + bias[N] + residual[M x N] )
// ------------------------------------------------------------------------------------------
constexpr int GEMM_BM = 128, GEMM_BK = 64, GEMM_STAGES = 3, GEMM_THREADS = 192;   // 3 stages (<= 99 KB): two CTAs per SM, one's epilogue overlaps the other's main loop

// implicit-GEMM geometry of a 3x3 / stride 1 / pad 1 convolution: the A operand is the NHWC activation itself, seen through a 3-D
// tensor map (C, W, H); an M tile is a Wbox x Hbox pixel block (Wbox*Hbox = 128) and each of the 9 taps is the same block shifted
// by (kx-1, ky-1) -- TMA zero-fills the out-of-image part, which IS the padding.  No im2col buffer.
struct ConvGeom { int mode; int Wimg, Himg, Wbox, Hbox, cblocks; };

template <int BN>
__global__ void __launch_bounds__(GEMM_THREADS, 1) k_gemm_bf16_tcgen05(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapB,
                                                                        const float* __restrict__ bias, const __nv_bfloat16* __restrict__ residual,
                                                                        __nv_bfloat16* __restrict__ out, int M, int N, int K, int relu, ConvGeom geo)
{
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    // carve: [stages x A tile 16 KB][stages x B tile BN*128 B][barriers][tmem ptr]
    uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    constexpr uint32_t A_BYTES = GEMM_BM * GEMM_BK * 2, B_BYTES = BN * GEMM_BK * 2;
    uint8_t* sA = smem;
    uint8_t* sB = smem + GEMM_STAGES * A_BYTES;
    uint64_t* full = (uint64_t*)(sB + GEMM_STAGES * B_BYTES);
    uint64_t* empty = full + GEMM_STAGES;
    uint64_t* tmem_full = empty + GEMM_STAGES;
    uint32_t* tmem_ptr = (uint32_t*)(tmem_full + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int tile_m = blockIdx.x, tile_n = blockIdx.y;
    const int num_k = K / GEMM_BK;
    int px0 = 0, py0 = 0;
    if (geo.mode) { const int tilesX = geo.Wimg / geo.Wbox; py0 = (tile_m / tilesX) * geo.Hbox; px0 = (tile_m % tilesX) * geo.Wbox; }

    if (warp == 0 && lane == 0) {
        for (int s = 0; s < GEMM_STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
        mbar_init(tmem_full, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_ptr)), "n"(BN) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    const uint32_t tmem_base = *tmem_ptr;

    if (warp == 0) {
        if (lane == 0) {
            // ===== TMA producer =====
            for (int kb = 0; kb < num_k; ++kb) {
                const int s = kb % GEMM_STAGES;
                const uint32_t ph = (kb / GEMM_STAGES) & 1;
                mbar_wait(&empty[s], ph ^ 1);
                mbar_expect_tx(&full[s], A_BYTES + B_BYTES);
                if (geo.mode) {
                    const int tap = kb / geo.cblocks, cb = kb - tap * geo.cblocks;
                    tma_load_3d(&mapA, &full[s], sA + s * A_BYTES, cb * GEMM_BK, px0 + tap % 3 - 1, py0 + tap / 3 - 1);
                } else
                    tma_load_2d(&mapA, &full[s], sA + s * A_BYTES, kb * GEMM_BK, tile_m * GEMM_BM);
                tma_load_2d(&mapB, &full[s], sB + s * B_BYTES, kb * GEMM_BK, tile_n * BN);
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            // ===== MMA issuer =====
            // instruction descriptor: D=f32 (bit 4), A=B=bf16 (bits 7, 10), K-major both, N>>3 at 17, M>>4 at 24
            const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(GEMM_BM >> 4) << 24);
            for (int kb = 0; kb < num_k; ++kb) {
                const int s = kb % GEMM_STAGES;
                const uint32_t ph = (kb / GEMM_STAGES) & 1;
                mbar_wait(&full[s], ph);
                tcgen05_fence_after();
                const uint64_t adesc = make_smem_desc(smem_u32(sA + s * A_BYTES));
                const uint64_t bdesc = make_smem_desc(smem_u32(sB + s * B_BYTES));
#pragma unroll
                for (int k = 0; k < GEMM_BK / 16; ++k)          // UMMA_K = 16 bf16 = 32 bytes: advance the start address by 2 (16-byte units)
                    umma_bf16(tmem_base, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), idesc, (kb | k) ? 1u : 0u);
                tcgen05_commit(&empty[s]);                      // frees the ring slot when these MMAs retire
            }
            tcgen05_commit(tmem_full);                          // accumulator complete
        }
    } else {
        // ===== epilogue: warps 2..5 own TMEM lane quarters (warp id % 4) =====
        const int q = warp & 3;
        mbar_wait(tmem_full, 0);
        tcgen05_fence_after();
        const int rt = q * 32 + lane;
        const int row = geo.mode ? (py0 + rt / geo.Wbox) * geo.Wimg + px0 + rt % geo.Wbox : tile_m * GEMM_BM + rt;
#pragma unroll 1
        for (int c0 = 0; c0 < BN; c0 += 32) {
            uint32_t r[32];
            tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)c0, r);
            const int col = tile_n * BN + c0;
            if (row < M) {
                __nv_bfloat16* optr = out + (size_t)row * N + col;
                const __nv_bfloat16* rptr = residual ? residual + (size_t)row * N + col : nullptr;
#pragma unroll
                for (int v = 0; v < 4; ++v) {                  // 4 x 16 bytes = 32 bf16
                    uint4 res4 = make_uint4(0, 0, 0, 0);
                    if (rptr) res4 = *reinterpret_cast<const uint4*>(rptr + v * 8);
                    const __nv_bfloat16* rb = reinterpret_cast<const __nv_bfloat16*>(&res4);
                    uint4 o4;
                    __nv_bfloat16* ob = reinterpret_cast<__nv_bfloat16*>(&o4);
#pragma unroll
                    const float4 b0 = __ldg(reinterpret_cast<const float4*>(bias + col + v * 8)), b1 = __ldg(reinterpret_cast<const float4*>(bias + col + v * 8 + 4));
                    const float bb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
                    for (int e = 0; e < 8; ++e) {
                        float x = __uint_as_float(r[v * 8 + e]) + bb[e];
                        if (rptr) x += __bfloat162float(rb[e]);
                        if (relu) x = fmaxf(x, 0.f);
                        ob[e] = __float2bfloat16(x);
                    }
                    *reinterpret_cast<uint4*>(optr + v * 8) = o4;
                }
            }
        }
    }
    tcgen05_fence_before();
    __syncthreads();
    if (warp == 1) {
        tcgen05_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(BN) : "memory");
    }
}

// ------------------------------------------------------------------------------------------
// data-movement kernels around the GEMMs (NHWC bf16)
// ------------------------------------------------------------------------------------------
// im2col: A[m][(ky*kw + kx)*Cin + c], K padded with zeros to Kpad; 8 channels (16 B) per thread when Cin % 8 == 0
__global__ void k_im2col(const __nv_bfloat16* __restrict__ in, int Hin, int Win, int Cin, int Hout, int Wout, int kh, int kw, int stride,
                         int pad, int Kpad, __nv_bfloat16* __restrict__ A)
{
    const int K = kh * kw * Cin;
    const size_t total8 = (size_t)Hout * Wout * (Kpad / 8);
    for (size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x; t < total8; t += (size_t)gridDim.x * blockDim.x) {
        const int k8 = (int)(t % (Kpad / 8));
        const size_t m = t / (Kpad / 8);
        const int ox = (int)(m % Wout), oy = (int)(m / Wout);
        uint4 v = make_uint4(0, 0, 0, 0);
        const int k0 = k8 * 8;
        if ((Cin & 7) == 0) {
            if (k0 < K) {
                const int tap = k0 / Cin, c = k0 - tap * Cin;
                const int ky = tap / kw, kx = tap - ky * kw;
                const int iy = oy * stride - pad + ky, ix = ox * stride - pad + kx;
                if (iy >= 0 && iy < Hin && ix >= 0 && ix < Win) v = *reinterpret_cast<const uint4*>(in + ((size_t)iy * Win + ix) * Cin + c);
            }
        } else {
            __nv_bfloat16* e = reinterpret_cast<__nv_bfloat16*>(&v);
            for (int j = 0; j < 8; ++j) {
                const int k = k0 + j;
                if (k < K) {
                    const int tap = k / Cin, c = k - tap * Cin;
                    const int ky = tap / kw, kx = tap - ky * kw;
                    const int iy = oy * stride - pad + ky, ix = ox * stride - pad + kx;
                    if (iy >= 0 && iy < Hin && ix >= 0 && ix < Win) e[j] = in[((size_t)iy * Win + ix) * Cin + c];
                }
            }
        }
        *reinterpret_cast<uint4*>(A + m * Kpad + k0) = v;
    }
}
// 3x3 / stride 2 max pool, TensorFlow "same" padding (extra row/column at the END): window [2o, 2o+2] clipped
__global__ void k_maxpool3s2(const __nv_bfloat16* __restrict__ in, int Hin, int Win, int C, int Hout, int Wout, __nv_bfloat16* __restrict__ out)
{
    const size_t total = (size_t)Hout * Wout * C;
    for (size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (size_t)gridDim.x * blockDim.x) {
        const int c = (int)(t % C); const size_t p = t / C;
        const int ox = (int)(p % Wout), oy = (int)(p / Wout);
        float best = -INFINITY;
        for (int ky = 0; ky < 3; ++ky)
            for (int kx = 0; kx < 3; ++kx) {
                const int iy = 2 * oy + ky, ix = 2 * ox + kx;
                if (iy < Hin && ix < Win) best = fmaxf(best, __bfloat162float(in[((size_t)iy * Win + ix) * C + c]));
            }
        out[t] = __float2bfloat16(best);
    }
}
// FPN top-down: out = lateral + nearest-upsample2(top)
__global__ void k_upsample_add(const __nv_bfloat16* __restrict__ lateral, const __nv_bfloat16* __restrict__ top, int H, int W, int C,
                               __nv_bfloat16* __restrict__ out)
{
    const size_t total = (size_t)H * W * C;
    for (size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (size_t)gridDim.x * blockDim.x) {
        const int c = (int)(t % C); const size_t p = t / C;
        const int x = (int)(p % W), y = (int)(p / W);
        float v = __bfloat162float(lateral[t]) + __bfloat162float(top[((size_t)(y / 2) * (W / 2) + x / 2) * C + c]);
        out[t] = __float2bfloat16(v);
    }
}
// 1x1 / stride 2 sub-sampling (P6 = MaxPool 1x1 stride 2 of P5; also the A operand of strided 1x1 convolutions)
__global__ void k_subsample2(const __nv_bfloat16* __restrict__ in, int Hin, int Win, int C, __nv_bfloat16* __restrict__ out)
{
    const int Hout = Hin / 2, Wout = Win / 2;
    const size_t total8 = (size_t)Hout * Wout * (C / 8);
    for (size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x; t < total8; t += (size_t)gridDim.x * blockDim.x) {
        const int c8 = (int)(t % (C / 8)); const size_t p = t / (C / 8);
        const int ox = (int)(p % Wout), oy = (int)(p / Wout);
        *reinterpret_cast<uint4*>(out + p * C + c8 * 8) = *reinterpret_cast<const uint4*>(in + ((size_t)(2 * oy) * Win + 2 * ox) * C + c8 * 8);
    }
}
// letter-boxed network input: uint8 RGB HxW -> bf16 NHWC SxS (bilinear resize to fit, zero padding, mean pixel subtracted)
__global__ void k_mold_input(const uchar4* __restrict__ rgb, int W, int H, int S, float scale, int offx, int offy, int newW, int newH,
                             __nv_bfloat16* __restrict__ out)
{
    const size_t total = (size_t)S * S;
    for (size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (size_t)gridDim.x * blockDim.x) {
        const int x = (int)(t % S), y = (int)(t / S);
        float r = 0, g = 0, b = 0;
        const int lx = x - offx, ly = y - offy;
        if (lx >= 0 && lx < newW && ly >= 0 && ly < newH) {
            float sx = fminf(fmaxf((lx + 0.5f) / scale - 0.5f, 0.f), (float)(W - 1)), sy = fminf(fmaxf((ly + 0.5f) / scale - 0.5f, 0.f), (float)(H - 1));
            int x0 = (int)sx, y0 = (int)sy, x1 = min(x0 + 1, W - 1), y1 = min(y0 + 1, H - 1);
            float fx = sx - x0, fy = sy - y0;
            uchar4 a = rgb[y0 * W + x0], bb = rgb[y0 * W + x1], c = rgb[y1 * W + x0], d = rgb[y1 * W + x1];
            r = (a.x * (1 - fx) + bb.x * fx) * (1 - fy) + (c.x * (1 - fx) + d.x * fx) * fy - 123.7f;      // MEAN_PIXEL (mrcnn config)
            g = (a.y * (1 - fx) + bb.y * fx) * (1 - fy) + (c.y * (1 - fx) + d.y * fx) * fy - 116.8f;
            b = (a.z * (1 - fx) + bb.z * fx) * (1 - fy) + (c.z * (1 - fx) + d.z * fx) * fy - 103.9f;
        }
        out[t * 3] = __float2bfloat16(r); out[t * 3 + 1] = __float2bfloat16(g); out[t * 3 + 2] = __float2bfloat16(b);
    }
}

// ------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                    const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static PFN_encodeTiled g_encode = nullptr;
static std::string g_cnn_err;

static bool ensure_encode()
{
    if (g_encode) return true;
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres) != cudaSuccess || !fn) { g_cnn_err = "cuTensorMapEncodeTiled not available"; return false; }
    g_encode = (PFN_encodeTiled)fn;
    return true;
}
// 2-D row-major [rows x K] bf16, box = {64, boxRows}, SWIZZLE_128B
static bool make_map(CUtensorMap* m, const void* ptr, uint64_t rows, uint64_t K, uint32_t boxRows)
{
    cuuint64_t dims[2] = {K, rows};
    cuuint64_t strides[1] = {K * 2};
    cuuint32_t box[2] = {64, boxRows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = g_encode(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                          CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { g_cnn_err = "cuTensorMapEncodeTiled failed: " + std::to_string((int)r); return false; }
    return true;
}

// 3-D map over an NHWC activation (C, W, H), box = {64, Wbox, Hbox}, SWIZZLE_128B
static bool make_map_nhwc(CUtensorMap* m, const void* ptr, int C, int W, int H, int Wbox, int Hbox)
{
    cuuint64_t dims[3] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H};
    cuuint64_t strides[2] = {(cuuint64_t)C * 2, (cuuint64_t)W * C * 2};
    cuuint32_t box[3] = {64, (cuuint32_t)Wbox, (cuuint32_t)Hbox};
    cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = g_encode(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(ptr), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                          CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { g_cnn_err = "cuTensorMapEncodeTiled (3-D) failed: " + std::to_string((int)r); return false; }
    return true;
}

// Tensor maps are pure functions of (pointer, geometry): the backbone's buffers never move, so every map is encoded ONCE (first forward) and
// reused by all later launches (round 1 encoded two maps per GEMM per forward on the host: 224 driver calls in front of 112 launches).
static std::map<std::array<uint64_t, 6>, CUtensorMap> g_mapCache;
static bool cached_map(CUtensorMap* m, const void* ptr, uint64_t rows, uint64_t K, uint32_t boxRows)
{
    const std::array<uint64_t, 6> key = {(uint64_t)(uintptr_t)ptr, rows, K, boxRows, 0, 0};
    auto it = g_mapCache.find(key);
    if (it != g_mapCache.end()) { *m = it->second; return true; }
    if (!make_map(m, ptr, rows, K, boxRows)) return false;
    if (g_mapCache.size() < 4096) g_mapCache[key] = *m;
    return true;
}
static bool cached_map_nhwc(CUtensorMap* m, const void* ptr, int Cin, int Wimg, int Himg, int Wbox, int Hbox)
{
    const std::array<uint64_t, 6> key = {(uint64_t)(uintptr_t)ptr, (uint64_t)Cin, (uint64_t)Wimg, (uint64_t)Himg, (uint64_t)Wbox, (uint64_t)Hbox + 1};
    auto it = g_mapCache.find(key);
    if (it != g_mapCache.end()) { *m = it->second; return true; }
    if (!make_map_nhwc(m, ptr, Cin, Wimg, Himg, Wbox, Hbox)) return false;
    if (g_mapCache.size() < 4096) g_mapCache[key] = *m;
    return true;
}

template <int BN>
static size_t gemm_smem_bytes() { return (size_t)GEMM_STAGES * (GEMM_BM * GEMM_BK * 2 + BN * GEMM_BK * 2) + 256 + 1024; }

const char* cnn_last_error() { return g_cnn_err.c_str(); }

// D = relu?(A * B^T + bias + residual); all device pointers; K % 64 == 0, N % 64 == 0
// conv3x3 != nullptr: A is an NHWC activation [Himg x Wimg x Cin] and the GEMM is the implicit 3x3/s1/p1 convolution (K = 9*Cin)
int launch_gemm_bf16(const void* A, const void* B, const float* bias, const void* residual, void* out, int M, int N, int K, int relu, cudaStream_t s,
                     const int* conv3x3 /* Wimg, Himg, Cin */ = nullptr)
{
    if (!ensure_encode()) return -1;
    if (K % 64 || N % 64 || M <= 0) { g_cnn_err = "gemm: need K % 64 == 0 and N % 64 == 0"; return -2; }
    const int mtiles = (M + GEMM_BM - 1) / GEMM_BM;
    // fill the machine: with few M tiles prefer the narrow N tile (twice the CTAs)
    const int BN = (N % 128 == 0 && mtiles * (N / 128) >= 148) ? 128 : 64;
    CUtensorMap mA, mB;
    ConvGeom geo; memset(&geo, 0, sizeof geo);
    if (conv3x3) {
        const int Wimg = conv3x3[0], Himg = conv3x3[1], Cin = conv3x3[2];
        geo.mode = 1; geo.Wimg = Wimg; geo.Himg = Himg; geo.Wbox = Wimg >= 128 ? 128 : Wimg; geo.Hbox = 128 / geo.Wbox; geo.cblocks = Cin / 64;
        if (Wimg % geo.Wbox || Himg % geo.Hbox || Cin % 64 || K != 9 * Cin) { g_cnn_err = "conv3x3: unsupported geometry"; return -2; }
        if (!cached_map_nhwc(&mA, A, Cin, Wimg, Himg, geo.Wbox, geo.Hbox)) return -3;
    } else if (!cached_map(&mA, A, (uint64_t)M, (uint64_t)K, GEMM_BM)) return -3;
    if (!cached_map(&mB, B, (uint64_t)N, (uint64_t)K, (uint32_t)BN)) return -3;
    dim3 grid(mtiles, N / BN);
    prof_mark(s, BN == 128 ? "k_gemm_bf16_tcgen05_n128" : "k_gemm_bf16_tcgen05_n64");
    if (BN == 128) {
        static bool attr = false;
        if (!attr) { cudaFuncSetAttribute(k_gemm_bf16_tcgen05<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)gemm_smem_bytes<128>()); attr = true; }
        k_gemm_bf16_tcgen05<128><<<grid, GEMM_THREADS, gemm_smem_bytes<128>(), s>>>(mA, mB, bias, (const __nv_bfloat16*)residual, (__nv_bfloat16*)out, M, N, K, relu, geo);
    } else {
        static bool attr = false;
        if (!attr) { cudaFuncSetAttribute(k_gemm_bf16_tcgen05<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)gemm_smem_bytes<64>()); attr = true; }
        k_gemm_bf16_tcgen05<64><<<grid, GEMM_THREADS, gemm_smem_bytes<64>(), s>>>(mA, mB, bias, (const __nv_bfloat16*)residual, (__nv_bfloat16*)out, M, N, K, relu, geo);
    }
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) { g_cnn_err = std::string("gemm launch: ") + cudaGetErrorString(e); return -4; }
    return 0;
}

// ---- backbone ---------------------------------------------------------------------------
struct ConvLayer {
    int Cin, Cout, k, stride, pad, Kpad;
    size_t wOff, bOff;      // offsets into the weight / bias pools (elements)
};

struct Backbone {
    int S;                                    // square network input (1024)
    std::vector<ConvLayer> layers;
    __nv_bfloat16* dW = nullptr; float* dB = nullptr;
    std::vector<float> hW; std::vector<float> hB;            // fp32 master copy (already bf16-rounded) for the reference check
    __nv_bfloat16 *input = nullptr, *bufA = nullptr, *bufB = nullptr, *bufC = nullptr, *bufS = nullptr, *col = nullptr;
    __nv_bfloat16 *C2 = nullptr, *C3 = nullptr, *C4 = nullptr, *C5 = nullptr, *P[5] = {nullptr, nullptr, nullptr, nullptr, nullptr}, *lat = nullptr, *td = nullptr;
    double flops = 0;
    int gemms = 0;
};

static uint32_t lcg(uint32_t& s) { s = s * 1664525u + 1013904223u; return s; }
static float urand(uint32_t& s) { return (float)(lcg(s) >> 8) * (1.0f / 16777216.0f) * 2.f - 1.f; }
static float bf16_round(float f) { return __bfloat162float(__float2bfloat16(f)); }

static int add_conv(Backbone* b, int Cin, int Cout, int k, int stride, int pad, uint32_t& seed, float gain)
{
    ConvLayer L; L.Cin = Cin; L.Cout = Cout; L.k = k; L.stride = stride; L.pad = pad;
    int K = k * k * Cin; L.Kpad = (K + 63) / 64 * 64;
    L.wOff = b->hW.size(); L.bOff = b->hB.size();
    float sc = gain * sqrtf(2.0f / (float)K);                        // He-style so activations stay O(1) through 101 layers
    b->hW.resize(L.wOff + (size_t)Cout * L.Kpad, 0.f);
    for (int o = 0; o < Cout; ++o)
        for (int kk = 0; kk < K; ++kk) b->hW[L.wOff + (size_t)o * L.Kpad + kk] = bf16_round(urand(seed) * sc * 1.7320508f);
    b->hB.resize(L.bOff + Cout);
    for (int o = 0; o < Cout; ++o) b->hB[L.bOff + o] = urand(seed) * 0.05f;
    b->layers.push_back(L);
    return (int)b->layers.size() - 1;
}

// runs conv layer li on in[Hin x Win x Cin] -> out[Hout x Wout x Cout]
static int run_conv(Backbone* b, int li, const __nv_bfloat16* in, int Hin, int Win, __nv_bfloat16* out, const __nv_bfloat16* residual, int relu, cudaStream_t s,
                    int* HoutP = nullptr, int* WoutP = nullptr)
{
    const ConvLayer& L = b->layers[li];
    const int Hout = (Hin + 2 * L.pad - L.k) / L.stride + 1, Wout = (Win + 2 * L.pad - L.k) / L.stride + 1;
    const int M = Hout * Wout;
    const __nv_bfloat16* A = in;
    const int wbox = Win >= 128 ? 128 : Win;
    const bool implicit3 = L.k == 3 && L.stride == 1 && L.pad == 1 && (L.Cin % 64) == 0 && Win <= 128 * 1024 && (128 % wbox) == 0 && (Win % wbox) == 0 &&
                           (Hin % (128 / wbox)) == 0 && wbox >= 8;
    if (implicit3) {
        int g3[3] = {Win, Hin, L.Cin};
        int rc = launch_gemm_bf16(in, b->dW + L.wOff, b->dB + L.bOff, residual, out, M, L.Cout, L.Kpad, relu, s, g3);
        b->flops += 2.0 * M * (double)L.Cout * (double)(9 * L.Cin);
        b->gemms++;
        if (HoutP) *HoutP = Hout;
        if (WoutP) *WoutP = Wout;
        return rc;
    }
    if (!(L.k == 1 && L.stride == 1)) {
        if (L.k == 1 && L.stride == 2 && (L.Cin % 8) == 0) {
            prof_mark(s, "k_subsample2"); k_subsample2<<<592, 256, 0, s>>>(in, Hin, Win, L.Cin, b->col);
        } else {
            prof_mark(s, "k_im2col"); k_im2col<<<1184, 256, 0, s>>>(in, Hin, Win, L.Cin, Hout, Wout, L.k, L.k, L.stride, L.pad, L.Kpad, b->col);
        }
        A = b->col;
    }
    int rc = launch_gemm_bf16(A, b->dW + L.wOff, b->dB + L.bOff, residual, out, M, L.Cout, L.Kpad, relu, s);
    b->flops += 2.0 * M * (double)L.Cout * (double)(L.k * L.k * L.Cin);
    b->gemms++;
    if (HoutP) *HoutP = Hout;
    if (WoutP) *WoutP = Wout;
    return rc;
}

}  // namespace mfb

using namespace mfb;

// ==========================================================================================
// C ABI (declared in include/maskfusion_b200.h)
// ==========================================================================================
struct mf_backbone { Backbone b; cudaStream_t stream; std::vector<int> plan; int stem, fpnLat[4], fpnOut[4]; std::vector<int> blockConv; };

extern "C" const char* mf_cnn_last_error(void) { return cnn_last_error(); }

extern "C" int mf_gemm_bf16(const void* dA, const void* dB, const float* dBias, const void* dResidual, void* dOut, int M, int N, int K, int relu, void* stream)
{
    int rc = launch_gemm_bf16(dA, dB, dBias, dResidual, dOut, M, N, K, relu, (cudaStream_t)stream);
    return rc;
}

// implicit-GEMM 3x3 / stride 1 / pad 1 convolution on an NHWC bf16 activation (weights [Cout][3][3][Cin] bf16)
extern "C" int mf_conv3x3_bf16(const void* dIn, const void* dW, const float* dBias, const void* dResidual, void* dOut, int H, int W, int Cin, int Cout,
                               int relu, void* stream)
{
    int g3[3] = {W, H, Cin};
    return launch_gemm_bf16(dIn, dW, dBias, dResidual, dOut, H * W, Cout, 9 * Cin, relu, (cudaStream_t)stream, g3);
}

// ResNet-101 (stages 3,4,23,3; stride in the first 1x1 of each stage as in Keras/matterport) + FPN(256)
extern "C" mf_backbone* mf_backbone_create(int input_size, unsigned seed, void* stream)
{
    if (input_size % 64) { g_cnn_err = "input size must be a multiple of 64 (mrcnn: IMAGE_MAX_DIM=1024)"; return nullptr; }
    mf_backbone* h = new mf_backbone;
    Backbone* b = &h->b;
    b->S = input_size; h->stream = (cudaStream_t)stream;
    uint32_t sd = seed ? seed : 1u;
    h->stem = add_conv(b, 3, 64, 7, 2, 3, sd, 1.0f);
    const int nblocks[4] = {3, 4, 23, 3}, mid[4] = {64, 128, 256, 512};
    int cin = 64;
    for (int st = 0; st < 4; ++st)
        for (int blk = 0; blk < nblocks[st]; ++blk) {
            const int stride = (blk == 0 && st > 0) ? 2 : 1;
            const int f = mid[st], cout = f * 4;
            h->blockConv.push_back(add_conv(b, cin, f, 1, stride, 0, sd, 1.0f));
            h->blockConv.push_back(add_conv(b, f, f, 3, 1, 1, sd, 1.0f));
            h->blockConv.push_back(add_conv(b, f, cout, 1, 1, 0, sd, 0.5f));           // damped: residual sums keep O(1) variance
            h->blockConv.push_back(blk == 0 ? add_conv(b, cin, cout, 1, stride, 0, sd, 0.7f) : -1);
            cin = cout;
        }
    const int cdim[4] = {256, 512, 1024, 2048};
    for (int i = 0; i < 4; ++i) h->fpnLat[i] = add_conv(b, cdim[i], 256, 1, 1, 0, sd, 0.7f);
    for (int i = 0; i < 4; ++i) h->fpnOut[i] = add_conv(b, 256, 256, 3, 1, 1, sd, 1.0f);
    // device pools
    const int S = b->S;
    std::vector<__nv_bfloat16> wbf(b->hW.size());
    for (size_t i = 0; i < wbf.size(); ++i) wbf[i] = __float2bfloat16(b->hW[i]);
    bool ok = cudaMalloc(&b->dW, wbf.size() * 2) == cudaSuccess && cudaMalloc(&b->dB, b->hB.size() * 4) == cudaSuccess;
    const size_t big = (size_t)(S / 2) * (S / 2) * 64;                         // stem output == largest activation (elements): C1 512x512x64 = C2 256x256x256
    ok = ok && cudaMalloc(&b->input, (size_t)S * S * 3 * 2) == cudaSuccess;
    ok = ok && cudaMalloc(&b->bufA, big * 2) == cudaSuccess && cudaMalloc(&b->bufB, big * 2) == cudaSuccess && cudaMalloc(&b->bufC, big * 2) == cudaSuccess && cudaMalloc(&b->bufS, big * 2) == cudaSuccess;
    const size_t colElems = (size_t)(S / 4) * (S / 4) * 9 * 256;               // largest im2col: FPN P2 3x3 on 256x256x256 (and C2 3x3 64ch is smaller); stem: 512*512*192
    ok = ok && cudaMalloc(&b->col, colElems * 2) == cudaSuccess;
    const int fs[4] = {S / 4, S / 8, S / 16, S / 32};
    ok = ok && cudaMalloc(&b->C2, (size_t)fs[0] * fs[0] * 256 * 2) == cudaSuccess && cudaMalloc(&b->C3, (size_t)fs[1] * fs[1] * 512 * 2) == cudaSuccess &&
         cudaMalloc(&b->C4, (size_t)fs[2] * fs[2] * 1024 * 2) == cudaSuccess && cudaMalloc(&b->C5, (size_t)fs[3] * fs[3] * 2048 * 2) == cudaSuccess;
    for (int i = 0; i < 4; ++i) ok = ok && cudaMalloc(&b->P[i], (size_t)fs[i] * fs[i] * 256 * 2) == cudaSuccess;
    ok = ok && cudaMalloc(&b->P[4], (size_t)(fs[3] / 2) * (fs[3] / 2) * 256 * 2) == cudaSuccess;
    ok = ok && cudaMalloc(&b->lat, (size_t)fs[0] * fs[0] * 256 * 2) == cudaSuccess && cudaMalloc(&b->td, (size_t)fs[0] * fs[0] * 256 * 2) == cudaSuccess;
    if (!ok) { g_cnn_err = "backbone: cudaMalloc failed"; delete h; return nullptr; }
    cudaMemcpy(b->dW, wbf.data(), wbf.size() * 2, cudaMemcpyHostToDevice);
    cudaMemcpy(b->dB, b->hB.data(), b->hB.size() * 4, cudaMemcpyHostToDevice);
    return h;
}

extern "C" void mf_backbone_destroy(mf_backbone* h)
{
    if (!h) return;
    Backbone* b = &h->b;
    void* ptrs[] = {b->dW, b->dB, b->input, b->bufA, b->bufB, b->bufC, b->bufS, b->col, b->C2, b->C3, b->C4, b->C5, b->P[0], b->P[1], b->P[2], b->P[3], b->P[4], b->lat, b->td};
    for (void* p : ptrs) if (p) cudaFree(p);
    delete h;
}

extern "C" int mf_backbone_num_layers(mf_backbone* h) { return h ? (int)h->b.layers.size() : -1; }
// layer table: Cin Cout k stride pad Kpad
extern "C" int mf_backbone_layer(mf_backbone* h, int i, int* out6)
{
    if (!h || i < 0 || i >= (int)h->b.layers.size()) return -1;
    const ConvLayer& L = h->b.layers[i];
    out6[0] = L.Cin; out6[1] = L.Cout; out6[2] = L.k; out6[3] = L.stride; out6[4] = L.pad; out6[5] = L.Kpad;
    return 0;
}
// weights [Cout x Kpad] fp32 (bf16-representable), (ky,kx,cin) order along K; bias [Cout]
extern "C" int mf_backbone_get_weights(mf_backbone* h, int i, float* w, float* bias)
{
    if (!h || i < 0 || i >= (int)h->b.layers.size()) return -1;
    const ConvLayer& L = h->b.layers[i];
    memcpy(w, h->b.hW.data() + L.wOff, (size_t)L.Cout * L.Kpad * sizeof(float));
    memcpy(bias, h->b.hB.data() + L.bOff, (size_t)L.Cout * sizeof(float));
    return 0;
}

// forward on an already-moulded input (device, NHWC bf16 S x S x 3).  Outputs stay on the device (P2..P6, NHWC bf16).
extern "C" int mf_backbone_forward(mf_backbone* h, const void* d_input)
{
    if (!h) return -1;
    Backbone* b = &h->b; cudaStream_t s = h->stream;
    b->flops = 0; b->gemms = 0;
    const int S = b->S;
    int H, W;
    // C1: 7x7/2 + ReLU, max-pool 3x3/2
    if (run_conv(b, h->stem, (const __nv_bfloat16*)d_input, S, S, b->bufA, nullptr, 1, s, &H, &W)) return -2;
    prof_mark(s, "k_maxpool3s2"); k_maxpool3s2<<<1184, 256, 0, s>>>(b->bufA, H, W, 64, H / 2, W / 2, b->bufB);
    H /= 2; W /= 2;
    __nv_bfloat16* x = b->bufB;                      // current block input
    __nv_bfloat16* pool[3] = {b->bufA, b->bufC, b->bufS};
    const int nblocks[4] = {3, 4, 23, 3};
    __nv_bfloat16* stageOut[4] = {b->C2, b->C3, b->C4, b->C5};
    size_t bi = 0;
    for (int st = 0; st < 4; ++st)
        for (int blk = 0; blk < nblocks[st]; ++blk, bi += 4) {
            const int c1 = h->blockConv[bi], c2 = h->blockConv[bi + 1], c3 = h->blockConv[bi + 2], sc = h->blockConv[bi + 3];
            // pick three scratch buffers different from x
            __nv_bfloat16* t[3]; int n = 0;
            __nv_bfloat16* all[4] = {b->bufA, b->bufB, b->bufC, b->bufS};
            for (int k = 0; k < 4 && n < 3; ++k) if (all[k] != x) t[n++] = all[k];
            int H1, W1;
            if (run_conv(b, c1, x, H, W, t[0], nullptr, 1, s, &H1, &W1)) return -2;
            if (run_conv(b, c2, t[0], H1, W1, t[1], nullptr, 1, s)) return -2;
            const __nv_bfloat16* shortcut = x;
            if (sc >= 0) { if (run_conv(b, sc, x, H, W, t[2], nullptr, 0, s)) return -2; shortcut = t[2]; }
            const bool last = blk == nblocks[st] - 1;
            __nv_bfloat16* y = last ? stageOut[st] : t[0];
            if (run_conv(b, c3, t[1], H1, W1, y, shortcut, 1, s)) return -2;
            x = y; H = H1; W = W1;
            (void)pool;
        }
    // FPN
    const int fs[4] = {S / 4, S / 8, S / 16, S / 32};
    __nv_bfloat16* Cs[4] = {b->C2, b->C3, b->C4, b->C5};
    // P5 lateral
    if (run_conv(b, h->fpnLat[3], Cs[3], fs[3], fs[3], b->td, nullptr, 0, s)) return -2;
    __nv_bfloat16* top = b->td;                       // running top-down map (pre-3x3)
    __nv_bfloat16* tdBuf[2] = {b->bufA, b->bufB};
    if (run_conv(b, h->fpnOut[3], top, fs[3], fs[3], b->P[3], nullptr, 0, s)) return -2;
    for (int i = 2; i >= 0; --i) {
        if (run_conv(b, h->fpnLat[i], Cs[i], fs[i], fs[i], b->lat, nullptr, 0, s)) return -2;
        __nv_bfloat16* nt = tdBuf[i & 1];
        prof_mark(s, "k_upsample_add"); k_upsample_add<<<1184, 256, 0, s>>>(b->lat, top, fs[i], fs[i], 256, nt);
        top = nt;
        if (run_conv(b, h->fpnOut[i], top, fs[i], fs[i], b->P[i], nullptr, 0, s)) return -2;
    }
    prof_mark(s, "k_subsample2"); k_subsample2<<<64, 256, 0, s>>>(b->P[3], fs[3], fs[3], 256, b->P[4]);       // P6
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) { g_cnn_err = std::string("backbone forward: ") + cudaGetErrorString(e); return -3; }
    return 0;
}
extern "C" double mf_backbone_flops(mf_backbone* h) { return h ? h->b.flops : 0; }
extern "C" int mf_backbone_num_gemms(mf_backbone* h) { return h ? h->b.gemms : 0; }
extern "C" void* mf_backbone_input_buffer(mf_backbone* h) { return h ? h->b.input : nullptr; }
extern "C" void* mf_backbone_stream(mf_backbone* h) { return h ? (void*)h->stream : nullptr; }
// level 0..3 = C2..C5, 4..8 = P2..P6; returns device pointer, fills dims (H, W, C)
extern "C" void* mf_backbone_output(mf_backbone* h, int level, int* dims3)
{
    if (!h) return nullptr;
    Backbone* b = &h->b; const int S = b->S;
    const int fs[5] = {S / 4, S / 8, S / 16, S / 32, S / 64};
    const int cdim[4] = {256, 512, 1024, 2048};
    if (level >= 0 && level < 4) { dims3[0] = dims3[1] = fs[level]; dims3[2] = cdim[level]; __nv_bfloat16* c[4] = {b->C2, b->C3, b->C4, b->C5}; return c[level]; }
    if (level >= 4 && level < 9) { dims3[0] = dims3[1] = fs[level - 4]; dims3[2] = 256; return b->P[level - 4]; }
    return nullptr;
}
extern "C" int mf_backbone_download(mf_backbone* h, int level, void* host_bf16)
{
    int d[3];
    void* p = mf_backbone_output(h, level, d);
    if (!p) return -1;
    cudaStreamSynchronize(h->stream);
    return cudaMemcpy(host_bf16, p, (size_t)d[0] * d[1] * d[2] * 2, cudaMemcpyDeviceToHost) == cudaSuccess ? 0 : -2;
}
// letter-box + normalise a 640x480 (or any) RGBA8 device image into the network input (MaskRCNN.py.in mold_inputs)
extern "C" int mf_backbone_mold(mf_backbone* h, const void* d_rgba, int W, int H)
{
    if (!h) return -1;
    Backbone* b = &h->b; const int S = b->S;
    float scale = fminf((float)S / (float)W, (float)S / (float)H);
    int newW = (int)lroundf(W * scale), newH = (int)lroundf(H * scale);
    int offx = (S - newW) / 2, offy = (S - newH) / 2;
    prof_mark(h->stream, "k_mold_input");
    k_mold_input<<<1184, 256, 0, h->stream>>>((const uchar4*)d_rgba, W, H, S, scale, offx, offy, newW, newH, b->input);
    return cudaGetLastError() == cudaSuccess ? 0 : -2;
}
