// mf_seg.cu -- geometric depth-edge segmentation kernels (sm_100a)
//   edge-ness (concavity + distance) <- computeGeometricSegmentation_Kernel, Core/Cuda/segmentation.cu:122-177
//   threshold / invert               <- segmentation.cu:257-269
//   binary close                     <- dilate_Kernel / erode_Kernel, segmentation.cu:217-255, host :334-354
// Edge-ness and threshold are fused (one pass over the level-0 tracking maps, which is what
// the reference feeds it: MfSegmentation.cpp:149-151, REUSE_FILTERED_MAPS).
#include "mf_common.cuh"
#include "mf_kernels.h"

namespace mfb {

__global__ void k_geometric_edges(const float4* __restrict__ vmap, const float4* __restrict__ nmap, int W, int H, float wD, float wC,
                                  float thr, float* __restrict__ edge, uint8_t* __restrict__ binary)
{
    int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y;
    if (x >= W || y >= H) return;
    float e = 1.0f;
    if (!(x < 1 || x >= W - 1 || y < 1 || y >= H - 1)) {
        float4 v4 = vmap[y * W + x], n4 = nmap[y * W + x];
        float3 v = make_float3(v4.x, v4.y, v4.z), n = make_float3(n4.x, n4.y, n4.z);
        if (!(v.z <= 0.0f)) {
            float c = 0.0f, d = 0.0f;
            const int ox[8] = {-1, 0, 1, -1, 1, -1, 0, 1}, oy[8] = {-1, -1, -1, 0, 0, 1, 1, 1};
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                float4 a = vmap[(y + oy[k]) * W + x + ox[k]], b = nmap[(y + oy[k]) * W + x + ox[k]];
                float3 dd = make_float3(a.x - v.x, a.y - v.y, a.z - v.z);
                float dn = dot3(dd, n);
                float ct = (dn < 0) ? 0.0f : 1 - dot3(make_float3(b.x, b.y, b.z), n);
                c = fmaxf(ct, c);
                d = fmaxf(fabsf(dn), d);
            }
            c = fmaxf(c, 0.0f);
            c *= wC; d *= wD;
            e = fminf(1.0f, c > d ? c : d);
        }
    }
    edge[y * W + x] = e;
    binary[y * W + x] = e > thr ? 255 : 0;
}

__global__ void k_morph(const uint8_t* __restrict__ in, uint8_t* __restrict__ out, int W, int H, int r, int dilate)
{
    int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y;
    if (x >= W || y >= H) return;
    int x1 = max(x - r, 0), y1 = max(y - r, 0), x2 = min(x + r, W - 1), y2 = min(y + r, H - 1);
    uint8_t res = dilate ? 0 : 255;
    for (int cy = y1; cy <= y2; ++cy)
        for (int cx = x1; cx <= x2; ++cx) {
            if (cy == y && cx == x) continue;
            uint8_t v = in[cy * W + cx];
            if (dilate && v == 255) res = 255;
            if (!dilate && v == 0) res = 0;
        }
    out[y * W + x] = res;
}
__global__ void k_invert(const uint8_t* __restrict__ in, int n, uint8_t* __restrict__ out)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = (uint8_t)(255 - in[i]);
}

void launch_geometric_edges(const float4* vmap, const float4* nmap, int W, int H, float wD, float wC, float thr, float* edge, uint8_t* binary, cudaStream_t s)
{
    dim3 b(32, 8), g((W + 31) / 32, (H + 7) / 8);
    prof_mark(s, "k_geometric_edges"); k_geometric_edges<<<g, b, 0, s>>>(vmap, nmap, W, H, wD, wC, thr, edge, binary);
}
void launch_morph_close_invert(uint8_t* data, uint8_t* buf, int W, int H, int radius, int iterations, uint8_t* inverted, cudaStream_t s)
{
    dim3 b(32, 8), g((W + 31) / 32, (H + 7) / 8);
    for (int i = 0; i < iterations; ++i) {
        prof_mark(s, "k_morph"); k_morph<<<g, b, 0, s>>>(data, buf, W, H, radius, 1);
        prof_mark(s, "k_morph"); k_morph<<<g, b, 0, s>>>(buf, data, W, H, radius, 0);
    }
    prof_mark(s, "k_invert"); k_invert<<<(W * H + 255) / 256, 256, 0, s>>>(data, W * H, inverted);
}

}  // namespace mfb
