/*
 * oracle/orc_maps.c -- CPU ORACLE (test infrastructure, not product):
 * restatement of Core/Cuda/cudafuncs.cu map / pyramid kernels and the depth
 * bilateral shader.  Build flags (oracle/Makefile): -O2 -ffp-contract=off.
 */
#include "orc.h"
#include <math.h>
#include <string.h>
#include <stdlib.h>

static inline float qnan(void) { union { uint32_t u; float f; } c; c.u = 0x7fffffffu; return c.f; } /* cudafuncs.cu:130 */
static inline int imin(int a, int b) { return a < b ? a : b; }
static inline int imax(int a, int b) { return a > b ? a : b; }

/* ---- deterministic exp / acos: plain + - * / sqrt only, so the CUDA build
 *      (compiled with -fmad=false) reproduces them bit for bit. GLSL leaves
 *      exp()/acos() precision implementation-defined; this is the written rule. */
float orc_expf(float x)
{
    if (!(x > -87.0f)) return (x != x) ? x : 0.0f;
    if (x > 88.0f) return INFINITY;
    float t = x * 1.44269504088896341f;
    float n = floorf(t + 0.5f);
    float r = (x - n * 0.693359375f) - n * (-2.12194440e-4f);
    /* degree-6 minimax on [-ln2/2, ln2/2] (Cephes expf coefficients) */
    float p = 1.9875691500e-4f;
    p = p * r + 1.3981999507e-3f;
    p = p * r + 8.3334519073e-3f;
    p = p * r + 4.1665795894e-2f;
    p = p * r + 1.6666665459e-1f;
    p = p * r + 5.0000001201e-1f;
    float r2 = r * r;
    float y = (p * r2 + r) + 1.0f;
    union { uint32_t u; float f; } s;
    s.u = (uint32_t)((int)n + 127) << 23;
    return y * s.f;
}

float orc_acosf(float x)
{
    float a = fabsf(x);
    if (!(a <= 1.0f)) return qnan();
    /* Abramowitz & Stegun 4.4.46 (|err| <= 2e-8) */
    float p = -0.0012624911f;
    p = p * a + 0.0066700901f;
    p = p * a + -0.0170881256f;
    p = p * a + 0.0308918810f;
    p = p * a + -0.0501743046f;
    p = p * a + 0.0889789874f;
    p = p * a + -0.2145988016f;
    p = p * a + 1.5707963050f;
    float r = sqrtf(1.0f - a) * p;
    return x < 0.0f ? 3.14159265358979f - r : r;
}

void orc_pose_inverse(const float* T, float* Ti)
{
    /* rigid inverse [R^T | -R^T t] in fp32 (the reference calls Eigen's general
     * 4x4 inverse, Model.cpp:668, ModelProjection.cpp:114; rule fixed in DESIGN.md) */
    float R[9] = { T[0], T[1], T[2], T[4], T[5], T[6], T[8], T[9], T[10] };
    float t[3] = { T[3], T[7], T[11] };
    for (int r = 0; r < 3; ++r) {
        for (int c = 0; c < 3; ++c) Ti[r * 4 + c] = R[c * 3 + r];
        Ti[r * 4 + 3] = -((R[0 * 3 + r] * t[0] + R[1 * 3 + r] * t[1]) + R[2 * 3 + r] * t[2]);
    }
    Ti[12] = 0; Ti[13] = 0; Ti[14] = 0; Ti[15] = 1;
}

void orc_pose_mul(const float* A, const float* B, float* C)
{
    float o[16];
    for (int r = 0; r < 4; ++r)
        for (int c = 0; c < 4; ++c) {
            float s = 0;
            for (int k = 0; k < 4; ++k) s += A[r * 4 + k] * B[k * 4 + c];
            o[r * 4 + c] = s;
        }
    memcpy(C, o, sizeof o);
}

orc_cam orc_cam_level(orc_cam c, int level)
{
    int div = 1 << level;                                   /* types.cuh:94-98 */
    orc_cam r = { c.fx / div, c.fy / div, c.cx / div, c.cy / div };
    return r;
}

/* depth_bilateral_metric.frag:30-76; N14: texel (cx,cy) sampled nearest */
void orc_bilateral(const float* depth, float* out, int W, int H)
{
    const float sigma_space2_inv_half = 0.024691358f;
    const float sigma_color2_inv_half = 555.556f;
    const int R = 6, D = R * 2 + 1;
#pragma omp parallel for schedule(static)
    for (int y = 0; y < H; ++y)
        for (int x = 0; x < W; ++x) {
            float value = depth[y * W + x];
            if (value <= 0.03f) { out[y * W + x] = 0; continue; }
            int tx = imin(x - D / 2 + D, W);
            int ty = imin(y - D / 2 + D, H);
            float sum1 = 0, sum2 = 0;
            for (int cy = imax(y - D / 2, 0); cy < ty; ++cy)
                for (int cx = imax(x - D / 2, 0); cx < tx; ++cx) {
                    float tmp = depth[cy * W + cx];
                    float dx = (float)x - (float)cx, dy = (float)y - (float)cy;
                    float space2 = dx * dx + dy * dy;
                    float dc = value - tmp;
                    float color2 = dc * dc;
                    float weight = orc_expf(-(space2 * sigma_space2_inv_half + color2 * sigma_color2_inv_half));
                    sum1 += tmp * weight;
                    sum2 += weight;
                }
            out[y * W + x] = sum1 / sum2;
        }
}

static const float GAUSS5[25] = { 1, 4, 6, 4, 1, 4, 16, 24, 16, 4, 6, 24, 36, 24, 6,
                                  4, 16, 24, 16, 4, 1, 4, 6, 4, 1 };   /* cudafuncs.cu:516-520 */

/* cudafuncs.cu:333-364 (N9: exclusive clamp cols-1, int weight sum) */
void orc_pyrdown_gauss_f(const float* src, int sw, int sh, float* dst)
{
    int dw = sw / 2, dh = sh / 2;
    const int D = 5;
    #pragma omp parallel for schedule(static)
    for (int y = 0; y < dh; ++y)
        for (int x = 0; x < dw; ++x) {
            int tx = imin(2 * x - D / 2 + D, sw - 1);
            int ty = imin(2 * y - D / 2 + D, sh - 1);
            float sum = 0; int count = 0;
            for (int cy = imax(0, 2 * y - D / 2); cy < ty; ++cy)
                for (int cx = imax(0, 2 * x - D / 2); cx < tx; ++cx) {
                    float v = src[cy * sw + cx];
                    if (!isnan(v)) {
                        float g = GAUSS5[(ty - cy - 1) * 5 + (tx - cx - 1)];
                        sum += v * g;
                        count = (int)((float)count + g);
                    }
                }
            dst[y * dw + x] = sum / (float)count;
        }
}

/* cudafuncs.cu:534-564; float->uchar store: truncation, NaN (0/0) -> 0 */
void orc_pyrdown_gauss_u8(const uint8_t* src, int sw, int sh, uint8_t* dst)
{
    int dw = sw / 2, dh = sh / 2;
    const int D = 5;
    #pragma omp parallel for schedule(static)
    for (int y = 0; y < dh; ++y)
        for (int x = 0; x < dw; ++x) {
            int tx = imin(2 * x - D / 2 + D, sw - 1);
            int ty = imin(2 * y - D / 2 + D, sh - 1);
            float sum = 0; int count = 0;
            for (int cy = imax(0, 2 * y - D / 2); cy < ty; ++cy)
                for (int cx = imax(0, 2 * x - D / 2); cx < tx; ++cx) {
                    uint8_t v = src[cy * sw + cx];
                    if (v > 0) {
                        float g = GAUSS5[(ty - cy - 1) * 5 + (tx - cx - 1)];
                        sum += (float)v * g;
                        count = (int)((float)count + g);
                    }
                }
            float r = sum / (float)count;
            dst[y * dw + x] = (r != r) ? 0 : (uint8_t)(int)r;
        }
}

/* cudafuncs.cu:109-134: integer pixel coords (N1); invalid => x=NaN, z=0.
 * The reference leaves y (and stale planes) untouched; the oracle writes 0. */
void orc_vmap(const float* depth, int W, int H, orc_cam cam, float cutoff, float* vmap)
{
    float fx_inv = 1.f / cam.fx, fy_inv = 1.f / cam.fy;
    #pragma omp parallel for schedule(static)
    for (int v = 0; v < H; ++v)
        for (int u = 0; u < W; ++u) {
            float z = depth[v * W + u];
            if (z > 0.0f && z < cutoff) {
                vmap[(0 * H + v) * W + u] = z * ((float)u - cam.cx) * fx_inv;
                vmap[(1 * H + v) * W + u] = z * ((float)v - cam.cy) * fy_inv;
                vmap[(2 * H + v) * W + u] = z;
            } else {
                vmap[(0 * H + v) * W + u] = qnan();
                vmap[(1 * H + v) * W + u] = 0;
                vmap[(2 * H + v) * W + u] = 0;
            }
        }
}

/* cudafuncs.cu:152-189; normalized() = v * rsqrtf(dot) (operators.cuh);
 * oracle rule: v / sqrtf(dot) (IEEE), tolerance documented for _ref compare */
void orc_nmap(const float* vmap, int W, int H, float* nmap)
{
    const float* X = vmap; const float* Y = vmap + H * W; const float* Z = vmap + 2 * H * W;
    #pragma omp parallel for schedule(static)
    for (int v = 0; v < H; ++v)
        for (int u = 0; u < W; ++u) {
            float* nx = &nmap[(0 * H + v) * W + u];
            float* ny = &nmap[(1 * H + v) * W + u];
            float* nz = &nmap[(2 * H + v) * W + u];
            *nx = qnan(); *ny = 0; *nz = 0;
            if (u == W - 1 || v == H - 1) continue;
            int i00 = v * W + u, i01 = v * W + u + 1, i10 = (v + 1) * W + u;
            if (isnan(X[i00]) || isnan(X[i01]) || isnan(X[i10])) continue;
            float ax = X[i01] - X[i00], ay = Y[i01] - Y[i00], az = Z[i01] - Z[i00];
            float bx = X[i10] - X[i00], by = Y[i10] - Y[i00], bz = Z[i10] - Z[i00];
            float cx = ay * bz - az * by;
            float cy = az * bx - ax * bz;
            float cz = ax * by - ay * bx;
            float len = sqrtf((cx * cx + cy * cy) + cz * cz);
            *nx = cx / len; *ny = cy / len; *nz = cz / len;
        }
}

/* cudafuncs.cu:271-311 */
void orc_copy_maps(const float* vt, const float* nt, int W, int H, float* vmap, float* nmap)
{
    #pragma omp parallel for schedule(static)
    for (int y = 0; y < H; ++y)
        for (int x = 0; x < W; ++x) {
            const float* vs = &vt[(y * W + x) * 4];
            const float* ns = &nt[(y * W + x) * 4];
            int ok = !(vs[2] == 0);
            for (int p = 0; p < 3; ++p) {
                vmap[(p * H + y) * W + x] = ok ? vs[p] : qnan();
                nmap[(p * H + y) * W + x] = ok ? ns[p] : qnan();
            }
        }
}

/* cudafuncs.cu:366-417 */
void orc_resize_map(const float* in, int sw, int sh, int normalize, float* out)
{
    int dw = sw / 2, dh = sh / 2;
    #pragma omp parallel for schedule(static)
    for (int y = 0; y < dh; ++y)
        for (int x = 0; x < dw; ++x) {
            int xs = 2 * x, ys = 2 * y;
            float n[3]; int bad = 0;
            for (int p = 0; p < 3; ++p) {
                const float* pl = in + p * sh * sw;
                float a = pl[ys * sw + xs], b = pl[ys * sw + xs + 1];
                float c = pl[(ys + 1) * sw + xs], d = pl[(ys + 1) * sw + xs + 1];
                if (p == 0 && (isnan(a) || isnan(b) || isnan(c) || isnan(d))) { bad = 1; break; }
                n[p] = (((a + b) + c) + d) / 4;
            }
            if (bad) {
                out[(0 * dh + y) * dw + x] = qnan();
                out[(1 * dh + y) * dw + x] = 0;
                out[(2 * dh + y) * dw + x] = 0;
                continue;
            }
            if (normalize) {
                float len = sqrtf((n[0] * n[0] + n[1] * n[1]) + n[2] * n[2]);
                n[0] /= len; n[1] /= len; n[2] /= len;
            }
            for (int p = 0; p < 3; ++p) out[(p * dh + y) * dw + x] = n[p];
        }
}

/* cudafuncs.cu:207-249 (in place, as RGBDOdometry.cpp:180-182 calls it) */
void orc_transform_maps(float* vmap, float* nmap, int W, int H, const float* R, const float* t)
{
    int P = W * H;
    #pragma omp parallel for schedule(static)
    for (int i = 0; i < P; ++i) {
        float x = vmap[i];
        if (!isnan(x)) {
            float y = vmap[P + i], z = vmap[2 * P + i];
            vmap[i]         = ((R[0] * x + R[1] * y) + R[2] * z) + t[0];
            vmap[P + i]     = ((R[3] * x + R[4] * y) + R[5] * z) + t[1];
            vmap[2 * P + i] = ((R[6] * x + R[7] * y) + R[8] * z) + t[2];
        }
        float nx = nmap[i];
        if (!isnan(nx)) {
            float ny = nmap[P + i], nz = nmap[2 * P + i];
            nmap[i]         = (R[0] * nx + R[1] * ny) + R[2] * nz;
            nmap[P + i]     = (R[3] * nx + R[4] * ny) + R[5] * nz;
            nmap[2 * P + i] = (R[6] * nx + R[7] * ny) + R[8] * nz;
        }
    }
}

/* cudafuncs.cu:602-613 */
void orc_vertices_to_depth(const float* vt, int W, int H, float cutoff, float* depth)
{
    #pragma omp parallel for schedule(static)
    for (int i = 0; i < W * H; ++i) {
        float z = vt[i * 4 + 2];
        depth[i] = (z > cutoff || z <= 0) ? qnan() : z;
    }
}

/* cudafuncs.cu:626-639: texel (x,y,z) = the three uploaded channels in order */
void orc_rgb_to_intensity(const uint8_t* rgb, int W, int H, uint8_t* out)
{
    #pragma omp parallel for schedule(static)
    for (int i = 0; i < W * H; ++i) {
        float v = ((float)rgb[i * 3 + 0] * 0.114f + (float)rgb[i * 3 + 1] * 0.299f) + (float)rgb[i * 3 + 2] * 0.587f;
        out[i] = (uint8_t)(int)v;
    }
}

/* cudafuncs.cu:658-683, taps :690-696; float->short store truncates */
void orc_sobel(const uint8_t* src, int W, int H, int16_t* dx, int16_t* dy)
{
    static const float gx[9] = { 0.52201f, 0.00000f, -0.52201f, 0.79451f, -0.00000f, -0.79451f, 0.52201f, 0.00000f, -0.52201f };
    static const float gy[9] = { 0.52201f, 0.79451f, 0.52201f, 0.00000f, 0.00000f, 0.00000f, -0.52201f, -0.79451f, -0.52201f };
    #pragma omp parallel for schedule(static)
    for (int y = 0; y < H; ++y)
        for (int x = 0; x < W; ++x) {
            float dxv = 0, dyv = 0; int k = 8;
            for (int j = imax(y - 1, 0); j <= imin(y + 1, H - 1); ++j)
                for (int i = imax(x - 1, 0); i <= imin(x + 1, W - 1); ++i) {
                    dxv += (float)src[j * W + i] * gx[k];
                    dyv += (float)src[j * W + i] * gy[k];
                    --k;
                }
            dx[y * W + x] = (int16_t)(int)dxv;
            dy[y * W + x] = (int16_t)(int)dyv;
        }
}

/* cudafuncs.cu:718-736 */
void orc_project_points(const float* depth, int W, int H, orc_cam cam, float* cloud)
{
    float ifx = 1.0f / cam.fx, ify = 1.0f / cam.fy;
    #pragma omp parallel for schedule(static)
    for (int y = 0; y < H; ++y)
        for (int x = 0; x < W; ++x) {
            float z = depth[y * W + x];
            cloud[(y * W + x) * 3 + 0] = ((float)x - cam.cx) * z * ifx;
            cloud[(y * W + x) * 3 + 1] = ((float)y - cam.cy) * z * ify;
            cloud[(y * W + x) * 3 + 2] = z;
        }
}
