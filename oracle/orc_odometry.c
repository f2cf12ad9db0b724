/*
 * oracle/orc_odometry.c -- CPU ORACLE (test infrastructure, not product):
 * restatement of Core/Cuda/reduce.cu (icpKernel, residualKernel, rgbKernel,
 * so3Kernel) and of the host driver Core/Utils/RGBDOdometry.cpp +
 * OdometryProvider.h.  Reductions accumulate in double (order independent);
 * the reference accumulates fp32 in launch-config dependent order (N8), so
 * comparisons against it and against the CUDA build are tolerance based.
 */
#include "orc.h"
#include "orc_odom.h"
#include <math.h>
#include <float.h>
#include <string.h>
#include <stdlib.h>

static inline int rn(float v) { return (int)lrintf(v); }   /* __float2int_rn */

static inline void m3v(const float* R, const float* v, float* o)
{
    o[0] = (R[0] * v[0] + R[1] * v[1]) + R[2] * v[2];
    o[1] = (R[3] * v[0] + R[4] * v[1]) + R[5] * v[2];
    o[2] = (R[6] * v[0] + R[7] * v[1]) + R[8] * v[2];
}

static void accumulate(const float* row, int n, int found, double* out)
{
    int k = 0;
    for (int i = 0; i < n; ++i)
        for (int j = i; j < n + 1; ++j) out[k++] += (double)row[i] * (double)row[j];
    out[k++] += (double)row[n] * (double)row[n];
    out[k] += found ? 1.0 : 0.0;
}

/* Multi-threaded evaluation with the SEQUENTIAL sums kept bit for bit: the per-pixel rows (the expensive part: gathers, divisions,
 * square roots) are computed by all threads into a buffer of 8 floats per pixel (row[0..n], ..., flag in [7]); the accumulators are
 * then distributed over the threads, each one summing ITS accumulator over the pixels in the original order -- the same additions
 * in the same order as accumulate() in a single loop (a pixel that contributes nothing adds +0.0, the identity, and is skipped). */
static void accumulate_rows(const float* rows, int P, int n, double* out)
{
    int pi[32], pj[32], nq = 0;
    for (int i = 0; i < n; ++i)
        for (int j = i; j < n + 1; ++j) { pi[nq] = i; pj[nq] = j; ++nq; }
    pi[nq] = n; pj[nq] = n; ++nq;
#pragma omp parallel for schedule(static, 1)
    for (int q = 0; q <= nq; ++q) {
        double s = 0.0;
        if (q < nq) {
            const int a = pi[q], b = pj[q];
            /* the product of two floats is exact in double: the sum is the fp64 sum of the exact products (the reference's kernels
             * contract product and add into an fp32 FMA, i.e. also add the unrounded product).  An fp64 sum in ANY order agrees with
             * this one to ~1e-15 relative, so after the rounding to float that the reference's 29-float result record imposes, a
             * parallel fp64 reduction (the CUDA path) reproduces these values bit for bit except with probability ~1e-7 per value. */
            for (int p = 0; p < P; ++p) { const float* r = rows + (size_t)p * 8; if (r[7] != 0.f) s += (double)r[a] * (double)r[b]; }
        } else
            for (int p = 0; p < P; ++p) s += rows[(size_t)p * 8 + 7] != 0.f ? 1.0 : 0.0;
        out[q] = s;
    }
}

/* reduce.cu:259-444 */
void orc_icp_step(const float* Rcurr, const float* tcurr, const float* vmap_curr, const float* nmap_curr,
                  const float* Rprev_inv, const float* tprev, orc_cam cam,
                  const float* vmap_g_prev, const float* nmap_g_prev,
                  float distThres, float angleThres, int W, int H, double* out29)
{
    int P = W * H;
    float* rows = (float*)malloc((size_t)P * 8 * sizeof(float));
#pragma omp parallel for schedule(static)
    for (int i = 0; i < P; ++i) {
        float row[7] = { 0, 0, 0, 0, 0, 0, 0 };
        int found = 0;
        float vcurr[3] = { vmap_curr[i], vmap_curr[P + i], vmap_curr[2 * P + i] };
        float vg[3], tmp[3], vcp[3];
        m3v(Rcurr, vcurr, vg);
        vg[0] += tcurr[0]; vg[1] += tcurr[1]; vg[2] += tcurr[2];
        tmp[0] = vg[0] - tprev[0]; tmp[1] = vg[1] - tprev[1]; tmp[2] = vg[2] - tprev[2];
        m3v(Rprev_inv, tmp, vcp);
        float px = vcp[0] * cam.fx / vcp[2] + cam.cx;
        float py = vcp[1] * cam.fy / vcp[2] + cam.cy;
        /* NaN / out-of-int-range projections: __float2int_rn(NaN)=0 on the GPU;
         * such pixels are rejected below by the isnan(ncurr.x) test (N8). */
        int ux = (px != px) ? 0 : (px > 1e9f ? 1000000000 : (px < -1e9f ? -1000000000 : rn(px)));
        int uy = (py != py) ? 0 : (py > 1e9f ? 1000000000 : (py < -1e9f ? -1000000000 : rn(py)));
        if (!(ux < 0 || uy < 0 || ux >= W || uy >= H || vcp[2] < 0)) {
            int j = uy * W + ux;
            float vp[3] = { vmap_g_prev[j], vmap_g_prev[P + j], vmap_g_prev[2 * P + j] };
            float ncurr[3] = { nmap_curr[i], nmap_curr[P + i], nmap_curr[2 * P + i] };
            float ng[3];
            m3v(Rcurr, ncurr, ng);
            float np[3] = { nmap_g_prev[j], nmap_g_prev[P + j], nmap_g_prev[2 * P + j] };
            float d[3] = { vp[0] - vg[0], vp[1] - vg[1], vp[2] - vg[2] };
            float dist = sqrtf((d[0] * d[0] + d[1] * d[1]) + d[2] * d[2]);
            float c[3] = { ng[1] * np[2] - ng[2] * np[1], ng[2] * np[0] - ng[0] * np[2], ng[0] * np[1] - ng[1] * np[0] };
            float sine = sqrtf((c[0] * c[0] + c[1] * c[1]) + c[2] * c[2]);
            found = (sine < angleThres && dist <= distThres && !isnan(ncurr[0]) && !isnan(np[0]));
            if (found) {
                float s_cp[3], d_cp[3], n_cp[3], t2[3];
                m3v(Rprev_inv, tmp, s_cp);
                t2[0] = vp[0] - tprev[0]; t2[1] = vp[1] - tprev[1]; t2[2] = vp[2] - tprev[2];
                m3v(Rprev_inv, t2, d_cp);
                m3v(Rprev_inv, np, n_cp);
                row[0] = n_cp[0]; row[1] = n_cp[1]; row[2] = n_cp[2];
                row[3] = s_cp[1] * n_cp[2] - s_cp[2] * n_cp[1];
                row[4] = s_cp[2] * n_cp[0] - s_cp[0] * n_cp[2];
                row[5] = s_cp[0] * n_cp[1] - s_cp[1] * n_cp[0];
                row[6] = (n_cp[0] * (s_cp[0] - d_cp[0]) + n_cp[1] * (s_cp[1] - d_cp[1])) + n_cp[2] * (s_cp[2] - d_cp[2]);
            }
        }
        float* r = rows + (size_t)i * 8;
        for (int k = 0; k < 7; ++k) r[k] = row[k];
        r[7] = found ? 1.f : 0.f;
    }
    accumulate_rows(rows, P, 6, out29);
    free(rows);
}

/* reduce.cu:774-997 */
void orc_rgb_residual(float minScale, const int16_t* dIdx, const int16_t* dIdy,
                      const float* lastDepth, const float* nextDepth,
                      const uint8_t* lastImage, const uint8_t* nextImage,
                      orc_dataterm* corresImg, float maxDepthDelta, const float* kt,
                      const float* K, int W, int H, int* count, int* sigmaSum)
{
    int cnt = 0, sig = 0;                 /* integer sums: any order gives the same totals */
#pragma omp parallel for schedule(static) reduction(+ : cnt, sig)
    for (int k = 0; k < W * H; ++k) {
        int i = k / W, j0 = k - i * W;
        orc_dataterm c; memset(&c, 0, sizeof c);
        if (j0 < W - 5 && i < H - 1) {
            int valid = 1;
            for (int u = (i - 2 > 0 ? i - 2 : 0); u < (i + 2 < H ? i + 2 : H); ++u)
                for (int v = (j0 - 2 > 0 ? j0 - 2 : 0); v < (j0 + 2 < W ? j0 + 2 : W); ++v)
                    valid = valid && (nextImage[u * W + v] > 0);
            if (valid) {
                int valx = dIdx[k], valy = dIdy[k];
                float mTwo = (float)((valx * valx) + (valy * valy));
                if (mTwo >= minScale) {
                    int y = i, x = j0;
                    float d1 = nextDepth[k];
                    if (!isnan(d1)) {
                        float td1 = d1 * ((K[6] * x + K[7] * y) + K[8]) + kt[2];
                        float fu = (d1 * ((K[0] * x + K[1] * y) + K[2]) + kt[0]) / td1;
                        float fv = (d1 * ((K[3] * x + K[4] * y) + K[5]) + kt[1]) / td1;
                        int u0 = (fu != fu || fabsf(fu) > 1e9f) ? -1 : rn(fu);
                        int v0 = (fv != fv || fabsf(fv) > 1e9f) ? -1 : rn(fv);
                        if (u0 >= 0 && v0 >= 0 && u0 < W && v0 < H) {
                            float d0 = lastDepth[v0 * W + u0];
                            if (d0 > 0 && fabsf(td1 - d0) <= maxDepthDelta && lastImage[v0 * W + u0] != 0) {
                                c.zx = (int16_t)u0; c.zy = (int16_t)v0; c.ox = (int16_t)x; c.oy = (int16_t)y;
                                c.diff = (float)nextImage[k] - (float)lastImage[v0 * W + u0];
                                c.valid = 1;
                                cnt += 1;
                                sig += (int)(c.diff * c.diff);
                            }
                        }
                    }
                }
            }
        }
        corresImg[k] = c;
    }
    *count = cnt; *sigmaSum = sig;
}

/* reduce.cu:529-713 */
void orc_rgb_step(const orc_dataterm* corres, float sigma, const float* cloud, float fx, float fy,
                  const int16_t* dIdx, const int16_t* dIdy, float sobelScale, int W, int H, double* out29)
{
    const int P = W * H;
    float* rows = (float*)malloc((size_t)P * 8 * sizeof(float));
#pragma omp parallel for schedule(static)
    for (int i = 0; i < P; ++i) {
        const orc_dataterm* c = &corres[i];
        float row[7] = { 0, 0, 0, 0, 0, 0, 0 };
        if (c->valid) {
            float w = sigma + fabsf(c->diff);
            w = w > 1.19209290E-07F ? 1.0f / w : 1.0f;
            if (sigma == -1) w = 1;
            row[6] = -w * c->diff;
            const float* cp = &cloud[(c->zy * W + c->zx) * 3];
            float invz = (float)(1.0 / (double)cp[2]);
            float dIdx_v = w * sobelScale * (float)dIdx[c->oy * W + c->ox];
            float dIdy_v = w * sobelScale * (float)dIdy[c->oy * W + c->ox];
            float v0 = dIdx_v * fx * invz;
            float v1 = dIdy_v * fy * invz;
            float v2 = -(v0 * cp[0] + v1 * cp[1]) * invz;
            row[0] = v0; row[1] = v1; row[2] = v2;
            row[3] = -cp[2] * v1 + cp[1] * v2;
            row[4] = cp[2] * v0 - cp[0] * v2;
            row[5] = -cp[1] * v0 + cp[0] * v1;
        }
        float* r = rows + (size_t)i * 8;
        for (int k = 0; k < 7; ++k) r[k] = row[k];
        r[7] = c->valid ? 1.f : 0.f;
    }
    accumulate_rows(rows, P, 6, out29);
    free(rows);
}

/* reduce.cu:999-1202 */
static void grad_u8(const uint8_t* img, int W, int x, int y, float* gx, float* gy)
{
    float actu = img[y * W + x], back = img[y * W + x - 1], fore = img[y * W + x + 1];
    *gx = ((back + actu) / 2.0f) - ((fore + actu) / 2.0f);
    back = img[(y - 1) * W + x]; fore = img[(y + 1) * W + x];
    *gy = ((back + actu) / 2.0f) - ((fore + actu) / 2.0f);
}

void orc_so3_step(const uint8_t* lastImage, const uint8_t* nextImage, const float* B, const float* kinv,
                  const float* krlr, int W, int H, double* out11)
{
    const int P = W * H;
    float* rows = (float*)malloc((size_t)P * 8 * sizeof(float));
#pragma omp parallel for schedule(static)
    for (int k = 0; k < P; ++k) {
        int y = k / W, x = k - y * W;
        float ur[3] = { (float)x, (float)y, 1.0f }, wr[3];
        m3v(B, ur, wr);
        float fx_ = wr[0] / wr[2], fy_ = wr[1] / wr[2];
        int wx = (fx_ != fx_ || fabsf(fx_) > 1e9f) ? -1 : rn(fx_);
        int wy = (fy_ != fy_ || fabsf(fy_) > 1e9f) ? -1 : rn(fy_);
        int found = (wx >= 1 && wx < W - 1 && wy >= 1 && wy < H - 1 && x >= 1 && x < W - 1 && y >= 1 && y < H - 1);
        float row[4] = { 0, 0, 0, 0 };
        if (found) {
            float gnx, gny, glx, gly;
            grad_u8(nextImage, W, wx, wy, &gnx, &gny);
            grad_u8(lastImage, W, x, y, &glx, &gly);
            float gx = (gnx + glx) / 2.0f, gy = (gny + gly) / 2.0f;
            float p[3];
            m3v(kinv, ur, p);
            float z2 = p[2] * p[2];
            float a = krlr[0], b = krlr[1], c = krlr[2], d = krlr[3], e = krlr[4], f = krlr[5], g = krlr[6], h = krlr[7], i = krlr[8];
            float fy = (float)y, fxx = (float)x;
            float l[3] = { ((p[2] * (d * gy + a * gx)) - (gy * g * fy) - (gx * g * fxx)) / z2,
                           ((p[2] * (e * gy + b * gx)) - (gy * h * fy) - (gx * h * fxx)) / z2,
                           ((p[2] * (f * gy + c * gx)) - (gy * i * fy) - (gx * i * fxx)) / z2 };
            row[0] = l[1] * p[2] - l[2] * p[1];
            row[1] = l[2] * p[0] - l[0] * p[2];
            row[2] = l[0] * p[1] - l[1] * p[0];
            row[3] = -((float)nextImage[wy * W + wx] - (float)lastImage[k]);
        }
        float* r = rows + (size_t)k * 8;
        for (int q = 0; q < 4; ++q) r[q] = row[q];
        r[4] = r[5] = r[6] = 0.f;
        r[7] = found ? 1.f : 0.f;
    }
    accumulate_rows(rows, P, 3, out11);
    free(rows);
}

/* ------------------------------------------------------------------------- */
/* Host maths                                                                */
/* ------------------------------------------------------------------------- */

/* Eigen LDLT semantics (pivoting on the largest diagonal, pivots at or below
 * DBL_MIN solve to 0): RGBDOdometry.cpp:313,451-459 */
void orc_ldlt_solve(const double* Ain, const double* b, int n, double* x)
{
    double A[36], y[6]; int perm[6];
    for (int i = 0; i < n * n; ++i) A[i] = Ain[i];
    for (int i = 0; i < n; ++i) perm[i] = i;
    int kend = n;
    for (int k = 0; k < n; ++k) {
        int piv = k; double best = fabs(A[k * n + k]);
        for (int i = k + 1; i < n; ++i) if (fabs(A[i * n + i]) > best) { best = fabs(A[i * n + i]); piv = i; }
        if (piv != k) {
            for (int j = 0; j < n; ++j) { double t = A[k * n + j]; A[k * n + j] = A[piv * n + j]; A[piv * n + j] = t; }
            for (int j = 0; j < n; ++j) { double t = A[j * n + k]; A[j * n + k] = A[j * n + piv]; A[j * n + piv] = t; }
            int t = perm[k]; perm[k] = perm[piv]; perm[piv] = t;
        }
        double d = A[k * n + k];
        if (fabs(d) <= DBL_MIN) { kend = k; break; }
        for (int i = k + 1; i < n; ++i) A[i * n + k] /= d;
        for (int i = k + 1; i < n; ++i)
            for (int j = k + 1; j <= i; ++j) {
                A[i * n + j] -= A[i * n + k] * d * A[j * n + k];
                A[j * n + i] = A[i * n + j];
            }
    }
    for (int i = kend; i < n; ++i) { for (int j = 0; j < i; ++j) if (j >= kend) A[i * n + j] = 0; }
    for (int i = 0; i < n; ++i) y[i] = b[perm[i]];
    for (int i = 0; i < n; ++i) for (int j = 0; j < i && j < kend; ++j) y[i] -= A[i * n + j] * y[j];
    for (int i = 0; i < n; ++i) {
        double d = (i < kend) ? A[i * n + i] : 0.0;
        y[i] = (fabs(d) > DBL_MIN) ? y[i] / d : 0.0;
    }
    for (int i = n - 1; i >= 0; --i) for (int j = i + 1; j < n; ++j) if (i < kend) y[i] -= A[j * n + i] * y[j];
    for (int i = 0; i < n; ++i) x[perm[i]] = y[i];
}

/* R-SINCOS: libm's and CUDA's double sin/cos are both within 1 ulp but not bit-identical; the update angles of the
 * Gauss-Newton steps are tiny, so both sides evaluate one fixed Taylor polynomial (Horner, no contraction) for |x| <= 0.8 and
 * fall back to the library beyond (never reached by a converging tracker; a mismatch there would be a 1-ulp effect). */
void orc_det_sincos(double x, double* s, double* c)
{
    if (!(fabs(x) <= 0.8)) { *s = sin(x); *c = cos(x); return; }
    const double z = x * x;
    double ps = -1.0 / 51090942171709440000.0;                    /* -1/21! */
    ps = ps * z + 1.0 / 121645100408832000.0;                      /*  1/19! */
    ps = ps * z - 1.0 / 355687428096000.0;                         /* -1/17! */
    ps = ps * z + 1.0 / 1307674368000.0;                           /*  1/15! */
    ps = ps * z - 1.0 / 6227020800.0;                              /* -1/13! */
    ps = ps * z + 1.0 / 39916800.0;                                /*  1/11! */
    ps = ps * z - 1.0 / 362880.0;                                  /* -1/9!  */
    ps = ps * z + 1.0 / 5040.0;                                    /*  1/7!  */
    ps = ps * z - 1.0 / 120.0;                                     /* -1/5!  */
    ps = ps * z + 1.0 / 6.0;                                       /*  1/3!  */
    *s = x - x * (z * ps);
    double pc = 1.0 / 2432902008176640000.0;                       /*  1/20! */
    pc = pc * z - 1.0 / 6402373705728000.0;                        /* -1/18! */
    pc = pc * z + 1.0 / 20922789888000.0;                          /*  1/16! */
    pc = pc * z - 1.0 / 87178291200.0;                             /* -1/14! */
    pc = pc * z + 1.0 / 479001600.0;                               /*  1/12! */
    pc = pc * z - 1.0 / 3628800.0;                                 /* -1/10! */
    pc = pc * z + 1.0 / 40320.0;                                   /*  1/8!  */
    pc = pc * z - 1.0 / 720.0;                                     /* -1/6!  */
    pc = pc * z + 1.0 / 24.0;                                      /*  1/4!  */
    pc = pc * z - 1.0 / 2.0;                                       /* -1/2!  */
    *c = 1.0 + z * pc;
}

/* OdometryProvider.h:32-66 */
void orc_rodrigues(const double* src, double* R)
{
    double rx = src[0], ry = src[1], rz = src[2];
    double theta = sqrt(rx * rx + ry * ry + rz * rz);
    for (int k = 0; k < 9; ++k) R[k] = (k % 4 == 0) ? 1.0 : 0.0;
    if (theta >= DBL_EPSILON) {
        const double I[9] = { 1, 0, 0, 0, 1, 0, 0, 0, 1 };
        double c, s; orc_det_sincos(theta, &s, &c);
        double c1 = 1. - c, it = theta ? 1. / theta : 0.;
        rx *= it; ry *= it; rz *= it;
        double rrt[9] = { rx * rx, rx * ry, rx * rz, rx * ry, ry * ry, ry * rz, rx * rz, ry * rz, rz * rz };
        double rxm[9] = { 0, -rz, ry, rz, 0, -rx, -ry, rx, 0 };
        for (int k = 0; k < 9; ++k) R[k] = c * I[k] + c1 * rrt[k] + s * rxm[k];
    }
}

static void inv3d(const double* M, double* o)
{
    double c00 = M[4] * M[8] - M[5] * M[7], c01 = M[5] * M[6] - M[3] * M[8], c02 = M[3] * M[7] - M[4] * M[6];
    double det = M[0] * c00 + M[1] * c01 + M[2] * c02, id = 1.0 / det;
    o[0] = c00 * id; o[1] = (M[2] * M[7] - M[1] * M[8]) * id; o[2] = (M[1] * M[5] - M[2] * M[4]) * id;
    o[3] = c01 * id; o[4] = (M[0] * M[8] - M[2] * M[6]) * id; o[5] = (M[2] * M[3] - M[0] * M[5]) * id;
    o[6] = c02 * id; o[7] = (M[1] * M[6] - M[0] * M[7]) * id; o[8] = (M[0] * M[4] - M[1] * M[3]) * id;
}
static void mul3d(const double* A, const double* B, double* C)
{
    double o[9];
    for (int r = 0; r < 3; ++r) for (int c = 0; c < 3; ++c) o[r * 3 + c] = A[r * 3] * B[c] + A[r * 3 + 1] * B[3 + c] + A[r * 3 + 2] * B[6 + c];
    memcpy(C, o, sizeof o);
}
static void inv3f(const float* M, float* o)
{
    float c00 = M[4] * M[8] - M[5] * M[7], c01 = M[5] * M[6] - M[3] * M[8], c02 = M[3] * M[7] - M[4] * M[6];
    float det = (M[0] * c00 + M[1] * c01) + M[2] * c02, id = 1.0f / det;
    o[0] = c00 * id; o[1] = (M[2] * M[7] - M[1] * M[8]) * id; o[2] = (M[1] * M[5] - M[2] * M[4]) * id;
    o[3] = c01 * id; o[4] = (M[0] * M[8] - M[2] * M[6]) * id; o[5] = (M[2] * M[3] - M[0] * M[5]) * id;
    o[6] = c02 * id; o[7] = (M[1] * M[6] - M[0] * M[7]) * id; o[8] = (M[0] * M[4] - M[1] * M[3]) * id;
}

/* ------------------------------------------------------------------------- */
/* RGBDOdometry object                                                       */
/* ------------------------------------------------------------------------- */
orc_odom* orc_odom_create(int W, int H, orc_cam cam)
{
    orc_odom* o = (orc_odom*)calloc(1, sizeof(orc_odom));
    o->W = W; o->H = H; o->cam = cam;
    o->vtex_tmp = (float*)calloc((size_t)W * H * 4, sizeof(float));
    for (int l = 0; l < 3; ++l) {
        size_t P = (size_t)(W >> l) * (H >> l);
        o->vmap_g[l] = (float*)calloc(P * 3, sizeof(float));
        o->nmap_g[l] = (float*)calloc(P * 3, sizeof(float));
        o->lastDepth[l] = (float*)calloc(P, sizeof(float));
        o->nextDepth[l] = (float*)calloc(P, sizeof(float));
        o->lastImage[l] = (uint8_t*)calloc(P, 1);
        o->nextImage[l] = (uint8_t*)calloc(P, 1);
        o->lastNextImage[l] = (uint8_t*)calloc(P, 1);
        o->dIdx[l] = (int16_t*)calloc(P, sizeof(int16_t));
        o->dIdy[l] = (int16_t*)calloc(P, sizeof(int16_t));
        o->cloud[l] = (float*)calloc(P * 3, sizeof(float));
        o->corres[l] = (orc_dataterm*)calloc(P, sizeof(orc_dataterm));
    }
    return o;
}

void orc_odom_destroy(orc_odom* o)
{
    if (!o) return;
    free(o->vtex_tmp);
    for (int l = 0; l < 3; ++l) {
        free(o->vmap_g[l]); free(o->nmap_g[l]); free(o->lastDepth[l]); free(o->nextDepth[l]);
        free(o->lastImage[l]); free(o->nextImage[l]); free(o->lastNextImage[l]);
        free(o->dIdx[l]); free(o->dIdy[l]); free(o->cloud[l]); free(o->corres[l]);
    }
    free(o);
}

static void intensity_from(const uint8_t* img, int stride, int W, int H, uint8_t* out)
{
    for (int i = 0; i < W * H; ++i) {
        const uint8_t* p = img + (size_t)i * stride;
        float v = ((float)p[0] * 0.114f + (float)p[1] * 0.299f) + (float)p[2] * 0.587f;
        out[i] = (uint8_t)(int)v;
    }
}

/* RGBDOdometry.cpp:217-225 */
void orc_odom_init_first_rgb(orc_odom* o, const uint8_t* rgb3)
{
    intensity_from(rgb3, 3, o->W, o->H, o->lastNextImage[0]);
    for (int l = 0; l + 1 < 3; ++l) orc_pyrdown_gauss_u8(o->lastNextImage[l], o->W >> l, o->H >> l, o->lastNextImage[l + 1]);
}

/* RGBDOdometry.cpp:153-185 */
void orc_odom_init_icp_model(orc_odom* o, const float* vtex4, const float* ntex4, const float* pose)
{
    int W = o->W, H = o->H;
    memcpy(o->vtex_tmp, vtex4, (size_t)W * H * 4 * sizeof(float));
    orc_copy_maps(vtex4, ntex4, W, H, o->vmap_g[0], o->nmap_g[0]);
    for (int l = 1; l < 3; ++l) {
        orc_resize_map(o->vmap_g[l - 1], W >> (l - 1), H >> (l - 1), 0, o->vmap_g[l]);
        orc_resize_map(o->nmap_g[l - 1], W >> (l - 1), H >> (l - 1), 1, o->nmap_g[l]);
    }
    float R[9] = { pose[0], pose[1], pose[2], pose[4], pose[5], pose[6], pose[8], pose[9], pose[10] };
    float t[3] = { pose[3], pose[7], pose[11] };
    for (int l = 0; l < 3; ++l) orc_transform_maps(o->vmap_g[l], o->nmap_g[l], W >> l, H >> l, R, t);
}

/* RGBDOdometry.cpp:187-215 (populateRGBDData; both calls read vmaps_tmp, N/A masks: N10) */
static void populate(orc_odom* o, const uint8_t* img, int stride, float** depths, uint8_t** images)
{
    int W = o->W, H = o->H;
    orc_vertices_to_depth(o->vtex_tmp, W, H, 6.0f /* maxDepthRGB, RGBDOdometry.cpp:34 */, depths[0]);
    for (int l = 0; l + 1 < 3; ++l) orc_pyrdown_gauss_f(depths[l], W >> l, H >> l, depths[l + 1]);
    intensity_from(img, stride, W, H, images[0]);
    for (int l = 0; l + 1 < 3; ++l) orc_pyrdown_gauss_u8(images[l], W >> l, H >> l, images[l + 1]);
}
void orc_odom_init_rgb_model(orc_odom* o, const uint8_t* image4) { populate(o, image4, 4, o->lastDepth, o->lastImage); }
void orc_odom_init_rgb(orc_odom* o, const uint8_t* rgb3) { populate(o, rgb3, 3, o->nextDepth, o->nextImage); }

/* Sensitivity probe (tests only, off by default): scale every reduced sum by (1 + eps * u), u in [-1, 1] from a counter hash.
 * eps = 1e-7 is the size of an fp32 tree-reduction's rounding, i.e. what separates two correct implementations of the same
 * sums; tests/test_cpu.py uses it to document how far such noise moves a tracked pose (object models: near-singular systems). */
static double g_sum_perturb = 0.0;
static unsigned g_sum_counter = 0;
void orc_debug_set_sum_perturb(double eps) { g_sum_perturb = eps; g_sum_counter = 0; }
static double perturbed(double v)
{
    if (g_sum_perturb == 0.0) return v;
    unsigned h = ++g_sum_counter * 2654435761u; h ^= h >> 15; h *= 2246822519u; h ^= h >> 13;
    return v * (1.0 + g_sum_perturb * ((double)(h & 0xffffff) / 8388607.5 - 1.0));
}

static void unpack29(const double* s, float* A, float* b)
{
    int k = 0;
    for (int i = 0; i < 6; ++i)
        for (int j = i; j < 7; ++j) {
            float v = (float)perturbed(s[k++]);
            if (j == 6) b[i] = v; else A[j * 6 + i] = A[i * 6 + j] = v;
        }
}

/* RGBDOdometry.cpp:227-497.  pose (row-major 4x4) is updated in place; the
 * incremental transform is returned in transformOut. vmaps/nmaps: frame
 * pyramids (planar). */
void orc_odom_track(orc_odom* o, float* const* frame_vmaps, float* const* frame_nmaps,
                    const orc_track_params* prm, float* pose, float* transformOut)
{
    const int W = o->W, H = o->H;
    int icp = !prm->rgbOnly && prm->icpWeight > 0;
    int rgb = prm->rgbOnly || prm->icpWeight < 100;
    float Rprev[9] = { pose[0], pose[1], pose[2], pose[4], pose[5], pose[6], pose[8], pose[9], pose[10] };
    float tprev[3] = { pose[3], pose[7], pose[11] };
    float Rcurr[9], tcurr[3];
    memcpy(Rcurr, Rprev, sizeof Rprev); memcpy(tcurr, tprev, sizeof tprev);

    if (rgb) for (int l = 0; l < 3; ++l) orc_sobel(o->nextImage[l], W >> l, H >> l, o->dIdx[l], o->dIdy[l]);

    double resultR[9] = { 1, 0, 0, 0, 1, 0, 0, 0, 1 };
    if (prm->so3) {
        int lv = 2; orc_cam c = orc_cam_level(o->cam, lv);
        float R_lr[9] = { 1, 0, 0, 0, 1, 0, 0, 0, 1 };
        double K[9] = { c.fx, 0, c.cx, 0, c.fy, c.cy, 0, 0, 1 }, Kinv[9];
        inv3d(K, Kinv);
        float lastError = FLT_MAX / 2, lastCount = FLT_MAX / 2;
        double lastResultR[9]; memcpy(lastResultR, resultR, sizeof resultR);
        for (int i = 0; i < 10; ++i) {
            double hom[9], krlr[9]; float homf[9], kinvf[9], krlrf[9];
            mul3d(K, resultR, krlr); mul3d(krlr, Kinv, hom);
            for (int k = 0; k < 9; ++k) { homf[k] = (float)hom[k]; kinvf[k] = (float)Kinv[k]; krlrf[k] = (float)krlr[k]; }
            double s[11];
            orc_so3_step(o->lastNextImage[lv], o->nextImage[lv], homf, kinvf, krlrf, W >> lv, H >> lv, s);
            float res0 = (float)s[9], res1 = (float)s[10];
            o->lastSO3Error = sqrtf(res0) / res1; o->lastSO3Count = res1;
            if (o->lastSO3Error < lastError && fabsf(lastError - o->lastSO3Count) < 0.001f) break;
            else if (o->lastSO3Error > lastError + 0.001f) {
                o->lastSO3Error = lastError; o->lastSO3Count = lastCount;
                memcpy(resultR, lastResultR, sizeof resultR); break;
            }
            lastError = o->lastSO3Error; lastCount = o->lastSO3Count; memcpy(lastResultR, resultR, sizeof resultR);
            float jtj[9], jtr[3]; int k = 0;
            for (int a = 0; a < 3; ++a) for (int b2 = a; b2 < 4; ++b2) { float v = (float)s[k++]; if (b2 == 3) jtr[a] = v; else jtj[b2 * 3 + a] = jtj[a * 3 + b2] = v; }
            double Ad[9], bd[3], delta[3];
            for (int q = 0; q < 9; ++q) Ad[q] = jtj[q];
            for (int q = 0; q < 3; ++q) bd[q] = jtr[q];
            orc_ldlt_solve(Ad, bd, 3, delta);
            for (int q = 0; q < 3; ++q) delta[q] = (double)(float)delta[q];
            double ru[9]; orc_rodrigues(delta, ru);
            float ruf[9], n[9];
            for (int q = 0; q < 9; ++q) ruf[q] = (float)ru[q];
            for (int r = 0; r < 3; ++r) for (int cc = 0; cc < 3; ++cc) n[r * 3 + cc] = (ruf[r * 3] * R_lr[cc] + ruf[r * 3 + 1] * R_lr[3 + cc]) + ruf[r * 3 + 2] * R_lr[6 + cc];
            memcpy(R_lr, n, sizeof n);
            for (int q = 0; q < 9; ++q) resultR[q] = R_lr[q];
        }
    }

    int iterations[3] = { prm->fastOdom ? 3 : 10, prm->pyramid ? 5 : 0, prm->pyramid ? 4 : 0 };
    float Rprev_inv[9]; inv3f(Rprev, Rprev_inv);
    double resultRt[16] = { 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1 };
    if (prm->so3) for (int r = 0; r < 3; ++r) for (int c = 0; c < 3; ++c) resultRt[r * 4 + c] = resultR[r * 3 + c];
    float trR[9] = { 1, 0, 0, 0, 1, 0, 0, 0, 1 }, trT[3] = { 0, 0, 0 };   /* 'transform' (Isometry3f) */
    const float sobelScale = (float)(1.0 / pow(2.0, 3));                   /* RGBDOdometry.cpp:31-32 */
    const float minGrad[3] = { 5, 3, 1 };

    for (int l = 2; l >= 0; --l) {
        int w = W >> l, h = H >> l; orc_cam c = orc_cam_level(o->cam, l);
        if (rgb) orc_project_points(o->lastDepth[l], w, h, c, o->cloud[l]);
        double K[9] = { c.fx, 0, c.cx, 0, c.fy, c.cy, 0, 0, 1 }, Kinv[9];
        inv3d(K, Kinv);
        o->lastRGBError = FLT_MAX;
        for (int j = 0; j < iterations[l]; ++j) {
            /* Rt = resultRt.inverse() */
            double R3[9] = { resultRt[0], resultRt[1], resultRt[2], resultRt[4], resultRt[5], resultRt[6], resultRt[8], resultRt[9], resultRt[10] };
            double Ri[9]; inv3d(R3, Ri);
            double ti[3];
            for (int r = 0; r < 3; ++r) ti[r] = -(Ri[r * 3] * resultRt[3] + Ri[r * 3 + 1] * resultRt[7] + Ri[r * 3 + 2] * resultRt[11]);
            double KRK[9], tmpm[9]; mul3d(K, Ri, tmpm); mul3d(tmpm, Kinv, KRK);
            float krk[9]; for (int q = 0; q < 9; ++q) krk[q] = (float)KRK[q];
            double Kt[3];
            for (int r = 0; r < 3; ++r) Kt[r] = K[r * 3] * ti[0] + K[r * 3 + 1] * ti[1] + K[r * 3 + 2] * ti[2];
            float kt[3] = { (float)Kt[0], (float)Kt[1], (float)Kt[2] };
            int sigma = 0, rgbSize = 0;
            if (rgb) {
                float minScale = (float)(pow(minGrad[l], 2.0) / pow((double)sobelScale, 2.0));
                orc_rgb_residual(minScale, o->dIdx[l], o->dIdy[l], o->lastDepth[l], o->nextDepth[l], o->lastImage[l],
                                 o->nextImage[l], o->corres[l], 0.07f, kt, krk, w, h, &rgbSize, &sigma);
            }
            float tmpError = (float)(sqrt((double)sigma) / (double)rgbSize);
            float sigmaVal = (tmpError == 0) ? 1 : (float)rgbSize;
            if (prm->rgbOnly && tmpError > o->lastRGBError) break;
            o->lastRGBError = tmpError; o->lastRGBCount = (float)rgbSize;
            if (prm->rgbOnly) sigmaVal = -1;

            float A_icp[36], b_icp[6], A_rgb[36], b_rgb[6];
            memset(A_icp, 0, sizeof A_icp); memset(b_icp, 0, sizeof b_icp);
            memset(A_rgb, 0, sizeof A_rgb); memset(b_rgb, 0, sizeof b_rgb);
            if (icp) {
                double s[29];
                orc_icp_step(Rcurr, tcurr, frame_vmaps[l], frame_nmaps[l], Rprev_inv, tprev, c, o->vmap_g[l], o->nmap_g[l],
                             0.10f, (float)sin(20.f * 3.14159254f / 180.f), w, h, s);
                unpack29(s, A_icp, b_icp);
                o->lastICPError = sqrtf((float)s[27]) / (float)s[28]; o->lastICPCount = (float)s[28];
            }
            if (rgb) {
                double s[29];
                orc_rgb_step(o->corres[l], sigmaVal, o->cloud[l], c.fx, c.fy, o->dIdx[l], o->dIdy[l], sobelScale, w, h, s);
                unpack29(s, A_rgb, b_rgb);
            }
            double A[36], b[6], result[6];
            if (icp && rgb) {
                double wgt = prm->icpWeight;
                for (int q = 0; q < 36; ++q) A[q] = (double)A_rgb[q] + wgt * wgt * (double)A_icp[q];
                for (int q = 0; q < 6; ++q) b[q] = (double)b_rgb[q] + wgt * (double)b_icp[q];
            } else if (icp) {
                for (int q = 0; q < 36; ++q) A[q] = A_icp[q];
                for (int q = 0; q < 6; ++q) b[q] = b_icp[q];
            } else {
                for (int q = 0; q < 36; ++q) A[q] = A_rgb[q];
                for (int q = 0; q < 6; ++q) b[q] = b_rgb[q];
            }
            memcpy(o->lastA, A, sizeof A); memcpy(o->lastb, b, sizeof b);
            orc_ldlt_solve(A, b, 6, result);

            /* computeUpdateSE3, OdometryProvider.h:69-90 */
            double Rt[16] = { 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1 }, Rup[9];
            orc_rodrigues(&result[3], Rup);
            for (int r = 0; r < 3; ++r) { for (int cc = 0; cc < 3; ++cc) Rt[r * 4 + cc] = Rup[r * 3 + cc]; Rt[r * 4 + 3] = result[r]; }
            double nr[16];
            for (int r = 0; r < 4; ++r) for (int cc = 0; cc < 4; ++cc) {
                double sacc = 0; for (int k = 0; k < 4; ++k) sacc += Rt[r * 4 + k] * resultRt[k * 4 + cc];
                nr[r * 4 + cc] = sacc;
            }
            memcpy(resultRt, nr, sizeof nr);
            for (int r = 0; r < 3; ++r) { for (int cc = 0; cc < 3; ++cc) trR[r * 3 + cc] = (float)resultRt[r * 4 + cc]; trT[r] = (float)resultRt[r * 4 + 3]; }
            /* currentT = [Rprev|tprev] * transform.inverse()   (RGBDOdometry.cpp:466-474) */
            float iR[9], iT[3];
            for (int r = 0; r < 3; ++r) for (int cc = 0; cc < 3; ++cc) iR[r * 3 + cc] = trR[cc * 3 + r];
            for (int r = 0; r < 3; ++r) iT[r] = -((iR[r * 3] * trT[0] + iR[r * 3 + 1] * trT[1]) + iR[r * 3 + 2] * trT[2]);
            for (int r = 0; r < 3; ++r) {
                for (int cc = 0; cc < 3; ++cc) Rcurr[r * 3 + cc] = (Rprev[r * 3] * iR[cc] + Rprev[r * 3 + 1] * iR[3 + cc]) + Rprev[r * 3 + 2] * iR[6 + cc];
                tcurr[r] = ((Rprev[r * 3] * iT[0] + Rprev[r * 3 + 1] * iT[1]) + Rprev[r * 3 + 2] * iT[2]) + tprev[r];
            }
        }
    }
    float dx = tcurr[0] - tprev[0], dy = tcurr[1] - tprev[1], dz = tcurr[2] - tprev[2];
    if (rgb && sqrtf((dx * dx + dy * dy) + dz * dz) > 0.3f) {
        memcpy(Rcurr, Rprev, sizeof Rprev); memcpy(tcurr, tprev, sizeof tprev);
        float I9[9] = { 1, 0, 0, 0, 1, 0, 0, 0, 1 }; memcpy(trR, I9, sizeof I9); trT[0] = trT[1] = trT[2] = 0;
    }
    if (prm->so3) for (int l = 0; l < 3; ++l) { uint8_t* t = o->lastNextImage[l]; o->lastNextImage[l] = o->nextImage[l]; o->nextImage[l] = t; }
    for (int r = 0; r < 3; ++r) { for (int cc = 0; cc < 3; ++cc) pose[r * 4 + cc] = Rcurr[r * 3 + cc]; pose[r * 4 + 3] = tcurr[r]; }
    if (transformOut) {
        for (int q = 0; q < 16; ++q) transformOut[q] = (q % 5 == 0) ? 1.0f : 0.0f;
        for (int r = 0; r < 3; ++r) { for (int cc = 0; cc < 3; ++cc) transformOut[r * 4 + cc] = trR[r * 3 + cc]; transformOut[r * 4 + 3] = trT[r]; }
    }
}
