/*
 * oracle/orc_segmentation.c -- CPU ORACLE (test infrastructure, not product):
 * restatement of Core/Cuda/segmentation.cu (edge-ness map, threshold, invert,
 * binary morphology).  fmax()/fmin() follow CUDA semantics (NaN operand is
 * ignored), which is what makes stale y/z planes of invalid map pixels
 * irrelevant: their x plane is NaN and poisons every dot product.
 */
#include "orc.h"
#include <math.h>
#include <string.h>

static inline void get3(const float* m, int W, int H, int x, int y, float* o)
{
    size_t P = (size_t)W * H;
    o[0] = m[(size_t)y * W + x]; o[1] = m[P + (size_t)y * W + x]; o[2] = m[2 * P + (size_t)y * W + x];
}
static inline float dot3(const float* a, const float* b) { return (a[0] * b[0] + a[1] * b[1]) + a[2] * b[2]; }

/* segmentation.cu:103-120 */
static float concavity(const float* vmap, const float* nmap, int W, int H, const float* v, const float* n, int xn, int yn)
{
    float vn[3], nn[3], d[3];
    get3(vmap, W, H, xn, yn, vn); get3(nmap, W, H, xn, yn, nn);
    d[0] = vn[0] - v[0]; d[1] = vn[1] - v[1]; d[2] = vn[2] - v[2];
    if (dot3(d, n) < 0) return 0;
    return 1 - dot3(nn, n);
}
static float distance_term(const float* vmap, int W, int H, const float* v, const float* n, int xn, int yn)
{
    float vn[3], d[3];
    get3(vmap, W, H, xn, yn, vn);
    d[0] = vn[0] - v[0]; d[1] = vn[1] - v[1]; d[2] = vn[2] - v[2];
    return fabsf(dot3(d, n));
}

/* segmentation.cu:122-177 */
void orc_geometric_edges(const float* vmap, const float* nmap, int W, int H, float wD, float wC, float* out)
{
    static const int ox[8] = { -1, 0, 1, -1, 1, -1, 0, 1 }, oy[8] = { -1, -1, -1, 0, 0, 1, 1, 1 };
    for (int y = 0; y < H; ++y)
        for (int x = 0; x < W; ++x) {
            if (x < 1 || x >= W - 1 || y < 1 || y >= H - 1) { out[y * W + x] = 1.0f; continue; }
            float v[3], n[3];
            get3(vmap, W, H, x, y, v); get3(nmap, W, H, x, y, n);
            if (v[2] <= 0.0f) { out[y * W + x] = 1.0f; continue; }
            float c = 0.0f, d = 0.0f;
            for (int k = 0; k < 8; ++k) c = fmaxf(concavity(vmap, nmap, W, H, v, n, x + ox[k], y + oy[k]), c);
            c = fmaxf(c, 0.0f);
            c *= wC;
            for (int k = 0; k < 8; ++k) d = fmaxf(distance_term(vmap, W, H, v, n, x + ox[k], y + oy[k]), d);
            d *= wD;
            float e = c > d ? c : d;               /* max(c,d), both non-NaN here */
            out[y * W + x] = fminf(1.0f, e);
        }
}

void orc_threshold(const float* in, int n, float thr, uint8_t* out) { for (int i = 0; i < n; ++i) out[i] = in[i] > thr ? 255 : 0; }
void orc_invert(const uint8_t* in, int n, uint8_t* out) { for (int i = 0; i < n; ++i) out[i] = (uint8_t)(255 - in[i]); }

/* segmentation.cu:217-255 (centre pixel skipped), host loop :334-354 */
static void morph(const uint8_t* in, uint8_t* out, int W, int H, int r, int dilate)
{
    for (int y = 0; y < H; ++y)
        for (int x = 0; x < W; ++x) {
            int x1 = x - r < 0 ? 0 : x - r, y1 = y - r < 0 ? 0 : y - r;
            int x2 = x + r > W - 1 ? W - 1 : x + r, y2 = y + r > H - 1 ? H - 1 : y + r;
            uint8_t res = dilate ? 0 : 255;
            for (int cy = y1; cy <= y2 && res == (dilate ? 0 : 255); ++cy)
                for (int cx = x1; cx <= x2; ++cx) {
                    if (cy == y && cx == x) continue;
                    if (dilate && in[cy * W + cx] == 255) { res = 255; break; }
                    if (!dilate && in[cy * W + cx] == 0) { res = 0; break; }
                }
            out[y * W + x] = res;
        }
}
void orc_morph_close(uint8_t* data, uint8_t* buf, int W, int H, int radius, int iterations)
{
    for (int i = 0; i < iterations; ++i) { morph(data, buf, W, H, radius, 1); morph(buf, data, W, H, radius, 0); }
}
