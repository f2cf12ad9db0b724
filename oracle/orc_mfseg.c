/*
 * oracle/orc_mfseg.c -- CPU ORACLE (test infrastructure, not product): restatement of the
 * CPU part of MfSegmentation::performSegmentation (Core/Segmentation/MfSegmentation.cpp:208-538)
 * -- BASELINE.json configs[0] times exactly this function on one 640x480 frame.
 *
 * OpenCV calls restated (OpenCV is not available as a C++ library here; pinned 3.4.1 upstream):
 *   connectedComponentsWithStats(img, labels, stats, centroids, 4): two-pass union-find, final labels in
 *       raster order of each component's first pixel (label numbering does not influence the result);
 *       stats = left, top, width, height, area.
 *   morphologyEx(MORPH_CLOSE, ellipse (2r+1)^2, iterations): dilate x iterations then erode x iterations,
 *       border taps ignored (morphologyDefaultBorderValue), iterations == 0 => copy.
 * removeEdgeIslands (default false, no setter reaches it from the GUI) is not restated.
 */
#include "orc_mfseg.h"
#include <math.h>
#include <stdlib.h>
#include <string.h>

/* BASELINE configs[0] / SURVEY 8(d)(i) times this file on ONE core (the reference is single-threaded here) and with "a
 * straightforward OpenMP split" over the host cores: every per-pixel sweep below is an `omp parallel for` that is only active
 * when orc_mfseg_set_threads(n > 1) was called; integer histograms use atomic adds (order-free), the Jacobi sweeps read the
 * previous labels only, the two-pass connected-component labelling stays sequential.  Results are identical for any n. */
static int g_threads = 1;
void orc_mfseg_set_threads(int n) { g_threads = n < 1 ? 1 : n; }
#define PAR _Pragma("omp parallel for schedule(static) num_threads(g_threads) if (g_threads > 1)")

static int uf_find(int* p, int x) { while (p[x] != x) { p[x] = p[p[x]]; x = p[x]; } return x; }

/* 4-connected components of the non-zero pixels; labels[] gets 0 for background, 1..n-1 otherwise;
 * stats[label*5 + {0:left,1:top,2:width,3:height,4:area}].  Returns n (including background). */
int orc_connected_components4(const uint8_t* img, int W, int H, int32_t* labels, int32_t** statsOut)
{
    int P = W * H;
    int* parent = (int*)malloc((size_t)(P / 2 + 2) * sizeof(int));
    int next = 1;
    parent[0] = 0;
    for (int y = 0; y < H; ++y)
        for (int x = 0; x < W; ++x) {
            int i = y * W + x;
            if (!img[i]) { labels[i] = 0; continue; }
            int up = (y > 0 && img[i - W]) ? labels[i - W] : 0;
            int left = (x > 0 && img[i - 1]) ? labels[i - 1] : 0;
            if (!up && !left) { parent[next] = next; labels[i] = next++; }
            else if (up && left) {
                int a = uf_find(parent, up), b = uf_find(parent, left);
                int r = a < b ? a : b;
                parent[a] = r; parent[b] = r;
                labels[i] = r;
            } else labels[i] = up ? up : left;
        }
    int* remap = (int*)calloc((size_t)next, sizeof(int));
    int n = 1;
    for (int l = 1; l < next; ++l) if (uf_find(parent, l) == l) remap[l] = n++;
    int32_t* stats = (int32_t*)malloc((size_t)n * 5 * sizeof(int32_t));
    for (int l = 0; l < n; ++l) { stats[l * 5] = W; stats[l * 5 + 1] = H; stats[l * 5 + 2] = -1; stats[l * 5 + 3] = -1; stats[l * 5 + 4] = 0; }
    for (int y = 0; y < H; ++y)
        for (int x = 0; x < W; ++x) {
            int i = y * W + x;
            int l = labels[i] ? remap[uf_find(parent, labels[i])] : 0;
            labels[i] = l;
            int32_t* s = stats + l * 5;
            if (x < s[0]) s[0] = x;
            if (y < s[1]) s[1] = y;
            if (x > s[2]) s[2] = x;      /* right, converted to width below */
            if (y > s[3]) s[3] = y;
            s[4]++;
        }
    for (int l = 0; l < n; ++l) {
        int32_t* s = stats + l * 5;
        if (s[4] == 0) { s[0] = s[1] = s[2] = s[3] = 0; }
        else { s[2] = s[2] - s[0] + 1; s[3] = s[3] - s[1] + 1; }
    }
    free(parent); free(remap);
    *statsOut = stats;
    return n;
}

/* cv::getStructuringElement(MORPH_ELLIPSE, (2r+1, 2r+1)) */
static void ellipse_element(int r, uint8_t* k)
{
    int n = 2 * r + 1, c = r;
    double inv_r2 = r ? 1.0 / ((double)r * r) : 0;
    memset(k, 0, (size_t)n * n);
    for (int i = 0; i < n; ++i) {
        int dy = i - r, j1 = 0, j2 = 0;
        if (abs(dy) <= r) {
            int dx = (int)lrint(c * sqrt((r * r - dy * dy) * inv_r2));
            j1 = c - dx > 0 ? c - dx : 0;
            j2 = c + dx + 1 < n ? c + dx + 1 : n;
        }
        for (int j = j1; j < j2; ++j) k[i * n + j] = 1;
    }
}
static void morph_gray(const uint8_t* in, uint8_t* out, int W, int H, int r, const uint8_t* k, int dilate)
{
    int n = 2 * r + 1;
    PAR
    for (int y = 0; y < H; ++y)
        for (int x = 0; x < W; ++x) {
            int best = dilate ? 0 : 255;
            for (int i = 0; i < n; ++i)
                for (int j = 0; j < n; ++j) {
                    if (!k[i * n + j]) continue;
                    int yy = y + i - r, xx = x + j - r;
                    if (yy < 0 || yy >= H || xx < 0 || xx >= W) continue;
                    int v = in[yy * W + xx];
                    if (dilate ? v > best : v < best) best = v;
                }
            out[y * W + x] = (uint8_t)best;
        }
}
void orc_morph_close_ellipse(uint8_t* img, int W, int H, int r, int iterations)
{
    if (iterations <= 0) return;
    uint8_t* k = (uint8_t*)malloc((size_t)(2 * r + 1) * (2 * r + 1));
    uint8_t* tmp = (uint8_t*)malloc((size_t)W * H);
    ellipse_element(r, k);
    for (int i = 0; i < iterations; ++i) { morph_gray(img, tmp, W, H, r, k, 1); memcpy(img, tmp, (size_t)W * H); }
    for (int i = 0; i < iterations; ++i) { morph_gray(img, tmp, W, H, r, k, 0); memcpy(img, tmp, (size_t)W * H); }
    free(k); free(tmp);
}

void orc_mfseg_state_init(orc_mfseg_state* s, int W, int H)
{
    memset(s, 0, sizeof *s);
    s->semanticIgnoreMap = (uint8_t*)calloc((size_t)W * H, 1);
    s->maskToID[255] = 255; s->maskToID[0] = 0;                      /* MfSegmentation.cpp:70-71 */
    s->minMaskModelOverlap = 0.05f; s->minMappedComponentSize = 160; s->personClassID = 255; s->removeEdges = 1;
}
void orc_mfseg_state_free(orc_mfseg_state* s) { free(s->semanticIgnoreMap); s->semanticIgnoreMap = 0; }

/* MfSegmentation.cpp:208-538.  edgesInv is overwritten (cv8UC1Buffer). */
void orc_mfseg_cpu(orc_mfseg_state* st, const orc_mfseg_in* in, orc_mfseg_out* out)
{
    const int W = in->W, H = in->H; const size_t total = (size_t)W * H;
    uint8_t* buf = in->edgesInv;
    const int nMasks = in->nMasks, nModels = in->nModels;
    const size_t minNewMaskPixels = (size_t)(in->minRelSizeNew * total), maxNewMaskPixels = (size_t)(in->maxRelSizeNew * total);
    uint8_t* seg = out->fullSegmentation;
    memset(seg, 0, total);
    out->hasNewLabel = 0; out->newClassID = -1;
    for (int m = 0; m < nModels; ++m) { out->isEmpty[m] = 1; out->pixelCount[m] = 0; }
    uint8_t modelIDToIndex[256]; memset(modelIDToIndex, 0, sizeof modelIDToIndex);
    uint8_t modelIndexToID[257]; memset(modelIndexToID, 0, sizeof modelIndexToID);
    for (int m = 0; m < nModels; ++m) { modelIDToIndex[in->modelIDs[m]] = (uint8_t)m; modelIndexToID[m] = in->modelIDs[m]; }
    if (in->allowNew) { modelIDToIndex[in->nextModelID] = (uint8_t)nModels; modelIndexToID[nModels] = in->nextModelID; }

    /* :221-235 ignore map */
    if (nMasks) {
        PAR
        for (size_t i = 0; i < total; ++i) {
            if (in->classIDs[in->mask[i]] == st->personClassID) { st->semanticIgnoreMap[i] = 255; buf[i] = 0; }
            else st->semanticIgnoreMap[i] = 0;
        }
    } else {
        PAR
        for (size_t i = 0; i < total; ++i) if (st->semanticIgnoreMap[i]) buf[i] = 0;
    }
    /* :238-239 */
    int32_t* labels = (int32_t*)malloc(total * sizeof(int32_t));
    int32_t* stats = 0;
    int nComponents = orc_connected_components4(buf, W, H, labels, &stats);

    /* :243-291 remove edges: 5 Jacobi sweeps (reads the previous sweep's labels, writes a copy) */
    if (st->removeEdges) {
        const int small_thr = 50, iters = 5;
        int32_t* r = (int32_t*)malloc(total * sizeof(int32_t));
        static const int oy[8] = { -1, -1, -1, 0, 0, 1, 1, 1 }, ox[8] = { -1, 0, 1, -1, 1, -1, 0, 1 };
        for (int it = 0; it < iters; ++it) {
            memcpy(r, labels, total * sizeof(int32_t));
            PAR
            for (int y = 1; y < H - 1; ++y)
                for (int x = 1; x < W - 1; ++x) {
                    int c = r[y * W + x];
                    float d = in->depth[y * W + x];
                    if (c == 0 || stats[c * 5 + 4] < small_thr) {
                        for (int k = 0; k < 8; ++k) {
                            int yy = y + oy[k], xx = x + ox[k];
                            int n = labels[yy * W + xx];
                            if (n != 0 && fabsf(in->depth[yy * W + xx] - d) < 0.008 && stats[n * 5 + 4] > small_thr) { r[y * W + x] = n; break; }
                        }
                    }
                }
            memcpy(labels, r, total * sizeof(int32_t));
        }
        free(r);
    }

    /* :303-346 overlaps */
    int* mapComponentToMask = (int*)calloc((size_t)nComponents, sizeof(int));
    int* maskComponentPixels = (int*)calloc((size_t)(nMasks > 0 ? nMasks : 1), sizeof(int));
    int* compMaskOverlap = (int*)calloc((size_t)nComponents * (nMasks > 0 ? nMasks : 1), sizeof(int));
    int* compModelOverlap = (int*)calloc((size_t)nComponents * (nModels + 1), sizeof(int));
    /* histograms: one private copy per thread, merged in thread order (integer sums: order-free) */
    {
        const size_t nA = (size_t)nComponents * (nModels + 1), nB = nMasks ? (size_t)nComponents * nMasks : 0;
        #pragma omp parallel num_threads(g_threads) if (g_threads > 1)
        {
            int* ha = g_threads > 1 ? (int*)calloc(nA + nB + 1, sizeof(int)) : compModelOverlap;
            int* hb = g_threads > 1 ? ha + nA : compMaskOverlap;
            #pragma omp for schedule(static) nowait
            for (size_t i = 0; i < total; ++i) {
                ha[(size_t)labels[i] * (nModels + 1) + modelIDToIndex[in->projectedIDs[i]]]++;
                if (nMasks) hb[(size_t)labels[i] * nMasks + in->mask[i]]++;
            }
            if (g_threads > 1) {
                #pragma omp critical
                {
                    for (size_t q = 0; q < nA; ++q) compModelOverlap[q] += ha[q];
                    for (size_t q = 0; q < nB; ++q) compMaskOverlap[q] += hb[q];
                }
                free(ha);
            }
        }
    }
    if (nMasks) {
        const float overlap_threshold = 0.65f;
        for (int c = 1; c < nComponents; ++c) {
            int csize = stats[c * 5 + 4];
            if (csize > st->minMappedComponentSize) {
                int t = (int)(overlap_threshold * csize);
                for (int m = 1; m < nMasks; ++m)
                    if (compMaskOverlap[(size_t)c * nMasks + m] > t) { mapComponentToMask[c] = m; maskComponentPixels[m] += csize; }
            } else mapComponentToMask[c] = 0;
        }
    }
    PAR
    for (size_t i = 0; i < total; ++i) seg[i] = (uint8_t)mapComponentToMask[labels[i]];
    PAR
    for (size_t i = 0; i < total; ++i) if (st->semanticIgnoreMap[i]) seg[i] = 255;           /* :360-362 */

    if (nMasks) {
        orc_morph_close_ellipse(seg, W, H, in->morphMaskRadius, in->morphMaskIterations);    /* :424-426 */
        for (int m = 1; m < nMasks; ++m) { st->maskToID[m] = 0; if (in->classIDs[m] == st->personClassID) st->maskToID[m] = 255; }
        /* overlap of each (closed) mask with each projected model */
        unsigned* maskOverlap = (unsigned*)calloc((size_t)nModels * 256, sizeof(unsigned));
        #pragma omp parallel num_threads(g_threads) if (g_threads > 1)
        {
            unsigned* hm = g_threads > 1 ? (unsigned*)calloc((size_t)nModels * 256, sizeof(unsigned)) : maskOverlap;
            #pragma omp for schedule(static) nowait
            for (size_t i = 0; i < total; ++i) {
                uint8_t mk = seg[i];
                for (int b = 0; b < nModels; ++b) if (in->projectedIDs[i] == in->modelIDs[b]) hm[b * 256 + mk]++;
            }
            if (g_threads > 1) {
                #pragma omp critical
                for (int q = 0; q < nModels * 256; ++q) maskOverlap[q] += hm[q];
                free(hm);
            }
        }
        for (int midx = 1; midx < nMasks; ++midx) {
            if (st->maskToID[midx] == 255) continue;
            int bestModelIndex = 0; unsigned bestOverlap = 0;
            int maskClassID = in->classIDs[midx];
            for (int j = 1; j < nModels; ++j) { unsigned ov = maskOverlap[j * 256 + midx]; if (ov > bestOverlap) { bestOverlap = ov; bestModelIndex = j; } }
            int matches = in->modelClassIDs[bestModelIndex] == maskClassID;
            if (bestOverlap < st->minMaskModelOverlap * maskComponentPixels[midx]) bestModelIndex = 0;
            if (bestModelIndex != 0 && matches) {
                st->maskToID[midx] = in->modelIDs[bestModelIndex];
                out->isEmpty[bestModelIndex] = 0; out->pixelCount[bestModelIndex] = maskComponentPixels[midx];
            } else if (!out->hasNewLabel && in->allowNew && (size_t)maskComponentPixels[midx] > minNewMaskPixels &&
                       (size_t)maskComponentPixels[midx] < maxNewMaskPixels && bestModelIndex == 0) {
                st->maskToID[midx] = in->nextModelID; out->hasNewLabel = 1; out->newClassID = maskClassID;
            } else st->maskToID[midx] = 255;
        }
        free(maskOverlap);
    }
    PAR
    for (size_t i = 0; i < total; ++i) seg[i] = st->maskToID[seg[i]];                         /* :495-496 */

    /* :498-522 unused components absorbed by the model they overlap */
    for (int c = 1; c < nComponents; ++c) {
        if (mapComponentToMask[c] != 0) continue;
        int model_index = 0, overlap = compModelOverlap[(size_t)c * (nModels + 1)];
        for (int m = 1; m < nModels; ++m) { int v = compModelOverlap[(size_t)c * (nModels + 1) + m]; if (v > overlap) { overlap = v; model_index = m; } }
        int model_id = modelIndexToID[model_index];
        if (model_id > 0 && overlap > 0.6f * stats[c * 5 + 4]) {
            int x1 = stats[c * 5], x2 = x1 + stats[c * 5 + 2], y1 = stats[c * 5 + 1], y2 = y1 + stats[c * 5 + 3];
            for (int y = y1; y <= y2 && y < H; ++y)
                for (int x = x1; x <= x2 && x < W; ++x)
                    if (labels[y * W + x] == c) seg[y * W + x] = (uint8_t)model_id;
        }
    }
    if (out->labels) memcpy(out->labels, labels, total * sizeof(int32_t));
    out->nComponents = nComponents;
    free(labels); free(stats); free(mapComponentToMask); free(maskComponentPixels); free(compMaskOverlap); free(compModelOverlap);
}
