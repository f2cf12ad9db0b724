/*
 * oracle/orc_pipeline.c -- CPU ORACLE (test infrastructure, not product):
 * restatement of MaskFusion::processFrame (Core/MaskFusion.cpp:200-607),
 * MaskFusion::predict (:616-628), Model::performTracking / initICP / fuse / clean
 * (Core/Model/Model.cpp:391-464, 466-772), computeFusionWeight (:449-464).
 *
 * Schedule note: the reference calls predict() twice per frame (:423 and :569).
 * The first call's outputs are overwritten by the second before anything reads
 * them when closeLoops==false (the only reachable configuration);
 * the oracle therefore evaluates predict() once, at :569.
 */
#include "orc_pipeline.h"
#include <math.h>
#include <float.h>
#include <limits.h>
#include <stdlib.h>
#include <string.h>

void orc_config_defaults(orc_config* c, int width, int height)
{
    memset(c, 0, sizeof *c);
    c->width = width; c->height = height;
    if (width == 640) { c->fx = 528; c->fy = 528; c->cx = 320; c->cy = 240; }        /* MainController.cpp:124-125 */
    else { c->fx = 528.f * width / 640.f; c->fy = c->fx; c->cx = width / 2.f; c->cy = height / 2.f; }
    c->depthCutoff = 4.0f; c->maxDepthProcessed = 20.0f; c->icpWeight = 20.0f;
    c->rgbOnly = 0; c->pyramid = 1; c->fastOdom = 0; c->so3 = 1; c->frameToFrameRGB = 0;
    c->confGlobal = 10.0f; c->confObject = 0.01f;
    c->timeDelta = INT_MAX / 2;
    c->outlierCoeff = 0.1f;
    c->capacityGlobal = 3072 * 3072; c->capacityObject = 1024 * 1024;
    c->enableMultipleModels = 0; c->trackAllModels = 0; c->modelSpawnOffset = 22;
    c->minRelSizeNew = 0.015f; c->maxRelSizeNew = 0.4f;
    c->segThreshold = 0.3f; c->segWeightDistance = 150.f; c->segWeightConvexity = 2.8f;
    c->segMorphEdgeIterations = 0; c->segMorphEdgeRadius = 1; c->segMorphMaskIterations = 0; c->segMorphMaskRadius = 2;
}

static void ident(float* T) { for (int i = 0; i < 16; ++i) T[i] = (i % 5 == 0) ? 1.f : 0.f; }

static orc_model* model_create(orc_mf* h, int id, float conf, int fillIn, int capacity)
{
    size_t P = (size_t)h->cfg.width * h->cfg.height;
    orc_model* m = (orc_model*)calloc(1, sizeof *m);
    m->id = id; m->classID = -1; m->confThreshold = conf; m->maxDepth = FLT_MAX; m->capacity = capacity;
    ident(m->pose); ident(m->lastPose); ident(m->initialC2Winv);
    m->surf[0] = (float*)calloc((size_t)capacity * 12, sizeof(float));
    m->surf[1] = (float*)calloc((size_t)capacity * 12, sizeof(float));
    m->allowFillIn = fillIn;
    m->idx = (uint32_t*)calloc(P, 4); m->vertConf = (float*)calloc(P * 4, 4); m->colorTime = (float*)calloc(P * 4, 4); m->normRad = (float*)calloc(P * 4, 4);
    m->splatImage = (uint8_t*)calloc(P * 4, 1); m->splatVertex = (float*)calloc(P * 4, 4); m->splatNormal = (float*)calloc(P * 4, 4); m->splatTime = (uint16_t*)calloc(P, 2);
    m->fillVertex = (float*)calloc(P * 4, 4); m->fillNormal = (float*)calloc(P * 4, 4); m->fillImage = (uint8_t*)calloc(P * 4, 1);
    m->updateId = (uint8_t*)calloc(P, 1); m->best = (uint32_t*)calloc(P, 4); m->meas = (float*)calloc(P * 12, 4);
    m->odom = orc_odom_create(h->cfg.width, h->cfg.height, h->cam);
    return m;
}
static void model_destroy(orc_model* m)
{
    if (!m) return;
    free(m->surf[0]); free(m->surf[1]); free(m->idx); free(m->vertConf); free(m->colorTime); free(m->normRad);
    free(m->splatImage); free(m->splatVertex); free(m->splatNormal); free(m->splatTime);
    free(m->fillVertex); free(m->fillNormal); free(m->fillImage); free(m->updateId); free(m->best); free(m->meas);
    orc_odom_destroy(m->odom); free(m->log); free(m);
}

orc_mf* orc_mf_create(const orc_config* cfg)
{
    orc_mf* h = (orc_mf*)calloc(1, sizeof *h);
    h->cfg = *cfg;
    h->cam.fx = cfg->fx; h->cam.fy = cfg->fy; h->cam.cx = cfg->cx; h->cam.cy = cfg->cy;
    h->tick = 1;                                        /* MaskFusion.h: tick(1) */
    size_t P = (size_t)cfg->width * cfg->height;
    h->rgb = (uint8_t*)calloc(P * 3, 1); h->depthRaw = (float*)calloc(P, 4); h->depthFilt = (float*)calloc(P, 4); h->mask = (uint8_t*)calloc(P, 1);
    for (int l = 0; l < 3; ++l) {
        size_t Pl = (size_t)(cfg->width >> l) * (cfg->height >> l);
        h->depthPyr[l] = (float*)calloc(Pl, 4); h->maskPyr[l] = (uint8_t*)calloc(Pl, 1);
        h->vmap[l] = (float*)calloc(Pl * 3, 4); h->nmap[l] = (float*)calloc(Pl * 3, 4);
    }
    h->projKeys = (uint64_t*)calloc(P, 8); h->projectedIDs = (uint8_t*)calloc(P, 1); h->fullSeg = (uint8_t*)calloc(P, 1);
    h->edgeMap = (float*)calloc(P, 4); h->edgeBin = (uint8_t*)calloc(P, 1); h->edgeBuf = (uint8_t*)calloc(P, 1);
    orc_mfseg_state_init(&h->seg, cfg->width, cfg->height);
    h->nextID = 0;
    h->models[0] = model_create(h, h->nextID++, cfg->confGlobal, 1, cfg->capacityGlobal);   /* MaskFusion.cpp:80-81 */
    h->nmodels = 1;
    return h;
}

void orc_mf_destroy(orc_mf* h)
{
    if (!h) return;
    for (int i = 0; i < h->nmodels; ++i) model_destroy(h->models[i]);
    free(h->rgb); free(h->depthRaw); free(h->depthFilt); free(h->mask);
    free(h->projKeys); free(h->projectedIDs); free(h->fullSeg); free(h->edgeMap); free(h->edgeBin); free(h->edgeBuf);
    orc_mfseg_state_free(&h->seg);
    for (int l = 0; l < 3; ++l) { free(h->depthPyr[l]); free(h->maskPyr[l]); free(h->vmap[l]); free(h->nmap[l]); }
    free(h);
}

orc_model* orc_mf_model(orc_mf* h, int i) { return (i >= 0 && i < h->nmodels) ? h->models[i] : 0; }
const float* orc_model_surfels(const orc_model* m) { return m->surf[m->target]; }

/* Model::generateCUDATextures, Model.cpp:350-389 */
void orc_generate_frame_maps(orc_mf* h)
{
    int W = h->cfg.width, H = h->cfg.height;
    memcpy(h->depthPyr[0], h->depthFilt, (size_t)W * H * 4);
    memcpy(h->maskPyr[0], h->mask, (size_t)W * H);
    for (int l = 1; l < 3; ++l) {
        orc_pyrdown_gauss_f(h->depthPyr[l - 1], W >> (l - 1), H >> (l - 1), h->depthPyr[l]);
        orc_pyrdown_gauss_u8(h->maskPyr[l - 1], W >> (l - 1), H >> (l - 1), h->maskPyr[l]);
    }
    for (int l = 0; l < 3; ++l) {
        orc_vmap(h->depthPyr[l], W >> l, H >> l, orc_cam_level(h->cam, l), h->cfg.depthCutoff, h->vmap[l]);
        orc_nmap(h->vmap[l], W >> l, H >> l, h->nmap[l]);
    }
}

/* Model::performTracking + initICP, Model.cpp:391-447 */
void orc_model_track(orc_mf* h, orc_model* m, float* transformOut)
{
    int W = h->cfg.width, H = h->cfg.height;
    int doFillIn = m->allowFillIn ? orc_requires_fill_in(m->splatImage, W, H, 0.75f) : 0;   /* MaskFusion.cpp:630-648 */
    memcpy(m->lastPose, m->pose, sizeof m->pose);
    if (doFillIn) {
        orc_odom_init_icp_model(m->odom, m->fillVertex, m->fillNormal, m->pose);
        orc_odom_init_rgb_model(m->odom, m->fillImage);
    } else {
        orc_odom_init_icp_model(m->odom, m->splatVertex, m->splatNormal, m->pose);
        orc_odom_init_rgb_model(m->odom, (h->cfg.frameToFrameRGB && m->allowFillIn) ? m->fillImage : m->splatImage);
    }
    orc_odom_init_rgb(m->odom, h->rgb);
    orc_track_params p = { h->cfg.rgbOnly, h->cfg.icpWeight, h->cfg.pyramid, h->cfg.fastOdom, h->cfg.so3 };
    orc_odom_track(m->odom, h->vmap, h->nmap, &p, m->pose, transformOut);
}

/* Model::rodrigues2 (Model.cpp:890-932) without the SVD re-orthonormalisation
 * (input is already a product of rotations; deviation <= 1e-7) */
static void rodrigues2(const float* R, float* out)
{
    double rx = R[7] - R[5], ry = R[2] - R[6], rz = R[3] - R[1];
    double s = sqrt((rx * rx + ry * ry + rz * rz) * 0.25);
    double c = ((double)(R[0] + R[4] + R[8]) - 1) * 0.5;
    c = c > 1. ? 1. : c < -1. ? -1. : c;
    double theta = acos(c);
    if (s < 1e-5) {
        double t;
        if (c > 0) rx = ry = rz = 0;
        else {
            t = (R[0] + 1) * 0.5; rx = sqrt(t > 0 ? t : 0.0);
            t = (R[4] + 1) * 0.5; ry = sqrt(t > 0 ? t : 0.0) * (R[1] < 0 ? -1.0 : 1.0);
            t = (R[8] + 1) * 0.5; rz = sqrt(t > 0 ? t : 0.0) * (R[2] < 0 ? -1.0 : 1.0);
            if (fabs(rx) < fabs(ry) && fabs(rx) < fabs(rz) && (R[5] > 0) != (ry * rz > 0)) rz = -rz;
            theta /= sqrt(rx * rx + ry * ry + rz * rz);
            rx *= theta; ry *= theta; rz *= theta;
        }
    } else {
        double vth = 1 / (2 * s); vth *= theta; rx *= vth; ry *= vth; rz *= vth;
    }
    out[0] = (float)rx; out[1] = (float)ry; out[2] = (float)rz;
}

/* Model::computeFusionWeight, Model.cpp:449-464 */
float orc_model_fusion_weight(const orc_model* m, float weightMultiplier)
{
    float inv[16], diff[16];
    orc_pose_inverse(m->pose, inv);
    orc_pose_mul(inv, m->lastPose, diff);
    float R[9] = { diff[0], diff[1], diff[2], diff[4], diff[5], diff[6], diff[8], diff[9], diff[10] };
    float tn = sqrtf((diff[3] * diff[3] + diff[7] * diff[7]) + diff[11] * diff[11]);
    float rv[3]; rodrigues2(R, rv);
    float rn = sqrtf((rv[0] * rv[0] + rv[1] * rv[1]) + rv[2] * rv[2]);
    float weighting = tn > rn ? tn : rn;
    const float largest = 0.01f, minWeight = 0.5f;
    if (weighting > largest) weighting = largest;
    float w = 1.0f - (weighting / largest);
    return (w > minWeight ? w : minWeight) * weightMultiplier;
}

void orc_model_predict_indices(orc_mf* h, orc_model* m, int time)
{
    orc_predict_indices(m->surf[m->target], m->count, m->pose, h->cam, h->cfg.width, h->cfg.height,
                        h->cfg.maxDepthProcessed, time, h->cfg.timeDelta, m->idx, m->vertConf, m->colorTime, m->normRad);
}

/* Model::fuse, Model.cpp:466-647 (headless: lastBoundingBox empty, N7). The
 * reference writes the updated VBO to the other buffer and swaps; the oracle
 * updates in place (identical contents). */
void orc_model_fuse(orc_mf* h, orc_model* m, int time, float depthCutoff, float weightMultiplier)
{
    float maxDepth = depthCutoff < m->maxDepth ? depthCutoff : m->maxDepth;
    orc_data_associate(h->rgb, h->depthRaw, h->depthFilt, h->mask, m->idx, m->vertConf, m->normRad, m->pose, h->cam,
                       h->cfg.width, h->cfg.height, maxDepth, time, orc_model_fusion_weight(m, weightMultiplier),
                       (uint8_t)m->id, m->updateId, m->best, m->meas);
    orc_fuse_update(m->surf[m->target], m->count, m->updateId, m->best, m->meas, h->cfg.width, h->cfg.height, time);
}

/* Model::clean, Model.cpp:649-772 */
void orc_model_clean(orc_mf* h, orc_model* m, int time)
{
    int other = 1 - m->target;
    m->count = orc_clean(m->surf[m->target], m->count, m->updateId, m->meas, m->idx, m->vertConf, m->colorTime,
                         h->depthFilt, h->mask, m->pose, h->cam, h->cfg.width, h->cfg.height, time, h->cfg.timeDelta,
                         m->confThreshold, h->cfg.outlierCoeff, (uint8_t)m->id, m->surf[other], m->capacity);
    m->target = other;
}

void orc_model_combined_predict(orc_mf* h, orc_model* m, int time, int maxTime)
{
    orc_combined_predict(m->surf[m->target], m->count, m->pose, h->cam, h->cfg.width, h->cfg.height,
                         h->cfg.maxDepthProcessed, m->confThreshold, time, maxTime, h->cfg.timeDelta,
                         m->splatImage, m->splatVertex, m->splatNormal, m->splatTime);
}

/* Model::performFillIn, Model.cpp:976-984 (lost == false) */
void orc_model_fill_in(orc_mf* h, orc_model* m)
{
    if (!m->allowFillIn) return;
    orc_fill_in(m->splatVertex, m->splatNormal, m->splatImage, h->depthFilt, h->rgb, h->cam, h->cfg.width, h->cfg.height,
                0, h->cfg.frameToFrameRGB, m->fillVertex, m->fillNormal, m->fillImage);
}

static void predict(orc_mf* h)
{
    for (int i = 0; i < h->nmodels; ++i) {
        orc_model_combined_predict(h, h->models[i], h->tick, h->tick);
        orc_model_fill_in(h, h->models[i]);
    }
}

/* Eigen::Quaternionf(Matrix3f) (Shepperd's method as in Eigen/Geometry/Quaternion.h) */
static void rot_to_quat(const float* R, float* q /* x y z w */)
{
    float t = R[0] + R[4] + R[8];
    if (t > 0) {
        t = sqrtf(t + 1.0f); q[3] = 0.5f * t; t = 0.5f / t;
        q[0] = (R[7] - R[5]) * t; q[1] = (R[2] - R[6]) * t; q[2] = (R[3] - R[1]) * t;
    } else {
        int i = 0; if (R[4] > R[0]) i = 1; if (R[8] > R[i * 4]) i = 2;
        int j = (i + 1) % 3, k = (j + 1) % 3;
        t = sqrtf(R[i * 4] - R[j * 4] - R[k * 4] + 1.0f);
        q[i] = 0.5f * t; t = 0.5f / t;
        q[3] = (R[k * 3 + j] - R[j * 3 + k]) * t;
        q[j] = (R[j * 3 + i] + R[i * 3 + j]) * t;
        q[k] = (R[k * 3 + i] + R[i * 3 + k]) * t;
    }
}

static void log_pose(orc_model* m, int64_t ts, const float* T)
{
    if (m->nlog == m->caplog) { m->caplog = m->caplog ? m->caplog * 2 : 64; m->log = (double*)realloc(m->log, (size_t)m->caplog * 8 * sizeof(double)); }
    float R[9] = { T[0], T[1], T[2], T[4], T[5], T[6], T[8], T[9], T[10] }, q[4];
    rot_to_quat(R, q);
    double* o = m->log + (size_t)m->nlog * 8;
    o[0] = (double)ts; o[1] = T[3]; o[2] = T[7]; o[3] = T[11]; o[4] = q[0]; o[5] = q[1]; o[6] = q[2]; o[7] = q[3];
    m->nlog++;
}

/* upload + filterDepth (MaskFusion.cpp:212-230, 650-657); mask NULL == -static (all zero) */
void orc_mf_set_frame(orc_mf* h, const uint8_t* rgb3, const float* depth, const uint8_t* mask)
{
    int W = h->cfg.width, H = h->cfg.height; size_t P = (size_t)W * H;
    memcpy(h->rgb, rgb3, P * 3);
    memcpy(h->depthRaw, depth, P * 4);
    orc_bilateral(h->depthRaw, h->depthFilt, W, H);
    if (mask) memcpy(h->mask, mask, P); else memset(h->mask, 0, P);
}

/* GlobalProjection::project + downloadDirect (GlobalProjection.cpp:43-111): every model splatted with the fixed
 * confidence threshold 12 (:61) into one depth-tested image; draw order = list order, then surfel id. */
void orc_mf_global_projection(orc_mf* h)
{
    int W = h->cfg.width, H = h->cfg.height; size_t P = (size_t)W * H;
    uint32_t base[257]; base[0] = 0;
    orc_global_projection_begin(W, H, h->projKeys);
    for (int i = 0; i < h->nmodels; ++i) {
        orc_model* m = h->models[i];
        orc_global_projection_add(m->surf[m->target], m->count, m->pose, h->cam, W, H, h->cfg.depthCutoff, 12.0f, h->tick, h->tick,
                                  h->cfg.timeDelta, base[i], h->projKeys);
        base[i + 1] = base[i] + (uint32_t)m->count;
    }
    for (size_t p = 0; p < P; ++p) {
        uint8_t id = 0;
        if (h->projKeys[p] != ~0ull) {
            uint32_t d = (uint32_t)(h->projKeys[p] & 0xffffffffu);
            for (int i = 0; i < h->nmodels; ++i) if (d >= base[i] && d < base[i + 1]) { id = (uint8_t)h->models[i]->id; break; }
        }
        h->projectedIDs[p] = id;
    }
}

/* MfSegmentation::performSegmentation (MfSegmentation.cpp:83-538): GPU-kernel part restated by orc_geometric_edges /
 * orc_threshold / orc_morph_close / orc_invert (uses the level-0 TRACKING maps, :149-151), CPU part by orc_mfseg_cpu. */
void orc_mf_segmentation(orc_mf* h, const uint8_t* mask, const int32_t* classIDs, int nMasks, int allowNew, orc_mfseg_out* out)
{
    int W = h->cfg.width, H = h->cfg.height; size_t P = (size_t)W * H;
    orc_geometric_edges(h->vmap[0], h->nmap[0], W, H, h->cfg.segWeightDistance, h->cfg.segWeightConvexity, h->edgeMap);
    orc_threshold(h->edgeMap, (int)P, h->cfg.segThreshold, h->edgeBin);
    orc_morph_close(h->edgeBin, h->edgeBuf, W, H, h->cfg.segMorphEdgeRadius, h->cfg.segMorphEdgeIterations);
    orc_invert(h->edgeBin, (int)P, h->edgeBuf);
    uint8_t ids[256]; int32_t cls[256];
    for (int i = 0; i < h->nmodels; ++i) { ids[i] = (uint8_t)h->models[i]->id; cls[i] = h->models[i]->classID; }
    orc_mfseg_in in;
    memset(&in, 0, sizeof in);
    in.W = W; in.H = H; in.edgesInv = h->edgeBuf; in.depth = h->depthRaw; in.mask = mask; in.nMasks = nMasks; in.classIDs = classIDs;
    in.projectedIDs = h->projectedIDs; in.nModels = h->nmodels; in.modelIDs = ids; in.modelClassIDs = cls;
    in.nextModelID = h->nextID; in.allowNew = allowNew; in.minRelSizeNew = h->cfg.minRelSizeNew; in.maxRelSizeNew = h->cfg.maxRelSizeNew;
    in.morphMaskRadius = h->cfg.segMorphMaskRadius; in.morphMaskIterations = h->cfg.segMorphMaskIterations;
    out->fullSegmentation = h->fullSeg; out->labels = 0;
    orc_mfseg_cpu(&h->seg, &in, out);
}

/* MaskFusion::getNextModelID(assign=true), MaskFusion.cpp:712-730 */
static uint8_t next_model_id_assign(orc_mf* h)
{
    uint8_t next = h->nextID;
    for (;;) {
        h->nextID++;
        int occupied = 0;
        for (int i = 0; i < h->nmodels; ++i) if (h->nextID == h->models[i]->id) occupied = 1;
        if (!occupied) break;
    }
    return next;
}

int orc_mf_process_frame_ex(orc_mf* h, const uint8_t* rgb3, const float* depth, int64_t timestamp, const uint8_t* maskIn,
                            const int32_t* classIDs, int nMasks)
{
    int W = h->cfg.width, H = h->cfg.height; size_t P = (size_t)W * H;
    const int multi = h->cfg.enableMultipleModels;
    orc_mf_set_frame(h, rgb3, depth, multi ? h->mask : 0);          /* multi: textureMask keeps the last segmentation until :297 */
    orc_model* g = h->models[0];
    if (h->tick == 1) {
        g->count = orc_init_model(h->rgb, h->depthRaw, h->depthFilt, h->cam, W, H, h->tick, h->cfg.maxDepthProcessed,
                                  g->surf[g->target], g->capacity);
        orc_odom_init_first_rgb(g->odom, h->rgb);
    } else {
        orc_generate_frame_maps(h);
        orc_model_track(h, g, 0);
        for (int i = 1; i < h->nmodels; ++i) {                       /* MaskFusion.cpp:258-276 */
            orc_model* m = h->models[i];
            if (m->nonstatic || h->cfg.trackAllModels) {
                float T[16];
                orc_model_track(h, m, T);
                float d = sqrtf((T[3] * T[3] + T[7] * T[7]) + T[11] * T[11]);
                if (d > 0.2f) {                                      /* inactivateModel (:268-272) */
                    model_destroy(m);
                    for (int k = i; k + 1 < h->nmodels; ++k) h->models[k] = h->models[k + 1];
                    h->nmodels--; --i;
                }
            } else {
                float nw[16];
                orc_pose_mul(m->initialC2Winv, g->pose, nw);          /* updateStaticPose, Model.h:263 */
                memcpy(m->lastPose, m->pose, sizeof m->pose); memcpy(m->pose, nw, sizeof nw);
            }
        }
        if (multi) {
            orc_mf_global_projection(h);                             /* :289-290 */
            if (h->spawnOffset < h->cfg.modelSpawnOffset) h->spawnOffset++;
            orc_mfseg_out so;
            orc_mf_segmentation(h, maskIn, classIDs, nMasks, h->spawnOffset >= h->cfg.modelSpawnOffset, &so);
            memcpy(h->mask, h->fullSeg, P);                          /* textureMask upload, :297 */
            h->lastHasNewLabel = so.hasNewLabel;
            orc_model* nm = 0;
            if (so.hasNewLabel) {                                    /* spawnObjectModel + moveNewModelToList, :313-334, :671-684 */
                uint8_t id = next_model_id_assign(h);
                nm = model_create(h, id, h->cfg.confObject, 0, h->cfg.capacityObject);
                orc_odom_init_first_rgb(nm->odom, h->rgb);
                float ginv[16]; orc_pose_inverse(g->pose, ginv);
                orc_pose_mul(nm->pose, ginv, nm->initialC2Winv);     /* makeStatic, Model.h:264 */
                nm->isStatic = 1;
                h->spawnOffset = 0;
                nm->maxDepth = 30.0f + 30.0f * 1.2f;                 /* getMaxDepth(depthMean=30, depthStd=30), :292,328 */
                nm->classID = so.newClassID;
                h->models[h->nmodels++] = nm;
            }
            for (int i = 1; i < h->nmodels; ++i) h->models[i]->maxDepth = 30.0f + 30.0f * 1.2f;     /* :337-341 */
            if (nm) {                                                /* :344-353 */
                orc_model_predict_indices(h, nm, h->tick);
                orc_model_fuse(h, nm, h->tick, h->cfg.maxDepthProcessed, 100.0f);
                orc_model_clean(h, nm, h->tick);
            }
            for (int i = 1; i < h->nmodels; ++i) {                   /* :369-374 */
                float f = (float)h->models[i]->age / 25.0f;
                h->models[i]->confThreshold = f < 4.5f ? f : 4.5f;
            }
        }
        if (!h->cfg.rgbOnly) {
            for (int i = 0; i < h->nmodels; ++i) orc_model_predict_indices(h, h->models[i], h->tick);
            for (int i = 0; i < h->nmodels; ++i) orc_model_fuse(h, h->models[i], h->tick, h->cfg.depthCutoff, 1.0f);
            for (int i = 0; i < h->nmodels; ++i) orc_model_predict_indices(h, h->models[i], h->tick);
            for (int i = 0; i < h->nmodels; ++i) orc_model_clean(h, h->models[i], h->tick);
        }
    }
    predict(h);
    h->tick++;
    g = h->models[0];
    for (int i = 0; i < h->nmodels; ++i) {
        orc_model* m = h->models[i];
        float T[16];
        if (i == 0) memcpy(T, g->pose, sizeof T);
        else { float inv[16]; orc_pose_inverse(m->pose, inv); orc_pose_mul(g->pose, inv, T); }
        log_pose(m, timestamp, T);
        m->age++;
    }
    return 0;
}

int orc_mf_process_frame(orc_mf* h, const uint8_t* rgb3, const float* depth, int64_t timestamp)
{
    return orc_mf_process_frame_ex(h, rgb3, depth, timestamp, 0, 0, 0);
}
