// mf_host.cu -- host orchestration of the dense pipeline: MaskFusion::processFrame schedule
// (Core/MaskFusion.cpp:200-607) and the Model methods it drives (Core/Model/Model.cpp).
// Everything is enqueued on one CUDA stream; the only host<->device synchronisation inside
// a frame is the read-back of the tracked poses (one per frame, not one per Gauss-Newton
// iteration as in the reference).
#include "mf_host.h"
#include <math.h>
#include <float.h>
#include <string.h>
#include <stdio.h>
#include <limits.h>
#include <stdlib.h>

namespace mfb {

void cudaCheck(cudaError_t e, const char* where)
{
    if (e != cudaSuccess) throw CudaError{std::string(where) + ": " + cudaGetErrorString(e)};
}

// --------------------------------------------------------------------------------------
// in-stream stage timer: one CUDA event per mark on the launching stream; the interval up
// to the next mark is attributed to the mark's name.  Off by default (zero overhead).
// --------------------------------------------------------------------------------------
Profiler* g_prof = nullptr;
void prof_mark(cudaStream_t s, const char* name)
{
    Profiler* p = g_prof;
    if (!p || !p->on) return;
    if (p->used == (int)p->events.size()) {
        cudaEvent_t e; cudaEventCreate(&e); p->events.push_back(e); p->names.push_back(nullptr);
    }
    p->names[p->used] = name;
    cudaEventRecord(p->events[p->used], s);
    p->used++;
}
void Profiler::resolve()
{
    // caller has synchronised the stream
    for (int i = 0; i + 1 < used; ++i) {
        if (!names[i]) continue;
        float ms = 0;
        if (cudaEventElapsedTime(&ms, events[i], events[i + 1]) == cudaSuccess) { auto& a = acc[names[i]]; a.first += 1; a.second += ms; }
    }
    used = 0;
}

Mat4 rigidInverse(const Mat4& T)
{
    // [R^T | -R^T t] in fp32 (reference: Eigen::Matrix4f::inverse(); rule fixed in DESIGN.md)
    Mat4 o = Mat4::identity();
    const float* m = T.m;
    float R[9] = {m[0], m[1], m[2], m[4], m[5], m[6], m[8], m[9], m[10]}, t[3] = {m[3], m[7], m[11]};
    for (int r = 0; r < 3; ++r) {
        for (int c = 0; c < 3; ++c) o.m[r * 4 + c] = R[c * 3 + r];
        o.m[r * 4 + 3] = -((R[0 * 3 + r] * t[0] + R[1 * 3 + r] * t[1]) + R[2 * 3 + r] * t[2]);
    }
    return o;
}
Mat4 mul(const Mat4& A, const Mat4& B)
{
    Mat4 o;
    for (int r = 0; r < 4; ++r)
        for (int c = 0; c < 4; ++c) {
            float s = 0;
            for (int k = 0; k < 4; ++k) s += A.m[r * 4 + k] * B.m[k * 4 + c];
            o.m[r * 4 + c] = s;
        }
    return o;
}
Rt toRt(const Mat4& T) { Rt r; for (int i = 0; i < 12; ++i) r.m[i] = T.m[i]; return r; }

// Model::rodrigues2 (Model.cpp:890-932) without the SVD re-orthonormalisation
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

// Eigen::Quaternionf(Matrix3f) as used by the pose log (MaskFusion.cpp:585-590)
static void rotToQuat(const float* R, float* q)
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

// ======================================================================================
// Model
// ======================================================================================
Model::Model(MaskFusion* o, unsigned char id_, float conf, bool enableFillIn, int cap, int ownerRank_, bool ghost)
    : owner(o), ownerRank(ownerRank_), owned(!ghost), id(id_), pose(Mat4::identity()), lastPose(Mat4::identity()), initialC2Winv(Mat4::identity()),
      confidenceThreshold(conf), maxDepth(FLT_MAX), fillIn(enableFillIn), capacity((uint32_t)cap), lastTransform(Mat4::identity())
{
    const int W = o->W, H = o->H, P = o->P;
    cudaStream_t s = o->stream;
    if (ghost) return;
    // in-place clean (default): ONE copy of the store; the ping-pong pair only for the A/B path MFB200_CLEAN_INPLACE=0
    for (int b = 0; b < (o->cleanInPlace ? 1 : 2); ++b) { pos[b].alloc(capacity); col[b].alloc(capacity); nrm[b].alloc(capacity); }
    count.alloc(2); count.zero(s);
    cudaCheck(cudaMallocHost((void**)&hCount, 2 * sizeof(uint32_t)), "cudaMallocHost"); hCount[0] = hCount[1] = 0;
    cudaCheck(cudaMallocHost((void**)&hTrackOut, 40 * sizeof(float)), "cudaMallocHost");
    key.alloc(P); launch_fill_u64(key, KEY_EMPTY, P, s);
    idx.alloc(P); vertConf.alloc(P); colorTime.alloc(P); normRad.alloc(P); cleanTex.alloc((size_t)P); cleanTex.zero(s);
    idx.zero(s); vertConf.zero(s); colorTime.zero(s); normRad.zero(s);
    splatImage.alloc(P); splatVertex.alloc(P); splatNormal.alloc(P); splatTime.alloc(P);
    splatImage.zero(s); splatVertex.zero(s); splatNormal.zero(s); splatTime.zero(s);
    nonBlack.alloc(1); nonBlack.zero(s);
    if (fillIn) { fillImage.alloc(P); fillVertex.alloc(P); fillNormal.alloc(P); fillImage.zero(s); fillVertex.zero(s); fillNormal.zero(s); }
    aflag.alloc(P); abest.alloc(P); aflag.zero(s); abest.zero(s);
    for (int k = 0; k < 3; ++k) { meas[k].alloc(P); meas[k].zero(s); }
    slot.alloc(capacity); launch_fill_u32(slot, 0xffffffffu, capacity, s);
    keep.alloc((size_t)capacity + P);
    size_t nblk = ((size_t)capacity + P + 511) / 512 + 1;
    blockSums.alloc(nblk); blockSums2.alloc(nblk);
    if (o->cleanInPlace) {
        cleanTicket.alloc(4); cleanTicket.zero(s); cleanLoaded.alloc(nblk); cleanLoaded.zero(s);
        cudaCheck(cudaMallocHost((void**)&hCleanStat, 2 * sizeof(uint32_t)), "cudaMallocHost"); hCleanStat[0] = hCleanStat[1] = 0;
    }
    cand.alloc((size_t)capacity + P); candCount.alloc(1); candCount.zero(s);
    for (int l = 0; l < 3; ++l) {
        size_t Pl = (size_t)(W >> l) * (H >> l);
        vmapG[l].alloc(Pl); nmapG[l].alloc(Pl); cloud[l].alloc(Pl); lastDepth[l].alloc(Pl); lastImage[l].alloc(Pl); corres[l].alloc(l == 0 ? Pl + (size_t)o->numSMs * 512 : 1);   // level 0 only: scratch when the correspondences do not fit in shared memory
        vmapG[l].zero(s); nmapG[l].zero(s); lastDepth[l].zero(s); lastImage[l].zero(s);
    }
    if (id_ != 0) for (int l = 0; l < 3; ++l) { validBits[l].alloc(((size_t)(W >> l) * (H >> l) + 31) / 32 + 1); validBits[l].zero(s); }
    lastNextImage2.alloc((size_t)(W >> 2) * (H >> 2)); lastNextImage2.zero(s);
    trackState.alloc(1); trackState.zero(s);
    partial.alloc((size_t)TRACK_MAX_BLOCKS * 128); partial.zero(s);     // 2 x 512 rows x 32 x 16 B (flagged words); flags start at 0: never a valid flag
    dpose.alloc(1); pushPose();
    o->launches += 3;
}

void Model::pushPose()
{
    if (!owned || !dpose.p) return;          // ghosts hold no device state; during construction the buffer appears last
    launch_set_pose(dpose, pose.m, lastPose.m, owner->stream);
    owner->launches += 1;
}

Model::~Model()
{
    if (hCount) cudaFreeHost(hCount);
    if (hTrackOut) cudaFreeHost(hTrackOut);
    if (hCleanStat) cudaFreeHost(hCleanStat);
}

unsigned Model::lastCount()
{
    if (!owned) throw CudaError{"model is owned by another rank (sharded mode): no surfel store here"};
    cudaCheck(cudaMemcpyAsync(hCount, dCount(), sizeof(uint32_t), cudaMemcpyDeviceToHost, owner->stream), "count D2H");
    owner->sync();
    return hCount[0];
}

// Model::initialise (Model.cpp:240-285) fed by MaskFusion::computeFeedbackBuffers (MaskFusion.cpp:187-198)
void Model::initialise(int time)
{
    MaskFusion* o = owner;
    launch_init_model(o->rgb, o->depthRaw, o->depthFilt, o->cam, o->W, o->H, time, o->cfg.maxDepthProcessed, o->initFlagR, o->initFlagF,
                      blockSums, blockSums2, capacity, current(), dCount(), o->stream);
    o->launches += 5;
}

// model side of Model::initICP (Model.cpp:391-409): predicted (or fill-in) maps -> pyramids in the model-global frame
void Model::prepareTracking()
{
    MaskFusion* o = owner;
    cudaStream_t s = o->stream;
    const int W = o->W, H = o->H;
    float denom = (float)((H / 20) * (W / 20));
    float4* v[3] = {vmapG[0].p, vmapG[1].p, vmapG[2].p};
    float4* n[3] = {nmapG[0].p, nmapG[1].p, nmapG[2].p};
    const uint32_t* nb = fillIn ? nonBlack.p : nullptr;
    launch_model_maps(splatVertex, splatNormal, fillIn ? fillVertex.p : splatVertex.p, fillIn ? fillNormal.p : splatNormal.p, nb, denom, W, H,
                      dpose, 6.0f /* maxDepthRGB, RGBDOdometry.cpp:34 */, v, n, lastDepth[0], s);
    o->launches += 1;
    if (validBits[0].p && o->trackValidBits) {
        uint32_t* b3[3] = {validBits[0].p, validBits[1].p, validBits[2].p};
        launch_valid_bits3(n, W, H, b3, s);
        o->launches += 1;
    }
    const bool rgb = o->cfg.rgbOnly || o->cfg.icpWeight < 100;
    if (rgb) {
        launch_intensity_select(splatImage, fillIn ? fillImage.p : splatImage.p, nb, denom, (o->cfg.frameToFrameRGB && fillIn) ? 1 : 0, o->P, lastImage[0], s);
        launch_pyrdown2_pair(lastDepth[0], lastDepth[1], lastDepth[2], lastImage[0], lastImage[1], lastImage[2], W, H, s);
        const float* d3[3] = {lastDepth[0].p, lastDepth[1].p, lastDepth[2].p};
        float4* c3[3] = {cloud[0].p, cloud[1].p, cloud[2].p};
        launch_project_points3(d3, W, H, o->cam, c3, s);
        o->launches += 3;
    }
}

float Model::computeFusionWeight(float weightMultiplier) const
{
    Mat4 diff = mul(rigidInverse(pose), lastPose);       // Model::getLastTransform, Model.h:239
    const float* d = diff.m;
    float R[9] = {d[0], d[1], d[2], d[4], d[5], d[6], d[8], d[9], d[10]};
    float tn = sqrtf((d[3] * d[3] + d[7] * d[7]) + d[11] * d[11]);
    float rv[3]; rodrigues2(R, rv);
    float rn = sqrtf((rv[0] * rv[0] + rv[1] * rv[1]) + rv[2] * rv[2]);
    float weighting = tn > rn ? tn : rn;
    const float largest = 0.01f, minWeight = 0.5f;
    if (weighting > largest) weighting = largest;
    float w = 1.0f - (weighting / largest);
    return (w > minWeight ? w : minWeight) * weightMultiplier;
}

void Model::predictIndices(int time, float depthCutoff, int timeDelta, bool forClean)
{
    MaskFusion* o = owner;
    if (forClean && o->fuseIndexIntoClean) {                          // lazily: see idxDeferred
        idxDeferred = true; idxTime = time; idxDelta = timeDelta; idxDepth = depthCutoff;
        return;
    }
    idxDeferred = false;
    launch_predict_indices(current(), dCount(), dpose, o->cam, o->W, o->H, depthCutoff, time, timeDelta, key, idx, vertConf,
                           colorTime, normRad, forClean ? cleanTex.p : nullptr, confidenceThreshold, o->stream);
    if (forClean) { cleanTexTime = time; cleanTexConf = confidenceThreshold; }
    o->launches += 2;
}

void Model::flushIndex()
{
    if (!idxDeferred) return;
    idxDeferred = false;
    MaskFusion* o = owner;
    launch_predict_indices(current(), dCount(), dpose, o->cam, o->W, o->H, idxDepth, idxTime, idxDelta, key, idx, vertConf, colorTime, normRad, cleanTex.p, confidenceThreshold, o->stream);
    cleanTexTime = idxTime; cleanTexConf = confidenceThreshold;
    o->launches += 2;
}

void Model::fuse(int time, float depthCutoff, float weightMultiplier)
{
    MaskFusion* o = owner;
    flushIndex();
    float md = depthCutoff < maxDepth ? depthCutoff : maxDepth;      // Model.cpp:527 (headless: bounding box empty, N7)
    float4* m[3] = {meas[0].p, meas[1].p, meas[2].p};
    launch_associate(o->rgb, o->depthRaw, o->depthFilt, o->mask, idx, vertConf, normRad, dpose, o->cam, o->W, o->H, md, time,
                     weightMultiplier, id, aflag, abest, m, slot, o->stream);
    launch_fuse_update(aflag, abest, m, slot, o->P, time, current(), o->stream);
    o->launches += 3;
}

void Model::clean(int time, int timeDelta, float /*depthCutoff*/)
{
    MaskFusion* o = owner;
    float4* m[3] = {meas[0].p, meas[1].p, meas[2].p};
    const bool inPlace = o->cleanInPlace;
    // In-place compaction moves only the surfels behind the first removal: 22 us instead of the copy's 68 when removals sit in the young tail
    // of the store (the steady state), but its ticketed hand-over needs ~140 us when most of a 4.5 M store moves (a removal near the front:
    // e.g. the first frames after a map upload).  Both produce the same store, so the choice is free: a large store uses the copy into
    // a second plane set (allocated on first need) for the frame that FOLLOWS one in which more than 40 % of it moved -- the statistic
    // comes back with an asynchronous 8-byte copy and is read without waiting (a stale value only delays the switch).
    bool pingPong = false;
    if (inPlace && capacity >= (1u << 20) && hCleanStat && hCleanStat[1] > 0) {
        const uint32_t first = hCleanStat[0], nb = hCleanStat[1];
        pingPong = first < nb && (uint64_t)(nb - first) * 10 > (uint64_t)nb * 4;
    }
    if (pingPong && !pos[1 - target].p) { pos[1 - target].alloc(capacity); col[1 - target].alloc(capacity); nrm[1 - target].alloc(capacity); }
    int other = (inPlace && !pingPong) ? target : 1 - target, otherCount = 1 - countSel;
    // the pending index projection rides in pass 1 when it uses this call's time gate (always, in the frame schedule)
    const bool fused = idxDeferred && idxTime == time && idxDelta == timeDelta && (size_t)capacity + (size_t)o->P < 0x80000000ull;   // bit 31 of a candidate entry is a flag
    if (!fused) flushIndex();
    IndexFused f{key.p, idx.p, vertConf.p, colorTime.p, normRad.p, cleanTex.p, idxDepth};
    idxDeferred = false;
    // the packed window texels carry two tests evaluated with a time and a confidence threshold: valid for this call when the index
    // map is resolved inside it, or was resolved with the same two values (the frame schedule); else the window reads the images
    const bool packedOK = fused || (cleanTexTime == time && cleanTexConf == confidenceThreshold);
    CleanWindowImages win{packedOK ? cleanTex.p : nullptr, vertConf.p, colorTime.p, idx.p};
    if (++cleanEpoch == 0) ++cleanEpoch;                              // 0 = "never published"
    CleanInPlace ip{cleanTicket.p, cleanLoaded.p, cleanTicket.p + 1, cleanEpoch, pingPong};
    launch_clean(planes(target), planes(other), dCount(), count.p + otherCount, capacity, aflag, m, dpose, o->cam, o->W, o->H,
                 time, timeDelta, confidenceThreshold, o->cfg.outlierCoeff, id, win, o->depthFilt, o->mask, keep, blockSums,
                 cand, candCount, o->stream, fused ? &f : nullptr, inPlace ? &ip : nullptr);
    if (inPlace && capacity >= (1u << 20))
        cudaCheck(cudaMemcpyAsync(hCleanStat, cleanTicket.p + 1, 2 * sizeof(uint32_t), cudaMemcpyDeviceToHost, o->stream), "clean statistic D2H");
    target = other; countSel = otherCount;
    o->launches += fused ? 6 : 5;
}

void Model::combinedPredict(float depthCutoff, int time, int maxTime, int timeDelta)
{
    MaskFusion* o = owner;
    launch_combined_predict(current(), dCount(), dpose, o->cam, o->W, o->H, depthCutoff, confidenceThreshold, time, maxTime,
                            timeDelta, o->rayTab, key, splatImage, splatVertex, splatNormal, splatTime, fillIn ? 1 : 0, o->depthFilt, o->rgb, 0,
                            o->cfg.frameToFrameRGB ? 1 : 0, fillImage, fillVertex, fillNormal, fillIn ? nonBlack.p : nullptr, o->stream, capacity);
    o->launches += 2;
}

// ======================================================================================
// MaskFusion
// ======================================================================================
MaskFusion::MaskFusion(const mf_config& c, int dev, cudaStream_t st) : cfg(c), device(dev)
{
    cudaCheck(cudaSetDevice(dev), "cudaSetDevice");
    cudaDeviceProp prop;
    cudaCheck(cudaGetDeviceProperties(&prop, dev), "cudaGetDeviceProperties");
    numSMs = prop.multiProcessorCount;
    set_num_sms(numSMs);
    if (const char* env = getenv("MFB200_FUSE_INDEX")) fuseIndexIntoClean = env[0] != '0';
    if (const char* env = getenv("MFB200_TRACK_BITS")) trackValidBits = env[0] != '0';
    if (const char* env = getenv("MFB200_CLEAN_INPLACE")) cleanInPlace = env[0] != '0';
    W = c.width; H = c.height; P = W * H;
    if (W % 4 || H % 4) throw CudaError{"width and height must be multiples of 4 (3-level pyramid)"};
    cam = Cam{c.fx, c.fy, c.cx, c.cy};
    ownStream = (st == nullptr);
    if (ownStream) cudaCheck(cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking), "cudaStreamCreate"); else stream = st;
    g_prof = &prof;
    for (int k = 0; k < 2; ++k) { inBuf[k].alloc(packetBytes()); inBuf[k].zero(stream); rgbBuf[k].alloc(P); depthFiltBuf[k].alloc(P); }
    selectSet(0);
    mask.alloc(P); mask.zero(stream);
    cudaCheck(cudaStreamCreateWithFlags(&preStream, cudaStreamNonBlocking), "cudaStreamCreate");
    cudaCheck(cudaEventCreateWithFlags(&preDone, cudaEventDisableTiming), "cudaEventCreate");
    cudaCheck(cudaEventCreateWithFlags(&inputsCopied, cudaEventDisableTiming), "cudaEventCreate");
    cudaCheck(cudaEventCreateWithFlags(&evMain, cudaEventDisableTiming), "cudaEventCreate");
    cudaCheck(cudaEventCreateWithFlags(&evComm, cudaEventDisableTiming), "cudaEventCreate");
    multiOverlap = true;             // validated bit-exact (single process and 2-rank NCCL); MFB200_MULTI_OVERLAP=0 keeps everything on one stream (A/B)
    if (const char* env = getenv("MFB200_MULTI_OVERLAP")) multiOverlap = env[0] != '0';
    for (int l = 0; l < 3; ++l) {
        size_t Pl = (size_t)(W >> l) * (H >> l);
        if (l > 0) depthPyr[l].alloc(Pl);
        vmap[l].alloc(Pl); nmap[l].alloc(Pl); nextImage[l].alloc(Pl); nextGrad[l].alloc(Pl); rgbValid[l].alloc(Pl);
    }
    edgeMap.alloc(P); edgeBinary.alloc(P); edgeBuf.alloc(P); edgeInv.alloc(P);
    dJobs.alloc(TRACK_MAX_JOBS); trackBars.alloc(TRACK_MAX_JOBS * 32);
    cudaCheck(cudaMallocHost((void**)&hJobs, TRACK_MAX_JOBS * sizeof(TrackJob)), "cudaMallocHost");
    initFlagR.alloc(P); initFlagF.alloc(P);
    scratch.alloc((size_t)P * 4);
    rayTab.alloc(P); launch_ray_table(cam, W, H, rayTab, stream);
    if (c.enableMultipleModels) {
        dRes.alloc(1); dRes.zero(stream);
        cudaCheck(cudaMallocHost((void**)&hRes, sizeof(FrameResult)), "cudaMallocHost"); memset(hRes, 0, sizeof(FrameResult));
        cudaCheck(cudaEventCreateWithFlags(&resEvt, cudaEventDisableTiming), "cudaEventCreate");
        poseTable.alloc((size_t)MF_MAX_MODELS * 32); poseTable.zero(stream);
        gathered.alloc((size_t)64 * MF_MAX_MODELS * 32); gathered.zero(stream);
        // component histograms for the worst case (every second pixel its own component): the counts of a frame live on the device
        compModel.alloc(((size_t)P / 2 + 2) * MF_MAX_MODELS); compMask.alloc(((size_t)P / 2 + 2) * 256);
        projKeys.alloc(P); launch_fill_u64(projKeys, KEY_EMPTY, P, stream); projectedIDs.alloc(P); projectedIDs.zero(stream);
        ccL.alloc(P); ccDense.alloc(P); ccLabA.alloc(P); ccLabB.alloc(P); ccArea.alloc((size_t)P + 1); ccBox.alloc(((size_t)P / 2 + 2) * 4); mapToMask.alloc((size_t)P / 2 + 2); absorbId.alloc((size_t)P / 2 + 2);
        maskPixels.alloc(256); ccCounter.alloc(1); ccCounter.zero(stream);
        segTmp.alloc(P); ignoreMap.alloc(P); ignoreMap.zero(stream);
        tblIdToIndex.alloc(256); tblIndexToId.alloc(256); tblIsModel.alloc(256); tblMaskToID.alloc(256); tblIsPerson.alloc(256);
        tblIdToIndex.zero(stream); tblIndexToId.zero(stream); tblIsModel.zero(stream); tblIsPerson.zero(stream);
        tblMaskToID.zero(stream); cudaCheck(cudaMemsetAsync(tblMaskToID.p + 255, 255, 1, stream), "memset");   // maskToID[255] = 255, MfSegmentation.cpp:70-71 (persists across frames)
        maskOverlap.alloc((size_t)MF_MAX_MODELS * 256);
    }
    models.emplace_back(new Model(this, nextID++, c.confGlobal, true, c.capacityGlobal));    // MaskFusion.cpp:80-81
    sync();
}

MaskFusion::~MaskFusion()
{
    cudaStreamSynchronize(stream);
    if (g_prof == &prof) g_prof = nullptr;
    for (cudaEvent_t e : prof.events) cudaEventDestroy(e);
    if (trackDone) cudaEventDestroy(trackDone);
    if (preStream) { cudaStreamSynchronize(preStream); cudaStreamDestroy(preStream); }
    if (preDone) cudaEventDestroy(preDone);
    if (inputsCopied) cudaEventDestroy(inputsCopied);
    if (evMain) cudaEventDestroy(evMain);
    if (evComm) cudaEventDestroy(evComm);
    models.clear();
    inactiveModels.clear();
    if (hJobs) cudaFreeHost(hJobs);
    if (hRes) cudaFreeHost(hRes);
    if (resEvt) cudaEventDestroy(resEvt);
    if (bbFrameReady) cudaEventDestroy(bbFrameReady);
    if (bbMoldDone) cudaEventDestroy(bbMoldDone);
    if (ownStream) cudaStreamDestroy(stream);
}

void MaskFusion::sync()
{
    if (prof.on) prof_mark(stream, nullptr);            // closes the last open interval
    cudaCheck(cudaStreamSynchronize(stream), "cudaStreamSynchronize");
    if (prof.on) prof.resolve();
    finalisePending();
}

// textureRGB / textureDepthMetric (+ FrameData::mask, classIDs) upload into the current input set (MaskFusion.cpp:212-217)
void MaskFusion::uploadInputs(const uint8_t* rgbIn, const float* depthIn, const uint8_t* maskIn, int64_t timestamp, bool onDevice, cudaStream_t s)
{
    if (!s) s = stream;
    cudaMemcpyKind kind = onDevice ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice;
    g_prof = &prof;
    prof_mark(s, onDevice ? "copy_d2d_frame" : "copy_h2d_frame");
    if (onDevice && inputReady) { cudaCheck(cudaStreamWaitEvent(s, inputReady, 0), "cudaStreamWaitEvent"); inputReady = nullptr; }   // caller's producer
    cudaCheck(cudaMemcpyAsync(rgb3, rgbIn, (size_t)P * 3, kind, s), "rgb upload");
    cudaCheck(cudaMemcpyAsync(depthRaw, depthIn, (size_t)P * sizeof(float), kind, s), "depth upload");
    if (cfg.enableMultipleModels) {
        // instance masks + their class list ride with the images: the header is a kernel argument (no staging buffer to keep alive)
        FrameHdr h; memset(&h, 0, sizeof h);
        h.timestamp = timestamp;
        if (maskIn && !classIDs.empty()) {
            if (classIDs.size() > 256) throw CudaError{"more than 256 mask labels"};
            cudaCheck(cudaMemcpyAsync(frameMask, maskIn, (size_t)P, kind, s), "mask upload");
            h.nMasks = (int)classIDs.size();
            for (size_t i = 0; i < classIDs.size(); ++i) h.classIDs[i] = classIDs[i];
        }
        launch_frame_header(h, dHdr, s);
        launches += 1;
    }
    // host inputs belong to the caller again when processFrame returns (the reference uploads synchronously): see processFrame.
    // Device inputs copied on the pre-processing stream: the context stream waits for the copies, so whatever the caller queues
    // there after this call (e.g. the producer of the next frame writing the same buffers) is ordered behind them.
    if (!onDevice) { cudaCheck(cudaEventRecord(inputsCopied, s), "cudaEventRecord"); copyPending = true; }
    else if (s != stream) { cudaCheck(cudaEventRecord(inputsCopied, s), "cudaEventRecord"); cudaCheck(cudaStreamWaitEvent(stream, inputsCopied, 0), "cudaStreamWaitEvent"); }
}

// filterDepth (MaskFusion.cpp:650-657) + the RGBA copy every later pass reads
void MaskFusion::preprocess(cudaStream_t s)
{
    if (!s) s = stream;
    launch_unpack_rgb(rgb3, rgb, P, s);
    launch_bilateral(depthRaw, depthFilt, W, H, s);
    launches += 2;
    frameMapsValid = false; intensityValid = false;
}

void MaskFusion::setFrame(const uint8_t* rgbIn, const float* depthIn, const uint8_t* maskIn, bool onDevice, cudaStream_t s)
{
    if (!s) s = stream;
    // stage-wise test entry (mf_set_frame): the optional mask goes straight into textureMask like the reference's upload
    if (maskIn) cudaCheck(cudaMemcpyAsync(mask, maskIn, (size_t)P, onDevice ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice, s), "mask upload");
    uploadInputs(rgbIn, depthIn, nullptr, 0, onDevice, s);
    preprocess(s);
}

// Model::generateCUDATextures (Model.cpp:350-389): level 0 aliases the filtered depth.
// The mask pyramid of the reference is dead (N10) and not built.
void MaskFusion::generateCUDATextures(cudaStream_t s)
{
    if (!s) s = stream;
    const float* d[3] = {depthFilt, depthPyr[1].p, depthPyr[2].p};
    float4* v3[3] = {vmap[0].p, vmap[1].p, vmap[2].p};
    float4* n3[3] = {nmap[0].p, nmap[1].p, nmap[2].p};
    launch_pyrdown2_f(depthFilt, W, H, depthPyr[1], depthPyr[2], s);
    launch_vmap_nmap3(d, W, H, cam, cfg.depthCutoff, v3, n3, s);
    launches += 2;
    frameMapsValid = true;
}

// frame side of RGBDOdometry::initRGB (RGBDOdometry.cpp:212-215): intensity pyramid, Sobel derivatives, photometric validity; shared by all models
void MaskFusion::frameIntensity(cudaStream_t s)
{
    if (!s) s = stream;
    const bool rgbTerm = cfg.rgbOnly || cfg.icpWeight < 100;
    launch_intensity(rgb, P, nextImage[0], s);
    launch_pyrdown2_u8(nextImage[0], W, H, nextImage[1], nextImage[2], s);
    launches += 2;
    if (rgbTerm) {
        const uint8_t* i3[3] = {nextImage[0].p, nextImage[1].p, nextImage[2].p};
        short2* g3[3] = {nextGrad[0].p, nextGrad[1].p, nextGrad[2].p};
        uint8_t* r3[3] = {rgbValid[0].p, rgbValid[1].p, rgbValid[2].p};
        launch_sobel3(i3, W, H, g3, r3, s);
        launches += 1;
    }
    intensityValid = true;
}

// Model::performTracking for a batch of models: one launch sequence, blockIdx.y = model.
// viaResult: the multi-model schedule reads the tracked poses back inside the frame's FrameResult (applyFrameResult) instead of the
// -static path's dedicated copy behind the tracking kernel.
void MaskFusion::trackModels(const std::vector<Model*>& ms, bool viaResult)
{
    if (ms.empty()) return;
    if ((int)ms.size() > TRACK_MAX_JOBS) throw CudaError{"too many tracked models for one batch"};
    const bool rgbTerm = cfg.rgbOnly || cfg.icpWeight < 100;
    if (!frameMapsValid) generateCUDATextures();
    if ((rgbTerm || cfg.so3) && !intensityValid) frameIntensity();
    unsigned lightMask = 0;          // jobs with a validity bitmask (object models): small share of the tracker's grid
    for (size_t j = 0; j < ms.size(); ++j) {
        Model* m = ms[j];
        if (!viaResult) m->lastPose = m->pose;                       // Model.cpp:430 (multi-model: applied with the result)
        m->prepareTracking();
        TrackJob& J = hJobs[j];
        for (int l = 0; l < 3; ++l) {
            J.vmapC[l] = vmap[l]; J.nmapC[l] = nmap[l]; J.nextImage[l] = nextImage[l]; J.nextGrad[l] = nextGrad[l]; J.rgbValid[l] = rgbValid[l];
            J.vmapG[l] = m->vmapG[l]; J.nmapG[l] = m->nmapG[l]; J.lastDepth[l] = m->lastDepth[l]; J.lastImage[l] = m->lastImage[l];
            J.cloud[l] = m->cloud[l]; J.corres[l] = m->corres[l];
        }
        J.lastNextImage2 = m->lastNextImage2; J.st = m->trackState; J.partial = m->partial; J.bar = trackBars.p + j * 32;
        J.dpose = m->dpose;
        for (int l = 0; l < 3; ++l) J.validBits[l] = (m->validBits[l].p && trackValidBits) ? m->validBits[l].p : nullptr;
        if (J.validBits[0] != nullptr) lightMask |= 1u << j;
    }
    if (preWaitPending) {            // the frame's preprocessing ran on preStream: the tracker is the first consumer on the main stream
        cudaCheck(cudaStreamWaitEvent(stream, preDone, 0), "cudaStreamWaitEvent");
        preWaitPending = false;
    }
    prof_mark(stream, "copy_jobs");
    // hJobs is reused every frame: the next frameBegin first waits (finalisePending) for an event recorded behind this copy
    cudaCheck(cudaMemcpyAsync(dJobs, hJobs, ms.size() * sizeof(TrackJob), cudaMemcpyHostToDevice, stream), "jobs upload");
    launches += launch_tracking(dJobs, (int)ms.size(), W, H, cam, cfg.rgbOnly != 0, cfg.icpWeight, cfg.pyramid != 0, cfg.fastOdom != 0,
                                cfg.so3 != 0, numSMs, trackBars, stream, lightMask);
    prof_mark(stream, "copy_pose_d2h");
    for (Model* m : ms) {
        if (!viaResult)
            cudaCheck(cudaMemcpyAsync(m->hTrackOut, (const char*)m->trackState.p + offsetof(TrackState, out), 40 * sizeof(float),
                                      cudaMemcpyDeviceToHost, stream), "pose D2H");
        if (cfg.so3)   // std::swap(lastNextImage, nextImage) (RGBDOdometry.cpp:484-488): only level 2 is ever read
            cudaCheck(cudaMemcpyAsync(m->lastNextImage2, nextImage[2], (size_t)(W >> 2) * (H >> 2), cudaMemcpyDeviceToDevice, stream), "so3 swap");
    }
    if (viaResult) return;
    if (!trackDone) cudaCheck(cudaEventCreateWithFlags(&trackDone, cudaEventDisableTiming), "cudaEventCreate");
    cudaCheck(cudaEventRecord(trackDone, stream), "cudaEventRecord");
    pendingModels = ms; pendingTrack = true;
}

// host copies of the tracked poses (and, multi-model, of everything else a frame decides): waits for an event that lies MID-frame
// (behind the tracking kernel / the vote kernel), not for the rest of the frame
void MaskFusion::finalisePending()
{
    if (pendingTrack) {
        cudaCheck(cudaEventSynchronize(trackDone), "cudaEventSynchronize");
        for (Model* m : pendingModels) {
            memcpy(m->pose.m, m->hTrackOut, 16 * sizeof(float));
            memcpy(m->lastTransform.m, m->hTrackOut + 16, 16 * sizeof(float));
        }
        pendingTrack = false; pendingModels.clear();
    }
    if (pendingResult) applyFrameResult();
    if (pendingLog) { pendingLog = false; logPoses(pendingTimestamp); }
}

// MaskFusion.cpp:577-592: one pose-log entry per model per frame
void MaskFusion::logPoses(int64_t timestamp)
{
    Model* g = models[0].get();
    for (size_t i = 0; i < models.size(); ++i) {
        Model* m = models[i].get();
        Mat4 T = (i == 0) ? g->pose : mul(g->pose, rigidInverse(m->pose));     // MaskFusion.cpp:581-583
        float R[9] = {T.m[0], T.m[1], T.m[2], T.m[4], T.m[5], T.m[6], T.m[8], T.m[9], T.m[10]}, q[4];
        rotToQuat(R, q);
        double e[8] = {(double)timestamp, T.m[3], T.m[7], T.m[11], q[0], q[1], q[2], q[3]};
        m->poseLog.insert(m->poseLog.end(), e, e + 8);
    }
}

void MaskFusion::predict()
{
    for (auto& m : models) if (m->owned) m->combinedPredict(cfg.maxDepthProcessed, tick, tick, cfg.timeDelta);   // MaskFusion.cpp:616-628
}

// GlobalProjection::project (GlobalProjection.cpp:43-107): all models into one depth-tested key image; the key's low word
// is (model list index << 26 | surfel id), i.e. the reference's draw order.  The ID image stays on the device.
void MaskFusion::globalProjection() { segTables(); projectLocal(); projectResolve(); }

// id <-> list-index tables of this frame's model list, as a kernel argument (no staging buffer, no synchronisation)
void MaskFusion::segTables()
{
    if (models.size() > MF_MAX_MODELS - 1) throw CudaError{"global projection / segmentation support up to 63 models"};
    SegTables t; memset(&t, 0, sizeof t);
    for (size_t i = 0; i < models.size(); ++i) { t.idToIndex[models[i]->id] = (uint8_t)i; t.indexToId[i] = models[i]->id; t.isModel[models[i]->id] = 1; }
    launch_seg_tables(t, tblIdToIndex, tblIndexToId, tblIsModel, stream);
    launches += 1;
}

// local half: every model whose surfels live here goes into the key image (ghost models arrive through the min all-reduce)
void MaskFusion::projectLocal()
{
    if (models.size() > MF_MAX_MODELS - 1) throw CudaError{"global projection supports up to 63 models"};
    for (size_t i = 0; i < models.size(); ++i) {
        Model* m = models[i].get();
        if (!m->owned) continue;
        launch_splat_project_only(m->current(), m->dCount(), m->dpose, cam, W, H, cfg.depthCutoff, 12.0f /* :61 */, tick, tick,
                                  cfg.timeDelta, (uint32_t)i << 26, rayTab, projKeys, stream, m->capacity);
        launches += 1;
    }
}

void MaskFusion::projectResolve()
{
    launch_proj_resolve(projKeys, P, tblIndexToId, projectedIDs, stream);
    launches += 1;
}

// MfSegmentation::performSegmentation (MfSegmentation.cpp:83-538) with the CPU tail on the GPU (mf_seg.cu) and NO host round trip:
// the number of components, the number of masks and their classes are read by the kernels from device memory; the mask -> model vote
// (:433-492) is a kernel; its decision (new label?) travels to the host inside the FrameResult and is applied by applyFrameResult().
void MaskFusion::performSegmentation(bool allowNew)
{
    const int nModels = (int)models.size();
    if (!frameMapsValid) generateCUDATextures();
    // edge-ness -> threshold -> close -> invert (MfSegmentation.cpp:149-208)
    launch_geometric_edges(vmap[0], nmap[0], W, H, cfg.segWeightDistance, cfg.segWeightConvexity, cfg.segThreshold, edgeMap, edgeBinary, stream);
    launch_morph_close_invert(edgeBinary, edgeBuf, W, H, cfg.segMorphEdgeRadius, cfg.segMorphEdgeIterations, edgeInv, stream);
    launches += 2 + 2 * cfg.segMorphEdgeIterations;
    // ignore map (:221-235)
    launch_person_table(dHdr, personClassID, tblIsPerson, stream);
    launch_apply_ignore(frameMask, tblIsPerson, dHdr, P, ignoreMap, edgeInv, stream);
    // connected components + 5 edge-removal sweeps (:238-291)
    launch_cc(edgeInv, W, H, ccL, ccDense, ccLabA, ccArea, ccBox, ccCounter, stream);
    launch_remove_edges(ccLabA, ccLabB, depthRaw, ccArea, W, H, 5, stream);
    int* lab = ccLabB;                                     // odd number of sweeps ends in B
    // overlap histograms (:303-346)
    launch_clear_hist(ccCounter, dHdr, nModels, compModel, compMask, stream);
    maskPixels.zero(stream);
    launch_seg_hist(lab, projectedIDs, frameMask, P, tblIdToIndex, nModels, dHdr, compModel, compMask, stream);
    launch_component_map(ccCounter, ccArea, compModel, compMask, nModels, dHdr, tblIndexToId, minMappedComponentSize, mapToMask, absorbId, maskPixels, stream);
    launch_seg_assign(lab, mapToMask, ignoreMap, P, segTmp, stream);
    launches += 16;
    // closing of the mask-id image with an elliptic element (:424-426, inside `if (nMasks)`); edgeBuf is free again at this point
    launches += launch_morph_close_ellipse(segTmp, edgeBuf, W, H, cfg.segMorphMaskRadius, cfg.segMorphMaskIterations, dHdr, stream);
    // mask -> model vote (:433-492)
    cudaCheck(cudaMemsetAsync(maskOverlap, 0, (size_t)nModels * 256 * sizeof(unsigned), stream), "memset");
    launch_mask_overlap(segTmp, projectedIDs, tblIdToIndex, tblIsModel, P, maskOverlap, stream);
    VoteParams vp; memset(&vp, 0, sizeof vp);
    vp.nModels = nModels; vp.allowNew = allowNew ? 1 : 0; vp.personClassID = personClassID;
    vp.minNew = (unsigned)(size_t)(cfg.minRelSizeNew * (size_t)P); vp.maxNew = (unsigned)(size_t)(cfg.maxRelSizeNew * (size_t)P);
    vp.minMaskModelOverlap = minMaskModelOverlap; vp.nextModelID = getNextModelID(false);
    for (int i = 0; i < nModels; ++i) { vp.modelClass[i] = models[i]->classID; vp.modelID[i] = models[i]->id; }
    launch_vote(dHdr, vp, maskPixels, maskOverlap, ccCounter, tblMaskToID, dRes, stream);
    launch_seg_final(segTmp, lab, mapToMask, absorbId, tblMaskToID, ccBox, P, W, mask, stream);      // writes textureMask directly (:297)
    launches += 3;
}

// MaskFusion::getNextModelID (MaskFusion.cpp:712-730)
unsigned char MaskFusion::getNextModelID(bool assign)
{
    unsigned char next = nextID;
    if (assign) {
        if (models.size() == 256) throw CudaError{"getNextModelID(): maximum amount of models is already in use (256)"};
        while (true) {
            nextID++;
            bool occupied = false;
            for (auto& m : models) if (nextID == m->id) occupied = true;
            if (!occupied) break;
        }
    }
    return next;
}

// MaskFusion::spawnObjectModel + moveNewModelToList (MaskFusion.cpp:671-690)
Model* MaskFusion::spawnObjectModel()
{
    Model* g = models[0].get();
    // sharded mode: the new store goes to the least-loaded rank; every rank evaluates the same rule on the same replicated list
    int64_t loads[64] = {0};
    // load of a rank = (number of models it tracks, their surfel capacity): every tracked model walks the whole image in the tracker, so
    // the model count dominates; the capacity breaks ties (the background's store is the big one)
    for (auto& m : models) loads[m->ownerRank] += ((int64_t)1 << 32) + m->capacity;
    const int ownerRank = world > 1 ? pickOwner(loads, world) : 0;
    models.emplace_back(new Model(this, getNextModelID(true), cfg.confObject, false, cfg.capacityObject, ownerRank, ownerRank != rank));
    Model* nm = models.back().get();
    if (nm->owned) {
        // newModel->getFrameOdometry().initFirstRGB(textureRGB)
        if (!intensityValid) {
            launch_intensity(rgb, P, nextImage[0], stream);
            launch_pyrdown2_u8(nextImage[0], W, H, nextImage[1], nextImage[2], stream);
            launches += 3; intensityValid = true;
        }
        cudaCheck(cudaMemcpyAsync(nm->lastNextImage2, nextImage[2], (size_t)(W >> 2) * (H >> 2), cudaMemcpyDeviceToDevice, stream), "initFirstRGB");
    }
    nm->makeStatic(g->pose);
    return nm;
}

int MaskFusion::pickOwner(const int64_t* loads, int world)
{
    int best = 0;
    for (int r = 1; r < world; ++r) if (loads[r] <= loads[best]) best = r;
    return best;
}

void MaskFusion::configureShard(int rank_, int world_)
{
    if (world_ < 1 || world_ > 64 || rank_ < 0 || rank_ >= world_) throw CudaError{"configureShard: need 0 <= rank < world <= 64"};
    if (tick != 1) throw CudaError{"configureShard: must be called before the first frame"};
    if (world_ > 1 && !cfg.enableMultipleModels) throw CudaError{"configureShard: a -static run has one model and does not shard (run replicas instead)"};
    rank = rank_; world = world_;
    if (rank != 0) {          // the background model lives on rank 0
        Model* g = models[0].get();
        models[0].reset(new Model(this, g->id, cfg.confGlobal, true, cfg.capacityGlobal, 0, true));
        sync();
    }
}

// communicator of the shards (mf_shard_comm_init): from here on the three exchanges of a frame are NCCL calls on the context's stream
void MaskFusion::initShardComm(const unsigned char* id128, int rank_, int world_)
{
    if (world == 1 && world_ > 1) configureShard(rank_, world_);
    if (rank_ != rank || world_ != world) throw CudaError{"initShardComm: rank / world differ from configureShard"};
    cudaCheck(cudaSetDevice(device), "cudaSetDevice");
    shard.init(id128, rank_, world_);
    shardNccl = true;
}

extern "C" int mf_backbone_mold(struct mf_backbone* h, const void* d_rgba, int W, int H);
extern "C" int mf_backbone_forward(struct mf_backbone* h, const void* d_input);
extern "C" void* mf_backbone_input_buffer(struct mf_backbone* h);
extern "C" void* mf_backbone_stream(struct mf_backbone* h);

void MaskFusion::attachBackbone(void* bb, int everyK)
{
    backbone = bb; backboneEvery = bb ? (everyK > 0 ? everyK : 1) : 0;
    if (bb && !bbFrameReady) {
        cudaCheck(cudaEventCreateWithFlags(&bbFrameReady, cudaEventDisableTiming), "cudaEventCreate");
        cudaCheck(cudaEventCreateWithFlags(&bbMoldDone, cudaEventDisableTiming), "cudaEventCreate");
    }
}

// every k-th frame: RGBA image of this frame -> letter-boxed network input -> backbone forward, all on the backbone's stream
void MaskFusion::runBackbone(cudaStream_t producer)
{
    if (!backbone || backboneEvery <= 0 || (tick % backboneEvery) != 0) return;
    if (!producer) producer = stream;
    mf_backbone* bb = (mf_backbone*)backbone;
    cudaStream_t bs = (cudaStream_t)mf_backbone_stream(bb);
    cudaCheck(cudaEventRecord(bbFrameReady, producer), "cudaEventRecord");          // the RGBA copy of this frame exists
    cudaCheck(cudaStreamWaitEvent(bs, bbFrameReady, 0), "cudaStreamWaitEvent");
    if (mf_backbone_mold(bb, rgb, W, H) != 0) throw CudaError{"backbone: mold_inputs failed"};
    cudaCheck(cudaEventRecord(bbMoldDone, bs), "cudaEventRecord");                   // the frame's image is free again once the input is molded
    bbMoldPending = true;
    if (mf_backbone_forward(bb, mf_backbone_input_buffer(bb)) != 0) throw CudaError{"backbone: forward failed"};
}

// the models of the frame in flight as the lifecycle kernels see them
LifeParams MaskFusion::lifeParams() const
{
    LifeParams lp; memset(&lp, 0, sizeof lp);
    if (models.size() > MF_MAX_MODELS) throw CudaError{"more than 64 models"};
    lp.nModels = (int)models.size(); lp.rank = rank;
    for (size_t i = 0; i < models.size(); ++i) {
        const Model* m = models[i].get();
        LifeModel& L = lp.m[i];
        L.tracked = m->tracked ? 1 : 0; L.owned = m->owned ? 1 : 0; L.ownerRank = m->ownerRank;
        if (m->owned) {
            L.dpose = m->dpose.p; L.count = m->dCount();
            L.trackOut = reinterpret_cast<const float*>(reinterpret_cast<const char*>(m->trackState.p) + offsetof(TrackState, out));
        }
        memcpy(L.initialC2Winv, m->initialC2Winv.m, sizeof L.initialC2Winv);
    }
    return lp;
}

// external transport (no NCCL communicator): this rank's rows to the host / the gathered rows of all ranks back to the device
void MaskFusion::getShardPoses(float* out)
{
    cudaCheck(cudaMemcpyAsync(out, poseTable.p, (size_t)MF_MAX_MODELS * 32 * sizeof(float), cudaMemcpyDeviceToHost, stream), "rows D2H");
    cudaCheck(cudaStreamSynchronize(stream), "cudaStreamSynchronize");
}
void MaskFusion::setShardPoses(const float* all)
{
    cudaCheck(cudaMemcpyAsync(gathered.p, all, (size_t)world * MF_MAX_MODELS * 32 * sizeof(float), cudaMemcpyHostToDevice, stream), "rows H2D");
    cudaCheck(cudaStreamSynchronize(stream), "cudaStreamSynchronize");
}

bool MaskFusion::processFrame(const uint8_t* rgbIn, const float* depthIn, int64_t timestamp, const uint8_t* maskIn, const Mat4* inPose,
                              float weightMultiplier, bool bootstrap, bool onDevice)
{
    if (world > 1 && !shardNccl) throw CudaError{"processFrame: this context is one shard of several without a communicator; call mf_shard_comm_init, or drive the frame_begin/project/end phases and move the rows / keys yourself"};
    frameBegin(rgbIn, depthIn, timestamp, maskIn, inPose, bootstrap, onDevice);
    frameProject();
    frameEnd(weightMultiplier);
    if (copyPending) {               // the caller's host buffers are free again on return, as with the reference's synchronous upload
        cudaCheck(cudaEventSynchronize(inputsCopied), "cudaEventSynchronize");
        copyPending = false;
    }
    return false;
}

// MaskFusion.cpp:200-276: upload, filter, first-frame initialisation or tracking of every model whose store lives here
void MaskFusion::frameBegin(const uint8_t* rgbIn, const float* depthIn, int64_t timestamp, const uint8_t* maskIn, const Mat4* inPose, bool bootstrap,
                            bool onDevice)
{
    const bool multi = cfg.enableMultipleModels != 0;
    if (world > 1 && inPose) throw CudaError{"sharded mode tracks every frame (no external poses)"};
    if (multi && bootstrap && inPose) throw CudaError{"bootstrap poses are supported by the -static schedule only"};
    finalisePending();                                  // previous frame: tracked poses, pose log, (multi) inactivations and the deferred spawn
    fTimestamp = timestamp; fHasPose = inPose != nullptr; if (inPose) fInPose = *inPose; fBootstrap = bootstrap;
    const bool tracking = tick > 1 && (bootstrap || !inPose);
    // -static tracking frames: upload + bilateral + pyramids + maps + intensity/Sobel of THIS frame go to preStream and into the other
    // input set, so they run next to the surfel passes of the previous frame that are still queued on the main stream (those read the
    // previous frame's images; the copy engine and the issue-bound bilateral overlap well with the HBM-bound clean/scatter).
    // Safe without further events: finalisePending() above has waited for the previous frame's tracker, the last reader of the maps.
    const bool overlap = !multi && world == 1 && tracking && !prof.on;
    // multi-model frames (MFB200_MULTI_OVERLAP=1): the same idea -- the frame's inputs and preprocessing run on preStream next to the
    // previous frame's fusion / clean / prediction tail.  With a communicator ALL collectives are issued on preStream (one stream per
    // communicator: their order is the same on every rank by construction) and tied to the main stream by events.
    const bool moverlap = multi && multiOverlap && tracking && !prof.on && (world == 1 || shardNccl);
    if (overlap) {
        selectSet(curSet ^ 1);
        uploadInputs(rgbIn, depthIn, nullptr, timestamp, onDevice, preStream);
        preprocess(preStream);
        generateCUDATextures(preStream);
        if (cfg.rgbOnly || cfg.icpWeight < 100 || cfg.so3) frameIntensity(preStream);
        cudaCheck(cudaEventRecord(preDone, preStream), "cudaEventRecord");
        preWaitPending = true;
    } else if (moverlap) {
        selectSet(curSet ^ 1);
        if (spawnedInApply) {
            // the spawn that finalisePending() just carried out reads the previous frame's intensity pyramid (initFirstRGB), which this
            // frame's preprocessing overwrites: on such (rare) frames the preprocessing waits for the main stream
            cudaCheck(cudaEventRecord(evMain, stream), "cudaEventRecord");
            cudaCheck(cudaStreamWaitEvent(preStream, evMain, 0), "cudaStreamWaitEvent");
            spawnedInApply = false;
        }
        if (bbMoldPending) { cudaCheck(cudaStreamWaitEvent(preStream, bbMoldDone, 0), "cudaStreamWaitEvent"); bbMoldPending = false; }
        if (shardNccl) {
            if (rank == 0) uploadInputs(rgbIn, depthIn, maskIn, timestamp, onDevice, preStream);
            shard.broadcast(inBuf[curSet].p, packetBytes(), 0, preStream);
        } else
            uploadInputs(rgbIn, depthIn, maskIn, timestamp, onDevice, preStream);
        preprocess(preStream);
        runBackbone(preStream);
        generateCUDATextures(preStream);
        if (cfg.rgbOnly || cfg.icpWeight < 100 || cfg.so3) frameIntensity(preStream);
        cudaCheck(cudaEventRecord(preDone, preStream), "cudaEventRecord");
        preWaitPending = true;
    } else {
        // multi-model: the previous frame's input set stays intact (a spawn decided by that frame has just been carried out from it)
        if (multi) selectSet(curSet ^ 1);
        spawnedInApply = false;
        if (bbMoldPending) { cudaCheck(cudaStreamWaitEvent(stream, bbMoldDone, 0), "cudaStreamWaitEvent"); bbMoldPending = false; }   // an older frame's image is being read
        if (shardNccl) {
            // object-sharded: rank 0 holds the loader; the frame packet (images + mask + header, one buffer) goes to every rank over NVLink
            if (rank == 0) uploadInputs(rgbIn, depthIn, maskIn, timestamp, onDevice);
            prof_mark(stream, "nccl_broadcast_packet");
            shard.broadcast(inBuf[curSet].p, packetBytes(), 0, stream);
        } else
            uploadInputs(rgbIn, depthIn, maskIn, timestamp, onDevice);     // -static: textureMask stays all zero (MaskFusion.cpp:223-230)
        preprocess();
        runBackbone();
    }
    commOnPre = moverlap && shardNccl;
    Model* g = models[0].get();
    fTracked = false;
    for (auto& m : models) m->tracked = false;
    if (tick == 1) {
        if (g->owned) {
            g->initialise(tick);
            // globalModel->getFrameOdometry().initFirstRGB (MaskFusion.cpp:238)
            launch_intensity(rgb, P, nextImage[0], stream);
            launch_pyrdown2_u8(nextImage[0], W, H, nextImage[1], nextImage[2], stream);
            cudaCheck(cudaMemcpyAsync(g->lastNextImage2, nextImage[2], (size_t)(W >> 2) * (H >> 2), cudaMemcpyDeviceToDevice, stream), "initFirstRGB");
            launches += 3;
        }
    } else if (tracking) {
        if (!frameMapsValid) generateCUDATextures();
        // MaskFusion.cpp:247-276: the global model and every tracked object share one batched launch sequence
        std::vector<Model*> tracked;
        for (size_t i = 0; i < models.size(); ++i) {
            Model* m = models[i].get();
            m->tracked = (i == 0 || m->nonstatic || cfg.trackAllModels);
            if (m->owned && m->tracked) tracked.push_back(m);
        }
        trackModels(tracked, multi);
        if (preWaitPending) {            // no model is tracked here (a shard without stores): the later passes still read this frame's maps
            cudaCheck(cudaStreamWaitEvent(stream, preDone, 0), "cudaStreamWaitEvent");
            preWaitPending = false;
        }
        fTracked = true;
        if (multi) {
            // pose rows of the models tracked here; with a communicator every rank receives every rank's rows (all-gather)
            launch_pack_rows(lifeParams(), poseTable, stream);
            launches += 1;
            if (shardNccl) {
                prof_mark(stream, "nccl_allgather_poses");
                if (commOnPre) {
                    cudaCheck(cudaEventRecord(evMain, stream), "cudaEventRecord"); cudaCheck(cudaStreamWaitEvent(preStream, evMain, 0), "cudaStreamWaitEvent");
                    shard.allGatherFloats(poseTable, gathered, (size_t)MF_MAX_MODELS * 32, preStream);
                    cudaCheck(cudaEventRecord(evComm, preStream), "cudaEventRecord"); cudaCheck(cudaStreamWaitEvent(stream, evComm, 0), "cudaStreamWaitEvent");
                } else
                    shard.allGatherFloats(poseTable, gathered, (size_t)MF_MAX_MODELS * 32, stream);
            }
        } else if (bootstrap && inPose) finalisePending();     // -static bootstrap: the host composes the tracked pose with the given one
    }
}

// MaskFusion.cpp:257-290 after the poses are known everywhere: inactivation, static poses (device side), local part of the ID projection
void MaskFusion::frameProject()
{
    if (!fTracked) return;
    if (!cfg.enableMultipleModels) {
        if (fBootstrap && fHasPose) { Model* g = models[0].get(); g->overridePose(mul(g->pose, fInPose)); }
        return;
    }
    launch_lifecycle(lifeParams(), world > 1 ? gathered.p : poseTable.p, dRes, stream);
    launches += 1;
    projectLocal();                                                            // :289-290
}

// MaskFusion.cpp:290-607: segmentation (replicated: every rank holds the same merged key image), fusion of the local stores, prediction
void MaskFusion::frameEnd(float weightMultiplier)
{
    const bool multi = cfg.enableMultipleModels != 0;
    Model* g = models[0].get();
    fWeight = weightMultiplier;
    if (tick > 1) {
        if (fTracked) {
            if (multi) {
                if (shardNccl) {
                    prof_mark(stream, "nccl_allreduce_keys");
                    if (commOnPre) {
                        cudaCheck(cudaEventRecord(evMain, stream), "cudaEventRecord"); cudaCheck(cudaStreamWaitEvent(preStream, evMain, 0), "cudaStreamWaitEvent");
                        shard.allReduceMinU64(projKeys, (size_t)P, preStream);
                        cudaCheck(cudaEventRecord(evComm, preStream), "cudaEventRecord"); cudaCheck(cudaStreamWaitEvent(stream, evComm, 0), "cudaStreamWaitEvent");
                    } else
                        shard.allReduceMinU64(projKeys, (size_t)P, stream);
                }
                segTables();
                projectResolve();
                if (spawnOffset < cfg.modelSpawnOffset) spawnOffset++;
                performSegmentation(spawnOffset >= cfg.modelSpawnOffset);
                // what the host needs from this frame, in one copy behind the vote kernel; picked up by the next finalisePending()
                prof_mark(stream, "copy_result_d2h");
                cudaCheck(cudaMemcpyAsync(hRes, dRes.p, sizeof(FrameResult), cudaMemcpyDeviceToHost, stream), "result D2H");
                cudaCheck(cudaEventRecord(resEvt, stream), "cudaEventRecord");
                pendingResult = true;
                for (size_t i = 1; i < models.size(); ++i) models[i]->maxDepth = 30.0f + 30.0f * 1.2f;   // getMaxDepth(30, 30), :292,337-341
                for (size_t i = 1; i < models.size(); ++i) {                           // :369-374
                    float f = (float)models[i]->age / 25.0f;
                    models[i]->confidenceThreshold = f < 4.5f ? f : 4.5f;
                }
            }
        } else {
            g->overridePose(fInPose);
        }
        if (!cfg.rgbOnly) {
            for (auto& m : models) if (m->owned) m->predictIndices(tick, cfg.maxDepthProcessed, cfg.timeDelta, false);
            for (auto& m : models) if (m->owned) m->fuse(tick, cfg.depthCutoff, weightMultiplier);
            for (auto& m : models) if (m->owned) m->predictIndices(tick, cfg.maxDepthProcessed, cfg.timeDelta);
            for (auto& m : models) if (m->owned) m->clean(tick, cfg.timeDelta, cfg.maxDepthProcessed);
        }
    }
    predict();          // MaskFusion.cpp:569 (the call at :423 is dead in open-loop mode: its outputs are overwritten here)
    fTick = tick;
    tick++;
    if (pendingResult) { /* the pose-log entry is written by applyFrameResult, after the spawn it may carry out */ }
    else if (pendingTrack) { pendingLog = true; pendingTimestamp = fTimestamp; }    // -static: the entry is written when the pose arrives
    else logPoses(fTimestamp);
    for (auto& m : models) m->age++;
}

// Everything the host learns from a multi-model frame, applied at the start of the next one (or by any query in between), in the
// order of the reference's frame: tracked poses (Model.cpp:427-447) -> inactivation (MaskFusion.cpp:268-272, 699-713) -> spawn of the
// model MfSegmentation asked for and its first fusion FROM THAT FRAME'S DATA (:313-353; the input set, textureMask and the intensity
// pyramid of the frame are still untouched) -> pose log (:577-592).  Identical on every rank of a sharded run (replicated inputs).
void MaskFusion::applyFrameResult()
{
    cudaCheck(cudaEventSynchronize(resEvt), "cudaEventSynchronize");
    pendingResult = false;
    const FrameResult& R = *hRes;
    for (size_t i = 0; i < models.size(); ++i) {
        Model* m = models[i].get();
        if (!m->tracked && i == 0) continue;
        m->lastPose = m->pose;
        memcpy(m->pose.m, R.poses[i], 16 * sizeof(float));
        if (m->tracked) memcpy(m->lastTransform.m, R.poses[i] + 16, 16 * sizeof(float));
    }
    for (size_t i = models.size(); i-- > 1;) {
        if (!R.dead[i]) continue;
        Model* m = models[i].get();
        const unsigned cnt = m->owned ? R.deadCount[i] : modelKeepMinSurfels;
        if (!enableSmartModelDelete || (cnt >= modelKeepMinSurfels && m->confidenceThreshold > modelKeepConfThreshold)) {
            if (m->owned) { launch_set_count(m->dCount(), cnt, stream); launches += 1; }       // the store as it was when the model left
            inactiveModels.push_back(std::move(models[i]));
        }
        models.erase(models.begin() + i);
    }
    if (R.hasNewLabel) {                                                                       // :313-334
        Model* nm = spawnObjectModel();
        spawnedInApply = true;
        spawnOffset = 0;
        nm->classID = R.newClassID;
        nm->maxDepth = 30.0f + 30.0f * 1.2f;
        if (nm->owned) {                                                                       // :344-353, with the frame's own tick / weight
            const int t = fTick;
            nm->predictIndices(t, cfg.maxDepthProcessed, cfg.timeDelta);
            nm->fuse(t, cfg.maxDepthProcessed, 100.0f);
            nm->clean(t, cfg.timeDelta, cfg.maxDepthProcessed);
            nm->confidenceThreshold = 0.0f;                                                    // age 0 (:369-374)
            if (!cfg.rgbOnly) {
                nm->predictIndices(t, cfg.maxDepthProcessed, cfg.timeDelta, false);
                nm->fuse(t, cfg.depthCutoff, fWeight);
                nm->predictIndices(t, cfg.maxDepthProcessed, cfg.timeDelta);
                nm->clean(t, cfg.timeDelta, cfg.maxDepthProcessed);
            }
            nm->combinedPredict(cfg.maxDepthProcessed, t, t, cfg.timeDelta);
        }
        nm->confidenceThreshold = 0.0f;
        nm->age = 1;
    }
    logPoses(R.timestamp);
}

}  // namespace mfb
