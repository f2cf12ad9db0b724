// mf_kernels.h -- host-callable launchers of the sm_100a kernels and the device-side
// structures they share with the host classes (mf_host.cu).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "mf_common.cuh"

namespace mfb {

struct SurfelPlanes { float4* pos; float4* col; float4* nrm; };

// photometric correspondence record (reference: DataTerm, Core/Cuda/types.cuh:75-81)
struct DataTerm { short2 zero; short2 one; float diff; int valid; };

#define TRACK_MAX_JOBS 32        // tracked models of one batched launch (configs[4]: 16 objects + background)
#define TRACK_MAX_BLOCKS 1024
struct TrackPoses { float p[TRACK_MAX_JOBS][16]; };

// Pose of one model as the kernels see it: DEVICE resident, so that the passes after tracking (index map, association,
// fusion, clean, splat) are enqueued without waiting for the tracked pose on the host.  Written by the tracking kernel's
// epilogue or by k_set_pose (host-driven poses); both run the same derivePose().
//   pose  : Model::pose, camera -> model frame, row-major [R|t]
//   tinv  : pose.inverse() as [R^T | -R^T t] (rule R-INV)
//   fusionW : Model::computeFusionWeight(1.0), Model.cpp:449-464 (the caller's weightMultiplier is applied by the kernel)
struct DevPose { Rt pose; Rt tinv; float fusionW; float pad[7]; };

#ifdef __CUDACC__
// Model::rodrigues2 (Model.cpp:890-932) without the SVD re-orthonormalisation (rule R-SVD); R row-major 3x3
__device__ inline void rodrigues2Dev(const float* R, float* out)
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

// pose (row-major 4x4, rigid) + the pose before this frame's update -> everything the surfel passes read.
// One thread.  tinv: [R^T | -R^T t] in fp32 with the operation order of the host rule (R-INV);
// fusionW: max(1 - max(|t|, |rotation vector|) of (pose^-1 * lastPose) / 0.01, 0.5)  (Model.cpp:449-464)
__device__ inline void derivePose(DevPose* d, const float* m, const float* last)
{
    float inv[16];
    {
        const float R[9] = {m[0], m[1], m[2], m[4], m[5], m[6], m[8], m[9], m[10]}, t[3] = {m[3], m[7], m[11]};
        for (int r = 0; r < 3; ++r) {
            for (int c = 0; c < 3; ++c) inv[r * 4 + c] = R[c * 3 + r];
            inv[r * 4 + 3] = -((R[0 * 3 + r] * t[0] + R[1 * 3 + r] * t[1]) + R[2 * 3 + r] * t[2]);
        }
        inv[12] = 0.f; inv[13] = 0.f; inv[14] = 0.f; inv[15] = 1.f;
    }
    for (int k = 0; k < 12; ++k) { d->pose.m[k] = m[k]; d->tinv.m[k] = inv[k]; }
    // diff = pose^-1 * lastPose, full 4x4 product in the order of the host routine (s += A[r][k] * B[k][c], k = 0..3)
    float diff[16];
    for (int r = 0; r < 4; ++r)
        for (int c = 0; c < 4; ++c) {
            float s = 0;
            for (int k = 0; k < 4; ++k) s += inv[r * 4 + k] * last[k * 4 + c];
            diff[r * 4 + c] = s;
        }
    const float R[9] = {diff[0], diff[1], diff[2], diff[4], diff[5], diff[6], diff[8], diff[9], diff[10]};
    float tn = sqrtf((diff[3] * diff[3] + diff[7] * diff[7]) + diff[11] * diff[11]);
    float rv[3]; rodrigues2Dev(R, rv);
    float rn = sqrtf((rv[0] * rv[0] + rv[1] * rv[1]) + rv[2] * rv[2]);
    float weighting = tn > rn ? tn : rn;
    const float largest = 0.01f, minWeight = 0.5f;
    if (weighting > largest) weighting = largest;
    float w = 1.0f - (weighting / largest);
    d->fusionW = w > minWeight ? w : minWeight;
}
#endif

// ---- device-resident frame bookkeeping of the multi-model schedule: nothing in a frame waits for the host ----
// FrameHdr: what the loader knows about the frame besides the images; in the object-sharded mode it is the tail of the broadcast
// frame packet, so the ranks that never saw the host inputs read it where the kernels read it: on the device.
struct FrameHdr { long long timestamp; int nMasks; int pad; int classIDs[256]; };          // nMasks == 0: the frame carries no instance masks
// FrameResult: everything the host learns from a frame (one asynchronous copy behind the vote kernel; read at the START of the next
// frame): the spawn decision of MfSegmentation, which tracked models jumped > 0.2 m (MaskFusion.cpp:268-272), every model's pose.
#define MF_MAX_MODELS 64
struct FrameResult {
    int hasNewLabel, newClassID, nMasks, nComponents; long long timestamp;
    int dead[MF_MAX_MODELS]; unsigned deadCount[MF_MAX_MODELS];
    float poses[MF_MAX_MODELS][32];                       // pose (row-major 4x4) | last incremental transform
};
struct SegTables { unsigned char idToIndex[256], indexToId[256], isModel[256]; };
struct VoteParams {
    int nModels, allowNew, personClassID; unsigned minNew, maxNew; float minMaskModelOverlap; unsigned char nextModelID;
    int modelClass[MF_MAX_MODELS]; unsigned char modelID[MF_MAX_MODELS];
};
struct LifeModel { int tracked, owned, ownerRank, pad; DevPose* dpose; const float* trackOut; unsigned* count; float initialC2Winv[16]; };
struct LifeParams { int nModels, rank; LifeModel m[MF_MAX_MODELS]; };

// Gauss-Newton state of one tracked model; lives in device memory for the whole frame
struct TrackState {
    float Rprev[9], tprev[3], RprevInv[9];
    float Rcurr[9], tcurr[3];
    float trR[9], trT[3];
    double resultRt[16];
    double resultR[9], lastResultR[9];
    float R_lr[9];
    float so3LastError, so3LastCount; int so3Done;
    float krk[9], kt[3];
    float sigmaVal; int levelBreak;
    float lastICPError, lastICPCount, lastRGBError, lastRGBCount, lastSO3Error, lastSO3Count;
    double lastA[36], lastb[6];
    float out[40];
    unsigned ticket[4];
};

struct TrackJob {
    const float4* vmapC[3]; const float4* nmapC[3];        // frame maps (shared by all models)
    const uint8_t* nextImage[3]; const short2* nextGrad[3]; const uint8_t* rgbValid[3];
    const float4* vmapG[3]; const float4* nmapG[3];        // model maps in the model-global frame
    const float* lastDepth[3]; const uint8_t* lastImage[3];
    const uint8_t* lastNextImage2;
    const float4* cloud[3];
    DataTerm* corres[3];
    DevPose* dpose;       // initial pose in, tracked pose + derived quantities out
    TrackState* st;
    float* partial;       // 2 x (TRACK_MAX_BLOCKS / 2) rows of 32 x 16 bytes: per-CTA partial sums (fp64 value + flags), ping-pong between reductions
    unsigned* bar;        // grid barrier counter of this job (own 128-byte line)
    const uint32_t* validBits[3];   // object models: one bit per model-map pixel, set where the model normal is valid (nullptr: not used)
};

void set_num_sms(int n);
// in-stream stage timer hook (mf_host.cu): records a CUDA event on `s`; the time until the next mark is attributed to `name`
void prof_mark(cudaStream_t s, const char* name);

// ---- mf_frame.cu ----
void launch_unpack_rgb(const uint8_t* rgb3, uchar4* out, int P, cudaStream_t s);
void launch_bilateral(const float* depth, float* out, int W, int H, cudaStream_t s);
// fused variants: two pyramid levels per launch; three levels of a per-pixel kernel per launch (blockIdx.z = level)
void launch_pyrdown2_f(const float* src, int sw, int sh, float* dst1, float* dst2, cudaStream_t s);
void launch_pyrdown2_u8(const uint8_t* src, int sw, int sh, uint8_t* dst1, uint8_t* dst2, cudaStream_t s);
void launch_pyrdown2_pair(const float* srcF, float* dstF1, float* dstF2, const uint8_t* srcU, uint8_t* dstU1, uint8_t* dstU2, int sw, int sh, cudaStream_t s);
void launch_vmap_nmap3(const float* const* depth, int W, int H, Cam cam, float cutoff, float4* const* vmap, float4* const* nmap, cudaStream_t s);
void launch_sobel3(const uint8_t* const* img, int W, int H, short2* const* grad, uint8_t* const* rgbValid, cudaStream_t s);
void launch_project_points3(const float* const* depth, int W, int H, Cam cam, float4* const* cloud, cudaStream_t s);
void launch_intensity(const uchar4* img, int P, uint8_t* out, cudaStream_t s);
void launch_intensity_select(const uchar4* imgPred, const uchar4* imgFill, const uint32_t* nonBlack, float denom, int forceFill, int P, uint8_t* out, cudaStream_t s);
float track_min_scale(int level);          // gradient-magnitude gate of the photometric term at a pyramid level (RGBDOdometry.cpp:44-49)
void launch_model_maps(const float4* srcVp, const float4* srcNp, const float4* srcVf, const float4* srcNf, const uint32_t* nonBlack, float denom,
                       int W, int H, const DevPose* dpose, float maxDepthRGB, float4* const* v, float4* const* n, float* depth0, cudaStream_t s);
void launch_valid_bits3(const float4* const* nmap, int W, int H, uint32_t* const* bits, cudaStream_t s);    // bit i of level l = !isnan(nmap[l][i].x)
void launch_map_to_planar(const float4* m, int P, float* out, cudaStream_t s);

// ---- mf_surfel.cu ----
void launch_fill_u32(uint32_t* p, uint32_t v, size_t n, cudaStream_t s);
void launch_fill_u64(uint64_t* p, uint64_t v, size_t n, cudaStream_t s);
void launch_set_pose(DevPose* d, const float* pose16, const float* lastPose16, cudaStream_t s);      // host-driven pose -> DevPose
void launch_predict_indices(const SurfelPlanes& sp, const uint32_t* count, const DevPose* dpose, Cam cam, int W, int H, float maxDepth, int time,
                            int timeDelta, uint64_t* key, uint32_t* idx, float4* vertConf, float4* colorTime, float4* normRad, float4* cleanTex, float cleanConf,
                            cudaStream_t s);
void launch_associate(const uchar4* rgb, const float* depthRaw, const float* depthFilt, const uint8_t* mask, const uint32_t* idx,
                      const float4* vertConf, const float4* normRad, const DevPose* dpose, Cam cam, int W, int H, float maxDepth, int time,
                      float weightMultiplier, uint8_t maskID, uint8_t* flag, uint32_t* best, float4* const* meas, uint32_t* slot, cudaStream_t s);
void launch_fuse_update(const uint8_t* flag, const uint32_t* best, float4* const* meas, uint32_t* slot, int P, int time,
                        const SurfelPlanes& sp, cudaStream_t s);
// index-map outputs of a Model::predictIndices that is carried out INSIDE the clean pass (one stream over the store instead of two)
// what the index-map window of Model::clean reads: the packed 16-byte texels (written by the index resolve for THIS call's time and
// confidence threshold) or, packed == nullptr, the index-map images themselves
struct CleanWindowImages { const float4* packed; const float4* vertConf; const float4* colorTime; const uint32_t* idx; };
struct IndexFused { uint64_t* key; uint32_t* idx; float4* vertConf; float4* colorTime; float4* normRad; float4* cleanTex; float maxDepth; };
// in-place ordered compaction of Model::clean (k_clean_compact): ticket (reset by the sums pass), one published-epoch word per 512-entry
// sub-block, the first sub-block that holds a removal (written by the scan), the epoch of this call (never 0, changes with every call)
struct CleanInPlace { uint32_t* ticket; uint32_t* loaded; uint32_t* firstMoved; uint32_t epoch; bool pingPong; };   // firstMoved[0..1]: first moved sub-block, number of sub-blocks; pingPong: copy into `dst` instead (same result)
void launch_clean(const SurfelPlanes& src, const SurfelPlanes& dst, const uint32_t* count, uint32_t* newCount, uint32_t capacity,
                  const uint8_t* aflag, float4* const* meas, const DevPose* dpose, Cam cam, int W, int H, int time, int timeDelta, float confThreshold,
                  float outlierCoeff, uint8_t maskID, const CleanWindowImages& win,
                  const float* depthFilt, const uint8_t* mask, uint8_t* keep, uint32_t* blockSums, uint32_t* cand, uint32_t* candCount, cudaStream_t s,
                  const IndexFused* fused = nullptr, const CleanInPlace* inplace = nullptr);
void launch_combined_predict(const SurfelPlanes& sp, const uint32_t* count, const DevPose* dpose, Cam cam, int W, int H, float maxDepth,
                             float confThreshold, int time, int maxTime, int timeDelta, const float4* rayTab, uint64_t* key, uchar4* image, float4* vertexConf,
                             float4* normalRad, uint16_t* timeTex, int doFill, const float* depthFilt, const uchar4* rgb, int ptVN, int ptImg,
                             uchar4* fillImage, float4* fillVertex, float4* fillNormal, uint32_t* nonBlackSamples, cudaStream_t s, uint32_t capacity = 0);
void launch_ray_table(Cam cam, int W, int H, float4* tab, cudaStream_t s);      // pixel-centre viewing rays (combo_splat.frag:39-45), once per context
void launch_init_model(const uchar4* rgb, const float* depthRaw, const float* depthFilt, Cam cam, int W, int H, int time, float maxDepth,
                       uint8_t* fr, uint8_t* ff, uint32_t* sumR, uint32_t* sumF, uint32_t capacity, const SurfelPlanes& sp, uint32_t* count,
                       cudaStream_t s);
void launch_planes_to_aos(const SurfelPlanes& sp, uint32_t n, float4* out, cudaStream_t s);
void launch_aos_to_planes(const float4* in, uint32_t n, const SurfelPlanes& sp, cudaStream_t s);

// ---- mf_track.cu ----
void track_shares(int nJobs, unsigned lightMask, int totalCTAs, int ratio, int* G);   // CTAs per model of the persistent tracking grid
int launch_tracking(TrackJob* d_jobs, int nJobs, int W, int H, Cam cam, bool rgbOnly, float icpWeight,
                    bool pyramid, bool fastOdom, bool so3, int numSMs, unsigned* bars, cudaStream_t s, unsigned lightMask = false);
int debug_track_timing(long long* out, int cap);
void launch_icp_only(const float4* vmapC, const float4* nmapC, const float4* vmapG, const float4* nmapG, int W, int H, Cam cam,
                     const TrackPoses& pp, float* partial, unsigned* ticket, float* out29, int numSMs, cudaStream_t s);

// ---- mf_seg.cu ----
void launch_geometric_edges(const float4* vmap, const float4* nmap, int W, int H, float wD, float wC, float thr, float* edge, uint8_t* binary, cudaStream_t s);
int launch_morph_close_ellipse(uint8_t* data, uint8_t* buf, int W, int H, int radius, int iterations, const FrameHdr* onlyIfMasks, cudaStream_t s);   // MfSegmentation.cpp:424-426; returns the number of launches
void launch_morph_close_invert(uint8_t* data, uint8_t* buf, int W, int H, int radius, int iterations, uint8_t* inverted, cudaStream_t s);

}  // namespace mfb

namespace mfb {
// ---- mf_seg.cu: GPU segmentation tail + global projection resolve ----
void launch_cc(const uint8_t* img, int W, int H, int* L, int* dense, int* lab, int* area, int* box, uint32_t* counter, cudaStream_t s);   // box: left, top, right, bottom per component
void launch_remove_edges(int* labA, int* labB, const float* depth, const int* area, int W, int H, int iterations, cudaStream_t s);
void launch_seg_tables(const SegTables& t, uint8_t* idToIndex, uint8_t* indexToId, uint8_t* isModel, cudaStream_t s);
void launch_frame_header(const FrameHdr& h, FrameHdr* d, cudaStream_t s);                       // single-process path: the header by value
void launch_person_table(const FrameHdr* hdr, int personClassID, uint8_t* isPerson, cudaStream_t s);
void launch_clear_hist(const uint32_t* ccCounter, const FrameHdr* hdr, int nModels, int* compModel, int* compMask, cudaStream_t s);
void launch_seg_hist(const int* lab, const uint8_t* projID, const uint8_t* mask, int P, const uint8_t* idToIndex, int nModels, const FrameHdr* hdr,
                     int* compModel, int* compMask, cudaStream_t s);
void launch_component_map(const uint32_t* ccCounter, const int* area, const int* compModel, const int* compMask, int nModels, const FrameHdr* hdr,
                          const uint8_t* indexToId, int minMapped, int* mapToMask, int* absorb, int* maskPixels, cudaStream_t s);
void launch_vote(const FrameHdr* hdr, const VoteParams& vp, const int* maskPixels, const unsigned* maskOverlap, const uint32_t* ccCounter,
                 uint8_t* maskToID, FrameResult* res, cudaStream_t s);
void launch_seg_assign(const int* lab, const int* mapToMask, const uint8_t* ignore, int P, uint8_t* seg, cudaStream_t s);
void launch_mask_overlap(const uint8_t* seg, const uint8_t* projID, const uint8_t* idToIndex, const uint8_t* isModelId, int P, unsigned* maskOverlap, cudaStream_t s);
void launch_seg_final(const uint8_t* seg, const int* lab, const int* mapToMask, const int* absorb, const uint8_t* maskToID, const int* box, int P, int W,
                      uint8_t* out, cudaStream_t s);
void launch_apply_ignore(const uint8_t* mask, const uint8_t* isPerson, const FrameHdr* hdr, int P, uint8_t* ignore, uint8_t* edges, cudaStream_t s);
void launch_proj_resolve(uint64_t* key, int P, const uint8_t* indexToId, uint8_t* out, cudaStream_t s);
void launch_splat_project_only(const SurfelPlanes& sp, const uint32_t* count, const DevPose* dpose, Cam cam, int W, int H, float maxDepth, float confThreshold,
                               int time, int maxTime, int timeDelta, uint32_t drawBase, const float4* rayTab, uint64_t* key, cudaStream_t s, uint32_t capacity = 0);
}  // namespace mfb
