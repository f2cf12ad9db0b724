// mf_host.h -- host classes of the B200-native dense pipeline.  Class and method names
// mirror the reference so that its call sites read the same:
//   mfb::MaskFusion  <-> MaskFusion   (Core/MaskFusion.h:47-70)
//   mfb::Model       <-> Model        (Core/Model/Model.h:128-164)
// GPUTexture* arguments of the reference become device buffers owned by these classes;
// Eigen::Matrix4f becomes Mat4 (row-major here, column-major at the C ABI).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <memory>
#include <string>
#include <vector>
#include <map>
#include "../../include/maskfusion_b200.h"
#include "mf_kernels.h"

namespace mfb {

struct Mat4 {
    float m[16];
    static Mat4 identity() { Mat4 r; for (int i = 0; i < 16; ++i) r.m[i] = (i % 5 == 0) ? 1.f : 0.f; return r; }
};
Mat4 rigidInverse(const Mat4& T);
Mat4 mul(const Mat4& A, const Mat4& B);
Rt toRt(const Mat4& T);

struct CudaError { std::string what; };
void cudaCheck(cudaError_t e, const char* where);

template <typename T>
struct DevBuf {
    T* p = nullptr; size_t n = 0;
    DevBuf() {}
    DevBuf(const DevBuf&) = delete;
    DevBuf& operator=(const DevBuf&) = delete;
    ~DevBuf() { if (p) cudaFree(p); }
    void alloc(size_t count) { if (p) cudaFree(p); p = nullptr; n = count; if (count) cudaCheck(cudaMalloc((void**)&p, count * sizeof(T)), "cudaMalloc"); }
    void zero(cudaStream_t s) { if (n) cudaCheck(cudaMemsetAsync(p, 0, n * sizeof(T), s), "memset"); }
    operator T*() const { return p; }
};

struct Profiler {
    bool on = false; int used = 0;
    std::vector<cudaEvent_t> events; std::vector<const char*> names;
    std::map<std::string, std::pair<long, double>> acc;      // name -> (count, total ms)
    void resolve();
};
extern Profiler* g_prof;

class MaskFusion;

// object-sharded mode: NCCL communicator of the ranks that share one replay (mf_sched.cu; libnccl is opened at run time)
struct ShardComm {
    void* comm = nullptr; int rank = 0, world = 1, version = 0; size_t bytesMoved = 0; long calls = 0;
    ~ShardComm();
    void init(const unsigned char* id128, int rank, int world);
    void broadcast(void* buf, size_t bytes, int root, cudaStream_t s);                               // frame packet (MaskFusion.cpp:212-217)
    void allGatherFloats(const float* send, float* recv, size_t countPerRank, cudaStream_t s);       // pose rows (MaskFusion.cpp:257-276)
    void allReduceMinU64(uint64_t* buf, size_t count, cudaStream_t s);                               // ID-projection keys (GlobalProjection.cpp:66-95)
};
void shardUniqueId(unsigned char* out128);
void launch_pack_rows(const LifeParams& lp, float* table, cudaStream_t s);
void launch_lifecycle(const LifeParams& lp, const float* gathered, FrameResult* res, cudaStream_t s);
void launch_set_count(uint32_t* c, uint32_t v, cudaStream_t s);

class Model {
public:
    // ghost = replica of a model whose surfel store lives on another rank: pose, ids and age only, no device buffers
    Model(MaskFusion* owner, unsigned char id, float confidenceThresh, bool enableFillIn, int capacity, int ownerRank = 0, bool ghost = false);
    ~Model();
    Model(const Model&) = delete;

    // ---- reference API (Core/Model/Model.h:128-164) ----
    void initialise(int time);                                        // Model::initialise (+ computeFeedbackBuffers)
    void prepareTracking();                                           // Model::initICP (model side)
    void predictIndices(int time, float depthCutoff, int timeDelta, bool forClean = true);  // Model::predictIndices (forClean: also the packed window texels)
    void fuse(int time, float depthCutoff, float weightMultiplier);   // Model::fuse
    void clean(int time, int timeDelta, float depthCutoff);           // Model::clean
    void combinedPredict(float depthCutoff, int time, int maxTime, int timeDelta);   // Model::combinedPredict + performFillIn
    float computeFusionWeight(float weightMultiplier) const;          // Model::computeFusionWeight
    void overridePose(const Mat4& p) { lastPose = pose; pose = p; pushPose(); }
    void makeStatic(const Mat4& globalPose) { initialC2Winv = mul(pose, rigidInverse(globalPose)); isStatic = true; }
    void updateStaticPose(const Mat4& globalPose) { overridePose(mul(initialC2Winv, globalPose)); }
    void pushPose();                                                  // host pose/lastPose -> device-resident DevPose (k_set_pose)
    bool allowsFillIn() const { return fillIn; }
    unsigned lastCount();                                             // Model::lastCount (synchronises)
    SurfelPlanes planes(int b) const { return SurfelPlanes{pos[b].p, col[b].p, nrm[b].p}; }
    SurfelPlanes current() const { return planes(target); }
    uint32_t* dCount() const { return count.p + countSel; }

    MaskFusion* owner;
    int ownerRank = 0; bool owned = true;                             // sharded mode: which rank holds the surfels
    unsigned char id; int classID = -1;
    Mat4 pose, lastPose, initialC2Winv;
    bool isStatic = false, nonstatic = false; unsigned age = 0;
    float confidenceThreshold, maxDepth;
    bool fillIn;
    uint32_t capacity;
    int target = 0, countSel = 0;
    DevBuf<float4> pos[2], col[2], nrm[2];
    DevBuf<uint32_t> count;                 // [2] ping-pong, device-resident (no host round trip in the loop)
    uint32_t* hCount = nullptr;             // pinned mirror
    // index map
    DevBuf<uint64_t> key;
    DevBuf<uint32_t> idx; DevBuf<float4> vertConf, colorTime, normRad, cleanTex;
    // prediction + fill-in
    DevBuf<uchar4> splatImage, fillImage; DevBuf<float4> splatVertex, splatNormal, fillVertex, fillNormal; DevBuf<uint16_t> splatTime;
    DevBuf<uint32_t> nonBlack;
    // association
    DevBuf<uint8_t> aflag; DevBuf<uint32_t> abest; DevBuf<float4> meas[3]; DevBuf<uint32_t> slot;
    DevBuf<uint8_t> keep; DevBuf<uint32_t> blockSums, blockSums2, cand, candCount;
    uint32_t* hCleanStat = nullptr;          // pinned: {first moved sub-block, sub-blocks} of the previous clean (adaptive in-place / copy choice)
    DevBuf<uint32_t> cleanTicket, cleanLoaded; uint32_t cleanEpoch = 0;     // in-place compaction of Model::clean: [0] ticket, [1] first moved sub-block; published epochs
    // tracking
    DevBuf<float4> vmapG[3], nmapG[3], cloud[3];
    DevBuf<float> lastDepth[3]; DevBuf<uint8_t> lastImage[3]; DevBuf<uint8_t> lastNextImage2;
    DevBuf<DataTerm> corres[3];
    DevBuf<uint32_t> validBits[3];           // object models: validity bitmask of nmapG (see k_valid_bits3)
    DevBuf<TrackState> trackState; DevBuf<float> partial;
    DevBuf<DevPose> dpose;                  // what every kernel reads: pose, inverse, fusion weight (device resident)
    float* hTrackOut = nullptr;             // pinned: pose(16) transform(16) stats(8)
    Mat4 lastTransform;
    std::vector<double> poseLog;            // 8 doubles per entry
    bool tracked = false;                   // took part in the tracking launch of the frame in flight (deferred bookkeeping)
    // predictIndices(forClean) is lazy: when Model::clean follows with the same time gate (the frame schedule, MaskFusion.cpp:550-562) the
    // projection rides inside the clean pass; any reader of the index map in between (fuse, the read-backs) runs it stand-alone first
    bool idxDeferred = false; int idxTime = 0, idxDelta = 0; float idxDepth = 0.f;
    int cleanTexTime = -1; float cleanTexConf = -1.f;    // the time / confidence threshold the packed window texels (cleanTex) were written with
    void flushIndex();
};

class MaskFusion {
public:
    MaskFusion(const mf_config& cfg, int device, cudaStream_t stream);
    ~MaskFusion();

    // bool MaskFusion::processFrame(FrameDataPointer, const Eigen::Matrix4f* inPose, float weightMultiplier, bool bootstrap)
    bool processFrame(const uint8_t* rgb, const float* depth, int64_t timestamp, const uint8_t* mask, const Mat4* inPose,
                      float weightMultiplier, bool bootstrap, bool inputsOnDevice);
    // upload + filterDepth; generateCUDATextures; frame side of initRGB.  `s` = stream to enqueue on (nullptr: the main stream)
    void setFrame(const uint8_t* rgb, const float* depth, const uint8_t* mask, bool onDevice, cudaStream_t s = nullptr);
    void uploadInputs(const uint8_t* rgb, const float* depth, const uint8_t* mask, int64_t timestamp, bool onDevice, cudaStream_t s = nullptr);
    void preprocess(cudaStream_t s = nullptr);
    void generateCUDATextures(cudaStream_t s = nullptr);                                          // Model::generateCUDATextures
    void frameIntensity(cudaStream_t s = nullptr);                                                // RGBDOdometry::initRGB (frame side) + Sobel + validity
    void trackModels(const std::vector<Model*>& ms, bool viaResult = false);                      // performTracking for a batch
    void predict();                                                                               // MaskFusion::predict
    void sync();
    // The tracked pose reaches the host through an asynchronous copy + event recorded right after the tracking kernel.  The -static
    // schedule has no host decision that depends on it, so processFrame returns with the rest of the frame enqueued and the pose is
    // picked up (finalisePending) by the next processFrame / getPose / sync; the multi-model schedule finalises inside the frame.
    void finalisePending();
    void logPoses(int64_t timestamp);
    bool pendingTrack = false, pendingLog = false; int64_t pendingTimestamp = 0; std::vector<Model*> pendingModels; cudaEvent_t trackDone = nullptr;
    // ---- multi-model path (MaskFusion.cpp:287-375) ----
    void globalProjection();                                                                      // GlobalProjection::project + downloadDirect (stays on the device)
    void segTables();
    void performSegmentation(bool allowNew);                                                      // MfSegmentation::performSegmentation (result -> FrameResult)
    unsigned char getNextModelID(bool assign);                                                    // MaskFusion::getNextModelID
    Model* spawnObjectModel();                                                                    // MaskFusion::spawnObjectModel + moveNewModelToList
    void setFrameClasses(const int32_t* ids, int n) { classIDs.assign(ids, ids + n); }            // FrameData::classIDs

    // ---- a frame in three phases; between them sit the two exchanges of the object-sharded mode:
    //   frameBegin  inputs (+ frame-packet broadcast), preprocessing, tracking of the models whose store lives here, pose rows
    //   [pose-row all-gather]
    //   frameProject  device-side lifecycle (inactivation, static poses), local part of the global ID projection
    //   [64-bit MIN all-reduce of the projection keys]
    //   frameEnd  resolve, segmentation + vote (device driven), FrameResult copy, fusion of the local stores, prediction
    // With an NCCL communicator (initShardComm) the exchanges are issued from here on the context's stream and NOTHING in a frame
    // waits for the host; without one the caller moves the rows / keys itself between the phase calls (any transport: the gloo tests).
    void configureShard(int rank, int world);
    void initShardComm(const unsigned char* id128, int rank, int world);
    void frameBegin(const uint8_t* rgb, const float* depth, int64_t timestamp, const uint8_t* mask, const Mat4* inPose, bool bootstrap, bool onDevice);
    void getShardPoses(float* out);                             // [MF_MAX_MODELS][32] rows of this rank (external transport; synchronises)
    void setShardPoses(const float* gathered);                  // [world][MF_MAX_MODELS][32] -> device (external transport)
    void frameProject();
    void frameEnd(float weightMultiplier);
    // deferred bookkeeping of the multi-model schedule: applied at the start of the next frame (or by any query in between)
    void applyFrameResult();
    LifeParams lifeParams() const;
    ShardComm shard; bool shardNccl = false;
    // Mask R-CNN backbone on the frame path (MaskRCNN::executeSequential, MaskRCNN.cpp:147-151, is called from MfSegmentation.cpp:130):
    // every k-th frame the RGB image is letter-boxed into the backbone's input and the ResNet-101-FPN forward is enqueued on the
    // backbone's own stream, next to the dense pipeline of the same GPU (the reference runs its network as a ~5 Hz sidecar)
    void* backbone = nullptr; int backboneEvery = 0; cudaEvent_t bbFrameReady = nullptr, bbMoldDone = nullptr; bool bbMoldPending = false;
    void attachBackbone(void* bb, int everyK);
    void runBackbone(cudaStream_t producer = nullptr);
    // multi-model frames with the inputs / preprocessing (and, sharded, every collective) on preStream: MFB200_MULTI_OVERLAP=1
    bool multiOverlap = false, spawnedInApply = false, commOnPre = false; cudaEvent_t evMain = nullptr, evComm = nullptr;
    FrameResult* hRes = nullptr; DevBuf<FrameResult> dRes; cudaEvent_t resEvt = nullptr; bool pendingResult = false;
    DevBuf<float> poseTable, gathered;
    float fWeight = 1.f; int fTick = 0; bool fTracked = false;
    std::vector<std::unique_ptr<Model>> inactiveModels;        // MaskFusion::inactiveModels (MaskFusion.cpp:699-713): kept for exportPoses / savePly
    bool enableSmartModelDelete = true; unsigned modelKeepMinSurfels = 4000; float modelKeepConfThreshold = 0.3f;   // MaskFusion.h:398,414-415
    void projectLocal(); void projectResolve();                 // the two halves of globalProjection()
    static int pickOwner(const int64_t* loads, int world);      // least-loaded rank (by owned surfel capacity), ties -> highest rank
    int rank = 0, world = 1;
    int64_t fTimestamp = 0; Mat4 fInPose; bool fHasPose = false, fBootstrap = false;

    mf_config cfg; Cam cam; int W, H, P; int device; cudaStream_t stream; bool ownStream;
    int numSMs = 148;
    bool fuseIndexIntoClean = true;         // Model::predictIndices rides inside the following Model::clean (one stream over the store); MFB200_FUSE_INDEX=0: two passes (A/B)
    bool cleanInPlace = true;               // Model::clean compacts the store in place, touching only the tail behind the first removal; MFB200_CLEAN_INPLACE=0: ping-pong copy of the whole store (A/B)
    bool trackValidBits = true;             // MFB200_TRACK_BITS=0 switches it off: object models carry a validity bitmask of their model maps for the tracker's early reject
    int tick = 1;
    int64_t launches = 0;
    std::vector<std::unique_ptr<Model>> models;
    unsigned char nextID = 0;
    // frame
    // Frame inputs exist twice: in the -static schedule the upload and preprocessing of frame t+1 run on their own stream
    // (preStream) while the surfel passes of frame t, which still read frame t's images, occupy the main stream.
    // The loader's data of one frame is ONE contiguous device buffer -- rgb (3P) | raw depth (4P) | instance mask (P) | FrameHdr -- which
    // is exactly the frame packet the object-sharded mode broadcasts (MaskFusion.cpp:212-217).
    DevBuf<uint8_t> inBuf[2]; DevBuf<uchar4> rgbBuf[2]; DevBuf<float> depthFiltBuf[2];
    uint8_t* rgb3 = nullptr; uchar4* rgb = nullptr; float* depthRaw = nullptr; float* depthFilt = nullptr;   // the current set
    uint8_t* frameMask = nullptr; FrameHdr* dHdr = nullptr;                                                   // FrameData::mask / classIDs of the current set
    int curSet = 0;
    size_t packetBytes() const { return (size_t)P * 8 + sizeof(FrameHdr); }
    void selectSet(int k)
    {
        curSet = k; rgb3 = inBuf[k].p; depthRaw = reinterpret_cast<float*>(inBuf[k].p + (size_t)P * 3); frameMask = inBuf[k].p + (size_t)P * 7;
        dHdr = reinterpret_cast<FrameHdr*>(inBuf[k].p + (size_t)P * 8); rgb = rgbBuf[k]; depthFilt = depthFiltBuf[k];
    }
    cudaStream_t preStream = nullptr; cudaEvent_t preDone = nullptr, inputsCopied = nullptr; bool preWaitPending = false, copyPending = false;
    cudaEvent_t inputReady = nullptr;        // caller's producer event for device inputs (mf_set_input_event): waited on before the next frame's copies
    DevBuf<uint8_t> mask;
    DevBuf<float> depthPyr[3]; DevBuf<float4> vmap[3], nmap[3];
    DevBuf<uint8_t> nextImage[3]; DevBuf<short2> nextGrad[3]; DevBuf<uint8_t> rgbValid[3];
    DevBuf<float> edgeMap; DevBuf<uint8_t> edgeBinary, edgeBuf, edgeInv;
    DevBuf<TrackJob> dJobs; TrackJob* hJobs = nullptr; DevBuf<unsigned> trackBars;
    DevBuf<uint8_t> initFlagR, initFlagF;
    DevBuf<float> scratch;                  // read-back staging
    DevBuf<float4> rayTab;                  // viewing ray of every pixel centre (camera constant): read by the splat rasteriser
    bool frameMapsValid = false, intensityValid = false;
    Profiler prof;
    // multi-model state
    std::vector<int32_t> classIDs;          // of the frame being processed
    int spawnOffset = 0;
    DevBuf<uint64_t> projKeys; DevBuf<uint8_t> projectedIDs;
    DevBuf<int> ccL, ccDense, ccLabA, ccLabB, ccArea, ccBox, mapToMask, absorbId, maskPixels, compModel, compMask;
    DevBuf<uint32_t> ccCounter; DevBuf<unsigned> maskOverlap;
    DevBuf<uint8_t> segTmp, ignoreMap, tblIdToIndex, tblIndexToId, tblIsModel, tblMaskToID, tblIsPerson;
    float minMaskModelOverlap = 0.05f; int minMappedComponentSize = 160; int personClassID = 255;   // MfSegmentation.cpp:43, MfSegmentation.h:58
};

}  // namespace mfb
