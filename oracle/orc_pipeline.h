/* oracle/orc_pipeline.h -- CPU ORACLE (test infrastructure, not product):
 * stateful restatement of MaskFusion::processFrame (Core/MaskFusion.cpp:200-607)
 * and Model (Core/Model/Model.cpp) on top of the stateless passes in orc.h. */
#ifndef ORC_PIPELINE_H
#define ORC_PIPELINE_H
#include "orc.h"
#include "orc_odom.h"
#include "orc_mfseg.h"
#ifdef __cplusplus
extern "C" {
#endif

/* GUI-effective defaults ("Config / flags") */
typedef struct {
    int32_t width, height;
    float fx, fy, cx, cy;
    float depthCutoff;          /* GUI.h:194 (4.0)                 */
    float maxDepthProcessed;    /* MaskFusion.cpp:57 (20.0)        */
    float icpWeight;            /* GUI.h:195 (20.0)                */
    int32_t rgbOnly, pyramid, fastOdom, so3, frameToFrameRGB;
    float confGlobal, confObject;   /* MainController.cpp:215-216  */
    int32_t timeDelta;          /* INT_MAX/2, MainController.cpp:399 */
    float outlierCoeff;         /* GUI.h:196 (0.1)                 */
    int32_t capacityGlobal, capacityObject;   /* surfels, Core/CMakeLists.txt:27-28 */
    int32_t enableMultipleModels;   /* 0 == -static                */
    int32_t trackAllModels;
    int32_t modelSpawnOffset;   /* GUI.h:347 (22)                  */
    float minRelSizeNew, maxRelSizeNew;
    float segThreshold, segWeightDistance, segWeightConvexity;   /* GUI.h:367-374 */
    int32_t segMorphEdgeIterations, segMorphEdgeRadius, segMorphMaskIterations, segMorphMaskRadius;
} orc_config;

void orc_config_defaults(orc_config* c, int width, int height);

typedef struct orc_model {
    int id; int classID;
    float pose[16], lastPose[16], initialC2Winv[16];
    int isStatic, age;
    int nonstatic;
    float confThreshold, maxDepth;
    int capacity, count;
    float* surf[2]; int target;           /* ping-pong VBOs (Model.h:285) */
    int allowFillIn;
    /* index map (ModelProjection sparse* textures) */
    uint32_t* idx; float* vertConf; float* colorTime; float* normRad;
    /* combinedPredict outputs */
    uint8_t* splatImage; float* splatVertex; float* splatNormal; uint16_t* splatTime;
    /* fill-in outputs */
    float* fillVertex; float* fillNormal; uint8_t* fillImage;
    /* data-association scratch (newUnstableBuffer + update maps) */
    uint8_t* updateId; uint32_t* best; float* meas;
    orc_odom* odom;
    /* pose log: ts, x y z qx qy qz qw */
    int nlog, caplog; double* log;
} orc_model;

typedef struct orc_mf {
    orc_config cfg; orc_cam cam;
    int tick;
    uint8_t* rgb; float* depthRaw; float* depthFilt; uint8_t* mask;
    float* depthPyr[3]; uint8_t* maskPyr[3]; float* vmap[3]; float* nmap[3];
    int nmodels; orc_model* models[256];
    uint8_t nextID;
    int spawnOffset;
    /* multi-model state */
    uint64_t* projKeys; uint8_t* projectedIDs; uint8_t* fullSeg;
    float* edgeMap; uint8_t* edgeBin; uint8_t* edgeBuf;
    orc_mfseg_state seg;
    int lastHasNewLabel;
} orc_mf;

orc_mf* orc_mf_create(const orc_config* cfg);
void orc_mf_destroy(orc_mf* h);
/* rgb: HxWx3 u8, depth: HxW f32 metres.  inPose unused (NULL). */
void orc_mf_set_frame(orc_mf* h, const uint8_t* rgb3, const float* depth, const uint8_t* mask);
int orc_mf_process_frame_ex(orc_mf* h, const uint8_t* rgb3, const float* depth, int64_t timestamp, const uint8_t* mask,
                            const int32_t* classIDs, int nMasks);
void orc_mf_global_projection(orc_mf* h);
void orc_mf_segmentation(orc_mf* h, const uint8_t* mask, const int32_t* classIDs, int nMasks, int allowNew, orc_mfseg_out* out);
int orc_mf_process_frame(orc_mf* h, const uint8_t* rgb3, const float* depth, int64_t timestamp);
orc_model* orc_mf_model(orc_mf* h, int i);
const float* orc_model_surfels(const orc_model* m);     /* == getModelBuffer(): vbos[target] */

/* Model-level entry points (Model.h:128-164), usable stand-alone by tests */
void orc_model_predict_indices(orc_mf* h, orc_model* m, int time);
void orc_model_fuse(orc_mf* h, orc_model* m, int time, float depthCutoff, float weightMultiplier);
void orc_model_clean(orc_mf* h, orc_model* m, int time);
void orc_model_combined_predict(orc_mf* h, orc_model* m, int time, int maxTime);
void orc_model_fill_in(orc_mf* h, orc_model* m);
void orc_model_track(orc_mf* h, orc_model* m, float* transformOut);
float orc_model_fusion_weight(const orc_model* m, float weightMultiplier);
void orc_generate_frame_maps(orc_mf* h);                /* Model::generateCUDATextures */

#ifdef __cplusplus
}
#endif
#endif
