/* oracle/orc_mfseg.h -- CPU ORACLE (test infrastructure): MfSegmentation CPU tail
 * (Core/Segmentation/MfSegmentation.cpp:208-538, MfSegmentation.h:40-58). */
#ifndef ORC_MFSEG_H
#define ORC_MFSEG_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

typedef struct {
    uint8_t* semanticIgnoreMap;          /* persists across frames (:221-235) */
    uint8_t maskToID[256];
    float minMaskModelOverlap;           /* 0.05, MfSegmentation.cpp:43 */
    int32_t minMappedComponentSize;      /* 160 */
    int32_t personClassID;               /* 255 */
    int32_t removeEdges;                 /* true */
} orc_mfseg_state;

typedef struct {
    int32_t W, H;
    uint8_t* edgesInv;                   /* thresholded + closed + inverted edge map (255 = not an edge); overwritten */
    const float* depth;                  /* frame->depth (raw metric) */
    const uint8_t* mask; int32_t nMasks; const int32_t* classIDs;   /* frame->mask / frame->classIDs (nMasks == classIDs.size(), 0 => no masks) */
    const uint8_t* projectedIDs;         /* GlobalProjection::getProjectedModelIDs */
    int32_t nModels; const uint8_t* modelIDs; const int32_t* modelClassIDs;
    uint8_t nextModelID; int32_t allowNew;
    float minRelSizeNew, maxRelSizeNew;
    int32_t morphMaskRadius, morphMaskIterations;
} orc_mfseg_in;

typedef struct {
    uint8_t* fullSegmentation;           /* HxW model id per pixel (255 = ignored) */
    int32_t hasNewLabel, newClassID;
    int32_t isEmpty[256], pixelCount[256];
    int32_t* labels;                     /* optional: component labels after edge removal */
    int32_t nComponents;
} orc_mfseg_out;

void orc_mfseg_set_threads(int n);        /* 1 (default) = the reference's single-threaded sweeps; n > 1 = OpenMP split, same results */
void orc_mfseg_state_init(orc_mfseg_state* s, int W, int H);
void orc_mfseg_state_free(orc_mfseg_state* s);
void orc_mfseg_cpu(orc_mfseg_state* st, const orc_mfseg_in* in, orc_mfseg_out* out);
int orc_connected_components4(const uint8_t* img, int W, int H, int32_t* labels, int32_t** stats);
void orc_morph_close_ellipse(uint8_t* img, int W, int H, int r, int iterations);

#ifdef __cplusplus
}
#endif
#endif
