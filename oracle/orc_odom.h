/* oracle/orc_odom.h -- CPU ORACLE (test infrastructure): RGBDOdometry state
 * (Core/Utils/RGBDOdometry.h:95-150). */
#ifndef ORC_ODOM_H
#define ORC_ODOM_H
#include "orc.h"
#ifdef __cplusplus
extern "C" {
#endif
struct orc_odom {
    int W, H; orc_cam cam;
    float* vtex_tmp;                 /* vmaps_tmp: interleaved predicted vertices (RGBDOdometry.cpp:158-161) */
    float* vmap_g[3]; float* nmap_g[3];
    float* lastDepth[3]; float* nextDepth[3];
    uint8_t* lastImage[3]; uint8_t* nextImage[3]; uint8_t* lastNextImage[3];
    int16_t* dIdx[3]; int16_t* dIdy[3];
    float* cloud[3];
    orc_dataterm* corres[3];
    float lastICPError, lastICPCount, lastRGBError, lastRGBCount, lastSO3Error, lastSO3Count;
    double lastA[36], lastb[6];
};
orc_odom* orc_odom_create(int W, int H, orc_cam cam);
void orc_odom_destroy(orc_odom* o);
void orc_odom_init_first_rgb(orc_odom* o, const uint8_t* rgb3);
void orc_odom_init_icp_model(orc_odom* o, const float* vtex4, const float* ntex4, const float* pose);
void orc_odom_init_rgb_model(orc_odom* o, const uint8_t* image4);
void orc_odom_init_rgb(orc_odom* o, const uint8_t* rgb3);
void orc_odom_track(orc_odom* o, float* const* frame_vmaps, float* const* frame_nmaps,
                    const orc_track_params* prm, float* pose, float* transformOut);
void orc_ldlt_solve(const double* A, const double* b, int n, double* x);
void orc_rodrigues(const double* src, double* R9);
#ifdef __cplusplus
}
#endif
#endif
