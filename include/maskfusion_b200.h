/*
 * include/maskfusion_b200.h -- C ABI of the B200-native MaskFusion dense pipeline.
 *
 * The reference has no FFI: its hot path sits behind C++ methods of libmaskfusion.so
 * that take Eigen / OpenCV / OpenGL types.  This header is
 * the flat, toolchain-neutral boundary a maintainer binds instead; each entry point
 * names the reference method it replaces.  Plain pointers and sizes only: no torch,
 * Eigen, OpenCV or GL types.  All functions return 0 on success, non-zero on error
 * (text via mf_last_error()); nothing calls exit() (the reference does on CUDA
 * errors, Core/Cuda/convenience.cuh:76-83).
 *
 * Conventions
 *   rgb    : H x W x 3 uint8, as FrameData::rgb (CV_8UC3)      Core/FrameData.h:37
 *   depth  : H x W float32 metres, as FrameData::depth (CV_32FC1) Core/FrameData.h:38
 *   mask   : H x W uint8 (optional external segmentation)      Core/FrameData.h:36
 *   poses  : float[16] COLUMN-major == Eigen::Matrix4f storage (Core/Model/Model.h:263-264)
 *   surfels: 12 floats each, position.xyz conf | colour unused initTime lastTime |
 *            normal.xyz radius                                 Core/Model/Model.h:190-192
 *   "tex4" : H x W x 4 float32, the layout of the reference's RGBA32F GL textures
 *   planar maps: 3*H x W float32 (x plane, y plane, z plane)   Core/Cuda/cudafuncs.cu:124-126
 */
#ifndef MASKFUSION_B200_H
#define MASKFUSION_B200_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

#define MF_ABI_VERSION 1

/* Effective parameters of one run: the reference spreads these over MaskFusion's
 * 23-argument constructor (Core/MaskFusion.h:47-53), ~60 setters and the GUI
 * defaults pushed every frame (GUI/MainController.cpp:528-571, GUI/Tools/GUI.h). */
typedef struct mf_config {
    int32_t width, height;                 /* Resolution::setResolution, MainController.cpp:117 */
    float fx, fy, cx, cy;                  /* Intrinsics::setIntrinics,  MainController.cpp:124-126 */
    float depthCutoff;                     /* GUI.h:194 = 4.0 */
    float maxDepthProcessed;               /* MaskFusion.cpp:57 = 20.0 */
    float icpWeight;                       /* GUI.h:195 = 20.0 (100 => ICP only, RGBDOdometry.cpp:236-237) */
    int32_t rgbOnly, pyramid, fastOdom, so3, frameToFrameRGB;
    float confGlobal, confObject;          /* MainController.cpp:215-216 = 10, 0.01 */
    int32_t timeDelta;                     /* MainController.cpp:399 = INT_MAX/2 (open loop) */
    float outlierCoeff;                    /* GUI.h:196 = 0.1 */
    int32_t capacityGlobal, capacityObject;/* Core/CMakeLists.txt:27-28 (runtime here, not compile time) */
    int32_t enableMultipleModels;          /* 0 == "-static" */
    int32_t trackAllModels;                /* GUI.h:344 */
    int32_t modelSpawnOffset;              /* GUI.h:347 = 22 */
    float minRelSizeNew, maxRelSizeNew;    /* GUI.h:345-346 */
    float segThreshold, segWeightDistance, segWeightConvexity;           /* GUI.h:367-374 */
    int32_t segMorphEdgeIterations, segMorphEdgeRadius, segMorphMaskIterations, segMorphMaskRadius;
} mf_config;

typedef struct mf_context mf_context;

const char* mf_last_error(void);
int mf_abi_version(void);
void mf_config_defaults(mf_config* cfg, int width, int height);

/* MaskFusion::MaskFusion (Core/MaskFusion.h:47-53).  device = CUDA ordinal
 * (reference: MASKFUSION_GPU_SLAM, MaskFusion.cpp:89-90); stream = cudaStream_t the
 * whole pipeline is enqueued on (NULL = a private non-blocking stream). */
mf_context* mf_create(const mf_config* cfg, int device, void* stream);
void mf_destroy(mf_context* ctx);

/* bool MaskFusion::processFrame(FrameDataPointer, const Eigen::Matrix4f* inPose,
 *      float weightMultiplier, bool bootstrap)                 Core/MaskFusion.h:69-70
 * Host buffers; the H2D copies are part of the call. mask / in_pose may be NULL. */
int mf_process_frame(mf_context* ctx, const uint8_t* rgb, const float* depth, int64_t timestamp,
                     const uint8_t* mask, const float* in_pose, float weight_multiplier, int bootstrap);
/* Same, inputs already resident in device memory (rgb: packed 3 bytes/pixel).
 * Lifetime / ordering of device inputs: the library copies them into its own frame set at the START of the call; on -static
 * tracking frames that copy runs on a private pre-processing stream next to the previous frame's surfel passes, NOT behind
 * the work queued on the context stream.  Therefore (a) the buffers must be completely written when the call is made -- or
 * the producer's completion event is handed over with mf_set_input_event before the call; (b) work the caller enqueues on
 * the CONTEXT stream after the call returns is ordered after the copies (the context stream waits for them), so the buffers
 * may be overwritten from there; a caller writing them from another stream waits for mf_sync or the next call's return. */
int mf_process_frame_device(mf_context* ctx, const void* d_rgb, const void* d_depth, int64_t timestamp,
                            const void* d_mask, const float* in_pose, float weight_multiplier, int bootstrap);
/* cudaEvent_t that the NEXT mf_process_frame_device waits on (on whichever stream it copies from) before it reads its
 * device inputs; consumed by that call.  NULL clears it. */
int mf_set_input_event(mf_context* ctx, void* cuda_event);
int mf_sync(mf_context* ctx);                       /* wait for everything enqueued so far */
int mf_tick(mf_context* ctx);                       /* MaskFusion::getTick */
int64_t mf_kernel_launches(mf_context* ctx);        /* kernels launched since creation */

/* ---- model list (MaskFusion::getModels) ---- */
int mf_model_count(mf_context* ctx);
int mf_model_id(mf_context* ctx, int i);
int mf_get_pose(mf_context* ctx, int i, float pose16[16]);          /* Model::getPose */
int mf_set_pose(mf_context* ctx, int i, const float pose16[16]);    /* Model::overridePose */
int mf_model_surfel_count(mf_context* ctx, int i);                  /* Model::lastCount */
int mf_model_set_conf_threshold(mf_context* ctx, int i, float t);   /* Model::setConfidenceThreshold */
int mf_download_surfels(mf_context* ctx, int i, float* out, int max_surfels);   /* Model::downloadMap, Model.cpp:944-974 */
int mf_upload_surfels(mf_context* ctx, int i, const float* in, int n);          /* test / bootstrap hook */
int mf_pose_log_size(mf_context* ctx, int i);                                   /* Model::getPoseLog */
int mf_get_pose_log(mf_context* ctx, int i, double* out8, int max_entries);     /* ts x y z qx qy qz qw, MaskFusion.cpp:850-879 */

/* ---- per-stage entry points (same names / order as Core/Model/Model.h:128-164) ---- */
int mf_set_frame(mf_context* ctx, const uint8_t* rgb, const float* depth, const uint8_t* mask);   /* upload + filterDepth + Model::generateCUDATextures */
int mf_model_perform_tracking(mf_context* ctx, int i, float transform16[16]);                      /* Model::performTracking */
int mf_model_predict_indices(mf_context* ctx, int i, int time);                                    /* Model::predictIndices */
int mf_model_fuse(mf_context* ctx, int i, int time, float depth_cutoff, float weight_multiplier);  /* Model::fuse */
int mf_model_clean(mf_context* ctx, int i, int time);                                              /* Model::clean */
int mf_model_combined_predict(mf_context* ctx, int i, int time, int max_time);                     /* Model::combinedPredict + performFillIn */
int mf_model_init_from_frame(mf_context* ctx, int i, int time);                                    /* computeFeedbackBuffers + Model::initialise */

/* ---- read-back of intermediate products, in the reference's layouts ---- */
int mf_download_filtered_depth(mf_context* ctx, float* out);                                       /* textureDepthMetricFiltered */
int mf_download_frame_maps(mf_context* ctx, int level, float* depth, float* vmap_planar, float* nmap_planar);  /* GPUSetup::{depth,vertex_map,normal_map}_tmp */
int mf_download_model_maps(mf_context* ctx, int i, int level, float* vmap_planar, float* nmap_planar);          /* RGBDOdometry::vmaps_g_prev_/nmaps_g_prev_ */
int mf_download_index_map(mf_context* ctx, int i, uint32_t* idx, float* vert_conf4, float* color_time4, float* norm_rad4); /* ModelProjection sparse* textures */
int mf_download_prediction(mf_context* ctx, int i, uint8_t* image4, float* vertex_conf4, float* normal_rad4, uint16_t* time); /* splat textures */
int mf_download_fill_in(mf_context* ctx, int i, uint8_t* image4, float* vertex4, float* normal4);  /* FillIn textures */
int mf_download_association(mf_context* ctx, int i, uint8_t* update_id, uint32_t* best, float* meas12); /* x-major pixel order, Model.cpp:179-183 */
int mf_download_track_stats(mf_context* ctx, int i, double* A36, double* b6, float* err_count6);   /* RGBDOdometry::lastA/lastb, lastICPError,... */
int mf_download_edge_map(mf_context* ctx, float* edge, uint8_t* binary);                           /* MfSegmentation floatEdgeMap / binary edge map */
/* test hook: morphological close of a host image (W x H of the context, in place). ellipse != 0: cv::morphologyEx(MORPH_CLOSE, MORPH_ELLIPSE) of the
 * mask-id image (MfSegmentation.cpp:424-426); ellipse == 0: the binary edge-map close (segmentation.cu:217-255,334-354), `inverted` = 255 - result. */
/* Mask R-CNN backbone on the frame path (MaskRCNN::executeSequential, MaskRCNN.cpp:147-151, called at MfSegmentation.cpp:130): every k-th
 * processFrame letter-boxes the frame's RGB image into `backbone`'s input (mf_backbone_mold) and enqueues its forward on the backbone's
 * own stream, concurrently with the dense pipeline on the same GPU.  backbone = handle of mf_backbone_create; NULL detaches. */
int mf_attach_backbone(mf_context* ctx, void* backbone, int every_k);
int mf_debug_track_timing(int64_t* out, int cap);   /* profiling builds only (-DMF_TRACK_TIMING): (tag, SM clock) pairs of the last tracking launch; else 0 */
int mf_morph_close(mf_context* ctx, uint8_t* image, int radius, int iterations, int ellipse, uint8_t* inverted);

/* ---- stand-alone kernels exposed for parity tests (device work, host buffers) ---- */
/* ---- multi-model inputs / outputs ---- */
int mf_set_frame_classes(mf_context* ctx, const int32_t* class_ids, int n);   /* FrameData::classIDs of the NEXT processFrame call (classIDs[mask value]; entry 0 = background), Core/FrameData.h:40 */
int mf_download_segmentation(mf_context* ctx, uint8_t* mask, uint8_t* projected_ids);   /* SegmentationResult::fullSegmentation (textureMask) and GlobalProjection::getProjectedModelIDs */
int mf_model_class_id(mf_context* ctx, int i);                                /* Model::getClassID */

/* ---- object-sharded mode: one context per GPU/rank, object Models (and their surfel stores) partitioned over the ranks
 *      (BASELINE.json north_star "Object Models ... shard one-per-GPU"); the couplings of a frame are exactly those of
 *      MaskFusion.cpp:212-217 (every model reads the frame), :257-276 (tracked poses decide inactivation; static objects follow the
 *      global pose), GlobalProjection.cpp:66-95 (one depth-tested ID image) and MaskFusion.cpp:296-297 (one segmentation, evaluated
 *      identically on every rank from the merged image).
 *
 *      (1) In-library exchange (the product path): after mf_shard_comm_init every rank calls mf_shard_process_frame once per frame;
 *          only rank 0 passes inputs.  The library issues, on the context's stream and without any host synchronisation inside the
 *          frame: ncclBroadcast of the frame packet (rgb | depth | mask | header, one buffer), ncclAllGather of the pose rows of the
 *          tracked models, ncclAllReduce(ncclMin, ncclUint64) of the ID-projection keys.  libnccl.so.2 is opened at run time.
 *          Bootstrap: rank 0 calls mf_shard_unique_id and distributes the 128 bytes by any side channel (the tests and bench.py use
 *          torch.distributed's store); the communicator lives in the context.
 *      (2) Transport-agnostic phases (tests over gloo, one GPU or none of the NCCL requirements): the caller moves the data itself
 *              [frame to every rank]  mf_set_frame_classes; mf_shard_frame_begin;
 *              mf_shard_get_poses -> [all-gather of 64 x 32 floats] -> mf_shard_set_poses;
 *              mf_shard_project -> [MIN all-reduce of the uint64 keys at mf_shard_projection_keys];
 *              mf_shard_frame_end
 *      With world == 1 the three phase calls are exactly mf_process_frame. ---- */
int mf_shard_configure(mf_context* ctx, int rank, int world);                 /* before the first frame; rank 0 owns the background model */
int mf_shard_unique_id(uint8_t* out128);                                      /* ncclGetUniqueId (rank 0) */
int mf_shard_comm_init(mf_context* ctx, const uint8_t* id128, int rank, int world);   /* ncclCommInitRank on the context's device (+ mf_shard_configure) */
int mf_shard_process_frame(mf_context* ctx, const uint8_t* rgb, const float* depth, int64_t timestamp, const uint8_t* mask,
                           const int32_t* class_ids, int n_class_ids, float weight_multiplier,
                           int inputs_on_device);   /* inputs are read on rank 0 only (class_ids always from the host) */
int mf_shard_stats(mf_context* ctx, int64_t* out4);                           /* collective bytes moved, NCCL calls, communicator size, NCCL version */
int mf_shard_frame_begin(mf_context* ctx, const void* rgb, const void* depth, int64_t timestamp, const void* mask, int on_device);
int mf_shard_get_poses(mf_context* ctx, float* out_64x32, int capacity_models);   /* returns nModels; row i = pose(16) lastTransform(16) of model i if it is tracked here, else 0 */
int mf_shard_set_poses(mf_context* ctx, const float* gathered_world_x_64_x32);
int mf_shard_project(mf_context* ctx);                                        /* device-side lifecycle after tracking + local models into the key image */
void* mf_shard_projection_keys(mf_context* ctx);                              /* device pointer, width*height uint64 (depth bits << 32 | model index << 26 | surfel) */
int mf_shard_frame_end(mf_context* ctx, float weight_multiplier);
int mf_model_owner(mf_context* ctx, int i);                                   /* rank holding model i's surfels */
/* CTAs of the persistent tracking launch per tracked model (host only): bit j of light_mask marks an object model (validity bitmask, nearly all
 * pixels culled), the others are full-frame models and get `ratio` times the share.  Model::performTracking of a batch of models (Model.cpp:427-447)
 * is ONE launch here; this is how its grid is dealt.  n_jobs <= 32. */
int mf_track_shares(int n_jobs, unsigned light_mask, int total_ctas, int ratio, int* shares);
int mf_shard_pick_owner(const int64_t* loads, int world);                     /* placement rule for a new model: least owned capacity, ties -> highest rank (host only) */

/* in-stream CUDA-event stage timer (replaces the reference's TICK/TOCK Stopwatch, Core/Utils/Stopwatch.h:46-54) */
int mf_set_profiling(mf_context* ctx, int on);
int mf_get_stage_times(mf_context* ctx, char* buf, int bufsize);   /* lines: "name count total_ms" */
int mf_debug_set_poses(mf_context* ctx, int i, const float pose16[16], const float last_pose16[16]);   /* sets Model::pose and Model::lastPose verbatim */
int mf_icp_step(mf_context* ctx, int i, int level, const float Rcurr9[9], const float tcurr3[3], float out29[29]);  /* icpStep, reduce.cu:446-525 */

/* ---- Mask R-CNN backbone: ResNet-101 + FPN as tcgen05/TMEM GEMMs (replaces the dense part of the Keras/TF sidecar,
 *      Core/Segmentation/MaskRCNN/MaskRCNN.py.in:55-58,101-111; weights are synthetic/seeded: no COCO weights offline) ---- */
typedef struct mf_backbone mf_backbone;
const char* mf_cnn_last_error(void);
/* out[MxN] = relu?(A[MxK] * B[NxK]^T + bias[N] + residual[MxN]); bf16 device pointers, K % 64 == 0, N % 64 == 0 */
int mf_gemm_bf16(const void* dA, const void* dB, const float* dBias, const void* dResidual, void* dOut, int M, int N, int K, int relu, void* stream);
/* implicit-GEMM 3x3/s1/p1 convolution, NHWC bf16, weights [Cout][3][3][Cin]; the activation is read through a 3-D TMA map (no im2col) */
int mf_conv3x3_bf16(const void* dIn, const void* dW, const float* dBias, const void* dResidual, void* dOut, int H, int W, int Cin, int Cout, int relu, void* stream);
mf_backbone* mf_backbone_create(int input_size, unsigned seed, void* stream);
void mf_backbone_destroy(mf_backbone* h);
int mf_backbone_num_layers(mf_backbone* h);
int mf_backbone_layer(mf_backbone* h, int i, int* cin_cout_k_stride_pad_kpad);
int mf_backbone_get_weights(mf_backbone* h, int i, float* w_cout_kpad, float* bias);
int mf_backbone_mold(mf_backbone* h, const void* d_rgba, int W, int H);            /* letter-box + mean subtraction -> network input */
void* mf_backbone_stream(mf_backbone* h);   /* the cudaStream_t the backbone enqueues on */
void* mf_backbone_input_buffer(mf_backbone* h);
int mf_backbone_forward(mf_backbone* h, const void* d_input_nhwc_bf16);
void* mf_backbone_output(mf_backbone* h, int level, int* dims_hwc);               /* 0..3 = C2..C5, 4..8 = P2..P6 (device, NHWC bf16) */
int mf_backbone_download(mf_backbone* h, int level, void* host_bf16);
double mf_backbone_flops(mf_backbone* h);
int mf_backbone_num_gemms(mf_backbone* h);

/* ---- image-directory loader ("-dir", GUI/Tools/ImageLogReader.{h,cpp}; GUI/MainController.cpp:150-176) ----
 * colour .png/.ppm/.jpg, depth 16-bit .png (x 0.001), masks 8-bit .png/.pgm + "<mask>.txt" (class ids, boxes); .exr depth is refused
 * (no OpenEXR in this build).  hasMore() lets the last frame through (ImageLogReader.cpp:326), unlike the .klg reader. */
typedef struct mf_dir mf_dir;
mf_dir* mf_dir_open(const char* color_dir, const char* depth_dir, const char* mask_dir /* NULL: no masks */, int index_width /* <=0: 4 */,
                    const char* color_prefix, const char* depth_prefix, const char* mask_prefix);   /* ImageLogReader::ImageLogReader */
int mf_dir_num_frames(mf_dir* r);                                       /* ImageLogReader::getNumFrames */
int mf_dir_has_more(mf_dir* r);                                         /* ImageLogReader::hasMore */
int mf_dir_has_masks(mf_dir* r);                                        /* LogReader::hasMasks */
int mf_dir_set_max_masks(mf_dir* r, int n);                             /* ImageLogReader::setMaxMasks ("-nm") */
int mf_dir_size(mf_dir* r, int* width, int* height);                    /* size of the first colour image */
/* ImageLogReader::getNext + loadFrameFromDrive: rgb HxWx3, depth HxW metres; mask/class_ids/boxes may be NULL.  *n_class_ids: in =
 * capacity, out = ids read (classIDs[0] == 0, ImageLogReader.cpp:306); boxes = cv::Rect x,y,w,h per object.  Returns 1 if a mask was
 * delivered, 0 if not, < 0 on error; timestamp = index * 1000 / 24 (ImageLogReader.cpp:283). */
int mf_dir_get_next(mf_dir* r, uint8_t* rgb, float* depth, uint8_t* mask, int32_t* class_ids, int32_t* boxes, int* n_class_ids, int64_t* timestamp);
void mf_dir_close(mf_dir* r);

/* MaskFusion::exportPoses (Core/MaskFusion.cpp:849-881): writes <export_dir>poses-<model id>.txt for every active model
 * and for every inactivated model the reference's keep rule retained (inactivateModel, MaskFusion.cpp:699-713: smart delete keeps a model
 * with >= 4000 surfels and confidence threshold > 0.3) ("seconds x y z qx qy qz qw", fixed notation, 6 decimals); returns the number of files. */
int mf_export_poses(mf_context* ctx, const char* export_dir);

/* PLY export of one model's surfels as MaskFusion::savePly writes it (Core/MaskFusion.cpp:733-848): vertices with conf > threshold,
 * binary little endian, x y z | r g b | -nx -ny -nz | radius.  surfels = n x 12 floats as mf_download_surfels returns them. */
int mf_write_ply(const char* path, const float* surfels, int n, float conf_threshold);

/* PreSegmentation::performSegmentation (Core/Segmentation/PreSegmentation.cpp:28-90, the "precomputed masks" performer; host code in the
 * reference too): mask values -> model ids through the persistent table `mapping` (256 bytes, zero-initialised by the caller before the first
 * frame; the reference's function-static vector), the first unseen value in raster order becomes next_model_id when allow_new.  Outputs: the full
 * segmentation (W*H), has_new_label, and per model in list order (the new one last) superPixelCount, depthMean, depthStd (mean absolute deviation),
 * the inputs of Model::setMaxDepth (MaskFusion.cpp:291,337-341).  model_ids: ids of the live models, background first.  Returns the number of
 * entries written.  Not wired into the device-driven schedule (its statistics are sequential float sums in raster order). */
int mf_pre_segmentation(const uint8_t* mask, const float* depth, int W, int H, const uint8_t* model_ids, int n_models, int next_model_id,
                        int allow_new, uint8_t* mapping, uint8_t* full_segmentation, int* has_new_label, uint32_t* super_pixel_count,
                        float* depth_mean, float* depth_std);

/* Mask R-CNN post-processing (Core/Segmentation/MaskRCNN/helpers.py:70-98 generate_id_image): detections (masks HxWxN u8, N fastest;
 * scores; class ids; rois N x 4) -> id image HxW (ids 1..n in export order, later detections overwrite earlier ones), exported class ids
 * and rois.  class_filter / special_assignments may be NULL with count 0.  Returns the number of exported detections. */
int mf_generate_id_image(const uint8_t* masks, int H, int W, int N, const float* scores, const int32_t* class_ids, const int32_t* rois,
                         double min_score, const int32_t* class_filter, int n_filter, const int32_t* special_assignments, int n_special,
                         uint8_t* id_image, int32_t* exported_class_ids, int32_t* exported_rois);

/* baseline JPEG -> 8-bit RGB exactly as libjpeg's default decode path produces it (islow IDCT, fancy upsampling; mf_jpeg.cu).
 * out == NULL: only the size.  Used by both loaders; exported for the decoder's own parity test. */
int mf_decode_jpeg(const uint8_t* data, int size, uint8_t* out, int capacity, int* width, int* height);

/* OpenEXR scan-line file (HALF / FLOAT channels, compression NONE / RLE / ZIPS / ZIP) -> the float depth image the -dir reader delivers:
 * what the reference keeps of cv::imread(path, IMREAD_UNCHANGED), GUI/Tools/ImageLogReader.cpp:251-258 (CV_32FC1 as is, element 0 = the
 * B channel of a CV_32FC3 result).  capacity in floats; out == NULL: only the size.  Exported for the decoder's own parity test. */
int mf_decode_exr_depth(const uint8_t* data, int size, float* out, int capacity, int* width, int* height);

/* ---- .klg log reader / writer (GUI/Tools/KlgLogReader.cpp:29-113) ---- */
typedef struct mf_klg mf_klg;
mf_klg* mf_klg_open(const char* path, int width, int height, int flip_colors);
int mf_klg_num_frames(mf_klg* k);
int mf_klg_has_more(mf_klg* k);                                         /* KlgLogReader::hasMore (N11: last frame never read) */
int mf_klg_get_next(mf_klg* k, uint8_t* rgb, float* depth, int64_t* timestamp);   /* KlgLogReader::getNext + readFrame */
void mf_klg_close(mf_klg* k);
int mf_klg_write(const char* path, int width, int height, int num_frames, const int64_t* timestamps,
                 const uint16_t* depth_mm, const uint8_t* rgb);          /* raw (uncompressed) log */

#ifdef __cplusplus
}
#endif
#endif
