/*
 * image_stitching.h - C ABI of the B200-native compositing path (libisb.so).
 *
 * The reference (a1q123456/image_stitching) has no library API: image_stitching/image_stitching.h:4-8
 * declares nothing and the compositing path is an inline loop in main()
 * (image_stitching/image_stitching.cpp:1086-1229) that drives OpenCV's
 * cv::detail::RotationWarper / ExposureCompensator / MultiBandBlender virtual interfaces.
 * This header gives that path the entry points a maintainer would bind instead: one C function
 * per OpenCV method the loop calls (same names, argument order, units, ROI/mask conventions and
 * assertion behaviour), plus the fused whole-loop call isb_compose().  Every function cites the
 * reference line it replaces.
 *
 * Conventions
 *  - All matrices are row-major.  K and R are 3x3 float32, exactly what the reference passes after
 *    `cameras[i].K().convertTo(K, CV_32F)` (image_stitching.cpp:1135-1136, 1150-1151).
 *  - Images are 8UC3 (or 8UC1 masks) interleaved HWC with a byte pitch; 16SC3 where OpenCV uses CV_16SC3.
 *  - Every data pointer may be a HOST pointer or a CUDA DEVICE pointer; the library detects which
 *    (cudaPointerGetAttributes).  Host data is copied to device staging buffers (fastest from pinned memory);
 *    device data is used in place.
 *  - Functions return ISB_OK (0) or a negative code whose value is the OpenCV error class the reference
 *    would have thrown (cv::Error::Code); isb_last_error() gives the message (thread-local).
 *    No exception crosses this boundary.  Handles are not thread-safe (same as the reference's objects).
 *  - There is no CPU fallback: without a CUDA device every compute entry point fails with
 *    ISB_ERR_GPU_API.
 */
#ifndef IMAGE_STITCHING_B200_H
#define IMAGE_STITCHING_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(_WIN32)
#define ISB_API __declspec(dllexport)
#else
#define ISB_API __attribute__((visibility("default")))
#endif

/* ---- status codes (values follow cv::Error::Code) -------------------------------------------- */
enum {
    ISB_OK = 0,
    ISB_ERR_NO_MEM = -4,          /* StsNoMem */
    ISB_ERR_BAD_ARG = -5,         /* StsBadArg */
    ISB_ERR_NULL_PTR = -27,       /* StsNullPtr */
    ISB_ERR_UNMATCHED_SIZES = -209,
    ISB_ERR_OUT_OF_RANGE = -211,
    ISB_ERR_NOT_IMPLEMENTED = -213,
    ISB_ERR_ASSERT = -215,        /* StsAssert: what CV_Assert raises */
    ISB_ERR_GPU_NOT_SUPPORTED = -216,
    ISB_ERR_GPU_API = -217,       /* GpuApiCallError */
    ISB_ERR_IO = -2               /* StsError: file could not be opened / parsed */
};

enum { ISB_WARP_SPHERICAL = 0, ISB_WARP_CYLINDRICAL = 1 };      /* warp_type, image_stitching.cpp:64,919-971 */
enum { ISB_INTER_NEAREST = 0, ISB_INTER_LINEAR = 1 };            /* cv::INTER_NEAREST / cv::INTER_LINEAR */
enum { ISB_BORDER_CONSTANT = 0, ISB_BORDER_REFLECT = 2 };        /* cv::BORDER_CONSTANT / cv::BORDER_REFLECT */
enum { ISB_EULER_XYZ = 0, ISB_EULER_YXZ, ISB_EULER_ZXY, ISB_EULER_ZYX, ISB_EULER_YZX, ISB_EULER_XZY }; /* euler_order.h:3-11 */

ISB_API const char* isb_last_error(void);
ISB_API const char* isb_version(void);
/* number of CUDA devices visible (0 => every compute call fails loudly) */
ISB_API int isb_device_count(void);
/* stream all subsequent work of the calling thread's handles is enqueued on (cudaStream_t; NULL = default stream) */
ISB_API int isb_set_stream(void* cuda_stream);
/* the measurement switches (ISB_PDL, ISB_BLEND_PIPE, ISB_BLEND_TMA, ISB_STAGED_STORES) are read from the
 * environment once per process; a tool that changes them afterwards calls this to have them re-read */
ISB_API void isb_reload_env(void);
/* counts kernel launches made by this library since the last reset (bench.py's `gpu_launches`) */
ISB_API long long isb_launch_count(int reset);

/* ============================================================================================
 * Camera types and host helpers (quaternion.h, euler.h, serializer.cpp)
 * ============================================================================================ */

/* cv::detail::CameraParams as the compositing loop uses it (image_stitching.cpp:1123-1136). */
typedef struct isb_camera {
    double focal, aspect, ppx, ppy;
    float R[9]; /* CV_32F after image_stitching.cpp:626-638 */
    float t[3];
} isb_camera;

/* CameraParams::K() converted to CV_32F: [[focal,0,ppx],[0,focal*aspect,ppy],[0,0,1]] */
ISB_API void isb_camera_K(const isb_camera* cam, float K[9]);

/* Quaternion<double> members the reference uses (quaternion.h:260-322, 564-596, 172-239, 241-258, 464-478, 480-544).
 * q = (x, y, z, w). */
ISB_API void isb_quat_from_rotation_matrix(const double R[9], double q[4]);
ISB_API void isb_quat_to_rotation_matrix(const double q[4], double R[9]);
ISB_API void isb_quat_from_euler(const double euler_xyz[3], int order, double q[4]);
ISB_API void isb_quat_from_axis_angle(const double axis[3], double angle, double q[4]);
ISB_API void isb_quat_multiply(const double a[4], const double b[4], double out[4]);
ISB_API void isb_quat_slerp(const double a[4], const double b[4], double t, double out[4]);
/* the reference's EXIF pose fix-up (image_stitching.cpp:485-517): R3x3 -> quaternion -> component flips
 * (portrait: (y,x,-z,w); landscape: (-x,y,-z,w)) -> rotation matrix */
ISB_API void isb_pose_from_cam_transform(const double R_in[9], int is_portrait, double R_out[9]);

/* euler.h:4-133 / :135-300 (double) */
ISB_API int isb_rotation_matrix_to_euler(const double R[9], int order, double euler_xyz[3]);
ISB_API int isb_euler_to_rotation_matrix(const double euler_xyz[3], int order, double R[9]);

/* serializer.cpp.  Matrices come back as float (deserializeMatrix yields CV_32F, :69-111) or
 * double (parseMatrixStr yields CV_64F square, :22-36). */
ISB_API int isb_parse_matrix_str(const char* s, double* out, int capacity, int* side);
ISB_API int isb_serialize_matrix(const double* m, int rows, int cols, int is_f32, char* buf, size_t cap);
ISB_API int isb_deserialize_matrix(const char* s, float* out, int capacity, int* rows, int* cols);
/* cams.data (serializer.cpp:113-167) and indices.data (:169-193); path NULL => "./cams.data" / "./indices.data" */
ISB_API int isb_save_cams(const char* path, const isb_camera* cams, int n);
ISB_API int isb_load_cams(const char* path, isb_camera* cams, int capacity, int* n);
ISB_API int isb_save_indices(const char* path, const int* idx, int n);
ISB_API int isb_load_indices(const char* path, int* idx, int capacity, int* n);

/* ============================================================================================
 * cv::detail::RotationWarper  (created by WarperCreator::create(scale), image_stitching.cpp:973,1117)
 * ============================================================================================ */
typedef struct isb_warper isb_warper;

ISB_API isb_warper* isb_warper_create(int kind, float scale);
ISB_API void isb_warper_destroy(isb_warper* w);
ISB_API float isb_warper_get_scale(const isb_warper* w);
ISB_API int isb_warper_set_scale(isb_warper* w, float scale);

/* Rect roi = warper->warpRoi(sz, K, R)  (image_stitching.cpp:1138) -> rect = {x, y, width, height} */
ISB_API int isb_warper_warp_roi(isb_warper* w, int src_w, int src_h, const float K[9], const float R[9], int rect_xywh[4]);
/* Point2f warpPoint(pt, K, R) / warpPointBackward(pt, K, R) */
ISB_API int isb_warper_warp_point(isb_warper* w, const float pt[2], const float K[9], const float R[9], float out[2]);
ISB_API int isb_warper_warp_point_backward(isb_warper* w, const float pt[2], const float K[9], const float R[9], float out[2]);
/* Rect buildMaps(src_size, K, R, xmap, ymap): xmap/ymap are rect.height x rect.width float32 with the given
 * pitch in BYTES, computed by the same device code the fused warp uses (there the maps are never stored). */
ISB_API int isb_warper_build_maps(isb_warper* w, int src_w, int src_h, const float K[9], const float R[9],
                                  float* xmap, float* ymap, size_t pitch_bytes, int rect_xywh[4]);
/* Point warp(src, K, R, interp_mode, border_mode, dst)  (image_stitching.cpp:1154,1159,985,988).
 * src: 8UC1 or 8UC3 (channels = 1|3).  dst must hold rect.height rows of rect.width px (rect from
 * isb_warper_warp_roi).  Supported (what the reference calls): INTER_LINEAR+BORDER_REFLECT,
 * INTER_NEAREST+BORDER_CONSTANT, and the two cross combinations. */
ISB_API int isb_warper_warp(isb_warper* w, const uint8_t* src, int src_w, int src_h, int channels, size_t src_pitch,
                            const float K[9], const float R[9], int interp_mode, int border_mode, uint8_t* dst,
                            size_t dst_pitch, int corner_xy[2]);

/* void warpBackward(src, K, R, interp_mode, border_mode, dst_size, dst): the inverse of warp().  src is a warped image of
 * exactly warpRoi(dst_size) pixels (OpenCV asserts it), dst the original camera frame of dst_w x dst_h pixels.  The
 * forward map's atan2f / acosf are evaluated per pixel on the device with glibc's float algorithms (bit-identical maps). */
ISB_API int isb_warper_warp_backward(isb_warper* w, const uint8_t* src, int src_w, int src_h, int channels, size_t src_pitch,
                                     const float K[9], const float R[9], int interp_mode, int border_mode, int dst_w, int dst_h,
                                     uint8_t* dst, size_t dst_pitch);

/* ============================================================================================
 * cv::detail::BlocksGainCompensator - apply side only  (image_stitching.cpp:1162).
 * feed() (gain estimation) stays on the reference CPU path; its result enters through set_mat_gains
 * (== ExposureCompensator::setMatGains).
 * ============================================================================================ */
typedef struct isb_compensator isb_compensator;

ISB_API isb_compensator* isb_compensator_create(int block_w, int block_h);
ISB_API void isb_compensator_destroy(isb_compensator* c);
ISB_API int isb_compensator_set_mat_gains(isb_compensator* c, int n, const float* const* gains, const int* gain_w,
                                          const int* gain_h);
ISB_API int isb_compensator_get_mat_gain(const isb_compensator* c, int index, float* out, int capacity, int* gain_w, int* gain_h);
/* compensator->apply(index, corner, image, mask): image 8UC3 in place; mask is accepted and ignored, as OpenCV does */
ISB_API int isb_compensator_apply(isb_compensator* c, int index, const int corner_xy[2], uint8_t* image, int w, int h,
                                  size_t pitch, const uint8_t* mask, size_t mask_pitch);

/* ============================================================================================
 * Seam-mask preparation of the loop: dilate(3x3) -> resize(INTER_LINEAR_EXACT) -> AND
 * (image_stitching.cpp:1169-1171).  mask_warped (w x h) is updated in place.
 * ============================================================================================ */
ISB_API int isb_seam_mask_apply(const uint8_t* seam_mask, int seam_w, int seam_h, size_t seam_pitch,
                                uint8_t* mask_warped, int w, int h, size_t pitch);

/* ============================================================================================
 * Ingest pre-steps of the loop: the rotation every decoded image gets and the
 * compose-scale down-sizing.  Both are bit-exact integer operations.
 * ============================================================================================ */
enum { ISB_ROTATE_90_CLOCKWISE = 0, ISB_ROTATE_180 = 1 };      /* cv::ROTATE_90_CLOCKWISE / cv::ROTATE_180 */
/* rotate(full_img, tmp, ROTATE_90_CLOCKWISE | ROTATE_180)  (image_stitching.cpp:1093-1103, 569-581).
 * dst is (src_h x src_w) pixels wide x high for ROTATE_90_CLOCKWISE, (src_w x src_h) for ROTATE_180. */
ISB_API int isb_rotate(const uint8_t* src, int src_w, int src_h, int channels, size_t src_pitch, int rotate_code,
                       uint8_t* dst, size_t dst_pitch);
/* cv::resize(full_img, img, Size(), compose_scale, compose_scale, INTER_LINEAR_EXACT)  (image_stitching.cpp:1143-1146).
 * Pass fx, fy > 0 for the scale-factor form (dst_w = cvRound(src_w * fx), which the caller computes) or fx = fy = 0
 * for the dsize form (as used for the seam mask, :1170). */
ISB_API int isb_resize_linear_exact(const uint8_t* src, int src_w, int src_h, int channels, size_t src_pitch, uint8_t* dst,
                                    int dst_w, int dst_h, size_t dst_pitch, double fx, double fy);

/* ============================================================================================
 * cv::detail::MultiBandBlender  (image_stitching.cpp:1173-1225)
 * ============================================================================================ */
typedef struct isb_blender isb_blender;

/* Rect resultRoi(corners, sizes)  (image_stitching.cpp:1176) */
ISB_API int isb_result_roi(const int* corners_xy, const int* sizes_wh, int n, int rect_xywh[4]);
/* the reference's band-count rule (image_stitching.cpp:1177-1183): returns -1 when blend_width < 1 (Blender::NO) */
ISB_API int isb_num_bands_for(int dst_w, int dst_h, float blend_strength);

ISB_API isb_blender* isb_blender_create(int num_bands); /* MultiBandBlender(try_gpu, num_bands=5, CV_32F) */
ISB_API void isb_blender_destroy(isb_blender* b);
ISB_API int isb_blender_set_num_bands(isb_blender* b, int num_bands);
ISB_API int isb_blender_num_bands(const isb_blender* b);        /* the REQUESTED value, as OpenCV returns */
ISB_API int isb_blender_actual_num_bands(const isb_blender* b); /* after prepare(): min(requested, ceil(log2(max(w,h)))) */
/* blender->prepare(corners, sizes) == prepare(resultRoi(corners, sizes)) */
ISB_API int isb_blender_prepare(isb_blender* b, const int* corners_xy, const int* sizes_wh, int n);
ISB_API int isb_blender_prepare_roi(isb_blender* b, const int rect_xywh[4]);
/* dst_roi_ (padded to 2^num_bands) and dst_roi_final_ */
ISB_API int isb_blender_get_rois(const isb_blender* b, int padded_xywh[4], int final_xywh[4]);
/* the rect feed() builds its pyramids on: {tl.x, tl.y, br.x, br.y} in panorama coordinates */
ISB_API int isb_blender_tile_rect(const isb_blender* b, int w, int h, int tl_x, int tl_y, int rect_tlbr[4]);
/* blender->feed(img CV_16SC3, mask CV_8U, tl) */
ISB_API int isb_blender_feed(isb_blender* b, const int16_t* img, size_t img_pitch, const uint8_t* mask, size_t mask_pitch,
                             int w, int h, int tl_x, int tl_y);
/* blender->blend(dst, dst_mask): dst CV_16SC3 and dst_mask CV_8U of dst_roi_final_ size.  Single use per prepare(). */
ISB_API int isb_blender_blend(isb_blender* b, int16_t* dst, size_t dst_pitch, uint8_t* dst_mask, size_t mask_pitch);

/* ============================================================================================
 * cv::detail::Blender (Blender::NO) and cv::detail::FeatherBlender  (image_stitching.cpp:1175-1191:
 * Blender::createDefault(blend_type), the blend_width < 1 fall-back to Blender::NO, FeatherBlender::setSharpness)
 * ============================================================================================ */
enum { ISB_BLENDER_NO = 0, ISB_BLENDER_FEATHER = 1, ISB_BLENDER_MULTI_BAND = 2 }; /* == cv::detail::Blender::{NO,FEATHER,MULTI_BAND} */
typedef struct isb_simple_blender isb_simple_blender;
/* Blender::createDefault(type) for type NO / FEATHER (MULTI_BAND is isb_blender above); sharpness as FeatherBlender(0.02f) */
ISB_API isb_simple_blender* isb_simple_blender_create(int type, float sharpness);
ISB_API void isb_simple_blender_destroy(isb_simple_blender* b);
ISB_API int isb_simple_blender_set_sharpness(isb_simple_blender* b, float sharpness); /* fb->setSharpness(1.f / blend_width) */
ISB_API float isb_simple_blender_sharpness(const isb_simple_blender* b);
ISB_API int isb_simple_blender_prepare(isb_simple_blender* b, const int* corners_xy, const int* sizes_wh, int n);
ISB_API int isb_simple_blender_prepare_roi(isb_simple_blender* b, const int rect_xywh[4]);
/* feed(img CV_16SC3, mask CV_8U, tl); the image rect must lie inside the prepared ROI (CV_Assert in the reference) */
ISB_API int isb_simple_blender_feed(isb_simple_blender* b, const int16_t* img, size_t img_pitch, const uint8_t* mask,
                                    size_t mask_pitch, int w, int h, int tl_x, int tl_y);
/* blend(dst CV_16SC3, dst_mask CV_8U) of the prepared ROI size.  Single use per prepare(). */
ISB_API int isb_simple_blender_blend(isb_simple_blender* b, int16_t* dst, size_t dst_pitch, uint8_t* dst_mask, size_t mask_pitch);
/* cv::detail::createWeightMap(mask, sharpness, weight): weight = min(1, sharpness * distanceTransform(mask, DIST_L1, 3)) */
ISB_API int isb_create_weight_map(const uint8_t* mask, size_t mask_pitch, int w, int h, float sharpness, float* weight,
                                  size_t weight_pitch);

/* cv::detail::Timelapser (image_stitching.cpp:1194-1215: Timelapser::createDefault(timelapse_type), initialize(corners,
 * sizes), process(img_warped_s, ones, corners[img_idx]), getDst()) */
enum { ISB_TIMELAPSER_AS_IS = 0, ISB_TIMELAPSER_CROP = 1 }; /* == cv::detail::Timelapser::{AS_IS, CROP} */
typedef struct isb_timelapser isb_timelapser;
ISB_API isb_timelapser* isb_timelapser_create(int type);
ISB_API void isb_timelapser_destroy(isb_timelapser* t);
ISB_API int isb_timelapser_initialize(isb_timelapser* t, const int* corners_xy, const int* sizes_wh, int n, int dst_roi_xywh[4]);
/* process(img CV_16SC3, mask (ignored by OpenCV), tl): clears the canvas, copies the pixels of img that fall inside dst_roi */
ISB_API int isb_timelapser_process(isb_timelapser* t, const int16_t* img, size_t img_pitch, int w, int h, int tl_x, int tl_y);
/* getDst(): CV_16SC3 of dst_roi size */
ISB_API int isb_timelapser_get_dst(isb_timelapser* t, int16_t* dst, size_t dst_pitch);

/* ============================================================================================
 * Fused path: the whole loop image_stitching.cpp:1086-1229 (warp, mask, gain, ->16S, seam mask,
 * prepare, feed x n, blend, saturate to 8U) without materialising xmap/ymap or the intermediates.
 * ============================================================================================ */
typedef struct isb_image { const uint8_t* data; int width, height; size_t pitch; } isb_image;   /* 8UC3 */
typedef struct isb_gainmap { const float* data; int width, height; } isb_gainmap;                /* f32 grid; data NULL = no gain */
typedef struct isb_mask { const uint8_t* data; int width, height; size_t pitch; } isb_mask;      /* 8UC1 seam mask; data NULL = all 255 */

enum { ISB_GATHER_PEER_STORES = 0, ISB_GATHER_COPY_ENGINE = 1, ISB_GATHER_LOCAL = 2 };
typedef struct isb_config {
    int warp_kind;            /* ISB_WARP_*  (warp_type, image_stitching.cpp:64) */
    float warped_image_scale; /* warper scale (image_stitching.cpp:1116-1117) */
    int num_bands;            /* MultiBandBlender::setNumBands (image_stitching.cpp:1183) */
    int strip_index;          /* this process's strip, 0 <= strip_index < strip_count */
    int strip_count;          /* 1 = whole panorama on this GPU */
    int cache_plan;           /* !=0: keep geometry (ROIs, trig tables, tile lists) across calls with equal cameras */
    int async_mode;           /* !=0: isb_composer_run() only enqueues (pinned host buffers must stay valid until
                                 isb_composer_sync()); lets two composers on two streams overlap one step's upload with
                                 the previous step's download */
    int gather_mode;          /* strip-sharded runs with a device-resident isb_pano (ISB_GATHER_*):
                                 PEER_STORES (0): the final kernel stores into isb_pano directly, as 16-byte vectors staged in
                                   shared memory - over NVLink when isb_pano is rank 0's peer-mapped panorama;
                                 COPY_ENGINE (1): the strip is composed into a local double-buffered block and pushed into
                                   isb_pano (rank 0's panorama) by the copy engine on a second stream, overlapping the next
                                   run's kernels; the rows have landed after isb_composer_sync() / behind isb_composer_join();
                                 LOCAL (2): isb_pano is this GPU's own memory (rank 0): direct stores. */
    int pipeline_depth;       /* > 1: that many runs may be in flight: consecutive isb_composer_run() calls are served by
                                 independent pyramid sets on internal streams and overlap on the device (the latency-bound coarse
                                 levels of one step under the issue-bound kernels of the next).  A run is then asynchronous with
                                 respect to the caller's stream - its isb_pano is valid behind isb_composer_join() (stream-ordered)
                                 or after isb_composer_sync() - and runs in flight must be given different output buffers.
                                 0 / 1: one run at a time on the caller's stream. */
    int use_blend_rule;       /* 0: MultiBandBlender with `num_bands` as given (setNumBands).  !=0: the reference's own blender
                                 set-up (image_stitching.cpp:1173-1193) from `blend_type` and `blend_strength`:
                                 blend_width = sqrt(dst_area) * blend_strength / 100; blend_width < 1 -> Blender::NO;
                                 MULTI_BAND -> num_bands = ceil(log(blend_width) / log(2)) - 1; FEATHER -> sharpness = 1 / blend_width.
                                 NO / FEATHER run the loop (warp, mask warp, gain, ->16S, seam mask, feed, blend) on the device
                                 with the per-call kernels; strips and pipeline_depth apply to MULTI_BAND only. */
    int blend_type;           /* ISB_BLENDER_NO / ISB_BLENDER_FEATHER / ISB_BLENDER_MULTI_BAND (blend_type, image_stitching.cpp:80) */
    float blend_strength;     /* blend_strength, image_stitching.cpp:81 (default 5) */
    double compose_scale;     /* ingest pre-steps of the loop inside the composer: when ingest_rotate != 0 or
                                 |compose_scale - 1| > 0.1 the isb_image arguments are the DECODED frames; every run first applies
                                 cv::rotate (image_stitching.cpp:1093-1103) and cv::resize(Size(), compose_scale, compose_scale,
                                 INTER_LINEAR_EXACT) (:1143-1146) on the device and warps the result.  isb_composer_plan() then
                                 takes the decoded sizes and derives sz = cvRound(rotated size * compose_scale) (:1130-1133); the
                                 cameras are the compose-scale ones (:1123-1125).  0 or 1: no resize. */
    int ingest_rotate;        /* 0: none; 1 + ISB_ROTATE_90_CLOCKWISE (portrait frames) or 1 + ISB_ROTATE_180 (:1093-1103) */
    int reserved[3];
} isb_config;

/* Output of isb_compose: the panorama (dst_roi_final_ size).  data/mask describe the FULL panorama buffer
 * (host, device or a peer-mapped device pointer); a strip-sharded call writes only rows
 * [strip_y0, strip_y1) of it.  data16 (CV_16SC3, what blend() returns) is optional. */
typedef struct isb_pano {
    uint8_t* data;  size_t pitch;       /* 8UC3 saturated (imwrite, image_stitching.cpp:1228) */
    uint8_t* mask;  size_t mask_pitch;  /* result_mask */
    int16_t* data16; size_t pitch16;    /* may be NULL */
    int roi_xywh[4];                    /* out: dst_roi_final_ */
    int strip_y0, strip_y1;             /* out: rows of the panorama this call produced */
} isb_pano;

typedef struct isb_composer isb_composer;
ISB_API isb_composer* isb_composer_create(const isb_config* cfg);
ISB_API void isb_composer_destroy(isb_composer* c);
/* Geometry pass of the first loop iteration (image_stitching.cpp:1116-1141 + :1176): corners[i], sizes[i], resultRoi. */
ISB_API int isb_composer_plan(isb_composer* c, const isb_camera* cams, const int* src_sizes_wh, int n,
                              int* corners_xy, int* sizes_wh, int dst_roi_xywh[4]);
/* Runs the loop on the planned geometry. gains / seam_masks may be NULL. */
ISB_API int isb_composer_run(isb_composer* c, const isb_image* imgs, const isb_gainmap* gains, const isb_mask* seam_masks,
                             int n, isb_pano* out);
/* waits for the last isb_composer_run() of this composer (needed only in async_mode / ISB_GATHER_COPY_ENGINE) */
ISB_API int isb_composer_sync(isb_composer* c);
/* stream-ordered variant for ISB_GATHER_COPY_ENGINE: work enqueued on the calling thread's stream after this call starts
 * after the strips of all previous runs have landed in isb_pano (no host synchronisation) */
ISB_API int isb_composer_join(isb_composer* c);
/* device time (ms) of the last run split by stage; names via isb_composer_stage_name; returns #stages */
ISB_API int isb_composer_last_timings(isb_composer* c, float* ms, int capacity);
ISB_API const char* isb_composer_stage_name(int stage);
/* algorithmic byte model of SURVEY.md 8(d) for the planned rig: S, M (valid warped px, counted on the device), A_p, B_alg */
ISB_API int isb_composer_byte_model(isb_composer* c, double* S_px, double* M_px, double* Ap_px, double* B_alg_bytes);
/* strip-sharded runs: rows [lo, hi] of source image `index` this strip reads (host sources are uploaded band-wise), and the
 * source bytes the last isb_composer_run() copied host -> device */
ISB_API int isb_composer_source_band(isb_composer* c, int index, int* row_lo, int* row_hi);
ISB_API long long isb_composer_last_h2d_bytes(isb_composer* c);
/* rows [y0, y1) of the panorama this composer's strip covers (after isb_composer_plan).  The planner cuts the panorama on the
 * 2^nb grid so that every strip has about the same WORK (pyramid area incl. halo + blended rows), which differs from the
 * plain arithmetic cuts of isb_strip_rows() whenever the images are not spread evenly over the rows. */
ISB_API int isb_composer_strip_rows(isb_composer* c, int* y0, int* y1);
/* one-shot convenience == create + plan + run + destroy */
ISB_API int isb_compose(const isb_image* imgs, const isb_camera* cams, const isb_gainmap* gains, const isb_mask* seam_masks,
                        int n, const isb_config* cfg, isb_pano* out);

/* ============================================================================================
 * crop() of the reference (image_stitching/cropper.cpp:116-209): the largest-interior-rectangle
 * heuristic - biggest external contour, its filled mask, a rectangle shrunk side by side until its border holds no exterior
 * pixel (checkInteriorExterior, cropper.cpp:6-104).  The reference derives its mask from the image (`gray > 0`, :118-124);
 * the north star runs it on the composited mask, so both forms are offered.  Returns the rectangle {x, y, width, height}
 * crop() would narrow the image to (source = source(croppingMask)); the caller crops by pointer arithmetic.
 * ISB_ERR_OUT_OF_RANGE when the mask is empty (the reference's contours.at(0) throws).
 * ============================================================================================ */
/* mask: 8UC1, non-zero = inside (isb_pano.mask); host or device pointer.  n_contour_points (may be NULL): size of the
 * chosen contour as cv::findContours(RETR_EXTERNAL, CHAIN_APPROX_NONE) reports it. */
ISB_API int isb_crop_rect(const uint8_t* mask, int width, int height, size_t pitch, int rect_xywh[4], int* n_contour_points);
/* the literal crop(source): image 8UC3 (is_16s = 0) or 16SC3 (is_16s = 1, what blend() returns; saturated to 8U first) */
ISB_API int isb_crop_rect_image(const void* image, int width, int height, size_t pitch_bytes, int is_16s, int rect_xywh[4],
                                int* n_contour_points);

/* ==============================================================================================
 * Output side: cv::imwrite(result_name, result) with the default "result.jpg" (image_stitching.cpp:81, :1228; the timelapse
 * frames of :1214 likewise).  OpenCV hands the image to libjpeg with its defaults - quality 95, YCbCr 4:2:0, baseline
 * sequential DCT (JDCT_ISLOW), the Annex K Huffman tables, no restart markers, JFIF 1.01 - after converting a 16S result with
 * saturate_cast<uchar>.  The whole pipeline is integer arithmetic; the device reproduces the file byte for byte.
 * image: 8UC3 (is_16s = 0) or 16SC3 (is_16s = 1, what blend() returns), BGR, host or device pointer.
 * out: host or device buffer of `capacity` bytes; *out_size receives the stream size.  If the buffer is too small (or NULL)
 * the call returns ISB_ERR_OUT_OF_RANGE with *out_size = the size needed.  quality: 1..100 (imwrite's default: 95).
 * Images beyond libjpeg's 65500-pixel limit per side are refused (ISB_ERR_OUT_OF_RANGE), as imwrite refuses them.
 * ============================================================================================ */
ISB_API int isb_jpeg_encode(const void* image, int width, int height, size_t pitch_bytes, int is_16s, int quality, uint8_t* out,
                            size_t capacity, size_t* out_size);
/* The encoder keeps its device work buffers (about 7 bytes per pixel) per calling thread between calls; this frees them. */
ISB_API int isb_jpeg_release_workspace(void);

/* Peer-memory plumbing for the fused "collapse + gather" of the strip-sharded path: rank 0 allocates the panorama
 * with isb_device_malloc and exports it; the other ranks open the handle and pass the returned pointer as
 * isb_pano.data / .mask, so that the final blend kernel stores its rows straight into rank 0's HBM over NVLink. */
ISB_API int isb_device_malloc(size_t bytes, void** dev_ptr);
ISB_API int isb_device_free(void* dev_ptr);
ISB_API int isb_ipc_get_handle(const void* dev_ptr, unsigned char handle[64]);
ISB_API int isb_ipc_open_handle(const unsigned char handle[64], void** dev_ptr);
ISB_API int isb_ipc_close_handle(void* dev_ptr);
/* cudaMemcpyAsync(cudaMemcpyDefault) on the library's stream, for moving results out of such buffers */
ISB_API int isb_memcpy(void* dst, const void* src, size_t bytes, int synchronize);

/* Arithmetic strip cuts (rows [y0,y1) of strip i of n, boundaries on the 2^nb grid, halo-aware): what the planner falls back
 * to for very short panoramas; the rows a composer actually produces are isb_composer_strip_rows() / isb_pano.strip_y0/1 */
ISB_API int isb_strip_rows(int padded_h, int final_h, int num_bands, int strip_index, int strip_count, int* y0, int* y1);

#ifdef __cplusplus
}
#endif
#endif /* IMAGE_STITCHING_B200_H */
