/* vsc.h -- C ABI of the B200-native hot path of Interactive Temporal Video Consistency.
 *
 * One shared library (libvsc_b200.so, hand-written CUDA for sm_100a) exports exactly these
 * symbols.  They are what the reference's two plug-in boundaries bind to:
 *
 *   (1) the onnxruntime custom-op library  (reference: src/ort_custom_ops)
 *         CorrelationKernel::ComputeCUDA   correlation_cuda.cc:29-138  ->  vsc_correlation_f32
 *         WarpKernel::ComputeCUDA          warp_cuda.cc:28-77          ->  vsc_warp_nchw_f32
 *   (2) the stabilization interface        (reference: src/stabilization)
 *         get_warp_result   flowconsistency.cuh:19-22  / .cu:301-312   ->  vsc_warp_hwc3
 *         get_adap_comb     flowconsistency.cuh:25-36  / .cu:314-333   ->  vsc_adap_comb
 *         get_consist_wt    flowconsistency.cuh:38-43  / .cu:335-348   ->  vsc_consist_wt
 *         get_bilinear      flowconsistency.cuh:15-17  / .cu:288-299   ->  vsc_bilinear
 *         get_consist_out   flowconsistency.cuh:45-52  / .cu:350-374   ->  vsc_consist_solve
 *         GPUImage::copyFromQImage  gpuimage.cpp:103-124 / gpuimage.cu:39-51,70-104  ->  vsc_rgba8_to_f32x3
 *         GPUImage::copyToQImage    gpuimage.cpp:127-135 / gpuimage.cu:54-67,107-135 ->  vsc_f32x3_to_rgba8
 *         VideoStabilizer::doOneStep  videostabilizer.cpp:167-265      ->  vsc_stage_a_fused + vsc_frame_solve,
 *                                                                          or the vsc_stabilizer_* object
 *
 * Conventions
 *   - plain C: pointers, ints, floats; no C++/torch/ORT types.  `stream` is a cudaStream_t
 *     passed as void* (NULL = legacy default stream).
 *   - every pointer named *_dev / documented "device" is device memory on the current device.
 *   - kernels are enqueued on `stream`; the functions never synchronise, never allocate device
 *     memory and never touch the default stream unless asked to (the reference does all three:
 *     correlation_cuda.cu:369-370,440-441; flowconsistency.cu:285,363-365).  Scratch memory is
 *     caller-owned (vsc_*_workspace_bytes) except inside the vsc_stabilizer object, which
 *     allocates once at creation.
 *   - return value: 0 (VSC_OK) or a negative VSC_E_* for argument errors, or a positive
 *     cudaError_t from the launch.  No exceptions cross this boundary, nothing calls exit().
 *   - images: float, interleaved HWC, data[(y*W + x)*C + c]  (gpuimage.h:15-20);
 *     op tensors: float NCHW, contiguous.
 *   - there is NO CPU fallback: without a CUDA device every compute entry point returns the
 *     cudaError_t of the failed launch.
 */
#ifndef VSC_VSC_H
#define VSC_VSC_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(_MSC_VER)
#define VSC_API __declspec(dllexport)
#else
#define VSC_API __attribute__((visibility("default")))
#endif

#define VSC_VERSION 100 /* 0.1.0 */

enum {
    VSC_OK = 0,
    VSC_E_INVALID = -1,   /* null pointer / non-positive size / unsupported channel count */
    VSC_E_WORKSPACE = -2, /* workspace missing or too small */
    VSC_E_STATE = -3,     /* stabilizer called out of order (e.g. step before 3 frames were pushed) */
    VSC_E_ALIGN = -4,     /* pointer not aligned to 4 bytes */
    /* .flo ingestion: one code per exception of the reference's ReadFlowFile (flowIO.cpp:31-78) and
     * initializeFlowImage (imagehelpers.cpp:42-50); vsc_error_string() returns the reference's message */
    VSC_E_FLO_OPEN = -5,    /* "ReadFlowFile: could not open" */
    VSC_E_FLO_HEADER = -6,  /* "ReadFlowFile: problem reading file" */
    VSC_E_FLO_TAG = -7,     /* "ReadFlowFile: wrong tag (possibly due to big-endian machine?)" */
    VSC_E_FLO_WIDTH = -8,   /* "ReadFlowFile: illegal width" */
    VSC_E_FLO_HEIGHT = -9,  /* "ReadFlowFile: illegal height" */
    VSC_E_FLO_SHORT = -10,  /* "ReadFlowFile: file is too short" */
    VSC_E_FLO_LONG = -11,   /* "ReadFlowFile: file is too long" */
    VSC_E_FLO_DIMS = -12    /* "Flow image size does not match image size" */
};

typedef void* vsc_stream_t; /* cudaStream_t */

VSC_API int vsc_version(void);
/* static string for VSC_E_* codes and cudaError_t values */
VSC_API const char* vsc_error_string(int code);
/* number of kernels this library has launched in this process (all entry points); the benchmark's
 * "gpu_launches" claim is read from here */
VSC_API uint64_t vsc_launch_count(void);

/* ------------------------------------------------------------------ ORT custom ops (NCHW fp32)
 *
 * custom::Correlation (domain "custom", attributes legacy, max_displacement; correlation.h:19-31)
 *   legacy == 0:  out[N, P, P, H, W],  P = 2*max_displacement+1
 *                 out[n,ph,pw,h,w] = sum_c in1[n,c,h,w] * in2[n,c,h+ph-md,w+pw-md], outside = 0
 *   legacy != 0:  out[N, P*P, H, W] = the same sums divided by C (correlation_cuda.cu:183-265);
 *                 the two layouts are byte-identical, only the 1/C scale differs.
 * in1/in2/out: device. */
VSC_API int vsc_correlation_f32(const float* in1, const float* in2, float* out, int N, int C, int H, int W,
    int max_displacement, int legacy, vsc_stream_t stream);

/* 0 (default): tiles staged by TMA when the tensors allow it (W % 4 == 0, 16-byte aligned bases; skewed 64x8
 * tiles whose lane pairs share their in2 rows for W >= 96, else 32x8), plain loads otherwise;  1: always the
 * plain-load stager;  2 / 3 / 5 / 6: TMA with 32x8 / 64x8 / skewed shared-row 64x8 / skewed shared-row 32x8 tiles
 * -- same arithmetic, bit-identical results;  4 / 7: the channel-split kernels that mode 0 uses for maps of at
 * most 12288 pixels (one pixel / one 4-pixel quad per thread, the latter when W % 4 == 0; 4 - 16 partial sums per value:
 * equal within rounding).  For tests. */
VSC_API int vsc_set_correlation_mode(int mode);

/* custom::Warp: masked bilinear backward warp (warp.cc:71-134 / warp_cuda.cu:29-84).
 * in [N,C,H,W], flow [N,2,H,W] (pixels), out [N,C,H,W]; out = 0 where the bilinear validity
 * mask is <= 0.999. */
VSC_API int vsc_warp_nchw_f32(const float* in, const float* flow, float* out, int N, int C, int H, int W,
    vsc_stream_t stream);

/* Kernel selection for vsc_warp_nchw_f32 (same results): 0 default (= 1), 1 = one pixel per thread over the
* flattened image (fastest on smooth flow), 2 = 32x8 pixel tiles with a shared 2x2 gather quad (faster on
 * scattered flow), 3 = the quad addressing on the flattened mapping; | (k << 4), k = 1..255: the linear kernel splits
 * the channels into k chunks (grid y) instead of choosing the split itself.  Process-wide; for tests and benchmarks. */
VSC_API int vsc_set_warp_mode(int mode);

/* ------------------------------------------------------------------ stabilization (HWC fp32, 3 channels)
 * flow_channels: 3 (model output u,v,0; videostabilizer.cpp:50-51) or 2 (.flo layout); only u,v are read. */
VSC_API int vsc_warp_hwc3(const float* in, const float* flow, float* out, int W, int H, int flow_channels,
    vsc_stream_t stream);

/* adapCmbIn may be NULL (it is only an input of vsc_consist_wt). */
VSC_API int vsc_adap_comb(const float* crntIn, const float* crntPr, const float* prevWarpIn, const float* prevWarpPr,
    const float* nextWarpIn, const float* nextWarpPr, float* adapCmbIn, float* adapCmbPr, const float* lastStabWarp,
    float alpha, int W, int H, vsc_stream_t stream);

VSC_API int vsc_consist_wt(const float* adapCmbIn, const float* crntIn, float* consWt, float beta, float gamma, int W,
    int H, vsc_stream_t stream);

/* resize without half-pixel offset; writes Co channels, reads with stride Ci (Co <= Ci) */
VSC_API int vsc_bilinear(const float* in, int Wi, int Hi, int Ci, float* out, int Wo, int Ho, int Co,
    vsc_stream_t stream);

/* get_consist_out: numIter sweeps of the screened-Poisson gradient descent on consisOut (in/out,
 * caller-initialised).  Deterministic Jacobi ordering (the reference's in-place update is a data
 * race, flowconsistency.cu:370).  workspace: device scratch of at least
 * vsc_consist_solve_workspace_bytes(W,H) bytes, contents undefined on entry and exit. */
VSC_API size_t vsc_consist_solve_workspace_bytes(int W, int H);
VSC_API int vsc_consist_solve(const float* crntPr, const float* prevStabWarp, const float* consWt, int numIter,
    float stepSize, float momFac, float* consisOut, int W, int H, void* workspace, size_t workspace_bytes,
    vsc_stream_t stream);

/* How the sweeps are executed (results are identical, bit for bit, in every mode):
 *   0  auto (default): temporally blocked passes (10 or 8 sweeps per launch plus one shorter even pass,
 *      intermediate sweeps kept on chip) for images of at least 128x48, single unblocked sweeps otherwise and
 *      for an odd remainder;
 *   1  unblocked sweeps only;   2  blocked passes whenever numIter >= 4, whatever the image size.
 * Two flag bits select variants of the blocked kernel (same results): | 0x10 = CTA-wide barrier instead of
 * neighbour-pair named barriers; | 0x20 = per-thread 4-byte staging instead of warp-cooperative 16-byte staging;
 * | 0x40 = vsc_frame_stabilize never takes its fused path; | 0x80 = no programmatic dependent launch
 * anywhere in the library; | (j << 12), j = 1 or 2: main blocked passes always of 8 / always of 10 sweeps
 * (default: 8; 10 on images of at least 0.9 Mpx only where that saves enough passes, e.g. 20 sweeps = 2 x 10); | (k << 8), k = 1..4, forces the band width of the
 * blocked kernel (512, 448, 384, 256 floats) instead of the cost model; | 0x8000 = the fully unrolled form of the
 * blocked kernel (stab_solver_stream.cu) instead of the 4-step loop (stab_solver_rolled.cu; needs 16-byte aligned
 * rows; by default it takes every pass except the 10-sweep passes of images of 4 Mpx and more), | 0x4000 = the 4-step
 * loop for those too; | 0x0800 = the sweeps are split into the fewest passes of nearly equal depth, odd depths 3..9
 * included (75 = 3 x 10 + 5 x 9; default: 8- or 10-sweep passes, a 2-sweep remainder merged into the last pass, an odd
 * sweep on its own: measured faster); | ((a + 1) << 16) | ((b + 1) << 22), a, b in 0..62: the first / last row chunk of the 4-step-loop kernel is
 * a / b rows shorter than the others (default 12 / 6); | (1 << 28) = the exchange ring of the 4-step-loop kernel in its scalar
 * layout (default: quad gather, [slot][column][level], 128-bit accesses); | (1 << 30) = the warps of one scheduler own adjacent
 * column blocks.
 * Process-wide; meant for tests and benchmarks. */
VSC_API int vsc_set_solver_mode(int mode);

/* RGBA8888 (device bytes) <-> float3.  to-float: float(u8)/255 (alpha ignored);
 * to-u8: (uint8)floor(|v|*255) without clamp (wraps mod 256), alpha byte = 1 (gpuimage.cu:39-67). */
VSC_API int vsc_rgba8_to_f32x3(const uint8_t* rgba_dev, float* out, int W, int H, vsc_stream_t stream);
VSC_API int vsc_f32x3_to_rgba8(const float* in, uint8_t* rgba_dev, int W, int H, vsc_stream_t stream);

/* Nearest-neighbour scaling of an RGBA8888 device image, sampled at pixel centres in 16.16 fixed point
 * (source x = (ix / 2 + x * ix) >> 16, ix = 65536 * srcW / dstW truncated; same in y): the device-side
 * replacement for the CPU QImage::scaled(..., Qt::FastTransformation) in FlowModel::run (flowmodel.cpp:126-131),
 * so that the flow network's input is produced from the frame already resident on the GPU. */
VSC_API int vsc_rgba8_scale_nearest(const uint8_t* src_dev, int srcW, int srcH, uint8_t* dst_dev, int dstW, int dstH,
    vsc_stream_t stream);

/* Fused "stage A" of doOneStep (videostabilizer.cpp:182-198): the five get_warp_result calls,
 * get_adap_comb and get_consist_wt in ONE pass -- no warped intermediates touch HBM.
 * Outputs adapCmbPr and consWt (and adapCmbIn if non-NULL). */
VSC_API int vsc_stage_a_fused(const float* origPrev, const float* origCur, const float* origNext,
    const float* procPrev, const float* procCur, const float* procNext, const float* lastStab, const float* flowFwd,
    const float* flowBwd, int flow_channels, float alpha, float beta, float gamma, float* adapCmbIn, float* adapCmbPr,
    float* consWt, int W, int H, vsc_stream_t stream);

/* Kernel selection for the fused stage A (vsc_stage_a_fused, vsc_frame_stabilize, the stream object; same results
 * bit for bit).  Low 4 bits: 0 default (= 3 with 128-thread CTAs), 1 = one row per CTA, 2 = a thread walks down a
 * chunk of rows, 3 = + the next row's flow prefetched, 4 = + the next row's loads issued before the current row's
 * arithmetic; bits 4-7 = log2(rows per CTA) or bits 12-19 = rows per CTA (neither: 8, fewer on small frames); | 0x100 =
 * 128-thread CTAs.
 * vsc_stage_a_fused itself only distinguishes 1 from the rest.  Process-wide; for tests and benchmarks. */
VSC_API int vsc_set_stage_a_mode(int mode);

/* hyperParams of the reference, same field order (videostabilizer.h:38-46) */
typedef struct vsc_hyper_params {
    float alpha;
    float beta;
    float gamma;
    int pyramidLevels;
    int numIter;
    float stepSize;
    float momFac;
} vsc_hyper_params;
/* defaults of VideoStabilizer::initHyperParams (videostabilizer.cpp:104-112) */
VSC_API void vsc_hyper_params_default(vsc_hyper_params* p);

/* The pyramid + solver part of doOneStep (videostabilizer.cpp:200-228): pyramid of
 * p->pyramidLevels levels (sizes halved with integer division), coarse-to-fine solve with
 * numIter/(j+1) sweeps at level j, result in consisOut (full resolution, need not be initialised).
 * workspace >= vsc_frame_solve_workspace_bytes(W,H,levels). */
VSC_API size_t vsc_frame_solve_workspace_bytes(int W, int H, int pyramidLevels);
VSC_API int vsc_frame_solve(const float* procCur, const float* adapCmbPr, const float* consWt,
    const vsc_hyper_params* p, float* consisOut, int W, int H, void* workspace, size_t workspace_bytes,
    vsc_stream_t stream);

/* Stage A + pyramid + solve in one call: everything doOneStep does between retrieveOpticalFlow and
 * copyToQImage (videostabilizer.cpp:182-228) on device images.  With two pyramid levels and even W, H the
 * adaptive combination, the consistency weight, the level-0 solver coefficients and the level-1 inputs come
 * out of ONE kernel (no adapCmbPr / consWt round trip through HBM, no separate down-scales); otherwise it is
 * vsc_stage_a_fused + vsc_frame_solve.  Results are bit-identical either way.
 * workspace >= vsc_frame_stabilize_workspace_bytes(W,H,levels). */
VSC_API size_t vsc_frame_stabilize_workspace_bytes(int W, int H, int pyramidLevels);
VSC_API int vsc_frame_stabilize(const float* origPrev, const float* origCur, const float* origNext,
    const float* procPrev, const float* procCur, const float* procNext, const float* lastStab, const float* flowFwd,
    const float* flowBwd, int flow_channels, const vsc_hyper_params* p, float* consisOut, int W, int H,
    void* workspace, size_t workspace_bytes, vsc_stream_t stream);

/* ------------------------------------------------------------------ per-stream pipeline object
 * Mirror of VideoStabilizer's recurrence (videostabilizer.cpp:136-153,167-265) for ONE video
 * stream on the current device: a sliding window [prev,cur,next] of original+processed frames,
 * the fp32 lastStabilizedFrame state, all persistent device buffers, a compute stream and a copy
 * stream with pinned staging (frames are uploaded asynchronously while the previous frame is
 * being solved).  Independent streams = independent objects, one per GPU for multi-GPU. */
typedef struct vsc_stabilizer vsc_stabilizer;

VSC_API int vsc_stabilizer_create(vsc_stabilizer** out, int W, int H, int flow_channels);
/* the same with flow batches (main.cpp:48-63 `-b batchSize`, videostabilizer.cpp:65-71,136-153,176,269-273): the
 * window holds 2k + batchSize frames (k = 1; batchSize 1..16), lastStabilizedFrame is initialised from the LAST
 * preloaded frame (the (2 + batchSize)-th push), a step consumes window[0..2] as before, and
 * vsc_stabilizer_flow_input accepts window indices up to 1 + batchSize, so that a flow session can fill its
 * [batchSize, H, W, 4] inputs with frames 1..batchSize / 2..batchSize+1 once every batchSize steps.
 * vsc_stabilizer_create == batchSize 1. */
VSC_API int vsc_stabilizer_create_batched(vsc_stabilizer** out, int W, int H, int flow_channels, int batchSize);
VSC_API int vsc_stabilizer_batch_size(const vsc_stabilizer* s);
/* frames currently in the window */
VSC_API int vsc_stabilizer_window_count(const vsc_stabilizer* s);
VSC_API void vsc_stabilizer_destroy(vsc_stabilizer* s);
/* live pointer, like VideoStabilizer::getHyperParams(); snapshotted once per step (:192) */
VSC_API vsc_hyper_params* vsc_stabilizer_hyper_params(vsc_stabilizer* s);
/* loadFrame(): append one (original, processed) RGBA8888 frame pair from HOST memory to the
 * window (upload + u8->f32 on the copy stream).  The third push after creation/reset also
 * initialises lastStabilizedFrame from that processed frame (preloadProcessedFrames :152).
 * Host buffers may be pageable (staged through internal pinned buffers) or pinned
 * (vsc_host_alloc / cudaHostRegister: copied directly).  A pageable buffer may be reused as soon as the
 * call returns.  A PINNED buffer is read by an asynchronous copy that may still be pending when this call --
 * and any following vsc_stabilizer_step -- has returned: it may be overwritten only after
 * vsc_stabilizer_sync (or vsc_stabilizer_wait_uploads, which waits for the uploads alone).  The same holds
 * for pinned flows of vsc_stabilizer_step_host_flow, and a pinned out_rgba_host is complete only after
 * vsc_stabilizer_sync. */
VSC_API int vsc_stabilizer_push_frame(vsc_stabilizer* s, const uint8_t* orig_rgba_host,
    const uint8_t* proc_rgba_host);
/* Stream contract of every *_dev argument of the vsc_stabilizer_* calls: the call enqueues work on the object's
 * own compute stream (vsc_stabilizer_compute_stream, created cudaStreamNonBlocking) and returns at once.  The
 * caller must (a) make that stream wait for whatever stream produced a *_dev input (cudaEventRecord on the
 * producer + cudaStreamWaitEvent on the compute stream) BEFORE the call, (b) keep the memory alive and unmodified
 * until the step has run (vsc_stabilizer_sync, or an event recorded on the compute stream after the call), and
 * (c) order any consumer of a *_dev output after the compute stream in the same way.
 *
 * doOneStep(): stabilise window[1] with flow cur->next (flowFwd) and the flow the reference uses
 * as cur->prev (flowBwd), both DEVICE HWC images of flow_channels channels at frame resolution.
 * Writes the RGBA8888 result to out_rgba_host (if non-NULL; valid after vsc_stabilizer_sync),
 * updates lastStabilizedFrame and pops the window front.  Requires at least 3 frames in the window. */
VSC_API int vsc_stabilizer_step(vsc_stabilizer* s, const float* flowFwd_dev, const float* flowBwd_dev,
    uint8_t* out_rgba_host);
/* same, flows given at a lower resolution (FLOWDOWNSCALE, flowmodel.cpp:156-165): upsampled with
 * vsc_bilinear semantics (values not rescaled, as the reference) inside the step */
VSC_API int vsc_stabilizer_step_lowres_flow(vsc_stabilizer* s, const float* flowFwd_dev, const float* flowBwd_dev,
    int flowW, int flowH, uint8_t* out_rgba_host);
/* same, flows given in HOST memory (precomputed-flow mode, BASELINE config 4: what FileStabilizer reads from
 * frame_%06d.flo / frame_%06d_bwd.flo with ReadFlowFile, stabilizefiles.cpp:134-149): HWC float, flow_channels
 * per pixel (2 for .flo), at flowW x flowH; uploaded on the copy stream (pinned buffers directly, pageable ones
 * through pinned staging), up-sampled like _step_lowres_flow if smaller than the frame. */
VSC_API int vsc_stabilizer_step_host_flow(vsc_stabilizer* s, const float* flowFwd_host, const float* flowBwd_host,
    int flowW, int flowH, uint8_t* out_rgba_host);
/* same, flows read from <flow_dir>/frame_%06d.flo (currentFrame + 1: previous -> current ... see
 * stabilizefiles.cpp:139-144) and <flow_dir>/frame_%06d_bwd.flo (currentFrame): what
 * FileStabilizer::retrieveOpticalFlow + doOneStep do for frame `currentFrame`.  The stabilizer must have been
 * created with flow_channels == 2; files of another size than the frame fail with VSC_E_FLO_DIMS. */
VSC_API int vsc_stabilizer_step_flow_files(vsc_stabilizer* s, const char* flow_dir, int currentFrame,
    uint8_t* out_rgba_host);
/* optional: start reading the .flo pair of `currentFrame` on a worker thread into the second set of pinned
 * landing buffers, so that the next vsc_stabilizer_step_flow_files(s, flow_dir, currentFrame, ...) finds it in
 * memory (file reading overlaps the previous frame's GPU work).  Read errors surface from that step call. */
VSC_API int vsc_stabilizer_prefetch_flow_files(vsc_stabilizer* s, const char* flow_dir, int currentFrame);
/* FlowModel::run input path without the host (flowmodel.cpp:121-150): writes window frame `window_index`
 * (0 = previous, 1 = current, 2 = next, ... up to 1 + batchSize; the ORIGINAL stream) as RGBA8888 at netW x netH into dst_dev -- e.g. the
 * flow session's bound input tensor -- scaling with vsc_rgba8_scale_nearest when the size differs.  Enqueued on
 * the compute stream; the window must hold that frame. */
VSC_API int vsc_stabilizer_flow_input(vsc_stabilizer* s, int window_index, uint8_t* dst_dev, int netW, int netH);
VSC_API int vsc_stabilizer_sync(vsc_stabilizer* s);
/* blocks until every host->device copy enqueued so far (frames of vsc_stabilizer_push_frame, flows of
 * vsc_stabilizer_step_host_flow) has left its source buffer: after it, pinned input buffers may be overwritten.
 * Does not wait for any kernel. */
VSC_API int vsc_stabilizer_wait_uploads(vsc_stabilizer* s);
/* device pointer to the fp32 result of the last step (W*H*3 floats), for tests */
VSC_API const float* vsc_stabilizer_last_output_dev(vsc_stabilizer* s);
/* enqueue a device-to-device copy of that result into dst_dev on the compute stream */
VSC_API int vsc_stabilizer_copy_last_output(vsc_stabilizer* s, float* dst_dev);
VSC_API vsc_stream_t vsc_stabilizer_compute_stream(vsc_stabilizer* s);
/* clears the window and the recurrence (seek) */
VSC_API int vsc_stabilizer_reset(vsc_stabilizer* s);

/* ---- .flo ingestion (host side; precomputed-flow mode) -------------------------------------------------
 * vsc_flo_read: ReadFlowFile (flowIO.cpp:31-78) into a caller buffer of dst_capacity_floats floats (use
 * vsc_host_alloc memory to make the following upload asynchronous); *width / *height receive the header
 * values.  A buffer smaller than width*height*2 fails with VSC_E_WORKSPACE after the header was read, so a
 * caller may size the buffer from vsc_flo_read_header.  Errors: VSC_E_FLO_*, one per reference exception.
 * vsc_flo_frame_path: "<dir>/frame_%06d.flo" or "<dir>/frame_%06d_bwd.flo" (stabilizefiles.cpp:99-101,141-144). */
VSC_API int vsc_flo_read_header(const char* path, int* width, int* height);
VSC_API int vsc_flo_read(const char* path, float* dst, size_t dst_capacity_floats, int* width, int* height);
VSC_API int vsc_flo_frame_path(const char* flow_dir, int frame, int backward, char* out, size_t out_capacity);

/* pinned host memory for frame buffers (cudaHostAlloc / cudaFreeHost) */
VSC_API int vsc_host_alloc(void** p, size_t bytes);
VSC_API int vsc_host_free(void* p);

#ifdef __cplusplus
}
#endif
#endif /* VSC_VSC_H */
