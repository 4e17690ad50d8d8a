/*
 * unmicst_b200 — C-ABI of the B200-native UnMicst probability-map engine.
 *
 * This is the drop-in boundary for the one hot path of HMS-IDAC/UnMicst:
 *   tile -> UNet2D forward -> ramp-weighted stitch -> uint8 probability maps.
 * Every entry point names the reference interface it replaces (file:line in the
 * reference tree).  Plain pointers and sizes only; no C++ or torch types; no
 * exceptions cross the boundary.  All functions return 0 (UMX_OK) or a negative
 * UMX_E* code; the message is available from umx_last_error() (thread-local).
 *
 * Ownership: the caller owns every pointer it passes (host pageable, host pinned
 * or device memory — the library asks the driver which).  The library owns all
 * device memory, streams and events it creates; umx_destroy releases them.
 * Threading: a handle is bound to one device and must be driven by one thread at
 * a time; different handles may run concurrently from different threads (ctypes
 * drops the GIL for the duration of a call).  There is no global mutable state.
 */
#ifndef UNMICST_B200_H
#define UNMICST_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define UMX_ABI_VERSION 1

/* status codes */
#define UMX_OK            0
#define UMX_EINVAL       -1   /* bad argument / unsupported model description   */
#define UMX_ENOTENSOR    -2   /* a tensor the graph needs is missing / misshaped */
#define UMX_ECUDA        -3   /* CUDA runtime or driver error                    */
#define UMX_ENOMEM       -4   /* host or device allocation failed                */
#define UMX_ENODEVICE    -5   /* no usable sm_100 device                         */

/* graph generations */
#define UMX_GRAPH_LEGACY  0   /* UnMicst.py:51-187                                */
#define UMX_GRAPH_V2      1   /* UnMicst1-5.py:55-237, UnMicst2.py, UnMicstCyto2.py */

/* image sample types accepted by umx_infer_image */
#define UMX_U8   0
#define UMX_U16  1
#define UMX_F32  2
#define UMX_F64  3

/* arithmetic of the wide convolutions (umx_opts.precision / umx_model_desc.precision) */
#define UMX_PREC_DEFAULT  0   /* library default (currently UMX_PREC_SPLIT3 where the tensor path applies) */
#define UMX_PREC_FP32     1   /* every layer in fp32 FMA on the CUDA cores (exact reference arithmetic)     */
#define UMX_PREC_SPLIT3   2   /* tcgen05 fp16 hi/lo split, 3 MMAs per product, fp32 accumulate (~fp32)      */
#define UMX_PREC_SINGLE   3   /* tcgen05 fp16 operands, 1 MMA per product, fp32 accumulate                  */
#define UMX_PREC_MIXED    4   /* per layer: umx_model_desc.reserved[0] (low) / [1] (high) is a bit mask over the   */
                              /* ops in plan order (= entry order of umx_profile_read); set bits run with 1 MMA    */
                              /* per product, the other tensor-path layers keep the hi/lo split                     */

typedef struct umx_handle umx_handle;

/* hp.data of a model folder (UnMicst1-5.py:57-67) plus the graph generation. */
typedef struct umx_model_desc {
    int32_t abi_version;     /* = UMX_ABI_VERSION */
    int32_t graph;           /* UMX_GRAPH_*       */
    int32_t im_size;         /* hp['imSize']   S  */
    int32_t n_channels;      /* hp['nChannels'] C */
    int32_t n_classes;       /* hp['nClasses'] K  */
    int32_t n_out0;          /* hp['nOut0']       */
    int32_t n_layers;        /* hp['nLayers']     */
    int32_t feat_maps_fact;  /* hp['featMapsFact']*/
    int32_t down_samp_fact;  /* hp['downSampFact'] (must be 2) */
    int32_t ks;              /* hp['ks'] in {1,3,5} */
    int32_t n_extra_convs;   /* hp['nExtraConvs'] */
    int32_t precision;       /* UMX_PREC_*        */
    int32_t max_batch_tiles; /* tiles per forward launch group; 0 = library default */
    int32_t reserved[3];
} umx_model_desc;

/* One named fp32 variable of the checkpoint in TensorFlow layout
 * (conv kernels HWIO, conv-transpose kernels HW[out][in], vectors [C]). */
typedef struct umx_tensor {
    const char*  name;
    const float* data;       /* host pointer, row-major */
    int32_t      ndim;
    int64_t      shape[4];
} umx_tensor;

/* Host-side pre-map applied per sample before (x-mean)/std, in float64, in the
 * reference's operation order (UnMicst1-5.py:813-821, UnMicst.py:627-631):
 *   x = sample * in_scale                       (img_as_float: 1/65535, 1/255 or 1)
 *   if rescale: x = clip(x, imin, imax); x = (x-imin)/(imax-imin) * (omax-omin) + omin   */
typedef struct umx_premap {
    double  in_scale;
    int32_t rescale;
    int32_t pad_;
    double  imin, imax, omin, omax;
} umx_premap;

typedef struct umx_opts {
    int32_t tile_row0;       /* band of PI2D tile rows [tile_row0, tile_row1) this call owns; */
    int32_t tile_row1;       /* tile_row1 <= 0 means "to the last tile row"                    */
    int32_t precision;       /* 0, or the handle's own UMX_PREC_* (a different value is refused with UMX_EINVAL: the    */
                             /* arithmetic is fixed when the handle is built; make a second handle for a check pass)    */
    int32_t flags;           /* UMX_F_*                                                        */
    const umx_premap* premap;/* NULL = samples are already the float image the network sees (one map, or one per plane) */
    int64_t out_plane_stride;/* elements between class planes of out_u8/out_f32; 0 = H*W       */
    int32_t out_row_base;    /* image row that out_* row 0 corresponds to (band-local buffers) */
    int32_t infer_h;         /* --scalingFactor: run the network on skimage.transform.resize(img, (infer_h, infer_w))  */
    int32_t infer_w;         /* (UnMicst1-5.py:813-815), resampled on the fly from the H x W samples; 0 = H / W        */
    int32_t reserved[3];
} umx_opts;

#define UMX_F_NO_SYNC    1   /* return without waiting for the device (outputs must be device/pinned) */
#define UMX_F_PREMAP_PER_PLANE 4  /* umx_opts.premap points to one umx_premap per network input channel (unmicst-duo stretches each channel  */
                                  /* with its own min/max, UnMicst2.py:760-788) instead of one shared map                          */
#define UMX_F_CONTINUE   8   /* this call continues the image of the previous call on this handle (same image and options,   */
                             /* tile_row0 == the previous tile_row1): the tile row above the seam is still on the device and  */
                             /* is reused instead of recomputed — how a caller streams a slide band by band (e.g. into a     */
                             /* BigTIFF writer, UnMicst1-5.py:852-862) at no extra cost                                      */
#define UMX_F_STITCH_REPLACE 16 /* PI2D mode 'replace' (PartitionOfImage.py:99-100): in overlaps the tile patched last (row-major order) */
                                /* wins and no ramp weights are applied; default is 'accumulate' (:95-98, what the CLI uses)           */
#define UMX_F_FP16_QUANT 32  /* first quantisation exactly as the reference evaluates it: PI2D returns float16, so               */
                             /* np.uint8(255 * PM) (UnMicst1-5.py:848) is floor(fp16(255 * fp16(p))): p >= 0.99976 gives 255 (the  */
                             /* default, floor(255 * p) in fp32, gives 254 there); the fp16 ACCUMULATION of PI2D is not emulated  */
#define UMX_F_CLI_QUANT  2   /* out_u8 = the page the reference CLI writes (UnMicst1-5.py:848-853): uint8(255*p), resize back */
                             /* to the H x W grid of img when infer_h/infer_w differ, then uint8(255*x) a second time;       */
                             /* out_u8 is then [K][H][W] (rows of umx_band_out_rows for a band); out_f32 must be NULL        */

/* Per-kernel timing gathered with CUDA events on the handle's stream. */
typedef struct umx_prof_entry {
    char     name[48];
    int64_t  launches;
    double   ms_total;       /* sum of event-timed durations */
    double   flops;          /* algorithmic FLOPs over those launches   */
    double   bytes;          /* algorithmic HBM bytes over those launches */
} umx_prof_entry;

/* Number of CUDA devices visible; 0 when there is none (never an error).
 * Replaces toolbox/GPUselect.py:4-22 together with umx_device_free_mem. */
int umx_device_count(void);
int umx_device_free_mem(int device, int64_t* free_bytes, int64_t* total_bytes);

/* Build the graph and load weights: UNet2D.singleImageInferenceSetup
 * (UnMicst1-5.py:656-681: setupWithHP + Saver.restore).  Folds batch-norm,
 * merges the v2 shortcut, repacks for the kernels, allocates workspaces. */
int umx_create(const umx_model_desc* desc, const umx_tensor* weights, int32_t n_weights,
               int32_t device, umx_handle** out);

/* umx_create with the arithmetic chosen per op and per concat source: op_terms[i] (i = op index in plan order, the entry
 * order of umx_profile_read / umx_describe_plan) = t0 | t1 << 2, where t0 / t1 say which hi/lo correction terms the op
 * adds to a_hi*w_hi for its first / second source: bit 0 = a_hi*w_lo (weight rounding), bit 1 = a_lo*w_hi (activation
 * rounding); 15 = the full split, 0 = one MMA per product, -1 = whatever desc->precision implies.  Buffers and weights
 * get a lo plane only where a term reads it.  This is what the Python layer's `auto` calibration produces. */
int umx_create_ex(const umx_model_desc* desc, const umx_tensor* weights, int32_t n_weights, int32_t device,
                  const int32_t* op_terms, int32_t n_op_terms, umx_handle** out);

/* Calibration aid: on a handle built with UMX_PREC_SPLIT3 (every lo plane exists) switch the correction terms of one op
 * (same encoding) for the following calls; the results equal those of a handle built with umx_create_ex for these terms. */
int umx_set_op_terms(umx_handle* h, int32_t op_index, int32_t terms);

/* UNet2D.singleImageInferenceCleanup (UnMicst1-5.py:684-685). */
void umx_destroy(umx_handle* h);

/* Session.run(UNet2D.nn, {tfData: tiles, tfTraining: 0}) (UnMicst1-5.py:704):
 * tiles [n,S,S,C] fp32 NHWC -> probs [n,S,S,K] fp32 softmax.  Any n >= 0.
 * `precision`: 0 or the handle's own UMX_PREC_* (anything else: UMX_EINVAL, see umx_opts.precision). */
int umx_forward_tiles(umx_handle* h, const float* tiles_nhwc, int32_t n_tiles, float* probs_nhwc,
                      int32_t precision);

/* UNet2D.singleImageInference for every class at once (UnMicst1-5.py:687-710 with
 * PI2D.setup/getPatch/createOutput/patchOutput/getValidOutput,
 * toolbox/PartitionOfImage.py:23-122) plus np.uint8(255*p) (UnMicst1-5.py:848):
 * img [C][H][W] samples of `dtype` (plane stride in elements; 0 = H*W) ->
 * out_u8 [K][rows][W] = floor(255*p) and/or out_f32 [K][rows][W] = p, for the
 * image rows covered by the tile-row band in `opts` (all rows by default). */
int umx_infer_image(umx_handle* h, const void* img, int32_t dtype, int32_t n_planes, int32_t H, int32_t W,
                    int64_t plane_stride, double mean, double std_dev,
                    uint8_t* out_u8, float* out_f32, const umx_opts* opts);

/* One image of a umx_infer_images batch: the arguments of umx_infer_image, per image. */
typedef struct umx_image {
    const void* img;         /* [n_planes][H][W] samples */
    int32_t dtype, n_planes, H, W;
    int64_t plane_stride;    /* elements; 0 = H*W */
    const umx_premap* premap;/* NULL or ONE map for this image */
    uint8_t* out_u8;         /* [K][H][W] floor(255 p) (twice quantised with UMX_F_CLI_QUANT) or NULL */
    float*   out_f32;        /* [K][H][W] or NULL */
} umx_image;

/* Many small images through one resident model — the TMA "dearray" loop of batchUNet2DTMACycif.py:539-569, where the
 * reference runs singleImageInference core by core.  Tiles of as many images as fit a launch group share every network
 * launch, so small cores fill the GPU; results equal umx_infer_image per image bit for bit.  flags: UMX_F_CLI_QUANT. */
int umx_infer_images(umx_handle* h, const umx_image* images, int32_t n_images, double mean, double std_dev, int32_t flags);

/* Image rows [row0,row1) produced by a tile-row band (host-side helper for sharding). */
int umx_band_rows(umx_handle* h, int32_t H, int32_t tile_row0, int32_t tile_row1, int32_t* row0, int32_t* row1);

/* Same for UMX_F_CLI_QUANT output when the network runs at infer_h rows and the pages are resized back to raw_h rows:
 * the raw-grid rows [row0,row1) whose resize window lies inside what the band computes (bands stay disjoint and
 * complete, and need no exchange).  With raw_h == infer_h this equals umx_band_rows. */
int umx_band_out_rows(umx_handle* h, int32_t infer_h, int32_t tile_row0, int32_t tile_row1, int32_t raw_h,
                      int32_t* row0, int32_t* row1);

/* min and max of skimage.transform.resize(img_as_float(plane), (out_h, out_w)) — rescale_intensity's in_range for the
 * tools that stretch the resized image (UnMicst.py:627-631) — computed on the device; plane: H x W samples, host or
 * device memory.  in_scale = img_as_float's factor (1/65535, 1/255 or 1). */
int umx_resample_minmax(umx_handle* h, const void* plane, int32_t dtype, int32_t H, int32_t W, int32_t out_h, int32_t out_w,
                        double in_scale, double* min_out, double* max_out);

/* Use the caller's CUDA stream (cudaStream_t as an integer) for all compute; 0 = library stream. */
int umx_set_stream(umx_handle* h, uint64_t cuda_stream);

/* Per-kernel event timing: enable, run, then read back (returns number of entries). */
int umx_profile_enable(umx_handle* h, int32_t on);
int umx_profile_read(umx_handle* h, umx_prof_entry* out, int32_t capacity, int32_t reset);

/* Debug/bring-up: copy the activation buffer written by op `name` ("ld1.conv0", "lu3.convT", ...)
 * during the last forward, first n_tiles tiles, to host fp32 NHWC (hi+lo planes are summed).
 * Returns the element count per tile, or a negative UMX_E* code. */
int64_t umx_debug_buffer(umx_handle* h, const char* name, int32_t n_tiles, float* out, int64_t capacity);

/* Kernels launched by this handle since creation (umx_*: claim for "gpu_launches"). */
int64_t umx_launch_count(umx_handle* h);

/* How op `op_index` was lowered (diagnostics, calibration): out[0..11] = on the tensor path, halo mode, CTA pairs, patch /
 * ring slots, weight slots, taps per weight slot, weights resident in shared memory (2 = resident with the N-concatenated
 * a_hi x [w_hi | w_lo] MMA), conv-transpose px merge, planes per
 * activation slot, planes per weight slot, correction terms of source 0, of source 1. */
int umx_op_info(umx_handle* h, int32_t op_index, int32_t* out, int32_t capacity);

/* Pinned host memory for callers that want full-rate async copies. */
void* umx_host_alloc(int64_t bytes);
void  umx_host_free(void* p);

/* Host-only: text description (one line per op) of the plan umx_create would build for this model: graph
 * construction as in UNet2D.setup (UnMicst1-5.py:55-237, UnMicst.py:51-187), BN folding, the raw-input rewrite and
 * the kernel family of every op.  Needs no GPU.  Returns characters written or a negative UMX_E* code. */
int64_t umx_describe_plan(const umx_model_desc* desc, const umx_tensor* weights, int32_t n_weights, char* out, int64_t capacity);

/* Host-only: TIFF 6.0 LZW decoder for the channel-page reader (the codec tifffile / imagecodecs give the reference's
 * skio.imread(img_num=...) / tifffile.imread(key=...) calls, UnMicst1-5.py:794-797).  Returns the bytes written
 * (at most dst_capacity), or a negative UMX_E* code for a corrupt stream.  Needs no GPU. */
int64_t umx_tiff_lzw_decode(const uint8_t* src, int64_t src_bytes, uint8_t* dst, int64_t dst_capacity);

const char* umx_last_error(void);
const char* umx_version(void);

#ifdef __cplusplus
}
#endif
#endif /* UNMICST_B200_H */
