/*
 * micgpu.h -- C ABI of libmicgpu.so, the B200 (sm_100a) codec path for MIC
 * (pappuks/medical-image-codec).  Plain pointers and sizes only; this is what a
 * cgo / ctypes binding links against.
 *
 * Conventions (same as the reference's C twin, ojph/mic_decompress_c.h:24-49,
 * ojph/mic_parallel.h:49-55): 0 = success, negative = error class; the caller
 * owns every buffer; all host<->device copies complete inside the call, so Go
 * may pass pointers into Go-managed memory; calls are thread-safe (one
 * internal lock per device context).
 *
 * There is NO CPU fallback: every entry point fails with MICGPU_E_CUDA when no
 * CUDA device is usable.
 */
#ifndef MICGPU_H
#define MICGPU_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- error classes -------------------------------------------------------- */
#define MICGPU_OK 0
#define MICGPU_E_HEADER (-1)      /* bad magic / args / truncated container (C twin -1) */
#define MICGPU_E_NCOUNT (-2)      /* corrupt ncount header (C twin -2) */
#define MICGPU_E_ALLOC (-3)
#define MICGPU_E_DTABLE (-4)      /* corrupt decode table (C twin -4) */
#define MICGPU_E_BITSTREAM (-6)   /* bit reader over-read (C twin -6) */
#define MICGPU_E_RLE (-8)         /* RLE stream malformed / too short */
#define MICGPU_E_SIZE (-9)        /* output capacity / geometry mismatch */
#define MICGPU_E_UNSUPPORTED (-10)
#define MICGPU_E_INCOMPRESSIBLE (-11) /* ErrIncompressible (fseu16.go:33) */
#define MICGPU_E_USE_RLE (-12)        /* ErrUseRLE (fseu16.go:36) */
#define MICGPU_E_INTERNAL (-13)       /* FSE normalisation / table construction failed */
#define MICGPU_E_CUDA (-20)       /* CUDA runtime failure or no device */

/* unit kinds for micgpu_decoder_add_unit */
#define MICGPU_KIND_SPATIAL 0     /* DecompressSingleFrame (multiframecompress.go:97) */
#define MICGPU_KIND_RLE 1         /* decompressResidualFrame (multiframecompress.go:165) */

/* ---- library ---------------------------------------------------------------- */
int micgpu_device_count(void);
/* Text of the last error raised on the calling thread. */
const char *micgpu_last_error(void);
/* Pinned host memory helpers (optional; pageable buffers work, slower). */
void *micgpu_host_alloc(size_t bytes);
void micgpu_host_free(void *p);
/* Devices the one-call batch entry points spread their units over: the list replaces the worker
 * pools of the reference (goroutines in parallelstrips.go:292-321, `Workers` in wsicompress.go:112-115, pthreads in
 * ojph/mic_parallel.c:131-189).  devices == NULL or n <= 0 selects every visible device.  Returns the number of devices
 * in use (> 0) or an error.  Without this call everything runs on the calling thread's current device.  Affects
 * micgpu_pics_decompress_batch, micgpu_mic2_decompress and micgpu_wsi_decompress_tile_range: each device gets a
 * contiguous unit range balanced by compressed bytes, its own host thread, stream and scratch; nothing is exchanged
 * between devices except the carry frames of a temporal MIC2 stack (peer reads). */
int micgpu_init(const int *devices, int n);
/* The partition those calls use, exported for callers that shard by process instead (one rank per GPU): `parts`
 * contiguous ranges of n units balanced by sizes[] (compressed bytes from the container's own offset table:
 * parallelstrips.go:115-122, multiframe.go:72-78, wsiformat.go:145-155); cuts[] receives parts + 1 boundaries. */
int micgpu_partition_by_bytes(const uint64_t *sizes, uint64_t n, int parts, uint64_t *cuts);
/* Release the per-device default contexts used by the one-shot calls (and forget the device list). */
void micgpu_shutdown(void);

/* ---- batch decoder: plan once from HOST copies of the streams, run many ---- */
typedef struct micgpu_decoder micgpu_decoder;

micgpu_decoder *micgpu_decoder_create(int device);
void micgpu_decoder_destroy(micgpu_decoder *d);
/* Forget the current plan. */
int micgpu_decoder_begin(micgpu_decoder *d);
/* Add one FSE frame.  `frame` is a host pointer (only its first bytes are read
 * for planning); comp_off is where the frame will sit inside the compressed
 * device buffer passed to micgpu_decoder_run_device, out_off the first output
 * element (uint16) it decodes to.  SPATIAL: width x height pixels.
 * RLE: width*height is the output capacity in elements.  Returns the unit index. */
int micgpu_decoder_add_unit(micgpu_decoder *d, const uint8_t *frame, size_t frame_len, uint64_t comp_off, int kind,
                            uint32_t width, uint32_t height, uint64_t out_off);
/* Add every strip of a PICS container (parallelstrips.go:270-330).  The whole
 * blob is assumed to be copied at comp_off; pixels land at out_off. */
int micgpu_decoder_add_pics(micgpu_decoder *d, const uint8_t *pics, size_t len, uint64_t comp_off, uint64_t out_off,
                            int *width, int *height);
/* Add every strip of a PICA container (parallelstripsadaptive.go:143-212): content-adaptive strip rows, per-strip
 * predictor flag (bit 0 = gradient-adaptive, deltagradrlecompressu16.go:70-133). */
int micgpu_decoder_add_pica(micgpu_decoder *d, const uint8_t *pica, size_t len, uint64_t comp_off, uint64_t out_off,
                            int *width, int *height);
/* Add every frame of an independent-mode MIC2 container (multiframecompress.go:227-262). */
int micgpu_decoder_add_mic2(micgpu_decoder *d, const uint8_t *mic2, size_t len, uint64_t comp_off, uint64_t out_off,
                            int *width, int *height, int *frames, int *temporal);
/* Add frames [first_frame, first_frame + frame_count) of a MIC2 container: the shard of one rank when a stack is spread
 * over GPUs (frame offsets come from the container's own table, multiframe.go:72-78).  Frame first_frame
 * lands at out_off.  Independent mode: nothing else to do.  Temporal mode (multiframecompress.go:236-258): a range that
 * starts at frame 0 decodes to pixels; a later range decodes to running sums relative to a ZERO carry, and the caller
 * finishes it with micgpu_temporal_add_carry once it has the absolute last frame of the previous range -- the only
 * exchange step of the whole path, and it is associative: carry(r) = sum of the last frames of ranges < r (mod 2^16). */
int micgpu_decoder_add_mic2_range(micgpu_decoder *d, const uint8_t *mic2, size_t len, uint64_t comp_off, uint64_t out_off,
                                  int first_frame, int frame_count, int *width, int *height, int *frames, int *temporal);
/* d_frames[f][i] += d_carry[i] (mod 2^16) for nframes frames of frame_px pixels, both in device memory. */
int micgpu_temporal_add_carry(void *d_frames, const void *d_carry, uint64_t frame_px, int nframes, void *cuda_stream);
/* The same step fused with the exchange: d_peer_last[q] (q < npeers <= 16) is the last frame of earlier range q, readable
 * from this GPU -- local memory, or the memory of a peer GPU opened with micgpu_ipc_open; the kernel reads them over
 * NVLink, forms the carry and finishes the local frames in one pass.  The caller makes sure the peers have finished
 * writing those frames (a host barrier after their decode) and keeps them alive until this stream has run. */
int micgpu_temporal_add_carry_peers(void *d_frames, const void *const *d_peer_last, int npeers, uint64_t frame_px, int nframes,
                                    void *cuda_stream);
/* Plain device buffers (cudaMalloc: whole allocations, which is what CUDA IPC can export) and their IPC handles
 * (64 opaque bytes) for the exchange above between one-process-per-GPU ranks of one node. */
void *micgpu_device_alloc(size_t bytes);
void micgpu_device_free(void *p);
int micgpu_ipc_export(const void *d_ptr, void *handle64);
int micgpu_ipc_open(const void *handle64, void **d_ptr);
int micgpu_ipc_close(void *d_ptr);
/* Size scratch for the plan.  Must be called after the last add_*. */
int micgpu_decoder_commit(micgpu_decoder *d);
int micgpu_decoder_unit_count(const micgpu_decoder *d);
/* Decode the planned batch.  d_comp: device buffer holding the streams
 * (64-byte aligned, readable for comp_bytes + 256 bytes); d_out: device buffer
 * of out_elems uint16.  cuda_stream: a cudaStream_t (NULL = default stream).
 * Asynchronous with respect to the host. */
int micgpu_decoder_run_device(micgpu_decoder *d, const void *d_comp, size_t comp_bytes, void *d_out, size_t out_elems,
                              void *cuda_stream);
/* Wait for the last run and fetch per-unit status (0 = ok, else MICGPU_E_*).
 * Returns the first non-zero status, or 0. */
int micgpu_decoder_unit_status(micgpu_decoder *d, int *status, int n, void *cuda_stream);
/* Kernels launched by the last run_device call. */
int micgpu_decoder_last_launches(const micgpu_decoder *d);
/* Per-kernel timing with CUDA events on the launch stream (off by default).  After a run,
 * micgpu_decoder_kernel_times waits for it and returns the number of kernels, their names
 * (';'-separated) and durations in milliseconds. */
int micgpu_decoder_set_profiling(micgpu_decoder *d, int on);
int micgpu_decoder_kernel_times(micgpu_decoder *d, char *names, size_t names_cap, float *ms, int cap);
/* Convenience: copy `comp` (host) to the device, run, copy `out_elems` uint16 back. */
int micgpu_decoder_run_host(micgpu_decoder *d, const uint8_t *comp, size_t comp_bytes, uint16_t *out, size_t out_elems);

/* ---- container-level one-shot calls (host buffers in, host buffers out) ----- */
/* DecompressParallelStrips (parallelstrips.go:270).  pixels_out holds cap_px
 * uint16; *width/*height are set from the header. */
int micgpu_pics_decompress(const uint8_t *pics, size_t len, uint16_t *pixels_out, size_t cap_px, int *width, int *height);
/* n independent PICS images in one launch sequence; outs[i] holds caps[i] uint16.
 * status[i] (optional) receives the per-image result. */
int micgpu_pics_decompress_batch(int n, const uint8_t *const *blobs, const size_t *lens, uint16_t *const *outs,
                                 const size_t *caps, int *status);
/* DecompressSingleFrame (multiframecompress.go:97): any of the 1/2/4/8-state or rANS-8 streams. */
int micgpu_decompress_single_frame(const uint8_t *frame, size_t len, uint16_t *pixels_out, int width, int height);
/* DecompressSingleFrameGrad (multiframecompress.go:129-142): the gradient-adaptive predictor of
 * deltagradrlecompressu16.go:70-133 / gradPredict (deltagradcompressu16.go:149-166). */
int micgpu_decompress_single_frame_grad(const uint8_t *frame, size_t len, uint16_t *pixels_out, int width, int height);
/* DecompressParallelStripsAdaptive (parallelstripsadaptive.go:143). */
int micgpu_pica_decompress(const uint8_t *pica, size_t len, uint16_t *pixels_out, size_t cap_px, int *width, int *height);
/* DecompressMultiFrame / DecompressFrame (multiframecompress.go:227,266). */
int micgpu_mic2_decompress(const uint8_t *mic2, size_t len, uint16_t *frames_out, size_t cap_px, int *width, int *height,
                           int *frames, int *temporal);
int micgpu_mic2_decompress_frame(const uint8_t *mic2, size_t len, int frame_idx, uint16_t *pixels_out, size_t cap_px,
                                 int *width, int *height);

/* ---- MIC3 whole-slide containers and MICR RGB payloads ------------------------ */
#define MICGPU_WSI_MAX_LEVELS 32
typedef struct micgpu_wsi_info {   /* WSIHeader + WSILevel (wsiformat.go:39-66) */
  int width, height, tile_w, tile_h, channels, bits_per_sample, color_transform, n_levels;
  uint64_t total_tiles;
  int level_w[MICGPU_WSI_MAX_LEVELS], level_h[MICGPU_WSI_MAX_LEVELS];
  int tiles_x[MICGPU_WSI_MAX_LEVELS], tiles_y[MICGPU_WSI_MAX_LEVELS], first_tile[MICGPU_WSI_MAX_LEVELS];
} micgpu_wsi_info;
/* ReadWSIHeader (wsicompress.go:299-305) */
int micgpu_wsi_read_header(const uint8_t *mic3, size_t len, micgpu_wsi_info *info);
/* DecompressWSITile (wsicompress.go:175-217): pixels of one tile, edge tiles cropped; RGB8 interleaved,
 * grey8 or grey16 little-endian bytes. */
int micgpu_wsi_decompress_tile(const uint8_t *mic3, size_t len, int level, int tile_x, int tile_y, uint8_t *out, size_t cap,
                               int *width, int *height);
/* n tiles of one container in one launch sequence (every plane of every tile is one unit). */
int micgpu_wsi_decompress_tiles(const uint8_t *mic3, size_t len, int n, const int *levels, const int *tile_xs, const int *tile_ys,
                                uint8_t *const *outs, const size_t *caps, int *widths, int *heights, int *status);
/* DecompressWSIRegion (wsicompress.go:220-296) */
int micgpu_wsi_decompress_region(const uint8_t *mic3, size_t len, int level, int x, int y, int w, int h, uint8_t *out, size_t cap,
                                 int *out_w, int *out_h);
/* ---- region / viewport serving from a resident slide (DecompressWSIRegion wsicompress.go:220-296) -------
 * micgpu_wsi_open validates the header, the level descriptors and the tile table once and keeps the tile data area in
 * device memory; every later call parses nothing but the blob headers of the tiles it touches.  A batch of rectangles
 * -- any mix of pyramid levels -- is ONE launch sequence: the union of the touched tiles is decoded once, cropped and
 * composed on the device, and only the requested pixels are copied out.  Rectangles are clamped to their level exactly
 * like DecompressWSIRegion; out_ws / out_hs receive the clamped sizes.  The caller's mic3 buffer may be released after
 * micgpu_wsi_open returns.  Calls on one handle are serialised; use one handle per device for multi-GPU serving. */
typedef struct micgpu_wsi_slide micgpu_wsi_slide;
micgpu_wsi_slide *micgpu_wsi_open(int device, const uint8_t *mic3, size_t len);
void micgpu_wsi_close(micgpu_wsi_slide *s);
int micgpu_wsi_slide_info(const micgpu_wsi_slide *s, micgpu_wsi_info *info);
int micgpu_wsi_slide_region(micgpu_wsi_slide *s, int level, int x, int y, int w, int h, uint8_t *out, size_t cap, int *out_w, int *out_h);
int micgpu_wsi_slide_regions(micgpu_wsi_slide *s, int n, const int *levels, const int *xs, const int *ys, const int *ws, const int *hs,
                             uint8_t *const *outs, const size_t *caps, int *out_ws, int *out_hs, int *status);
/* the same with DEVICE output pointers (pixels stay on the GPU for a renderer) */
int micgpu_wsi_slide_regions_device(micgpu_wsi_slide *s, int n, const int *levels, const int *xs, const int *ys, const int *ws,
                                    const int *hs, void *const *d_outs, const size_t *caps, int *out_ws, int *out_hs, int *status);
int micgpu_wsi_slide_stats(const micgpu_wsi_slide *s, uint64_t *requests, uint64_t *tiles_decoded, int *last_launches);
/* ---- tile ranges: plan once, run many (viewer / batch conversion path) ------------------------------------------
 * Tiles [first_tile, first_tile + n_tiles) of the container's tile table (level l starts at first_tile[l], row-major:
 * wsiformat.go:145-155), each decoded as a FULL tile_w x tile_h block of pixels, tile_bytes apart (edge tiles keep the
 * zero padding the encoder added, wsicompress.go:529-556; crop with the level extent from micgpu_wsi_read_header).
 * The plan parses the container once (every compressed plane = one unit; constant planes cost nothing; raw planes are
 * copies); a run needs the bytes [span_off, span_off + span_len) of the container in device memory and writes
 * out_bytes of pixels -- no host work per tile. */
typedef struct micgpu_wsi_plan micgpu_wsi_plan;
micgpu_wsi_plan *micgpu_wsi_plan_tiles(int device, const uint8_t *mic3, size_t len, uint64_t first_tile, uint64_t n_tiles);
void micgpu_wsi_plan_destroy(micgpu_wsi_plan *p);
int micgpu_wsi_plan_info(const micgpu_wsi_plan *p, uint64_t *span_off, uint64_t *span_len, uint64_t *tile_bytes, uint64_t *out_bytes,
                         int *n_units);
/* d_span: device copy of mic3[span_off .. span_off+span_len) (+256 readable bytes); d_out: out_bytes. Asynchronous. */
int micgpu_wsi_plan_run_device(micgpu_wsi_plan *p, const void *d_span, void *d_out, void *cuda_stream);
/* Waits for the last run; tile_status[i] (optional, n entries) = 0 or the error of tile first_tile + i. */
int micgpu_wsi_plan_status(micgpu_wsi_plan *p, int *tile_status, int n, void *cuda_stream);
int micgpu_wsi_plan_launches(const micgpu_wsi_plan *p);
/* on >= 0: switch per-kernel CUDA-event timing of the plan on/off; on < 0: fetch names (';'-separated) and durations
 * of the last run, as micgpu_decoder_kernel_times does. */
int micgpu_wsi_plan_kernel_times(micgpu_wsi_plan *p, int on, char *names, size_t names_cap, float *ms, int cap);
/* Host buffers in and out: one H2D copy of the span, the kernels, one D2H copy per device (micgpu_init). */
int micgpu_wsi_decompress_tile_range(const uint8_t *mic3, size_t len, uint64_t first_tile, uint64_t n_tiles, uint8_t *out, size_t cap,
                                     int *status);
/* DecompressRGB (rgbcompress.go:31-33) */
int micgpu_rgb_decompress(const uint8_t *blob, size_t len, int width, int height, uint8_t *rgb_out);

/* ---- WaveletV2 streams ---------------------------------------------------------- */
/* WaveletV2[SIMD]RLEFSEDecompressU16 (waveletfsecompressu16.go:374,493): rows/cols come from the 11-byte header. */
int micgpu_wavelet_v2_decompress(const uint8_t *blob, size_t len, uint16_t *pixels_out, size_t cap_px, int *rows, int *cols);
/* The V1 wavelet layouts: WaveletFSEDecompressU16 (with_rle = 0, waveletfsecompressu16.go:124-163) and
 * WaveletRLEFSEDecompressU16 (with_rle = 1, :624-669): coefficients in raster order, interleaved in-place lifting
 * (waveletInverse2DRegion :180-189).  (The reference replaced these layouts by WaveletV2.) */
int micgpu_wavelet_v1_decompress(const uint8_t *blob, size_t len, int with_rle, uint16_t *pixels_out, size_t cap_px, int *rows, int *cols);
/* The canonical-Huffman back end of the legacy streams.  A Huffman stream has no magic byte, so these are
 * explicit entry points.  micgpu_huff_decompress = CanHuffmanDecompressU16.Init + ReadTable + Decompress
 * (canhuffmandecompressu16.go:31-108): *n_out receives the symbol count of the header (also when cap is too small:
 * MICGPU_E_SIZE).  micgpu_delta_rle_huff_decompress = DeltaRleHuffDecompressU16.Decompress
 * (deltarlehuffdecompressu16.go:19-39): Huffman -> RLE -> inverse avg(top,left) predictor.
 * micgpu_decoder_add_huff_unit queues such a stream in a plan beside FSE units (kind as for micgpu_decoder_add_unit:
 * SPATIAL = Delta+RLE symbols of a width x height image, RLE = an RLE stream of up to `width` words).
 * Limits: maxCodeLength <= 16 (the reference's encoder stops at 14 before it adds the delimiter); a stream that asks for
 * bits past its end returns MICGPU_E_BITSTREAM (Go reads stale window bits or panics on the slice). */
int micgpu_huff_decompress(const uint8_t *stream, size_t len, uint16_t *symbols_out, size_t cap, size_t *n_out);
int micgpu_delta_rle_huff_decompress(const uint8_t *stream, size_t len, uint16_t *pixels_out, int width, int height);
int micgpu_decoder_add_huff_unit(micgpu_decoder *d, const uint8_t *stream, size_t len, uint64_t comp_off, int kind,
                                 uint32_t width, uint32_t height, uint64_t out_off);
/* The encoder: CanHuffmanCompressU16.Init + Compress (canhuffmancompressu16.go:46-81) and the test composition
 * DeltaRleCompressU16 -> CanHuffman (fseu16_test.go:881-889).  Histogram and bit emission on the device, the code
 * construction (<= 65536 list entries) on the host between them.  *out_len receives the stream size (also with
 * MICGPU_E_SIZE).  Symbols of equal frequency are ordered by a stable sort (ascending symbol, the delimiter last) where
 * the reference's sort.Slice is unstable: the bytes can differ from Go's there; either decoder reads either stream. */
int micgpu_huff_compress(const uint16_t *symbols, size_t n, uint8_t *out, size_t cap, size_t *out_len);
int micgpu_delta_rle_huff_compress(const uint16_t *pixels, int width, int height, uint16_t max_value, uint8_t *out, size_t cap,
                                   size_t *out_len);
/* ... and their encoders, WaveletFSECompressU16 (with_rle = 0, :71-123) / WaveletRLEFSECompressU16 (with_rle = 1, :551-623):
 * levels are clamped to [1, 4] as the reference does. */
int micgpu_wavelet_v1_compress(const uint16_t *pixels, int rows, int cols, uint16_t max_value, int levels, int with_rle,
                               uint8_t *out, size_t cap, size_t *out_len);
int micgpu_wavelet_v2_decompress_batch(int n, const uint8_t *const *blobs, const size_t *lens, uint16_t *const *outs, const size_t *caps,
                                       int *rows, int *cols, int *status);

/* ---- encode direction -------------------------------------------------------------- */
/* Stage calls (used by the parity tests): DeltaRleCompressU16.Compress (deltarlecompressu16.go:24) and
 * RleCompressU16.Init(len,1,maxValue)+Compress (rlecompressu16.go:15-93). */
int micgpu_delta_rle_compress(const uint16_t *pixels, int width, int height, uint16_t max_value, uint16_t *out, size_t cap, size_t *out_len);
int micgpu_rle_compress(const uint16_t *in, size_t n, uint16_t max_value, uint16_t *out, size_t cap, size_t *out_len);
/* CompressSingleFrame (nstates 2), CompressSingleFrame4State (4), CompressSingleFrame8State (8) with the reference's
 * fallback ladder (multiframecompress.go:15-93); nstates 1 = Delta+RLE + FSECompressU16; nstates MICGPU_CODER_RANS8 =
 * Delta+RLE + RANSCompressU16EightState (rans8state.go:31, magic [0xFF,0x08], no fallback).  maxValue is the caller's. */
#define MICGPU_CODER_RANS8 108
int micgpu_compress_single_frame(const uint16_t *pixels, int width, int height, uint16_t max_value, int nstates, uint8_t *out, size_t cap,
                                 size_t *out_len);
/* CompressParallelStrips / 4State / 8State (parallelstrips.go:55,128,199); num_strips must be > 0 (GOMAXPROCS is the caller's). */
int micgpu_pics_compress(const uint16_t *pixels, int width, int height, uint16_t max_value, int num_strips, int nstates, uint8_t *out,
                         size_t cap, size_t *out_len);
int micgpu_pics_compress_batch(int n, const uint16_t *const *pixels, int width, int height, const uint16_t *max_values, int num_strips,
                               int nstates, uint8_t *const *outs, const size_t *caps, size_t *out_lens, int *status);
/* CompressSingleFrameGrad (multiframecompress.go:111-127): gradient-adaptive Delta + RLE, 2-state FSE with 1-state fallback. */
int micgpu_compress_single_frame_grad(const uint16_t *pixels, int width, int height, uint16_t max_value, uint8_t *out, size_t cap,
                                      size_t *out_len);
/* CompressParallelStripsAdaptive (parallelstripsadaptive.go:54-139): adaptive strip rows (equal-cost partition of the
 * inter-row |delta| sums), both predictors per strip, the smaller frame kept; num_strips must be > 0. */
int micgpu_pica_compress(const uint16_t *pixels, int width, int height, uint16_t max_value, int num_strips, uint8_t *out, size_t cap,
                         size_t *out_len);
/* adaptiveStripBoundaries (parallelstripsadaptive.go:214-289): starts_out holds min(num_strips, height) ints. */
int micgpu_pica_boundaries(const uint16_t *pixels, int width, int height, int num_strips, int *starts_out, int *n_out);
/* CompressMultiFrame (multiframecompress.go:179): frames contiguous, independent or temporal (ZigZag residual) mode. */
int micgpu_mic2_compress(const uint16_t *frames, int width, int height, int nframes, uint16_t max_value, int temporal, uint8_t *out,
                         size_t cap, size_t *out_len);
/* WaveletV2[SIMD]RLEFSECompressU16 (waveletfsecompressu16.go:303,427); note the (rows, cols) argument order of the Go API. */
int micgpu_wavelet_v2_compress(const uint16_t *pixels, int rows, int cols, uint16_t max_value, int levels, uint8_t *out, size_t cap,
                               size_t *out_len);
int micgpu_wavelet_v2_compress_batch(int n, const uint16_t *const *pixels, int rows, int cols, const uint16_t *max_values, int levels,
                                     uint8_t *const *outs, const size_t *caps, size_t *out_lens, int *status);
/* CompressRGB (rgbcompress.go:25) and CompressWSI (wsicompress.go:27; tile_w/tile_h 0 = 256, pyramid_levels <= 0 = auto). */
int micgpu_rgb_compress(const uint8_t *rgb, int width, int height, uint8_t *out, size_t cap, size_t *out_len);
int micgpu_wsi_compress(const uint8_t *pixels, int width, int height, int channels, int bits_per_sample, int tile_w, int tile_h,
                        int pyramid_levels, uint8_t *out, size_t cap, size_t *out_len);
void micgpu_encoder_shutdown(void);

/* ---- .mic file wrappers (cmd/mic-compress/main.go:26-91 writes them, cmd/mic-wasm/main.go:52-131 reads them) -------- */
/* 1 MIC1 (single 16-bit frame), 2 MIC2, 3 MIC3, 4 MICR (RGB), 5 PICS, 0 unknown: route to the call of that container. */
int micgpu_file_kind(const uint8_t *file, size_t len);
/* MIC1 = "MIC1", width, height, pipeline (1 = Delta+RLE+FSE), length (u32 LE each), one frame. */
int micgpu_mic1_compress(const uint16_t *pixels, int width, int height, uint16_t max_value, int nstates, uint8_t *out, size_t cap,
                         size_t *out_len);
int micgpu_mic1_decompress(const uint8_t *file, size_t len, uint16_t *pixels_out, size_t cap_px, int *width, int *height);
/* MICR = "MICR", width, height (u32 LE), CompressRGB blob. */
int micgpu_micr_compress(const uint8_t *rgb, int width, int height, uint8_t *out, size_t cap, size_t *out_len);
int micgpu_micr_decompress(const uint8_t *file, size_t len, uint8_t *rgb_out, size_t cap_bytes, int *width, int *height);
/* ojph/mic_compress_c.h:26-37 (maxValue derived from the pixels, no fallback ladder) */
int mic_compress_two_state(const uint16_t *pixels, int width, int height, uint8_t *out, size_t out_cap, size_t *out_len);
int mic_compress_four_state(const uint16_t *pixels, int width, int height, uint8_t *out, size_t out_cap, size_t *out_len);
int mic_compress_eight_state(const uint16_t *pixels, int width, int height, uint8_t *out, size_t out_cap, size_t *out_len);

/* ---- drop-in symbols of the reference C twin (same signatures) -------------- */
/* ojph/mic_decompress_c.h:24-49 */
int mic_decompress_two_state(const uint8_t *compressed, size_t compressed_len, uint16_t *pixels_out, int width, int height);
int mic_decompress_two_state_simd(const uint8_t *compressed, size_t compressed_len, uint16_t *pixels_out, int width, int height);
int mic_decompress_four_state(const uint8_t *compressed, size_t compressed_len, uint16_t *pixels_out, int width, int height);
int mic_decompress_four_state_simd(const uint8_t *compressed, size_t compressed_len, uint16_t *pixels_out, int width, int height);
int mic_decompress_eight_state(const uint8_t *compressed, size_t compressed_len, uint16_t *pixels_out, int width, int height);
int mic_decompress_eight_state_simd(const uint8_t *compressed, size_t compressed_len, uint16_t *pixels_out, int width, int height);
/* ojph/mic_parallel.h:49-55 (max_threads is accepted and ignored: one batched launch) */
int mic_decompress_parallel(const uint8_t *compressed, size_t compressed_len, uint16_t *pixels_out, int width, int height, int max_threads);
int mic_decompress_parallel_scalar(const uint8_t *compressed, size_t compressed_len, uint16_t *pixels_out, int width, int height, int max_threads);

#ifdef __cplusplus
}
#endif
#endif /* MICGPU_H */
