/*
 * mic_oracle.h -- CPU oracle for the MIC hot path (TEST INFRASTRUCTURE ONLY).
 *
 * A plain-C restatement of the Go semantics of pappuks/medical-image-codec for
 * the path named in BASELINE.json: Delta(avg(top,left),escape) -> RLE-u16 ->
 * FSE (1/2/4/8-state tANS, rANS-8), temporal ZigZag, YCoCg-R, WaveletV2 5/3,
 * pyramid, and the PICS / MIC2 / MIC3 containers.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load this library.  The product (libmicgpu.so)
 * never links or calls it.
 *
 * Parity pin: validated against the compiled reference C twin (oracle/_ref,
 * built from /root/reference/ojph/mic_{compress,decompress}_c.c,
 * mic_parallel.c) byte-for-byte on 2/4/8-state streams, and against every
 * known-answer value in the reference's own tests (tests/test_oracle_*.py).
 * Stream byte-identity versus the *Go* encoder for 1-state, rANS-8, temporal,
 * wavelet and MIC3 is pinned only by restatement-from-source (the reference
 * holds no golden compressed vectors).
 *
 * Every function cites the reference file:line it follows (paths relative to
 * the reference repository root).
 */
#ifndef MIC_ORACLE_H
#define MIC_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* error codes (all entry points return >= 0 on success) */
#define ORC_ERR_INCOMPRESSIBLE (-1) /* fseu16.go:33 ErrIncompressible */
#define ORC_ERR_USE_RLE (-2)        /* fseu16.go:36 ErrUseRLE */
#define ORC_ERR_CORRUPT (-3)
#define ORC_ERR_ARG (-4)
#define ORC_ERR_INTERNAL (-6)

/* entropy coder selectors */
#define ORC_FSE1 1
#define ORC_FSE2 2
#define ORC_FSE4 4
#define ORC_FSE8 8
#define ORC_RANS8 108

void orc_free(void *p);

/* ---- L1: entropy coder -------------------------------------------------- */
/* FSECompressU16 / TwoState / FourState / EightState / RANSCompressU16EightState
 * (fsecompressu16.go:19, fse2state.go:22, fse4state.go:25, fse8state.go:32,
 * rans8state.go:32).  No fallback ladder. *out is malloc'ed. */
int orc_fse_compress(const uint16_t *in, size_t n, int coder, uint8_t **out, size_t *out_len);
/* FSEDecompressU16Auto (fse2state.go:102). *out is malloc'ed. */
int orc_fse_decompress_auto(const uint8_t *in, size_t len, uint16_t **out, size_t *out_len);
/* diagnostics: tableLog / symbolLen / norm[] chosen for an input */
int orc_fse_table_info(const uint16_t *in, size_t n, int *table_log, int *symbol_len, int32_t *norm_out /* 65536 or NULL */);

/* CanHuffmanCompressU16.Init+Compress / CanHuffmanDecompressU16.Init+ReadTable+Decompress (canhuffmancompressu16.go:46-81,
 * canhuffmandecompressu16.go:31-108; mic_oracle_huff.c).  Encoder bytes: ties between equal frequencies are ordered by a
 * stable sort here, by Go's unstable sort.Slice there (parity unpinned); the decoder is fully determined by the stream. */
int orc_huff_compress(const uint16_t *in, size_t n, uint8_t **out, size_t *out_len);
int orc_huff_decompress(const uint8_t *in, size_t len, uint16_t **out, size_t *out_len);
/* DeltaRleCompressU16 -> CanHuffman (fseu16_test.go:881-889) / DeltaRleHuffDecompressU16.Decompress (deltarlehuffdecompressu16.go:19-39) */
int orc_delta_rle_huff_compress(const uint16_t *px, int width, int height, uint16_t max_value, uint8_t **out, size_t *out_len);
int orc_delta_rle_huff_decompress(const uint8_t *in, size_t len, int width, int height, uint16_t *px_out);

/* ---- L2: transforms ----------------------------------------------------- */
/* RleCompressU16.Init(len,1,maxValue)+Compress (rlecompressu16.go:15-93) */
int orc_rle_compress(const uint16_t *in, size_t n, uint16_t max_value, uint16_t **out, size_t *out_len);
/* RleDecompressU16.Init+Decompress (rledecompressu16.go:21-97) */
int orc_rle_decompress(const uint16_t *in, size_t n, uint16_t **out, size_t *out_len);
/* DeltaRleCompressU16.Compress (deltarlecompressu16.go:24-68) */
int orc_delta_rle_compress(const uint16_t *px, int width, int height, uint16_t max_value, uint16_t **out, size_t *out_len);
/* DeltaRleDecompressU16.Decompress (deltarlecompressu16.go:69-128) */
int orc_delta_rle_decompress(const uint16_t *in, size_t n, int width, int height, uint16_t *px_out);
/* ZigZag / UnZigZag (deltazigzagcompressu16.go:108-116) */
uint16_t orc_zigzag(int16_t x);
int16_t orc_unzigzag(uint16_t x);
/* TemporalDeltaEncode / Decode (temporaldelta.go:11-39) */
void orc_temporal_encode(const uint16_t *cur, const uint16_t *prev, size_t n, uint16_t *out);
void orc_temporal_decode(const uint16_t *res, const uint16_t *prev, size_t n, uint16_t *out);
/* YCoCgRForward / Inverse (ycocgr.go:19-35, asm_generic.go:25-53) */
void orc_ycocg_forward(const uint8_t *rgb, size_t n, uint16_t *y, uint16_t *co, uint16_t *cg);
void orc_ycocg_inverse(const uint16_t *y, const uint16_t *co, const uint16_t *cg, size_t n, uint8_t *rgb);
/* Downsample2xRGB / Grey (wsipyramid.go:10-55); return 0 and set nw,nh (0 when degenerate) */
int orc_downsample2x_rgb(const uint8_t *src, int w, int h, uint8_t *dst, int *nw, int *nh);
int orc_downsample2x_grey(const uint16_t *src, int w, int h, uint16_t *dst, int *nw, int *nh);
/* wt53Forward2DSeparated / Inverse (waveletu16.go:162-257), in place on int32 */
void orc_wt53_forward_2d(int32_t *data, int rows, int cols, int full_cols);
void orc_wt53_inverse_2d(int32_t *data, int rows, int cols, int full_cols);
void orc_wt53_forward_1d(int32_t *data, int offset, int n, int stride);
void orc_wt53_inverse_1d(int32_t *data, int offset, int n, int stride);

/* ---- L3: single-unit pipelines ----------------------------------------- */
/* CompressSingleFrame[4State/8State] with the 8->4->2->1 ladder
 * (multiframecompress.go:15-93). nstates in {1(=plain FSECompressU16),2,4,8}. */
int orc_compress_single_frame(const uint16_t *px, int width, int height, uint16_t max_value, int nstates, uint8_t **out, size_t *out_len);
/* DecompressSingleFrame (multiframecompress.go:97) */
int orc_decompress_single_frame(const uint8_t *in, size_t len, int width, int height, uint16_t *px_out);
/* compressResidualFrame / decompressResidualFrame (multiframecompress.go:146-175) */
int orc_compress_residual_frame(const uint16_t *res, size_t n, uint16_t max_value, uint8_t **out, size_t *out_len);
int orc_decompress_residual_frame(const uint8_t *in, size_t len, uint16_t **out, size_t *out_len);
/* WaveletV2RLEFSECompressU16 / Decompress (waveletfsecompressu16.go:303-421) */
int orc_wavelet_v2_compress(const uint16_t *px, int rows, int cols, uint16_t max_value, int levels, uint8_t **out, size_t *out_len);
int orc_wavelet_v2_decompress(const uint8_t *in, size_t len, uint16_t **px_out, int *rows, int *cols);

/* WaveletFSECompressU16 / WaveletRLEFSECompressU16 and their decoders (V1 layouts, waveletfsecompressu16.go:71-189,551-669) */
int orc_wavelet_v1_compress(const uint16_t *px, int rows, int cols, uint16_t max_value, int levels, int with_rle, uint8_t **out, size_t *out_len);
int orc_wavelet_v1_decompress(const uint8_t *in, size_t len, int with_rle, uint16_t **px_out, int *rows, int *cols);

/* ---- L4: containers ------------------------------------------------------ */
/* CompressParallelStrips[4State/8State] / DecompressParallelStrips (parallelstrips.go:55-330) */
int orc_pics_compress(const uint16_t *px, int width, int height, uint16_t max_value, int num_strips, int nstates, uint8_t **out, size_t *out_len);
int orc_pics_decompress(const uint8_t *in, size_t len, uint16_t **px_out, int *width, int *height);
/* GradDeltaRleCompressU16 / GradDeltaRleDecompressU16 (deltagradrlecompressu16.go:26-133) */
int orc_grad_delta_rle_compress(const uint16_t *px, int width, int height, uint16_t max_value, uint16_t **out, size_t *out_len);
int orc_grad_delta_rle_decompress(const uint16_t *in, size_t n, int width, int height, uint16_t *px_out);
/* CompressSingleFrameGrad / DecompressSingleFrameGrad (multiframecompress.go:111-142) */
int orc_compress_single_frame_grad(const uint16_t *px, int width, int height, uint16_t max_value, uint8_t **out, size_t *out_len);
int orc_decompress_single_frame_grad(const uint8_t *in, size_t len, int width, int height, uint16_t *px_out);
/* adaptiveStripBoundaries, CompressParallelStripsAdaptive / DecompressParallelStripsAdaptive (parallelstripsadaptive.go:54-289) */
int orc_pica_boundaries(const uint16_t *px, int width, int height, int num_strips, int *starts /* >= min(num_strips, height) */);
int orc_pica_compress(const uint16_t *px, int width, int height, uint16_t max_value, int num_strips, uint8_t **out, size_t *out_len);
int orc_pica_decompress(const uint8_t *in, size_t len, uint16_t **px_out, int *width, int *height);
/* CompressMultiFrame / DecompressMultiFrame / DecompressFrame (multiframecompress.go:179-315) */
int orc_mic2_compress(const uint16_t *frames, int width, int height, int nframes, uint16_t max_value, int temporal, uint8_t **out, size_t *out_len);
int orc_mic2_decompress(const uint8_t *in, size_t len, uint16_t **frames_out, int *width, int *height, int *nframes, int *temporal);
int orc_mic2_decompress_frame(const uint8_t *in, size_t len, int frame_idx, uint16_t **px_out, int *width, int *height);
/* compressRGBTileBlob / decompressRGBTileBlob == CompressRGB / DecompressRGB (rgbcompress.go:25-33, wsicompress.go:319-484) */
int orc_rgb_compress(const uint8_t *rgb, int width, int height, int color_transform, uint8_t **out, size_t *out_len);
int orc_rgb_decompress(const uint8_t *in, size_t len, int width, int height, int color_transform, uint8_t *rgb_out);
/* compressWSIPlane / decompressWSIPlane (wsicompress.go:373-421, 487-524) */
int orc_wsi_plane_compress(const uint16_t *plane, int width, int height, uint8_t **out, size_t *out_len);
int orc_wsi_plane_decompress(const uint8_t *in, size_t len, int width, int height, uint16_t *out);
/* CompressWSI / DecompressWSITile / DecompressWSIRegion (wsicompress.go:27-305) */
int orc_wsi_compress(const uint8_t *pixels, int width, int height, int channels, int bits_per_sample,
                     int tile_w, int tile_h, int pyramid_levels, uint8_t **out, size_t *out_len);
int orc_wsi_decompress_tile(const uint8_t *in, size_t len, int level, int tx, int ty, uint8_t **out, size_t *out_len, int *tw, int *th);
int orc_wsi_decompress_region(const uint8_t *in, size_t len, int level, int x, int y, int w, int h, uint8_t **out, size_t *out_len, int *ow, int *oh);
/* ReadMIC3Header summary: fills up to max_levels*5 ints (w,h,tilesX,tilesY,firstTileIdx) */
int orc_wsi_header(const uint8_t *in, size_t len, int *width, int *height, int *tile_w, int *tile_h, int *channels, int *bps,
                   int *color_transform, int *nlevels, int *level_info, int max_levels, uint64_t *total_tiles);

#ifdef __cplusplus
}
#endif
#endif
