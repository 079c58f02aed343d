// mic_device.cuh -- shared device-side declarations for the micgpu kernels (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "mic_enc.h"
#include "mic_unit.h"

namespace micgpu {

// ---- launchers (defined in the k_*.cu files) -------------------------------
void launch_build_tables(MicUnit* d_units, int nunits, const uint8_t* d_comp, uint32_t* d_tabA, uint16_t* d_tabS,
                         uint8_t* d_scratch, unsigned long long scratch_stride, int max_log, int grid, cudaStream_t st);

// Split table build (k_table.cu): K1a one warp per unit parses the ncount header, K1b one CTA per unit builds the
// decode table in shared memory; d_norm / d_sym hold the present-symbol lists at each unit's tab_off.  Units the
// split path cannot take are built by the one-kernel path on fallback_grid CTAs (0 = none can occur).
void launch_parse_ncount(MicUnit* d_units, int nunits, const uint8_t* d_comp, int32_t* d_norm, uint16_t* d_sym, cudaStream_t st);
void launch_build_dtable(MicUnit* d_units, int nunits, const uint8_t* d_comp, uint32_t* d_tabA, uint16_t* d_tabS,
                         const int32_t* d_norm, const uint16_t* d_sym, unsigned int* d_queue, uint8_t* d_scratch,
                         unsigned long long scratch_stride, int max_log, int fallback_grid, int sm_count, cudaStream_t st);

// ANS decode of the units listed in d_list (all with the same state count).
// Output: the *state* stream (table indices, u16) at units[i].sym_off; symbols
// are recovered later through tabS.  smem_mode: 0 = u32 entries in shared
// memory, 1 = u16 nextState entries in shared memory, 2 = table stays in L2.
void launch_ans_decode(MicUnit* d_units, const int* d_list, int nlist, int nstates, const uint8_t* d_comp,
                       const uint32_t* d_tabA, uint16_t* d_states, int max_log, int smem_mode, int slots_per_cta,
                       int grid, cudaStream_t st);
size_t ans_decode_smem_bytes(int max_log, int smem_mode, int slots_per_cta);

// Thread-per-unit decode of 1-, 2- and 4-state streams (k_ans_serial.cu).  mode: 0 = 4-byte cells, 1 = 2-byte cells.
// The list is cut into segments (first, count) whose tables fit one CTA; d_slot_off[i] = shared-memory byte offset of
// list entry i inside its segment; smem = bytes of the largest segment; max_slots = entries of the largest one (<= 128).
void launch_ans_decode_serial(MicUnit* d_units, const int* d_list, const int2* d_segs, int nseg, const uint32_t* d_slot_off, int nstates,
                              const uint8_t* d_comp, const uint32_t* d_tabA, uint16_t* d_states, int mode, int max_slots,
                              size_t smem, int grid, cudaStream_t st);
size_t ans_serial_unit_bytes(int table_log, int mode);
size_t ans_serial_smem_bytes(int max_log, int mode, int slots);

// Canonical-Huffman streams (k_huff.cu): one CTA per listed unit parses the header, builds the code table in the unit's
// tabA region, writes an identity tabS and decodes the symbols to the unit's place in the state stream (self-synchronising
// subsequences, one per thread; serial != 0: one thread walks the chain, for A/B runs).
void launch_huff_decode(MicUnit* d_units, const int* d_list, int nlist, const uint8_t* d_comp, uint32_t* d_tabA, uint16_t* d_tabS,
                        uint16_t* d_states, int serial, int sm_count, cudaStream_t st);

// RLE expand (+ escape split for spatial units).  Spatial units produce the
// residual plane D (pitch wp) and the literal bit mask M; RLE units write
// their expanded stream straight to d_out.
// Units [ubase, ubase + nunits) are handed to the CTAs through the atomic counter *d_queue (reset by the launcher).
void launch_rle_expand(MicUnit* d_units, int ubase, int nunits, const uint16_t* d_states, const uint16_t* d_tabS,
                       uint16_t* d_D, uint32_t* d_M, uint16_t* d_out, int max_log, int grid, unsigned int* d_queue,
                       cudaStream_t st, int shape = 0);

// Inverse avg(top,left) predictor as an anti-diagonal wavefront; one CTA per spatial unit.
void launch_delta_wavefront(MicUnit* d_units, const int* d_list, int nlist, const uint16_t* d_D, const uint32_t* d_M,
                            uint16_t* d_out, int max_width, int max_height, cudaStream_t st, int redo_only = 0);
// The same predictor one row per step with lanes = 8-pixel column blocks and a warp scan over block functions
// (k_delta_scan.cu).  Units whose pixels wrap uint16 are flagged (MicUnit::k4_redo) for the wavefront kernel.
// Returns false (nothing launched) when the widest unit needs more than 32 warps.
bool launch_delta_rowscan(MicUnit* d_units, const int* d_list, int nlist, const uint16_t* d_D, const uint32_t* d_M,
                          uint16_t* d_out, int max_width, bool all_aligned, cudaStream_t st);
int delta_wavefront_threads(int max_width, int max_height);
// Gradient-adaptive predictor (k_grad.cu): one CTA per listed unit with MicUnit::predictor == 1 (the others return at once).
void launch_grad_wavefront(MicUnit* d_units, const int* d_list, int nlist, const uint16_t* d_D, const uint32_t* d_M, uint16_t* d_out,
                           int max_height, cudaStream_t st);
// PICA row statistics (adaptiveStripBoundaries): cost[y] = sum_x |px[y][x] - px[y-1][x]|, cost[0] = 0
void launch_row_costs(const uint16_t* d_px, int width, int height, unsigned long long* d_cost, cudaStream_t st);

// In-place frame-axis running sum for temporal MIC2 (frames contiguous, fpx pixels each).
void launch_temporal_accumulate(uint16_t* d_frames, unsigned long long fpx, int nframes, int first_is_residual, int sm_count,
                                cudaStream_t st);
// Last frames of the ranges before this one, readable from this GPU (local or peer memory), for the fused carry kernel.
struct PeerFrames {
  const uint16_t* last[16];
  int n;
  int aligned;     // every pointer, the frame buffer and frame_px allow 16 B accesses
};
void launch_temporal_add_carry_peers(uint16_t* d_frames, const PeerFrames& peers, unsigned long long fpx, int nframes, int sm_count,
                                     cudaStream_t st);
// frames[f] += carry (mod 2^16) for a shard of a temporal stack decoded relative to a zero carry
void launch_temporal_add_carry(uint16_t* d_frames, const uint16_t* d_carry, unsigned long long fpx, int nframes, int sm_count,
                               cudaStream_t st);

// MIC3 helpers (k_misc.cu)
struct PlaneFillJob {
  unsigned long long plane_off;   // element offset in the plane buffer
  unsigned long long comp_off;    // byte offset of the raw payload (kind 1)
  unsigned long long n;           // samples
  unsigned int kind;              // 0 constant, 1 raw little-endian u16
  unsigned int value;
};
struct TileBlitJob {
  unsigned long long plane_off[3];   // element offset of each plane (unused for a constant plane)
  unsigned long long dst_off;        // byte offset of the destination rectangle
  unsigned int tile_w, tile_h;
  unsigned int src_x, src_y, copy_w, copy_h;
  unsigned int dst_pitch;            // bytes per destination row
  unsigned int mode;                 // 0 YCoCg-R -> RGB8, 1 planar RGB -> RGB8, 2 grey8, 3 grey16
  unsigned int cmask;                // bit k: plane k is constant (plane modes 0/1, wsicompress.go:487-500): no plane is materialised
  unsigned short cval[4];            // its value
};
void launch_plane_fill(const PlaneFillJob* d_jobs, int njobs, const uint8_t* d_comp, uint16_t* d_planes, cudaStream_t st);
void launch_tile_blit(const TileBlitJob* d_jobs, int njobs, const uint16_t* d_planes, uint8_t* d_out, cudaStream_t st);

// WaveletV2 decode (k_wavelet.cu)
struct WaveletGeom {
  unsigned rows, cols;
  int levels;                 // levels actually applied (header byte 10; up to 255 for the V1 layouts, whose geometry is one segment)
  int nseg;                   // 1 + 3*levels subband segments in stream order
  unsigned seg_start[25], seg_y0[25], seg_x0[25], seg_w[25];
};
void launch_wavelet_decode(MicUnit* d_units, const int* d_unit_of_img, int nimg, const uint16_t* d_stream, int* d_flags,
                           int32_t* d_A, int32_t* d_B, uint16_t* d_px, const WaveletGeom& G, cudaStream_t st, int v1_layout = 0);

// Encode side (k_enc_rle.cu, k_enc_fse.cu)
void launch_enc_delta_rle(MicEncUnit* d_units, int nunits, const uint16_t* d_src, uint16_t* d_V, uint32_t* d_segs, uint16_t* d_S,
                          int grid, cudaStream_t st);
void launch_enc_tables(MicEncUnit* d_units, int nunits, const uint16_t* d_S, uint8_t* d_scratch, uint16_t* d_state_tab, uint2* d_sym_tt,
                       uint8_t* d_hdrs, int grid, cudaStream_t st);
unsigned long long enc_tables_scratch_per_cta();
void launch_enc_ans(MicEncUnit* d_units, const int* d_list, int nlist, int nstates, const uint16_t* d_S, const uint16_t* d_state_tab,
                    const uint2* d_sym_tt, uint32_t* d_T, int sm_count, cudaStream_t st, int max_table_log = 0, bool any_rans = false);
void launch_enc_pack(MicEncUnit* d_units, int nunits, const uint32_t* d_T, const uint8_t* d_hdrs, uint8_t* d_frames, int grid, cudaStream_t st);

// Encode front ends (k_enc_front.cu)
struct TilePlaneJob {
  unsigned long long img_off;     // byte offset of the level image in the image buffer
  unsigned long long plane_off;   // element offset of the tile's first plane
  unsigned int img_w, img_h, x0, y0, tile_w, tile_h;
  unsigned int mode;              // 0 RGB8 -> Y,Co,Cg ; 1 RGB8 -> R,G,B ; 2 grey8 ; 3 grey16
  unsigned int pad;
};
struct GatherJob { unsigned long long src_off, dst_off, len; };
void launch_temporal_residual(const uint16_t* d_frames, uint16_t* d_res, unsigned long long fpx, int nframes, int sm_count, cudaStream_t st);
void launch_downsample2x(const uint8_t* d_src, uint8_t* d_dst, unsigned w, unsigned h, unsigned ch, unsigned bytes_per_sample, int sm_count, cudaStream_t st);
void launch_tile_planes(const TilePlaneJob* d_jobs, int njobs, const uint8_t* d_images, uint16_t* d_planes, cudaStream_t st);
void launch_plane_stats(const uint16_t* d_planes, const unsigned long long* d_offs, unsigned long long npx, int nplanes, unsigned* d_stats, cudaStream_t st);
void launch_wavelet_forward(const uint16_t* d_px, int32_t* d_A, int32_t* d_B, int nimg, const WaveletGeom& G, int sm_count, cudaStream_t st);
// V1 layouts: interleaved in-place lifting on the top-left corner per level (result in d_A)
void launch_wavelet_forward_v1(const uint16_t* d_px, int32_t* d_A, int32_t* d_B, int nimg, unsigned rows, unsigned cols, int levels, int sm_count,
                               cudaStream_t st);
void launch_wavelet_pack(const int32_t* d_A, uint16_t* d_V, MicEncUnit* d_units, const int* d_unit_of_img, int nimg, const WaveletGeom& G, cudaStream_t st);
void launch_gather_bytes(const GatherJob* d_jobs, int njobs, const uint8_t* d_src, uint8_t* d_dst, cudaStream_t st);
void launch_plane_raw(const GatherJob* d_jobs, int njobs, const uint16_t* d_planes, uint8_t* d_dst, cudaStream_t st);

// ---- small device helpers ---------------------------------------------------
__device__ __forceinline__ uint32_t ld_u32_unaligned_safe(const uint8_t* p) {
  return (uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16) | ((uint32_t)p[3] << 24);
}
__device__ __forceinline__ int bit_len16(uint32_t v) { return 32 - __clz(v); }  // bits.Len16 for v < 65536

}  // namespace micgpu
