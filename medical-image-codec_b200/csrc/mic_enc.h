// mic_enc.h -- encode-side unit descriptor shared by host planning code and device kernels.
//
// One unit = one FSE frame to produce: a PICS strip / MIC2 frame / MIC3 plane (spatial: Delta+RLE+FSE,
// multiframecompress.go:15-93), a temporal residual frame (RLE+FSE, :146-162) or a WaveletV2 coefficient
// stream (RLE + 4-state FSE, waveletfsecompressu16.go:345-356).
#pragma once
#include <stdint.h>

enum MicEncKind {
  MIC_ENC_SPATIAL = 0,   // source = pixels (w x h); V = [maxValue, delta symbols / escapes]
  MIC_ENC_RLE = 1,       // source = ready-made symbol stream V (residuals / wavelet coefficients); RleCompressU16.Compress
  MIC_ENC_RAW = 2,       // source = ready-made symbol stream that goes to FSE as it is (WaveletFSECompressU16, waveletfsecompressu16.go:71-123)
};

enum MicEncStatus {
  MIC_ENC_OK = 0,
  MIC_ENC_INCOMPRESSIBLE = -1,   // fseu16.go:33 ErrIncompressible
  MIC_ENC_USE_RLE = -2,          // fseu16.go:36 ErrUseRLE
  MIC_ENC_INTERNAL = -6,         // normalisation / table construction failed (Go: "internal error", weight < 1, ...)
  MIC_ENC_UNSUPPORTED = -10,     // maxValue with fewer than 4 bits, maxValue == 0 (Go panics on 1<<-1)
  MIC_ENC_CAPACITY = -9,
};

struct MicEncUnit {
  // ---- host-filled ------------------------------------------------------
  unsigned long long src_off;   // element offset (u16) of the source (pixels or V) in the source buffer
  unsigned long long v_off;     // element offset of V in the pre-RLE scratch (spatial only; RLE kind reads the source)
  unsigned long long s_off;     // element offset of the RLE symbol stream S
  unsigned long long seg_off;   // entry offset in the segment-list scratch (2 u32 per entry)
  unsigned long long t_off;     // entry offset in the per-symbol (bits, nbBits) scratch
  unsigned long long tab_off;   // entry offset in the encode-table scratch (stateTable u16 x 2^16 max)
  unsigned long long tt_off;    // entry offset in the symbolTT scratch
  unsigned long long hdr_off;   // byte offset in the ncount-header scratch
  unsigned long long out_off;   // byte offset of this unit's frame in the frame scratch
  unsigned int width, height;   // spatial geometry; RLE kind: width = length of V, height = 1
  unsigned int max_value;       // caller's maxValue (spatial) / resMax or rleMaxVal (RLE kind);
                                // 0xFFFFFFFF: max(V) (resMax, multiframecompress.go:194-199); 0xFFFFFFFE: (1<<bits.Len16(max(V)))-1,
                                // depth >= 1 (rleMaxVal, waveletfsecompressu16.go:337-343)
  unsigned int kind;
  unsigned int nstates;         // requested tier 8/4/2/1
  unsigned int no_ladder;       // 1: a rejected tier is final (WaveletV2 uses FSECompressU16FourState with no fallback)
  unsigned int predictor;       // spatial kind: 0 = avg(top,left), 1 = gradient-adaptive (GradDeltaRleCompressU16, deltagradrlecompressu16.go:26-68)
  unsigned int rans;            // 1: RANSCompressU16EightState (rans8state.go:31): nstates 8, magic 0x08, no fallback ladder
  unsigned int v_cap, s_cap, out_cap;
  // ---- device-filled ----------------------------------------------------
  unsigned int v_len;           // symbols in V
  unsigned int nseg;            // RLE segments
  unsigned int s_len;           // symbols in S
  unsigned int symbol_len;      // highest symbol + 1
  unsigned int table_log;
  unsigned int hdr_len;         // ncount header bytes
  unsigned int bits_total;      // payload bits before the sentinel
  unsigned int frame_len;       // bytes of the finished frame
  unsigned int used_states;     // tier actually used (after the 8->4->2->1 ladder)
  unsigned int mid;             // RLE midCount
  int status;
};
