// mic_unit.h -- unit descriptor shared by host planning code and device kernels.
//
// A "unit" is one independent FSE frame of the MIC hot path: a PICS strip, a
// MIC2 frame (spatial or temporal residual), one plane of a MIC3 tile or a
// WaveletV2 coefficient stream (reference: multiframecompress.go:97,165;
// parallelstrips.go:292-321; wsicompress.go:487-524; waveletfsecompressu16.go:374-421).
#pragma once
#include <stdint.h>

enum MicUnitKind {
  MIC_KIND_SPATIAL = 0,  // Delta(avg(top,left),escape)+RLE symbol stream -> w*h pixels
  MIC_KIND_RLE = 1,      // RLE stream with 2-word length prefix (temporal residual / wavelet)
  MIC_KIND_RAW = 2,      // the FSE symbols are the payload (V1 wavelet without RLE, waveletfsecompressu16.go:124-163)
};

// MicUnit::rans (the entropy coder of the unit)
enum MicCoder {
  MIC_CODER_TANS = 0,    // FSE, 1/2/4/8 states
  MIC_CODER_RANS = 1,    // rANS-8 (magic [0xFF,0x08])
  MIC_CODER_HUFF = 2,    // canonical Huffman (canhuffmandecompressu16.go; no magic: chosen by the caller, k_huff.cu decodes it)
};

enum MicStatus {
  MIC_OK = 0,
  MIC_E_HEADER = -1,     // bad magic / truncated frame          (C twin: -1)
  MIC_E_NCOUNT = -2,     // readNCount corruption                (C twin: -2)
  MIC_E_ALLOC = -3,
  MIC_E_DTABLE = -4,     // buildDtable corruption               (C twin: -4)
  MIC_E_BITSTREAM = -6,  // bit reader over-read / zero last byte (C twin: -6)
  MIC_E_RLE = -8,        // RLE stream over-run or malformed header
  MIC_E_SIZE = -9,       // decoded size does not match the unit's geometry
  MIC_E_UNSUPPORTED = -10,
};

struct MicUnit {
  // ---- host-filled ------------------------------------------------------
  unsigned long long comp_off;  // byte offset of the FSE frame in the device compressed buffer
  unsigned long long sym_off;   // element offset (u16) in the state-stream scratch
  unsigned long long tab_off;   // entry offset in the decode-table scratch
  unsigned long long d_off;     // element offset (u16) in the residual plane scratch (pitch wp)
  unsigned long long m_off;     // word offset in the literal-mask scratch (pitch wp/32)
  unsigned long long out_off;   // element offset (u16) in the output buffer
  unsigned int comp_len;
  unsigned int kind;
  unsigned int width, height;   // spatial: image geometry; RLE: width = expected outlen, height = 1
  unsigned int wp;              // padded pitch (multiple of 32 elements, >= width + 8)
  unsigned int nstates;         // 1,2,4,8 (from the magic prefix, fse2state.go:102-116)
  unsigned int rans;            // MicCoder: 1 when magic is [0xFF,0x08], 2 for a canonical-Huffman stream
  unsigned int table_log;       // peeked from the ncount header (fsedecompressu16.go:61)
  unsigned int count;           // symbol count from the 6-byte prefix (0: 1-state, unknown)
  unsigned int sym_cap;         // capacity (elements) reserved at sym_off
  unsigned int align0;          // (output element address of pixel (0,0)) mod 8, set per run; row y of D and M
                                // is stored at padded column a_y + x with a_y = (align0 + y*width) mod 8
  unsigned int exact_len;       // RLE kind: 1 = the expanded length must equal `width` (a MIC2 residual frame is exactly
                                // one frame, multiframecompress.go:165-175); 0 = `width` is only a capacity
  unsigned int predictor;       // spatial kind: 0 = avg(top,left) (deltarlecompressu16.go), 1 = gradient-adaptive
                                // (deltagradrlecompressu16.go; PICA strips with picaFlagGradPredictor, DecompressSingleFrameGrad)
  unsigned int pad0;
  // ---- device-filled ----------------------------------------------------
  unsigned int bits_off;        // byte offset of the bitstream inside the frame
  unsigned int bits_len;        // bitstream length in bytes
  unsigned int npres;           // K1a -> K1b: symbols with a non-zero normalised count (0xFFFFFFFF: one-kernel fallback)
  unsigned int nsym;            // symbols actually decoded
  unsigned int thr;             // deltaThreshold (deltarlecompressu16.go:72-74)
  unsigned int delim;           // delimiterForOverflow
  unsigned int k4_redo;         // row-scan predictor kernel saw a uint16 wrap: the wavefront kernel redoes the unit
  int status;                   // MicStatus
};
