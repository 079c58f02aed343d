// k_huff.cu -- K2h: canonical-Huffman back end of the legacy streams.
//
// Replaces CanHuffmanDecompressU16.ReadTable / Decompress (canhuffmandecompressu16.go:36-137, bitreaderhuff.go) and, with
// K3/K4 behind it, DeltaRleHuffDecompressU16.Decompress (deltarlehuffdecompressu16.go:19-39, rlehuffdecompressu16.go:21-51).
// One CTA per stream:
//   1. header: u32 symbol count, u16 maxValue, u8 maxCodeLength, u16 list size, the symbol list (pixelDepth bits each) and
//      the code lengths (Len8(maxCodeLength) bits each), all MSB first; every thread reads its own entries straight from
//      the bit string (fields have fixed widths, so entry j sits at a computed bit offset);
//   2. canonical codes: symbols per length by atomics, first code of each length by the reference's recurrence
//      (CalculateSymbolStartForCodeLength, canhuffmancompressu16.go:312-333), the rank of an entry among the earlier
//      entries of its length by __match_any_sync inside a warp and running per-length counters across warps, then the
//      2^maxCodeLength code -> (symbol, length, delimiter) table filled warp-cooperatively (a short code owns a long
//      interval of it);
//   3. decode.  A Huffman stream is one serial chain of variable-length codewords, but the chain re-synchronises: a
//      decoder started at a wrong bit falls onto the true codeword boundaries after a few symbols.  The data bits are cut
//      into subsequences of HUFF_SUB bits, one per thread; every thread decodes its subsequence from its (guessed) start and
//      publishes where the first codeword past its end begins; threads whose predecessor ended somewhere else than they
//      started decode again from there, until nothing moves (thread k is exact after k rounds at the latest, in practice
//      after two); a scan over the symbol counts gives every subsequence its output offset and one more pass writes the
//      symbols.  A codeword here is the code plus, for the delimiter, the pixelDepth raw bits that follow it.
// The symbols go where K2 puts ANS states; K3 maps states through tabS, so this kernel also writes an identity tabS for
// the unit (the host reserves tableLog-16 regions for it: 65536 code-table cells in tabA, 65536 identity cells in tabS).
// A read past the end of the stream, which Go answers with stale window bits or a slice panic, is MIC_E_BITSTREAM here.
// The encode direction (histogram, bit emission) is at the end of the file.
#include "mic_device.cuh"

namespace micgpu {

namespace {

constexpr int HUFF_THREADS = 256;
constexpr int HUFF_SUB = 512;   // bits per subsequence (longest codeword: 32 bits)

// n <= 32 bits of the big-endian bit string that starts at the 4-byte aligned address w, from bit `pos`
__device__ __forceinline__ uint32_t peek_be(const uint32_t* __restrict__ w, uint32_t pos) {
  const uint32_t i = pos >> 5;
  const uint32_t hi = __byte_perm(w[i], 0, 0x0123), lo = __byte_perm(w[i + 1], 0, 0x0123);
  return __funnelshift_l(lo, hi, pos & 31u);   // 32 bits starting at pos
}
__device__ __forceinline__ uint32_t get_be(const uint32_t* __restrict__ w, uint32_t pos, int n) {
  return n ? peek_be(w, pos) >> (32 - n) : 0u;
}

struct HuffDec {
  const uint32_t* w;     // aligned base of the bit string
  const uint32_t* tab;   // symbol | codeLen << 16 | isDelimiter << 24
  int depth, max_len, mmd_shift;   // mmd_shift = 32 - (delimiter code length + depth): GetSymbolAfterDelimiter (:144-146) on a top-aligned window
  uint32_t depth_mask;
  // one codeword at bit `pos`: symbol and bits used (0 = an uncovered table slot: only a damaged header leaves any)
  __device__ __forceinline__ uint32_t step(uint32_t pos, uint32_t& sym) const {
    const uint32_t win = peek_be(w, pos);
    const uint32_t e = tab[max_len ? win >> (32 - max_len) : 0u];
    uint32_t used = (e >> 16) & 0xFFu;
    sym = e & 0xFFFFu;
    if (e >> 24) {   // the delimiter: the symbol follows its code in pixelDepth bits (JustDecodeNext, :122-137)
      sym = depth ? (win >> mmd_shift) & depth_mask : 0u;
      used += depth;
    }
    return used;
  }
};

}  // namespace

__global__ void __launch_bounds__(HUFF_THREADS)
k_huff_decode(MicUnit* __restrict__ units, const int* __restrict__ list, int nlist, const uint8_t* __restrict__ comp,
              uint32_t* __restrict__ tabA, uint16_t* __restrict__ tabS, uint16_t* __restrict__ states_out, int serial) {
  __shared__ uint32_t s_per[33], s_start[33], s_run[33];
  __shared__ uint32_t s_end[HUFF_THREADS], s_scan[HUFF_THREADS / 32];
  __shared__ uint32_t s_bad, s_delim, s_final;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

  for (int li = blockIdx.x; li < nlist; li += gridDim.x) {
    MicUnit* U = &units[list[li]];
    __syncthreads();
    if (U->status != MIC_OK) continue;
    const uint8_t* frame = comp + U->comp_off;
    const uint32_t flen = U->comp_len;
    const uintptr_t fa = reinterpret_cast<uintptr_t>(frame);
    const uint32_t* w = reinterpret_cast<const uint32_t*>(fa & ~(uintptr_t)3);
    const uint32_t o0 = (uint32_t)(fa & 3) * 8;          // bit offset of the stream inside the aligned words
    const uint32_t nbits = o0 + flen * 8;                // end of the stream in those coordinates (flen < 2^28: host)
    // ---- fixed header (ReadTable, :36-52) ------------------------------------------------------------------------
    const uint32_t count = get_be(w, o0, 32);
    const uint32_t max_value = get_be(w, o0 + 32, 16);
    const int max_len = (int)get_be(w, o0 + 48, 8);
    const uint32_t nl = get_be(w, o0 + 56, 16);
    const int depth = bit_len16(max_value);
    const uint32_t depth_mask = depth ? 0xFFFFFFFFu >> (32 - depth) : 0u;
    const uint32_t delim = (1u << depth) - 1u;
    const int len_bits = 32 - __clz((unsigned)max_len);
    const int both = max_len + depth;
    const uint32_t sym_at = o0 + 72, len_at = sym_at + nl * depth, data_at = len_at + nl * len_bits;
    if (flen < 9 || max_len > 16 || (unsigned long long)sym_at + (unsigned long long)nl * (depth + len_bits) > nbits ||
        count != U->count || count > U->sym_cap) {
      if (tid == 0) U->status = (flen >= 9 && max_len > 16) ? MIC_E_UNSUPPORTED : MIC_E_HEADER;
      continue;
    }
    uint32_t* tab = tabA + U->tab_off;
    uint16_t* ident = tabS + U->tab_off;
    uint16_t* out = states_out + U->sym_off;
    // ---- tables ---------------------------------------------------------------------------------------------------
    if (tid < 33) { s_per[tid] = 0; s_run[tid] = 0; s_start[tid] = 0; }
    if (tid == 0) { s_bad = 0; s_delim = 0; s_final = 0xFFFFFFFFu; }
    for (uint32_t i = tid; i < (1u << max_len); i += HUFF_THREADS) tab[i] = 0;     // uncovered slots stay {0, 0, false}
    for (uint32_t i = tid; i < 65536u; i += HUFF_THREADS) ident[i] = (uint16_t)i;
    __syncthreads();
    for (uint32_t j = tid; j < nl; j += HUFF_THREADS) {
      const uint32_t l = get_be(w, len_at + j * len_bits, len_bits);
      if (l > (uint32_t)max_len) s_bad = 1;   // Go: index out of range in CalculateSymbolsPerCodeLength
      else atomicAdd(&s_per[l], 1u);
    }
    __syncthreads();
    if (s_bad) {
      if (tid == 0) U->status = MIC_E_DTABLE;
      continue;
    }
    if (tid == 0) {   // CalculateSymbolStartForCodeLength (canhuffmancompressu16.go:312-333)
      int prev = 0;
      uint32_t nprev = 0;
      for (int i = 1; i <= max_len; i++) {
        const uint32_t ns = s_per[i];
        if (ns) {
          s_start[i] = prev ? (s_start[prev] + nprev) << (i - prev) : 0u;
          prev = i; nprev = ns;
        }
      }
    }
    __syncthreads();
    // ConstructCanHuffmanTable (:335-344) + the code -> symbol table (canhuffmandecompressu16.go:60-77), list order kept:
    // chunks of HUFF_THREADS entries, inside a chunk the warps take turns at the per-length counters
    for (uint32_t j0 = 0; j0 < nl; j0 += HUFF_THREADS) {
      const uint32_t j = j0 + tid;
      const bool have = j < nl;
      const uint32_t l = have ? get_be(w, len_at + j * len_bits, len_bits) : 33u + (uint32_t)lane;   // idle lanes match nobody
      const uint32_t sym = have ? get_be(w, sym_at + j * depth, depth) : 0u;
      uint32_t code = 0;
      for (int wv = 0; wv < HUFF_THREADS / 32; wv++) {
        if (warp == wv) {
          const uint32_t m = __match_any_sync(0xFFFFFFFFu, l);
          const uint32_t before = __popc(m & ((1u << lane) - 1u));
          const uint32_t base = have ? s_run[l] : 0u;
          __syncwarp();
          code = have ? s_start[l] + base + before : 0u;
          if (have && before == 0) s_run[l] = base + __popc(m);
        }
        __syncthreads();
      }
      uint32_t base = 0, span = 0, e = 0;
      if (have) {
        base = code << (max_len - l);
        span = 1u << (max_len - l);
        if ((unsigned long long)base + span > (1ull << max_len) || (l < 32 && (code >> l))) { s_bad = 1; span = 0; }   // Go: index out of range
        e = sym | (l << 16) | (sym == delim ? 1u << 24 : 0u);
        if (sym == delim) atomicMax(&s_delim, (j << 8) | l | 0x80000000u);   // the last delimiter entry sets maxMinusDelimiterCodeLength
      }
      for (int src = 0; src < 32; src++) {
        const uint32_t b = __shfl_sync(0xFFFFFFFFu, base, src), sp = __shfl_sync(0xFFFFFFFFu, span, src);
        const uint32_t ee = __shfl_sync(0xFFFFFFFFu, e, src);
        for (uint32_t i = lane; i < sp; i += 32) tab[b + i] = ee;
      }
    }
    __syncthreads();
    if (s_bad) {
      if (tid == 0) U->status = MIC_E_DTABLE;
      continue;
    }
    const int dlen = (int)(s_delim & 0xFFu);
    HuffDec D{w, tab, depth, max_len, 32 - dlen - depth, depth_mask};
    // ---- decode -----------------------------------------------------------------------------------------------------
    if (both == 0) {
      // maxValue 0 and a single zero-length code: every symbol is 0 and costs no bits
      for (uint32_t i = tid; i < count; i += HUFF_THREADS) out[i] = 0;
      if (tid == 0) { U->nsym = count; if (data_at > nbits) U->status = MIC_E_BITSTREAM; }
      continue;
    }
    if (serial) {
      if (tid == 0) {
        uint32_t pos = data_at, k = 0;
        int err = 0;
        for (; k < count; k++) {
          if (pos >= nbits) { err = MIC_E_BITSTREAM; break; }
          uint32_t sym;
          const uint32_t used = D.step(pos, sym);
          if (!used) { err = MIC_E_DTABLE; break; }
          out[k] = (uint16_t)sym;
          pos += used;
        }
        if (!err && (unsigned long long)pos + both > nbits) err = MIC_E_BITSTREAM;   // Go has read maxCodeLength + pixelDepth bits ahead
        U->nsym = count;
        if (err) U->status = err;
      }
      continue;
    }
    uint32_t tile_start = data_at, sym_base = 0;
    while (sym_base < count && tile_start < nbits) {
      const uint32_t my_lo = tile_start + (uint32_t)tid * HUFF_SUB;        // < 2^31 + 2^17
      const uint32_t my_hi = min(my_lo + (uint32_t)HUFF_SUB, nbits);
      uint32_t start = my_lo, cnt = 0, endp = my_lo;
      bool dirty = true;
      // rounds of speculative decoding until every subsequence starts where its predecessor ended
      for (int round = 0; round <= HUFF_THREADS; round++) {
        if (dirty) {
          uint32_t pos = start;
          cnt = 0;
          while (pos < my_hi) {
            uint32_t sym;
            const uint32_t used = D.step(pos, sym);
            pos += used ? used : 1u;   // an uncovered slot met from a guessed start must not stall the guess (see the writing pass)
            cnt++;
          }
          endp = pos;
        }
        s_end[tid] = max(endp, min(my_lo + (uint32_t)HUFF_SUB, nbits));    // an empty subsequence hands its start on
        __syncthreads();
        const uint32_t want = tid ? s_end[tid - 1] : tile_start;
        dirty = want != start;
        start = want;
        if (!__syncthreads_or(dirty ? 1 : 0)) break;
      }
      if (s_bad) break;
      // exclusive scan of the symbol counts -> output offsets
      uint32_t inc = cnt;
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) {
        const uint32_t t = __shfl_up_sync(0xFFFFFFFFu, inc, d);
        if (lane >= d) inc += t;
      }
      if (lane == 31) s_scan[warp] = inc;
      __syncthreads();
      uint32_t off = inc - cnt, total = 0;
#pragma unroll
      for (int wv = 0; wv < HUFF_THREADS / 32; wv++) {
        const uint32_t x = s_scan[wv];
        if (wv < warp) off += x;
        total += x;
      }
      // the writing pass
      {
        uint32_t pos = start, k = sym_base + off;
        while (pos < my_hi && k < count) {
          uint32_t sym;
          const uint32_t used = D.step(pos, sym);
          if (!used) s_bad = 2;              // the true chain met an uncovered slot (Go would emit zeros for ever without moving)
          out[k] = (uint16_t)sym;
          pos += used ? used : 1u;
          if (++k == count) s_final = pos;   // where the stream stands after its last symbol
        }
      }
      const uint32_t next_start = s_end[HUFF_THREADS - 1];
      __syncthreads();   // s_end / s_scan are rewritten by the next tile
      sym_base += total;
      tile_start = next_start;
    }
    __syncthreads();
    if (tid == 0) {
      U->nsym = count;
      if (s_bad) U->status = MIC_E_DTABLE;
      else if (count && (s_final == 0xFFFFFFFFu || (unsigned long long)s_final + both > nbits)) U->status = MIC_E_BITSTREAM;
      else if (!count && (unsigned long long)data_at + both > nbits) U->status = MIC_E_BITSTREAM;
    }
  }
}

void launch_huff_decode(MicUnit* d_units, const int* d_list, int nlist, const uint8_t* d_comp, uint32_t* d_tabA, uint16_t* d_tabS,
                        uint16_t* d_states, int serial, int sm_count, cudaStream_t st) {
  if (nlist <= 0) return;
  const int grid = nlist < sm_count * 4 ? nlist : sm_count * 4;
  k_huff_decode<<<grid, HUFF_THREADS, 0, st>>>(d_units, d_list, nlist, d_comp, d_tabA, d_tabS, d_states, serial);
}

}  // namespace micgpu

// ---- encode direction (CanHuffmanCompressU16.Compress, canhuffmancompressu16.go:52-81) ------------------------------
// The code itself (frequencies -> code lengths -> canonical codes) is a few thousand operations on at most 65536 list
// entries and stays on the host (micgpu_huff_host.cu); what scales with the input is the histogram and the bit
// emission.  enc[s] = code | codeLen << 24 | delimiter << 31 (GenerateAllSymbolTable, :83-106); a symbol outside the list
// is the delimiter's code followed by the symbol in pixelDepth bits, i.e. one field of codeLen + pixelDepth <= 32 bits.
namespace micgpu {

namespace {
constexpr int HENC_THREADS = 256;
constexpr int HENC_PER_THREAD = 16;
constexpr int HENC_CHUNK = HENC_THREADS * HENC_PER_THREAD;

__device__ __forceinline__ void huff_field(uint32_t e, uint32_t sym, int depth, uint32_t& v, uint32_t& nb) {
  nb = (e >> 24) & 0x7Fu;
  v = e & 0xFFFFFFu;
  if (e >> 31) {
    v = depth ? (v << depth) | sym : v;
    nb += (uint32_t)depth;
  }
}
}  // namespace

__global__ void __launch_bounds__(256) k_huff_hist(const uint16_t* __restrict__ sym, unsigned long long n, uint32_t* __restrict__ hist) {
  for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (unsigned long long)gridDim.x * blockDim.x)
    atomicAdd(&hist[sym[i]], 1u);
}

// bits of every chunk of HENC_CHUNK symbols
__global__ void __launch_bounds__(HENC_THREADS)
k_huff_chunk_bits(const uint16_t* __restrict__ sym, unsigned long long n, const uint32_t* __restrict__ enc, int depth,
                  unsigned long long* __restrict__ chunk_bits) {
  __shared__ uint32_t s_w[HENC_THREADS / 32];
  const unsigned long long base = (unsigned long long)blockIdx.x * HENC_CHUNK;
  uint32_t bits = 0;
  for (int k = 0; k < HENC_PER_THREAD; k++) {
    const unsigned long long i = base + (unsigned long long)k * HENC_THREADS + threadIdx.x;
    if (i < n) {
      uint32_t v, nb;
      huff_field(enc[sym[i]], sym[i], depth, v, nb);
      bits += nb;
    }
  }
  for (int d = 16; d; d >>= 1) bits += __shfl_xor_sync(0xFFFFFFFFu, bits, d);
  if ((threadIdx.x & 31) == 0) s_w[threadIdx.x >> 5] = bits;
  __syncthreads();
  if (threadIdx.x == 0) {
    uint32_t t = 0;
    for (int w = 0; w < HENC_THREADS / 32; w++) t += s_w[w];
    chunk_bits[blockIdx.x] = t;
  }
}

// exclusive scan of the chunk totals in place (one CTA; a 5 M-symbol stream has 1228 chunks); total[0] = sum
__global__ void __launch_bounds__(256) k_huff_scan(unsigned long long* __restrict__ chunk_bits, int nchunks, unsigned long long* __restrict__ total) {
  __shared__ unsigned long long s_part[256];
  const int per = (nchunks + 255) / 256;
  const int lo = threadIdx.x * per, hi = min(lo + per, nchunks);
  unsigned long long sum = 0;
  for (int i = lo; i < hi; i++) sum += chunk_bits[i];
  s_part[threadIdx.x] = sum;
  __syncthreads();
  unsigned long long before = 0;
  for (int t = 0; t < (int)threadIdx.x; t++) before += s_part[t];
  for (int i = lo; i < hi; i++) {
    const unsigned long long v = chunk_bits[i];
    chunk_bits[i] = before;
    before += v;
  }
  if (threadIdx.x == 255) total[0] = before;
}

// every symbol's field to its place in the MSB-first bit string that starts `first_bit` bits into out (zeroed but for the
// header the host wrote).  Thread t of a chunk owns symbols [16 t, 16 t + 16): consecutive fields, so whole words are
// assembled in a register; the first and last word of a thread's run are shared with its neighbours, hence atomicOr.
__global__ void __launch_bounds__(HENC_THREADS)
k_huff_emit(const uint16_t* __restrict__ sym, unsigned long long n, const uint32_t* __restrict__ enc, int depth,
            const unsigned long long* __restrict__ chunk_off, unsigned long long first_bit, uint32_t* __restrict__ out) {
  __shared__ uint32_t s_w[HENC_THREADS / 32];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const unsigned long long base = (unsigned long long)blockIdx.x * HENC_CHUNK + (unsigned long long)tid * HENC_PER_THREAD;
  uint32_t v[HENC_PER_THREAD], nb[HENC_PER_THREAD], mine = 0;
#pragma unroll
  for (int k = 0; k < HENC_PER_THREAD; k++) {
    v[k] = 0; nb[k] = 0;
    if (base + k < n) huff_field(enc[sym[base + k]], sym[base + k], depth, v[k], nb[k]);
    mine += nb[k];
  }
  uint32_t inc = mine;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    const uint32_t t = __shfl_up_sync(0xFFFFFFFFu, inc, d);
    if (lane >= d) inc += t;
  }
  if (lane == 31) s_w[warp] = inc;
  __syncthreads();
  uint32_t before = inc - mine;
  for (int w = 0; w < warp; w++) before += s_w[w];
  unsigned long long p = first_bit + chunk_off[blockIdx.x] + before;     // bit position of this thread's first field
  // 64-bit accumulator holding the bits of word `w` (top-aligned) and what spills into the next one
  unsigned long long w = p >> 5;
  uint32_t fill = (uint32_t)(p & 31);          // bits of word w that belong to earlier threads (or the header)
  unsigned long long acc = 0;                  // bits [fill, ...) of word w, top-aligned at bit 63 - fill
  auto flush = [&](uint32_t word) {            // word w is complete (or the run ends): big-endian bytes in memory
    atomicOr(&out[w], __byte_perm(word, 0, 0x0123));
  };
#pragma unroll
  for (int k = 0; k < HENC_PER_THREAD; k++) {
    if (nb[k]) {
      acc |= (unsigned long long)v[k] << (64 - fill - nb[k]);      // fill < 32, nb <= 32
      fill += nb[k];
      if (fill >= 32) {
        flush((uint32_t)(acc >> 32));
        acc <<= 32;
        fill -= 32;
        w++;
      }
    }
  }
  if (fill && (uint32_t)(acc >> 32)) flush((uint32_t)(acc >> 32));
}

void launch_huff_hist(const uint16_t* d_sym, unsigned long long n, uint32_t* d_hist, int sm_count, cudaStream_t st) {
  if (!n) return;
  const unsigned long long want = (n + 255) / 256;
  const int grid = (int)(want < (unsigned long long)sm_count * 8 ? want : (unsigned long long)sm_count * 8);
  k_huff_hist<<<grid, 256, 0, st>>>(d_sym, n, d_hist);
}

int huff_enc_chunks(unsigned long long n) { return (int)((n + HENC_CHUNK - 1) / HENC_CHUNK); }

void launch_huff_emit(const uint16_t* d_sym, unsigned long long n, const uint32_t* d_enc, int depth, unsigned long long* d_chunk,
                      unsigned long long* d_total, unsigned long long first_bit, uint32_t* d_out, int phase, cudaStream_t st) {
  const int nchunks = huff_enc_chunks(n);
  if (!nchunks) return;
  if (phase == 0) {
    k_huff_chunk_bits<<<nchunks, HENC_THREADS, 0, st>>>(d_sym, n, d_enc, depth, d_chunk);
    k_huff_scan<<<1, 256, 0, st>>>(d_chunk, nchunks, d_total);
  } else {
    k_huff_emit<<<nchunks, HENC_THREADS, 0, st>>>(d_sym, n, d_enc, depth, d_chunk, first_bit, d_out);
  }
}

}  // namespace micgpu
