// k_table.cu -- K1: ncount header parse + tANS / rANS decode-table build.
//
// Replaces readNCount (fsedecompressu16.go:48-167), buildDtable (:198-263) and
// buildRansDecTable (ransu16.go:77-135).  One persistent CTA per unit slot:
// warp 0 runs the (inherently serial) ncount bit parse out of a shared-memory
// window; the spread, the low-probability cells and the per-symbol state
// ranking are data-parallel:
//   * spread position j*step mod 2^L is a permutation, so cell p = pos(j)
//     receives the symbol whose cumulative range holds rank(j) = #live j' < j;
//   * nextState[u] = norm[sym] + #cells u' < u with the same symbol, computed
//     32 cells at a time with __match_any_sync.
// Output per unit: tabA[u] = newState | nbBits<<16 and tabS[u] = symbol.
#include <algorithm>

#include "mic_device.cuh"

namespace micgpu {

constexpr int K1_THREADS = 128;
constexpr unsigned K1_FALLBACK = 0xFFFFFFFFu;   // MicUnit::npres of a unit the split path leaves to k_build_tables
constexpr int K1_WIN = 8192;   // bytes of ncount header staged in shared memory (8 KB keeps 16 CTAs per SM resident)

struct BlockScan {
  // exclusive scan of one u32 per thread over K1_THREADS threads
  __device__ static unsigned run(unsigned v, unsigned* smem4, unsigned* total) {
    const unsigned lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    unsigned inc = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      unsigned t = __shfl_up_sync(0xffffffffu, inc, d);
      if (lane >= (unsigned)d) inc += t;
    }
    __syncthreads();
    if (lane == 31) smem4[warp] = inc;
    __syncthreads();
    unsigned base = 0, tot = 0;
#pragma unroll
    for (int w = 0; w < K1_THREADS / 32; w++) {
      unsigned s = smem4[w];
      if ((unsigned)w < warp) base += s;
      tot += s;
    }
    *total = tot;
    return base + inc - v;
  }
};

// Serial ncount parse, executed redundantly by every lane of warp 0 (uniform
// control flow; lane 0 writes).  Follows readNCount line by line; `b.off`,
// `iend`, `bitStream`, `bitCount` keep their Go meaning.
struct NCountReader {
  const uint8_t* g;   // global pointer to the ncount header (frame + prefix)
  int blen;           // bytes available (iend)
  uint8_t* win;       // shared window
  int win_base;       // first byte held in win (or -1)

  __device__ void refill(int off) {
    __syncwarp();
    win_base = off;
    const int lane = threadIdx.x & 31;
    int n = blen - off;
    if (n > K1_WIN) n = K1_WIN;
    for (int i = lane; i < n; i += 32) win[i] = g[off + i];
    __syncwarp();
  }
  __device__ uint32_t u32(int off) {
    if (win_base < 0 || off < win_base || off + 4 > win_base + K1_WIN) refill(off);
    const uint8_t* p = win + (off - win_base);
    return (uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16) | ((uint32_t)p[3] << 24);
  }
};

__global__ void __launch_bounds__(K1_THREADS)
k_build_tables(MicUnit* __restrict__ units, int nunits, const uint8_t* __restrict__ comp,
               uint32_t* __restrict__ tabA, uint16_t* __restrict__ tabS,
               uint8_t* __restrict__ scratch, unsigned long long scratch_stride, int max_log, int only_flagged) {
  __shared__ __align__(16) uint8_t s_win[K1_WIN];
  __shared__ unsigned s_scan[K1_THREADS / 32];
  __shared__ int s_status, s_symlen, s_consumed, s_tlog;

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  // Only symbols with a non-zero normalised count matter and there are at most 2^L of them, so
  // all scratch is indexed by "present index" and sized by the table, not by the 65536-symbol alphabet.
  const unsigned cap = 1u << max_log;                       // entries reserved per array
  uint8_t* my = scratch + (unsigned long long)blockIdx.x * scratch_stride;
  int32_t* norm = reinterpret_cast<int32_t*>(my);           // [cap] normalised count of present symbol i
  uint32_t* cumul = reinterpret_cast<uint32_t*>(norm + cap);   // [cap]
  uint32_t* symnext = cumul + cap;                          // [cap]
  uint16_t* pres_sym = reinterpret_cast<uint16_t*>(symnext + cap);  // [cap] symbol value of present symbol i
  uint16_t* sym_of_rank = pres_sym + cap;                   // [cap] present index owning spread rank k
  uint16_t* cell_idx = sym_of_rank + cap;                   // [cap] present index of table cell u

  for (int ui = blockIdx.x; ui < nunits; ui += gridDim.x) {
    MicUnit* U = &units[ui];
    __syncthreads();
    if (U->status != MIC_OK || U->rans == MIC_CODER_HUFF) continue;   // Huffman units bring no ncount header (k_huff.cu)
    if (only_flagged && U->npres != K1_FALLBACK) continue;   // the split kernels below already built this unit's tables
    const uint8_t* frame = comp + U->comp_off;
    const int flen = (int)U->comp_len;
    const int hdr = U->nstates > 1 ? 6 : 0;

    // ---------------- phase 1: serial ncount parse (warp 0) ----------------
    if (warp == 0) {
      int status = MIC_OK;
      int consumed = 0;
      uint32_t charnum = 0, npres = 0;
      int tlog = 0;
      NCountReader rd{frame + hdr, flen - hdr, s_win, -1};
      const int iend = rd.blen;
      if (iend < 4) status = MIC_E_NCOUNT;
      if (status == MIC_OK) {
        int off = 0;
        uint32_t bit_stream = rd.u32(off);
        uint32_t nb_bits = (bit_stream & 0xF) + 5;  // minTablelog
        if (nb_bits > 16 || (int)nb_bits > max_log) status = MIC_E_NCOUNT;  // containers emit <= maxTableLog (fseu16.go:24)
        bit_stream >>= 4;
        uint32_t bit_count = 4;
        tlog = (int)nb_bits;
        int32_t remaining = (int32_t)((1u << nb_bits) + 1);
        int32_t threshold = (int32_t)(1u << nb_bits);
        int32_t got_total = 0;
        nb_bits++;
        bool previous0 = false;
        while (status == MIC_OK && remaining > 1) {
          if (previous0) {
            uint32_t n0 = charnum;
            while ((bit_stream & 0xFFFF) == 0xFFFF) {
              n0 += 24;
              if (off < iend - 5) {
                off += 2;
                bit_stream = rd.u32(off) >> bit_count;
              } else {
                bit_stream >>= 16;
                bit_count += 16;
              }
              if (n0 > 65535u + 24u) break;
            }
            while ((bit_stream & 3) == 3) {
              n0 += 3;
              bit_stream >>= 2;
              bit_count += 2;
            }
            n0 += bit_stream & 3;
            bit_count += 2;
            if (n0 > 65535u) { status = MIC_E_NCOUNT; break; }
            charnum = n0;
            if (off <= iend - 7 || off + (int)(bit_count >> 3) <= iend - 4) {
              off += (int)(bit_count >> 3);
              bit_count &= 7;
              bit_stream = rd.u32(off) >> bit_count;
            } else {
              bit_stream >>= 2;
            }
          }
          const int32_t max = (2 * threshold - 1) - remaining;
          int32_t count;
          if (((int32_t)bit_stream & (threshold - 1)) < max) {
            count = (int32_t)bit_stream & (threshold - 1);
            bit_count += nb_bits - 1;
          } else {
            count = (int32_t)bit_stream & (2 * threshold - 1);
            if (count >= threshold) count -= max;
            bit_count += nb_bits;
          }
          count--;
          if (count < 0) { remaining += count; got_total -= count; }
          else { remaining -= count; got_total += count; }
          if (charnum > 65535u) { status = MIC_E_NCOUNT; break; }
          if (count != 0) {
            if (npres >= cap) { status = MIC_E_NCOUNT; break; }  // more present symbols than table cells
            if (lane == 0) { norm[npres] = count; pres_sym[npres] = (uint16_t)charnum; }
            npres++;
          }
          charnum++;
          previous0 = (count == 0);
          while (remaining < threshold) { nb_bits--; threshold >>= 1; }
          if (off <= iend - 7 || off + (int)(bit_count >> 3) <= iend - 4) {
            off += (int)(bit_count >> 3);
            bit_count &= 7;
          } else {
            bit_count -= (uint32_t)(8 * (iend - 4 - off));
            off = iend - 4;
          }
          if (off < 0 || off + 4 > iend) { status = MIC_E_NCOUNT; break; }
          bit_stream = rd.u32(off) >> (bit_count & 31);
        }
        if (status == MIC_OK) {
          if (charnum <= 1 || remaining != 1 || bit_count > 32 || got_total != (1 << tlog)) status = MIC_E_NCOUNT;
          consumed = off + (int)((bit_count + 7) >> 3);
          if (hdr + consumed >= flen) status = MIC_E_BITSTREAM;  // bitReader.init: "too short"
        }
      }
      if (lane == 0) { s_status = status; s_symlen = (int)npres; s_consumed = consumed; s_tlog = tlog; }
    }
    __syncthreads();
    if (s_status != MIC_OK) {
      if (tid == 0) U->status = s_status;
      continue;
    }
    const int L = s_tlog;
    const uint32_t S = 1u << L;
    const uint32_t symlen = (uint32_t)s_symlen;   // number of present symbols
    uint32_t* A = tabA + U->tab_off;
    uint16_t* Sy = tabS + U->tab_off;
    const bool rans = U->rans != 0;

    // ---------------- phase 2: cumul, low-prob cells, symbol-of-rank -------
    // thread t owns symbols [t*per, (t+1)*per)
    const uint32_t per = (symlen + K1_THREADS - 1) / K1_THREADS;
    const uint32_t s0 = min(symlen, (uint32_t)tid * per), s1 = min(symlen, s0 + per);
    unsigned my_sum = 0, my_low = 0;
    for (uint32_t s = s0; s < s1; s++) {
      int32_t v = norm[s];
      if (v > 0) my_sum += (unsigned)v;
      else if (v == -1) my_low++;
    }
    unsigned tot_sum, tot_low;
    unsigned base_sum = BlockScan::run(my_sum, s_scan, &tot_sum);
    unsigned base_low = BlockScan::run(my_low, s_scan, &tot_low);
    const uint32_t nlow = tot_low;
    const uint32_t high_threshold = S - 1 - nlow;  // cells above it are low-prob cells
    {
      unsigned c = base_sum, lr = base_low;
      for (uint32_t s = s0; s < s1; s++) {
        int32_t v = norm[s];
        cumul[s] = c;
        if (v > 0) {
          symnext[s] = (uint32_t)v;
          for (int32_t i = 0; i < v; i++) sym_of_rank[c + i] = (uint16_t)s;
          c += (unsigned)v;
        } else if (v == -1) {
          symnext[s] = 1;
          // tANS: lowprob symbols laid down from the top in symbol order (fsedecompressu16.go:207-211)
          // rANS: after all normal symbols, ascending (ransu16.go:116-129)
          uint32_t cell = rans ? (S - nlow + lr) : (S - 1 - lr);
          cell_idx[cell] = (uint16_t)s;
          if (rans) { A[cell] = ((uint32_t)L << 16); Sy[cell] = pres_sym[s]; }  // newState 0, nbBits = tableLog
          lr++;
        } else {
          symnext[s] = 0;
        }
      }
    }
    __syncthreads();

    if (rans) {
      // linear fill: slot k in [0, S-nlow) belongs to sym_of_rank[k]
      for (uint32_t k = tid; k < S - nlow; k += K1_THREADS) {
        uint32_t s = sym_of_rank[k];
        uint32_t x_next = (uint32_t)norm[s] + (k - cumul[s]);
        uint32_t nb = (uint32_t)L - (31u - __clz(x_next));
        uint32_t ns = (x_next << nb) - S;
        Sy[k] = pres_sym[s];
        A[k] = ns | (nb << 16);
      }
    } else {
      // ---------------- phase 3: spread (tANS) ------------------------------
      const uint32_t step = (S >> 1) + (S >> 3) + 3, mask = S - 1;
      const uint32_t jper = S / K1_THREADS ? S / K1_THREADS : 1;  // S >= 32; K1_THREADS=128 may exceed S
      const uint32_t j0 = min(S, (uint32_t)tid * jper), j1 = min(S, j0 + jper);
      unsigned live = 0;
      if (nlow) {
        for (uint32_t j = j0; j < j1; j++) live += (((j * step) & mask) <= high_threshold);
      } else {
        live = j1 - j0;
      }
      unsigned tot_live;
      unsigned rank = BlockScan::run(live, s_scan, &tot_live);
      for (uint32_t j = j0; j < j1; j++) {
        uint32_t pos = (j * step) & mask;
        if (pos <= high_threshold) cell_idx[pos] = sym_of_rank[rank++];
      }
      __syncthreads();
      // ---------------- phase 4: nextState ranking (warp 0) -----------------
      if (warp == 0) {
        int bad = 0;
        for (uint32_t u0 = 0; u0 < S; u0 += 32) {
          uint32_t u = u0 + lane;
          uint32_t sym = cell_idx[u];
          unsigned peers = __match_any_sync(0xffffffffu, sym);
          unsigned r = __popc(peers & ((1u << lane) - 1u));
          uint32_t basev = symnext[sym];
          __syncwarp();
          if (r == 0) symnext[sym] = basev + __popc(peers);
          __syncwarp();
          uint32_t next_state = basev + r;
          if (next_state == 0) { bad = 1; next_state = 1; }
          uint32_t nb = (uint32_t)L - (31u - __clz(next_state));
          uint32_t ns = (next_state << nb) - S;
          if (next_state >= 2 * S || ns >= S || (ns == u && nb == 0)) bad = 1;  // fsedecompressu16.go:252-258
          A[u] = (ns & 0xFFFF) | (nb << 16);
          Sy[u] = pres_sym[sym];
        }
        bad = __any_sync(0xffffffffu, bad);
        if (bad && lane == 0) s_status = MIC_E_DTABLE;
      }
    }
    __syncthreads();
    if (tid == 0) {
      int st = s_status;
      unsigned boff = (unsigned)(hdr + s_consumed);
      unsigned blen = (unsigned)flen - boff;
      if (st == MIC_OK && frame[flen - 1] == 0) st = MIC_E_BITSTREAM;  // bitreader.go:33-36
      U->bits_off = boff;
      U->bits_len = blen;
      U->table_log = (unsigned)L;
      U->status = st;
    }
  }
}

// ================================================================================================================
// Split path (the common case: tANS, tableLog <= 13, at most K1B_NPMAX present symbols).
//
// The single kernel above keeps a unit's CTA busy for the whole serial header parse (~350 cycles per symbol: byte
// loads from shared memory, loops for the zero runs and the threshold) and then ranks the table cells with ONE warp
// through global scratch.  All units must pass in one wave or the parse latency doubles, which pins 16 small CTAs
// on every SM.  Here the two halves have the shape that suits them:
//   K1a k_parse_ncount : one WARP per unit, every unit of the batch resident at once.  The parse is a chain, so what
//       matters is its length: a three-word register window over the header (refilled one word ahead, so no load
//       sits on the chain), the threshold loop as one find-leading-one, no byte assembly.  Near the end of a frame
//       the reference reader changes behaviour (fsedecompressu16.go:96-104,141-150); the fast loop stops 16 bytes
//       short of that and hands its state to the exact byte-wise reader.
//   K1b k_build_dtable : one CTA per unit, everything in shared memory, a handful of units per SM at a time.  The
//       spread is a permutation (cell j*step mod S gets the symbol owning live rank(j), found by binary search in the
//       cumulative counts, then walked); nextState = norm[sym] + #earlier cells of sym is computed by four warps
//       over quarter tables with __match_any_sync and per-warp counters, stitched by a prefix over the four.
// Units the split path cannot take (rANS, tableLog > 13, more present symbols than K1B_NPMAX) are flagged and built
// by k_build_tables(only_flagged).
constexpr int K1A_WIN_WORDS = 512;      // header window per warp (2 KB)
constexpr int K1A_WARPS = 4;
constexpr int K1B_THREADS = 256;
constexpr int K1B_NPMAX = 2048;
constexpr int K1B_RANKW = 4;            // warps that rank cells (each owns a quarter of the table)
constexpr int K1B_MAXLOG = 13;

__global__ void __launch_bounds__(32 * K1A_WARPS)
k_parse_ncount(MicUnit* __restrict__ units, int nunits, const uint8_t* __restrict__ comp, int32_t* __restrict__ g_norm,
               uint16_t* __restrict__ g_sym) {
  __shared__ __align__(16) uint32_t s_win_all[K1A_WARPS][K1A_WIN_WORDS];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int ui = blockIdx.x * K1A_WARPS + warp;
  if (ui >= nunits) return;
  MicUnit* U = &units[ui];
  if (U->status != MIC_OK || U->rans == MIC_CODER_HUFF) return;
  uint32_t* win = s_win_all[warp];
  const uint8_t* frame = comp + U->comp_off;
  const int flen = (int)U->comp_len;
  const int hdr = U->nstates > 1 ? 6 : 0;
  const uint8_t* g = frame + hdr;
  const int iend = flen - hdr;
  int32_t* norm = g_norm + U->tab_off;
  uint16_t* pres_sym = g_sym + U->tab_off;
  const uint32_t cap = 1u << U->table_log;      // entries reserved for this unit (table_log peeked by the host)

  int status = MIC_OK;
  int consumed = 0;
  uint32_t charnum = 0, npres = 0;
  int tlog = 0;
  if (iend < 4) status = MIC_E_NCOUNT;
  if (status == MIC_OK) {
    // word-aligned view of the header: byte b of the header is byte (b + a0) of the aligned stream
    const uintptr_t ga = reinterpret_cast<uintptr_t>(g);
    const uint32_t a0 = (uint32_t)(ga & 3u);
    const uint32_t* gw = reinterpret_cast<const uint32_t*>(ga - a0);
    const int nwords = (int)((a0 + (uint32_t)iend + 3u) >> 2);    // words that hold header/frame bytes
    int win_base = 0;                                             // first word held in win
    auto refill = [&](int w0) {
      __syncwarp();
      win_base = w0;
      for (int i = lane; i < K1A_WIN_WORDS; i += 32) win[i] = (w0 + i < nwords) ? __ldg(gw + w0 + i) : 0u;
      __syncwarp();
    };
    refill(0);
    const uint32_t first = __funnelshift_r(win[0], win[1], 8u * a0);
    uint32_t nb_bits = (first & 0xF) + 5;        // minTablelog
    if (nb_bits > 16 || nb_bits != U->table_log) status = MIC_E_NCOUNT;
    tlog = (int)nb_bits;
    int32_t remaining = (int32_t)((1u << nb_bits) + 1);
    int32_t threshold = (int32_t)(1u << nb_bits);
    int32_t got_total = 0;
    nb_bits++;
    bool previous0 = false;
    uint32_t bp = 8u * a0 + 4u;                  // absolute bit position in the aligned stream
    // ---- fast loop: pure sequential bit reader, valid while the reference reader is in its "far from the end" regime
    // and the header stays inside the first window (no refill in here: everything lives in registers; headers longer
    // than 2 KB finish in the exact reader below).  One symbol advances at most one word (a count field has <= 17 bits,
    // a zero-run step 16), so the window moves with three predicated moves and one shared-memory load per word.
    {
      // first bit the fast loop must not start a symbol at: the reference switches regime at off > iend - 7, and the
      // window holds K1A_WIN_WORDS words (two of them are look-ahead)
      long long lim = ((long long)iend - 7 - 16) * 8 + 8 * (long long)a0;
      lim = min(lim, (long long)(K1A_WIN_WORDS - 6) * 32);
      const uint32_t limit = lim < 0 ? 0u : (uint32_t)lim;
      uint32_t widx = bp >> 5;
      uint32_t w0 = win[widx], w1 = win[widx + 1], w2 = win[widx + 2];
#define K1A_STEP(NBITS)                                                 \
  {                                                                     \
    bp += (NBITS);                                                      \
    if ((bp >> 5) != widx) { widx++; w0 = w1; w1 = w2; w2 = win[widx + 2]; } \
  }
      while (status == MIC_OK && remaining > 1 && bp <= limit) {
        uint32_t bs = __funnelshift_r(w0, w1, bp);
        if (previous0) {
          // the zero run is only committed once it is complete: a run that would cross `limit` is redone by the exact reader
          const uint32_t bp_top = bp, widx_top = widx, s0 = w0, s1 = w1, s2 = w2;
          uint32_t n0 = charnum;
          bool bail = false;
          while ((bs & 0xFFFF) == 0xFFFF) {
            n0 += 24;
            if (bp + 16 > limit) { bail = true; break; }
            K1A_STEP(16)
            bs = __funnelshift_r(w0, w1, bp);
            if (n0 > 65535u + 24u) break;
          }
          if (bail) { bp = bp_top; widx = widx_top; w0 = s0; w1 = s1; w2 = s2; break; }
          uint32_t used = 0;
          while ((bs & 3) == 3) { n0 += 3; bs >>= 2; used += 2; }
          n0 += bs & 3;
          used += 2;
          if (n0 > 65535u) { status = MIC_E_NCOUNT; break; }
          charnum = n0;
          K1A_STEP(used)
          bs = __funnelshift_r(w0, w1, bp);
        }
        const int32_t max = (2 * threshold - 1) - remaining;
        const int32_t lowv = (int32_t)bs & (threshold - 1);
        int32_t count = (int32_t)bs & (2 * threshold - 1);
        if (count >= threshold) count -= max;
        uint32_t used = nb_bits;
        if (lowv < max) { count = lowv; used = nb_bits - 1; }
        count--;
        const int32_t mag = count < 0 ? -count : count;      // -1 means +1
        remaining -= mag;
        got_total += mag;
        if (charnum > 65535u) { status = MIC_E_NCOUNT; break; }
        if (count != 0) {
          if (npres >= cap) { status = MIC_E_NCOUNT; break; }  // more present symbols than table cells
          if (lane == 0) { norm[npres] = count; pres_sym[npres] = (uint16_t)charnum; }
          npres++;
        }
        charnum++;
        previous0 = (count == 0);
        if (remaining < threshold && remaining >= 1) {
          const uint32_t hb = 31u - (uint32_t)__clz(remaining);
          threshold = (int32_t)(1u << hb);
          nb_bits = hb + 1;
        }
        K1A_STEP(used)
      }
#undef K1A_STEP
    }
    // ---- exact reader for the tail (and for frames shorter than the margin): readNCount line by line -------------
    if (status == MIC_OK && remaining > 1) {
      // byte view of the header through the same window (refilled on demand)
      auto u32at = [&](int off) -> uint32_t {
        const uint32_t b0 = (uint32_t)off + a0;
        const int wi = (int)(b0 >> 2);
        if (wi < win_base || wi + 2 >= win_base + K1A_WIN_WORDS) refill(wi);
        return __funnelshift_r(win[wi - win_base], win[wi + 1 - win_base], 8u * (b0 & 3u));
      };
      int off = (int)((bp - 8u * a0) >> 3);
      uint32_t bit_count = (bp - 8u * a0) & 7u;
      uint32_t bit_stream = u32at(off) >> bit_count;
      while (status == MIC_OK && remaining > 1) {
        if (previous0) {
          uint32_t n0 = charnum;
          while ((bit_stream & 0xFFFF) == 0xFFFF) {
            n0 += 24;
            if (off < iend - 5) {
              off += 2;
              bit_stream = u32at(off) >> bit_count;
            } else {
              bit_stream >>= 16;
              bit_count += 16;
            }
            if (n0 > 65535u + 24u) break;
          }
          while ((bit_stream & 3) == 3) {
            n0 += 3;
            bit_stream >>= 2;
            bit_count += 2;
          }
          n0 += bit_stream & 3;
          bit_count += 2;
          if (n0 > 65535u) { status = MIC_E_NCOUNT; break; }
          charnum = n0;
          if (off <= iend - 7 || off + (int)(bit_count >> 3) <= iend - 4) {
            off += (int)(bit_count >> 3);
            bit_count &= 7;
            bit_stream = u32at(off) >> bit_count;
          } else {
            bit_stream >>= 2;
          }
        }
        const int32_t max = (2 * threshold - 1) - remaining;
        int32_t count;
        if (((int32_t)bit_stream & (threshold - 1)) < max) {
          count = (int32_t)bit_stream & (threshold - 1);
          bit_count += nb_bits - 1;
        } else {
          count = (int32_t)bit_stream & (2 * threshold - 1);
          if (count >= threshold) count -= max;
          bit_count += nb_bits;
        }
        count--;
        if (count < 0) { remaining += count; got_total -= count; }
        else { remaining -= count; got_total += count; }
        if (charnum > 65535u) { status = MIC_E_NCOUNT; break; }
        if (count != 0) {
          if (npres >= cap) { status = MIC_E_NCOUNT; break; }
          if (lane == 0) { norm[npres] = count; pres_sym[npres] = (uint16_t)charnum; }
          npres++;
        }
        charnum++;
        previous0 = (count == 0);
        while (remaining < threshold) { nb_bits--; threshold >>= 1; }
        if (off <= iend - 7 || off + (int)(bit_count >> 3) <= iend - 4) {
          off += (int)(bit_count >> 3);
          bit_count &= 7;
        } else {
          bit_count -= (uint32_t)(8 * (iend - 4 - off));
          off = iend - 4;
        }
        if (off < 0 || off + 4 > iend) { status = MIC_E_NCOUNT; break; }
        bit_stream = u32at(off) >> (bit_count & 31);
      }
      if (status == MIC_OK) {
        if (bit_count > 32) status = MIC_E_NCOUNT;
        consumed = off + (int)((bit_count + 7) >> 3);
      }
    } else if (status == MIC_OK) {
      // the fast loop finished the header: off = bp / 8 and bitCount = bp % 8 in the reference's terms
      const uint32_t hb = bp - 8u * a0;
      consumed = (int)((hb + 7) >> 3);
    }
    if (status == MIC_OK) {
      if (charnum <= 1 || remaining != 1 || got_total != (1 << tlog)) status = MIC_E_NCOUNT;
      if (hdr + consumed >= flen) status = MIC_E_BITSTREAM;  // bitReader.init: "too short"
    }
  }
  if (lane == 0) {
    if (status == MIC_OK && frame[flen - 1] == 0) status = MIC_E_BITSTREAM;  // bitreader.go:33-36
    const unsigned boff = (unsigned)(hdr + consumed);
    U->bits_off = boff;
    U->bits_len = (unsigned)flen - boff;
    U->status = status;
    // hand-over to K1b; units it cannot take are built by the one-kernel path
    const bool split_ok = U->rans == 0 && tlog <= K1B_MAXLOG && npres <= (uint32_t)K1B_NPMAX;
    U->npres = split_ok ? npres : K1_FALLBACK;
  }
}

__global__ void __launch_bounds__(K1B_THREADS)
k_build_dtable(MicUnit* __restrict__ units, int nunits, const int32_t* __restrict__ g_norm, const uint16_t* __restrict__ g_sym,
               uint32_t* __restrict__ tabA, uint16_t* __restrict__ tabS, int max_log, unsigned int* __restrict__ queue) {
  extern __shared__ __align__(16) uint8_t k1b_smem[];
  const uint32_t SMAX = 1u << max_log;
  const uint32_t NP = min((uint32_t)K1B_NPMAX, SMAX);
  uint16_t* s_cell = reinterpret_cast<uint16_t*>(k1b_smem);                 // [SMAX] present index of table cell u
  uint16_t* s_lrank = s_cell + SMAX;                                        // [SMAX] rank of the cell inside its quarter
  int32_t* s_norm = reinterpret_cast<int32_t*>(s_lrank + SMAX);             // [NP]
  uint32_t* s_cumul = reinterpret_cast<uint32_t*>(s_norm + NP);             // [NP + 1]
  uint16_t* s_sym = reinterpret_cast<uint16_t*>(s_cumul + NP + 4);          // [NP]
  uint16_t* s_cnt = s_sym + NP;                                             // [K1B_RANKW][NP]
  __shared__ unsigned s_scan[K1B_THREADS / 32];
  __shared__ int s_ui, s_bad;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

  auto block_scan = [&](unsigned v, unsigned* total) -> unsigned {   // exclusive scan over the CTA
    unsigned inc = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const unsigned t = __shfl_up_sync(0xffffffffu, inc, d);
      if (lane >= d) inc += t;
    }
    __syncthreads();
    if (lane == 31) s_scan[warp] = inc;
    __syncthreads();
    unsigned base = 0, tot = 0;
#pragma unroll
    for (int w = 0; w < K1B_THREADS / 32; w++) {
      const unsigned x = s_scan[w];
      if (w < warp) base += x;
      tot += x;
    }
    *total = tot;
    return base + inc - v;
  };

  while (true) {
    __syncthreads();
    if (tid == 0) { s_ui = (int)atomicAdd(queue, 1u); s_bad = 0; }
    __syncthreads();
    const int ui = s_ui;
    if (ui >= nunits) break;
    MicUnit* U = &units[ui];
    if (U->status != MIC_OK || U->npres == K1_FALLBACK || U->rans == MIC_CODER_HUFF) continue;
    const uint32_t L = U->table_log, S = 1u << L, np = U->npres;
    const int32_t* gn = g_norm + U->tab_off;
    const uint16_t* gs = g_sym + U->tab_off;
    // ---- lists into shared memory; cumulative counts of the normal symbols, ranks of the low-probability ones ----
    const uint32_t per = (np + K1B_THREADS - 1) / K1B_THREADS;
    const uint32_t s0 = min(np, (uint32_t)tid * per), s1 = min(np, s0 + per);
    unsigned my_sum = 0, my_low = 0;
    for (uint32_t i = s0; i < s1; i++) {
      const int32_t v = gn[i];
      s_norm[i] = v;
      s_sym[i] = gs[i];
      if (v > 0) my_sum += (unsigned)v;
      else if (v == -1) my_low++;
    }
    for (uint32_t i = tid; i < K1B_RANKW * NP / 2; i += K1B_THREADS) reinterpret_cast<uint32_t*>(s_cnt)[i] = 0;
    unsigned tot_sum, nlow;
    unsigned c = block_scan(my_sum, &tot_sum);
    unsigned lr = block_scan(my_low, &nlow);
    if (tot_sum + nlow != S) {            // cannot happen after a successful parse; keeps every index below in range
      if (tid == 0) U->status = MIC_E_DTABLE;
      continue;
    }
    const uint32_t high_threshold = S - 1 - nlow;   // cells above it are the low-probability cells (fsedecompressu16.go:207-211)
    for (uint32_t i = s0; i < s1; i++) {
      const int32_t v = s_norm[i];
      s_cumul[i] = c;
      if (v > 0) c += (unsigned)v;
      else if (v == -1) { s_cell[S - 1 - lr] = (uint16_t)i; lr++; }
    }
    if (tid == 0) s_cumul[np] = tot_sum;
    __syncthreads();
    // ---- spread: cell (j * step) mod S of live j gets the symbol that owns rank(j) = #live j' < j ----------------
    {
      const uint32_t step = (S >> 1) + (S >> 3) + 3, mask = S - 1;
      const uint32_t jper = (S + K1B_THREADS - 1) / K1B_THREADS;
      const uint32_t j0 = min(S, (uint32_t)tid * jper), j1 = min(S, j0 + jper);
      unsigned live = 0;
      if (nlow) {
        for (uint32_t j = j0; j < j1; j++) live += (((j * step) & mask) <= high_threshold);
      } else {
        live = j1 - j0;
      }
      unsigned tot_live;
      unsigned rank = block_scan(live, &tot_live);
      if (live) {
        // owner of `rank`: the last i with cumul[i] <= rank (zero-width entries never own a rank)
        uint32_t lo = 0, hi = np;          // invariant: cumul[lo] <= rank < cumul[hi]
        while (hi - lo > 1) {
          const uint32_t mid = (lo + hi) >> 1;
          if (s_cumul[mid] <= rank) lo = mid; else hi = mid;
        }
        uint32_t own = lo, next_c = s_cumul[own + 1];
        for (uint32_t j = j0; j < j1; j++) {
          const uint32_t pos = (j * step) & mask;
          if (pos <= high_threshold) {
            while (rank >= next_c) { own++; next_c = s_cumul[own + 1]; }
            s_cell[pos] = (uint16_t)own;
            rank++;
          }
        }
      }
    }
    __syncthreads();
    // ---- rank of every cell among the earlier cells of its symbol: four warps, a quarter table each ------------------
    const uint32_t qlen = S / K1B_RANKW;     // S >= 32: a multiple of 8 cells per warp; iterate 32 cells at a time
    if (warp < K1B_RANKW) {
      uint16_t* cnt = s_cnt + warp * NP;
      const uint32_t u0 = warp * qlen, u1 = u0 + qlen;
      for (uint32_t ub = u0; ub < u1; ub += 32) {
        const uint32_t u = ub + lane;
        const bool valid = u < u1;
        const uint32_t sym = valid ? s_cell[u] : (0xFFFF0000u | lane);   // idle lanes match nobody
        const unsigned peers = __match_any_sync(0xffffffffu, sym);
        const unsigned r = __popc(peers & ((1u << lane) - 1u));
        unsigned basev = 0;
        if (valid) basev = cnt[sym];
        __syncwarp();
        if (valid && r == 0) cnt[sym] = (uint16_t)(basev + __popc(peers));
        __syncwarp();
        if (valid) s_lrank[u] = (uint16_t)(basev + r);
      }
    }
    __syncthreads();
    // cnt[w][i] -> number of cells of symbol i in the quarters before w
    for (uint32_t i = tid; i < np; i += K1B_THREADS) {
      unsigned run = 0;
#pragma unroll
      for (int w = 0; w < K1B_RANKW; w++) {
        const unsigned x = s_cnt[w * NP + i];
        s_cnt[w * NP + i] = (uint16_t)run;
        run += x;
      }
    }
    __syncthreads();
    // ---- decode cells: nextState -> (nbBits, newState) (fsedecompressu16.go:243-258), coalesced stores -------------
    uint32_t* A = tabA + U->tab_off;
    uint16_t* Sy = tabS + U->tab_off;
    int bad = 0;
    for (uint32_t u = tid; u < S; u += K1B_THREADS) {
      const uint32_t i = s_cell[u];
      const int32_t v = s_norm[i];
      const uint32_t w = u / qlen;
      const uint32_t next_state = (v > 0 ? (uint32_t)v : 1u) + s_cnt[w * NP + i] + s_lrank[u];
      const uint32_t nb = L - (31u - __clz(next_state));
      const uint32_t ns = (next_state << nb) - S;
      if (next_state >= 2 * S || ns >= S || (ns == u && nb == 0)) bad = 1;
      A[u] = (ns & 0xFFFF) | (nb << 16);
      Sy[u] = s_sym[i];
    }
    if (bad) s_bad = 1;
    __syncthreads();
    if (tid == 0 && s_bad) U->status = MIC_E_DTABLE;
  }
}

size_t build_dtable_smem_bytes(int max_log) {
  const size_t S = (size_t)1 << max_log;
  const size_t NP = std::min<size_t>(K1B_NPMAX, S);
  return S * 2 * 2 + NP * 4 + (NP + 4) * 4 + NP * 2 + (size_t)K1B_RANKW * NP * 2;
}

void launch_build_tables(MicUnit* d_units, int nunits, const uint8_t* d_comp, uint32_t* d_tabA, uint16_t* d_tabS,
                         uint8_t* d_scratch, unsigned long long scratch_stride, int max_log, int grid, cudaStream_t st) {
  if (nunits <= 0) return;
  k_build_tables<<<grid, K1_THREADS, 0, st>>>(d_units, nunits, d_comp, d_tabA, d_tabS, d_scratch, scratch_stride, max_log, 0);
}

void launch_parse_ncount(MicUnit* d_units, int nunits, const uint8_t* d_comp, int32_t* d_norm, uint16_t* d_sym, cudaStream_t st) {
  if (nunits <= 0) return;
  k_parse_ncount<<<(nunits + K1A_WARPS - 1) / K1A_WARPS, 32 * K1A_WARPS, 0, st>>>(d_units, nunits, d_comp, d_norm, d_sym);
}

void launch_build_dtable(MicUnit* d_units, int nunits, const uint8_t* d_comp, uint32_t* d_tabA, uint16_t* d_tabS,
                         const int32_t* d_norm, const uint16_t* d_sym, unsigned int* d_queue, uint8_t* d_scratch,
                         unsigned long long scratch_stride, int max_log, int fallback_grid, int sm_count, cudaStream_t st) {
  if (nunits <= 0) return;
  cudaMemsetAsync(d_queue, 0, sizeof(unsigned int), st);
  const int blog = max_log < K1B_MAXLOG ? max_log : K1B_MAXLOG;
  const size_t smem = build_dtable_smem_bytes(blog);
  cudaFuncSetAttribute(k_build_dtable, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  const int per_sm = (int)std::max<size_t>(1, std::min<size_t>(8, (220 * 1024) / (smem + 1024)));
  const int grid = std::min(nunits, sm_count * per_sm);
  k_build_dtable<<<grid, K1B_THREADS, smem, st>>>(d_units, nunits, d_norm, d_sym, d_tabA, d_tabS, blog, d_queue);
  if (fallback_grid > 0)
    k_build_tables<<<fallback_grid, K1_THREADS, 0, st>>>(d_units, nunits, d_comp, d_tabA, d_tabS, d_scratch, scratch_stride, max_log, 1);
}

}  // namespace micgpu
