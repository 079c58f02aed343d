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
#include "mic_device.cuh"

namespace micgpu {

constexpr int K1_THREADS = 128;
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
               uint8_t* __restrict__ scratch, unsigned long long scratch_stride, int max_log) {
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
    if (U->status != MIC_OK) continue;
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

void launch_build_tables(MicUnit* d_units, int nunits, const uint8_t* d_comp, uint32_t* d_tabA, uint16_t* d_tabS,
                         uint8_t* d_scratch, unsigned long long scratch_stride, int max_log, int grid, cudaStream_t st) {
  if (nunits <= 0) return;
  k_build_tables<<<grid, K1_THREADS, 0, st>>>(d_units, nunits, d_comp, d_tabA, d_tabS, d_scratch, scratch_stride, max_log);
}

}  // namespace micgpu
