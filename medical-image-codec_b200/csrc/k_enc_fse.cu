// k_enc_fse.cu -- K6 + K7: FSE table construction and N-state tANS encoding, byte-identical to the Go encoder.
//
// K6 (k_enc_tables) replaces countSimple (fsecompressu16.go:438-462, countSimpleU16Asm), optimalTableLog
// (:480-518), normalizeCount / normalizeCount2 (:524-667), writeCount (:191-289) and buildCTable (:329-431).
// All arithmetic is the reference's (u64 fixed point 1<<62, rtbTable, first-max tie break, uint8 wrap in
// optimalTableLog).  Data-parallel parts: histogram, per-symbol probabilities with block reductions, cumulative
// sums, the spread (position j*step mod 2^L is a permutation) and the per-symbol state ranking; the ncount
// bit packing and the rare secondary normalisation are serial (one thread), as their state is carried
// symbol to symbol.
//
// K7 (k_enc_ans + k_enc_pack) replaces cStateU16.encode / compress / compress{2,4,8}State
// (fsecompressu16.go:95-187, fse2state.go:122-199, fse4state.go:100-191, fse8state.go:113-226).
// The N state chains are independent (symbol i belongs to state i mod N, walking from the last symbol to the
// first), so lane k of a warp runs chain k and records (bits, nbBits) per symbol; the bitstream is the
// concatenation of those fields in descending symbol order, i.e. a suffix sum of nbBits gives every field its bit
// position; a CTA packs them through a shared-memory word buffer and appends the final states (N-1 .. 0,
// tableLog bits each) and the 1-bit end mark (bitwriter.go:162-168).
#include <algorithm>
#include <cstdlib>

#include "mic_device.cuh"
#include "mic_enc.h"

namespace micgpu {

constexpr int T_THREADS = 256;

__device__ __forceinline__ unsigned high_bits(unsigned v) { return 31u - (unsigned)__clz(v); }   // bits.Len32(v)-1, v > 0

template <typename T, typename Op>
__device__ __forceinline__ T block_reduce(T v, T* s_tmp, Op op) {
  const unsigned lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) v = op(v, __shfl_xor_sync(0xffffffffu, v, d));
  __syncthreads();
  if (lane == 0) s_tmp[warp] = v;
  __syncthreads();
  T r = s_tmp[0];
#pragma unroll
  for (int w = 1; w < T_THREADS / 32; w++) r = op(r, s_tmp[w]);
  return r;
}

__device__ __forceinline__ unsigned block_excl_scan_t(unsigned v, unsigned* s_warp, unsigned* total) {
  const unsigned lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  unsigned inc = v;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    const unsigned t = __shfl_up_sync(0xffffffffu, inc, d);
    if (lane >= (unsigned)d) inc += t;
  }
  __syncthreads();
  if (lane == 31) s_warp[warp] = inc;
  __syncthreads();
  unsigned base = 0, tot = 0;
#pragma unroll
  for (int w = 0; w < T_THREADS / 32; w++) {
    const unsigned s = s_warp[w];
    if ((unsigned)w < warp) base += s;
    tot += s;
  }
  *total = tot;
  return base + inc - v;
}

__constant__ unsigned c_rtb[8] = {0, 473195, 504333, 520860, 550000, 700000, 750000, 830000};

// normalizeCount2 (fsecompressu16.go:582-667), serial
__device__ int normalize_count2(const unsigned* count, int* norm, unsigned symlen, unsigned n, unsigned tl) {
  const int NYA = -2;
  unsigned distributed = 0, total = n;
  const unsigned low_threshold = total >> tl;
  unsigned low_one = (total * 3) >> (tl + 1);
  for (unsigned i = 0; i < symlen; i++) {
    const unsigned c = count[i];
    if (c == 0) { norm[i] = 0; continue; }
    if (c <= low_threshold) { norm[i] = -1; distributed++; total -= c; continue; }
    if (c <= low_one) { norm[i] = 1; distributed++; total -= c; continue; }
    norm[i] = NYA;
  }
  if (distributed >= (1u << tl)) return MIC_ENC_INTERNAL;   // Go underflows toDistribute and spins for ~2^32 iterations
  unsigned to_distribute = (1u << tl) - distributed;
  if ((total / to_distribute) > low_one) {
    low_one = (total * 3) / (to_distribute * 2);
    for (unsigned i = 0; i < symlen; i++)
      if (norm[i] == NYA && count[i] <= low_one) { norm[i] = 1; distributed++; total -= count[i]; }
    if (distributed >= (1u << tl)) return MIC_ENC_INTERNAL;
    to_distribute = (1u << tl) - distributed;
  }
  if (distributed == symlen + 1) {
    unsigned max_v = 0, max_c = 0;
    for (unsigned i = 0; i < symlen; i++)
      if (count[i] > max_c) { max_v = i; max_c = count[i]; }
    norm[max_v] += (int)to_distribute;
    return 0;
  }
  if (total == 0) {
    for (unsigned i = 0; to_distribute > 0; i = (i + 1) % symlen)
      if (norm[i] > 0) { to_distribute--; norm[i]++; }
    return 0;
  }
  const unsigned long long v_step_log = 62 - (unsigned long long)tl;
  const unsigned long long mid = (1ull << (v_step_log - 1)) - 1;
  const unsigned long long r_step = (((1ull << v_step_log) * (unsigned long long)to_distribute) + mid) / (unsigned long long)total;
  unsigned long long tmp_total = mid;
  for (unsigned i = 0; i < symlen; i++) {
    if (norm[i] == NYA) {
      const unsigned long long end = tmp_total + (unsigned long long)count[i] * r_step;
      const unsigned weight = (unsigned)(end >> v_step_log) - (unsigned)(tmp_total >> v_step_log);
      if (weight < 1) return MIC_ENC_INTERNAL;
      norm[i] = (int)weight;
      tmp_total = end;
    }
  }
  return 0;
}

// writeCount (fsecompressu16.go:191-289), serial; returns header length or a negative status
__device__ int write_count(const int* norm, unsigned symlen, unsigned tl, uint8_t* out) {
  const int table_size = 1 << tl;
  bool previous0 = false;
  unsigned charnum = 0;
  unsigned bit_stream = tl - 5, bit_count = 4;
  int remaining = table_size + 1, threshold = table_size;
  unsigned nb_bits = tl + 1;
  unsigned outp = 0;
  while (remaining > 1) {
    if (previous0) {
      unsigned start = charnum;
      while (charnum < symlen && norm[charnum] == 0) charnum++;
      if (charnum >= symlen) return MIC_ENC_INTERNAL;
      while (charnum >= start + 24) {
        start += 24;
        bit_stream += 0xFFFFu << bit_count;
        out[outp] = (uint8_t)bit_stream; out[outp + 1] = (uint8_t)(bit_stream >> 8);
        outp += 2;
        bit_stream >>= 16;
      }
      while (charnum >= start + 3) {
        start += 3;
        bit_stream += 3u << bit_count;
        bit_count += 2;
      }
      bit_stream += (charnum - start) << bit_count;
      bit_count += 2;
      if (bit_count > 16) {
        out[outp] = (uint8_t)bit_stream; out[outp + 1] = (uint8_t)(bit_stream >> 8);
        outp += 2;
        bit_stream >>= 16;
        bit_count -= 16;
      }
    }
    if (charnum >= symlen) return MIC_ENC_INTERNAL;
    int count = norm[charnum];
    charnum++;
    const int max = (2 * threshold - 1) - remaining;
    if (count < 0) remaining += count; else remaining -= count;
    count++;
    if (count >= threshold) count += max;
    bit_stream += (unsigned)count << bit_count;
    bit_count += nb_bits;
    if (count < max) bit_count--;
    previous0 = (count == 1);
    if (remaining < 1) return MIC_ENC_INTERNAL;
    while (remaining < threshold) { nb_bits--; threshold >>= 1; }
    if (bit_count > 16) {
      out[outp] = (uint8_t)bit_stream; out[outp + 1] = (uint8_t)(bit_stream >> 8);
      outp += 2;
      bit_stream >>= 16;
      bit_count -= 16;
    }
  }
  out[outp] = (uint8_t)bit_stream; out[outp + 1] = (uint8_t)(bit_stream >> 8);
  outp += (bit_count + 7) / 8;
  if (charnum > symlen) return MIC_ENC_INTERNAL;
  return (int)outp;
}

// scratch per CTA (global): count u32[65536] | norm i32[65536] | cumul u32[65537] | posc u32[65537] | rank_sym u16[65536] | cell_sym u16[65536]
constexpr unsigned long long K6_SCRATCH = 65536ull * 4 * 2 + 65540ull * 4 * 2 + 65536ull * 2 * 2;

__global__ void __launch_bounds__(T_THREADS)
k_enc_tables(MicEncUnit* __restrict__ units, int nunits, const uint16_t* __restrict__ Sbuf, uint8_t* __restrict__ scratch,
             uint16_t* __restrict__ state_tab, uint2* __restrict__ sym_tt, uint8_t* __restrict__ hdrs) {
  __shared__ unsigned s_u[T_THREADS / 32];
  __shared__ unsigned long long s_ull[T_THREADS / 32];
  __shared__ int s_status, s_hdr;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  uint8_t* my = scratch + (unsigned long long)blockIdx.x * K6_SCRATCH;
  unsigned* count = reinterpret_cast<unsigned*>(my);
  int* norm = reinterpret_cast<int*>(count + 65536);
  unsigned* cumul = reinterpret_cast<unsigned*>(norm + 65536);     // [65537] table slots before symbol s (low-prob counts 1)
  unsigned* posc = cumul + 65540;                                  // [65537] positive-norm slots before symbol s
  uint16_t* rank_sym = reinterpret_cast<uint16_t*>(posc + 65540);  // symbol owning spread rank k
  uint16_t* cell_sym = rank_sym + 65536;                           // tableSymbol

  for (int ui = blockIdx.x; ui < nunits; ui += gridDim.x) {
    MicEncUnit* U = &units[ui];
    __syncthreads();
    if (U->status != MIC_ENC_OK) continue;
    const unsigned n = U->s_len;
    const uint16_t* S = Sbuf + U->s_off;
    // ---- histogram (countSimple) -------------------------------------------------------------
    unsigned mx = 0;
    for (unsigned i = tid; i < n; i += T_THREADS) mx = max(mx, (unsigned)S[i]);
    const unsigned symlen = block_reduce<unsigned>(mx, s_u, [](unsigned a, unsigned b) { return max(a, b); }) + 1;
    for (unsigned i = tid; i < symlen; i += T_THREADS) count[i] = 0;
    __syncthreads();
    for (unsigned i = tid; i < n; i += T_THREADS) atomicAdd(&count[S[i]], 1u);
    __syncthreads();
    unsigned mc = 0;
    for (unsigned i = tid; i < symlen; i += T_THREADS) mc = max(mc, count[i]);
    const unsigned max_count = block_reduce<unsigned>(mc, s_u, [](unsigned a, unsigned b) { return max(a, b); });
    // ---- reject tests shared by every tier (fsecompressu16.go:37-46) -----------------------------
    int status = MIC_ENC_OK;
    if (n <= 1) status = MIC_ENC_INCOMPRESSIBLE;
    else if (max_count == n) status = MIC_ENC_USE_RLE;
    else if (max_count == 1 || max_count < (n >> 15)) status = MIC_ENC_INCOMPRESSIBLE;
    if (status != MIC_ENC_OK) {
      if (tid == 0) { U->status = status; U->symbol_len = symlen; }
      continue;
    }
    // ---- optimalTableLog (uint8 arithmetic as in Go) ----------------------------------------------
    unsigned tl;
    {
      const unsigned min_bits_src = high_bits(n - 1) + 1, min_bits_sym = high_bits(symlen - 1) + 2;   // symlen >= 2 here
      const unsigned min_bits = (min_bits_src < min_bits_sym ? min_bits_src : min_bits_sym) & 0xFF;
      const unsigned max_bits_src = (high_bits(n - 1) - 2) & 0xFF;
      tl = 11;
      if (max_bits_src < tl) tl = max_bits_src;
      if (min_bits > tl) tl = min_bits;
      const unsigned density = n / symlen;
      if (symlen > 512 && density > 16 && tl < 13) tl = 13;
      else if (density > 64 && symlen > 256 && tl < 12) tl = 12;
      else if (density > 32 && symlen > 128 && tl < 12) tl = 12;
      if (max_bits_src < tl) tl = max_bits_src;
      if (tl < 5) tl = 5;
      if (tl > 16) tl = 16;
    }
    const unsigned Sz = 1u << tl;
    // ---- normalizeCount ---------------------------------------------------------------------------
    {
      const unsigned long long scale = 62 - (unsigned long long)tl;
      const unsigned long long step = (1ull << 62) / (unsigned long long)n;
      const unsigned long long v_step = 1ull << (scale - 20);
      const unsigned low_threshold = n >> tl;
      long long sum = 0;               // sum of proba and of the low-prob ones
      unsigned long long best = 0;     // (proba << 32) | (0xFFFFFFFF - index): max = largest proba, lowest index
      for (unsigned i = tid; i < symlen; i += T_THREADS) {
        const unsigned c = count[i];
        int nv = 0;
        if (c != 0) {
          if (c <= low_threshold) { nv = -1; sum += 1; }
          else {
            int proba = (int)(((unsigned long long)c * step) >> scale);
            if (proba < 8) {
              const unsigned long long rest_to_beat = v_step * (unsigned long long)c_rtb[proba];
              const unsigned long long v = (unsigned long long)c * step - ((unsigned long long)proba << scale);
              if (v > rest_to_beat) proba++;
            }
            nv = proba;
            sum += proba;
            const unsigned long long key = ((unsigned long long)(unsigned)proba << 32) | (0xFFFFFFFFu - i);
            if (proba > 0 && key > best) best = key;
          }
        }
        norm[i] = nv;
      }
      const long long tot = block_reduce<unsigned long long>((unsigned long long)sum, s_ull, [](unsigned long long a, unsigned long long b) { return a + b; });
      best = block_reduce<unsigned long long>(best, s_ull, [](unsigned long long a, unsigned long long b) { return a > b ? a : b; });
      const int still = (int)((long long)Sz - tot);
      // largestP starts at 0 and only a proba > 0 replaces it; with no candidate `largest` stays symbol 0
      const unsigned largest = best ? 0xFFFFFFFFu - (unsigned)(best & 0xFFFFFFFFu) : 0u;
      __syncthreads();
      if (tid == 0) {
        int st = MIC_ENC_OK;
        if (-still >= (norm[largest] >> 1)) st = normalize_count2(count, norm, symlen, n, tl);
        else norm[largest] += still;
        s_status = st;
      }
      __syncthreads();
      if (s_status != MIC_ENC_OK) {
        if (tid == 0) { U->status = s_status; U->symbol_len = symlen; U->table_log = 0; }   // 0: no tables -> the host must not retry a lower tier on them
        continue;
      }
    }
    // ---- writeCount (serial) ------------------------------------------------------------------------
    if (tid == 0) {
      const int h = write_count(norm, symlen, tl, hdrs + U->hdr_off);
      s_hdr = h;
    }
    // ---- buildCTable --------------------------------------------------------------------------------
    // cumul (low-prob symbols take one slot) and posc (slots of the spread ranks); thread t owns a contiguous symbol range
    unsigned nlow;
    {
      const unsigned per = (symlen + T_THREADS - 1) / T_THREADS;
      const unsigned a0 = min(symlen, (unsigned)tid * per), a1 = min(symlen, a0 + per);
      unsigned c_all = 0, c_pos = 0, c_low = 0;
      for (unsigned s = a0; s < a1; s++) {
        const int v = norm[s];
        if (v == -1) { c_all += 1; c_low += 1; }
        else if (v > 0) { c_all += (unsigned)v; c_pos += (unsigned)v; }
      }
      unsigned t_all, t_pos, t_low;
      unsigned b_all = block_excl_scan_t(c_all, s_u, &t_all);
      unsigned b_pos = block_excl_scan_t(c_pos, s_u, &t_pos);
      unsigned b_low = block_excl_scan_t(c_low, s_u, &t_low);
      nlow = t_low;
      if (t_all != Sz) {   // "expected cumul[s.symbolLen] == tableSize"
        if (tid == 0) U->status = MIC_ENC_INTERNAL;
        __syncthreads();
        continue;
      }
      for (unsigned s = a0; s < a1; s++) {
        const int v = norm[s];
        cumul[s] = b_all;
        posc[s] = b_pos;
        if (v == -1) {
          cell_sym[Sz - 1 - b_low] = (uint16_t)s;     // tableSymbol[highThreshold--] = u  (fsecompressu16.go:345-349)
          posc[s] = t_pos + b_low;                    // rANS: low-probability symbols sit after all normal ones (ransu16.go:175-188)
          b_all += 1; b_low += 1;
        } else if (v > 0) {
          for (int q = 0; q < v; q++) rank_sym[b_pos + q] = (uint16_t)s;
          b_all += (unsigned)v; b_pos += (unsigned)v;
        }
      }
    }
    __syncthreads();
    if (U->rans) {
      // buildRansEncTable (ransu16.go:139-197): freq, bias = slots before the symbol (normal symbols first, then the
      // low-probability ones), k0 = tableLog - highBits(freq); threshold = freq << k0 is recomputed by the encoder
      uint2* TTr = sym_tt + U->tt_off;
      for (unsigned i = tid; i < symlen; i += T_THREADS) {
        const int v = norm[i];
        uint2 e = make_uint2(0u, 0u);
        if (v == -1) { e.x = 1u | (tl << 20); e.y = posc[i]; }
        else if (v > 0) { e.x = (unsigned)v | ((tl - high_bits((unsigned)v)) << 20); e.y = posc[i]; }
        TTr[i] = e;
      }
      __syncthreads();
      if (tid == 0) {
        const int h = s_hdr;
        if (h < 0) U->status = h;
        U->hdr_len = h < 0 ? 0u : (unsigned)h;
        U->symbol_len = symlen;
        U->table_log = h < 0 ? 0u : tl;
      }
      continue;
    }
    {
      // spread: cell (j*step)&mask receives the symbol of rank #live(j' < j)  (fsecompressu16.go:373-395)
      const unsigned high_threshold = Sz - 1 - nlow;
      const unsigned step = (Sz >> 1) + (Sz >> 3) + 3, mask = Sz - 1;
      const unsigned jper = Sz / T_THREADS ? Sz / T_THREADS : 1;
      const unsigned j0 = min(Sz, (unsigned)tid * jper), j1 = min(Sz, j0 + jper);
      unsigned live = 0;
      if (nlow) { for (unsigned j = j0; j < j1; j++) live += (((j * step) & mask) <= high_threshold); }
      else live = j1 - j0;
      unsigned t_live;
      unsigned rank = block_excl_scan_t(live, s_u, &t_live);
      for (unsigned j = j0; j < j1; j++) {
        const unsigned pos = (j * step) & mask;
        if (pos <= high_threshold) cell_sym[pos] = rank_sym[rank++];
      }
    }
    __syncthreads();
    // stateTable[cumul[sym]++] = tableSize + u for u ascending (fsecompressu16.go:404-412): cells of one symbol are
    // ranked in table order, 32 cells per step with __match_any_sync; posc[] is recycled as the per-symbol counter
    for (unsigned i = tid; i < symlen; i += T_THREADS) posc[i] = 0;
    __syncthreads();
    uint16_t* ST = state_tab + U->tab_off;     // holds u; the encoder state value is tableSize + u
    if (warp == 0) {
      for (unsigned u0 = 0; u0 < Sz; u0 += 32) {
        const unsigned u = u0 + lane;
        const unsigned sym = cell_sym[u];
        const unsigned peers = __match_any_sync(0xffffffffu, sym);
        const unsigned r = __popc(peers & ((1u << lane) - 1u));
        const unsigned base = posc[sym];
        __syncwarp();
        if (r == 0) posc[sym] = base + __popc(peers);
        __syncwarp();
        ST[cumul[sym] + base + r] = (uint16_t)u;
      }
    }
    // symbolTT (fsecompressu16.go:414-431): `total` of the Go loop is cumul[s]
    uint2* TT = sym_tt + U->tt_off;
    for (unsigned i = tid; i < symlen; i += T_THREADS) {
      const int v = norm[i];
      uint2 e = make_uint2(0u, 0u);
      if (v == -1 || v == 1) {
        e.x = (tl << 16) - Sz;
        e.y = (unsigned)((int)cumul[i] - 1);
      } else if (v > 1) {
        const unsigned max_bits_out = tl - high_bits((unsigned)(v - 1));
        const unsigned min_state_plus = (unsigned)v << max_bits_out;
        e.x = (max_bits_out << 16) - min_state_plus;
        e.y = (unsigned)((int)cumul[i] - v);
      }
      TT[i] = e;
    }
    __syncthreads();
    if (tid == 0) {
      const int h = s_hdr;
      if (h < 0) U->status = h;
      U->hdr_len = h < 0 ? 0u : (unsigned)h;
      U->symbol_len = symlen;
      U->table_log = h < 0 ? 0u : tl;   // 0 tells the host that the tier-independent front half failed
    }
  }
}

// ---------------- K7a: the N state chains ------------------------------------------------------------------
// One warp per unit, lane k < N runs state k over symbols i = k (mod N) from the last to the first and stores
// (bits | nbBits<<16) per symbol; the final states go to T[n + k] (low tableLog bits, fsecompressu16.go:103-107).
template <int N>
__global__ void __launch_bounds__(128)
k_enc_ans(MicEncUnit* __restrict__ units, const int* __restrict__ list, int nlist, const uint16_t* __restrict__ Sbuf,
          const uint16_t* __restrict__ state_tab, const uint2* __restrict__ sym_tt, uint32_t* __restrict__ Tbuf) {
  const int k = threadIdx.x & 31;
  const int wglobal = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int wstride = gridDim.x * (blockDim.x >> 5);
  if (k >= N) return;
  for (int li = wglobal; li < nlist; li += wstride) {
    MicEncUnit* U = &units[list[li]];
    if (U->status != MIC_ENC_OK) continue;
    const unsigned n = U->s_len, tl = U->table_log, Sz = 1u << tl;
    const uint16_t* S = Sbuf + U->s_off;
    const uint16_t* ST = state_tab + U->tab_off;
    const uint2* TT = sym_tt + U->tt_off;
    uint32_t* T = Tbuf + U->t_off;
    if (N == 8 && U->rans) {
      // ransCompress8State (rans8state.go:106-220): states start at 0; xL = x + L, k = k0 - [xL < freq << k0] low bits leave,
      // x = bias + (xL >> k) - freq (ransEncodeStep, ransu16.go:203-213).  Same emission order as the tANS tiers.
      unsigned x = 0;
      if ((unsigned)k < n) {
        long long i = (long long)(n - 1) - (long long)((n - 1 - (unsigned)k) % N);
        unsigned sym1 = i >= N ? S[i - N] : 0u;
        uint2 tt = __ldg(TT + S[i]);
        for (; i >= 0; i -= N) {
          const unsigned sym2 = i >= 2 * N ? S[i - 2 * N] : 0u;   // symbols two steps ahead, table entries one step ahead
          const uint2 tt1 = __ldg(TT + sym1);
          const unsigned freq = tt.x & 0xFFFFFu, k0 = tt.x >> 20;
          const unsigned xL = x + Sz;
          const unsigned kk = k0 - (xL < (freq << k0) ? 1u : 0u);
          T[i] = (xL & ((1u << kk) - 1u)) | (kk << 16);
          x = tt.y + (xL >> kk) - freq;
          tt = tt1;
          sym1 = sym2;
        }
      }
      T[n + k] = x & (Sz - 1u);
      continue;
    }
    unsigned state = Sz;                                   // cStateU16.init: 1 << tableLog
    // No bounds checks inside the chain (a compare + branch per symbol cost 10-40 % of the encode): K6 builds the tables
    // from the histogram of this very stream, so every symbol has cells whenever the unit's status is still OK here.
    if ((unsigned)k < n) {
      long long i = (long long)(n - 1) - (long long)((n - 1 - (unsigned)k) % N);   // last index congruent to k mod N
      // The chain is state -> nbBits -> cell -> stateTable[cell] -> state.  The symbolTT entry of a step depends on the symbol
      // only, so it is loaded one step ahead (and the symbol two steps ahead): on a wide alphabet that entry comes from L2,
      // and fetched inside its own step it sat on the chain (WaveletV2 x16: k_enc_ans<4> 471 ms per call with it there).
      unsigned sym1 = i >= N ? S[i - N] : 0u;
      uint2 tt = __ldg(TT + S[i]);
      for (; i >= 0; i -= N) {
        const unsigned sym2 = i >= 2 * N ? S[i - 2 * N] : 0u;
        const uint2 tt1 = __ldg(TT + sym1);                // index 0 past the first symbol: a valid entry, never used
        const unsigned nb = (state + tt.x) >> 16;
        const int cell = (int)(state >> nb) + (int)tt.y;
        T[i] = (state & ((1u << nb) - 1u)) | (nb << 16);
        state = Sz + __ldg(ST + cell);
        tt = tt1;
        sym1 = sym2;
      }
    }
    T[n + k] = state & (Sz - 1u);
  }
}

// The same chains with the unit's stateTable in SHARED memory: one warp per CTA, the 32 lanes copy the table (2 B per cell,
// 16 KB for a 12-bit strip, 128 KB at tableLog 16), then lane k < N runs state k.  stateTable[cell] is the one load on the
// state -> state chain; through L2 it costs several hundred cycles per symbol (WaveletV2 x16: 400 ms per call for sixteen
// 5 M-symbol streams), from shared memory ~30.  Taken when every listed unit is resident at once (the latency-bound
// regime: wavelet streams, MIC2 frames, a PICS batch: at most four waves of resident units); many short units (MIC3 planes)
// keep the kernel above, whose 64 warps per SM hide the latency by numbers.
template <int N, int D>
__global__ void __launch_bounds__(32)
k_enc_ans_smem(MicEncUnit* __restrict__ units, const int* __restrict__ list, int nlist, const uint16_t* __restrict__ Sbuf,
               const uint16_t* __restrict__ state_tab, const uint2* __restrict__ sym_tt, uint32_t* __restrict__ Tbuf) {
  extern __shared__ __align__(16) uint16_t s_st[];
  const int k = threadIdx.x;
  for (int li = blockIdx.x; li < nlist; li += gridDim.x) {
    MicEncUnit* U = &units[list[li]];
    __syncwarp();                                          // the previous unit's chains are done with the table
    if (U->status != MIC_ENC_OK) continue;
    const unsigned n = U->s_len, tl = U->table_log, Sz = 1u << tl;
    const uint16_t* S = Sbuf + U->s_off;
    const uint2* TT = sym_tt + U->tt_off;
    uint32_t* T = Tbuf + U->t_off;
    {
      const uint4* src = reinterpret_cast<const uint4*>(state_tab + U->tab_off);   // tab_off is a multiple of 8192 entries
      uint4* dst = reinterpret_cast<uint4*>(s_st);
      for (unsigned j = (unsigned)k; j < Sz / 8u; j += 32u) dst[j] = __ldg(src + j);
    }
    __syncwarp();
    if (k >= N) continue;
    unsigned state = Sz;
    if ((unsigned)k < n) {
      long long i = (long long)(n - 1) - (long long)((n - 1 - (unsigned)k) % N);
      auto step = [&](long long at, uint2 tt) {
        const unsigned nb = (state + tt.x) >> 16;
        const int cell = (int)(state >> nb) + (int)tt.y;
        T[at] = (state & ((1u << nb) - 1u)) | (nb << 16);
        state = Sz + s_st[cell];
      };
      if (D > 1) {
        // symbolTT entries D steps ahead, symbols 2 D steps ahead, in register queues indexed by step mod D: on a wide
        // alphabet (wavelet coefficients, residuals) an entry comes from L2, several hundred cycles away from a chain
        // whose step is now ~60
        uint2 tq[D];
        unsigned sq[D];
#pragma unroll
        for (int j = 0; j < D; j++) {
          const long long a = i - (long long)j * N, b = i - (long long)(D + j) * N;
          tq[j] = __ldg(TT + (a >= 0 ? S[a] : 0u));
          sq[j] = b >= 0 ? S[b] : 0u;
        }
        while (i - (long long)(D - 1) * N >= 0) {
#pragma unroll
          for (int j = 0; j < D; j++) {
            const uint2 tt = tq[j];
            tq[j] = __ldg(TT + sq[j]);                         // the entry of step i - D N
            const long long c = i - (long long)2 * D * N;
            sq[j] = c >= 0 ? S[c] : 0u;                        // the symbol of the refill after that
            step(i, tt);
            i -= N;
          }
        }
      }
      // D == 1, and the last steps of a queued chain: symbol two steps ahead, entry one step ahead
      if (i >= 0) {
        unsigned sym1 = i >= N ? S[i - N] : 0u;
        uint2 tt = __ldg(TT + S[i]);
        for (; i >= 0; i -= N) {
          const unsigned sym2 = i >= 2 * N ? S[i - 2 * N] : 0u;
          const uint2 tt1 = __ldg(TT + sym1);
          step(i, tt);
          tt = tt1;
          sym1 = sym2;
        }
      }
    }
    T[n + k] = state & (Sz - 1u);
  }
}

// ---------------- K7b: bit positions + packing + frame assembly ------------------------------------------------
// frame = [0xFF, magic, count u32] (N > 1) | ncount header | bitstream.  Fields are appended LSB-first
// (bitwriter.go:50-53) in descending symbol order, then states N-1..0, then the end-mark bit.
__global__ void __launch_bounds__(T_THREADS)
k_enc_pack(MicEncUnit* __restrict__ units, int nunits, const uint32_t* __restrict__ Tbuf, const uint8_t* __restrict__ hdrs,
           uint8_t* __restrict__ frames) {
  constexpr int PER = 8;
  constexpr int CHUNK = PER * T_THREADS;           // symbols per iteration
  constexpr int WORDS = CHUNK * 16 / 32 + 8;       // worst case 16 bits per symbol
  __shared__ unsigned s_words[WORDS];
  __shared__ unsigned s_warp[T_THREADS / 32];
  const int tid = threadIdx.x;
  for (int ui = blockIdx.x; ui < nunits; ui += gridDim.x) {
    MicEncUnit* U = &units[ui];
    __syncthreads();
    if (U->status != MIC_ENC_OK) continue;
    const unsigned n = U->s_len, N = U->nstates, tl = U->table_log;
    const uint32_t* T = Tbuf + U->t_off;
    uint8_t* F = frames + U->out_off;                // 4-byte aligned
    const unsigned prefix = (N > 1 ? 6u : 0u) + U->hdr_len;
    // minimum lengths of the tiers (fsecompressu16.go:20, fse4state.go:26, fse8state.go:33)
    const unsigned reject = N == 8 ? 7u : N == 4 ? 3u : 1u;
    if (n <= reject) {
      if (tid == 0) U->status = MIC_ENC_INCOMPRESSIBLE;
      continue;
    }
    if (prefix + 2ull * n + 64 > U->out_cap) {
      if (tid == 0) U->status = MIC_ENC_CAPACITY;
      continue;
    }
    // prefix bytes
    if (N > 1 && tid < 6) {
      const uint8_t magic = N == 2 ? 0x02 : N == 4 ? 0x04 : (U->rans ? 0x08 : 0x84);
      F[tid] = tid == 0 ? 0xFF : tid == 1 ? magic : (uint8_t)(n >> (8 * (tid - 2)));
    }
    const uint8_t* H = hdrs + U->hdr_off;
    for (unsigned i = tid; i < U->hdr_len; i += T_THREADS) F[(N > 1 ? 6u : 0u) + i] = H[i];
    __syncthreads();
    // the bitstream starts at byte `prefix`; words are assembled 4-byte aligned, so the first word is seeded with
    // the prefix bytes that share it
    unsigned long long pos = 8ull * prefix;          // next free bit, relative to F
    unsigned carry = 0;                              // already-set low bits of word pos>>5
    {
      const unsigned w0 = (unsigned)(pos >> 5) * 4;
      for (unsigned b = w0; b < prefix; b++) carry |= (unsigned)F[b] << (8 * (b - w0));
    }
    const unsigned total_items = n + N;              // symbols (descending) then states N-1..0
    for (unsigned base = 0; base < total_items; base += CHUNK) {
      // item j (emission order): j < n -> symbol n-1-j ; else state N-1-(j-n) with tableLog bits
      unsigned bits[PER], nb[PER], sum = 0;
      const unsigned j0 = base + tid * PER;
#pragma unroll
      for (int q = 0; q < PER; q++) {
        const unsigned j = j0 + q;
        bits[q] = 0; nb[q] = 0;
        if (j < n) { const unsigned t = T[n - 1 - j]; bits[q] = t & 0xFFFFu; nb[q] = t >> 16; }
        else if (j < total_items) { bits[q] = T[n + (N - 1 - (j - n))]; nb[q] = tl; }
        sum += nb[q];
      }
      unsigned chunk_bits;
      unsigned off = block_excl_scan_t(sum, s_warp, &chunk_bits);
      const unsigned wbase = (unsigned)(pos >> 5);              // first word of this chunk
      const unsigned bit0 = (unsigned)(pos & 31);
      const unsigned nwords = (bit0 + chunk_bits + 31) >> 5;
      for (unsigned w = tid; w < nwords + 1; w += T_THREADS) s_words[w] = 0;
      __syncthreads();
      if (tid == 0) s_words[0] = carry;
      __syncthreads();
      unsigned p = bit0 + off;
#pragma unroll
      for (int q = 0; q < PER; q++) {
        if (nb[q]) {
          const unsigned w = p >> 5, sh = p & 31;
          atomicOr(&s_words[w], bits[q] << sh);
          if (sh + nb[q] > 32) atomicOr(&s_words[w + 1], bits[q] >> (32 - sh));
          p += nb[q];
        }
      }
      __syncthreads();
      const unsigned endbit = bit0 + chunk_bits;
      const unsigned full = endbit >> 5;                        // complete words
      unsigned* Fw = reinterpret_cast<unsigned*>(F) + wbase;
      for (unsigned w = tid; w < full; w += T_THREADS) Fw[w] = s_words[w];
      carry = s_words[full];
      pos += chunk_bits;
      __syncthreads();
    }
    // end mark + flushAlign (bitwriter.go:151-168)
    if (tid == 0) {
      const unsigned sh = (unsigned)(pos & 31);
      carry |= 1u << sh;
      const unsigned long long total_bits = pos + 1;
      const unsigned long long nbytes = (total_bits + 7) >> 3;            // frame length
      const unsigned w0 = (unsigned)(pos >> 5) * 4;
      for (unsigned long long b = w0; b < nbytes; b++) F[b] = (uint8_t)(carry >> (8 * (b - w0)));
      const unsigned long long body = nbytes - (N > 1 ? 6u : 0u);          // len(s.Out): header + bitstream
      U->bits_total = (unsigned)(pos - 8ull * prefix);
      U->frame_len = (unsigned)nbytes;
      U->used_states = N;
      if (body >= 2ull * n) U->status = MIC_ENC_INCOMPRESSIBLE;           // "if len(s.Out) >= len(in)*2"
    }
  }
}

void launch_enc_tables(MicEncUnit* d_units, int nunits, const uint16_t* d_S, uint8_t* d_scratch, uint16_t* d_state_tab, uint2* d_sym_tt,
                       uint8_t* d_hdrs, int grid, cudaStream_t st) {
  if (nunits <= 0) return;
  k_enc_tables<<<grid, T_THREADS, 0, st>>>(d_units, nunits, d_S, d_scratch, d_state_tab, d_sym_tt, d_hdrs);
}
unsigned long long enc_tables_scratch_per_cta() { return K6_SCRATCH + 256; }

void launch_enc_ans(MicEncUnit* d_units, const int* d_list, int nlist, int nstates, const uint16_t* d_S, const uint16_t* d_state_tab,
                    const uint2* d_sym_tt, uint32_t* d_T, int sm_count, cudaStream_t st, int max_table_log, bool any_rans) {
  if (nlist <= 0) return;
  // stateTable in shared memory when every listed unit can be resident at once (k_enc_ans_smem); max_table_log bounds the
  // tables of the list (the host's reservation, enc_plan), 0 = unknown.  rANS units use no stateTable.
  static const int smem_cfg = [] { const char* e = getenv("MICGPU_K7_SMEM"); return e ? atoi(e) : 1; }();
  if (smem_cfg && max_table_log >= 5 && max_table_log <= 16 && !any_rans) {
    const size_t bytes = (size_t)2 << max_table_log;
    const int per_sm = (int)std::min<size_t>((227 * 1024) / (bytes + 1024), 16);
    if (per_sm >= 1 && nlist <= 4 * sm_count * per_sm) {   // at most four waves of resident units
      auto go = [&](auto kern) {
        cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
        kern<<<nlist, 32, bytes, st>>>(d_units, d_list, nlist, d_S, d_state_tab, d_sym_tt, d_T);
      };
      // MICGPU_K7_DEPTH: how many steps ahead the symbolTT entries are fetched (8, or 1 for A/B runs).  WaveletV2 x16 encode per
      // call: 724 ms with the tables read through L1/L2, 663 with the entry one step ahead, 505 with the state table in shared
      // memory, 406 with the entries eight steps ahead.
      static const int depth_cfg = [] { const char* e = getenv("MICGPU_K7_DEPTH"); return e ? atoi(e) : 8; }();
      if (depth_cfg > 1) {
        switch (nstates) {
          case 1: go(k_enc_ans_smem<1, 8>); break;
          case 2: go(k_enc_ans_smem<2, 8>); break;
          case 4: go(k_enc_ans_smem<4, 8>); break;
          default: go(k_enc_ans_smem<8, 8>); break;
        }
      } else {
        switch (nstates) {
          case 1: go(k_enc_ans_smem<1, 1>); break;
          case 2: go(k_enc_ans_smem<2, 1>); break;
          case 4: go(k_enc_ans_smem<4, 1>); break;
          default: go(k_enc_ans_smem<8, 1>); break;
        }
      }
      return;
    }
  }
  const int warps_per_cta = 4;
  int grid = (nlist + warps_per_cta - 1) / warps_per_cta;
  if (grid > sm_count * 16) grid = sm_count * 16;
  switch (nstates) {
    case 1: k_enc_ans<1><<<grid, 128, 0, st>>>(d_units, d_list, nlist, d_S, d_state_tab, d_sym_tt, d_T); break;
    case 2: k_enc_ans<2><<<grid, 128, 0, st>>>(d_units, d_list, nlist, d_S, d_state_tab, d_sym_tt, d_T); break;
    case 4: k_enc_ans<4><<<grid, 128, 0, st>>>(d_units, d_list, nlist, d_S, d_state_tab, d_sym_tt, d_T); break;
    default: k_enc_ans<8><<<grid, 128, 0, st>>>(d_units, d_list, nlist, d_S, d_state_tab, d_sym_tt, d_T); break;
  }
}

void launch_enc_pack(MicEncUnit* d_units, int nunits, const uint32_t* d_T, const uint8_t* d_hdrs, uint8_t* d_frames, int grid, cudaStream_t st) {
  if (nunits <= 0) return;
  k_enc_pack<<<grid, T_THREADS, 0, st>>>(d_units, nunits, d_T, d_hdrs, d_frames);
}

}  // namespace micgpu
