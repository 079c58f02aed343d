// k_rle.cu -- K3: state->symbol translation, RLE expansion and escape split.
//
// Replaces RleDecompressU16.DecodeNext2/Decompress (rledecompressu16.go:59-97)
// and the symbol-stream half of DeltaRleDecompressU16.Decompress
// (deltarlecompressu16.go:69-128).  One CTA per unit, streaming the unit in
// windows of 4096 symbols:
//   A. all threads stage a window, translating ANS states to symbols via tabS;
//   B. warp 0 walks the run headers (the only serial step: a diff-run's payload
//      can look like a header) and expands each run 32 lanes wide into s_e;
//   C. spatial units only -- escape split, fully parallel: an element equal to
//      the delimiter is an escape *marker* iff its distance to the previous
//      non-delimiter element is odd (a literal may itself equal the delimiter,
//      so runs of delimiter values alternate marker/literal).  Ballots give the
//      per-window masks, a 128-entry prefix-max links windows.  Markers are
//      dropped; every other element is one pixel.  Pixels are compacted into
//      the residual plane D (row pitch wp) and literal pixels set a bit in M.
// RLE-kind units (temporal residuals, wavelet coefficients) copy s_e to d_out.
#include "mic_device.cuh"

namespace micgpu {

constexpr int K3_THREADS = 256;
constexpr int K3_WARPS = K3_THREADS / 32;
constexpr int IN_N = 4096;
constexpr int OUT_CH = 4096;
constexpr int NWIN = OUT_CH / 32;
constexpr int MAXR = 256;   // runs per chunk

struct WalkState {
  int ipos;        // next input symbol to parse
  int wbase, wend; // staged window [wbase, wend)
  unsigned c_rem;  // outputs left in the current run
  int kind;        // 0 same, 1 diff
  unsigned value;  // recurring value of a same-run
  int nout;        // elements produced by the last walk
  int nruns;       // runs listed by the last walk
  int done, err;
};

__global__ void __launch_bounds__(K3_THREADS)
k_rle_expand(MicUnit* __restrict__ units, int nunits, const uint16_t* __restrict__ states,
             const uint16_t* __restrict__ tabS, uint16_t* __restrict__ D, uint32_t* __restrict__ M,
             uint16_t* __restrict__ out, int tab_smem_log) {
  __shared__ uint16_t s_in[IN_N];
  __shared__ uint16_t s_e[OUT_CH];
  __shared__ uint16_t s_p[OUT_CH];
  __shared__ uint32_t s_non[NWIN], s_mark[NWIN];
  __shared__ int s_wprev[NWIN], s_pixbase[NWIN + 1];
  __shared__ uint16_t s_run_o[MAXR], s_run_n[MAXR];   // run list of the current chunk: output start, length
  __shared__ int s_run_src[MAXR];                     // >= 0: first literal in s_in (diff-run); < 0: -(value+1) (same-run)
  __shared__ WalkState ws;
  extern __shared__ __align__(16) uint16_t s_tab[];   // tabS of the current unit (when it fits: tab_smem_log >= tableLog)

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

  for (int ui = blockIdx.x; ui < nunits; ui += gridDim.x) {
    MicUnit* U = &units[ui];
    __syncthreads();
    if (U->status != MIC_OK) continue;
    const int nsym = (int)U->nsym;
    const uint16_t* st = states + U->sym_off;
    const uint16_t* Sy = tabS + U->tab_off;
    const bool spatial = U->kind == MIC_KIND_SPATIAL;
    const unsigned W = U->width, H = U->height, wp = U->wp, align0 = U->align0;
    const unsigned long long npx = (unsigned long long)W * H;

    if (nsym < (spatial ? 2 : 3)) {
      if (tid == 0) U->status = MIC_E_RLE;
      continue;
    }
    // Random 2-byte gathers from tabS through L1/L2 move a 32-byte sector per symbol (measured: K3 was
    // bound by exactly that), so the symbol table is staged in shared memory whenever it fits.
    const bool tab_in_smem = (int)U->table_log <= tab_smem_log;
    if (tab_in_smem) {
      const uint4* src = reinterpret_cast<const uint4*>(Sy);
      uint4* dst = reinterpret_cast<uint4*>(s_tab);
      const int n16 = (1 << U->table_log) / 8;
      for (int i = tid; i < n16; i += K3_THREADS) dst[i] = __ldg(src + i);
    }
    __syncthreads();
    // word 0 fixes midCount (rledecompressu16.go:21-30)
    const unsigned sym0 = Sy[st[0]];
    const int depth0 = bit_len16(sym0);
    if (depth0 == 0) {
      if (tid == 0) U->status = MIC_E_RLE;
      continue;
    }
    const unsigned mid = (1u << (depth0 - 1)) - 1u;
    unsigned long long outlen = 0;  // RLE kind: expected expanded length
    if (!spatial) {
      outlen = ((unsigned long long)Sy[st[1]] << 16) + Sy[st[2]];
      if (outlen > npx) {
        if (tid == 0) U->status = MIC_E_SIZE;
        continue;
      }
    }
    if (tid == 0) {
      ws.ipos = spatial ? 1 : 3;
      ws.wbase = 0; ws.wend = 0;
      ws.c_rem = 0; ws.kind = 0; ws.value = 0; ws.nout = 0; ws.done = 0; ws.err = 0;
    }
    unsigned thr = 0, delim = 0;
    unsigned long long pix = 0;        // pixels (spatial) or elements (RLE) emitted so far
    unsigned carry_m = 0;              // last element of the previous chunk was a marker
    bool first = true;
    __syncthreads();

    while (true) {
      // ---------------- A: (re)stage the symbol window ----------------------
      const int ipos = ws.ipos, wend0 = ws.wend;
      const bool restage = first || (wend0 < nsym && wend0 - ipos < IN_N / 2);
      __syncthreads();
      if (restage) {
        const int nb = min(nsym - ipos, IN_N);
        // two dependent global loads per symbol (state, then tabS[state], L1-resident): batch 8 of each
        for (int i0 = tid; i0 < nb; i0 += 8 * K3_THREADS) {
          unsigned stv[8];
#pragma unroll
          for (int q = 0; q < 8; q++) {
            const int i = i0 + q * K3_THREADS;
            stv[q] = i < nb ? __ldg(st + ipos + i) : 0u;
          }
          if (tab_in_smem) {
#pragma unroll
            for (int q = 0; q < 8; q++) {
              const int i = i0 + q * K3_THREADS;
              if (i < nb) s_in[i] = s_tab[stv[q]];
            }
          } else {
#pragma unroll
            for (int q = 0; q < 8; q++) {
              const int i = i0 + q * K3_THREADS;
              if (i < nb) s_in[i] = __ldg(Sy + stv[q]);
            }
          }
        }
        if (tid == 0) { ws.wbase = ipos; ws.wend = ipos + nb; }
      }
      __syncthreads();
      // ---------------- B: header walk + expansion (warp 0) -----------------
      if (warp == 0) {
        int ip = ws.ipos;
        const int wb = ws.wbase, we = ws.wend;
        unsigned c_rem = ws.c_rem, value = ws.value;
        int kind = ws.kind;
        int o = 0, done = 0, err = 0;
        int budget = OUT_CH;
        if (!spatial) {
          unsigned long long left = outlen - pix;
          if (left < (unsigned long long)budget) budget = (int)left;
          if (budget == 0) done = 1;
        }
        int nr = 0;
        while (o < budget && nr < MAXR) {
          if (c_rem == 0) {
            if (ip >= nsym) { done = 1; break; }
            if (ip >= we) break;
            const unsigned c = s_in[ip - wb];
            if (c <= mid) {
              if (c == 0 || ip + 1 >= nsym) { err = 1; done = 1; break; }
              if (ip + 1 >= we) break;
              value = s_in[ip + 1 - wb];
              kind = 0; c_rem = c; ip += 2;
            } else {
              kind = 1; c_rem = c - mid; ip += 1;
            }
          }
          int take = (int)min(c_rem, (unsigned)(budget - o));
          int src;
          if (kind == 1) {
            const int avail = we - ip;
            if (avail <= 0) {
              if (ip >= nsym) { err = 1; done = 1; }
              break;
            }
            take = min(take, avail);
            src = ip - wb;
            ip += take;
          } else {
            src = -(int)(value + 1u);
          }
          if (lane == 0) { s_run_o[nr] = (uint16_t)o; s_run_n[nr] = (uint16_t)take; s_run_src[nr] = src; }
          nr++;
          o += take;
          c_rem -= (unsigned)take;
        }
        if (!restage && o == 0 && nr == 0 && !done && !(we < nsym && we - ip < IN_N / 2)) { err = 1; done = 1; }  // no progress possible
        if (lane == 0) {
          ws.ipos = ip; ws.c_rem = c_rem; ws.kind = kind; ws.value = value;
          ws.nout = o; ws.nruns = nr; ws.done = done; ws.err = err;
        }
      }
      __syncthreads();
      const int nout = ws.nout;
      const int done = ws.done;
      if (ws.err) break;
      // ---------------- B2: expand the listed runs, all threads per run -------
      {
        const int nr = ws.nruns;
        for (int r = 0; r < nr; r++) {
          const int o = s_run_o[r], n = s_run_n[r], src = s_run_src[r];
          if (src >= 0) {
            for (int j = tid; j < n; j += K3_THREADS) s_e[o + j] = s_in[src + j];
          } else {
            const uint16_t v16 = (uint16_t)(-(src + 1));
            for (int j = tid; j < n; j += K3_THREADS) s_e[o + j] = v16;
          }
        }
      }
      __syncthreads();

      if (!spatial) {
        // ---------------- C (RLE kind): straight copy ------------------------
        uint16_t* dst = out + U->out_off + pix;
        for (int i = tid; i < nout; i += K3_THREADS) dst[i] = s_e[i];
        pix += (unsigned)nout;
        first = false;
        if (done || pix >= outlen) break;
        continue;
      }

      // ---------------- C (spatial): escape split ---------------------------
      const bool skip0 = first;  // element 0 of the stream is maxValue, not a pixel
      if (first) {
        if (nout < 1) break;
        const unsigned maxv = s_e[0];
        const int depth = bit_len16(maxv);
        if (depth == 0) break;   // reported below as a short stream
        thr = (1u << (depth - 1)) - 1u;
        delim = (1u << depth) - 1u;
        if (tid == 0) { U->thr = thr; U->delim = delim; }
      }
      first = false;
      // Fast path: no element of the chunk equals the delimiter (and no pending literal), so every
      // element is a plain pixel and the chunk goes to D without the marker/compaction passes.
      {
        bool has = false;
        for (int i = tid; i < nout; i += K3_THREADS) has |= (s_e[i] == delim) && !(i == 0 && skip0);
        const int anyd = __syncthreads_or(has ? 1 : 0) | (int)carry_m;
        if (!anyd) {
          const int skip = skip0 ? 1 : 0;
          int n = nout - skip;
          if (n < 0) n = 0;
          const int npix_chunk = n;
          if (pix + (unsigned)n > npx) n = (int)(npx - pix);
          if (tid < n) {
            unsigned long long gp = pix + (unsigned)tid;
            unsigned y = (unsigned)(gp / W), x = (unsigned)(gp - (unsigned long long)y * W);
            uint16_t* Du = D + U->d_off;
            for (int i = tid; i < n; i += K3_THREADS) {
              const unsigned ay = (align0 + (y & 7u) * (W & 7u)) & 7u;
              Du[(unsigned long long)y * wp + ay + x] = s_e[i + skip];
              x += K3_THREADS;
              while (x >= W) { x -= W; y++; }
            }
          }
          pix += (unsigned)npix_chunk;
          if (done || pix >= npx) break;
          continue;
        }
      }
      const int nwin = (nout + 31) >> 5;
      // C1: per-window masks of non-delimiter elements
      for (int w = warp; w < nwin; w += K3_WARPS) {
        const int e = w * 32 + lane;
        const bool valid = e < nout;
        bool isd = valid && s_e[e] == delim;
        if (e == 0 && (skip0 || carry_m)) isd = false;   // maxValue / forced literal
        const unsigned non = __ballot_sync(0xffffffffu, !isd);
        if (lane == 0) s_non[w] = non;
      }
      __syncthreads();
      // C2: exclusive prefix-max of "last non-delimiter position" over windows
      if (warp == 0) {
        int loc[4], run = -1;
#pragma unroll
        for (int q = 0; q < 4; q++) {
          const int w = lane * 4 + q;
          const unsigned non = w < nwin ? s_non[w] : 0u;
          loc[q] = non ? (w * 32 + 31 - __clz(non)) : -1;
          run = max(run, loc[q]);
        }
        int inc = run;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
          int t = __shfl_up_sync(0xffffffffu, inc, d);
          if (lane >= d) inc = max(inc, t);
        }
        int excl = __shfl_up_sync(0xffffffffu, inc, 1);
        if (lane == 0) excl = -1;
#pragma unroll
        for (int q = 0; q < 4; q++) {
          const int w = lane * 4 + q;
          if (w < nwin) s_wprev[w] = excl;
          excl = max(excl, loc[q]);
        }
      }
      __syncthreads();
      // C3: marker masks
      for (int w = warp; w < nwin; w += K3_WARPS) {
        const int e = w * 32 + lane;
        const bool valid = e < nout;
        bool isd = valid && s_e[e] == delim;
        if (e == 0 && (skip0 || carry_m)) isd = false;
        const unsigned non = s_non[w];
        const unsigned m = non & (0xffffffffu >> (31 - lane));
        const int lastnon = m ? (w * 32 + 31 - __clz(m)) : s_wprev[w];
        const bool marker = isd && (((e - lastnon) & 1) == 1);
        const unsigned mm = __ballot_sync(0xffffffffu, marker);
        if (lane == 0) s_mark[w] = mm;
      }
      __syncthreads();
      // C4: pixels per window -> exclusive scan
      if (warp == 0) {
        int cnt[4], sum = 0;
#pragma unroll
        for (int q = 0; q < 4; q++) {
          const int w = lane * 4 + q;
          int c = 0;
          if (w < nwin) {
            const int nvalid = min(32, nout - w * 32);
            unsigned vm = nvalid >= 32 ? 0xffffffffu : ((1u << nvalid) - 1u);
            if (w == 0 && skip0) vm &= ~1u;
            c = __popc(~s_mark[w] & vm);
          }
          cnt[q] = c;
          sum += c;
        }
        int inc = sum;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
          int t = __shfl_up_sync(0xffffffffu, inc, d);
          if (lane >= d) inc += t;
        }
        int excl = inc - sum;
#pragma unroll
        for (int q = 0; q < 4; q++) {
          const int w = lane * 4 + q;
          if (w < nwin) s_pixbase[w] = excl;
          excl += cnt[q];
        }
        if (lane == 31) s_pixbase[NWIN] = inc;   // total pixels of this chunk
      }
      __syncthreads();
      const int npix_chunk = s_pixbase[NWIN];
      // C5: compact pixels, flag literals
      for (int w = warp; w < nwin; w += K3_WARPS) {
        const int e = w * 32 + lane;
        const bool valid = e < nout && !(e == 0 && skip0);
        const unsigned mm = s_mark[w];
        const bool marker = (mm >> lane) & 1u;
        unsigned prevm;
        if (lane > 0) prevm = (mm >> (lane - 1)) & 1u;
        else prevm = w > 0 ? (s_mark[w - 1] >> 31) : (skip0 ? 0u : carry_m);
        if (valid && !marker) {
          unsigned vm = 0xffffffffu;
          if (w == 0 && skip0) vm &= ~1u;
          const int pl = s_pixbase[w] + __popc(~mm & vm & ((1u << lane) - 1u));
          const unsigned long long gp = pix + (unsigned)pl;
          if (gp < npx) {
            s_p[pl] = s_e[e];
            if (prevm) {   // literal pixel (deltarlecompressu16.go:105-106)
              const unsigned y = (unsigned)(gp / W), x = (unsigned)(gp - (unsigned long long)y * W);
              const unsigned pc = ((align0 + (y & 7u) * (W & 7u)) & 7u) + x;   // padded column
              atomicOr(&M[U->m_off + (unsigned long long)y * (wp >> 5) + (pc >> 5)], 1u << (pc & 31));
            }
          }
        }
      }
      __syncthreads();
      // C6: coalesced write of the compacted pixels into D (pitch wp)
      {
        int n = npix_chunk;
        if (pix + (unsigned)n > npx) n = (int)(npx - pix);
        if (tid < n) {
          unsigned long long gp = pix + (unsigned)tid;
          unsigned y = (unsigned)(gp / W), x = (unsigned)(gp - (unsigned long long)y * W);
          uint16_t* Du = D + U->d_off;
          for (int i = tid; i < n; i += K3_THREADS) {
            const unsigned ay = (align0 + (y & 7u) * (W & 7u)) & 7u;
            Du[(unsigned long long)y * wp + ay + x] = s_p[i];
            x += K3_THREADS;
            while (x >= W) { x -= W; y++; }
          }
        }
      }
      if (nout > 0) carry_m = (s_mark[(nout - 1) >> 5] >> ((nout - 1) & 31)) & 1u;
      pix += (unsigned)npix_chunk;
      if (done || pix >= npx) break;
    }
    __syncthreads();
    if (tid == 0) {
      if (ws.err) U->status = MIC_E_RLE;
      else if (spatial && pix < npx) U->status = MIC_E_RLE;        // stream ended before the last pixel
      else if (!spatial) {
        U->thr = (unsigned)pix;                                     // expanded length actually produced
        if (pix < outlen) U->status = MIC_E_RLE;
      }
    }
  }
}

void launch_rle_expand(MicUnit* d_units, int nunits, const uint16_t* d_states, const uint16_t* d_tabS,
                       uint16_t* d_D, uint32_t* d_M, uint16_t* d_out, int max_log, int grid, cudaStream_t st) {
  if (nunits <= 0) return;
  // stage tabS in shared memory up to tableLog 14 (32 KB); larger tables are gathered through L1/L2
  const int tab_log = max_log <= 14 ? max_log : 14;
  const size_t smem = (size_t)2 << tab_log;
  cudaFuncSetAttribute(k_rle_expand, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  k_rle_expand<<<grid, K3_THREADS, smem, st>>>(d_units, nunits, d_states, d_tabS, d_D, d_M, d_out, tab_log);
}

}  // namespace micgpu
