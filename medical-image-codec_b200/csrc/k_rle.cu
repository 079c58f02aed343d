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
//
// Instruction economy (the kernel was issue bound at 64 % with 77 thread instructions per symbol, profiles/):
//   * staging translates eight states per thread step (one 16 B load, eight table reads, one 16 B store);
//   * s_e is laid out destination-aligned: element o sits at s_e[ph + o] with ph chosen so that an aligned group of
//     eight elements is an aligned 16 B store in D (D's row phase follows the global pixel index, mic_unit.h);
//   * the common chunk -- no delimiter value in it -- is written optimistically, 16 B per thread step, with the
//     delimiter test fused into the same pass; a chunk that turns out to hold a delimiter is redone by the exact
//     marker/compaction passes (they rewrite the same D range);
//   * units are handed out through an atomic queue: CTAs are resident five per SM, a static stride left the last
//     wave half empty.
#include <algorithm>
#include <cstdlib>

#include "mic_device.cuh"

namespace micgpu {

// Launch shapes.  Every chunk costs ~6 CTA barriers around short phases and the header walk is one warp's serial work,
// so the trade is: bigger chunks = fewer barrier stalls per symbol, smaller CTAs = more walkers and more CTAs to overlap
// on one SM.  <256, 4096> (5 CTAs/SM) serves the strips; the smaller shapes are for run-heavy planes (MIC3 8-bit tiles,
// residual frames), where 7 of 8 warps wait for the walker (profiles/README.md).  MICGPU_K3_SHAPE picks one for A/B runs.
template <int THREADS, int CHUNK>
struct K3Shape {
  static constexpr int K3_THREADS = THREADS;
  static constexpr int K3_WARPS = THREADS / 32;
  static constexpr int IN_N = CHUNK;
  static constexpr int OUT_CH = CHUNK;
  static constexpr int NWIN = CHUNK / 32;
  static constexpr int WPL = (NWIN + 31) / 32;   // 32-element windows per lane in the single-warp scans
  static constexpr int MAXR = CHUNK / 16;        // runs per chunk
  static constexpr int MINB = (THREADS * 48 * 5 <= 65536 && CHUNK <= 4096) ? (65536 / (THREADS * 48) > 16 ? 16 : 65536 / (THREADS * 48)) : 4;
};

struct WalkState {
  int ipos;        // next input symbol to parse
  int wbase, wend; // staged window [wbase, wend)
  unsigned c_rem;  // outputs left in the current run
  int kind;        // 0 same, 1 diff
  unsigned value;  // recurring value of a same-run
  int nout;        // elements produced by the last walk
  int nruns;       // runs listed by the last walk
  int done, err;
  int mdirty;      // PW: the speculative masks are fresh, the true header mask has not been stitched yet
  int hcount, hidx; // PW: headers on the list, index of the next one (the header at ipos)
  int pwmode;      // PW: 1 = this window is walked in parallel (chosen at every predicted restage from the header density)
  int hseen;       // headers consumed since the last restage
};

// PW: parallel header walk.  The run headers of a stream form a chain (the position of a header follows from the one
// before it), which one warp used to follow run by run while the other warps waited: the bound of run-heavy streams
// (temporal residual frames, wavelet coefficient streams: a run every ~5 symbols; 8-bit tile planes).  The chain
// re-synchronises: followed from a wrong position it falls onto the true chain within a few runs (every diff-run of even
// length and every symbol above midCount re-aligns it).  So at every restage each warp follows the chain of its own
// 1/K3_WARPS of the window speculatively from the first symbol of that part (one shuffle per run, 32 symbols per
// shared-memory load) and leaves a bit mask of its headers; warp 0 then stitches the true chain: where the true entry
// of a part lies on the part's speculative chain, the rest of that part is adopted, else the chain is followed by hand
// until it meets the speculative one.  With the header bit mask of the window, the headers of a chunk are taken 32 at a
// time, one per lane: run lengths by a warp scan, short runs expanded by their lane, long ones listed for phase B2.
template <int THREADS, int CHUNK, bool PW>
__global__ void __launch_bounds__(THREADS, K3Shape<THREADS, CHUNK>::MINB)
k_rle_expand(MicUnit* __restrict__ units, int nunits, const uint16_t* __restrict__ states,
             const uint16_t* __restrict__ tabS, uint16_t* __restrict__ D, uint32_t* __restrict__ M,
             uint16_t* __restrict__ out, int tab_smem_log, unsigned int* __restrict__ queue, int ubase, int pw_min_headers) {
  using SH = K3Shape<THREADS, CHUNK>;
  constexpr int K3_THREADS = SH::K3_THREADS, K3_WARPS = SH::K3_WARPS, IN_N = SH::IN_N, OUT_CH = SH::OUT_CH, NWIN = SH::NWIN,
                WPL = SH::WPL, MAXR = SH::MAXR;
  __shared__ __align__(16) uint16_t s_in[IN_N];
  __shared__ __align__(16) uint16_t s_e_raw[OUT_CH + 16];
  // the compaction buffer of the exact (slow) path shares s_in: that path is rare and forces the window to be restaged
  uint16_t* const s_p = s_in;
  __shared__ int s_ui;
  __shared__ uint32_t s_non[NWIN], s_mark[NWIN];
  __shared__ int s_wprev[NWIN], s_pixbase[NWIN + 1];
  __shared__ uint16_t s_run_o[MAXR], s_run_n[MAXR];   // run list of the current chunk: output start, length
  __shared__ int s_run_src[MAXR];                     // >= 0: first literal in s_in (diff-run); < 0: -(value+1) (same-run)
  __shared__ WalkState ws;
  constexpr int SW = IN_N / K3_WARPS;                 // symbols of the window each warp follows speculatively (multiple of 32)
  __shared__ uint32_t s_spec[PW ? IN_N / 32 : 1];     // speculative header bits, one word per 32 window symbols
  __shared__ uint32_t s_hdr[PW ? IN_N / 32 : 1];      // true header bits of the window (valid from the entry of the last restage on)
  __shared__ int s_exit[PW ? K3_WARPS : 1];           // first position of each part's speculative chain beyond the part
  __shared__ uint16_t s_hpos[PW ? IN_N / 2 : 1];      // the true headers of the window in order (window-relative positions)
  extern __shared__ __align__(16) uint16_t s_tab[];   // tabS of the current unit (when it fits: tab_smem_log >= tableLog)

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

  while (true) {
    __syncthreads();
    if (tid == 0) s_ui = (int)atomicAdd(queue, 1u);
    __syncthreads();
    const int ui = s_ui;
    if (ui >= nunits) break;
    MicUnit* U = &units[ubase + ui];
    if (U->status != MIC_OK) continue;
    const int nsym = (int)U->nsym;
    const uint16_t* st = states + U->sym_off;
    const uint16_t* Sy = tabS + U->tab_off;
    const bool spatial = U->kind == MIC_KIND_SPATIAL;
    // strips of 12-bit and deeper pixels have a few runs per chunk (diff-runs of thousands): the one-warp walk is cheaper there
    const bool use_pw = PW && (!spatial || U->table_log <= 12);
    const unsigned W = U->width, H = U->height, wp = U->wp, align0 = U->align0;
    const unsigned long long npx = (unsigned long long)W * H;

    if (U->kind == MIC_KIND_RAW) {
      // no RLE layer: symbol i is output word i (capacity `width`)
      if ((unsigned long long)nsym > npx) {
        if (tid == 0) U->status = MIC_E_SIZE;
        continue;
      }
      uint16_t* dst = out + U->out_off;
      for (int i = tid; i < nsym; i += K3_THREADS) dst[i] = __ldg(Sy + __ldg(st + i));
      if (tid == 0) U->thr = (unsigned)nsym;
      continue;
    }
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
      if (outlen > npx || (U->exact_len && outlen != npx)) {
        if (tid == 0) U->status = MIC_E_SIZE;
        continue;
      }
    }
    if (tid == 0) {
      ws.ipos = spatial ? 1 : 3;
      ws.wbase = 0; ws.wend = 0;
      ws.c_rem = 0; ws.kind = 0; ws.value = 0; ws.nout = 0; ws.done = 0; ws.err = 0; ws.mdirty = 0; ws.hcount = 0; ws.hidx = 0; ws.pwmode = 0; ws.hseen = 0;
    }
    unsigned thr = 0, delim = 0;
    unsigned long long pix = 0;        // pixels (spatial) or elements (RLE) emitted so far
    unsigned carry_m = 0;              // last element of the previous chunk was a marker
    bool first = true;
    bool window_clobbered = false;
    __syncthreads();

    while (true) {
      // ---------------- A: (re)stage the symbol window ----------------------
      const int ipos = ws.ipos, wend0 = ws.wend;
      const bool pwm = PW && use_pw && ws.pwmode;      // walk this window in parallel (uniform: written before the last barrier)
      const bool restage = first || window_clobbered || (wend0 < nsym && wend0 - ipos < IN_N / 2);
      window_clobbered = false;
      __syncthreads();
      if (restage) {
        // window = [a0, a0 + nb): starts on a 16 B boundary of the state stream (sym_off is a multiple of 16 elements)
        const int a0 = ipos & ~7;
        const int nb = min(nsym - a0, IN_N);
        const uint4* st4 = reinterpret_cast<const uint4*>(st + a0);
        const int ng = nb >> 3;
        // eight states per thread step: one 16 B load, eight table reads, one 16 B store
        for (int v0 = tid; v0 < ng; v0 += 2 * K3_THREADS) {
          const int v1 = v0 + K3_THREADS;
          const uint4 sa = __ldg(st4 + v0);
          uint4 sb = make_uint4(0, 0, 0, 0);
          if (v1 < ng) sb = __ldg(st4 + v1);
          auto xl = [&](uint32_t w) -> uint32_t {
            if (tab_in_smem) return (uint32_t)s_tab[w & 0xFFFFu] | ((uint32_t)s_tab[w >> 16] << 16);
            return (uint32_t)__ldg(Sy + (w & 0xFFFFu)) | ((uint32_t)__ldg(Sy + (w >> 16)) << 16);
          };
          *reinterpret_cast<uint4*>(s_in + 8 * v0) = make_uint4(xl(sa.x), xl(sa.y), xl(sa.z), xl(sa.w));
          if (v1 < ng) *reinterpret_cast<uint4*>(s_in + 8 * v1) = make_uint4(xl(sb.x), xl(sb.y), xl(sb.z), xl(sb.w));
        }
        if (tid < (nb & 7)) {
          const int i = 8 * ng + tid;
          const unsigned sv = __ldg(st + a0 + i);
          s_in[i] = tab_in_smem ? s_tab[sv] : __ldg(Sy + sv);
        }
        if (tid == 0) { ws.wbase = a0; ws.wend = a0 + nb; ws.mdirty = 1; }
      }
      __syncthreads();
      if (pwm && restage) {
        // ---- speculative header chains, one part of the window per warp ----
        const int wlen = ws.wend - ws.wbase;
        const int pbase = warp * SW, pend = min(pbase + SW, wlen);
        int p = pbase;
        for (int blk = pbase; blk < pbase + SW; blk += 32) {
          const unsigned cl = blk + lane < wlen ? (unsigned)s_in[blk + lane] : 1u;
          unsigned mbits = 0;
          while (p < blk + 32 && p < pend) {
            const unsigned c = __shfl_sync(0xffffffffu, cl, p - blk);
            mbits |= 1u << (p - blk);
            p += (c == 0u || c <= mid) ? 2 : 1 + (int)(c - mid);
          }
          if (lane == 0) s_spec[blk >> 5] = mbits;
        }
        if (lane == 0) s_exit[warp] = p;
        __syncthreads();
      }
      // Destination-aligned layout: element o of the chunk sits at se[o] = s_e_raw[ph + o]; for spatial units ph makes
      // (index in s_e_raw) == (pixel index + align0) mod 8, so aligned groups of eight are aligned 16 B stores in D.
      const int skip = (spatial && first) ? 1 : 0;   // element 0 of the stream is maxValue, not a pixel
      const int ph = spatial ? (int)((pix + align0 + 8u - (unsigned)skip) & 7u) : 0;
      uint16_t* se = s_e_raw + ph;
      // ---------------- B: header walk + expansion (warp 0) -----------------
      // The walk is the serial part.  Run-heavy streams (temporal residuals: a run every ~5 elements, 1.25 M runs per
      // tomo frame) are bound by it -- 70-86 % of the kernel's samples are the other warps waiting at the barrier below.
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
        bool slow_once = false;
        int hseen = ws.hseen;
        if (pwm) {
          const int wlen = we - wb;
          int hidx = ws.hidx, hcount = ws.hcount;
          // ---- stitch the true header chain of the window (once per restage) ----
          if (ws.mdirty) {
            for (int i = lane; i < IN_N / 32; i += 32) s_hdr[i] = 0;
            __syncwarp();
            int cur = ip + ((c_rem > 0 && kind == 1) ? (int)c_rem : 0) - wb;    // the next header (after a carried diff-run)
            while (cur < wlen) {
              if ((s_spec[cur >> 5] >> (cur & 31)) & 1u) {
                const int k = cur / SW;
                const int w0 = cur >> 5, w1 = (k + 1) * SW / 32;
                for (int wi = w0 + lane; wi < w1; wi += 32) {
                  unsigned wv = s_spec[wi];
                  if (wi == w0) wv &= ~0u << (cur & 31);
                  s_hdr[wi] |= wv;
                }
                __syncwarp();
                cur = s_exit[k];
              } else {
                if (lane == 0) s_hdr[cur >> 5] |= 1u << (cur & 31);
                __syncwarp();
                const unsigned c = s_in[cur];
                cur += (c == 0u || c <= mid) ? 2 : 1 + (int)(c - mid);
              }
            }
            __syncwarp();
            // the headers in order: lane = mask word, positions by a scan over the word populations
            int total = 0;
            for (int wb32 = 0; wb32 < IN_N / 32; wb32 += 32) {
              unsigned mw = s_hdr[wb32 + lane];
              const int cnt = __popc(mw);
              int inc2 = cnt;
#pragma unroll
              for (int d = 1; d < 32; d <<= 1) {
                const int t = __shfl_up_sync(0xffffffffu, inc2, d);
                if (lane >= d) inc2 += t;
              }
              int o2 = total + inc2 - cnt;
              while (mw) {
                s_hpos[o2++] = (uint16_t)((wb32 + lane) * 32 + (__ffs(mw) - 1));
                mw &= mw - 1u;
              }
              total += __shfl_sync(0xffffffffu, inc2, 31);
            }
            __syncwarp();
            hcount = total;
            hidx = 0;
            if (lane == 0) ws.mdirty = 0;
          }
          // ---- the carried run (a chunk ended inside it) ----
          if (c_rem > 0 && o < budget) {
            int take = (int)min(c_rem, (unsigned)(budget - o));
            int src = -(int)(value + 1u);
            bool ok = true;
            if (kind == 1) {
              const int avail = we - ip;
              if (avail <= 0) {
                if (ip >= nsym) { err = 1; done = 1; }
                ok = false;
              } else {
                take = min(take, avail);
                src = ip - wb;
                ip += take;
              }
            }
            if (ok) {
              if (lane == 0) { s_run_o[nr] = (uint16_t)o; s_run_n[nr] = (uint16_t)take; s_run_src[nr] = src; }
              nr++;
              o += take;
              c_rem -= (unsigned)take;
            }
          }
          // ---- headers, 32 at a time ----
          while (!done && c_rem == 0 && o < budget && nr + 32 <= MAXR) {
            if (ip >= nsym) { done = 1; break; }
            const int hp = ip - wb;
            if (hp >= wlen) break;                                        // the next header lies beyond the window: restage
            if (hidx >= hcount || (int)s_hpos[hidx] != hp) { err = 1; done = 1; break; }   // cannot happen: ip is on the listed chain
            const int n = min(32, hcount - hidx);
            const bool valid = lane < n;
            const int pos = valid ? (int)s_hpos[hidx + lane] : 0;
            const unsigned c = valid ? (unsigned)s_in[pos] : 1u;
            const bool same = c <= mid;
            const int len = same ? (int)c : (int)(c - mid);
            const int apos = wb + pos;
            const bool bad = valid && (c == 0u || apos + 1 >= nsym);      // zero count, or a header that ends the stream
            const bool needre = valid && !bad && pos + 1 >= wlen;         // its value / payload starts beyond the window
            const int avail = same ? len : min(len, max(wlen - (pos + 1), 0));
            const unsigned v = (valid && same && pos + 1 < wlen) ? (unsigned)s_in[pos + 1] : 0u;
            int inc = valid ? avail : 0;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
              const int t = __shfl_up_sync(0xffffffffu, inc, d);
              if (lane >= d) inc += t;
            }
            const int excl = inc - (valid ? avail : 0);
            const int room = budget - o;
            const bool cut = valid && !same && len > avail;               // payload cut by the end of the window
            const bool over = valid && inc > room;
            const unsigned stopm = __ballot_sync(0xffffffffu, bad || needre || cut || over);
            const int f = stopm ? __ffs(stopm) - 1 : n;                   // lanes below f are emitted whole
            int take = lane < f ? avail : 0;
            // the stopping header: what the reference decoder would do with it at this point
            const int f_s = f < 32 ? f : 31;
            const int fbad = __shfl_sync(0xffffffffu, (int)bad, f_s), fneed = __shfl_sync(0xffffffffu, (int)needre, f_s);
            const int fexcl = __shfl_sync(0xffffffffu, excl, f_s), favail = __shfl_sync(0xffffffffu, avail, f_s);
            const int flen = __shfl_sync(0xffffffffu, len, f_s), fsame = __shfl_sync(0xffffffffu, (int)same, f_s);
            const int fpos = __shfl_sync(0xffffffffu, pos, f_s);
            const unsigned fv = __shfl_sync(0xffffffffu, v, f_s);
            bool stop = f < n;
            int ftake = 0;
            if (stop && !fbad && !fneed) ftake = min(favail, room - fexcl);      // partial run: the budget or the window cut it
            if (lane == f && ftake > 0) take = ftake;
            // ---- emit: short runs by their lane, long ones on the run list ----
            const int olane = o + excl;
            const bool emit = take > 0;
            const bool longr = emit && take > 16;
            const unsigned longm = __ballot_sync(0xffffffffu, longr);
            if (longr) {
              const int slot = nr + __popc(longm & ((1u << lane) - 1u));
              s_run_o[slot] = (uint16_t)olane; s_run_n[slot] = (uint16_t)take;
              s_run_src[slot] = same ? -(int)(v + 1u) : pos + 1;
            } else if (emit) {
              if (same) {
                for (int j = 0; j < take; j++) se[olane + j] = (uint16_t)v;
              } else {
                for (int j = 0; j < take; j++) se[olane + j] = s_in[pos + 1 + j];
              }
            }
            nr += __popc(longm);
            // totals and the position after the last whole run
            const int nwhole = f;
            int tot = 0, ipn = ip;
            if (nwhole > 0) {
              tot = __shfl_sync(0xffffffffu, inc, nwhole - 1);
              const int lpos = __shfl_sync(0xffffffffu, pos, nwhole - 1), llen = __shfl_sync(0xffffffffu, len, nwhole - 1);
              const int lsame = __shfl_sync(0xffffffffu, (int)same, nwhole - 1);
              ipn = wb + lpos + (lsame ? 2 : 1 + llen);
            }
            o += tot;
            ip = ipn;
            hidx += nwhole;
            if (stop) {
              if (fbad) { err = 1; done = 1; break; }
              if (fneed) break;                                            // ip is that header: the restage brings its payload
              hidx++;                                                      // the partial run's header is consumed
              // partial (possibly empty) piece of run f
              o += ftake;
              c_rem = (unsigned)(flen - ftake);
              kind = fsame ? 0 : 1;
              value = fv;
              ip = wb + fpos + (fsame ? 2 : 1 + ftake);
              break;
            }
          }
          hseen = hidx;
          if (lane == 0) { ws.hidx = hidx; ws.hcount = hcount; }
        } else
        while (o < budget && nr < MAXR) {
          // Long-run path: headers whose run is longer than 32 elements are followed one by one with uniform (broadcast)
          // loads -- one shared-memory load and a dozen instructions per header, ~70 cycles against ~200 for the
          // shuffle block below, which pays off only when a 32-symbol load covers several headers.  8-bit planes
          // (MIC3 tiles: midCount 127, a header every <= 124 symbols) spend most of this kernel's time in here.
          while (c_rem == 0 && o < budget && nr < MAXR && ip + 2 <= we && ip + 2 <= nsym) {
            const unsigned c = s_in[ip - wb];
            const int room = budget - o;
            if (c == 0) break;                                  // malformed: the scalar iteration reports it
            if (c <= mid) {
              if (c <= 32u) break;                              // short runs: the shuffle block
              const unsigned v = s_in[ip + 1 - wb];
              const int take = min((int)c, room);
              if (lane == 0) { s_run_o[nr] = (uint16_t)o; s_run_n[nr] = (uint16_t)take; s_run_src[nr] = -(int)(v + 1u); }
              nr++; o += take; ip += 2; hseen++;
              if (take < (int)c) { c_rem = c - (unsigned)take; kind = 0; value = v; }
            } else {
              const int len = (int)(c - mid);
              if (len <= 32) break;
              const int first = ip + 1;
              const int take = min(min(len, room), we - first);   // we - first >= 1 here
              if (lane == 0) { s_run_o[nr] = (uint16_t)o; s_run_n[nr] = (uint16_t)take; s_run_src[nr] = first - wb; }
              nr++; o += take; hseen++;
              if (take < len) { c_rem = (unsigned)(len - take); kind = 1; ip += 1 + take; }
              else ip += 1 + len;
            }
          }
          if (!(o < budget && nr < MAXR)) break;
          // Fast block: 32 symbols in registers, one per lane; the header chain inside them is followed with one
          // shuffle per run (c is warp-uniform after the shuffle, so every branch below is uniform).  The scalar
          // iteration further down costs ~450 cycles per run (two dependent shared-memory loads, three stores, a dozen
          // conditions); run-heavy streams -- temporal residuals have a run every ~5 symbols -- spent 86 % of this
          // kernel waiting for it (profiles/README.md).  Everything irregular (carry-in, window or stream end, a zero
          // count) is left to the scalar iteration.
          if (c_rem == 0 && !slow_once && ip + 33 <= we && ip + 33 <= nsym && nr + 16 <= MAXR) {
            const unsigned cl = s_in[ip - wb + lane];
            int p = 0;
            bool stop = false;
            while (p < 31 && !stop) {
              const unsigned c = __shfl_sync(0xffffffffu, cl, p);
              const unsigned v = __shfl_sync(0xffffffffu, cl, p + 1);
              if (c == 0) { slow_once = true; break; }          // malformed: the scalar path reports it
              const int room = budget - o;
              if (c <= mid) {
                const int take = min((int)c, room);
                if (take <= 32) {                                // short run: expanded right here, not listed
                  if (lane < take) se[o + lane] = (uint16_t)v;
                } else {
                  if (lane == 0) { s_run_o[nr] = (uint16_t)o; s_run_n[nr] = (uint16_t)take; s_run_src[nr] = -(int)(v + 1u); }
                  nr++;
                }
                o += take; p += 2; hseen++;
                if (take < (int)c) { c_rem = c - (unsigned)take; kind = 0; value = v; stop = true; }
              } else {
                const int len = (int)(c - mid);
                const int first = ip + p + 1;                    // stream index of the payload
                const int take = min(min(len, room), we - first);   // we - first >= 1 inside the block
                if (take <= 32) {
                  if (lane < take) se[o + lane] = s_in[first - wb + lane];
                } else {
                  if (lane == 0) { s_run_o[nr] = (uint16_t)o; s_run_n[nr] = (uint16_t)take; s_run_src[nr] = first - wb; }
                  nr++;
                }
                o += take;
                hseen++;
                if (take < len) { c_rem = (unsigned)(len - take); kind = 1; p += 1 + take; stop = true; }
                else p += 1 + len;
              }
              if (o >= budget) stop = true;
            }
            ip += p;
            if (!slow_once) continue;
          }
          slow_once = false;
          if (c_rem == 0) {
            if (ip >= nsym) { done = 1; break; }
            if (ip >= we) break;
            const unsigned c = s_in[ip - wb];
            if (c <= mid) {
              if (c == 0 || ip + 1 >= nsym) { err = 1; done = 1; break; }
              if (ip + 1 >= we) break;
              value = s_in[ip + 1 - wb];
              kind = 0; c_rem = c; ip += 2; hseen++;
            } else {
              kind = 1; c_rem = c - mid; ip += 1; hseen++;
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
        if (!restage && o == 0 && !done && !(we < nsym && we - ip < IN_N / 2)) { err = 1; done = 1; }  // no progress possible
        if (lane == 0) {
          ws.ipos = ip; ws.c_rem = c_rem; ws.kind = kind; ws.value = value;
          ws.nout = o; ws.nruns = nr; ws.done = done; ws.err = err;
          if (PW && use_pw) {
            // the next iteration restages (same predicate as at its top): choose how the new window is walked from the
            // header density of the stretch just consumed -- a header every <= 16 symbols pays for the parallel walk
            // (residual frames: one every ~5), fewer do not (wavelet streams, tile planes, strips)
            if (we < nsym && we - ip < IN_N / 2) { ws.pwmode = hseen >= pw_min_headers ? 1 : 0; ws.hseen = 0; }
            else ws.hseen = hseen;
          }
        }
      }
      __syncthreads();
      const int nout = ws.nout;
      const int done = ws.done;
      if (ws.err) break;
      // ---------------- B2: expand the listed (long) runs, one warp per run ----
      {
        // one warp per run: the per-run set-up is paid once instead of by all eight warps (the kernel is issue bound;
        // with ~6 runs per chunk the set-up was 30 % of all instructions), the copy itself is the same work either way
        const int nr = ws.nruns;
        uint32_t* dw = reinterpret_cast<uint32_t*>(s_e_raw);
        for (int r = warp; r < nr; r += K3_WARPS) {
          const int o = s_run_o[r], n = s_run_n[r], src = s_run_src[r];
          // 32-bit destination words; the odd head / tail elements go scalar
          const int d0 = ph + o, d1 = d0 + n;            // element range in s_e_raw
          const int w0 = (d0 + 1) >> 1, w1 = d1 >> 1;    // full words [w0, w1)
          if (src >= 0) {
            const int sh = src - d0;                      // source index = destination index + sh
            if (lane == 0 && (d0 & 1) && n > 0) s_e_raw[d0] = s_in[d0 + sh];
            if (lane == 1 && (d1 & 1) && d1 - 1 >= w0 * 2 && n > 0) s_e_raw[d1 - 1] = s_in[d1 - 1 + sh];
            if ((sh & 1) == 0) {
              const uint32_t* sw = reinterpret_cast<const uint32_t*>(s_in) + (sh >> 1);
#pragma unroll 4
              for (int w = w0 + lane; w < w1; w += 32) dw[w] = sw[w];
            } else {
#pragma unroll 4
              for (int w = w0 + lane; w < w1; w += 32)
                dw[w] = (uint32_t)s_in[2 * w + sh] | ((uint32_t)s_in[2 * w + 1 + sh] << 16);
            }
          } else {
            const uint32_t v16 = (uint32_t)(-(src + 1)) & 0xFFFFu;
            if (lane == 0 && (d0 & 1) && n > 0) s_e_raw[d0] = (uint16_t)v16;
            if (lane == 1 && (d1 & 1) && d1 - 1 >= w0 * 2 && n > 0) s_e_raw[d1 - 1] = (uint16_t)v16;
            const uint32_t v32 = v16 | (v16 << 16);
#pragma unroll 4
            for (int w = w0 + lane; w < w1; w += 32) dw[w] = v32;
          }
        }
      }
      __syncthreads();

      if (!spatial) {
        // ---------------- C (RLE kind): straight copy ------------------------
        uint16_t* dst = out + U->out_off + pix;
        for (int i = tid; i < nout; i += K3_THREADS) dst[i] = se[i];
        pix += (unsigned)nout;
        first = false;
        if (done || pix >= outlen) break;
        continue;
      }

      // ---------------- C (spatial): escape split ---------------------------
      const bool skip0 = first;  // element 0 of the stream is maxValue, not a pixel
      if (first) {
        if (nout < 1) break;
        const unsigned maxv = se[0];
        const int depth = bit_len16(maxv);
        if (depth == 0) break;   // reported below as a short stream
        thr = (1u << (depth - 1)) - 1u;
        delim = (1u << depth) - 1u;
        if (tid == 0) { U->thr = thr; U->delim = delim; }
      }
      first = false;
      // Optimistic path: if no element of the chunk equals the delimiter (and no literal is pending), every element
      // is a plain pixel.  The chunk is written to D in aligned groups of eight while the same pass looks for the
      // delimiter; a chunk that has one is redone below by the exact passes, which rewrite this D range.
      if (!carry_m) {
        int n = nout - skip;
        if (n < 0) n = 0;
        const int npix_chunk = n;
        if (pix + (unsigned)n > npx) n = (int)(npx - pix);
        const int e_lo = ph + skip, e_hi = e_lo + n;          // pixel elements in s_e_raw coordinates
        const int e_chk = ph + nout;                           // delimiter test covers every element of the chunk
        const int g0 = e_lo >> 3, g1 = (max(e_hi, e_chk) + 7) >> 3;
        uint16_t* Du = D + U->d_off;
        const uint32_t d2 = delim | (delim << 16);
        bool dirty = false;
        // pixel index of s_e_raw[0] is pbase (may be "negative" by less than 8 on the first chunk: use signed)
        const long long pbase = (long long)pix - (long long)e_lo;
        // (y, x) of the first pixel this thread's group can hold; npx < 2^32 (65535 x 65535), so 32-bit division
        int g = g0 + tid;
        unsigned y = 0, x = 0;
        if (g < g1) {
          const long long p0 = pbase + 8ll * g;
          const unsigned pb = p0 > 0 ? (unsigned)p0 : 0u;
          y = pb / W; x = pb - y * W;
        }
        for (; g < g1; g += K3_THREADS) {
          const int eb = 8 * g;
          const uint4 v = *reinterpret_cast<const uint4*>(s_e_raw + eb);
          const long long p0 = pbase + eb;
          const bool interior = eb >= e_lo && eb + 8 <= e_hi;
          if (interior && x + 8u <= W) {
            // zero-halfword test on v ^ delim (exact for 16-bit lanes)
            const uint32_t a = v.x ^ d2, b = v.y ^ d2, c = v.z ^ d2, d = v.w ^ d2;
            const uint32_t z = ((a - 0x00010001u) & ~a) | ((b - 0x00010001u) & ~b) | ((c - 0x00010001u) & ~c) | ((d - 0x00010001u) & ~d);
            dirty |= (z & 0x80008000u) != 0;
            const unsigned ay = (align0 + (y & 7u) * (W & 7u)) & 7u;
            *reinterpret_cast<uint4*>(Du + (unsigned long long)y * wp + ay + x) = v;
          } else {
            // partial group (chunk edge) or a group that straddles a row end: element by element, (yy, xx) walked
            const uint16_t* ve = s_e_raw + eb;
            const int i0 = p0 < 0 ? (int)(-p0) : 0;   // elements before pixel 0 (first chunk only)
            unsigned yy = y, xx = x;
#pragma unroll
            for (int i = 0; i < 8; i++) {
              const int e = eb + i;
              if (e >= e_lo && e < e_chk) dirty |= ve[i] == delim;
              if (i >= i0) {
                if (e >= e_lo && e < e_hi) {
                  const unsigned ay = (align0 + (yy & 7u) * (W & 7u)) & 7u;
                  Du[(unsigned long long)yy * wp + ay + xx] = ve[i];
                }
                if (++xx >= W) { xx = 0; yy++; }
              }
            }
          }
          // advance this thread's group by K3_THREADS groups
          if (W >= 8u * K3_THREADS && p0 >= 0) {
            x += 8u * K3_THREADS;
            if (x >= W) { x -= W; y++; }
          } else {
            const unsigned pn = (unsigned)(p0 + 8ll * K3_THREADS);
            y = pn / W; x = pn - y * W;
          }
        }
        if (!__syncthreads_or(dirty ? 1 : 0)) {
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
        bool isd = valid && se[e] == delim;
        if (e == 0 && (skip0 || carry_m)) isd = false;   // maxValue / forced literal
        const unsigned non = __ballot_sync(0xffffffffu, !isd);
        if (lane == 0) s_non[w] = non;
      }
      __syncthreads();
      // C2: exclusive prefix-max of "last non-delimiter position" over windows
      if (warp == 0) {
        int loc[WPL], run = -1;
#pragma unroll
        for (int q = 0; q < WPL; q++) {
          const int w = lane * WPL + q;
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
        for (int q = 0; q < WPL; q++) {
          const int w = lane * WPL + q;
          if (w < nwin) s_wprev[w] = excl;
          excl = max(excl, loc[q]);
        }
      }
      __syncthreads();
      // C3: marker masks
      for (int w = warp; w < nwin; w += K3_WARPS) {
        const int e = w * 32 + lane;
        const bool valid = e < nout;
        bool isd = valid && se[e] == delim;
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
        int cnt[WPL], sum = 0;
#pragma unroll
        for (int q = 0; q < WPL; q++) {
          const int w = lane * WPL + q;
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
        for (int q = 0; q < WPL; q++) {
          const int w = lane * WPL + q;
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
            s_p[pl] = se[e];
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
      window_clobbered = true;   // s_p == s_in
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

template <int THREADS, int CHUNK, bool PW = false>
static void launch_rle_expand_t(MicUnit* d_units, int ubase, int nunits, const uint16_t* d_states, const uint16_t* d_tabS, uint16_t* d_D,
                                uint32_t* d_M, uint16_t* d_out, int tab_log, int grid, unsigned int* d_queue, cudaStream_t st) {
  const size_t smem = (size_t)2 << tab_log;
  cudaFuncSetAttribute(k_rle_expand<THREADS, CHUNK, PW>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  // headers per consumed stretch (about half a window) from which a window is walked in parallel: one per 16 symbols
  static const int pw_div = [] { const char* e = getenv("MICGPU_PW_DIV"); int v = e ? atoi(e) : 32; return v < 1 ? 1 : v; }();
  k_rle_expand<THREADS, CHUNK, PW><<<grid, THREADS, smem, st>>>(d_units, nunits, d_states, d_tabS, d_D, d_M, d_out, tab_log, d_queue, ubase,
                                                                  CHUNK / pw_div);
}

void launch_rle_expand(MicUnit* d_units, int ubase, int nunits, const uint16_t* d_states, const uint16_t* d_tabS,
                       uint16_t* d_D, uint32_t* d_M, uint16_t* d_out, int max_log, int grid, unsigned int* d_queue,
                       cudaStream_t st, int shape) {
  if (nunits <= 0) return;
  cudaMemsetAsync(d_queue, 0, sizeof(unsigned int), st);
  // stage tabS in shared memory up to tableLog 14 (32 KB); larger tables are gathered through L1/L2
  const int tab_log = max_log <= 14 ? max_log : 14;
  static const int forced = [] { const char* e = getenv("MICGPU_K3_SHAPE"); return e ? atoi(e) : -1; }();
  if (forced >= 0) shape = forced;
  // the grid the caller sized is per 256-thread CTA (8 per SM at most); smaller CTAs come in proportionally larger numbers
  switch (shape) {
    case 1: launch_rle_expand_t<128, 4096>(d_units, ubase, nunits, d_states, d_tabS, d_D, d_M, d_out, tab_log, std::min(nunits, grid * 2), d_queue, st); break;
    case 2: launch_rle_expand_t<128, 2048>(d_units, ubase, nunits, d_states, d_tabS, d_D, d_M, d_out, tab_log, std::min(nunits, grid * 2), d_queue, st); break;
    case 3: launch_rle_expand_t<64, 2048>(d_units, ubase, nunits, d_states, d_tabS, d_D, d_M, d_out, tab_log, std::min(nunits, grid * 4), d_queue, st); break;
    case 4: launch_rle_expand_t<64, 1024>(d_units, ubase, nunits, d_states, d_tabS, d_D, d_M, d_out, tab_log, std::min(nunits, grid * 4), d_queue, st); break;
    // parallel header walk (run-heavy streams)
    case 5: launch_rle_expand_t<256, 4096, true>(d_units, ubase, nunits, d_states, d_tabS, d_D, d_M, d_out, tab_log, grid, d_queue, st); break;
    case 6: launch_rle_expand_t<128, 2048, true>(d_units, ubase, nunits, d_states, d_tabS, d_D, d_M, d_out, tab_log, std::min(nunits, grid * 2), d_queue, st); break;
    default: launch_rle_expand_t<256, 4096>(d_units, ubase, nunits, d_states, d_tabS, d_D, d_M, d_out, tab_log, grid, d_queue, st); break;
  }
}

}  // namespace micgpu
