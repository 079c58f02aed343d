// k_delta_scan.cu -- K4: inverse avg(top,left) predictor, one ROW per step, lanes = 8-pixel column blocks.
//
// Replaces DeltaRleDecompressU16.DecodeNextSymbol[NC] (deltarlecompressu16.go:102-128) and the C twin's
// delta_decode_simd.  out(x,y) = ((out(x-1,y) + out(x,y-1)) >> 1) + diff is serial along a row, which is why
// k_delta.cu runs an anti-diagonal wavefront (thread = row).  That mapping pays a fill/drain of one 8-pixel block per
// row -- 44 % of all lane-steps on a 2577x256 strip, 89 % on a 256x256 MIC3 tile -- and moves a 16 B block from lane
// to lane through shared memory every step.
//
// Here lane j owns pixel columns [8j, 8j+8) for the whole unit, a warp owns 256 columns, and every lane of the CTA
// works on the SAME row (warps chained left to right, one row apart).  `top` is then the lane's own previous result
// (registers, no exchange at all), and the only thing that crosses lanes is `left`: ONE pixel per block per row.  It is
// obtained without serialising the row by composing the blocks as functions of their incoming left pixel:
//     x -> ((x + t) >> 1) + e   composed over k pixels is   x -> ((x + a) >> k) + b     (nested floors compose),
// a literal pixel (escape) or column 0 makes the block constant, row 0 is x -> x + b.  Functions of the form
// ((x + a) >> p) + b are closed under composition; once p exceeds 16 the function takes two values on [0, 65535]
// and is renormalised to p = 17 (or to a constant), so everything fits 32-bit registers.  A warp scan (5 rounds of 3
// shuffles) gives every lane the composition of all blocks to its left; applied to the pixel entering the warp it yields
// the lane's `left`, and the lane then runs the real 8-pixel recurrence.  The symbolic form ignores the reference's
// uint16 wrap (deltarlecompressu16.go:123); it never triggers on a stream an encoder produced (the decoder reproduces
// pixels that exist), and the numeric pass detects it: such a unit is flagged and redone by the wavefront kernel, which
// wraps pixel by pixel.
//
// Memory: D (residuals; literal pixels hold their raw value) and the literal mask M are read in the padded
// coordinates K3 wrote them in (column a_y + x with a_y = (align0 + y*W) mod 8, mic_unit.h), i.e. with aligned 16 B
// loads (two per lane when a_y != 0, re-phased by a warp-uniform funnel shift), one row ahead of use; pixels leave as
// aligned 16 B stores (chunk j = tail of block j-1 + head of block j, one 4-register shuffle when a_y != 0).
// Warps hand their last block of each row (16 B: the carry pixel and the tail the next warp stores) through a small
// shared-memory ring guarded by two monotonic row counters.
#include <cstdlib>

#include "mic_device.cuh"

#ifndef MICGPU_K4_FENCE
#define MICGPU_K4_FENCE 1
#endif

namespace micgpu {

namespace {

constexpr int PTHR = 31;       // Fn::p of the threshold form
constexpr int TINF = 1 << 20;  // threshold no pixel reaches: the function is the constant b
constexpr int HAND_R = 16;     // ring slots per warp boundary

// A function of the incoming left pixel x in [0, 65535], in one of two forms (all 32-bit):
//   shift form      p <= 15:    f(x) = ((x + a) >> p) + b,  0 <= a < 2^p
//   threshold form  p == PTHR:  f(x) = b + [x >= a]          (a = TINF: the constant b)
// A composition whose shift reaches 16 takes at most two values on [0, 65535] and is stored in threshold form.
struct Fn { int a, p, b; };

__device__ __forceinline__ int fn_apply(const Fn& f, int x) {
  const int sh = f.p == PTHR ? 0 : f.p;
  const int s = ((x + f.a) >> sh) + f.b;
  const int t = f.b + (x >= f.a ? 1 : 0);
  return f.p == PTHR ? t : s;
}

// h = g o f (f is applied first).  Branch-free: lanes hold functions of different forms.
__device__ __forceinline__ Fn fn_compose(const Fn& f, const Fn& g) {
  const bool ft = f.p == PTHR, gt = g.p == PTHR;
  const int fp = ft ? 0 : f.p, gp = gt ? 0 : g.p;
  // --- f in threshold form: g sees f.b or f.b + 1
  const int v0 = fn_apply(g, f.b), v1 = fn_apply(g, f.b + 1);
  // --- both in shift form: ((x + f.a) >> fp) + m) >> gp with m = f.b + g.a = q * 2^gp + rem
  const int m = f.b + g.a;
  const int q = m >> gp;                                  // floor
  const int rem = m - (q << gp);                          // 0 <= rem < 2^gp
  const int pp = fp + gp;                                 // <= 30
  const int aa = f.a + (rem << fp);                       // < 2^fp + 2^(fp+gp) <= 2^31 - ...
  const int q2 = (int)((unsigned)aa >> pp);               // 0 or 1
  const int r2 = aa - (q2 << pp);
  const int bb = g.b + q + q2;
  const int tss = (1 << pp) - r2;                         // pp >= 16: bb + [x >= tss]
  // --- f in shift form, g in threshold form: b + [((x + f.a) >> fp) + f.b >= g.a]  =  b + [x >= ((g.a - f.b) << fp) - f.a]
  int need = g.a - f.b;                                   // value ((x + f.a) >> fp) has to reach
  need = need < 0 ? 0 : (need > (1 << 18) >> fp ? (1 << 18) >> fp : need);   // beyond reach either way: keep the shift in range
  int tst = (need << fp) - f.a;
  tst = g.a >= TINF ? TINF : tst;
  Fn h;
  if (ft) {
    h.a = (v1 != v0) ? f.a : TINF; h.p = PTHR; h.b = v0;
  } else if (gt) {
    h.a = tst > 65535 ? TINF : tst; h.p = PTHR; h.b = g.b;
  } else if (pp >= 16) {
    h.a = tss > 65535 ? TINF : tss; h.p = PTHR; h.b = bb;
  } else {
    h.a = r2; h.p = pp; h.b = bb;
  }
  return h;
}

__device__ __forceinline__ bool fn_is_const(const Fn& f) { return f.p == PTHR && f.a >= TINF; }

__device__ __forceinline__ uint32_t fsh16(uint32_t lo, uint32_t hi) { return __funnelshift_r(lo, hi, 16); }

// concat(p, c)[8 - delta .. 16 - delta) in u16 units, four packed registers each (delta warp-uniform, 1..7)
__device__ __forceinline__ uint4 rephase8(const uint4& p, const uint4& c, int delta) {
  uint4 o;
  switch (delta) {
    case 1: o.x = fsh16(p.w, c.x); o.y = fsh16(c.x, c.y); o.z = fsh16(c.y, c.z); o.w = fsh16(c.z, c.w); break;
    case 2: o.x = p.w; o.y = c.x; o.z = c.y; o.w = c.z; break;
    case 3: o.x = fsh16(p.z, p.w); o.y = fsh16(p.w, c.x); o.z = fsh16(c.x, c.y); o.w = fsh16(c.y, c.z); break;
    case 4: o.x = p.z; o.y = p.w; o.z = c.x; o.w = c.y; break;
    case 5: o.x = fsh16(p.y, p.z); o.y = fsh16(p.z, p.w); o.z = fsh16(p.w, c.x); o.w = fsh16(c.x, c.y); break;
    case 6: o.x = p.y; o.y = p.z; o.z = p.w; o.w = c.x; break;
    case 7: o.x = fsh16(p.x, p.y); o.y = fsh16(p.y, p.z); o.z = fsh16(p.z, p.w); o.w = fsh16(p.w, c.x); break;
    default: o = c; break;
  }
  return o;
}

__device__ __forceinline__ unsigned ld_volatile_s(const unsigned* p) {
  unsigned v;
  asm volatile("ld.volatile.shared.u32 %0, [%1];" : "=r"(v) : "r"((unsigned)__cvta_generic_to_shared(p)) : "memory");
  return v;
}
// the same with the shared-window address computed once, outside the polling loop (the compiler re-derives it from the
// generic pointer in every iteration otherwise: 8 instructions per poll instead of 3)
__device__ __forceinline__ unsigned ld_volatile_sa(unsigned sa) {
  unsigned v;
  asm volatile("ld.volatile.shared.u32 %0, [%1];" : "=r"(v) : "r"(sa) : "memory");
  return v;
}
__device__ __forceinline__ void st_volatile_s(unsigned* p, unsigned v) {
  asm volatile("st.volatile.shared.u32 [%0], %1;" ::"r"((unsigned)__cvta_generic_to_shared(p)), "r"(v) : "memory");
}

}  // namespace

template <int MAXT, int MINB>
__global__ void __launch_bounds__(MAXT, MINB)
k_delta_rowscan(MicUnit* __restrict__ units, const int* __restrict__ list, int nlist, const uint16_t* __restrict__ D,
                const uint32_t* __restrict__ M, uint16_t* __restrict__ out) {
  extern __shared__ __align__(16) uint8_t s_scan_raw[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  uint4* hand = reinterpret_cast<uint4*>(s_scan_raw);                          // [nwarps][HAND_R] last block of a row
  unsigned* rows_done = reinterpret_cast<unsigned*>(hand + nwarps * HAND_R);   // [nwarps] rows finished by each warp
  if (threadIdx.x < nwarps) rows_done[threadIdx.x] = 0;
  __syncthreads();

  MicUnit* U = &units[list[blockIdx.x]];
  if (U->status != MIC_OK || U->predictor != 0u) return;   // gradient-predictor units belong to k_grad_wavefront
  const int W = (int)U->width, H = (int)U->height;
  const unsigned wp = U->wp;
  const int thr = (int)U->thr;
  const int align0 = (int)U->align0 & 7;
  const int dW = W & 7;
  const bool aligned_unit = dW == 0 && align0 == 0;
  const int nchunks = aligned_unit ? (W >> 3) : ((W + 7 + 7) >> 3);    // 16 B output chunks a row can touch
  if (warp * 32 >= nchunks) return;                                    // nothing in these columns
  const int nw_used = (nchunks + 31) >> 5;
  const bool has_next = warp + 1 < nw_used;
  const uint16_t* Du = D + U->d_off;
  const uint32_t* Mu = M + U->m_off;
  uint16_t* Ou = out + U->out_off;
  const int j = warp * 32 + lane;             // block / chunk index of this lane
  const int x0 = 8 * j;
  const int nvalid = x0 >= W ? 0 : (W - x0 < 8 ? W - x0 : 8);
  const unsigned mwords = wp >> 5;

  const unsigned rd_prev_sa = (unsigned)__cvta_generic_to_shared(&rows_done[warp > 0 ? warp - 1 : 0]);
  int top[8];
#pragma unroll
  for (int i = 0; i < 8; i++) top[i] = 0;
  unsigned wrapped = 0;

  // residual chunks and mask words of the row about to be processed (loaded one row ahead)
  uint4 nd0 = make_uint4(0, 0, 0, 0), nd1 = make_uint4(0, 0, 0, 0);
  uint32_t nm0 = 0, nm1 = 0;
  auto prefetch = [&](int y) {
    if (y < H && j < nchunks) {
      const int a = (align0 + y * dW) & 7;
      const uint4* Drow = reinterpret_cast<const uint4*>(Du + (size_t)y * wp);
      nd0 = __ldg(Drow + j);
      if (a) nd1 = __ldg(Drow + j + 1);
      const uint32_t* Mrow = Mu + (size_t)y * mwords;
      const unsigned pp = (unsigned)(x0 + a);
      nm0 = __ldg(Mrow + (pp >> 5));
      nm1 = ((pp & 31u) > 24u) ? __ldg(Mrow + (pp >> 5) + 1) : 0u;
    }
  };
  prefetch(0);

  for (int y = 0; y < H; y++) {
    const int a = (align0 + y * dW) & 7;       // phase of this row (warp-uniform: every lane is on row y)
    // ---- this row's residuals and literal bits; start the next row's loads ------------------------------------
    const uint4 dd = a ? rephase8(nd0, nd1, 8 - a) : nd0;   // concat(nd0, nd1)[a .. a + 8)
    const unsigned mbits = __funnelshift_r(nm0, nm1, (unsigned)(x0 + a) & 31u) & 0xFFu;
    prefetch(y + 1);
    int d[8];
    d[0] = dd.x & 0xFFFF; d[1] = dd.x >> 16; d[2] = dd.y & 0xFFFF; d[3] = dd.y >> 16;
    d[4] = dd.z & 0xFFFF; d[5] = dd.z >> 16; d[6] = dd.w & 0xFFFF; d[7] = dd.w >> 16;
    if (nvalid != 8) {   // last block of the row / lanes past it: columns >= W hold stale scratch; make them e = 0
#pragma unroll
      for (int i = 0; i < 8; i++) d[i] = i < nvalid ? d[i] : thr;
    }

    // ---- block function: incoming left pixel -> last pixel of the block --------------------------------------
    // pixel op: x -> ((x + top) >> hs) + e with hs = 0 on row 0 (pred = left; top[] is zero there).
    // PLAIN rows (no literal in the warp's 256 columns, no partial block, not row 0 -- almost all of them): every block is
    // the shift form with p = 8 (7 for the block that starts at column 0, whose first pixel predicts from `top` alone:
    // its incoming x is 0 by construction, so x -> x + top + e stands in for the constant), shifts are immediates, the
    // first scan round is shift o shift -> threshold and the second threshold o threshold.
    const int hs = y == 0 ? 0 : 1;
    // (a partial last block and the lanes past the row end run the plain path too: nobody consumes their function, their
    // pixels past W are never stored, and with e = 0 they stay in range)
    const bool plain = y > 0 && !__any_sync(0xffffffffu, (mbits & ((1u << nvalid) - 1u)) != 0u);
    Fn f;
    if (plain) {
      const int c0 = x0 == 0 ? 1 : 0;
      int fa = top[0];
#pragma unroll
      for (int i = 1; i < 8; i++) fa += ((d[i - 1] - thr + top[i]) << i) >> c0;   // exact: a multiple of 2
      const int pw = 8 - c0;
      const int q = fa >> pw;
      f.a = fa - (q << pw);
      f.b = d[7] - thr + q;
      f.p = pw;
      // round 1: blocks (j-1, j): ((x + a1) >> p1 + m) >> 8 with m = b1 + a2 -> threshold form (p1 + 8 >= 15)
      {
        const int ga = __shfl_up_sync(0xffffffffu, f.a, 1), gpw = __shfl_up_sync(0xffffffffu, f.p, 1), gb = __shfl_up_sync(0xffffffffu, f.b, 1);
        // own block is never the column-0 block here (lane >= 1), so its p is 8
        const int m = gb + f.a;
        const int q1 = m >> 8, rem = m - (q1 << 8);
        const int pp = gpw + 8;                              // 15 or 16
        const int aa = ga + (rem << gpw);                    // < 2^pp
        const int tss = (1 << pp) - aa;                      // value = bb + [x >= tss]   (x + aa < 2^(pp+1) for x < 2^15; see below)
        const int bb = f.b + q1;
        if (lane >= 1) {
          // pp == 15 only for the pair that starts at column 0, whose x is 0: [x + aa >= 2^15] is still exact there
          f.a = tss > 65535 ? TINF : tss; f.p = PTHR; f.b = bb;
        }
      }
      // round 2: lanes >= 3 compose threshold o threshold; lane 2 composes lane 0's shift form with its threshold form
      {
        const int ga = __shfl_up_sync(0xffffffffu, f.a, 2), gpw = __shfl_up_sync(0xffffffffu, f.p, 2), gb = __shfl_up_sync(0xffffffffu, f.b, 2);
        // g (applied first) in threshold form: own f sees gb or gb + 1
        const int v0 = f.b + (gb >= f.a ? 1 : 0), v1 = f.b + (gb + 1 >= f.a ? 1 : 0);
        // g in shift form (lane 2 only): b + [((x + ga) >> gp) + gb >= f.a]
        const int gp = gpw == PTHR ? 0 : gpw;
        int need = f.a - gb;
        need = need < 0 ? 0 : (need > ((1 << 18) >> gp) ? ((1 << 18) >> gp) : need);
        int tst = (need << gp) - ga;
        tst = (f.a >= TINF || tst > 65535) ? TINF : tst;
        if (lane >= 2) {
          const bool gthr = gpw == PTHR;
          f.a = gthr ? ((v1 != v0) ? ga : TINF) : tst;
          f.b = gthr ? v0 : f.b;
        }
      }
    } else {
      int fa = 0, fpw = 0, fb = 0;
      bool isc = false;
#pragma unroll
      for (int i = 0; i < 8; i++) {
        const bool valid = i < nvalid;
        const bool lit = (mbits >> i) & 1u;
        const int e = d[i] - thr;
        const int t = top[i];
        const bool reset = valid && (lit || (x0 + i) == 0);
        const bool step = valid && !reset && !isc;
        const int cv = ((fb + t) >> hs) + e;              // running constant
        const int na = fa + (fb + t) * (1 << fpw);        // running shift form
        fa = step ? na : fa;
        fpw = step ? fpw + hs : fpw;
        fb = reset ? (lit ? d[i] : t + e) : (valid ? (isc ? cv : e) : fb);
        isc = isc || reset;
      }
      // normalise: 0 <= a < 2^p (p <= 8 here)
      const int q = fa >> fpw;
      f.a = isc ? TINF : fa - (q << fpw);
      f.b = isc ? fb : fb + q;
      f.p = isc ? PTHR : fpw;
    }
    // ---- inclusive scan over the warp.  Two rounds compose four blocks (32 pixels): the result is almost always a
    // constant, i.e. the left pixel no longer matters, and the remaining rounds are skipped unless some lane still
    // depends on its input (row 0 always does: x -> x + b).
    auto scan_round = [&](int dist) {
      Fn g;
      g.a = __shfl_up_sync(0xffffffffu, f.a, dist);
      g.p = __shfl_up_sync(0xffffffffu, f.p, dist);
      g.b = __shfl_up_sync(0xffffffffu, f.b, dist);
      const Fn h = fn_compose(g, f);
      const bool take = lane >= dist;
      f.a = take ? h.a : f.a; f.p = take ? h.p : f.p; f.b = take ? h.b : f.b;
    };
    if (!plain) {
      scan_round(1);
      scan_round(2);
    }
    if (!__all_sync(0xffffffffu, lane < 3 || nvalid != 8 || fn_is_const(f))) {
      scan_round(4);
      scan_round(8);
      scan_round(16);
    }
    Fn ex;
    ex.a = __shfl_up_sync(0xffffffffu, f.a, 1);
    ex.p = __shfl_up_sync(0xffffffffu, f.p, 1);
    ex.b = __shfl_up_sync(0xffffffffu, f.b, 1);
    if (lane == 0) { ex.a = 0; ex.p = 0; ex.b = 0; }
    // ---- the block the previous warp finished on this row: carry pixel + the tail this warp stores --------------
    uint4 hb = make_uint4(0, 0, 0, 0);
    if (warp > 0) {
      while (ld_volatile_sa(rd_prev_sa) <= (unsigned)y) { }
      hb = hand[(warp - 1) * HAND_R + (y & (HAND_R - 1))];
    }
    int x = fn_apply(ex, (int)(hb.w >> 16));
    // ---- the real recurrence -----------------------------------------------------------------------------------
    int v[8];
    if (plain) {
#pragma unroll
      for (int i = 0; i < 8; i++) {
        const int t = top[i];
        const int pred = (i == 0 && x0 == 0) ? t : ((x + t) >> 1);
        int val = pred + d[i] - thr;
        wrapped |= (unsigned)val >> 16;
        val &= 0xFFFF;
        x = val;
        v[i] = val;
        top[i] = val;
      }
    } else {
#pragma unroll
      for (int i = 0; i < 8; i++) {
        const bool valid = i < nvalid;
        const bool lit = (mbits >> i) & 1u;
        const int t = top[i];
        const int pred = (x0 + i) == 0 ? t : ((x + t) >> hs);      // row 0: top is 0, so column 0 predicts 0
        int val = lit ? d[i] : pred + d[i] - thr;
        wrapped |= valid ? ((unsigned)val >> 16) : 0u;
        val &= 0xFFFF;
        x = valid ? val : x;
        v[i] = valid ? val : 0;
        top[i] = v[i];
      }
    }
    const uint4 blk = make_uint4((unsigned)v[0] | ((unsigned)v[1] << 16), (unsigned)v[2] | ((unsigned)v[3] << 16),
                                 (unsigned)v[4] | ((unsigned)v[5] << 16), (unsigned)v[6] | ((unsigned)v[7] << 16));
    // ---- hand the warp's last block to the next warp ----------------------------------------------------------
    if (has_next) {
      if (y >= HAND_R) {   // the slot is free once the consumer has finished row y - HAND_R
        while (ld_volatile_s(&rows_done[warp + 1]) + HAND_R <= (unsigned)y) __nanosleep(64);   // not on anybody's critical path
      }
      if (lane == 31) hand[warp * HAND_R + (y & (HAND_R - 1))] = blk;
      // the slot must be visible to the next warp before the row counter: an acquire-release fence at CTA scope is enough
      // (__threadfence_block() compiles to the sequentially consistent MEMBAR.SC.CTA, which also waits for the warp's
      // global stores in flight)
#if MICGPU_K4_FENCE == 0
      __threadfence_block();
#else
      asm volatile("fence.acq_rel.cta;" ::: "memory");
#endif
    }
    // row y of this warp is published (the next warp may read the slot; this warp has read the previous warp's slot)
    if (lane == 31) st_volatile_s(&rows_done[warp], (unsigned)y + 1u);
    // ---- store: chunk j = pixels [8j - a, 8j - a + 8) of the row ------------------------------------------------
    uint4 chunk = blk;
    if (a) {
      uint4 pb;
      pb.x = __shfl_up_sync(0xffffffffu, blk.x, 1); pb.y = __shfl_up_sync(0xffffffffu, blk.y, 1);
      pb.z = __shfl_up_sync(0xffffffffu, blk.z, 1); pb.w = __shfl_up_sync(0xffffffffu, blk.w, 1);
      if (lane == 0) pb = hb;
      chunk = rephase8(pb, blk, a);
    }
    if (j < nchunks) {
      const int xs = x0 - a;                                  // first pixel of the chunk
      uint16_t* dst = Ou + (size_t)y * W + xs;                // 16 B aligned by the definition of a
      if (xs >= 0 && xs + 8 <= W) {
        *reinterpret_cast<uint4*>(dst) = chunk;
      } else {
        const unsigned w4[4] = {chunk.x, chunk.y, chunk.z, chunk.w};
#pragma unroll
        for (int k = 0; k < 8; k++) {
          const int xx = xs + k;
          if (xx >= 0 && xx < W) dst[k] = (uint16_t)(w4[k >> 1] >> (16 * (k & 1)));
        }
      }
    }
  }
  if (__any_sync(0xffffffffu, wrapped != 0) && lane == 0) U->k4_redo = 1;
}

size_t delta_rowscan_smem_bytes(int warps) { return (size_t)warps * (HAND_R * 16 + 4) + 16; }

// one CTA per listed spatial unit; warps = ceil(chunks / 32) of the widest unit (at most 32: W <= 8178)
template <int MAXT, int MINB>
static void launch_rowscan_t(MicUnit* d_units, const int* d_list, int nlist, const uint16_t* d_D, const uint32_t* d_M, uint16_t* d_out,
                             int warps, cudaStream_t st) {
  k_delta_rowscan<MAXT, MINB><<<nlist, 32 * warps, delta_rowscan_smem_bytes(warps), st>>>(d_units, d_list, nlist, d_D, d_M, d_out);
}

bool launch_delta_rowscan(MicUnit* d_units, const int* d_list, int nlist, const uint16_t* d_D, const uint32_t* d_M,
                          uint16_t* d_out, int max_width, bool all_aligned, cudaStream_t st) {
  if (nlist <= 0) return true;
  // all_aligned: every unit has W % 8 == 0 and a 16 B aligned first pixel (MIC3 tile planes): no chunk beyond W / 8, so a
  // 256-wide tile is ONE warp per CTA (an idle second warp would halve the resident units: registers are per CTA)
  const int nchunks = all_aligned ? (max_width >> 3) : ((max_width + 14) >> 3);
  const int warps = (nchunks + 31) >> 5;
  if (warps > 32) return false;     // wider than one CTA can chain: the caller uses the wavefront kernel
  // register budget by CTA size: three 12-warp CTAs (a 2577-wide strip is 11 warps) or eight 4-warp CTAs per SM
  if (warps <= 4) launch_rowscan_t<128, 8>(d_units, d_list, nlist, d_D, d_M, d_out, warps, st);
  else if (warps <= 12) launch_rowscan_t<384, 3>(d_units, d_list, nlist, d_D, d_M, d_out, warps, st);
  else launch_rowscan_t<1024, 1>(d_units, d_list, nlist, d_D, d_M, d_out, warps, st);
  return true;
}

}  // namespace micgpu
