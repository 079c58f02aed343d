// k_enc_rle.cu -- K5: forward predictor + escape coding, and the exact RLE-u16 encoder.
//
// Replaces DeltaRleCompressU16.Compress (deltarlecompressu16.go:24-68) and the stateful buffer machine
// RleCompressU16.{Init,Encode,Flush,Compress} (rlecompressu16.go:15-93); C twin delta_rle_encode.
//
// E1 (k_enc_delta): V = [maxValue, symbols...]: per pixel pred = avg(top,left) -- or the gradient-adaptive predictor of
//   deltagradrlecompressu16.go:26-68 when MicEncUnit::predictor is 1 -- from the SOURCE pixels, so it is
//   embarrassingly parallel; an escaped pixel contributes two symbols (delim, raw), positions by block scan.
// E2 (k_enc_rle_*): the reference encoder is a serial state machine, but its output has a closed form:
//   * every maximal run of >= 3 equal symbols is coded in "same" mode, everything between two such runs is a
//     "diff" stretch (runs of 1-2 equal symbols stay literals);
//   * the forced flush at len(buf) >= midCount-1 (rlecompressu16.go:58) cuts a run of L symbols into
//     K = (L >= mid ? (L-mid)/P + 1 : 0) pieces of P = midCount-3 and a final piece of L-K*P;
//   * a diff stretch of m literals followed by a same-run sees m+2 symbols in diff mode (the run's first two
//     stay in the buffer), so K = (m+2 >= mid ? (m+2-mid)/P + 1 : 0) pieces of P literals and a final piece of
//     m-K*P; a trailing stretch (Flush) uses m instead of m+2.
//   So: flag segment starts from a 5-symbol neighbourhood, compact them, scan their output sizes, and let one
//   warp per segment write headers and payload.  tests/test_gpu_encode.py checks the streams word for word
//   against the CPU oracle on long runs, 2-runs, forced flushes and run/stretch boundaries at every residue.
#include "mic_device.cuh"
#include "mic_enc.h"

namespace micgpu {

constexpr int E_THREADS = 256;

__device__ __forceinline__ unsigned block_excl_scan256(unsigned v, unsigned* s_warp, unsigned* total) {
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
  for (int w = 0; w < E_THREADS / 32; w++) {
    const unsigned s = s_warp[w];
    if ((unsigned)w < warp) base += s;
    tot += s;
  }
  *total = tot;
  return base + inc - v;
}

// ---------------- E1: Delta + escape ------------------------------------------------------------------
__global__ void __launch_bounds__(E_THREADS)
k_enc_delta(MicEncUnit* __restrict__ units, int nunits, const uint16_t* __restrict__ src, uint16_t* __restrict__ Vbuf) {
  __shared__ unsigned s_warp[E_THREADS / 32];
  for (int ui = blockIdx.x; ui < nunits; ui += gridDim.x) {
    MicEncUnit* U = &units[ui];
    __syncthreads();
    if (U->kind != MIC_ENC_SPATIAL || U->status != MIC_ENC_OK) continue;
    const unsigned W = U->width, H = U->height;
    const unsigned long long npx = (unsigned long long)W * H;
    const uint16_t* px = src + U->src_off;
    uint16_t* V = Vbuf + U->v_off;
    const unsigned maxv = U->max_value;
    const bool grad = U->predictor == 1u;
    const int depth = 32 - __clz(maxv);
    if (maxv == 0 || maxv > 65535 || depth < 4) {   // Go: 1<<-1 panics for 0; depths 1-3 give midCount <= 3 and a broken RLE
      if (threadIdx.x == 0) U->status = MIC_ENC_UNSUPPORTED;
      continue;
    }
    const unsigned thr = (1u << (depth - 1)) - 1u, delim = (1u << depth) - 1u;
    if (threadIdx.x == 0) V[0] = (uint16_t)maxv;
    unsigned long long out = 1;
    constexpr int PER = 8;
    for (unsigned long long base = 0; base < npx; base += (unsigned long long)PER * E_THREADS) {
      const unsigned long long p0 = base + (unsigned long long)threadIdx.x * PER;
      unsigned sym[PER], raw[PER];
      unsigned esc = 0, cnt = 0;
      if (p0 < npx) {
        unsigned y = (unsigned)(p0 / W), x = (unsigned)(p0 - (unsigned long long)y * W);
#pragma unroll
        for (int q = 0; q < PER; q++) {
          const unsigned long long p = p0 + q;
          if (p < npx) {
            int pred = 0;
            if (x > 0 && y > 0) {
              if (grad) {   // gradPredict (deltagradcompressu16.go:149-166), neighbours from the source like the avg case
                const int w = px[p - 1], n = px[p - W], nw = px[p - W - 1];
                const int ne = x + 1 < W ? (int)px[p - W + 1] : nw;
                const int g = abs(w - nw) + abs(n - nw), limit = g >> 1;
                pred = ((w + n) >> 1) + min(max((ne - nw) >> 3, -limit), limit);
              } else {
                pred = ((int)px[p - 1] + (int)px[p - W]) >> 1;
              }
            }
            else if (x > 0) pred = px[p - 1];
            else if (y > 0) pred = px[p - W];
            const int v = px[p];
            const int diff = v - pred;
            const unsigned ad = (unsigned)(diff < 0 ? -diff : diff);
            if ((ad & 0xFFFFu) >= thr) { esc |= 1u << q; sym[q] = delim; raw[q] = (unsigned)v; cnt += 2; }
            else { sym[q] = (unsigned)((int)thr + diff); cnt += 1; }
            if (++x == W) { x = 0; y++; }
          }
        }
      }
      unsigned tot;
      const unsigned off = block_excl_scan256(cnt, s_warp, &tot);
      if (out + tot > U->v_cap) {
        if (threadIdx.x == 0) U->status = MIC_ENC_CAPACITY;
        break;
      }
      if (p0 < npx) {
        uint16_t* o = V + out + off;
#pragma unroll
        for (int q = 0; q < PER; q++) {
          if (p0 + q < npx) {
            *o++ = (uint16_t)sym[q];
            if (esc & (1u << q)) *o++ = (uint16_t)raw[q];
          }
        }
      }
      out += tot;
    }
    __syncthreads();
    if (threadIdx.x == 0 && U->status == MIC_ENC_OK) U->v_len = (unsigned)out;
  }
}

// ---------------- E2a: segment starts ---------------------------------------------------------------
__device__ __forceinline__ const uint16_t* enc_stream(const MicEncUnit* U, const uint16_t* src, const uint16_t* Vbuf) {
  return U->kind == MIC_ENC_SPATIAL ? Vbuf + U->v_off : src + U->src_off;
}

__global__ void __launch_bounds__(E_THREADS)
k_enc_rle_segments(MicEncUnit* __restrict__ units, int nunits, const uint16_t* __restrict__ src, const uint16_t* __restrict__ Vbuf,
                   uint32_t* __restrict__ segs) {
  __shared__ unsigned s_warp[E_THREADS / 32];
  __shared__ unsigned s_max[E_THREADS / 32];
  for (int ui = blockIdx.x; ui < nunits; ui += gridDim.x) {
    MicEncUnit* U = &units[ui];
    __syncthreads();
    if (U->status != MIC_ENC_OK) continue;
    if (U->kind == MIC_ENC_RAW) {          // no RLE layer: S = V (k_enc_rle_emit copies it)
      if (threadIdx.x == 0) { U->v_len = U->width; U->nseg = 0; U->mid = 0; }
      continue;
    }
    if (U->kind == MIC_ENC_RLE) {
      if (threadIdx.x == 0) U->v_len = U->width;
      // rleMaxVal / resMax = max over V when the caller asks for it (multiframecompress.go:194-199)
      if (U->max_value >= 0xFFFFFFFEu) {
        const bool pow2 = U->max_value == 0xFFFFFFFEu;
        const uint16_t* v = src + U->src_off;
        unsigned m = 0;
        for (unsigned i = threadIdx.x; i < U->width; i += E_THREADS) m = max(m, (unsigned)v[i]);
        m = __reduce_max_sync(0xffffffffu, m);
        if ((threadIdx.x & 31) == 0) s_max[threadIdx.x >> 5] = m;
        __syncthreads();
        if (threadIdx.x == 0) {
          for (int w = 1; w < E_THREADS / 32; w++) m = max(m, s_max[w]);
          if (pow2) { int dep = 32 - __clz(m); if (dep < 1) dep = 1; m = (1u << dep) - 1u; }
          U->max_value = m;
        }
      }
      __syncthreads();
    }
    const unsigned n = U->kind == MIC_ENC_SPATIAL ? U->v_len : U->width;
    const uint16_t* V = enc_stream(U, src, Vbuf);
    uint32_t* sg = segs + U->seg_off * 2;
    // midCount comes from word 0 of the stream: the delimiter for Delta+RLE, maxValue otherwise
    const unsigned maxv = U->max_value;
    const int depth = 32 - __clz(maxv);
    if (maxv == 0 || maxv > 65535 || depth < 4) {
      if (threadIdx.x == 0) U->status = MIC_ENC_UNSUPPORTED;
      continue;
    }
    unsigned nseg = 0;
    constexpr int PER = 8;
    for (unsigned base = 0; base < n; base += PER * E_THREADS) {
      const unsigned i0 = base + threadIdx.x * PER;
      // neighbourhood i0-2 .. i0+PER+1
      unsigned w[PER + 4];
#pragma unroll
      for (int q = 0; q < PER + 4; q++) {
        const long long i = (long long)i0 + q - 2;
        w[q] = (i >= 0 && i < (long long)n) ? V[i] : 0x10000u + (unsigned)(q & 1) + (i < 0 ? 2u : 4u);  // never equal to a symbol or each other
      }
      unsigned startmask = 0, longmask = 0, cnt = 0;
      bool prev_long = false;
      {
        // long[i0-1]
        const long long i = (long long)i0 - 1;
        if (i >= 0 && i0 < n) {   // threads past the end of the stream have nothing to flag (and must not read V[i0-3])
          // needs V[i0-3]: reload
          const unsigned a = i >= 2 ? V[i - 2] : 0x20000u, b = w[0], c = w[1], d2 = w[2], e = w[3];
          prev_long = (a == b && b == c) || (b == c && c == d2) || (c == d2 && d2 == e);
          (void)a;
        }
      }
#pragma unroll
      for (int q = 0; q < PER; q++) {
        const unsigned i = i0 + q;
        if (i < n) {
          const unsigned a = w[q], b = w[q + 1], c = w[q + 2], d2 = w[q + 3], e = w[q + 4];
          const bool lg = (a == b && b == c) || (b == c && c == d2) || (c == d2 && d2 == e);
          const bool st = (i == 0) || (lg != prev_long) || (lg && c != b);
          if (st) { startmask |= 1u << q; cnt++; }
          if (lg) longmask |= 1u << q;
          prev_long = lg;
        }
      }
      unsigned tot;
      unsigned off = block_excl_scan256(cnt, s_warp, &tot);
#pragma unroll
      for (int q = 0; q < PER; q++)
        if (startmask & (1u << q)) {
          sg[2 * (nseg + off)] = (i0 + q) | ((longmask >> q) & 1u ? 0x80000000u : 0u);
          off++;
        }
      nseg += tot;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      U->nseg = nseg;
      U->mid = (1u << (depth - 1)) - 1u;
    }
  }
}

// ---------------- E2b: segment output sizes -> offsets ------------------------------------------------
__device__ __forceinline__ unsigned seg_pieces(unsigned t, unsigned mid) {   // forced flushes among t symbols seen in one mode
  return t >= mid ? (t - mid) / (mid - 3) + 1 : 0;
}

__global__ void __launch_bounds__(E_THREADS)
k_enc_rle_offsets(MicEncUnit* __restrict__ units, int nunits, uint32_t* __restrict__ segs) {
  __shared__ unsigned s_warp[E_THREADS / 32];
  for (int ui = blockIdx.x; ui < nunits; ui += gridDim.x) {
    MicEncUnit* U = &units[ui];
    __syncthreads();
    if (U->status != MIC_ENC_OK) continue;
    if (U->kind == MIC_ENC_RAW) {
      if (threadIdx.x == 0) {
        if (U->width > U->s_cap) U->status = MIC_ENC_CAPACITY;
        U->s_len = U->width;
      }
      continue;
    }
    const unsigned n = U->v_len, nseg = U->nseg, mid = U->mid, P = mid - 3;
    uint32_t* sg = segs + U->seg_off * 2;
    unsigned out = U->kind == MIC_ENC_SPATIAL ? 1u : 3u;   // word 0 (+ the 2 length words of Compress, rlecompressu16.go:85-87)
    for (unsigned base = 0; base < nseg; base += E_THREADS) {
      const unsigned k = base + threadIdx.x;
      unsigned size = 0;
      if (k < nseg) {
        const unsigned e0 = sg[2 * k], start = e0 & 0x7fffffffu;
        const unsigned end = k + 1 < nseg ? (sg[2 * (k + 1)] & 0x7fffffffu) : n;
        const unsigned m = end - start;
        if (e0 & 0x80000000u) {
          size = 2 * (seg_pieces(m, mid) + 1);
        } else {
          const unsigned K = seg_pieces(m + (k + 1 < nseg ? 2u : 0u), mid);
          size = K * (P + 1) + (m - K * P) + 1;
        }
      }
      unsigned tot;
      const unsigned off = block_excl_scan256(size, s_warp, &tot);
      if (k < nseg) sg[2 * k + 1] = out + off;
      out += tot;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      if (out > U->s_cap) U->status = MIC_ENC_CAPACITY;
      U->s_len = out;
    }
  }
}

// ---------------- E2c: emit -----------------------------------------------------------------------------
__global__ void __launch_bounds__(E_THREADS)
k_enc_rle_emit(MicEncUnit* __restrict__ units, int nunits, const uint16_t* __restrict__ src, const uint16_t* __restrict__ Vbuf,
               const uint32_t* __restrict__ segs, uint16_t* __restrict__ Sbuf) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int ui = blockIdx.x; ui < nunits; ui += gridDim.x) {
    const MicEncUnit* U = &units[ui];
    if (U->status != MIC_ENC_OK) continue;
    const unsigned n = U->v_len, nseg = U->nseg, mid = U->mid, P = mid - 3;
    const uint16_t* V = enc_stream(U, src, Vbuf);
    const uint32_t* sg = segs + U->seg_off * 2;
    uint16_t* S = Sbuf + U->s_off;
    if (U->kind == MIC_ENC_RAW) {
      for (unsigned i = threadIdx.x; i < U->width; i += E_THREADS) S[i] = V[i];
      continue;
    }
    if (threadIdx.x == 0) {
      if (U->kind == MIC_ENC_SPATIAL) {
        S[0] = (uint16_t)((1u << (32 - __clz(U->max_value))) - 1u);          // Init(width,height,delimiter)
      } else {
        S[0] = (uint16_t)U->max_value; S[1] = (uint16_t)(n >> 16); S[2] = (uint16_t)n;
      }
    }
    for (unsigned k = warp; k < nseg; k += E_THREADS / 32) {
      const unsigned e0 = sg[2 * k], start = e0 & 0x7fffffffu, o = sg[2 * k + 1];
      const unsigned end = k + 1 < nseg ? (sg[2 * (k + 1)] & 0x7fffffffu) : n;
      const unsigned m = end - start;
      if (e0 & 0x80000000u) {
        const unsigned K = seg_pieces(m, mid);
        const uint16_t val = V[start];
        for (unsigned q = lane; q <= K; q += 32) {
          S[o + 2 * q] = (uint16_t)(q < K ? P : m - K * P);
          S[o + 2 * q + 1] = val;
        }
      } else {
        const unsigned K = seg_pieces(m + (k + 1 < nseg ? 2u : 0u), mid);
        for (unsigned q = lane; q <= K; q += 32) S[o + q * (P + 1)] = (uint16_t)(mid + (q < K ? P : m - K * P));
        for (unsigned j = lane; j < m; j += 32) {
          const unsigned q = min(j / P, K);
          S[o + q * (P + 1) + 1 + (j - q * P)] = V[start + j];
        }
      }
    }
  }
}

void launch_enc_delta_rle(MicEncUnit* d_units, int nunits, const uint16_t* d_src, uint16_t* d_V, uint32_t* d_segs, uint16_t* d_S,
                          int grid, cudaStream_t st) {
  if (nunits <= 0) return;
  k_enc_delta<<<grid, E_THREADS, 0, st>>>(d_units, nunits, d_src, d_V);
  k_enc_rle_segments<<<grid, E_THREADS, 0, st>>>(d_units, nunits, d_src, d_V, d_segs);
  k_enc_rle_offsets<<<grid, E_THREADS, 0, st>>>(d_units, nunits, d_segs);
  k_enc_rle_emit<<<grid, E_THREADS, 0, st>>>(d_units, nunits, d_src, d_V, d_segs, d_S);
}

}  // namespace micgpu
