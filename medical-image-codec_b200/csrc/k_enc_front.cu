// k_enc_front.cu -- encode-side front ends: temporal residuals (K10), YCoCg-R forward + tile extraction (K8, K11),
// 2x2 box pyramid (K11), 5/3 forward lifting + subband gather + coefficient escape coding (K9), plane statistics
// and the byte gather that compacts finished frames.
#include "mic_device.cuh"

namespace micgpu {

// ---- K10 forward: TemporalDeltaEncode (temporaldelta.go:11-23): ZigZag(int16(cur - prev)) -------------------------
// Two uint16 lanes per word: lane-wise subtract mod 2^16 and lane-wise ZigZag of the int16 difference.
__device__ __forceinline__ uint32_t sub2_u16(uint32_t a, uint32_t b) {
  return ((a | 0x80008000u) - (b & 0x7FFF7FFFu)) ^ ((a ^ ~b) & 0x80008000u);
}
__device__ __forceinline__ uint32_t zigzag2_i16(uint32_t d) {
  return ((d << 1) & 0xFFFEFFFEu) ^ (((d >> 15) & 0x00010001u) * 0xFFFFu);   // (d << 1) ^ (d >> 15) in both halves
}

__global__ void __launch_bounds__(256)
k_temporal_residual(const uint16_t* __restrict__ frames, uint16_t* __restrict__ res, unsigned long long fpx, int nframes, int vec) {
  const unsigned long long total = fpx * (unsigned long long)(nframes - 1);
  if (vec) {   // eight residuals per thread: two 16 B loads, one 16 B store (frame size and pointers are 16 B multiples)
    const unsigned long long ngrp = total >> 3, gpf = fpx >> 3;
    const uint4* F = reinterpret_cast<const uint4*>(frames);
    uint4* R = reinterpret_cast<uint4*>(res);
    for (unsigned long long gi = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; gi < ngrp; gi += (unsigned long long)gridDim.x * blockDim.x) {
      const uint4 p = __ldg(F + gi), c = __ldg(F + gi + gpf);
      R[gi] = make_uint4(zigzag2_i16(sub2_u16(c.x, p.x)), zigzag2_i16(sub2_u16(c.y, p.y)), zigzag2_i16(sub2_u16(c.z, p.z)), zigzag2_i16(sub2_u16(c.w, p.w)));
    }
    return;
  }
  for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (unsigned long long)gridDim.x * blockDim.x) {
    const int d = (int)(short)((int)frames[i + fpx] - (int)frames[i]);
    res[i] = (uint16_t)((d << 1) ^ (d >> 15));
  }
}

// ---- K11: Downsample2xRGB / Downsample2xGrey (wsipyramid.go:10-55) ---------------------------------------------------
// `ch` interleaved 8-bit channels, or one 16-bit channel when bytes_per_sample == 2.
__global__ void __launch_bounds__(256)
k_downsample2x(const uint8_t* __restrict__ src, uint8_t* __restrict__ dst, unsigned w, unsigned h, unsigned ch, unsigned bytes_per_sample) {
  const unsigned nw = w / 2, nh = h / 2;
  const unsigned long long total = (unsigned long long)nw * nh * ch;
  for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (unsigned long long)gridDim.x * blockDim.x) {
    const unsigned c = (unsigned)(i % ch);
    const unsigned long long p = i / ch;
    const unsigned x = (unsigned)(p % nw), y = (unsigned)(p / nw);
    const unsigned long long s00 = ((unsigned long long)(2 * y) * w + 2 * x) * ch + c, s10 = s00 + ch, s01 = s00 + (unsigned long long)w * ch, s11 = s01 + ch;
    if (bytes_per_sample == 1) {
      dst[i] = (uint8_t)(((unsigned)src[s00] + src[s10] + src[s01] + src[s11] + 2) / 4);
    } else {
      const uint16_t* s16 = reinterpret_cast<const uint16_t*>(src);
      reinterpret_cast<uint16_t*>(dst)[i] = (uint16_t)(((unsigned)s16[s00] + s16[s10] + s16[s01] + s16[s11] + 2) / 4);
    }
  }
}

// ---- K8 forward + extractTileRGB (wsicompress.go:529-555, ycocgr.go:19-26, asm_generic.go:25-38) ------------------
// One job = one tile: reads the (zero padded) tile from a level image and writes its planes (u16, tile_w*tile_h each).
__global__ void __launch_bounds__(256)
k_tile_planes(const TilePlaneJob* __restrict__ jobs, int njobs, const uint8_t* __restrict__ images, uint16_t* __restrict__ planes) {
  for (int j = blockIdx.y; j < njobs; j += gridDim.y) {
    const TilePlaneJob J = jobs[j];
    const unsigned long long npx = (unsigned long long)J.tile_w * J.tile_h;
    const uint8_t* img = images + J.img_off;
    uint16_t* P = planes + J.plane_off;
    for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < npx; i += (unsigned long long)gridDim.x * blockDim.x) {
      const unsigned ty = (unsigned)(i / J.tile_w), tx = (unsigned)(i - (unsigned long long)ty * J.tile_w);
      const unsigned sx = J.x0 + tx, sy = J.y0 + ty;
      const bool in = sx < J.img_w && sy < J.img_h;
      const unsigned long long s = (unsigned long long)sy * J.img_w + sx;
      if (J.mode <= 1) {
        int r = 0, g = 0, b = 0;
        if (in) { r = img[3 * s]; g = img[3 * s + 1]; b = img[3 * s + 2]; }
        if (J.mode == 0) {
          const int co = r - b, t = b + (co >> 1), cg = g - t, yv = t + (cg >> 1);
          P[i] = (uint16_t)yv;
          P[npx + i] = (uint16_t)((co << 1) ^ (co >> 15));      // ZigZag(int16) (deltazigzagcompressu16.go:108-111)
          P[2 * npx + i] = (uint16_t)((cg << 1) ^ (cg >> 15));
        } else {
          P[i] = (uint16_t)r; P[npx + i] = (uint16_t)g; P[2 * npx + i] = (uint16_t)b;
        }
      } else if (J.mode == 2) {
        P[i] = in ? img[s] : 0;
      } else {
        P[i] = in ? (uint16_t)(img[2 * s] | (img[2 * s + 1] << 8)) : 0;
      }
    }
  }
}

// ---- plane statistics: max and "all equal to element 0" (compressWSIPlane wsicompress.go:374-385) ------------------
__global__ void __launch_bounds__(256)
k_plane_stats(const uint16_t* __restrict__ planes, const unsigned long long* __restrict__ offs, unsigned long long npx, int nplanes,
              unsigned* __restrict__ stats /* [2*nplanes]: max, not_constant */) {
  const int p = blockIdx.y;
  if (p >= nplanes) return;
  const uint16_t* P = planes + offs[p];
  const unsigned first = P[0];
  unsigned mx = 0, diff = 0;
  for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < npx; i += (unsigned long long)gridDim.x * blockDim.x) {
    const unsigned v = P[i];
    mx = max(mx, v);
    diff |= (v != first);
  }
  mx = __reduce_max_sync(0xffffffffu, mx);
  diff = __reduce_or_sync(0xffffffffu, diff);
  if ((threadIdx.x & 31) == 0) {
    atomicMax(&stats[2 * p], mx);
    if (diff) atomicOr(&stats[2 * p + 1], 1u);
  }
}

// ---- K9 forward: wt53Forward2DSeparated (waveletu16.go:26-71,162-208), element-wise ---------------------------------
// predict d[j] = x[2j+1] - ((x[2j] + xr) >> 1), xr = x[2j+2] or x[2j] at the right edge
// update  s[m] = x[2m] + ((dL + dR + 2) >> 2), dR = d[m] (or d[m-1], or 0), dL = d[m-1] (or dR)
struct FwdLift {
  template <typename F>
  __device__ static void pair(F X, unsigned n, unsigned m, int* s_out, int* d_out, bool* has_d) {
    const unsigned n_high = n / 2;
    auto D = [&](unsigned j) {
      const int l = X(2 * j), r = (2 * j + 2 < n) ? X(2 * j + 2) : l;
      return X(2 * j + 1) - ((l + r) >> 1);
    };
    const bool hc = m < n_high, hp = m > 0;
    const int dc = hc ? D(m) : 0, dp = hp ? D(m - 1) : 0;
    const int dR = hc ? dc : (hp ? dp : 0);
    const int dL = hp ? dp : dR;
    *s_out = X(2 * m) + ((dL + dR + 2) >> 2);
    *d_out = dc;
    *has_d = hc;
  }
};

// rows: A -> B over the r x c corner, output de-interleaved [low | high] along x
__global__ void __launch_bounds__(256)
k_wt53_fwd_rows(const int32_t* __restrict__ A, int32_t* __restrict__ B, unsigned r, unsigned c, unsigned cols, unsigned long long img_stride) {
  const unsigned m = blockIdx.x * blockDim.x + threadIdx.x;
  const unsigned y = blockIdx.y;
  const unsigned n_low = (c + 1) / 2;
  if (y >= r || m >= n_low) return;
  const int32_t* in = A + (unsigned long long)blockIdx.z * img_stride + (unsigned long long)y * cols;
  int32_t* out = B + (unsigned long long)blockIdx.z * img_stride + (unsigned long long)y * cols;
  int s, d; bool hd;
  FwdLift::pair([&](unsigned i) { return in[i]; }, c, m, &s, &d, &hd);
  out[m] = s;
  if (hd) out[n_low + m] = d;
}

// columns: B -> A over the r x c corner, output de-interleaved along y
__global__ void __launch_bounds__(256)
k_wt53_fwd_cols(const int32_t* __restrict__ B, int32_t* __restrict__ A, unsigned r, unsigned c, unsigned cols, unsigned long long img_stride) {
  const unsigned x = blockIdx.x * blockDim.x + threadIdx.x;
  const unsigned m = blockIdx.y;
  const unsigned n_low = (r + 1) / 2;
  if (x >= c || m >= n_low) return;
  const int32_t* in = B + (unsigned long long)blockIdx.z * img_stride;
  int32_t* out = A + (unsigned long long)blockIdx.z * img_stride;
  int s, d; bool hd;
  FwdLift::pair([&](unsigned i) { return in[(unsigned long long)i * cols + x]; }, r, m, &s, &d, &hd);
  out[(unsigned long long)m * cols + x] = s;
  if (hd) out[(unsigned long long)(n_low + m) * cols + x] = d;
}

__global__ void __launch_bounds__(256)
k_u16_to_i32(const uint16_t* __restrict__ px, int32_t* __restrict__ A, unsigned long long n) {
  for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (unsigned long long)gridDim.x * blockDim.x) A[i] = px[i];
}

__device__ __forceinline__ void seg_locate_enc(const WaveletGeom& G, unsigned k, unsigned* y, unsigned* x) {
  int s = 0;
#pragma unroll 1
  for (int i = 1; i < G.nseg; i++)
    if (k >= G.seg_start[i]) s = i;
  const unsigned r = k - G.seg_start[s], w = G.seg_w[s], ry = r / w;
  *y = G.seg_y0[s] + ry;
  *x = G.seg_x0[s] + (r - ry * w);
}

// collectSubbandOrder + waveletCoeffsToU16 (waveletfsecompressu16.go:28-41,202-241).  A coefficient beyond +-32767 becomes
// the triple 65535, hi16, lo16 and shifts everything after it, so output positions are a prefix sum -- but such
// coefficients are rare (16-bit data, several levels).  Three kernels: a flag pass (any escape in the image?  the flag is
// parked in MicEncUnit::v_len, which the RLE stage only writes later), the plain pass for images without one (word k is
// coefficient k of the subband order: every CTA of the grid works, eight coefficients per thread), and the chunked
// one-CTA-per-image scan for the images that do have one (it was the only path: 58 ms for two mammograms).
__global__ void __launch_bounds__(256)
k_wavelet_pack_flag(const int32_t* __restrict__ A, MicEncUnit* __restrict__ units, const int* __restrict__ unit_of_img, unsigned total) {
  const int img = blockIdx.y;
  const int32_t* in = A + (unsigned long long)img * total;
  int has = 0;
  for (unsigned k = blockIdx.x * blockDim.x + threadIdx.x; k < total; k += gridDim.x * blockDim.x) {
    const int v = in[k];
    has |= (v < -32767 || v > 32767);
  }
  if (__syncthreads_or(has) && threadIdx.x == 0) atomicOr(&units[unit_of_img[img]].v_len, 1u);
}

__global__ void __launch_bounds__(256)
k_wavelet_pack_plain(const int32_t* __restrict__ A, uint16_t* __restrict__ Vbuf, MicEncUnit* __restrict__ units,
                     const int* __restrict__ unit_of_img, WaveletGeom G) {
  const int img = blockIdx.y;
  MicEncUnit* U = &units[unit_of_img[img]];
  if (U->v_len) return;                      // has escapes: k_wavelet_pack
  const unsigned total = G.rows * G.cols;
  const int32_t* in = A + (unsigned long long)img * total;
  uint16_t* V = Vbuf + U->src_off;
  for (unsigned k0 = (blockIdx.x * blockDim.x + threadIdx.x) * 8u; k0 < total; k0 += gridDim.x * blockDim.x * 8u) {
    // eight consecutive words of the subband order: locate the first, then walk along the subband row
    unsigned y, x;
    seg_locate_enc(G, k0, &y, &x);
    int sgm = 0;
#pragma unroll 1
    for (int i = 1; i < G.nseg; i++)
      if (k0 >= G.seg_start[i]) sgm = i;
    unsigned xend = G.seg_x0[sgm] + G.seg_w[sgm];
    unsigned send = sgm + 1 < G.nseg ? G.seg_start[sgm + 1] : total;
    const unsigned n = min(8u, total - k0);
    for (unsigned q = 0; q < n; q++) {
      const unsigned k = k0 + q;
      if (k >= send) {                        // next subband (empty subbands share a start: re-locate)
        seg_locate_enc(G, k, &y, &x);
        sgm = 0;
#pragma unroll 1
        for (int i = 1; i < G.nseg; i++)
          if (k >= G.seg_start[i]) sgm = i;
        xend = G.seg_x0[sgm] + G.seg_w[sgm];
        send = sgm + 1 < G.nseg ? G.seg_start[sgm + 1] : total;
      }
      const int v = in[(unsigned long long)y * G.cols + x];
      V[k] = (uint16_t)((v >> 31) ^ (v << 1));          // zigzagEncode16
      if (++x >= xend) { x = G.seg_x0[sgm]; y++; }
    }
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) U->width = total;   // actual length of V; the RLE stage reads it from here
}

__global__ void __launch_bounds__(256)
k_wavelet_pack(const int32_t* __restrict__ A, uint16_t* __restrict__ Vbuf, MicEncUnit* __restrict__ units, const int* __restrict__ unit_of_img,
               WaveletGeom G) {
  __shared__ unsigned s_warp[8];
  const int img = blockIdx.x;
  MicEncUnit* U = &units[unit_of_img[img]];
  if (!U->v_len) return;                     // no escapes: k_wavelet_pack_plain wrote the stream
  const unsigned total = G.rows * G.cols;
  const int32_t* in = A + (unsigned long long)img * total;
  uint16_t* V = Vbuf + U->src_off;
  unsigned out = 0;
  const unsigned cap = U->width;     // capacity reserved for V (3 words per coefficient worst case)
  for (unsigned base = 0; base < total; base += 256) {
    const unsigned k = base + threadIdx.x;
    int v = 0;
    unsigned cnt = 0;
    if (k < total) {
      unsigned y, x;
      seg_locate_enc(G, k, &y, &x);
      v = in[(unsigned long long)y * G.cols + x];
      cnt = (v >= -32767 && v <= 32767) ? 1u : 3u;
    }
    // block exclusive scan
    const unsigned lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    unsigned inc = cnt;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) { const unsigned t = __shfl_up_sync(0xffffffffu, inc, d); if (lane >= (unsigned)d) inc += t; }
    __syncthreads();
    if (lane == 31) s_warp[warp] = inc;
    __syncthreads();
    unsigned pre = 0, tot = 0;
#pragma unroll
    for (int w = 0; w < 8; w++) { const unsigned s = s_warp[w]; if ((unsigned)w < warp) pre += s; tot += s; }
    const unsigned o = out + pre + inc - cnt;
    if (cnt && o + cnt <= cap) {
      if (cnt == 1) V[o] = (uint16_t)((v >> 31) ^ (v << 1));          // zigzagEncode16
      else { V[o] = 65535; V[o + 1] = (uint16_t)((unsigned)v >> 16); V[o + 2] = (uint16_t)(unsigned)v; }
    }
    out += tot;
  }
  if (threadIdx.x == 0) {
    if (out > cap) U->status = MIC_ENC_CAPACITY;
    U->width = out;      // actual length of V; the RLE stage reads it from here
  }
}

// ---- byte gather: compacts finished frames / blobs into one contiguous buffer --------------------------------------
__global__ void __launch_bounds__(256)
k_gather_bytes(const GatherJob* __restrict__ jobs, int njobs, const uint8_t* __restrict__ src, uint8_t* __restrict__ dst) {
  for (int j = blockIdx.y; j < njobs; j += gridDim.y) {
    const GatherJob J = jobs[j];
    const uint8_t* s = src + J.src_off;
    uint8_t* d = dst + J.dst_off;
    for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < J.len; i += (unsigned long long)gridDim.x * blockDim.x) d[i] = s[i];
  }
}

// ---- raw plane (mode 3): u16 -> little-endian bytes -----------------------------------------------------------------
__global__ void __launch_bounds__(256)
k_plane_raw(const GatherJob* __restrict__ jobs, int njobs, const uint16_t* __restrict__ planes, uint8_t* __restrict__ dst) {
  for (int j = blockIdx.y; j < njobs; j += gridDim.y) {
    const GatherJob J = jobs[j];                 // src_off: element offset of the plane, len: samples
    const uint16_t* s = planes + J.src_off;
    uint8_t* d = dst + J.dst_off;
    for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < J.len; i += (unsigned long long)gridDim.x * blockDim.x) {
      const unsigned v = s[i];
      d[2 * i] = (uint8_t)v; d[2 * i + 1] = (uint8_t)(v >> 8);
    }
  }
}

void launch_temporal_residual(const uint16_t* d_frames, uint16_t* d_res, unsigned long long fpx, int nframes, int sm_count, cudaStream_t st) {
  if (nframes <= 1 || fpx == 0) return;
  const int vec = (fpx % 8 == 0) && (reinterpret_cast<uintptr_t>(d_frames) % 16 == 0) && (reinterpret_cast<uintptr_t>(d_res) % 16 == 0);
  k_temporal_residual<<<sm_count * 8, 256, 0, st>>>(d_frames, d_res, fpx, nframes, vec);
}
void launch_downsample2x(const uint8_t* d_src, uint8_t* d_dst, unsigned w, unsigned h, unsigned ch, unsigned bytes_per_sample, int sm_count, cudaStream_t st) {
  if (w / 2 == 0 || h / 2 == 0) return;
  k_downsample2x<<<sm_count * 8, 256, 0, st>>>(d_src, d_dst, w, h, ch, bytes_per_sample);
}
void launch_tile_planes(const TilePlaneJob* d_jobs, int njobs, const uint8_t* d_images, uint16_t* d_planes, cudaStream_t st) {
  if (njobs <= 0) return;
  k_tile_planes<<<dim3(8, njobs < 65535 ? njobs : 65535), 256, 0, st>>>(d_jobs, njobs, d_images, d_planes);
}
void launch_plane_stats(const uint16_t* d_planes, const unsigned long long* d_offs, unsigned long long npx, int nplanes, unsigned* d_stats, cudaStream_t st) {
  if (nplanes <= 0) return;
  cudaMemsetAsync(d_stats, 0, (size_t)nplanes * 2 * sizeof(unsigned), st);
  for (int p0 = 0; p0 < nplanes; p0 += 65535) {
    const int np = nplanes - p0 < 65535 ? nplanes - p0 : 65535;
    k_plane_stats<<<dim3(4, np), 256, 0, st>>>(d_planes, d_offs + p0, npx, np, d_stats + 2 * p0);
  }
}
void launch_wavelet_forward(const uint16_t* d_px, int32_t* d_A, int32_t* d_B, int nimg, const WaveletGeom& G, int sm_count, cudaStream_t st) {
  const unsigned total = G.rows * G.cols;
  k_u16_to_i32<<<sm_count * 8, 256, 0, st>>>(d_px, d_A, (unsigned long long)total * nimg);
  unsigned r = G.rows, c = G.cols;
  for (int l = 0; l < G.levels; l++) {
    k_wt53_fwd_rows<<<dim3(((c + 1) / 2 + 255) / 256, r, nimg), 256, 0, st>>>(d_A, d_B, r, c, G.cols, total);
    k_wt53_fwd_cols<<<dim3((c + 255) / 256, (r + 1) / 2, nimg), 256, 0, st>>>(d_B, d_A, r, c, G.cols, total);
    r = (r + 1) / 2;
    c = (c + 1) / 2;
  }
}
// V1 layouts (waveletForward2DRegion, waveletfsecompressu16.go:167-177): wt53Forward1D (waveletu16.go:26-73) IN PLACE on
// interleaved samples (even = low, odd = high), every row of the r x c corner, then every column; the next level takes the
// top-left (r+1)/2 x (c+1)/2 corner of that interleaved buffer.  Out of place here, one output sample per thread:
//   d[2i+1] = x[2i+1] - ((x[2i] + x[2i+2]) >> 1)            (x[2i] again at the right edge)
//   s[2i]   = x[2i]   + ((dL + dR + 2) >> 2)                dR = d[2i+1] (d[2i-1] at the end of an odd-length signal, 0 if alone), dL = d[2i-1] (dR at i == 0)
// AXIS 0: along columns (stride = row pitch), AXIS 1: along rows.
template <int AXIS>
__global__ void __launch_bounds__(256)
k_wt53_il_fwd(const int32_t* __restrict__ A, int32_t* __restrict__ B, unsigned r, unsigned c, unsigned cols, unsigned long long img_stride) {
  const unsigned x = blockIdx.x * blockDim.x + threadIdx.x;
  const unsigned y = blockIdx.y;
  if (x >= c || y >= r) return;
  const int32_t* in = A + (unsigned long long)blockIdx.z * img_stride;
  int32_t* out = B + (unsigned long long)blockIdx.z * img_stride;
  const unsigned n = AXIS == 0 ? r : c;
  const unsigned p = AXIS == 0 ? y : x;
  auto X = [&](unsigned q) -> int {
    return AXIS == 0 ? in[(unsigned long long)q * cols + x] : in[(unsigned long long)y * cols + q];
  };
  auto odd = [&](unsigned q) -> int {              // high-pass sample at odd position q
    const int l = X(q - 1), rr = q + 1 < n ? X(q + 1) : l;
    return X(q) - ((l + rr) >> 1);
  };
  int res;
  if (n < 2) res = X(p);
  else if (p & 1u) res = odd(p);
  else {
    int dR;
    if (p + 1 < n) dR = odd(p + 1);
    else dR = p > 0 ? odd(p - 1) : 0;
    const int dL = p > 0 ? odd(p - 1) : dR;
    res = X(p) + ((dL + dR + 2) >> 2);
  }
  out[(unsigned long long)y * cols + x] = res;
}

void launch_wavelet_forward_v1(const uint16_t* d_px, int32_t* d_A, int32_t* d_B, int nimg, unsigned rows, unsigned cols, int levels, int sm_count,
                               cudaStream_t st) {
  const unsigned total = rows * cols;
  k_u16_to_i32<<<sm_count * 8, 256, 0, st>>>(d_px, d_A, (unsigned long long)total * nimg);
  unsigned r = rows, c = cols;
  for (int l = 0; l < levels; l++) {
    const dim3 grid((c + 255) / 256, r, nimg);
    k_wt53_il_fwd<1><<<grid, 256, 0, st>>>(d_A, d_B, r, c, cols, total);     // rows: A -> B (the corner only)
    k_wt53_il_fwd<0><<<grid, 256, 0, st>>>(d_B, d_A, r, c, cols, total);     // columns: B -> A
    r = (r + 1) / 2;
    c = (c + 1) / 2;
  }
}
void launch_wavelet_pack(const int32_t* d_A, uint16_t* d_V, MicEncUnit* d_units, const int* d_unit_of_img, int nimg, const WaveletGeom& G, cudaStream_t st) {
  if (nimg <= 0) return;
  const unsigned total = G.rows * G.cols;
  const unsigned fb = (total + 256 * 16 - 1) / (256 * 16);
  k_wavelet_pack_flag<<<dim3(fb < 32 ? 32 : (fb > 1024 ? 1024 : fb), nimg), 256, 0, st>>>(d_A, d_units, d_unit_of_img, total);
  const unsigned blocks = (total / 8 + 255) / 256;
  k_wavelet_pack_plain<<<dim3(blocks < 2048 ? (blocks ? blocks : 1) : 2048, nimg), 256, 0, st>>>(d_A, d_V, d_units, d_unit_of_img, G);
  k_wavelet_pack<<<nimg, 256, 0, st>>>(d_A, d_V, d_units, d_unit_of_img, G);
}
void launch_gather_bytes(const GatherJob* d_jobs, int njobs, const uint8_t* d_src, uint8_t* d_dst, cudaStream_t st) {
  if (njobs <= 0) return;
  k_gather_bytes<<<dim3(4, njobs < 65535 ? njobs : 65535), 256, 0, st>>>(d_jobs, njobs, d_src, d_dst);
}
void launch_plane_raw(const GatherJob* d_jobs, int njobs, const uint16_t* d_planes, uint8_t* d_dst, cudaStream_t st) {
  if (njobs <= 0) return;
  k_plane_raw<<<dim3(4, njobs < 65535 ? njobs : 65535), 256, 0, st>>>(d_jobs, njobs, d_planes, d_dst);
}

}  // namespace micgpu
