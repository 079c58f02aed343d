// k_wavelet.cu -- K9 (decode half): WaveletV2 coefficient unpack, subband scatter and inverse 5/3 lifting.
//
// Replaces u16ToWaveletCoeffs + zigzagDecode16 (waveletfsecompressu16.go:45-58,543-546),
// scatterSubbandOrder (:243-282), wt53Inverse2DSeparated[SIMD] (waveletu16.go:212-257,401-508) and the
// AVX2 wt53Inv{Predict,Update} kernels.  Lifting is written element-wise: every output pair of a row
// (or column) is computed by one thread from five neighbouring inputs (the two update results it
// needs are recomputed instead of communicated), so both passes are plain coalesced streaming
// kernels with no serial walk; levels ping-pong between two int32 planes.
#include <cuda.h>     // CUtensorMap types only: the encoder is fetched through cudaGetDriverEntryPoint, nothing links against libcuda

#include <cstdlib>

#include "mic_device.cuh"

namespace micgpu {

// -------- coefficient stream -> Mallat layout ------------------------------------------------------
// Subband order (collectSubbandOrder :202-241): LL of the coarsest level row-major, then per level
// coarse->fine HL, LH, HH.  seg[] holds, for each of the 1+3*levels segments, its first linear index and
// its rectangle; built on the host (at most 25 entries).
__global__ void __launch_bounds__(256)
k_wavelet_has_escape(const MicUnit* __restrict__ units, const int* __restrict__ unit_of_img, const uint16_t* __restrict__ stream,
                     int* __restrict__ flags) {
  const int img = blockIdx.y;
  const MicUnit* U = &units[unit_of_img[img]];
  if (U->status != MIC_OK) return;
  const uint16_t* e = stream + U->out_off;
  const unsigned n = U->thr;   // expanded length (K3 stores it in thr for RLE-kind units)
  int has = 0;
  for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) has |= (e[i] == 65535);
  if (__syncthreads_or(has) && threadIdx.x == 0) atomicOr(&flags[img], 1);
}

__device__ __forceinline__ void seg_locate(const WaveletGeom& G, unsigned k, unsigned* y, unsigned* x) {
  int s = 0;
#pragma unroll 1
  for (int i = 1; i < G.nseg; i++)
    if (k >= G.seg_start[i]) s = i;
  const unsigned r = k - G.seg_start[s];
  const unsigned w = G.seg_w[s];
  const unsigned ry = r / w;
  *y = G.seg_y0[s] + ry;
  *x = G.seg_x0[s] + (r - ry * w);
}

// ALT: subbands of level k go to plane (k & 1) -- the layout the fused per-level kernel below wants (a level reads its
// four quadrants from one plane and writes the other, where the result is the LL quadrant of the next finer level).
template <bool ALT>
__global__ void __launch_bounds__(256)
k_wavelet_scatter(MicUnit* __restrict__ units, const int* __restrict__ unit_of_img, const uint16_t* __restrict__ stream,
                  const int* __restrict__ flags, int32_t* __restrict__ planeA, int32_t* __restrict__ planeB, WaveletGeom G) {
  const int img = blockIdx.y;
  MicUnit* U = &units[unit_of_img[img]];
  if (U->status != MIC_OK) return;
  const uint16_t* e = stream + U->out_off;
  const unsigned n = U->thr;
  const unsigned total = G.rows * G.cols;
  int32_t* A = planeA + (unsigned long long)img * total;
  int32_t* Bp = planeB + (unsigned long long)img * total;
  // segment s: 0 = LL of the coarsest level (levels-1); 1 + 3j .. 3 + 3j = HL, LH, HH of level levels-1-j
  auto plane_of = [&](unsigned k) -> int32_t* {
    if (!ALT) return A;
    int sgm = 0;
#pragma unroll 1
    for (int i = 1; i < G.nseg; i++)
      if (k >= G.seg_start[i]) sgm = i;
    const int level = sgm == 0 ? G.levels - 1 : G.levels - 1 - (sgm - 1) / 3;
    return (level & 1) ? Bp : A;
  };
  if (!flags[img]) {
    // no escape triples: coefficient k is word k (u16ToWaveletCoeffs fast case)
    if (n < total) {
      if (blockIdx.x == 0 && threadIdx.x == 0) U->status = MIC_E_SIZE;
      return;
    }
    // eight consecutive words of the subband order per thread step: one segment search, then a walk along the subband
    // row (the search is ~50 instructions; per element it made this kernel issue bound at 5 % of the DRAM rate)
    for (unsigned k0 = (blockIdx.x * blockDim.x + threadIdx.x) * 8u; k0 < total; k0 += gridDim.x * blockDim.x * 8u) {
      int sgm = 0;
#pragma unroll 1
      for (int i = 1; i < G.nseg; i++)
        if (k0 >= G.seg_start[i]) sgm = i;
      unsigned y, x;
      seg_locate(G, k0, &y, &x);
      unsigned xend = G.seg_x0[sgm] + G.seg_w[sgm];
      unsigned send = sgm + 1 < G.nseg ? G.seg_start[sgm + 1] : total;
      int32_t* pl = plane_of(k0);
      const unsigned n = min(8u, total - k0);
      unsigned w8[8];
      if (n == 8 && ((reinterpret_cast<uintptr_t>(e + k0) & 15u) == 0)) {
        const uint4 v = *reinterpret_cast<const uint4*>(e + k0);
        w8[0] = v.x & 0xFFFFu; w8[1] = v.x >> 16; w8[2] = v.y & 0xFFFFu; w8[3] = v.y >> 16;
        w8[4] = v.z & 0xFFFFu; w8[5] = v.z >> 16; w8[6] = v.w & 0xFFFFu; w8[7] = v.w >> 16;
      } else {
#pragma unroll
        for (unsigned q = 0; q < 8; q++) w8[q] = q < n ? e[k0 + q] : 0u;
      }
#pragma unroll
      for (unsigned q = 0; q < 8; q++) {
        if (q < n) {
          const unsigned k = k0 + q;
          if (k >= send) {                      // next subband (empty subbands share a start: search again)
            sgm = 0;
#pragma unroll 1
            for (int i = 1; i < G.nseg; i++)
              if (k >= G.seg_start[i]) sgm = i;
            seg_locate(G, k, &y, &x);
            xend = G.seg_x0[sgm] + G.seg_w[sgm];
            send = sgm + 1 < G.nseg ? G.seg_start[sgm + 1] : total;
            pl = plane_of(k);
          }
          const unsigned u = w8[q];
          pl[(unsigned long long)y * G.cols + x] = (int32_t)((u >> 1) ^ (0u - (u & 1u)));   // zigzagDecode16
          if (++x >= xend) { x = G.seg_x0[sgm]; y++; }
        }
      }
    }
  } else if (blockIdx.x == 0 && threadIdx.x == 0) {
    // escape triples 65535,hi,lo shift every later coefficient (waveletfsecompressu16.go:50-55): positions are
    // only discoverable serially.  Rare (|coefficient| > 32767); one thread walks the stream.
    unsigned i = 0, k = 0;
    while (i < n && k < total) {
      int32_t v;
      if (e[i] != 65535) { const unsigned u = e[i]; v = (int32_t)((u >> 1) ^ (0u - (u & 1u))); i++; }
      else {
        if (i + 2 >= n) break;
        v = (int32_t)(((unsigned)e[i + 1] << 16) | (unsigned)e[i + 2]);
        i += 3;
      }
      unsigned y, x;
      seg_locate(G, k, &y, &x);
      plane_of(k)[(unsigned long long)y * G.cols + x] = v;
      k++;
    }
    if (k != total) U->status = MIC_E_SIZE;
  }
}

// -------- inverse lifting, element-wise (wt53Inverse1D waveletu16.go:75-122) -------------------------
// in: de-interleaved [low | high] along the transformed axis; out: interleaved samples.
// even[m] = s[m] - ((dL + dR + 2) >> 2),  odd[m] = d[m] + ((even[m] + evenR) >> 1)
struct Lift {
  __device__ static int even(int s, int d_prev, int d_cur, bool has_prev, bool has_cur) {
    const int dR = has_cur ? d_cur : (has_prev ? d_prev : 0);
    const int dL = has_prev ? d_prev : dR;
    return s - ((dL + dR + 2) >> 2);
  }
};

// V1 layouts (waveletInverse2DRegion, waveletfsecompressu16.go:180-189): the 1-D lifting works IN PLACE on interleaved
// samples (even = low, odd = high; wt53Inverse1D waveletu16.go:75-122) and a level covers the top-left r x c corner of the
// interleaved buffer.  Out of place here, one output sample per thread:
//   x[2i]   = v[2i]   - ((dL + dR + 2) >> 2)   dR = v[2i+1] (or v[2i-1] at the end of an odd-length signal, 0 if n == 1), dL = v[2i-1] (or dR at i == 0)
//   x[2i+1] = v[2i+1] + ((x[2i] + x[2i+2]) >> 1)   (x[2i] again at the right edge)
// AXIS 0: along columns (stride = row pitch), AXIS 1: along rows.
template <int AXIS>
__global__ void __launch_bounds__(256)
k_wt53_il_inv(const int32_t* __restrict__ A, int32_t* __restrict__ B, unsigned r, unsigned c, unsigned cols, unsigned long long img_stride) {
  const unsigned x = blockIdx.x * blockDim.x + threadIdx.x;
  const unsigned y = blockIdx.y;
  if (x >= c || y >= r) return;
  const int32_t* in = A + (unsigned long long)blockIdx.z * img_stride;
  int32_t* out = B + (unsigned long long)blockIdx.z * img_stride;
  const unsigned n = AXIS == 0 ? r : c;            // length of the lifted signal
  const unsigned p = AXIS == 0 ? y : x;            // position in it
  auto V = [&](unsigned q) -> int {
    return AXIS == 0 ? in[(unsigned long long)q * cols + x] : in[(unsigned long long)y * cols + q];
  };
  auto even = [&](unsigned q) -> int {             // restored even sample at even position q
    int dR;
    if (q + 1 < n) dR = V(q + 1);
    else dR = q > 0 ? V(q - 1) : 0;
    const int dL = q > 0 ? V(q - 1) : dR;
    return V(q) - ((dL + dR + 2) >> 2);
  };
  int res;
  if (n < 2) res = V(p);
  else if ((p & 1u) == 0) res = even(p);
  else {
    const int xl = even(p - 1);
    const int xr = p + 1 < n ? even(p + 1) : xl;
    res = V(p) + ((xl + xr) >> 1);
  }
  out[(unsigned long long)y * cols + x] = res;
}

// column pass: A -> B over the top-left r x c region (full row pitch `cols`)
__global__ void __launch_bounds__(256)
k_wt53_inv_cols(const int32_t* __restrict__ A, int32_t* __restrict__ B, unsigned r, unsigned c, unsigned cols, unsigned long long img_stride) {
  const unsigned x = blockIdx.x * blockDim.x + threadIdx.x;
  const unsigned m = blockIdx.y;            // output rows 2m, 2m+1
  if (x >= c) return;
  const int32_t* in = A + (unsigned long long)blockIdx.z * img_stride;
  int32_t* out = B + (unsigned long long)blockIdx.z * img_stride;
  if (r < 2) { if (m == 0) out[x] = in[x]; return; }
  const unsigned n_low = (r + 1) / 2, n_high = r / 2;
  if (m >= n_low) return;
  auto S = [&](unsigned i) { return in[(unsigned long long)i * cols + x]; };
  auto Dh = [&](unsigned j) { return in[(unsigned long long)(n_low + j) * cols + x]; };
  const bool hp = m > 0, hc = m < n_high;
  const int dp = hp ? Dh(m - 1) : 0, dc = hc ? Dh(m) : 0;
  const int ev = Lift::even(S(m), dp, dc, hp, hc);
  out[(unsigned long long)(2 * m) * cols + x] = ev;
  if (hc) {
    int evr = ev;
    if (2 * m + 2 < r) {
      const bool hc2 = m + 1 < n_high;
      evr = Lift::even(S(m + 1), dc, hc2 ? Dh(m + 1) : 0, true, hc2);
    }
    out[(unsigned long long)(2 * m + 1) * cols + x] = dc + ((ev + evr) >> 1);
  }
}

// row pass: B -> A over the r x c region; FINAL writes uint16 pixels (pitch `cols`) instead
template <bool FINAL>
__global__ void __launch_bounds__(256)
k_wt53_inv_rows(const int32_t* __restrict__ B, int32_t* __restrict__ A, uint16_t* __restrict__ px, unsigned r, unsigned c, unsigned cols,
                unsigned long long img_stride) {
  const unsigned m = blockIdx.x * blockDim.x + threadIdx.x;   // output columns 2m, 2m+1
  const unsigned y = blockIdx.y;
  if (y >= r) return;
  const int32_t* in = B + (unsigned long long)blockIdx.z * img_stride + (unsigned long long)y * cols;
  const unsigned long long obase = (unsigned long long)blockIdx.z * img_stride + (unsigned long long)y * cols;
  if (c < 2) {
    if (m == 0) { if (FINAL) px[obase] = (uint16_t)in[0]; else A[obase] = in[0]; }
    return;
  }
  const unsigned n_low = (c + 1) / 2, n_high = c / 2;
  if (m >= n_low) return;
  const bool hp = m > 0, hc = m < n_high;
  const int dp = hp ? in[n_low + m - 1] : 0, dc = hc ? in[n_low + m] : 0;
  const int ev = Lift::even(in[m], dp, dc, hp, hc);
  int od = 0;
  if (hc) {
    int evr = ev;
    if (2 * m + 2 < c) {
      const bool hc2 = m + 1 < n_high;
      evr = Lift::even(in[m + 1], dc, hc2 ? in[n_low + m + 1] : 0, true, hc2);
    }
    od = dc + ((ev + evr) >> 1);
  }
  if (FINAL) {
    px[obase + 2 * m] = (uint16_t)ev;
    if (hc) px[obase + 2 * m + 1] = (uint16_t)od;
  } else {
    A[obase + 2 * m] = ev;
    if (hc) A[obase + 2 * m + 1] = od;
  }
}

__global__ void __launch_bounds__(256)
k_i32_to_u16(const int32_t* __restrict__ A, uint16_t* __restrict__ px, unsigned long long n) {
  for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (unsigned long long)gridDim.x * blockDim.x)
    px[i] = (uint16_t)A[i];
}

// -------- fused inverse level: TMA-staged tile, column + row lifting in shared memory, 16 B stores ---------------------
// One CTA produces a 64 x 64 output tile of one level.  The four Mallat quadrants it needs (LL/HL/LH/HH boxes of 34
// rows x 36 columns: 32 + halo, the width rounded up to a 16 B multiple) arrive as four 3-D TMA box loads
// (cp.async.bulk.tensor, plane = [image][row][col] int32; coordinates left of / above the plane are zero-filled) on one
// mbarrier.  Phase 1 lifts the columns of every box column into V (shared), phase 2 lifts the rows out of V and stores
// eight results per thread: two 16 B int32 stores, or one 16 B store of uint16 pixels on the finest level.  The
// element formulas are the ones of k_wt53_inv_cols / k_wt53_inv_rows above (wt53Inverse1D, waveletu16.go:75-122) with
// the accessors pointed at the boxes, so odd sizes and the mirrored borders behave identically.  A level reads plane X
// and writes plane Y != X (its output is the LL quadrant of the next finer level, whose detail subbands the scatter
// put into Y), so no CTA overwrites what another still has to read.
constexpr int WT_T = 32;                 // low-pass samples per tile side (output tile 64 x 64)
constexpr int WT_BW = WT_T + 8;          // box width (columns): 32 + halo 2 + up to 3 columns of slack, because the innermost TMA
                                         // coordinate has to be a multiple of 16 B (measured: an odd x faults with "illegal instruction",
                                         // tools/probe/tma_probe.cu); the high-pass boxes start at (nlc + n0 - 1) & ~3
constexpr int WT_VW = WT_T + 4;          // columns of the lifted rows kept per output row
constexpr int WT_BH = WT_T + 2;          // box height (rows)

struct alignas(64) TmaDesc { unsigned char bytes[128]; };   // CUtensorMap (driver API type, kept opaque here)

template <bool FINAL>
__global__ void __launch_bounds__(256)
k_wt53_inv_level_tma(const __grid_constant__ TmaDesc tmap, int32_t* __restrict__ outp, uint16_t* __restrict__ px, unsigned r, unsigned c,
                     unsigned cols, unsigned long long img_stride) {
  struct alignas(128) Box { int32_t v[WT_BH][WT_BW]; };        // every TMA destination starts on a 128 B boundary
  __shared__ Box s_boxes[4];                                      // LL, HL, LH, HH
  __shared__ __align__(16) int32_t s_vlow[2 * WT_T][WT_VW];      // column-lifted rows: low-pass columns n0 + j
  __shared__ __align__(16) int32_t s_vhigh[2 * WT_T][WT_VW];     // high-pass columns n0 - 1 + j
  __shared__ __align__(8) unsigned long long s_bar;
  const unsigned tid = threadIdx.x;
  const unsigned nlr = (r + 1) / 2, nhr = r / 2, nlc = (c + 1) / 2, nhc = c / 2;
  const int m0 = (int)blockIdx.y * WT_T, n0 = (int)blockIdx.x * WT_T;
  const unsigned img = blockIdx.z;
  const unsigned bar = (unsigned)__cvta_generic_to_shared(&s_bar);
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // the TMA unit (async proxy) must see the initialised barrier
  }
  __syncthreads();
  if (tid == 0) {
    const unsigned bytes = 4u * WT_BH * WT_BW * 4u;
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
    const int cx[4] = {n0, ((int)nlc + n0 - 1) & ~3, n0, ((int)nlc + n0 - 1) & ~3};
    const int cy[4] = {m0, m0, (int)nlr + m0 - 1, (int)nlr + m0 - 1};
#pragma unroll
    for (int q = 0; q < 4; q++) {
      const unsigned dst = (unsigned)__cvta_generic_to_shared(&s_boxes[q].v[0][0]);
      asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
                   ::"r"(dst), "l"(&tmap), "r"(cx[q]), "r"(cy[q]), "r"((int)img), "r"(bar)
                   : "memory");
    }
  }
  {
    unsigned done = 0;
    while (!done)
      asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0, 0x989680;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                   : "=r"(done) : "r"(bar) : "memory");
  }
  // ---- phase 1: columns.  item = (low-pass row i of the tile, box column j, which box pair) ---------------------
  // low-pass columns use LL (s) and LH (d); high-pass columns use HL (s) and HH (d)
  const unsigned dxh = (unsigned)((int)nlc + n0 - 1) & 3u;     // the high-pass boxes start dxh columns early (16 B aligned x)
  for (unsigned it = tid; it < (unsigned)(WT_T * WT_VW * 2); it += blockDim.x) {
    const unsigned half = it / (WT_T * WT_VW), rem = it - half * (WT_T * WT_VW);
    const unsigned i = rem / WT_VW, j = rem - i * WT_VW;
    const unsigned m = (unsigned)m0 + i;               // low-pass row index of the level
    if (m >= nlr) continue;
    const unsigned jb = half ? j + dxh : j;            // column inside the box
    const int32_t (*Sb)[WT_BW] = s_boxes[half].v;          // LL or HL
    const int32_t (*Db)[WT_BW] = s_boxes[2 + half].v;      // LH or HH; box row i holds high-pass row m - 1
    const bool hp = m > 0, hc = m < nhr;
    const int dp = hp ? Db[i][jb] : 0, dc = hc ? Db[i + 1][jb] : 0;
    const int ev = Lift::even(Sb[i][jb], dp, dc, hp, hc);
    int od = 0;
    if (hc) {
      int evr = ev;
      if (2 * m + 2 < r) {
        const bool hc2 = m + 1 < nhr;
        evr = Lift::even(Sb[i + 1][jb], dc, hc2 ? Db[i + 2][jb] : 0, true, hc2);
      }
      od = dc + ((ev + evr) >> 1);
    }
    int32_t (*V)[WT_VW] = half ? s_vhigh : s_vlow;
    V[2 * i][j] = ev;
    V[2 * i + 1][j] = od;
  }
  __syncthreads();
  // ---- phase 2: rows.  item = (output row yy of the tile, group of four low-pass columns) -> eight output columns ----
  int32_t* oimg = outp + (unsigned long long)img * img_stride;
  uint16_t* pimg = px + (unsigned long long)img * img_stride;
  for (unsigned it = tid; it < (unsigned)(2 * WT_T * (WT_T / 4)); it += blockDim.x) {
    const unsigned yy = it / (WT_T / 4), g = it - yy * (WT_T / 4);
    const unsigned Y = 2u * (unsigned)m0 + yy;
    if (Y >= r) continue;
    int o[8];
    bool any = false;
#pragma unroll
    for (int q = 0; q < 4; q++) {
      const unsigned nn = 4 * g + q;                   // low-pass column of the tile
      const unsigned n = (unsigned)n0 + nn;
      o[2 * q] = 0; o[2 * q + 1] = 0;
      if (n >= nlc) continue;
      any = true;
      const bool hp = n > 0, hc = n < nhc;
      const int dp = hp ? s_vhigh[yy][nn] : 0, dc = hc ? s_vhigh[yy][nn + 1] : 0;
      const int ev = Lift::even(s_vlow[yy][nn], dp, dc, hp, hc);
      int od = 0;
      if (hc) {
        int evr = ev;
        if (2 * n + 2 < c) {
          const bool hc2 = n + 1 < nhc;
          evr = Lift::even(s_vlow[yy][nn + 1], dc, hc2 ? s_vhigh[yy][nn + 2] : 0, true, hc2);
        }
        od = dc + ((ev + evr) >> 1);
      }
      o[2 * q] = ev; o[2 * q + 1] = od;
    }
    if (!any) continue;
    const unsigned X = 2u * ((unsigned)n0 + 4 * g);     // first output column of the group (a multiple of 8)
    const unsigned long long at = (unsigned long long)Y * cols + X;
    if (X + 8 <= c) {
      if (FINAL) {
        *reinterpret_cast<uint4*>(pimg + at) = make_uint4((unsigned)(o[0] & 0xFFFF) | ((unsigned)o[1] << 16), (unsigned)(o[2] & 0xFFFF) | ((unsigned)o[3] << 16),
                                                          (unsigned)(o[4] & 0xFFFF) | ((unsigned)o[5] << 16), (unsigned)(o[6] & 0xFFFF) | ((unsigned)o[7] << 16));
      } else {
        *reinterpret_cast<int4*>(oimg + at) = make_int4(o[0], o[1], o[2], o[3]);
        *reinterpret_cast<int4*>(oimg + at + 4) = make_int4(o[4], o[5], o[6], o[7]);
      }
    } else {
#pragma unroll
      for (int q = 0; q < 8; q++)
        if (X + q < c) { if (FINAL) pimg[at + q] = (uint16_t)o[q]; else oimg[at + q] = o[q]; }
    }
  }
}

// cuTensorMapEncodeTiled through the runtime's driver entry point (no -lcuda): 3-D int32 tensor [nimg][rows][cols]
static bool make_plane_tmap(TmaDesc* out, const int32_t* plane, unsigned rows, unsigned cols, unsigned nimg) {
  static_assert(sizeof(CUtensorMap) == sizeof(TmaDesc), "CUtensorMap is 128 bytes");
  typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                               const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
  static EncodeFn fn = [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess) p = nullptr;
    return (EncodeFn)p;
  }();
  if (!fn) return false;
  const cuuint64_t dims[3] = {cols, rows, nimg};
  const cuuint64_t strides[2] = {(cuuint64_t)cols * 4, (cuuint64_t)rows * cols * 4};   // bytes, dims 1 and 2
  const cuuint32_t box[3] = {WT_BW, WT_BH, 1};
  const cuuint32_t estr[3] = {1, 1, 1};
  return fn(reinterpret_cast<CUtensorMap*>(out), CU_TENSOR_MAP_DATA_TYPE_INT32, 3, const_cast<int32_t*>(plane), dims, strides, box, estr,
            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

static bool wavelet_tma_enabled() {
  static const bool v = [] { const char* e = getenv("MICGPU_WAVELET_TMA"); return !(e && e[0] == '0'); }();
  return v;
}

// copy of the untouched part is not needed: both passes only rewrite the r x c corner and level l+1's
// corner is inside level l's.
void launch_wavelet_decode(MicUnit* d_units, const int* d_unit_of_img, int nimg, const uint16_t* d_stream, int* d_flags,
                           int32_t* d_A, int32_t* d_B, uint16_t* d_px, const WaveletGeom& G, cudaStream_t st, int v1_layout) {
  if (nimg <= 0) return;
  const unsigned total = G.rows * G.cols;
  if (v1_layout) {
    // coefficients in raster order (G holds ONE segment covering the image), interleaved in-place levels
    cudaMemsetAsync(d_flags, 0, nimg * sizeof(int), st);
    const unsigned eb1 = (total + 256 * 16 - 1) / (256 * 16);
    k_wavelet_has_escape<<<dim3(eb1 < 32 ? 32 : (eb1 > 1024 ? 1024 : eb1), nimg), 256, 0, st>>>(d_units, d_unit_of_img, d_stream, d_flags);
    const unsigned sb1 = (total / 8 + 255) / 256 + 1;
    k_wavelet_scatter<false><<<dim3(sb1 < 1024 ? sb1 : 1024, nimg), 256, 0, st>>>(d_units, d_unit_of_img, d_stream, d_flags, d_A, d_B, G);
    unsigned vr[256], vc[256];
    unsigned r1 = G.rows, c1 = G.cols;
    for (int l = 0; l < G.levels; l++) { vr[l] = r1; vc[l] = c1; r1 = (r1 + 1) / 2; c1 = (c1 + 1) / 2; }
    for (int l = G.levels - 1; l >= 0; l--) {
      const dim3 grid((vc[l] + 255) / 256, vr[l], nimg);
      k_wt53_il_inv<0><<<grid, 256, 0, st>>>(d_A, d_B, vr[l], vc[l], G.cols, total);
      k_wt53_il_inv<1><<<grid, 256, 0, st>>>(d_B, d_A, vr[l], vc[l], G.cols, total);
    }
    k_i32_to_u16<<<1024, 256, 0, st>>>(d_A, d_px, (unsigned long long)total * nimg);
    return;
  }
  cudaMemsetAsync(d_flags, 0, nimg * sizeof(int), st);
  const unsigned eb = (total + 256 * 16 - 1) / (256 * 16);   // ~16 words per thread
  k_wavelet_has_escape<<<dim3(eb < 32 ? 32 : (eb > 1024 ? 1024 : eb), nimg), 256, 0, st>>>(d_units, d_unit_of_img, d_stream, d_flags);
  const unsigned sb = (total / 8 + 255) / 256 + 1;   // eight words per thread
  unsigned dr[10], dc[10];
  unsigned r = G.rows, c = G.cols;
  for (int l = 0; l < G.levels; l++) { dr[l] = r; dc[l] = c; r = (r + 1) / 2; c = (c + 1) / 2; }
  // fused TMA path: rows of 16 B multiples for int32 planes and uint16 pixels, every level at least 2 x 2
  bool tma = wavelet_tma_enabled() && G.levels >= 1 && (G.cols % 8) == 0 && (reinterpret_cast<uintptr_t>(d_A) % 16) == 0 &&
             (reinterpret_cast<uintptr_t>(d_B) % 16) == 0 && (reinterpret_cast<uintptr_t>(d_px) % 16) == 0;
  for (int l = 0; l < G.levels; l++) tma = tma && dr[l] >= 2 && dc[l] >= 2;
  TmaDesc mapA, mapB;
  if (tma) tma = make_plane_tmap(&mapA, d_A, G.rows, G.cols, (unsigned)nimg) && make_plane_tmap(&mapB, d_B, G.rows, G.cols, (unsigned)nimg);
  if (tma) {
    k_wavelet_scatter<true><<<dim3(sb < 1024 ? sb : 1024, nimg), 256, 0, st>>>(d_units, d_unit_of_img, d_stream, d_flags, d_A, d_B, G);
    for (int l = G.levels - 1; l >= 0; l--) {
      const unsigned rr = dr[l], cc = dc[l];
      const dim3 grid(((cc + 1) / 2 + WT_T - 1) / WT_T, ((rr + 1) / 2 + WT_T - 1) / WT_T, nimg);
      const TmaDesc& in = (l & 1) ? mapB : mapA;          // level l lives in plane (l & 1)
      int32_t* outp = (l & 1) ? d_A : d_B;                // and writes the other one
      if (l == 0) k_wt53_inv_level_tma<true><<<grid, 256, 0, st>>>(in, outp, d_px, rr, cc, G.cols, total);
      else k_wt53_inv_level_tma<false><<<grid, 256, 0, st>>>(in, outp, d_px, rr, cc, G.cols, total);
    }
    return;
  }
  k_wavelet_scatter<false><<<dim3(sb < 1024 ? sb : 1024, nimg), 256, 0, st>>>(d_units, d_unit_of_img, d_stream, d_flags, d_A, d_B, G);
  if (G.levels == 0) {
    // no transform applied (rows < 2 or cols < 2): coefficients are the pixels
    k_i32_to_u16<<<1024, 256, 0, st>>>(d_A, d_px, (unsigned long long)total * nimg);
    return;
  }
  for (int l = G.levels - 1; l >= 0; l--) {
    const unsigned rr = dr[l], cc = dc[l];
    k_wt53_inv_cols<<<dim3((cc + 255) / 256, (rr + 1) / 2, nimg), 256, 0, st>>>(d_A, d_B, rr, cc, G.cols, total);
    const dim3 g2(((cc + 1) / 2 + 255) / 256, rr, nimg);
    if (l == 0) k_wt53_inv_rows<true><<<g2, 256, 0, st>>>(d_B, d_A, d_px, rr, cc, G.cols, total);
    else k_wt53_inv_rows<false><<<g2, 256, 0, st>>>(d_B, d_A, d_px, rr, cc, G.cols, total);
  }
}

}  // namespace micgpu
