// k_wavelet.cu -- K9 (decode half): WaveletV2 coefficient unpack, subband scatter and inverse 5/3 lifting.
//
// Replaces u16ToWaveletCoeffs + zigzagDecode16 (waveletfsecompressu16.go:45-58,543-546),
// scatterSubbandOrder (:243-282), wt53Inverse2DSeparated[SIMD] (waveletu16.go:212-257,401-508) and the
// AVX2 wt53Inv{Predict,Update} kernels.  Lifting is written element-wise: every output pair of a row
// (or column) is computed by one thread from five neighbouring inputs (the two update results it
// needs are recomputed instead of communicated), so both passes are plain coalesced streaming
// kernels with no serial walk; levels ping-pong between two int32 planes.
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

__global__ void __launch_bounds__(256)
k_wavelet_scatter(MicUnit* __restrict__ units, const int* __restrict__ unit_of_img, const uint16_t* __restrict__ stream,
                  const int* __restrict__ flags, int32_t* __restrict__ planeA, WaveletGeom G) {
  const int img = blockIdx.y;
  MicUnit* U = &units[unit_of_img[img]];
  if (U->status != MIC_OK) return;
  const uint16_t* e = stream + U->out_off;
  const unsigned n = U->thr;
  const unsigned total = G.rows * G.cols;
  int32_t* A = planeA + (unsigned long long)img * total;
  if (!flags[img]) {
    // no escape triples: coefficient k is word k (u16ToWaveletCoeffs fast case)
    if (n < total) {
      if (blockIdx.x == 0 && threadIdx.x == 0) U->status = MIC_E_SIZE;
      return;
    }
    for (unsigned k = blockIdx.x * blockDim.x + threadIdx.x; k < total; k += gridDim.x * blockDim.x) {
      const unsigned u = e[k];
      unsigned y, x;
      seg_locate(G, k, &y, &x);
      A[(unsigned long long)y * G.cols + x] = (int32_t)((u >> 1) ^ (0u - (u & 1u)));   // zigzagDecode16
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
      A[(unsigned long long)y * G.cols + x] = v;
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

// copy of the untouched part is not needed: both passes only rewrite the r x c corner and level l+1's
// corner is inside level l's.
void launch_wavelet_decode(MicUnit* d_units, const int* d_unit_of_img, int nimg, const uint16_t* d_stream, int* d_flags,
                           int32_t* d_A, int32_t* d_B, uint16_t* d_px, const WaveletGeom& G, cudaStream_t st) {
  if (nimg <= 0) return;
  const unsigned total = G.rows * G.cols;
  cudaMemsetAsync(d_flags, 0, nimg * sizeof(int), st);
  k_wavelet_has_escape<<<dim3(32, nimg), 256, 0, st>>>(d_units, d_unit_of_img, d_stream, d_flags);
  const unsigned sb = (total + 255) / 256;
  k_wavelet_scatter<<<dim3(sb < 1024 ? sb : 1024, nimg), 256, 0, st>>>(d_units, d_unit_of_img, d_stream, d_flags, d_A, G);
  unsigned dr[10], dc[10];
  unsigned r = G.rows, c = G.cols;
  for (int l = 0; l < G.levels; l++) { dr[l] = r; dc[l] = c; r = (r + 1) / 2; c = (c + 1) / 2; }
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
