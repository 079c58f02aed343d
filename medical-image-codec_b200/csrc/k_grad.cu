// k_grad.cu -- the gradient-adaptive predictor: K4g decode wavefront and the PICA row-cost kernel.
//
// Replaces GradDeltaRleDecompressU16.Decompress (deltagradrlecompressu16.go:70-133) / gradPredict
// (deltagradcompressu16.go:149-166) behind DecompressSingleFrameGrad (multiframecompress.go:129-142) and the strips of a
// PICA container that carry picaFlagGradPredictor (parallelstripsadaptive.go:186-190), and the row statistics of
// adaptiveStripBoundaries (:229-241) for the encoder.
//
// The predictor reads W, N, NW and NE:  avg = (W + N) >> 1, corrected by (NE - NW) >> 3 clamped to +-((|W-NW| + |N-NW|) >> 1).
// The clamp makes it non-linear in W, so the block-function scan of k_delta_scan.cu does not apply; the dependences
// (x-1, y) and (x+1, y-1) are the classic JPEG-LS/CALIC skew: row y can trail row y-1 by two pixels.
//
// Mapping: one CTA per unit, one WARP per band of 32 rows, lane r = row 32 b + r, and at step t lane r works on pixel
// x = t - 2 r.  What the lane above produced in the previous step is exactly this lane's NE, so the three neighbours of
// the row above live in a three-register window fed by ONE shuffle per step; W is the lane's own previous result.  Bands
// chain through global memory: lane 0 reads the last row of the band above (32 pixels per block of 32 steps, one
// coalesced volatile load, handed out by a uniform shuffle) once the producing warp has published that far in a
// shared-memory progress counter.  Warp w takes bands w, w + NW, ...; every band depends only on the one before it, so
// the spin-waits cannot deadlock (all warps of a CTA are resident).
//
// Residuals and literal bits come from the same D / M planes K3 writes for the avg predictor (padded column a_y + x,
// mic_unit.h); a lane loads the 32 residuals of a block into registers before the block starts (the loop is fully
// unrolled, so the registers are statically indexed).
#include "mic_device.cuh"

namespace micgpu {

namespace {

constexpr int GRAD_WARPS = 8;

__device__ __forceinline__ int grad_predict(int w, int n, int nw, int ne) {
  const int avg = (w + n) >> 1;
  const int g = abs(w - nw) + abs(n - nw);
  const int limit = g >> 1;
  int corr = (ne - nw) >> 3;                 // gradShift = 3 (deltagradcompressu16.go:147); arithmetic shift like Go's int32 >>
  corr = min(max(corr, -limit), limit);      // g == 0 -> limit == 0 -> corr == 0: the reference's early return
  return avg + corr;
}

__device__ __forceinline__ int ld_volatile_smem(const int* p) {
  int v;
  asm volatile("ld.volatile.shared.s32 %0, [%1];" : "=r"(v) : "r"((unsigned)__cvta_generic_to_shared(p)) : "memory");
  return v;
}
__device__ __forceinline__ void st_volatile_smem(int* p, int v) {
  asm volatile("st.volatile.shared.s32 [%0], %1;" ::"r"((unsigned)__cvta_generic_to_shared(p)), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned ld_volatile_u16(const uint16_t* p) {
  unsigned short v;
  asm volatile("ld.volatile.global.u16 %0, [%1];" : "=h"(v) : "l"(p) : "memory");
  return v;
}

}  // namespace

__global__ void __launch_bounds__(32 * GRAD_WARPS)
k_grad_wavefront(MicUnit* __restrict__ units, const int* __restrict__ list, int nlist, const uint16_t* __restrict__ D,
                 const uint32_t* __restrict__ M, uint16_t* __restrict__ out) {
  extern __shared__ int s_progress[];   // [bands]: pixels of the band's last row that are stored and visible
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const MicUnit* U = &units[list[blockIdx.x]];
  if (U->status != MIC_OK || U->predictor != 1u) return;
  const int W = (int)U->width, H = (int)U->height;
  const unsigned wp = U->wp, mwords = wp >> 5;
  const int thr = (int)U->thr;
  const int align0 = (int)U->align0 & 7, dW = W & 7;
  const uint16_t* Du = D + U->d_off;
  const uint32_t* Mu = M + U->m_off;
  uint16_t* Ou = out + U->out_off;
  const int bands = (H + 31) >> 5;
  for (int i = threadIdx.x; i < bands; i += blockDim.x) s_progress[i] = 0;
  __syncthreads();

  const int nblocks = (W + 62 + 31) >> 5;   // steps 0 .. W + 61, in blocks of 32
  for (int b = warp; b < bands; b += GRAD_WARPS) {
    const int y = 32 * b + lane;
    const bool rowvalid = y < H;
    const int yc = rowvalid ? y : H - 1;
    const int rlast = min(31, H - 1 - 32 * b);                 // lane of the band's last row
    const int a = (align0 + (yc & 7) * dW) & 7;                 // phase of this row in D / M
    const uint16_t* Drow = Du + (size_t)yc * wp + a;
    const uint32_t* Mrow = Mu + (size_t)yc * mwords;
    uint16_t* Orow = Ou + (size_t)yc * W;
    const uint16_t* above = Ou + (size_t)(32 * b - 1) * W;      // last row of the band above (b > 0 only)
    int w = 0, n = 0, nw = 0, ne = 0, last = 0;
    if (b > 0) {
      while (ld_volatile_smem(&s_progress[b - 1]) < 1) { }
      ne = (int)ld_volatile_u16(above);                          // becomes N of pixel 0 at step 0 (lane 0)
    }
    for (int k = 0; k < nblocks; k++) {
      const int t0 = 32 * k;
      const int x0 = t0 - 2 * lane;                              // this lane's pixel at the first step of the block
      // ---- the row above, for lane 0: pixels t0 + 1 .. t0 + 32 -------------------------------------------------------
      int above_reg = 0;
      if (b > 0) {
        const int need = min(t0 + 33, W);
        while (ld_volatile_smem(&s_progress[b - 1]) < need) { }
        const int xa = t0 + 1 + lane;
        if (xa < W) above_reg = (int)ld_volatile_u16(above + xa);
      }
      // ---- residuals and literal bits of the block -----------------------------------------------------------------
      unsigned res[32];
#pragma unroll
      for (int i = 0; i < 32; i++) {
        const int x = x0 + i;
        res[i] = (rowvalid && x >= 0 && x < W) ? (unsigned)__ldg(Drow + x) : 0u;
      }
      unsigned lits;
      {
        const int p0 = a + x0;                                   // padded column of step 0 (negative while the lane waits)
        const int wi = p0 >> 5;
        const unsigned m0 = (wi >= 0 && wi < (int)mwords) ? __ldg(Mrow + wi) : 0u;
        const unsigned m1 = (wi + 1 >= 0 && wi + 1 < (int)mwords) ? __ldg(Mrow + wi + 1) : 0u;
        lits = __funnelshift_r(m0, m1, (unsigned)p0 & 31u);
      }
      // ---- 32 steps ------------------------------------------------------------------------------------------------
#pragma unroll
      for (int i = 0; i < 32; i++) {
        const int x = x0 + i;
        int inc = __shfl_up_sync(0xffffffffu, last, 1);
        const int ab = __shfl_sync(0xffffffffu, above_reg, i);
        if (lane == 0) inc = ab;
        nw = n; n = ne; ne = inc;
        const bool active = rowvalid && x >= 0 && x < W;
        int pred;
        if (y == 0) pred = x == 0 ? 0 : w;                       // first row: left only (deltagradrlecompressu16.go:79-88)
        else if (x == 0) pred = n;                               // first column: top only (:94-101)
        else pred = grad_predict(w, n, nw, x + 1 < W ? ne : nw);
        const int d = (int)res[i];
        const int val = ((lits >> i) & 1u) ? d : ((pred + d - thr) & 0xFFFF);
        if (active) {
          w = val; last = val;
          Orow[x] = (uint16_t)val;
        }
      }
      // ---- publish the band's last row ------------------------------------------------------------------------------
      __threadfence_block();
      if (lane == rlast) st_volatile_smem(&s_progress[b], max(0, min(W, t0 + 32 - 2 * rlast)));
    }
  }
}

void launch_grad_wavefront(MicUnit* d_units, const int* d_list, int nlist, const uint16_t* d_D, const uint32_t* d_M, uint16_t* d_out,
                           int max_height, cudaStream_t st) {
  if (nlist <= 0) return;
  const size_t smem = (size_t)((max_height + 31) / 32 + 1) * sizeof(int);
  k_grad_wavefront<<<nlist, 32 * GRAD_WARPS, smem, st>>>(d_units, d_list, nlist, d_D, d_M, d_out);
}

// adaptiveStripBoundaries (parallelstripsadaptive.go:229-241): cost[y] = sum over x of |px[y][x] - px[y-1][x]| (cost[0] = 0).
// One warp per row, 16 B loads where the rows allow them.
__global__ void __launch_bounds__(256)
k_row_costs(const uint16_t* __restrict__ px, int width, int height, unsigned long long* __restrict__ cost) {
  const int lane = threadIdx.x & 31;
  const int y = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (y >= height) return;
  unsigned long long sum = 0;
  if (y >= 1) {
    const uint16_t* r1 = px + (size_t)y * width;
    const uint16_t* r0 = r1 - width;
    for (int x = lane; x < width; x += 32) {
      const int d = (int)__ldg(r1 + x) - (int)__ldg(r0 + x);
      sum += (unsigned)(d < 0 ? -d : d);
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) sum += __shfl_down_sync(0xffffffffu, sum, o);
  if (lane == 0) cost[y] = sum;
}

void launch_row_costs(const uint16_t* d_px, int width, int height, unsigned long long* d_cost, cudaStream_t st) {
  if (height <= 0) return;
  k_row_costs<<<(height + 7) / 8, 256, 0, st>>>(d_px, width, height, d_cost);
}

}  // namespace micgpu
