// k_delta.cu -- K4: inverse avg(top,left) predictor as an anti-diagonal wavefront.
//
// Replaces DeltaRleDecompressU16.DecodeNextSymbol[NC] (deltarlecompressu16.go:102-128)
// and the C twin's delta_decode_simd.  out = ((left+top)>>1) + diff is not
// linear in `left`, so there is no closed-form row scan; instead thread y owns
// row y and at step t reconstructs pixel (y, t-y): `left` is the thread's own
// previous result, `top` arrives from lane-1 by shuffle (it was produced one
// step earlier).  Warps chain through one shared-memory boundary row each,
// gated by a progress counter every K4_GATE columns, so a 256-row strip takes
// ~W+H steps.  Units taller than the CTA are processed in passes of blockDim
// rows; the last row of a pass seeds the next one.
//
// Inputs: residual plane D (row pitch wp, 16 B aligned rows; a literal pixel
// holds its raw value) and literal bit mask M (wp/32 words per row), both
// produced by K3.  Output rows are packed (pitch W): pixels are gathered eight
// at a time and written with 16 B stores once the row address is aligned.
#include "mic_device.cuh"

namespace micgpu {

constexpr int K4_GATE = 8;

__device__ __forceinline__ int ld_volatile_s32(const int* p) {
  return *reinterpret_cast<const volatile int*>(p);
}

__global__ void __launch_bounds__(1024)
k_delta_wavefront(MicUnit* __restrict__ units, const int* __restrict__ list, int nlist,
                  const uint16_t* __restrict__ D, const uint32_t* __restrict__ M, uint16_t* __restrict__ out,
                  int brow_pitch) {
  extern __shared__ __align__(16) uint16_t s_brow[];  // (nwarps+1) boundary rows of brow_pitch elements
  __shared__ int s_prog[33];                          // s_prog[w+1]: columns finished by warp w's last lane

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int nwarps = blockDim.x >> 5;
  const MicUnit* U = &units[list[blockIdx.x]];
  if (U->status != MIC_OK) return;
  const int W = (int)U->width, H = (int)U->height;
  const unsigned wp = U->wp;
  const int thr = (int)U->thr;
  const uint16_t* Du = D + U->d_off;
  const uint32_t* Mu = M + U->m_off;
  uint16_t* Ou = out + U->out_off;

  const int passes = (H + (int)blockDim.x - 1) / (int)blockDim.x;
  for (int p = 0; p < passes; p++) {
    const int y = p * (int)blockDim.x + tid;
    const int wy0 = p * (int)blockDim.x + warp * 32;   // first row of this warp
    __syncthreads();
    if (p > 0) {
      for (int i = tid; i < W; i += blockDim.x) s_brow[i] = s_brow[(size_t)nwarps * brow_pitch + i];
    }
    if (tid <= nwarps) s_prog[tid] = tid == 0 ? W : 0;
    __syncthreads();
    if (wy0 >= H) continue;   // whole warp idle this pass (still reaches the barriers above)

    const bool row_active = y < H;
    const bool has_top = wy0 > 0;                       // lane 0 has a row above it
    const uint16_t* brow_in = s_brow + (size_t)warp * brow_pitch;
    uint16_t* brow_out = s_brow + (size_t)(warp + 1) * brow_pitch;
    const int* prog_in = &s_prog[warp];
    int* prog_out = &s_prog[warp + 1];

    const uint16_t* Drow = Du + (size_t)(row_active ? y : 0) * wp;
    const uint32_t* Mrow = Mu + (size_t)(row_active ? y : 0) * (wp >> 5);
    uint16_t* Orow = Ou + (size_t)(row_active ? y : 0) * W;
    const int ea = (int)((reinterpret_cast<uintptr_t>(Orow) >> 1) & 7);
    const int hx = (8 - ea) & 7;                                   // first 16 B aligned column
    const int tail = hx <= W ? hx + ((W - hx) >> 3) * 8 : 0;       // columns >= tail are written scalar
    const int headx = hx <= W ? hx : W;

    unsigned long long dlo = 0, dhi = 0, n1lo = 0, n1hi = 0, n2lo = 0, n2hi = 0;
    unsigned mb = 0, m1 = 0, m2 = 0;
    unsigned long long olo = 0, ohi = 0;
    auto load_group = [&](int gx, unsigned long long& lo, unsigned long long& hi, unsigned& mbits) {
      if (row_active && gx < (int)wp) {
        const uint4 v = __ldg(reinterpret_cast<const uint4*>(Drow + gx));
        lo = (unsigned long long)v.x | ((unsigned long long)v.y << 32);
        hi = (unsigned long long)v.z | ((unsigned long long)v.w << 32);
        mbits = (__ldg(Mrow + (gx >> 5)) >> (gx & 24)) & 0xFFu;
      } else {
        lo = hi = 0; mbits = 0;
      }
    };
    load_group(0, n1lo, n1hi, m1);
    load_group(8, n2lo, n2hi, m2);

    unsigned left = 0, prev_out = 0;
    int x = -lane;
    const int steps = W + 31;
    for (int t = 0; t < steps; t++, x++) {
      if (has_top && (t & (K4_GATE - 1)) == 0 && t < W) {
        const int need = min(t + K4_GATE, W);
        while (ld_volatile_s32(prog_in) < need) { }
        __threadfence_block();   // order the boundary-row reads after the progress read
      }
      unsigned top = __shfl_up_sync(0xffffffffu, prev_out, 1);
      if (lane == 0) top = (has_top && t < W) ? brow_in[t] : 0u;
      const bool active = row_active && x >= 0 && x < W;
      if (x >= 0 && (x & 7) == 0) {
        dlo = n1lo; dhi = n1hi; mb = m1;
        n1lo = n2lo; n1hi = n2hi; m1 = m2;
        load_group(x + 16, n2lo, n2hi, m2);
      }
      const unsigned d = (unsigned)(dlo & 0xFFFFu);
      const unsigned lit = mb & 1u;
      if (x >= 0) {
        dlo = (dlo >> 16) | (dhi << 48);
        dhi >>= 16;
        mb >>= 1;
      }
      // predictor (deltarlecompressu16.go:112-126): (0,0)->0, row 0->left, col 0->top, else (left+top)>>1
      unsigned pred;
      const bool up = y > 0;
      if (x > 0 && up) pred = (left + top) >> 1;
      else if (x > 0) pred = left;
      else pred = up ? top : 0u;
      unsigned val = lit ? d : ((pred + d - (unsigned)thr) & 0xFFFFu);
      if (!active) val = 0;
      left = val;
      prev_out = val;
      if (active) {
        if (lane == 31) brow_out[x] = (uint16_t)val;
        if (x < headx || x >= tail) {
          Orow[x] = (uint16_t)val;
        } else {
          olo = (olo >> 16) | (ohi << 48);
          ohi = (ohi >> 16) | ((unsigned long long)val << 48);
          if (((x - hx) & 7) == 7) {
            uint4 v;
            v.x = (unsigned)olo; v.y = (unsigned)(olo >> 32); v.z = (unsigned)ohi; v.w = (unsigned)(ohi >> 32);
            *reinterpret_cast<uint4*>(Orow + x - 7) = v;
          }
        }
        if (lane == 31 && ((((x + 1) & (K4_GATE - 1)) == 0) || x == W - 1)) {
          __threadfence_block();
          *reinterpret_cast<volatile int*>(prog_out) = x + 1;
        }
      }
    }
  }
}

int delta_wavefront_threads(int max_width, int max_height) {
  int threads = (max_height + 31) / 32 * 32;
  if (threads > 1024) threads = 1024;
  if (threads < 32) threads = 32;
  const int pitch = (max_width + 7) / 8 * 8;
  // keep (nwarps+1) boundary rows within ~200 KB of shared memory
  while (threads > 32 && (size_t)(threads / 32 + 1) * pitch * sizeof(uint16_t) > 200u * 1024u) threads -= 32;
  return threads;
}

void launch_delta_wavefront(MicUnit* d_units, const int* d_list, int nlist, const uint16_t* d_D, const uint32_t* d_M,
                            uint16_t* d_out, int max_width, int max_height, cudaStream_t st) {
  if (nlist <= 0) return;
  const int threads = delta_wavefront_threads(max_width, max_height);
  const int nwarps = threads / 32;
  const int pitch = (max_width + 7) / 8 * 8;
  const size_t smem = (size_t)(nwarps + 1) * pitch * sizeof(uint16_t);
  cudaFuncSetAttribute(k_delta_wavefront, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  k_delta_wavefront<<<nlist, threads, smem, st>>>(d_units, d_list, nlist, d_D, d_M, d_out, pitch);
}

}  // namespace micgpu
