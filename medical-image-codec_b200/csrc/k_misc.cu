// k_misc.cu -- small elementwise kernels of the MIC path.
#include "mic_device.cuh"

namespace micgpu {

// K10: TemporalDeltaDecode chained over frames (temporaldelta.go:27-39,
// multiframecompress.go:236-258).  frame 0 holds pixels, frames 1..n-1 hold
// ZigZag residuals; out_f = out_{f-1} + UnZigZag(res_f) mod 2^16 is a running
// sum along the frame axis, kept in a register per pixel (in place).
__global__ void __launch_bounds__(256)
k_temporal_accumulate(uint16_t* __restrict__ frames, unsigned long long fpx, int nframes) {
  const unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x;
  for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < fpx; i += stride) {
    unsigned v = frames[i];
    for (int f = 1; f < nframes; f++) {
      const unsigned r = frames[(unsigned long long)f * fpx + i];
      const unsigned diff = (r >> 1) ^ (0u - (r & 1u));   // UnZigZag (deltazigzagcompressu16.go:113-116)
      v = (v + diff) & 0xFFFFu;
      frames[(unsigned long long)f * fpx + i] = (uint16_t)v;
    }
  }
}

void launch_temporal_accumulate(uint16_t* d_frames, unsigned long long fpx, int nframes, int sm_count, cudaStream_t st) {
  if (nframes <= 1 || fpx == 0) return;
  unsigned long long blocks = (fpx + 255) / 256;
  if (blocks > (unsigned long long)sm_count * 8) blocks = (unsigned long long)sm_count * 8;
  k_temporal_accumulate<<<(unsigned)blocks, 256, 0, st>>>(d_frames, fpx, nframes);
}

}  // namespace micgpu
