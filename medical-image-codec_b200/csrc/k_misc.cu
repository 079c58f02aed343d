// k_misc.cu -- small elementwise kernels of the MIC path.
#include "mic_device.cuh"

namespace micgpu {

// K10: TemporalDeltaDecode chained over frames (temporaldelta.go:27-39,
// multiframecompress.go:236-258).  frame 0 holds pixels, frames 1..n-1 hold
// ZigZag residuals; out_f = out_{f-1} + UnZigZag(res_f) mod 2^16 is a running
// sum along the frame axis, kept in a register per pixel (in place).
// first_is_residual: the group is a frame range that starts inside a temporal stack (a shard of it): frame 0 is a
// residual too and the sums are relative to a zero carry; k_temporal_add_carry finishes them once the carry frame
// (the absolute last frame of the previous shard) is known.
__global__ void __launch_bounds__(256)
k_temporal_accumulate(uint16_t* __restrict__ frames, unsigned long long fpx, int nframes, int first_is_residual) {
  const unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x;
  for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < fpx; i += stride) {
    unsigned v = frames[i];
    if (first_is_residual) {
      v = ((v >> 1) ^ (0u - (v & 1u))) & 0xFFFFu;
      frames[i] = (uint16_t)v;
    }
    for (int f = 1; f < nframes; f++) {
      const unsigned r = frames[(unsigned long long)f * fpx + i];
      const unsigned diff = (r >> 1) ^ (0u - (r & 1u));   // UnZigZag (deltazigzagcompressu16.go:113-116)
      v = (v + diff) & 0xFFFFu;
      frames[(unsigned long long)f * fpx + i] = (uint16_t)v;
    }
  }
}

void launch_temporal_accumulate(uint16_t* d_frames, unsigned long long fpx, int nframes, int first_is_residual, int sm_count,
                                cudaStream_t st) {
  if ((nframes <= 1 && !first_is_residual) || fpx == 0 || nframes <= 0) return;
  unsigned long long blocks = (fpx + 255) / 256;
  if (blocks > (unsigned long long)sm_count * 8) blocks = (unsigned long long)sm_count * 8;
  k_temporal_accumulate<<<(unsigned)blocks, 256, 0, st>>>(d_frames, fpx, nframes, first_is_residual);
}

// frames[f][i] += carry[i]  (mod 2^16): the one exchange step of a temporal stack sharded over GPUs
__global__ void __launch_bounds__(256)
k_temporal_add_carry(uint16_t* __restrict__ frames, const uint16_t* __restrict__ carry, unsigned long long fpx, int nframes) {
  const unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x;
  for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < fpx; i += stride) {
    const unsigned c = carry[i];
    for (int f = 0; f < nframes; f++) {
      const unsigned long long at = (unsigned long long)f * fpx + i;
      frames[at] = (uint16_t)(frames[at] + c);
    }
  }
}

// The exchange fused with its consumer: the carry of this range is the sum of the last frames of the ranges before it, and
// those frames sit in the memory of the peer GPUs.  One kernel reads them over NVLink (peer pointers opened through CUDA
// IPC, coalesced 16 B loads), forms the carry and finishes every local frame -- no staging copy, no separate collective.
__global__ void __launch_bounds__(256)
k_temporal_add_carry_peers(uint16_t* __restrict__ frames, PeerFrames peers, unsigned long long fpx, int nframes) {
  const unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x * 8ull;
  for (unsigned long long i0 = ((unsigned long long)blockIdx.x * blockDim.x + threadIdx.x) * 8ull; i0 < fpx; i0 += stride) {
    if (i0 + 8 <= fpx && peers.aligned) {
      uint32_t c[4] = {0, 0, 0, 0};     // two u16 lanes per word, added lane-wise mod 2^16
      for (int q = 0; q < peers.n; q++) {
        const uint4 v = *reinterpret_cast<const uint4*>(peers.last[q] + i0);
        const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int j = 0; j < 4; j++) c[j] = ((c[j] + (w[j] & 0xFFFFu)) & 0xFFFFu) | ((c[j] & 0xFFFF0000u) + (w[j] & 0xFFFF0000u));
      }
      for (int f = 0; f < nframes; f++) {
        uint4* p = reinterpret_cast<uint4*>(frames + (unsigned long long)f * fpx + i0);
        uint4 v = *p;
        uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int j = 0; j < 4; j++) w[j] = ((w[j] + (c[j] & 0xFFFFu)) & 0xFFFFu) | ((w[j] & 0xFFFF0000u) + (c[j] & 0xFFFF0000u));
        *p = make_uint4(w[0], w[1], w[2], w[3]);
      }
    } else {
      for (unsigned long long i = i0; i < fpx && i < i0 + 8; i++) {
        unsigned c = 0;
        for (int q = 0; q < peers.n; q++) c += peers.last[q][i];
        for (int f = 0; f < nframes; f++) {
          const unsigned long long at = (unsigned long long)f * fpx + i;
          frames[at] = (uint16_t)(frames[at] + c);
        }
      }
    }
  }
}

void launch_temporal_add_carry_peers(uint16_t* d_frames, const PeerFrames& peers, unsigned long long fpx, int nframes, int sm_count,
                                     cudaStream_t st) {
  if (nframes <= 0 || fpx == 0 || peers.n <= 0) return;
  unsigned long long blocks = (fpx / 8 + 255) / 256 + 1;
  if (blocks > (unsigned long long)sm_count * 8) blocks = (unsigned long long)sm_count * 8;
  k_temporal_add_carry_peers<<<(unsigned)blocks, 256, 0, st>>>(d_frames, peers, fpx, nframes);
}

void launch_temporal_add_carry(uint16_t* d_frames, const uint16_t* d_carry, unsigned long long fpx, int nframes, int sm_count,
                               cudaStream_t st) {
  if (nframes <= 0 || fpx == 0) return;
  unsigned long long blocks = (fpx + 255) / 256;
  if (blocks > (unsigned long long)sm_count * 8) blocks = (unsigned long long)sm_count * 8;
  k_temporal_add_carry<<<(unsigned)blocks, 256, 0, st>>>(d_frames, d_carry, fpx, nframes);
}

}  // namespace micgpu

namespace micgpu {

// ---- MIC3 plane fill: plane modes 0/1/3 (wsicompress.go:487-524) ---------------------------------
// job.kind 0: constant value; 1: raw little-endian u16 copied from the compressed buffer (unaligned).
__global__ void __launch_bounds__(256)
k_plane_fill(const PlaneFillJob* __restrict__ jobs, int njobs, const uint8_t* __restrict__ comp, uint16_t* __restrict__ planes) {
  for (int j = blockIdx.y; j < njobs; j += gridDim.y) {
    const PlaneFillJob J = jobs[j];
    uint16_t* dst = planes + J.plane_off;
    const unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x;
    if (J.kind == 0) {
      const uint16_t v = (uint16_t)J.value;
      for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < J.n; i += stride) dst[i] = v;
    } else {
      const uint8_t* src = comp + J.comp_off;
      for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < J.n; i += stride)
        dst[i] = (uint16_t)(src[2 * i] | (src[2 * i + 1] << 8));
    }
  }
}

// ---- MIC3 tile finish: planes -> pixels, YCoCg-R inverse, crop / region blit -----------------------
// Replaces YCoCgRInverse (ycocgr.go:28-35, asm_generic.go:40-53, ycocgRInverseSSSE3/NEON), the planar->RGB
// interleave (wsicompress.go:466-473), uint16ToBytes (:592-603), cropTile (:558-572) and the row copies of
// DecompressWSIRegion (:282-291).  One job = one rectangle of one tile.
__global__ void __launch_bounds__(256)
k_tile_blit(const TileBlitJob* __restrict__ jobs, int njobs, const uint16_t* __restrict__ planes, uint8_t* __restrict__ out) {
  for (int j = blockIdx.y; j < njobs; j += gridDim.y) {
    const TileBlitJob J = jobs[j];
    const unsigned long long npx = (unsigned long long)J.copy_w * J.copy_h;
    const unsigned long long plane_px = (unsigned long long)J.tile_w * J.tile_h;
    const uint16_t* p0 = planes + J.plane_off;
    uint8_t* dst = out + J.dst_off;
    for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < npx;
         i += (unsigned long long)gridDim.x * blockDim.x) {
      const unsigned ry = (unsigned)(i / J.copy_w), rx = (unsigned)(i - (unsigned long long)ry * J.copy_w);
      const unsigned long long s = (unsigned long long)(J.src_y + ry) * J.tile_w + (J.src_x + rx);
      uint8_t* d = dst + (unsigned long long)ry * J.dst_pitch;
      if (J.mode == 0 || J.mode == 1) {
        const int a = p0[s], b = p0[plane_px + s], c = p0[2 * plane_px + s];
        int r, g, bl;
        if (J.mode == 0) {                  // YCoCg-R inverse
          const int co = (int)(short)((b >> 1) ^ -(b & 1)), cg = (int)(short)((c >> 1) ^ -(c & 1));
          const int t = a - (cg >> 1);
          g = cg + t;
          bl = t - (co >> 1);
          r = co + bl;
        } else {                            // planar R,G,B
          r = a; g = b; bl = c;
        }
        d[3 * rx] = (uint8_t)r; d[3 * rx + 1] = (uint8_t)g; d[3 * rx + 2] = (uint8_t)bl;
      } else if (J.mode == 2) {             // grey 8-bit
        d[rx] = (uint8_t)p0[s];
      } else {                              // grey 16-bit little endian
        const unsigned v = p0[s];
        d[2 * rx] = (uint8_t)v; d[2 * rx + 1] = (uint8_t)(v >> 8);
      }
    }
  }
}

void launch_plane_fill(const PlaneFillJob* d_jobs, int njobs, const uint8_t* d_comp, uint16_t* d_planes, cudaStream_t st) {
  if (njobs <= 0) return;
  dim3 grid(8, njobs < 65535 ? njobs : 65535);
  k_plane_fill<<<grid, 256, 0, st>>>(d_jobs, njobs, d_comp, d_planes);
}

void launch_tile_blit(const TileBlitJob* d_jobs, int njobs, const uint16_t* d_planes, uint8_t* d_out, cudaStream_t st) {
  if (njobs <= 0) return;
  dim3 grid(8, njobs < 65535 ? njobs : 65535);
  k_tile_blit<<<grid, 256, 0, st>>>(d_jobs, njobs, d_planes, d_out);
}

}  // namespace micgpu
