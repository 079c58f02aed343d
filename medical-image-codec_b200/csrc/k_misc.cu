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
// two uint16 lanes per word: lane-wise add mod 2^16, lane-wise UnZigZag
__device__ __forceinline__ uint32_t add2_u16(uint32_t a, uint32_t b) {
  return ((a & 0x7FFF7FFFu) + (b & 0x7FFF7FFFu)) ^ ((a ^ b) & 0x80008000u);
}
__device__ __forceinline__ uint32_t unzigzag2_u16(uint32_t r) {
  return ((r >> 1) & 0x7FFF7FFFu) ^ ((r & 0x00010001u) * 0xFFFFu);   // (r >> 1) ^ -(r & 1) in both halves
}

__global__ void __launch_bounds__(256)
k_temporal_accumulate(uint16_t* __restrict__ frames, unsigned long long fpx, int nframes, int first_is_residual, int vec) {
  if (vec) {
    // eight pixels per thread: one 16 B load and one 16 B store per frame, the running sums stay in four registers
    const unsigned long long ngrp = fpx >> 3, stride = (unsigned long long)gridDim.x * blockDim.x;
    for (unsigned long long gi = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; gi < ngrp; gi += stride) {
      uint4* p0 = reinterpret_cast<uint4*>(frames) + gi;
      uint4 v = *p0;
      if (first_is_residual) {
        v.x = unzigzag2_u16(v.x); v.y = unzigzag2_u16(v.y); v.z = unzigzag2_u16(v.z); v.w = unzigzag2_u16(v.w);
        *p0 = v;
      }
      for (int f = 1; f < nframes; f++) {
        uint4* p = reinterpret_cast<uint4*>(frames + (unsigned long long)f * fpx) + gi;
        const uint4 r = *p;
        v.x = add2_u16(v.x, unzigzag2_u16(r.x)); v.y = add2_u16(v.y, unzigzag2_u16(r.y));
        v.z = add2_u16(v.z, unzigzag2_u16(r.z)); v.w = add2_u16(v.w, unzigzag2_u16(r.w));
        *p = v;
      }
    }
    return;
  }
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
  const int vec = (fpx % 8 == 0) && (reinterpret_cast<uintptr_t>(d_frames) % 16 == 0);
  unsigned long long blocks = ((vec ? fpx / 8 : fpx) + 255) / 256;
  if (blocks > (unsigned long long)sm_count * 8) blocks = (unsigned long long)sm_count * 8;
  if (blocks == 0) blocks = 1;
  k_temporal_accumulate<<<(unsigned)blocks, 256, 0, st>>>(d_frames, fpx, nframes, first_is_residual, vec);
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
// DecompressWSIRegion (:282-291).  One job = one rectangle of one tile; constant planes (modes 0/1) are carried in the
// job instead of being materialised.
//   * vector path (whole 16-pixel groups: full-width rectangles of tiles whose width is a multiple of 16, 16 B aligned
//     planes and destination -- every tile of the batch path, every interior tile of a region): a thread loads 2 x 16 B
//     per plane, converts 16 pixels and stores 3 x 16 B (RGB) / 1-2 x 16 B (grey);
//   * scalar path for everything else: rows over blockIdx.x, pixels over threads (no per-pixel division).
__device__ __forceinline__ void ycocg_inv(int y, int co_zz, int cg_zz, int& r, int& g, int& b) {
  const int co = (int)(short)((co_zz >> 1) ^ -(co_zz & 1)), cg = (int)(short)((cg_zz >> 1) ^ -(cg_zz & 1));
  const int t = y - (cg >> 1);
  g = cg + t;
  b = t - (co >> 1);
  r = co + b;
}

__global__ void __launch_bounds__(256)
k_tile_blit(const TileBlitJob* __restrict__ jobs, int njobs, const uint16_t* __restrict__ planes, uint8_t* __restrict__ out) {
  for (int j = blockIdx.y; j < njobs; j += gridDim.y) {
    const TileBlitJob J = jobs[j];
    const unsigned bpp = J.mode <= 1 ? 3u : (J.mode == 2 ? 1u : 2u);
    const unsigned nplanes = J.mode <= 1 ? 3u : 1u;
    uint8_t* dst = out + J.dst_off;
    const uint16_t* pl[3];
#pragma unroll
    for (unsigned k = 0; k < 3; k++) pl[k] = planes + J.plane_off[k < nplanes ? k : 0];
    bool vec = J.src_x == 0 && J.copy_w == J.tile_w && (J.tile_w & 15u) == 0 && J.dst_pitch == J.tile_w * bpp &&
               ((reinterpret_cast<uintptr_t>(dst) & 15u) == 0);
    for (unsigned k = 0; k < nplanes; k++)
      if (!((J.cmask >> k) & 1u)) vec = vec && ((reinterpret_cast<uintptr_t>(pl[k]) & 15u) == 0);
    if (vec) {
      // the rectangle is rows [src_y, src_y + copy_h) of the tile, contiguous on both sides
      const unsigned long long first = (unsigned long long)J.src_y * J.tile_w;
      const unsigned ngroups = (unsigned)(((unsigned long long)J.copy_w * J.copy_h) >> 4);
      for (unsigned gidx = blockIdx.x * blockDim.x + threadIdx.x; gidx < ngroups; gidx += gridDim.x * blockDim.x) {
        uint32_t v[3][8];   // 16 samples per plane, two per word
#pragma unroll
        for (unsigned k = 0; k < 3; k++) {
          if (k < nplanes) {
            if ((J.cmask >> k) & 1u) {
              const uint32_t c2 = (uint32_t)J.cval[k] * 0x00010001u;
#pragma unroll
              for (int q = 0; q < 8; q++) v[k][q] = c2;
            } else {
              const uint4* src = reinterpret_cast<const uint4*>(pl[k] + first) + 2ull * gidx;
              const uint4 a = __ldg(src), b = __ldg(src + 1);
              v[k][0] = a.x; v[k][1] = a.y; v[k][2] = a.z; v[k][3] = a.w;
              v[k][4] = b.x; v[k][5] = b.y; v[k][6] = b.z; v[k][7] = b.w;
            }
          }
        }
        if (J.mode <= 1) {
          uint8_t px[48];
#pragma unroll
          for (int q = 0; q < 16; q++) {
            const int a = (v[0][q >> 1] >> (16 * (q & 1))) & 0xFFFF, b = (v[1][q >> 1] >> (16 * (q & 1))) & 0xFFFF,
                      c = (v[2][q >> 1] >> (16 * (q & 1))) & 0xFFFF;
            int r, g, bl;
            if (J.mode == 0) ycocg_inv(a, b, c, r, g, bl);
            else { r = a; g = b; bl = c; }
            px[3 * q] = (uint8_t)r; px[3 * q + 1] = (uint8_t)g; px[3 * q + 2] = (uint8_t)bl;
          }
          uint4* o = reinterpret_cast<uint4*>(dst) + 3ull * gidx;
#pragma unroll
          for (int w = 0; w < 3; w++) {
            uint32_t x[4];
#pragma unroll
            for (int i = 0; i < 4; i++) {
              const int bidx = 16 * w + 4 * i;
              x[i] = (uint32_t)px[bidx] | ((uint32_t)px[bidx + 1] << 8) | ((uint32_t)px[bidx + 2] << 16) | ((uint32_t)px[bidx + 3] << 24);
            }
            o[w] = make_uint4(x[0], x[1], x[2], x[3]);
          }
        } else if (J.mode == 2) {
          uint32_t x[4];
#pragma unroll
          for (int i = 0; i < 4; i++)
            x[i] = (v[0][2 * i] & 0xFFu) | (((v[0][2 * i] >> 16) & 0xFFu) << 8) | ((v[0][2 * i + 1] & 0xFFu) << 16) | (((v[0][2 * i + 1] >> 16) & 0xFFu) << 24);
          reinterpret_cast<uint4*>(dst)[gidx] = make_uint4(x[0], x[1], x[2], x[3]);
        } else {
          uint4* o = reinterpret_cast<uint4*>(dst) + 2ull * gidx;
          o[0] = make_uint4(v[0][0], v[0][1], v[0][2], v[0][3]);
          o[1] = make_uint4(v[0][4], v[0][5], v[0][6], v[0][7]);
        }
      }
      continue;
    }
    for (unsigned ry = blockIdx.x; ry < J.copy_h; ry += gridDim.x) {
      const unsigned long long srow = (unsigned long long)(J.src_y + ry) * J.tile_w + J.src_x;
      uint8_t* d = dst + (unsigned long long)ry * J.dst_pitch;
      for (unsigned rx = threadIdx.x; rx < J.copy_w; rx += blockDim.x) {
        const unsigned long long s = srow + rx;
        int a = 0, b = 0, c = 0;
        a = (J.cmask & 1u) ? J.cval[0] : pl[0][s];
        if (nplanes == 3) {
          b = (J.cmask & 2u) ? J.cval[1] : pl[1][s];
          c = (J.cmask & 4u) ? J.cval[2] : pl[2][s];
        }
        if (J.mode <= 1) {
          int r, g, bl;
          if (J.mode == 0) ycocg_inv(a, b, c, r, g, bl);
          else { r = a; g = b; bl = c; }
          d[3 * rx] = (uint8_t)r; d[3 * rx + 1] = (uint8_t)g; d[3 * rx + 2] = (uint8_t)bl;
        } else if (J.mode == 2) {
          d[rx] = (uint8_t)a;
        } else {
          d[2 * rx] = (uint8_t)a; d[2 * rx + 1] = (uint8_t)(a >> 8);
        }
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
  // a 256x256 tile is 4096 groups of 16 pixels: 4 CTAs of 256 threads take four groups per thread
  dim3 grid(njobs >= 1024 ? 4 : 16, njobs < 65535 ? njobs : 65535);
  k_tile_blit<<<grid, 256, 0, st>>>(d_jobs, njobs, d_planes, d_out);
}

}  // namespace micgpu
