// k_delta.cu -- K4: inverse avg(top,left) predictor as a blocked anti-diagonal wavefront.
//
// Replaces DeltaRleDecompressU16.DecodeNextSymbol[NC] (deltarlecompressu16.go:102-128)
// and the C twin's delta_decode_simd.  out = ((left+top)>>1) + diff is not linear in
// `left`, so there is no closed-form row scan.  Row 0 IS linear (out = left + diff, with
// literal pixels as resets) and is reconstructed by a segmented warp scan; after that
// thread y owns row y and walks it in blocks of 8 pixels, one block per step, trailing the
// row above by one block (two when the 16-byte phase of the rows wraps):
//   * `left` is the thread's own previous pixel, the 8 `top` pixels arrive as four packed
//     registers shuffled from lane-1 (the block it finished in the previous step) and are
//     re-phased with the previously received block by the row-to-row alignment step
//     delta = W mod 8 (a warp-uniform funnel shift);
//   * rows are addressed in "padded" coordinates p = a_y + x with a_y the 8-pixel phase of
//     the OUTPUT row address, so every block is one aligned 16 B load of the residual plane D
//     (K3 writes D with the same phase) and one aligned 16 B store of pixels;
//   * warps chain through a K4_NB-block shared-memory ring each, handed over with full/empty mbarriers: the consumer
//     sleeps in mbarrier.try_wait instead of polling a counter (the polling loop was 29 % of all issued instructions
//     and the per-warp boundary rows kept residency at two CTAs per SM, profiles/README.md).
// The 8 pixels of a block are unrolled with compile-time register selection: ~11
// instructions per pixel per warp step versus ~130 for a one-pixel-per-step wavefront
// (measured, profiles/), at the price of a longer pipeline fill (8 columns of skew per row).
#include <cstdlib>

#include "mic_device.cuh"

namespace micgpu {

struct Grp {
  uint32_t w[4];  // 8 packed u16, element j in bits 16*(j&1) of w[j>>1]
};

__device__ __forceinline__ uint32_t fs16(uint32_t a, uint32_t b) { return __funnelshift_r(a, b, 16); }

// out = concat(p, c)[8-delta .. 16-delta) in u16 units (delta warp-uniform)
__device__ __forceinline__ Grp rephase(const Grp& p, const Grp& c, int delta) {
  Grp o;
  switch (delta) {
    case 0: o = c; break;
    case 1: o.w[0] = fs16(p.w[3], c.w[0]); o.w[1] = fs16(c.w[0], c.w[1]); o.w[2] = fs16(c.w[1], c.w[2]); o.w[3] = fs16(c.w[2], c.w[3]); break;
    case 2: o.w[0] = p.w[3]; o.w[1] = c.w[0]; o.w[2] = c.w[1]; o.w[3] = c.w[2]; break;
    case 3: o.w[0] = fs16(p.w[2], p.w[3]); o.w[1] = fs16(p.w[3], c.w[0]); o.w[2] = fs16(c.w[0], c.w[1]); o.w[3] = fs16(c.w[1], c.w[2]); break;
    case 4: o.w[0] = p.w[2]; o.w[1] = p.w[3]; o.w[2] = c.w[0]; o.w[3] = c.w[1]; break;
    case 5: o.w[0] = fs16(p.w[1], p.w[2]); o.w[1] = fs16(p.w[2], p.w[3]); o.w[2] = fs16(p.w[3], c.w[0]); o.w[3] = fs16(c.w[0], c.w[1]); break;
    case 6: o.w[0] = p.w[1]; o.w[1] = p.w[2]; o.w[2] = p.w[3]; o.w[3] = c.w[0]; break;
    default: o.w[0] = fs16(p.w[0], p.w[1]); o.w[1] = fs16(p.w[1], p.w[2]); o.w[2] = fs16(p.w[2], p.w[3]); o.w[3] = fs16(p.w[3], c.w[0]); break;
  }
  return o;
}

template <int J>
__device__ __forceinline__ uint32_t grp_get(const Grp& g) {
  return (J & 1) ? (g.w[J >> 1] >> 16) : (g.w[J >> 1] & 0xFFFFu);
}

constexpr int K4_NB = 8;      // blocks in a warp-to-warp boundary ring

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_inval(uint32_t bar) {
  asm volatile("mbarrier.inval.shared::cta.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
// blocks (suspended by the hardware, not spinning on issue slots) until the phase of the given parity completed
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n\t.reg .pred p;\n"
      "MBW_%=:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1, 0x989680;\n\t"   // suspend-time hint: wake on completion
      "@!p bra MBW_%=;\n\t"
      "}\n" ::"r"(bar), "r"(parity)
      : "memory");
}
constexpr int K4_RING_DEFAULT = 4;    // per-lane cp.async ring depth (blocks of 16 B) for the residual plane
constexpr int K4_MRING = 4;           // literal-mask words (4 blocks each) kept in flight per lane


template <int K4_RING>
__global__ void __launch_bounds__(1024)
k_delta_wavefront(MicUnit* __restrict__ units, const int* __restrict__ list, int nlist,
                  const uint16_t* __restrict__ D, const uint32_t* __restrict__ M, uint16_t* __restrict__ out,
                  int brow_pitch, int redo_only) {
  // dynamic shared memory: row_in | row_out (full rows, padded coordinates) | boundary rings | exchange slots |
  // residual ring | mask ring | mbarriers
  extern __shared__ __align__(16) uint16_t s_brow[];

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int nwarps = blockDim.x >> 5;
  const MicUnit* U = &units[list[blockIdx.x]];
  if (U->status != MIC_OK || U->predictor != 0u) return;   // gradient-predictor units belong to k_grad_wavefront
  if (redo_only && !U->k4_redo) return;   // only the units the row-scan kernel (k_delta_scan.cu) handed back
  const int W = (int)U->width, H = (int)U->height;
  const unsigned wp = U->wp;
  const int thr = (int)U->thr;
  const int align0 = (int)U->align0;
  const int delta = W & 7;
  const uint16_t* Du = D + U->d_off;
  const uint32_t* Mu = M + U->m_off;
  uint16_t* Ou = out + U->out_off;

  // ---------------- row 0: segmented prefix sum (warp 0) -----------------------------------
  if (warp == 0) {
    const int a0 = align0 & 7;
    unsigned carry = 0;
    for (int base = 0; base < W; base += 32) {
      const int x = base + lane;
      const bool valid = x < W;
      unsigned d = 0, flag = 0;
      if (valid) {
        d = Du[a0 + x];
        flag = (Mu[(a0 + x) >> 5] >> ((a0 + x) & 31)) & 1u;
      }
      unsigned val = flag ? d : (d - (unsigned)thr);   // row 0: pred = left (0 at x = 0)
#pragma unroll
      for (int off = 1; off < 32; off <<= 1) {
        const unsigned v2 = __shfl_up_sync(0xffffffffu, val, off);
        const unsigned f2 = __shfl_up_sync(0xffffffffu, flag, off);
        if (lane >= off) {
          if (!flag) val += v2;
          flag |= f2;
        }
      }
      if (!flag) val += carry;
      val &= 0xFFFFu;
      if (valid) {
        Ou[x] = (uint16_t)val;
        s_brow[a0 + x] = (uint16_t)val;
      }
      carry = __shfl_sync(0xffffffffu, val, 31);
    }
  }

  uint4* bring_all = reinterpret_cast<uint4*>(s_brow + 2 * (size_t)brow_pitch);        // nwarps rings of K4_NB blocks
  uint4* xch_all = bring_all + nwarps * K4_NB;
  uint4* dring_all = xch_all + nwarps * 64;
  uint32_t* mring_all = reinterpret_cast<uint32_t*>(dring_all + nwarps * (K4_RING * 32));
  uint64_t* bars = reinterpret_cast<uint64_t*>(mring_all + nwarps * (K4_MRING * 32));   // [warp][full K4_NB | empty K4_NB]
  const uint32_t bars_sa = (uint32_t)__cvta_generic_to_shared(bars);

  const int rows_per_pass = (int)blockDim.x;
  const int passes = (H - 1 + rows_per_pass - 1) / rows_per_pass;
  for (int p = 0; p < passes; p++) {
    const int y = 1 + p * rows_per_pass + tid;
    const int wy0 = 1 + p * rows_per_pass + warp * 32;   // first row of this warp
    __syncthreads();
    if (p > 0) {
      for (int i = tid; i < brow_pitch; i += blockDim.x) s_brow[i] = s_brow[(size_t)brow_pitch + i];
      if (tid < nwarps * 2 * K4_NB) mbar_inval(bars_sa + 8u * tid);
      __syncthreads();
    }
    if (tid < nwarps * 2 * K4_NB) mbar_init(bars_sa + 8u * tid, 1u);
    __syncthreads();
    if (wy0 >= H) continue;   // whole warp idle this pass (still reaches the barriers above)

    const bool row_active = y < H;
    const int a = (align0 + (y & 7) * delta) & 7;             // phase of this row: (align0 + y*W) mod 8
    const int nblk = row_active ? (a + W + 7) >> 3 : 0;
    const int wrap = (delta != 0 && a < delta) ? 1 : 0;       // phase wrapped relative to row y-1
    const int nblk_prev = (((a - delta) & 7) + W + 7) >> 3;   // blocks of row y-1
    // start step: lane l trails lane l-1 by one block, two when the phase wrapped
    int s = lane == 0 ? 0 : 1 + wrap;
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
      const int t2 = __shfl_up_sync(0xffffffffu, s, off);
      if (lane >= off) s += t2;
    }
    const int T = __reduce_max_sync(0xffffffffu, s + nblk);

    // input of lane 0: warp 0 reads the full row (row 0 / last row of the previous pass), the others their ring
    const uint16_t* row_in = s_brow;
    uint16_t* row_out = s_brow + brow_pitch;
    const uint4* ring_in = bring_all + warp * K4_NB;
    const uint32_t full_in = bars_sa + (uint32_t)warp * (2 * K4_NB * 8), empty_in = full_in + K4_NB * 8;
    const bool chained = warp + 1 < nwarps && wy0 + 32 < H;   // a warp below consumes this warp's last row
    uint4* ring_out = bring_all + (warp + 1) * K4_NB;
    const uint32_t full_out = bars_sa + (uint32_t)(warp + 1) * (2 * K4_NB * 8), empty_out = full_out + K4_NB * 8;

    const size_t yr = (size_t)(row_active ? y : 1);
    const uint16_t* Drow = Du + yr * wp;
    const uint32_t* Mrow = Mu + yr * (wp >> 5);
    uint16_t* Orow = Ou + yr * W;
    uint16_t* Oal = Orow - a;     // 16-byte aligned: block b is Oal[8b .. 8b+8)

    // lane -> lane+1 hand-over of the finished block through shared memory (double buffered):
    // one 16 B store, one warp barrier and one 16 B load per step instead of four shuffles
    uint4* xch = xch_all + warp * 64;
    xch[lane] = make_uint4(0, 0, 0, 0);
    xch[32 + lane] = make_uint4(0, 0, 0, 0);

    // Residual blocks are staged K4_RING-1 steps ahead with cp.async (LDGSTS) into a per-lane ring: the
    // 256 coupled rows of a unit advance at the pace of the slowest lane, so DRAM tail latency must be
    // hidden far deeper than two register-prefetched blocks can (measured: 4.7k cycles per step before).
    uint4* dring = dring_all + warp * (K4_RING * 32);
    // The literal mask travels the same way: word q of the row (blocks 4q..4q+3) is copied when block 4q is staged.
    // (It used to be a register load two words ahead; with 256 lock-stepped rows that exposed DRAM latency, 20 % of
    // the kernel's stall samples.)
    uint32_t* mring = mring_all + warp * (K4_MRING * 32);
    auto stage = [&](int pb) {
      if (pb >= 0 && pb < nblk) {
        const unsigned dst = (unsigned)__cvta_generic_to_shared(dring + (pb & (K4_RING - 1)) * 32 + lane);
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(Drow + 8 * pb));
        if ((pb & 3) == 0) {
          const unsigned mdst = (unsigned)__cvta_generic_to_shared(mring + ((pb >> 2) & (K4_MRING - 1)) * 32 + lane);
          asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(mdst), "l"(Mrow + (pb >> 2)));
        }
      }
      asm volatile("cp.async.commit_group;");
    };
#pragma unroll
    for (int q = 0; q < K4_RING - 1; q++) stage(-s + q);

    Grp tp = {{0, 0, 0, 0}}, g = {{0, 0, 0, 0}};
    unsigned left = 0;
    int b = -s;
    for (int t = 0; t < T; t++, b++) {
      __syncwarp();
      Grp rc;
      {
        const uint4 v = xch[((t + 1) & 1) * 32 + ((lane + 31) & 31)];   // lane-1's block of step t-1
        rc.w[0] = v.x; rc.w[1] = v.y; rc.w[2] = v.z; rc.w[3] = v.w;
      }
      if (lane == 0) {
        const int c = t + wrap;
        if (c < nblk_prev) {
          uint4 v;
          if (warp == 0) {
            v = *reinterpret_cast<const uint4*>(row_in + 8 * c);
            // lane 0 starts one block late relative to the row above when the phase wrapped: its first
            // window also needs block 0 of that row (the other lanes receive it during their start delay)
            if (t == 0 && wrap) {
              const uint4 v0 = *reinterpret_cast<const uint4*>(row_in);
              tp.w[0] = v0.x; tp.w[1] = v0.y; tp.w[2] = v0.z; tp.w[3] = v0.w;
            }
          } else {
            const int slot = c & (K4_NB - 1);
            mbar_wait(full_in + 8u * slot, (unsigned)(c / K4_NB) & 1u);   // blocks arrive in order: block c implies 0..c
            v = ring_in[slot];
            if (t == 0 && wrap) {
              const uint4 v0 = ring_in[0];
              tp.w[0] = v0.x; tp.w[1] = v0.y; tp.w[2] = v0.z; tp.w[3] = v0.w;
              mbar_arrive(empty_in);
            }
            mbar_arrive(empty_in + 8u * slot);
          }
          rc.w[0] = v.x; rc.w[1] = v.y; rc.w[2] = v.z; rc.w[3] = v.w;
        }
      }
      stage(b + K4_RING - 1);
      asm volatile("cp.async.wait_group %0;" ::"n"(K4_RING - 1));
      if (b >= 0 && b < nblk) {
        const Grp tw = rephase(tp, rc, delta);
        const uint4 dv = dring[(b & (K4_RING - 1)) * 32 + lane];
        const unsigned mbyte = (mring[((b >> 2) & (K4_MRING - 1)) * 32 + lane] >> ((b & 3) * 8)) & 0xFFu;
        Grp dg;
        dg.w[0] = dv.x; dg.w[1] = dv.y; dg.w[2] = dv.z; dg.w[3] = dv.w;
        const int x0 = 8 * b - a;
        if (mbyte == 0 && x0 >= 1 && x0 + 8 <= W) {
          // ---- interior block: 8 unrolled pixels, aligned 16 B store ----
          unsigned v[8];
#define PIX(J)                                                                        \
  {                                                                                   \
    const unsigned top = grp_get<J>(tw);                                              \
    const unsigned e = grp_get<J>(dg) - (unsigned)thr;                                \
    left = (((left + top) >> 1) + e) & 0xFFFFu;                                       \
    v[J] = left;                                                                      \
  }
          PIX(0) PIX(1) PIX(2) PIX(3) PIX(4) PIX(5) PIX(6) PIX(7)
#undef PIX
          g.w[0] = v[0] | (v[1] << 16); g.w[1] = v[2] | (v[3] << 16);
          g.w[2] = v[4] | (v[5] << 16); g.w[3] = v[6] | (v[7] << 16);
          *reinterpret_cast<uint4*>(Oal + 8 * b) = make_uint4(g.w[0], g.w[1], g.w[2], g.w[3]);
        } else {
          // ---- head / tail / literal block: per-pixel predicates, 2-byte stores ----
          unsigned v[8];
#define PIXS(J)                                                                       \
  {                                                                                   \
    const int x = x0 + J;                                                             \
    v[J] = 0;                                                                         \
    if (x >= 0 && x < W) {                                                            \
      const unsigned top = grp_get<J>(tw);                                            \
      const unsigned d = grp_get<J>(dg);                                              \
      const unsigned l = x == 0 ? top : left;          /* column 0: pred = top */     \
      left = ((mbyte >> J) & 1u) ? d : ((((l + top) >> 1) + d - (unsigned)thr) & 0xFFFFu); \
      v[J] = left;                                                                    \
      Orow[x] = (uint16_t)left;                                                       \
    }                                                                                 \
  }
          PIXS(0) PIXS(1) PIXS(2) PIXS(3) PIXS(4) PIXS(5) PIXS(6) PIXS(7)
#undef PIXS
          g.w[0] = v[0] | (v[1] << 16); g.w[1] = v[2] | (v[3] << 16);
          g.w[2] = v[4] | (v[5] << 16); g.w[3] = v[6] | (v[7] << 16);
        }
        if (lane == 31) {
          if (chained) {
            const int slot = b & (K4_NB - 1);
            mbar_wait(empty_out + 8u * slot, ((unsigned)(b / K4_NB) & 1u) ^ 1u);   // slot consumed K4_NB blocks ago
            ring_out[slot] = make_uint4(g.w[0], g.w[1], g.w[2], g.w[3]);
            mbar_arrive(full_out + 8u * slot);
          } else {
            *reinterpret_cast<uint4*>(row_out + 8 * b) = make_uint4(g.w[0], g.w[1], g.w[2], g.w[3]);
          }
        }
      }
      xch[(t & 1) * 32 + lane] = make_uint4(g.w[0], g.w[1], g.w[2], g.w[3]);
      tp = rc;
    }
  }
}

static int k4_ring() {
  static const int v = [] { const char* e = getenv("MICGPU_K4_RING"); return e && atoi(e) == 8 ? 8 : K4_RING_DEFAULT; }();
  return v;
}
static size_t k4_warp_bytes() { return K4_NB * 16 + 1024 + (size_t)k4_ring() * 512 + K4_MRING * 128 + 2 * K4_NB * 8; }

int delta_wavefront_threads(int max_width, int max_height) {
  int threads = (max_height - 1 + 31) / 32 * 32;   // row 0 is handled by the prologue scan
  if (threads > 1024) threads = 1024;
  if (threads < 32) threads = 32;
  const int pitch = (max_width + 16 + 7) / 8 * 8;
  // keep (nwarps+1) boundary rows within ~200 KB of shared memory
  while (threads > 32 && 2 * (size_t)pitch * sizeof(uint16_t) + (size_t)(threads / 32) * k4_warp_bytes() > 200u * 1024u) threads -= 32;
  return threads;
}

void launch_delta_wavefront(MicUnit* d_units, const int* d_list, int nlist, const uint16_t* d_D, const uint32_t* d_M,
                            uint16_t* d_out, int max_width, int max_height, cudaStream_t st, int redo_only) {
  if (nlist <= 0) return;
  const int threads = delta_wavefront_threads(max_width, max_height);
  const int nwarps = threads / 32;
  const int pitch = (max_width + 16 + 7) / 8 * 8;
  const size_t smem = 2 * (size_t)pitch * sizeof(uint16_t) + (size_t)nwarps * k4_warp_bytes();
  if (k4_ring() == 4) {
    cudaFuncSetAttribute(k_delta_wavefront<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    k_delta_wavefront<4><<<nlist, threads, smem, st>>>(d_units, d_list, nlist, d_D, d_M, d_out, pitch, redo_only);
  } else {
    cudaFuncSetAttribute(k_delta_wavefront<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    k_delta_wavefront<8><<<nlist, threads, smem, st>>>(d_units, d_list, nlist, d_D, d_M, d_out, pitch, redo_only);
  }
}

}  // namespace micgpu
