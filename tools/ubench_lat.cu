// ubench_lat.cu -- dependent-chain latencies of the instructions on K2's state->state chain (sm_100a).
// One warp, 8 active lanes (like k_ans_decode<8,*>), clock64 around an unrolled dependent chain.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/ubench_lat tools/ubench_lat.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

constexpr int ITERS = 2048;

template <int WHICH, int LANES>
__global__ void k(uint32_t* out, long long* cyc, uint32_t seed) {
  __shared__ uint16_t tab[8192];
  __shared__ uint32_t ring[40];
  const int lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < 8192; i += blockDim.x) tab[i] = (uint16_t)((i * 2654435761u + seed) >> 19);
  if (threadIdx.x < 40) ring[threadIdx.x] = threadIdx.x * 0x9E3779B9u + seed;
  __syncthreads();
  if (lane >= LANES) return;
  __shared__ __align__(16) uint8_t xch[64];
  uint32_t mlo = 0, mhi = 0;
  for (int f = 0; f < (lane & 7); f++) { if (f < 4) mlo |= 0xFFu << (8 * f); else mhi |= 0xFFu << (8 * (f - 4)); }
  uint32_t x = (seed + lane * 977u) & 8191u;
  uint32_t acc = 0;
  const uint32_t cmul = 0x01010101u << (8 * (lane & 3));
  long long t0 = clock64();
#pragma unroll 8
  for (int i = 0; i < ITERS; i++) {
    if (WHICH == 0) {            // LDS.U16 pointer chase
      x = tab[x & 8191u];
    } else if (WHICH == 1) {     // FLO chain
      x = (31u - __clz(x | 1u)) + (x << 3) + 77u + seed;
    } else if (WHICH == 2) {     // REDUX.SUM -> R chain (multiply, redux, extract)
      uint32_t s = __reduce_add_sync(0xFFu, (x & 15u) * cmul);
      x = ((s >> (8 * (lane & 3))) & 0xFFu) + (x >> 1) + 3u;
    } else if (WHICH == 3) {     // SHFL.UP chain
      uint32_t s = __shfl_up_sync(0xFFu, x, 1, 8);
      x = s + (x >> 1) + 3u;
    } else if (WHICH == 4) {     // SHFL.IDX chain
      uint32_t s = __shfl_sync(0xFFu, x, (x >> 2) & 7, 8);
      x = s + 3u;
    } else if (WHICH == 5) {     // LDS 2 words + funnel shift (ring extract) chain
      const uint32_t* w = reinterpret_cast<const uint32_t*>(reinterpret_cast<const uint8_t*>(ring) + ((x >> 3) & 0x7Cu));
      x = __funnelshift_r(w[0], w[1], x & 31u) + 5u;
    } else if (WHICH == 6) {     // plain ALU chain (IADD3/LOP3): 4 dependent ops per iteration
      x = ((x + 3u) ^ 0x55u) + (x >> 1);
    } else if (WHICH == 7) {     // the whole K2 round (MODE 1 shape), single warp, no contention
      const uint32_t nx = tab[x & 8191u];
      const uint32_t nb = 13u - (31u - __clz(nx | 1u)) & 15u;
      const uint32_t a = __reduce_add_sync(0xFFu, nb * cmul);
      const uint32_t b = __reduce_add_sync(0xFFu, nb * (cmul >> 8));
      const uint32_t tot = __reduce_add_sync(0xFFu, nb);
      const uint32_t before = (((lane & 4) ? b : a) >> (8 * (lane & 3))) & 0xFFu;
      const uint32_t lo = acc - before - nb;
      const uint32_t* w = reinterpret_cast<const uint32_t*>(reinterpret_cast<const uint8_t*>(ring) + ((lo >> 3) & 0x7Cu));
      const uint32_t bits = __funnelshift_r(w[0], w[1], lo & 31u) & ((1u << nb) - 1u);
      out[i * 8 + lane] = (uint16_t)x;
      x = ((nx << nb) + bits) & 8191u;
      acc -= tot;
    } else if (WHICH == 8) {     // STS.U8 -> LDS.64 -> masked dp4a (packed prefix) + 2 alu
      xch[lane] = (uint8_t)(x & 15u);
      __syncwarp();
      const uint2 v = *reinterpret_cast<const uint2*>(xch + (lane & ~7));
      const uint32_t before = __dp4a(v.x & mlo, 0x01010101u, __dp4a(v.y & mhi, 0x01010101u, 0u));
      x = before + (x >> 1) + 3u;
    } else if (WHICH == 9) {     // full packed round
      const uint32_t nx = tab[x & 8191u];
      uint32_t nb = 13u - (31u - __clz(nx | 1u)) & 15u;
      if (i > ITERS - (int)(seed & 3)) nb = 0;
      xch[lane + (i & 1) * 32] = (uint8_t)nb;
      __syncwarp();
      const uint2 v = *reinterpret_cast<const uint2*>(xch + (lane & ~7) + (i & 1) * 32);
      const uint32_t tot = __dp4a(v.x, 0x01010101u, __dp4a(v.y, 0x01010101u, 0u));
      const uint32_t before = __dp4a(v.x & mlo, 0x01010101u, __dp4a(v.y & mhi, 0x01010101u, 0u));
      const uint32_t lo = acc - before - nb;
      const uint32_t* w = reinterpret_cast<const uint32_t*>(reinterpret_cast<const uint8_t*>(ring) + ((lo >> 3) & 0x7Cu));
      const uint32_t bits = __funnelshift_r(w[0], w[1], lo & 31u) & ((1u << nb) - 1u);
      out[i * 32 + lane] = (uint16_t)x;
      x = ((nx << nb) + bits) & 8191u;
      acc -= tot;
    }
  }
  long long t1 = clock64();
  out[lane] = x + acc;
  if (lane == 0) cyc[WHICH + (LANES == 32 ? 16 : 0)] = t1 - t0;
}

int main() {
  uint32_t* out; long long* cyc;
  cudaMalloc(&out, ITERS * 32 * 4 + 64);
  cudaMallocManaged(&cyc, 64 * 8);
  const char* names[] = {"LDS.U16 chase", "FLO chain(+2 alu)", "IMAD+REDUX+SHF+LOP+2alu", "SHFL.UP+2alu", "SHFL.IDX+2alu(+shf,lop)",
                         "ring extract (4 alu + LDS + SHF.W + IADD)", "4 alu", "full K2 round (redux)",
                         "STS.U8+LDS.64+LOP+2xIDP+2alu", "full packed round"};
  for (int rep = 0; rep < 2; rep++) {
#define RUN(W) k<W, 8><<<1, 32>>>(out, cyc, 1); k<W, 32><<<1, 32>>>(out, cyc, 1);
    RUN(0) RUN(1) RUN(3) RUN(4) RUN(5) RUN(6) RUN(8) RUN(9)
    k<2, 8><<<1, 32>>>(out, cyc, 1); k<7, 8><<<1, 32>>>(out, cyc, 1);
    cudaDeviceSynchronize();
  }
  for (int i = 0; i < 10; i++)
    printf("%-45s  8 lanes %7.1f   32 lanes %7.1f cycles/iter\n", names[i], (double)cyc[i] / ITERS, (double)cyc[i + 16] / ITERS);
  printf("err=%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
