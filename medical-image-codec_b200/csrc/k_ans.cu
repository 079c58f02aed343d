// k_ans.cu -- K2: interleaved N-state tANS / rANS decode (N in {1,2,4,8}).
//
// Replaces decompress (fsedecompressu16.go:267-377), decompress2State
// (fse2state.go:203-308), decompress4State (fse4state.go:195-353),
// decompress8State (fse8state.go:230-380), ransDecompress8State
// (rans8state.go:223-412) and the amd64/arm64 assembly kernels.
//
// Mapping: one "slot" = N adjacent lanes of a warp decodes one unit; lane k owns
// state k.  All N states share ONE bitstream that is consumed round-robin
// (A,B,..), so within a round lane k reads its nbBits field right below the
// fields of lanes < k: the offset is an exclusive prefix sum of nbBits over the
// slot, obtained with one or two redux.sync.or over byte-packed lanes.
// One warp hosts exactly one slot (lanes >= N exit): with several slots per warp the
// partial-mask collectives are serialised by the compiler (WARPSYNC.COLLECTIVE) and a
// lock-step multi-slot warp has too little thread-level parallelism for this
// latency-bound chain -- measured 54 ms and 27 ms versus this layout (profiles/).
// The decode table lives in shared memory (4 B or 2 B per cell) and the
// bitstream is staged through a 128 B per-slot shared-memory ring that is
// refilled from registers loaded one 64 B half ahead (plus an L2 prefetch
// further ahead), so no global-memory latency sits on the state->state chain.
//
// Output is the *state* stream (the table index each symbol was emitted from);
// K3 maps states to symbols through tabS, which keeps the 16-bit symbol out of
// this kernel's shared-memory footprint.
//
// Bit coordinates: the reference reader starts at the last byte, skips the
// padding and the 1-bit sentinel and reads fields MSB-first walking to byte 0
// (bitreader.go:26-62).  With the stream viewed as one little-endian integer,
// a field of n bits read when P bits remain unread is bits [P-n, P).
#include "mic_device.cuh"

namespace micgpu {

constexpr int RING_WORDS = 32;    // ring[32] mirrors ring[0] so word pairs never wrap
constexpr int RING_STRIDE = 36;   // words reserved per slot (16 B aligned)
constexpr int HALF_WORDS = 16;

// Exclusive prefix sum and total of nb (<= 16) over the N active lanes of the warp (lanes >= N have exited, so MASK is
// the warp's whole active mask and each redux is one instruction).  Lane j multiplies its nb by a per-lane constant
// that replicates it into the 8-bit fields of all lanes k > j; redux.add then delivers EVERY lane's prefix at once
// (field k of the sum), so the post-processing is one shift and one mask instead of a masked horizontal sum.
// Field sums stay below 256 (7 * 16).  Fields 0-3 live in word A, fields 4-7 in word B.
template <int N>
struct PrefixConsts {
  uint32_t ca, cb;     // multipliers of this lane for words A and B
  uint32_t sh;         // shift that extracts this lane's field
  bool use_b;
  __device__ explicit PrefixConsts(int k) {
    ca = 0; cb = 0;
#pragma unroll
    for (int f = 0; f < 8; f++) {
      if (f > k && f < N) {
        if (f < 4) ca |= 1u << (8 * f);
        else cb |= 1u << (8 * (f - 4));
      }
    }
    sh = 8u * (unsigned)(k & 3);
    use_b = k >= 4;
  }
};

template <int N>
__device__ __forceinline__ uint32_t slot_prefix(uint32_t nb, const PrefixConsts<N>& pc, uint32_t* tot) {
  constexpr unsigned MASK = (N == 32) ? 0xffffffffu : ((1u << N) - 1u);
  if (N == 1) {
    *tot = nb;
    return 0;
  } else {
    const uint32_t a = __reduce_add_sync(MASK, nb * pc.ca);
    *tot = __reduce_add_sync(MASK, nb);
    uint32_t w = a;
    if (N > 4) {
      const uint32_t b = __reduce_add_sync(MASK, nb * pc.cb);
      w = pc.use_b ? b : a;
    }
    return (w >> pc.sh) & 0xFFu;
  }
}

// One warp = one slot = one unit at a time; only the first N lanes stay alive (lane k owns
// state k).  The state->state chain is latency bound and a warp issues in order, so throughput
// comes from having many warps (units) resident per SM, not from filling lanes.
template <int N, int MODE>
__global__ void __launch_bounds__(1024)
k_ans_decode(MicUnit* __restrict__ units, const int* __restrict__ list, int nlist, const uint8_t* __restrict__ comp,
             const uint32_t* __restrict__ tabA, uint16_t* __restrict__ states_out, int max_log, int slots_per_cta) {
  extern __shared__ __align__(16) uint8_t smem[];
  constexpr int PW = HALF_WORDS / N;  // ring words each lane carries for the in-flight half
  constexpr unsigned MASK = (N == 32) ? 0xffffffffu : ((1u << N) - 1u);
  const int slot = threadIdx.x >> 5;
  const int k = threadIdx.x & 31;
  if (k >= N) return;
  const PrefixConsts<N> pc(k);

  const size_t tbytes = MODE == 2 ? 0 : ((size_t)(1u << max_log) * (MODE == 0 ? 4 : 2));
  uint8_t* mytab = smem + (size_t)slot * tbytes;
  uint32_t* ring = reinterpret_cast<uint32_t*>(smem + (size_t)slots_per_cta * tbytes) + slot * RING_STRIDE;
  const uint32_t* T32 = reinterpret_cast<const uint32_t*>(mytab);
  const uint16_t* T16 = reinterpret_cast<const uint16_t*>(mytab);

  for (int li = blockIdx.x * slots_per_cta + slot; li < nlist; li += gridDim.x * slots_per_cta) {
    MicUnit* U = &units[list[li]];
    if (U->status != MIC_OK) continue;
    const int L = (int)U->table_log;
    const uint32_t S = 1u << L;
    const uint32_t* A = tabA + U->tab_off;
    __syncwarp(MASK);
    // ---- stage the decode table ------------------------------------------
    if (MODE == 0) {
      uint4* T4 = reinterpret_cast<uint4*>(mytab);
      const uint4* A4 = reinterpret_cast<const uint4*>(A);
      for (uint32_t i = k; i < S / 4; i += N) T4[i] = A4[i];
    } else if (MODE == 1) {
      uint2* T2 = reinterpret_cast<uint2*>(mytab);
      const uint4* A4 = reinterpret_cast<const uint4*>(A);
#pragma unroll 4
      for (uint32_t i = k; i < S / 4; i += N) {
        const uint4 e = A4[i];
        // nextState = (newState + S) >> nbBits  (inverse of fsedecompressu16.go:250-251)
        const uint32_t n0 = ((e.x & 0xFFFF) + S) >> (e.x >> 16), n1 = ((e.y & 0xFFFF) + S) >> (e.y >> 16);
        const uint32_t n2 = ((e.z & 0xFFFF) + S) >> (e.z >> 16), n3 = ((e.w & 0xFFFF) + S) >> (e.w >> 16);
        T2[i] = make_uint2(n0 | (n1 << 16), n2 | (n3 << 16));
      }
    }

    // ---- bitstream geometry ----------------------------------------------
    const uint8_t* bs = comp + U->comp_off + U->bits_off;
    const uint32_t blen = U->bits_len;
    const uintptr_t addr = reinterpret_cast<uintptr_t>(bs);
    const uint32_t* wbase = reinterpret_cast<const uint32_t*>(addr & ~(uintptr_t)63);
    const uint32_t shift = (uint32_t)(addr & 63) * 8;      // first data bit in ring coordinates
    const uint32_t lastb = bs[blen - 1];                   // non-zero (checked by K1)
    int P = (int)(shift + 8u * (blen - 1) + (31u - __clz(lastb | 1u)));  // unread bits are [shift, P); < 2^31 (frames < 256 MB)

    uint32_t pre[PW];
    auto load_half = [&](int hh) {
      if (hh >= 0) {
        const uint32_t* src = wbase + hh * HALF_WORDS + k * PW;
        if (PW == 2) {
          const uint2 v = __ldg(reinterpret_cast<const uint2*>(src));
          pre[0] = v.x; pre[1] = v.y;
        } else {
#pragma unroll
          for (int i = 0; i < PW / 4; i++) {
            const uint4 v = __ldg(reinterpret_cast<const uint4*>(src) + i);
            pre[4 * i] = v.x; pre[4 * i + 1] = v.y; pre[4 * i + 2] = v.z; pre[4 * i + 3] = v.w;
          }
        }
      } else {
#pragma unroll
        for (int i = 0; i < PW; i++) pre[i] = 0;
      }
    };
    auto store_half = [&](int hh) {
#pragma unroll
      for (int i = 0; i < PW; i++) ring[(hh * HALF_WORDS + k * PW + i) & (RING_WORDS - 1)] = pre[i];
      if (k == 0 && (hh & 1) == 0) ring[RING_WORDS] = pre[0];   // mirror of ring[0]
    };
    int cur_half = ((P - 1) >> 5) / HALF_WORDS;
    int cross = cur_half * (HALF_WORDS * 32);   // P <= cross  <=>  the top unread bit left half cur_half
    load_half(cur_half); store_half(cur_half);
    load_half(cur_half - 1); store_half(cur_half - 1);
    load_half(cur_half - 2);
    __syncwarp(MASK);

    const uint8_t* ringb = reinterpret_cast<const uint8_t*>(ring);
    // bits [lo, lo+nb) of the stream; nb == 0 yields 0
    auto extract = [&](int lo, uint32_t nb) -> uint32_t {
      const uint32_t* w = reinterpret_cast<const uint32_t*>(ringb + (((uint32_t)lo >> 3) & 0x7Cu));
      return __funnelshift_r(w[0], w[1], (uint32_t)lo & 31u) & ((1u << nb) - 1u);
    };
    // consume `tot` bits (warp-uniform)
    auto advance = [&](uint32_t tot) {
      P -= (int)tot;
      if (P <= cross) {
        // half cur_half is dead: overwrite it with half cur_half-2 (already in registers)
        store_half(cur_half - 2);
        cur_half -= 1;
        cross -= HALF_WORDS * 32;
        load_half(cur_half - 2);
        if (k == 0 && cur_half >= 6)
          asm volatile("prefetch.global.L2 [%0];" ::"l"(wbase + (cur_half - 6) * HALF_WORDS));
        __syncwarp(MASK);
      }
    };

    int err = 0;
    uint32_t state = 0;
    // ---- initial states: A first, tableLog bits each (fse8state.go:239-250) ----
    {
      const uint32_t tot = (uint32_t)(N * L);
      if ((uint32_t)P - shift < tot) {
        err = 1;
      } else {
        state = extract(P - (k + 1) * L, (uint32_t)L);
        advance(tot);
      }
    }
    uint16_t* out = states_out + U->sym_off;
    uint32_t nsym = 0;

    if (N == 1) {
      // 1-state: no symbol count; stop on bit exhaustion (fsedecompressu16.go:361-376)
      const uint32_t cap = U->sym_cap;
      while (!err) {
        uint32_t nb, ns;
        if (MODE == 0) { const uint32_t e = T32[state]; nb = e >> 16; ns = e & 0xFFFF; }
        else if (MODE == 1) { const uint32_t nx = T16[state]; nb = (uint32_t)L - (31u - __clz(nx | 1u)); ns = (nx << nb) - S; }
        else { const uint32_t e = __ldg(A + state); nb = e >> 16; ns = e & 0xFFFF; }
        if (P == (int)shift && nb > 0) {      // decoderU16.finished()
          if (state != 0) {
            if (nsym >= cap) { err = 2; break; }
            out[nsym++] = (uint16_t)state;  // final()
          }
          break;
        }
        if (nsym >= cap) { err = 2; break; }
        out[nsym++] = (uint16_t)state;
        if ((uint32_t)P - shift < nb) { err = 1; break; }   // partial over-read -> io.ErrUnexpectedEOF
        const uint32_t bits = extract(P - (int)nb, nb);
        state = ns + bits;
        advance(nb);
      }
    } else {
      const uint32_t count = U->count;
      if (count > U->sym_cap) err = 2;
      uint16_t* op = out + k;
      // One round = N symbols.  The loop body carries no error branch: an over-read only makes P run
      // below `shift`, which is remembered in `under` and reported after the loop (reads stay inside the ring).
      int under = 0;
      const uint32_t full = err ? 0u : count / N;
      // decode step without the ring check
      auto round_nocheck = [&]() {
        uint32_t nb, ns;
        if (MODE == 0) { const uint32_t e = T32[state]; nb = e >> 16; ns = e & 0xFFFF; }
        else if (MODE == 1) { const uint32_t nx = T16[state]; nb = (uint32_t)L - (31u - __clz(nx | 1u)); ns = (nx << nb) - S; }
        else { const uint32_t e = __ldg(A + state); nb = e >> 16; ns = e & 0xFFFF; }
        uint32_t tot;
        const uint32_t before = slot_prefix<N>(nb, pc, &tot);
        const uint32_t bits = extract(P - (int)before - (int)nb, nb);
        *op = (uint16_t)state;
        op += N;
        state = ns + bits;
        P -= (int)tot;
        under |= P - (int)shift;          // sign bit set once P < shift
      };
      // The ring is checked once per two rounds: two rounds consume at most 2*N*16 = 256 bits = 8 words, so a top that
      // just left half h still only reaches words of half h-1 before the next check.
      uint32_t r = 0;
      for (; r + 2 <= full; r += 2) {
        round_nocheck();
        round_nocheck();
        advance(0);
      }
      if (r < full) {
        round_nocheck();
        advance(0);
      }
      const uint32_t tail = err ? 0u : count - full * N;
      if (tail) {
        const bool active = (uint32_t)k < tail;
        uint32_t nb, ns;
        if (MODE == 0) { const uint32_t e = T32[state]; nb = e >> 16; ns = e & 0xFFFF; }
        else if (MODE == 1) { const uint32_t nx = T16[state]; nb = (uint32_t)L - (31u - __clz(nx | 1u)); ns = (nx << nb) - S; }
        else { const uint32_t e = __ldg(A + state); nb = e >> 16; ns = e & 0xFFFF; }
        (void)ns;
        if (!active) nb = 0;
        uint32_t tot;
        slot_prefix<N>(nb, pc, &tot);
        if (active) *op = (uint16_t)state;
        P -= (int)tot;
        under |= P - (int)shift;
      }
      if (under < 0 && !err) err = 1;
      nsym = count;
    }
    if (k == 0) {
      U->nsym = nsym;
      if (err) U->status = err == 1 ? MIC_E_BITSTREAM : MIC_E_SIZE;
    }
  }
}

size_t ans_decode_smem_bytes(int max_log, int smem_mode, int slots_per_cta) {
  size_t t = smem_mode == 2 ? 0 : ((size_t)(1u << max_log) * (smem_mode == 0 ? 4 : 2));
  return (size_t)slots_per_cta * (t + RING_STRIDE * 4);
}

template <int N, int MODE>
static void launch_one(MicUnit* d_units, const int* d_list, int nlist, const uint8_t* d_comp, const uint32_t* d_tabA,
                       uint16_t* d_states, int max_log, int slots, int grid, cudaStream_t st) {
  size_t smem = ans_decode_smem_bytes(max_log, MODE, slots);
  cudaFuncSetAttribute(k_ans_decode<N, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  k_ans_decode<N, MODE><<<grid, 32 * slots, smem, st>>>(d_units, d_list, nlist, d_comp, d_tabA, d_states, max_log, slots);
}

template <int N>
static void launch_n(MicUnit* d_units, const int* d_list, int nlist, const uint8_t* d_comp, const uint32_t* d_tabA,
                     uint16_t* d_states, int max_log, int mode, int slots, int grid, cudaStream_t st) {
  if (mode == 0) launch_one<N, 0>(d_units, d_list, nlist, d_comp, d_tabA, d_states, max_log, slots, grid, st);
  else if (mode == 1) launch_one<N, 1>(d_units, d_list, nlist, d_comp, d_tabA, d_states, max_log, slots, grid, st);
  else launch_one<N, 2>(d_units, d_list, nlist, d_comp, d_tabA, d_states, max_log, slots, grid, st);
}

void launch_ans_decode(MicUnit* d_units, const int* d_list, int nlist, int nstates, const uint8_t* d_comp,
                       const uint32_t* d_tabA, uint16_t* d_states, int max_log, int smem_mode, int slots_per_cta,
                       int grid, cudaStream_t st) {
  if (nlist <= 0) return;
  switch (nstates) {
    case 1: launch_n<1>(d_units, d_list, nlist, d_comp, d_tabA, d_states, max_log, smem_mode, slots_per_cta, grid, st); break;
    case 2: launch_n<2>(d_units, d_list, nlist, d_comp, d_tabA, d_states, max_log, smem_mode, slots_per_cta, grid, st); break;
    case 4: launch_n<4>(d_units, d_list, nlist, d_comp, d_tabA, d_states, max_log, smem_mode, slots_per_cta, grid, st); break;
    default: launch_n<8>(d_units, d_list, nlist, d_comp, d_tabA, d_states, max_log, smem_mode, slots_per_cta, grid, st); break;
  }
}

}  // namespace micgpu
