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
#include <cstdlib>
#include <type_traits>

#include <algorithm>
#include <cstdlib>

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


// ---- packed variant: 32/N units per warp ----------------------------------------------------------------------
// The one-unit-per-warp kernel above issues every instruction for N of 32 lanes.  With all 2048 strips of the
// headline batch resident (14 per SM, the shared-memory limit) its round costs 194 cycles of pure dependency chain
// but 264 cycles measured: 3.5 warps per scheduler x 43 instructions per round queue behind each other
// (tools/ubench_lat.cu, profiles/README.md).  Here lanes [sub*N, sub*N+N) of a warp own unit `sub`, so the same
// instruction stream advances 32/N units at once and the schedulers stop being a bottleneck.
// Cross-lane prefix: partial-mask redux would be serialised by the compiler, and 2 redux per unit do not scale to
// four units; instead every lane posts its nbBits as one byte in shared memory, reads its unit's N bytes back as
// one word (pair) and sums the masked bytes with dp4a -- same latency as the redux path (~50 cycles), 8
// instructions for any number of units per warp.
template <int N>
struct ByteMasks {
  uint32_t lo, hi;      // 0x01 in the bytes of lanes j < k of the unit's low / high word (dp4a multipliers)
  __device__ explicit ByteMasks(int k) {
    lo = 0; hi = 0;
#pragma unroll
    for (int f = 0; f < 8; f++) {
      if (f < k && f < N) {
        if (f < 4) lo |= 0x01u << (8 * f);
        else hi |= 0x01u << (8 * (f - 4));
      }
    }
  }
};

template <int N>
__device__ __forceinline__ uint32_t packed_prefix(uint32_t nb, uint8_t* my_byte, const uint8_t* unit_bytes,
                                                  const ByteMasks<N>& bm, uint32_t* tot) {
  *my_byte = (uint8_t)nb;
  __syncwarp();
  constexpr uint32_t ONES = 0x01010101u;
  if (N == 8) {
    const uint2 v = *reinterpret_cast<const uint2*>(unit_bytes);
    *tot = __dp4a(v.x, ONES, __dp4a(v.y, ONES, 0u));
    return __dp4a(v.x, bm.lo, __dp4a(v.y, bm.hi, 0u));   // the per-lane multiplier selects the lanes below this one
  } else if (N == 4) {
    const uint32_t v = *reinterpret_cast<const uint32_t*>(unit_bytes);
    *tot = __dp4a(v, ONES, 0u);
    return __dp4a(v, bm.lo, 0u);
  } else {
    const uint32_t v = *reinterpret_cast<const uint16_t*>(unit_bytes);
    *tot = (v & 0xFFu) + (v >> 8);
    return (v & 0xFFu) * bm.lo;
  }
}

// Ring maintenance of the packed kernel: straight-line, predicated, and asynchronous.
// The 128 B ring is four 32 B quarters; stream quarter q lives in slot q & 3.  Reads of one loop iteration (two
// rounds, at most C = 256 bits) touch [P - C, P).  When the top unread bit leaves quarter qtop, that slot receives
// quarter qtop - 4 through cp.async (global -> shared, no register and therefore no scoreboard wait in the loop: with
// four units per warp an LDG result consumed one iteration later stalled the whole warp ~140 cycles per refill,
// profiles/README.md).  A copy issued at check i is complete for every lane after check i + 2 (wait_group 2 +
// __syncwarp).  Proof of coverage: after check j the ring (landed or in flight) reaches below P_j + 256 - 1024; the
// rounds after check i may touch [P_i - C, P_i) and need what was issued up to check i - 2, which reaches below
// P_(i-2) - 768 <= P_i + 2C - 768 <= P_i - C because 3C <= 768.  Quarters below the start of the stream are never
// copied: only an over-reading (corrupt) stream looks there; it reads stale ring bytes and is rejected by P < shift.
template <int N>
__device__ __forceinline__ void ring_refill_async(int P, int& cross, uint32_t& ra, uint32_t& qo, int& hidx, const uint8_t*& srcp,
                                                  uint32_t rk, uint32_t mir, int pf_min) {
  constexpr int CP = 32 / N;   // bytes of a quarter each lane copies
  asm volatile(
      "{\n\t.reg .pred p, q, pl, pp;\n\t"
      "setp.le.s32 p, %5, %0;\n\t"
      "setp.ge.and.s32 pl, %3, 0, p;\n\t"
      "setp.eq.and.u32 q, %2, 0, pl;\n\t"            // slot 0: ring words 0..3 are mirrored at words 32..35
      "setp.ne.and.u32 q, %7, 0, q;\n\t"
      "@pl cp.async.ca.shared.global [%1], [%4], %9;\n\t"
      "@q cp.async.ca.shared.global [%1+128], [%4], %9;\n\t"
      "cp.async.commit_group;\n\t"
      "cp.async.wait_group 2;\n\t"
      "setp.ge.and.s32 pp, %3, %8, p;\n\t"
      "@pp prefetch.global.L2 [%4+-256];\n\t"
      "@p add.s32 %0, %0, -256;\n\t"
      "@p add.s32 %3, %3, -1;\n\t"
      "@p add.s64 %4, %4, -32;\n\t"
      "@p add.s32 %2, %2, -32;\n\t"
      "@p and.b32 %2, %2, 96;\n\t"
      "add.s32 %1, %6, %2;\n\t"
      "}\n"
      : "+r"(cross), "+r"(ra), "+r"(qo), "+r"(hidx), "+l"(srcp)
      : "r"(P), "r"(rk), "r"(mir), "r"(pf_min), "n"(CP)
      : "memory");
}

template <int N, int MODE>
__global__ void __launch_bounds__(1024)
k_ans_decode_packed(MicUnit* __restrict__ units, const int* __restrict__ list, int nlist, const uint8_t* __restrict__ comp,
                    const uint32_t* __restrict__ tabA, uint16_t* __restrict__ states_out, int max_log, int slots) {
  static_assert(N == 2 || N == 4 || N == 8, "packed decode is for the interleaved coders");
  static_assert(MODE == 0 || MODE == 1, "packed decode keeps its tables in shared memory");
  extern __shared__ __align__(16) uint8_t smem[];
  constexpr int CP = 32 / N;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  const int k = lane % N, sub = lane / N;
  const int slot = sub * nwarps + warp;          // slots are dealt round-robin over the warps
  const bool slot_ok = slot < slots;
  const int sslot = slot_ok ? slot : 0;          // lanes without a slot alias slot 0's ring (read-only)

  const ByteMasks<N> bm(k);
  // tableLog 16 in 2-byte cells: nextState needs 17 bits; the top one lives in a bit array behind the cells
  const bool l16 = MODE == 1 && max_log == 16;
  const size_t tbytes = (size_t)(1u << max_log) * (MODE == 0 ? 4 : 2) + (l16 ? (1u << 16) / 8 : 0);
  uint8_t* mytab = smem + (size_t)sslot * tbytes;
  uint32_t* myflags = reinterpret_cast<uint32_t*>(mytab + ((size_t)2 << max_log));
  uint32_t* ring = reinterpret_cast<uint32_t*>(smem + (size_t)slots * tbytes) + sslot * RING_STRIDE;
  uint8_t* xbase = smem + (size_t)slots * (tbytes + RING_STRIDE * 4);
  // Idle cell: a one-entry table for lanes that own no unit.  With L = 5 the entry decodes to nbBits = 0 and
  // newState = 0, so such lanes run the same instruction stream without consuming bits or leaving the cell.
  uint32_t* idle = reinterpret_cast<uint32_t*>(xbase);
  uint8_t* xch = xbase + 16 + warp * 64;         // byte exchange: 2 x 32 B per warp
  uint8_t* xmine = xch + lane;
  const uint8_t* xunit = xch + sub * N;
  const uint8_t* ringb = reinterpret_cast<const uint8_t*>(ring);
  const uint32_t ringsa = (uint32_t)__cvta_generic_to_shared(ring);
  const uint32_t rk = ringsa + (uint32_t)(k * CP);
  const int pf_min = k == 0 ? 8 : 0x7fffffff;    // lane 0 of a unit prefetches 8 quarters (256 B) ahead into L2
  const uint32_t mir = k * CP < 16 ? 1u : 0u;    // this lane's bytes of quarter 0 belong to ring words 0..3 (mirrored)
  if (threadIdx.x == 0) { idle[0] = MODE == 0 ? 0u : 32u; idle[1] = 0u; idle[2] = 0u; }   // [1]: idle cell of the direct format, [2]: idle flag word
  __syncthreads();

  for (int base = blockIdx.x * slots; base < nlist; base += gridDim.x * slots) {
    const int li = base + slot;
    MicUnit* U = nullptr;
    bool has = false;
    if (slot_ok && li < nlist) {
      U = &units[list[li]];
      has = U->status == MIC_OK;
    }
    __syncwarp();
    int L = 5;
    uint32_t S = 32, shift = 0, count = 0;
    int P = 0, cross = -(1 << 30);
    const uint8_t* srcp = nullptr;    // this lane's bytes of the next quarter to copy (index hidx)
    int hidx = -1;
    uint32_t qo = 0, ra = rk;

    auto extract = [&](int lo, uint32_t nb) -> uint32_t {
      const uint32_t* w = reinterpret_cast<const uint32_t*>(ringb + (((uint32_t)lo >> 3) & 0x7Cu));
      return __funnelshift_r(w[0], w[1], (uint32_t)lo & 31u) & ((1u << nb) - 1u);
    };

    int err = 0;
    uint32_t full = 0;
    bool d16_ok = true;
    if (has) {
      L = (int)U->table_log;
      S = 1u << L;
      const uint8_t* bs = comp + U->comp_off + U->bits_off;
      const uint32_t blen = U->bits_len;
      const uintptr_t addr = reinterpret_cast<uintptr_t>(bs);
      const uint8_t* wbase = reinterpret_cast<const uint8_t*>(addr & ~(uintptr_t)31);
      shift = (uint32_t)(addr & 31) * 8;         // first data bit in ring coordinates
      const uint32_t lastb = bs[blen - 1];       // non-zero (checked by K1)
      P = (int)(shift + 8u * (blen - 1) + (31u - __clz(lastb | 1u)));   // unread bits are [shift, P)
      const int qtop = (P - 1) >> 8;
      cross = qtop << 8;                         // P <= cross  <=>  the top unread bit left quarter qtop
      // quarters qtop .. qtop-3 fill the ring
#pragma unroll
      for (int j = 0; j < 4; j++) {
        const int q = qtop - j;
        if (q >= 0) {
          const uint32_t dst = rk + (uint32_t)((q & 3) << 5);
          const uint8_t* src = wbase + (size_t)q * 32 + k * CP;
          asm volatile("cp.async.ca.shared.global [%0], [%1], %2;" ::"r"(dst), "l"(src), "n"(CP) : "memory");
          if (mir && (q & 3) == 0)
            asm volatile("cp.async.ca.shared.global [%0], [%1], %2;" ::"r"(dst + 128u), "l"(src), "n"(CP) : "memory");
        }
      }
      hidx = qtop - 4;
      srcp = wbase + (ptrdiff_t)hidx * 32 + k * CP;
      qo = (uint32_t)((hidx & 3) << 5);
      ra = rk + qo;
      count = U->count;
      if (count > U->sym_cap) err = 2;
      // ---- stage the decode table while the ring copies fly -----------------------------
      const uint32_t* A = tabA + U->tab_off;
      if (MODE == 0) {
        uint4* T4 = reinterpret_cast<uint4*>(mytab);
        const uint4* A4 = reinterpret_cast<const uint4*>(A);
        for (uint32_t i = k; i < S / 4; i += N) T4[i] = A4[i];
      } else {
        // Direct 16-bit cells: nbBits in the low nibble, newState >> nbBits above it (newState is a multiple of
        // 2^nbBits).  They fit while newState >> nbBits < 4096, i.e. for every table up to tableLog 12 and for tableLog 13
        // unless one symbol holds more than 3/4 of the table; nbBits is then one AND away from the cell instead of a
        // find-leading-one (18 cycles on the state chain, tools/ubench_lat.cu).
        uint2* T2 = reinterpret_cast<uint2*>(mytab);
        const uint4* A4 = reinterpret_cast<const uint4*>(A);
        uint32_t wide = 0;
#pragma unroll 4
        for (uint32_t i = k; i < S / 4; i += N) {
          const uint4 e = A4[i];
          const uint32_t d0 = (e.x & 0xFFFF) >> (e.x >> 16), d1 = (e.y & 0xFFFF) >> (e.y >> 16);
          const uint32_t d2 = (e.z & 0xFFFF) >> (e.z >> 16), d3 = (e.w & 0xFFFF) >> (e.w >> 16);
          wide |= d0 | d1 | d2 | d3;
          T2[i] = make_uint2((d0 << 4) | (e.x >> 16) | (((d1 << 4) | (e.y >> 16)) << 16), (d2 << 4) | (e.z >> 16) | (((d3 << 4) | (e.w >> 16)) << 16));
        }
        d16_ok = wide < 4096u && L <= 15;
        if (l16)
          for (uint32_t j = k; j < (1u << 16) / 32; j += N) myflags[j] = 0;
      }
    }
    // one cell format per warp: the direct one if every unit of the warp allows it, else nextState cells for all of them
    const bool d16 = MODE == 1 && __all_sync(0xffffffffu, !has || d16_ok);
    __syncwarp();   // the first staging pass (and the zeroed flag words) is visible to every lane of the unit
    if (MODE == 1 && !d16 && has) {
      uint2* T2 = reinterpret_cast<uint2*>(mytab);
      const uint4* A4 = reinterpret_cast<const uint4*>(tabA + U->tab_off);
#pragma unroll 4
      for (uint32_t i = k; i < S / 4; i += N) {
        const uint4 e = A4[i];
        // nextState = (newState + S) >> nbBits  (inverse of fsedecompressu16.go:250-251)
        const uint32_t n0 = ((e.x & 0xFFFF) + S) >> (e.x >> 16), n1 = ((e.y & 0xFFFF) + S) >> (e.y >> 16);
        const uint32_t n2 = ((e.z & 0xFFFF) + S) >> (e.z >> 16), n3 = ((e.w & 0xFFFF) + S) >> (e.w >> 16);
        T2[i] = make_uint2((n0 & 0xFFFF) | (n1 << 16), (n2 & 0xFFFF) | (n3 << 16));
        if (l16) {
          const uint32_t hi = (n0 >> 16) | ((n1 >> 16) << 1) | ((n2 >> 16) << 2) | ((n3 >> 16) << 3);
          if (hi) atomicOr(&myflags[i >> 3], hi << ((i & 7u) * 4));
        }
      }
    }
    asm volatile("cp.async.commit_group;\n\tcp.async.wait_group 0;" ::: "memory");
    __syncwarp();
    // Per-unit ring maintenance; P and cross are uniform over a unit's lanes.
    auto refill = [&]() {
      ring_refill_async<N>(P, cross, ra, qo, hidx, srcp, rk, mir, pf_min);
      __syncwarp();
    };

    uint32_t state = 0;
    if (has && !err) {
      const uint32_t tot = (uint32_t)(N * L);
      if ((uint32_t)P - shift < tot) {
        err = 1;
      } else {
        state = extract(P - (k + 1) * L, (uint32_t)L);
        P -= (int)tot;
      }
    }
    refill();
    const bool live = has && !err;
    if (!live) { L = 5; S = 32; state = 0; }
    // lanes without a live unit decode the idle cell
    const uint32_t* T32 = live ? reinterpret_cast<const uint32_t*>(mytab) : idle;
    const uint16_t* T16 = live ? reinterpret_cast<const uint16_t*>(mytab) : reinterpret_cast<const uint16_t*>(idle + (d16 ? 1 : 0));
    const uint32_t* FL = live ? myflags : idle + 2;
    // one table cell -> (nbBits, newState); the format is uniform over the warp
    auto cell = [&](auto fmt, uint32_t st, uint32_t& nb, uint32_t& ns) {
      constexpr int F = decltype(fmt)::value;   // 0: u32 cells, 1: nextState cells, 2: direct 16-bit cells
      if (F == 0) { const uint32_t e = T32[st]; nb = e >> 16; ns = e & 0xFFFF; }
      else if (F == 1) {
        uint32_t nx = T16[st];
        if (l16) nx |= ((FL[st >> 5] >> (st & 31u)) & 1u) << 16;
        nb = (uint32_t)L - (31u - __clz(nx));    // nx >= 1 (K1)
        ns = (nx << nb) - S;
      }
      else { const uint32_t e = T16[st]; nb = e & 15u; ns = (e >> 4) << nb; }
    };
    auto cell_rt = [&](uint32_t st, uint32_t& nb, uint32_t& ns) {
      if (MODE == 0) cell(std::integral_constant<int, 0>{}, st, nb, ns);
      else if (d16) cell(std::integral_constant<int, 2>{}, st, nb, ns);
      else cell(std::integral_constant<int, 1>{}, st, nb, ns);
    };
    uint16_t* op = states_out + (has ? U->sym_off : 0) + k;
    const uint32_t full0 = live ? count / N : 0u;
    full = full0;
    const uint32_t maxfull = __reduce_max_sync(0xffffffffu, full);
    // The unguarded loop lets an over-reading (corrupt) unit run on to minfull: its reads stay inside the ring and the
    // table, its writes inside its own sym_cap, and P falls by at most 16 bits per symbol, so it cannot wrap below.
    uint32_t minfull = min(maxfull, __reduce_min_sync(0xffffffffu, live ? full : 0xffffffffu));
    if (__any_sync(0xffffffffu, count > (1u << 26))) minfull = 0;

    // ALL: every live unit of the warp is inside its stream (no per-round activity test)
    auto round = [&](auto all_tag, bool act, int buf, int oidx) {
      constexpr bool ALL = decltype(all_tag)::value;
      uint32_t nb, ns;
      cell_rt(state, nb, ns);
      if (!ALL && !act) nb = 0;
      uint32_t tot;
      const uint32_t before = packed_prefix<N>(nb, xmine + buf * 32, xunit + buf * 32, bm, &tot);
      const uint32_t bits = extract(P - (int)before - (int)nb, nb);
      if (ALL) {
        if (live) op[oidx * N] = (uint16_t)state;
        state = ns + bits;
      } else if (act) {
        op[oidx * N] = (uint16_t)state;
        state = ns + bits;
      }
      P -= (int)tot;
    };
    // Hot loop: the five ring words that can hold this round's fields ([P - 128, P) plus word alignment) are loaded
    // into registers as soon as P is known, i.e. before the table lookup of the round, so the bit extraction after
    // the prefix exchange is a register select + funnel shift instead of an address computation and a shared-memory
    // load on the state -> state chain (-24 cycles per round, tools/ubench_lat.cu).
    uint32_t W0 = 0, W1 = 0, W2 = 0, W3 = 0, W4 = 0, Pb = 0;
    auto load_window = [&]() {
      const uint32_t wl = (((uint32_t)(P - 1)) >> 5) - 4u;     // lowest word of the window
      const uint32_t* w = reinterpret_cast<const uint32_t*>(ringb + ((wl & 31u) << 2));
      W0 = w[0]; W1 = w[1]; W2 = w[2]; W3 = w[3]; W4 = w[4];    // words 32..35 mirror 0..3: no wrap inside the window
      Pb = (uint32_t)P - (wl << 5);                            // P relative to the window, in (128, 160]
    };
    auto round_win = [&](auto fmt, int buf, int oidx) {
      uint32_t nb, ns;
      cell(fmt, state, nb, ns);
      uint32_t tot;
      const uint32_t before = packed_prefix<N>(nb, xmine + buf * 32, xunit + buf * 32, bm, &tot);
      const uint32_t rel = Pb - before - nb;                   // bit offset of the field inside the window
      const bool j0 = rel & 32u, j1 = rel & 64u, j2 = rel & 128u;
      const uint32_t a0 = j0 ? W1 : W0, a1 = j0 ? W3 : W2, b0 = j0 ? W2 : W1, b1 = j0 ? W4 : W3;
      const uint32_t a = j1 ? a1 : a0, hi = j1 ? b1 : b0;
      const uint32_t lw = j2 ? W4 : a;                         // word 4 holds its field entirely: hi is irrelevant there
      const uint32_t bits = __funnelshift_r(lw, hi, rel) & ((1u << nb) - 1u);
      if (live) op[oidx * N] = (uint16_t)state;
      state = ns + bits;
      P -= (int)tot;
    };
    // ring check once per two rounds (two rounds consume at most 2*N*16 = 256 bits = one quarter)
    uint32_t r = 0;
    {
      auto hot = [&](auto fmt) {
        for (; r + 2 <= minfull; r += 2) {
          load_window();          // after the ring check of the previous iteration (its copies are then guaranteed)
          round_win(fmt, 0, 0);
          load_window();
          round_win(fmt, 1, 1);
          op += 2 * N;
          refill();
        }
      };
      if (MODE == 0) hot(std::integral_constant<int, 0>{});
      else if (d16) hot(std::integral_constant<int, 2>{});
      else hot(std::integral_constant<int, 1>{});
    }
    if (live && P < (int)shift) full = 0;
    for (; r + 2 <= maxfull; r += 2) {
      const bool a0 = r < full, a1 = r + 1 < full;
      round(std::false_type{}, a0, 0, 0);
      round(std::false_type{}, a1, 1, 1);
      op += (a0 ? N : 0) + (a1 ? N : 0);
      refill();
      if (P < (int)shift) full = 0;   // over-read: stop this unit (reads stayed inside the ring)
    }
    if (r < maxfull) {
      const bool a0 = r < full;
      round(std::false_type{}, a0, 0, 0);
      op += a0 ? N : 0;
      refill();
      if (P < (int)shift) full = 0;
    }
    {
      const uint32_t tail = (live && full == full0) ? count - full0 * N : 0u;
      const bool act = (uint32_t)k < tail;
      uint32_t nb, ns_unused;
      cell_rt(state, nb, ns_unused);
      if (!act) nb = 0;
      uint32_t tot;
      packed_prefix<N>(nb, xmine + 32, xunit + 32, bm, &tot);
      if (act) *op = (uint16_t)state;
      P -= (int)tot;
      if (live && (P < (int)shift || full != full0)) err = 1;
    }
    asm volatile("cp.async.wait_group 0;" ::: "memory");   // no copy may land in the ring after the next unit took it over
    if (has && k == 0) {
      U->nsym = count;
      if (err) U->status = err == 1 ? MIC_E_BITSTREAM : MIC_E_SIZE;
    }
  }
}

size_t ans_decode_smem_bytes(int max_log, int smem_mode, int slots_per_cta) {
  size_t t = smem_mode == 2 ? 0 : ((size_t)(1u << max_log) * (smem_mode == 0 ? 4 : 2));
  if (smem_mode == 1 && max_log == 16) t += (1u << 16) / 8;   // bit 16 of nextState, one bit per cell (packed kernel only)
  // + the packed kernel's idle cell (16 B) and byte-exchange area (64 B per warp; at least 4 slots share a warp)
  return (size_t)slots_per_cta * (t + RING_STRIDE * 4) + 16 + 64 * (size_t)((slots_per_cta + 3) / 4);
}

template <int N, int MODE>
static void launch_one(MicUnit* d_units, const int* d_list, int nlist, const uint8_t* d_comp, const uint32_t* d_tabA,
                       uint16_t* d_states, int max_log, int slots, int grid, cudaStream_t st) {
  size_t smem = ans_decode_smem_bytes(max_log, MODE, slots);
  cudaFuncSetAttribute(k_ans_decode<N, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  k_ans_decode<N, MODE><<<grid, 32 * slots, smem, st>>>(d_units, d_list, nlist, d_comp, d_tabA, d_states, max_log, slots);
}

template <int N, int MODE>
static void launch_packed(MicUnit* d_units, const int* d_list, int nlist, const uint8_t* d_comp, const uint32_t* d_tabA,
                          uint16_t* d_states, int max_log, int slots, int grid, cudaStream_t st) {
  constexpr int UPW = 32 / N;
  int warps = (slots + UPW - 1) / UPW;
  // MICGPU_K2_UPW: units per warp (default 32 / N).  Slots are dealt round-robin over the warps, so more warps simply
  // leave the upper lanes of each on the idle cell: fewer lanes per table lookup (fewer bank conflicts) against more
  // warps per scheduler.
  static const int upw_cfg = [] { const char* e = getenv("MICGPU_K2_UPW"); return e ? atoi(e) : 0; }();
  if (upw_cfg >= 1 && upw_cfg < UPW) warps = std::min(32, (slots + upw_cfg - 1) / upw_cfg);
  size_t smem = ans_decode_smem_bytes(max_log, MODE, slots);
  const int base_warps = (slots + 3) / 4;
  if (warps > base_warps) smem += 64 * (size_t)(warps - base_warps);   // byte-exchange area: 64 B per warp
  cudaFuncSetAttribute(k_ans_decode_packed<N, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  k_ans_decode_packed<N, MODE><<<grid, 32 * warps, smem, st>>>(d_units, d_list, nlist, d_comp, d_tabA, d_states, max_log, slots);
}

static bool use_packed() {
  static const bool v = [] { const char* e = getenv("MICGPU_K2_ONE_UNIT_PER_WARP"); return !(e && e[0] == '1'); }();
  return v;
}

template <int N>
static void launch_n(MicUnit* d_units, const int* d_list, int nlist, const uint8_t* d_comp, const uint32_t* d_tabA,
                     uint16_t* d_states, int max_log, int mode, int slots, int grid, cudaStream_t st) {
  if constexpr (N > 1) {
    if (mode != 2 && use_packed()) {
      if (mode == 0) launch_packed<N, 0>(d_units, d_list, nlist, d_comp, d_tabA, d_states, max_log, slots, grid, st);
      else launch_packed<N, 1>(d_units, d_list, nlist, d_comp, d_tabA, d_states, max_log, slots, grid, st);
      return;
    }
  }
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
