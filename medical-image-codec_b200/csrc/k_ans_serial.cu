// k_ans_serial.cu -- K2s: tANS decode of 1-, 2-, 4- and 8-state streams with ONE THREAD PER UNIT.
//
// Replaces decompress (fsedecompressu16.go:267-377), decompress2State (fse2state.go:203-308) and
// decompress4State (fse4state.go:195-353), and decompress8State (fse8state.go:230-380) for the streams the Go API emits by default:
// CompressParallelStrips, CompressSingleFrame (MIC2 frames, MIC3 tile planes) and compressResidualFrame all
// try the two-state coder first and fall back to the single-state one (multiframecompress.go:15-35,145-162).
//
// Why not lanes = states as in k_ans.cu: a round of the lane-parallel kernel costs ~170 cycles whatever N is
// (table lookup -> nbBits -> cross-lane prefix through shared memory -> window select -> next state), so a 2-state
// stream advances 2 symbols per 170 cycles.  With the N states of a unit in the registers of one thread the
// cross-lane exchange disappears: the bit position chains through N integer subtractions, the N table lookups of a
// round are issued back to back, and the fields of a round (<= 32 bits for N <= 2, <= 64 for N = 4) come out of a
// two- or three-register window with one select and one funnel shift each.  The chain per round is one
// shared-memory load plus ~8 dependent ALU instructions (~70 cycles for N <= 2).  Throughput is (resident units) /
// (round latency), and residency is bounded by shared memory (the decode tables), exactly as for k_ans.cu.
//
// Everything else follows k_ans.cu: tables in shared memory as 4-byte cells (newState | nbBits << 16) or 2-byte
// cells (direct: nbBits | (newState >> nbBits) << 4; or nextState, with a bit array for bit 16 when tableLog is
// 16), chosen per CTA; the bitstream of a unit in a 128 B shared-memory ring of four 32 B quarters refilled with
// cp.async; output = the state stream (table index of every symbol), K3 maps it to symbols through tabS.
// Bit coordinates as in k_ans.cu:26-29 (bitreader.go:26-62): a field of n bits read when P bits remain unread is
// bits [P-n, P) of the stream viewed as one little-endian integer.
#include <type_traits>

#include "mic_device.cuh"

namespace micgpu {

namespace {

constexpr int SERIAL_THREADS = 128;
// Ring of RQ quarters of 32 B (+ mirror of words 0..3).  Four quarters cover every table up to tableLog 13; wider tables
// take eight.  Why: the register bit buffer reads up to 96 bits below the bit position, so the rounds after a ring check
// touch [P - C - 96, P) with C <= 256 bits consumed per check.  With four quarters and the two youngest copies possibly
// still in flight (wait_group 2), what is guaranteed to have landed reaches down to the start of the quarter below the
// one P is in -- enough unless three consecutive checks consume more than 673 bits (48 symbols of > 14 bits: only
// tableLog >= 14 can do that, and near-incompressible 16-bit data does).  With eight quarters the landed part reaches
// 1280 bits below the top quarter.  The extra 128 B do not cost residency: such a table is 32 KB or more.
__host__ __device__ constexpr int serial_ring_quarters(int table_log) { return table_log >= 14 ? 8 : 4; }
__host__ __device__ constexpr int serial_ring_words(int table_log) { return serial_ring_quarters(table_log) * 8 + 4; }

// Branch-free ring maintenance of one thread's unit (see ring_refill_async in k_ans.cu for the coverage proof; the
// cadence is the same: at most 256 bits are consumed between two calls).  When the top unread bit has left quarter
// qtop, the slot of that quarter receives quarter qtop - 4: two 16 B cp.async, plus the mirror copy for ring slot 0.
template <int RQ>
__device__ __forceinline__ void serial_refill(int P, int& cross, uint32_t& ra, uint32_t& qo, int& hidx, const uint8_t*& srcp,
                                              uint32_t ring_sa) {
  asm volatile(
      "{\n\t.reg .pred p, pl, q, pp;\n\t"
      "setp.le.s32 p, %5, %0;\n\t"
      "setp.ge.and.s32 pl, %3, 0, p;\n\t"
      "setp.eq.and.u32 q, %2, 0, pl;\n\t"
      "@pl cp.async.ca.shared.global [%1], [%4], 16;\n\t"
      "@pl cp.async.ca.shared.global [%1+16], [%4+16], 16;\n\t"
      "@q cp.async.ca.shared.global [%1+%7], [%4], 16;\n\t"
      "cp.async.commit_group;\n\t"
      "cp.async.wait_group 2;\n\t"
      "setp.ge.and.s32 pp, %3, 8, p;\n\t"
      "@pp prefetch.global.L2 [%4+-256];\n\t"
      "@p add.s32 %0, %0, -256;\n\t"
      "@p add.s32 %3, %3, -1;\n\t"
      "@p add.s64 %4, %4, -32;\n\t"
      "@p add.s32 %2, %2, -32;\n\t"
      "@p and.b32 %2, %2, %8;\n\t"
      "add.s32 %1, %6, %2;\n\t"
      "}\n"
      : "+r"(cross), "+r"(ra), "+r"(qo), "+r"(hidx), "+l"(srcp)
      : "r"(P), "r"(ring_sa), "n"(RQ * 32), "n"((RQ - 1) * 32)
      : "memory");
}

// FMT: 0 = 4-byte cells, 1 = 2-byte nextState cells (+ bit 16 in a bit array when L16), 2 = 2-byte direct cells,
// 3 = split cells (u16 newState array + u8 nbBits array: 3 bytes per cell)
template <int FMT, bool L16>
struct Cells {
  const uint8_t* tab;
  const uint32_t* flags;
  uint32_t L, S;
  // -> nbBits and h = newState >> nbBits (newState is a multiple of 2^nbBits): the next state is (h << nb) | field, which
  // one funnel shift forms from h and the 32 stream bits that end where the field ends (decode_unit)
  __device__ __forceinline__ void get(uint32_t st, uint32_t& nb, uint32_t& h) const {
    if (FMT == 0) {
      const uint32_t e = reinterpret_cast<const uint32_t*>(tab)[st];
      nb = e >> 16; h = (e & 0xFFFFu) >> nb;
    } else if (FMT == 1) {
      uint32_t nx = reinterpret_cast<const uint16_t*>(tab)[st];
      if (L16) nx |= ((flags[st >> 5] >> (st & 31u)) & 1u) << 16;
      nb = L - (31u - __clz(nx));   // nextState >= 1 (K1)
      h = nx - (S >> nb);           // ((nx << nb) - S) >> nb
    } else if (FMT == 2) {
      const uint32_t e = reinterpret_cast<const uint16_t*>(tab)[st];
      nb = e & 15u; h = e >> 4;
    } else {
      // split cells: h (< 2^16 even for tableLog 16) in a u16 array, nbBits in a u8 array behind it -- two independent
      // loads and nothing to compute, where the 2-byte nextState format needs a flag word, a shift and a
      // find-leading-one on the state chain (157 against 112 cycles per round on tableLog-16 residual frames)
      h = reinterpret_cast<const uint16_t*>(tab)[st];
      nb = (tab + ((size_t)2 << L))[st];
    }
  }
};

template <int N, int FMT, bool L16, int RQ>
__device__ __forceinline__ void decode_unit(MicUnit* U, const uint8_t* __restrict__ comp, uint16_t* __restrict__ states_out,
                                            const uint8_t* mytab, const uint32_t* myflags, uint32_t* ring) {
  const Cells<FMT, L16> cells{mytab, myflags, U->table_log, 1u << U->table_log};
  const int L = (int)U->table_log;
  const uint8_t* bs = comp + U->comp_off + U->bits_off;
  const uint32_t blen = U->bits_len;
  const uintptr_t addr = reinterpret_cast<uintptr_t>(bs);
  const uint8_t* wbase = reinterpret_cast<const uint8_t*>(addr & ~(uintptr_t)31);
  const int shift = (int)(addr & 31) * 8;            // first data bit in ring coordinates
  const uint32_t lastb = bs[blen - 1];               // non-zero (checked by K1)
  int P = shift + 8 * (int)(blen - 1) + (31 - __clz(lastb | 1u));   // unread bits are [shift, P); < 2^31 (frames < 256 MB)
  const uint32_t ring_sa = (uint32_t)__cvta_generic_to_shared(ring);
  const uint8_t* ringb = reinterpret_cast<const uint8_t*>(ring);
  const int qtop = (P - 1) >> 8;
  int cross = qtop << 8;                             // P <= cross  <=>  the top unread bit left quarter qtop
  constexpr uint32_t WMASK = RQ * 8 - 1;   // ring words
#pragma unroll
  for (int j = 0; j < RQ; j++) {
    const int q = qtop - j;
    if (q >= 0) {
      const uint32_t dst = ring_sa + (uint32_t)((q & (RQ - 1)) << 5);
      const uint8_t* src = wbase + (size_t)q * 32;
      asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
      asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(dst + 16u), "l"(src + 16) : "memory");
      if ((q & (RQ - 1)) == 0) asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(dst + (uint32_t)(RQ * 32)), "l"(src) : "memory");
    }
  }
  int hidx = qtop - RQ;
  const uint8_t* srcp = wbase + (ptrdiff_t)hidx * 32;
  uint32_t qo = (uint32_t)((hidx & (RQ - 1)) << 5), ra = ring_sa + qo;
  asm volatile("cp.async.commit_group;\n\tcp.async.wait_group 0;" ::: "memory");

  // bits [lo, lo+nb) straight from the ring (start-up and tails; the hot loops use a register window)
  auto extract = [&](int lo, uint32_t nb) -> uint32_t {
    const uint32_t* w = reinterpret_cast<const uint32_t*>(ringb + (((uint32_t)lo >> 3) & (WMASK * 4u)));
    return __funnelshift_r(w[0], w[1], (uint32_t)lo & 31u) & ((1u << nb) - 1u);
  };
  auto refill = [&]() { serial_refill<RQ>(P, cross, ra, qo, hidx, srcp, ring_sa); };

  int err = 0;
  uint32_t st[N];
#pragma unroll
  for (int k = 0; k < N; k++) st[k] = 0;
  // ---- initial states: A first, tableLog bits each (fsedecompressu16.go:387-391, fse2state.go:216-222) ----
  if (P - shift < N * L) {
    err = 1;
  } else {
#pragma unroll
    for (int k = 0; k < N; k++) st[k] = extract(P - (k + 1) * L, (uint32_t)L);
    P -= N * L;
  }
  refill();
  uint16_t* out = states_out + U->sym_off;
  uint32_t nsym = 0;

  // Bit buffer in registers, top aligned: the next unread stream bit is bit 31 of BH, `cnt` bits of BH:BL are valid
  // (kept in (32, 64] between steps of at most 32 bits), the word below the buffer (WN, stream word `wn`) is loaded a step
  // before it is appended.  A field of nb bits is then simply the top nb bits of BH, so
  //   * the next state is ONE funnel shift away from the table cell: (h : BH) << nb, high half, with h = newState >> nb
  //     (the chain used to run lookup -> nb -> position -> select -> shift -> mask -> add);
  //   * consuming a field is one funnel shift and one shift; no position arithmetic, no word select;
  //   * no shared-memory load sits between the fields of one step and those of the next (reloading a window from the
  //     ring at every round put an address computation and a load on exactly that chain).
  // The bit position is implicit: P = 32 (wn + 1) + cnt; it is materialised for the ring checks only.
  uint32_t BH = 0, BL = 0, WN = 0, cnt = 0, wn = 0;
  const uint32_t* ringw = reinterpret_cast<const uint32_t*>(ringb);
  auto buf_load = [&]() {                       // from P
    const uint32_t wtop = ((uint32_t)(P - 1)) >> 5;
    const uint32_t r = (uint32_t)P - (wtop << 5);                 // valid bits of the top word, 1..32
    const uint32_t wt = ringw[wtop & WMASK], wb = ringw[(wtop - 1u) & WMASK];
    BH = __funnelshift_l(wb, wt, 32u - r);                         // (wt : wb) << (32 - r), high half
    BL = wb << ((32u - r) & 31u);
    if (r == 32u) BL = wb;
    cnt = r + 32u;
    wn = wtop - 2u;
    WN = ringw[wn & WMASK];
  };
  auto buf_sync = [&]() { P = (int)(((wn + 1u) << 5) + cnt); };   // the bit position the buffer stands for
  auto take = [&](uint32_t nb) {                                   // consume nb <= 16 bits
    BH = __funnelshift_l(BL, BH, nb);
    BL <<= nb;
    cnt -= nb;
  };
  auto top_up = [&]() {                                            // after a step of <= 32 bits: append the word below
    if (cnt <= 32u) {                                              // BL holds no valid bit any more (and is zero)
      const unsigned long long add = ((unsigned long long)WN << 32) >> cnt;
      BH |= (uint32_t)(add >> 32);
      BL = (uint32_t)add;
      cnt += 32u;
      wn -= 1u;
    }
    WN = ringw[wn & WMASK];
  };
  auto next_state = [&](uint32_t h, uint32_t nb) -> uint32_t { return __funnelshift_l(BH, h, nb); };

  if (N == 1) {
    // 1-state: no symbol count; the stream ends when the bits do (fsedecompressu16.go:351-376)
    const uint32_t cap = U->sym_cap;
    uint32_t s0 = st[0];
    // 16 symbols (<= 256 bits) per ring check, two per window; cannot reach the end of the bits inside the group
    while (!err && P - shift > 256 && nsym + 16 <= cap) {
      buf_load();
      uint32_t pk[8];                    // states are < 2^16: two per word, the 16 of a check leave as two 16 B stores
#pragma unroll
      for (int j = 0; j < 8; j++) {
        uint32_t nb, h;
        cells.get(s0, nb, h);
        const uint32_t e0 = s0;
        s0 = next_state(h, nb);
        take(nb);
        cells.get(s0, nb, h);
        pk[j] = __byte_perm(e0, s0, 0x5410);
        s0 = next_state(h, nb);
        take(nb);
        top_up();
      }
      *reinterpret_cast<uint4*>(out + nsym) = make_uint4(pk[0], pk[1], pk[2], pk[3]);       // nsym is a multiple of 16 here
      *reinterpret_cast<uint4*>(out + nsym + 8) = make_uint4(pk[4], pk[5], pk[6], pk[7]);
      nsym += 16;
      buf_sync();
      refill();
    }
    while (!err) {
      uint32_t nb, h;
      cells.get(s0, nb, h);
      if (P == shift && nb > 0) {        // decoderU16.finished()
        if (s0 != 0) {
          if (nsym >= cap) { err = 2; break; }
          out[nsym++] = (uint16_t)s0;    // final()
        }
        break;
      }
      if (nsym >= cap) { err = 2; break; }
      out[nsym++] = (uint16_t)s0;
      if ((uint32_t)(P - shift) < nb) { err = 1; break; }   // partial over-read -> io.ErrUnexpectedEOF
      s0 = (h << nb) + extract(P - (int)nb, nb);
      P -= (int)nb;
      refill();
    }
  } else {
    const uint32_t count = U->count;
    if (count > U->sym_cap) err = 2;
    const uint32_t full = err ? 0u : count / N;
    // one round = N symbols in stream order A, B, ...: emit the current states, then move each one
    auto round_win = [&](uint16_t* op) {
      uint32_t nb[N], h[N];
#pragma unroll
      for (int k = 0; k < N; k++) cells.get(st[k], nb[k], h[k]);
      if (N == 2) {
        *reinterpret_cast<uint32_t*>(op) = st[0] | (st[1] << 16);
      } else if (N == 4) {
        *reinterpret_cast<uint2*>(op) = make_uint2(st[0] | (st[1] << 16), st[2] | (st[3] << 16));
      } else {
        // sym_off is a multiple of 16 elements and op advances by whole rounds: 16 B aligned
        *reinterpret_cast<uint4*>(op) = make_uint4(st[0] | (st[1] << 16), st[2] | (st[3] << 16), st[4 % N] | (st[5 % N] << 16),
                                                   st[6 % N] | (st[7 % N] << 16));
      }
#pragma unroll
      for (int k = 0; k < N; k++) {
        st[k] = next_state(h[k], nb[k]);
        take(nb[k]);
        if ((k & 1) == 1 || k == N - 1) top_up();      // at most 32 bits since the last one
      }
    };
    // the same round with the emitted states packed two per word into pw[N / 2]: the hot loop stores the 16 symbols of a
    // ring check as two 16 B words (one PRMT per two symbols and two stores, against shift + or + store per round; a unit
    // alone on its SM -- a MIC2 frame -- runs 15 % faster with it, full warps the same: tools/k2s_variants.sh)
    auto round_pack = [&](uint32_t* pw) {
      uint32_t nb[N], h[N];
#pragma unroll
      for (int k = 0; k < N; k++) cells.get(st[k], nb[k], h[k]);
#pragma unroll
      for (int k = 0; k < N; k += 2) pw[k / 2] = __byte_perm(st[k], st[k + 1], 0x5410);
#pragma unroll
      for (int k = 0; k < N; k++) {
        st[k] = next_state(h[k], nb[k]);
        take(nb[k]);
        if ((k & 1) == 1 || k == N - 1) top_up();
      }
    };
    constexpr uint32_t RPC = 16 / N;    // rounds per ring check: RPC * N * 16 = 256 bits
    uint32_t r = 0;
    uint16_t* op = out;
    buf_load();
    for (; r + RPC <= full; r += RPC) {
      uint32_t pk[8];
#pragma unroll
      for (uint32_t j = 0; j < RPC; j++) round_pack(pk + j * (N / 2));
      // sym_off is a multiple of 16 elements and op advances by 16: both stores are 16 B aligned
      *reinterpret_cast<uint4*>(op) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
      *reinterpret_cast<uint4*>(op + 8) = make_uint4(pk[4], pk[5], pk[6], pk[7]);
      op += RPC * N;
      buf_sync();
      refill();
      if (P < shift) break;             // over-read (corrupt stream): reads stayed inside the ring, writes inside sym_cap
    }
    if (P >= shift) {
      for (; r < full; r++) {
        round_win(op);
        op += N;
        buf_sync();
        refill();
      }
      const uint32_t tail = err ? 0u : count - full * N;
#pragma unroll
      for (int k = 0; k < N; k++) {
        if ((uint32_t)k < tail) {
          uint32_t nb, h;
          cells.get(st[k], nb, h);
          op[k] = (uint16_t)st[k];
          P -= (int)nb;
        }
      }
    }
    if (P < shift && !err) err = 1;
    nsym = count;
  }
  asm volatile("cp.async.wait_group 0;" ::: "memory");   // no copy may land in the ring after the next unit took it over
  U->nsym = nsym;
  if (err) U->status = err == 1 ? MIC_E_BITSTREAM : MIC_E_SIZE;
}

}  // namespace

// mode: 0 = 4-byte cells, 1 = 2-byte cells, 3 = split cells.  The host cuts the unit list into SEGMENTS (first index, count) whose tables
// and rings fit one CTA together; slot_off[i] is the byte offset of list entry i inside its segment's shared memory
// (table, then -- tableLog 16 in 2-byte cells -- the bit array, then the ring).  Tables of different sizes share a
// CTA, so a batch of 4 KB and 8 KB tables (MIC3 luma / chroma planes) fills shared memory instead of paying the
// largest table for every unit.  Slots are dealt round-robin over the warps (slot = lane * nwarps + warp) so that
// every scheduler of the SM gets its share of the units.
template <int N>
__global__ void __launch_bounds__(SERIAL_THREADS)
k_ans_decode_serial(MicUnit* __restrict__ units, const int* __restrict__ list, const int2* __restrict__ segs, int nseg,
                    const uint32_t* __restrict__ slot_off, const uint8_t* __restrict__ comp, const uint32_t* __restrict__ tabA,
                    uint16_t* __restrict__ states_out, int mode) {
  extern __shared__ __align__(16) uint8_t smem[];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = blockDim.x >> 5;
  const int slot = lane * nwarps + warp;

  for (int sg = blockIdx.x; sg < nseg; sg += gridDim.x) {
    __syncthreads();   // the previous segment is done with the tables
    const int base = segs[sg].x, npass = segs[sg].y;
    // ---- stage the tables of this segment, all threads per table --------------------------------------------
    bool wide_any = false, any16 = false;
    for (int s = 0; s < npass; s++) {
      const MicUnit* V = &units[list[base + s]];
      if (V->status != MIC_OK) continue;
      const uint32_t L = V->table_log, S = 1u << L;
      const uint4* A4 = reinterpret_cast<const uint4*>(tabA + V->tab_off);
      uint8_t* T = smem + slot_off[base + s];
      if (mode == 0) {
        uint4* T4 = reinterpret_cast<uint4*>(T);
        for (uint32_t i = tid; i < S / 4; i += blockDim.x) T4[i] = __ldg(A4 + i);
      } else if (mode == 3) {
        uint2* T2 = reinterpret_cast<uint2*>(T);
        uint32_t* T1 = reinterpret_cast<uint32_t*>(T + ((size_t)2 << L));
        for (uint32_t i = tid; i < S / 4; i += blockDim.x) {
          const uint4 e = __ldg(A4 + i);
          const uint32_t h0 = (e.x & 0xFFFFu) >> (e.x >> 16), h1 = (e.y & 0xFFFFu) >> (e.y >> 16);
          const uint32_t h2 = (e.z & 0xFFFFu) >> (e.z >> 16), h3 = (e.w & 0xFFFFu) >> (e.w >> 16);
          T2[i] = make_uint2(h0 | (h1 << 16), h2 | (h3 << 16));
          T1[i] = (e.x >> 16) | ((e.y >> 16) << 8) | ((e.z >> 16) << 16) | ((e.w >> 16) << 24);
        }
      } else {
        // direct cells (k_ans.cu:469-472): valid while newState >> nbBits < 4096 and tableLog <= 15
        uint2* T2 = reinterpret_cast<uint2*>(T);
        uint32_t wide = L > 15 ? 4096u : 0u;
        any16 |= L > 15;
        for (uint32_t i = tid; i < S / 4; i += blockDim.x) {
          const uint4 e = __ldg(A4 + i);
          const uint32_t d0 = (e.x & 0xFFFF) >> (e.x >> 16), d1 = (e.y & 0xFFFF) >> (e.y >> 16);
          const uint32_t d2 = (e.z & 0xFFFF) >> (e.z >> 16), d3 = (e.w & 0xFFFF) >> (e.w >> 16);
          wide |= d0 | d1 | d2 | d3;
          T2[i] = make_uint2(((d0 << 4) | (e.x >> 16)) | (((d1 << 4) | (e.y >> 16)) << 16),
                             ((d2 << 4) | (e.z >> 16)) | (((d3 << 4) | (e.w >> 16)) << 16));
        }
        wide_any |= wide >= 4096u;
      }
    }
    // one cell format per CTA: direct cells if every table of the segment allows them, else nextState cells for all
    const bool d16 = mode == 1 && !__syncthreads_or(wide_any ? 1 : 0);
    const bool l16 = mode == 1 && any16;     // uniform: every thread looked at every unit of the segment
    if (mode == 1 && !d16) {
      for (int s = 0; s < npass; s++) {
        const MicUnit* V = &units[list[base + s]];
        if (V->status != MIC_OK) continue;
        const uint32_t L = V->table_log, S = 1u << L;
        const uint4* A4 = reinterpret_cast<const uint4*>(tabA + V->tab_off);
        uint8_t* T = smem + slot_off[base + s];
        uint2* T2 = reinterpret_cast<uint2*>(T);
        uint32_t* F = reinterpret_cast<uint32_t*>(T + ((size_t)2 << L));
        const bool u16 = L > 15;
        if (u16) {
          for (uint32_t j = tid; j < (1u << 16) / 32; j += blockDim.x) F[j] = 0;
          __syncthreads();
        }
        for (uint32_t i = tid; i < S / 4; i += blockDim.x) {
          const uint4 e = __ldg(A4 + i);
          // nextState = (newState + S) >> nbBits  (inverse of fsedecompressu16.go:250-251)
          const uint32_t n0 = ((e.x & 0xFFFF) + S) >> (e.x >> 16), n1 = ((e.y & 0xFFFF) + S) >> (e.y >> 16);
          const uint32_t n2 = ((e.z & 0xFFFF) + S) >> (e.z >> 16), n3 = ((e.w & 0xFFFF) + S) >> (e.w >> 16);
          T2[i] = make_uint2((n0 & 0xFFFF) | (n1 << 16), (n2 & 0xFFFF) | (n3 << 16));
          if (u16) {
            const uint32_t hi = (n0 >> 16) | ((n1 >> 16) << 1) | ((n2 >> 16) << 2) | ((n3 >> 16) << 3);
            if (hi) atomicOr(&F[i >> 3], hi << ((i & 7u) * 4));
          }
        }
      }
    }
    __syncthreads();
    // ---- one thread per unit ----------------------------------------------------------------------------
    if (slot < npass) {
      MicUnit* U = &units[list[base + slot]];
      if (U->status == MIC_OK) {
        const uint32_t L = U->table_log;
        const uint8_t* mytab = smem + slot_off[base + slot];
        const size_t cells = mode == 3 ? (size_t)3 << L : (size_t)(mode == 0 ? 4 : 2) << L;
        const uint32_t* myflags = reinterpret_cast<const uint32_t*>(mytab + cells);
        const size_t fl = (mode == 1 && L > 15) ? (1u << 16) / 8 : 0;
        uint32_t* ring = reinterpret_cast<uint32_t*>(const_cast<uint8_t*>(mytab) + cells + fl);
        if (L < 14) {   // four ring quarters (see serial_ring_quarters)
          if (mode == 0) decode_unit<N, 0, false, 4>(U, comp, states_out, mytab, myflags, ring);
          else if (mode == 3) decode_unit<N, 3, false, 4>(U, comp, states_out, mytab, myflags, ring);
          else if (d16) decode_unit<N, 2, false, 4>(U, comp, states_out, mytab, myflags, ring);
          else decode_unit<N, 1, false, 4>(U, comp, states_out, mytab, myflags, ring);
        } else {        // eight
          if (mode == 0) decode_unit<N, 0, false, 8>(U, comp, states_out, mytab, myflags, ring);
          else if (mode == 3) decode_unit<N, 3, false, 8>(U, comp, states_out, mytab, myflags, ring);
          else if (d16) decode_unit<N, 2, false, 8>(U, comp, states_out, mytab, myflags, ring);
          else if (l16 && L > 15) decode_unit<N, 1, true, 8>(U, comp, states_out, mytab, myflags, ring);
          else decode_unit<N, 1, false, 8>(U, comp, states_out, mytab, myflags, ring);
        }
      }
    }
  }
}

// shared-memory bytes of one unit: cells, the bit array for bit 16 of nextState (tableLog 16 in 2-byte cells), the ring
size_t ans_serial_unit_bytes(int table_log, int mode) {
  const size_t ring = (size_t)serial_ring_words(table_log) * 4;
  if (mode == 3) return ((size_t)3 << table_log) + ring;
  size_t t = (size_t)(1u << table_log) * (mode == 0 ? 4 : 2);
  if (mode == 1 && table_log == 16) t += (1u << 16) / 8;
  return t + ring;
}

size_t ans_serial_smem_bytes(int max_log, int mode, int slots) { return (size_t)slots * ans_serial_unit_bytes(max_log, mode); }

int ans_serial_threads(int slots) {
  // up to four warps so that each scheduler of the SM drives its own share of the units
  const int warps = slots >= 4 ? 4 : (slots < 1 ? 1 : slots);
  return 32 * warps;
}

void launch_ans_decode_serial(MicUnit* d_units, const int* d_list, const int2* d_segs, int nseg, const uint32_t* d_slot_off, int nstates,
                              const uint8_t* d_comp, const uint32_t* d_tabA, uint16_t* d_states, int mode, int max_slots,
                              size_t smem, int grid, cudaStream_t st) {
  if (nseg <= 0) return;
  const int threads = ans_serial_threads(max_slots);
  auto go = [&](auto kern) {
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    kern<<<grid, threads, smem, st>>>(d_units, d_list, d_segs, nseg, d_slot_off, d_comp, d_tabA, d_states, mode);
  };
  if (nstates == 1) go(k_ans_decode_serial<1>);
  else if (nstates == 2) go(k_ans_decode_serial<2>);
  else if (nstates == 4) go(k_ans_decode_serial<4>);
  else go(k_ans_decode_serial<8>);
}

}  // namespace micgpu
